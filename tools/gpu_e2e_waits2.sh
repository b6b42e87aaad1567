#!/bin/bash
# Wall vs on-CPU time of the decoder threads' phases (HMDEC_B200_STATS), single stream and saturated.  usage: tools/gpu_e2e_waits2.sh <tag>
TAG=${1:-dev}; S=bench_data/c3_ra10_2160p.bin; MT="timeout 120 frontend/_build/hmdec_mt"
mkdir -p gpurun_out
(
echo "== 1 thread"; HMDEC_B200_STATS=1 $MT -b $S --threads 1 --repeat 3 2>&1 | grep -v "hm_emit stats" | cut -c1-330
echo "== 1 thread, no hash"; HMDEC_B200_STATS=1 $MT -b $S --threads 1 --repeat 3 --no-hash 2>&1 | grep -v "hm_emit stats" | cut -c1-330
echo "== 1 thread, lazy, no planes, no hash"; HMDEC_B200_STATS=1 HMDEC_B200_LAZY_PLANES=1 $MT -b $S --threads 1 --repeat 3 --no-hash --no-planes 2>&1 | grep -v "hm_emit stats" | cut -c1-330
echo "== 32 threads"; HMDEC_B200_STATS=32 $MT -b $S --threads 32 --repeat 3 2>&1 | grep -v "hm_emit stats" | cut -c1-330
echo "== 16 threads"; HMDEC_B200_STATS=16 $MT -b $S --threads 16 --repeat 3 2>&1 | grep -v "hm_emit stats" | cut -c1-330
) > gpurun_out/${TAG}_e2e_waits2.log 2>&1
cat gpurun_out/${TAG}_e2e_waits2.log
