#!/bin/bash
# Where the decoder threads WAIT (HMDEC_B200_STATS=<n>: wall-clock accounting of the blocking points, the first n decoders = the
# harness's warm-up pass not counted), for several thread counts and with the verdict of a bitstream collected behind the next one.
# usage: tools/gpu_e2e_waits.sh <tag> [threads ...]
TAG=${1:-dev}; shift; S=bench_data/c3_ra10_2160p.bin; MT="timeout 120 frontend/_build/hmdec_mt"
mkdir -p gpurun_out
(
for T in ${@:-16 24 32}; do echo "== default, $T threads"; HMDEC_B200_STATS=$T $MT -b $S --threads $T --repeat 3 2>&1 | grep -v "hm_emit stats" | cut -c1-330; done
for T in 16 24; do echo "== overlap-verdict, $T threads"; HMDEC_B200_STATS=$T $MT -b $S --threads $T --repeat 3 --overlap-verdict 2>&1 | grep -v "hm_emit stats" | cut -c1-330; done
echo "== no hash, 24 threads"; HMDEC_B200_STATS=24 $MT -b $S --threads 24 --repeat 3 --no-hash 2>&1 | grep -v "hm_emit stats" | cut -c1-330
echo "== lazy planes, no planes, no hash, 16 threads"; HMDEC_B200_STATS=16 HMDEC_B200_LAZY_PLANES=1 $MT -b $S --threads 16 --repeat 3 --no-planes --no-hash 2>&1 | grep -v "hm_emit stats" | cut -c1-330
) > gpurun_out/${TAG}_e2e_waits.log 2>&1
cat gpurun_out/${TAG}_e2e_waits.log
