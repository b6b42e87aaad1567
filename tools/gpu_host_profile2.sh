#!/bin/bash
# Steady-state CPU profile of the decoder threads on the GPU box (warm-up pass = 1 of 9), 16 threads on 16 cores, with the callers of
# the libc samples.   usage: tools/gpu_host_profile2.sh <tag>
TAG=${1:-dev}
mkdir -p gpurun_out
gcc -O2 -shared -fPIC -o /tmp/libpcsample.so tools/pcsample.c || exit 1
PCS_OUT=/tmp/pcs_mt.txt LD_PRELOAD=/tmp/libpcsample.so frontend/_build/hmdec_mt -b bench_data/c3_ra10_2160p.bin --threads 16 --repeat 8 2>&1 | tail -1 | cut -c1-200 > gpurun_out/${TAG}_host_profile2.log
python tools/pcsample_resolve.py /tmp/pcs_mt.txt 45 >> gpurun_out/${TAG}_host_profile2.log
cut -c1-150 gpurun_out/${TAG}_host_profile2.log
