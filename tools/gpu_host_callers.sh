#!/bin/bash
# Who calls into libc / libcuda from the decoder threads of the e2e harness (ioctl, poll, memcpy ...): stack-scanning sampler
# (tools/pcsample_stack.c) on hmdec_mt, steady state.   usage: tools/gpu_host_callers.sh <tag> [threads] [repeat]
TAG=${1:-dev}; T=${2:-16}; R=${3:-6}; mkdir -p gpurun_out
gcc -O2 -shared -fPIC -o /tmp/libpcstack.so tools/pcsample_stack.c || exit 1
PCS_OUT=/tmp/pcs_stack_mt.txt LD_PRELOAD=/tmp/libpcstack.so frontend/_build/hmdec_mt -b bench_data/c3_ra10_2160p.bin --threads $T --repeat $R 2>&1 | tail -1 | cut -c1-200 > gpurun_out/${TAG}_host_callers.log
python tools/pcsample_stack_resolve.py /tmp/pcs_stack_mt.txt 60 >> gpurun_out/${TAG}_host_callers.log
cut -c1-230 gpurun_out/${TAG}_host_callers.log
