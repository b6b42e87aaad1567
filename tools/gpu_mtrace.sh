#!/bin/bash
# Which large mallocs are slow in the e2e harness (steady state, 16 threads)?   usage: tools/gpu_mtrace.sh <tag>
TAG=${1:-dev}; mkdir -p gpurun_out
gcc -O2 -shared -fPIC -o /tmp/libmtrace.so tools/mtrace.c -ldl || exit 1
MTRACE_OUT=/tmp/mtrace_mt.txt LD_PRELOAD=/tmp/libmtrace.so frontend/_build/hmdec_mt -b bench_data/c3_ra10_2160p.bin --threads 16 --repeat 6 2>&1 | tail -1 | cut -c1-160 > gpurun_out/${TAG}_mtrace.log
python tools/mtrace_resolve.py /tmp/mtrace_mt.txt 24 >> gpurun_out/${TAG}_mtrace.log
cut -c1-260 gpurun_out/${TAG}_mtrace.log
