#!/bin/bash
# Which calls of ours end in ioctl() on the GPU box, how often and for how long (tools/ioctl_trace.c).  usage: tools/gpu_ioctl_trace.sh <tag> [threads] [repeat]
TAG=${1:-dev}; T=${2:-16}; R=${3:-6}; mkdir -p gpurun_out
gcc -O2 -shared -fPIC -o /tmp/libioctltrace.so tools/ioctl_trace.c -ldl -lpthread || exit 1
IOT_OUT=/tmp/iot_mt.txt LD_PRELOAD=/tmp/libioctltrace.so frontend/_build/hmdec_mt -b bench_data/c3_ra10_2160p.bin --threads $T --repeat $R 2>&1 | tail -1 | cut -c1-200 > gpurun_out/${TAG}_ioctl_trace.log
python tools/ioctl_trace_resolve.py /tmp/iot_mt.txt 40 >> gpurun_out/${TAG}_ioctl_trace.log
cut -c1-230 gpurun_out/${TAG}_ioctl_trace.log
