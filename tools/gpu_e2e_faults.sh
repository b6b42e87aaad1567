#!/bin/bash
# e2e harness: fps, CPU seconds and page faults of the timed region, plain and with huge pages for malloc.  usage: tools/gpu_e2e_faults.sh <tag> [threads] [repeat]
TAG=${1:-dev}; T=${2:-32}; R=${3:-3}; mkdir -p gpurun_out; L=gpurun_out/${TAG}_e2e_faults.log; : > $L
run() { frontend/_build/hmdec_mt -b bench_data/c3_ra10_2160p.bin --threads $T --repeat $R > /tmp/mt.out 2>/tmp/mt.err; echo "rc=$? $(tail -1 /tmp/mt.out | grep -o '"fps": [0-9.]*\|"cpu_user_s": [0-9.]*\|"cpu_sys_s": [0-9.]*\|"minor_faults": [0-9]*\|"failures": [0-9]*' | tr '\n' ' ')"; }
echo "plain   $(run)" >> $L
echo "hugetlb $(GLIBC_TUNABLES=glibc.malloc.hugetlb=1 run)" >> $L
echo "plain   $(run)" >> $L
echo "no record pool $(HMDEC_B200_NO_RECORD_POOL=1 run)" >> $L
cat $L
