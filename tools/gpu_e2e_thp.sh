#!/bin/bash
# e2e harness A/B on one box: transparent huge pages for malloc (GLIBC_TUNABLES=glibc.malloc.hugetlb=1) off / on, 48 threads, and
# a loop of short runs that records every exit status (an exit-time crash would lose the result line).  usage: tools/gpu_e2e_thp.sh <tag>
TAG=${1:-dev}; mkdir -p gpurun_out; L=gpurun_out/${TAG}_e2e_thp.log; : > $L
echo "thp: $(cat /sys/kernel/mm/transparent_hugepage/enabled) defrag: $(cat /sys/kernel/mm/transparent_hugepage/defrag) cores: $(nproc)" >> $L
run() { frontend/_build/hmdec_mt -b bench_data/c3_ra10_2160p.bin --threads $1 --repeat $2 > /tmp/mt.out 2>/tmp/mt.err; echo "rc=$? $(tail -1 /tmp/mt.out | cut -c1-120) $(grep -o '"cpu_user_s.*' /tmp/mt.out | cut -c1-60)"; }
for i in 1 2 3; do
  echo "plain   $(run 32 4)" >> $L
  echo "hugetlb $(GLIBC_TUNABLES=glibc.malloc.hugetlb=1 run 32 4)" >> $L
done
echo "48 thr  $(run 48 3)" >> $L
for i in 1 2 3 4 5 6 7 8; do echo "short   $(run 32 1)" >> $L; done
cat $L
