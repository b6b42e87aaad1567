#!/bin/bash
# Round profile on the GPU box: plain run first (must exit 0), then the ncu launch list and one full capture of the top kernel.
# usage: tools/profile_run.sh <tag> <kernel-regex>
TAG=${1:-r01}; KREGEX=${2:-intra_kernel}
CMD="python tools/stage_times.py bench_data/c3_ra10_2160p.hmr.gz --reps 1"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:${KREGEX} -s 40 -c 3 -o gpurun_out/${TAG}_${KREGEX} -f $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
tail -3 gpurun_out/${TAG}_ncu1.log gpurun_out/${TAG}_ncu2.log
