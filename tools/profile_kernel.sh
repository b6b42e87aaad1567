#!/bin/bash
# ncu full capture of one kernel on selected pictures.  usage: tools/profile_kernel.sh <tag> <kernel-regex> <frames> [skip] [dump]
TAG=$1; K=$2; FR=$3; SKIP=${4:-2}; DUMP=${5:-bench_data/c3_ra10_2160p.hmr.gz}
CMD="python tools/replay_frames.py $DUMP --frames $FR --reps 2"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain.log; exit 1; }
cat gpurun_out/${TAG}_plain.log
ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c 1 -o gpurun_out/${TAG}_$K -f $CMD > gpurun_out/${TAG}_ncu.log 2>&1
tail -3 gpurun_out/${TAG}_ncu.log
