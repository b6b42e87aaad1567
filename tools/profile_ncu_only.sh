#!/bin/bash
# Only the ncu --set full captures of tools/profile_run.sh (after a plain run of the same command).  usage: tools/profile_ncu_only.sh <tag>
TAG=${1:-r01}
python tools/replay_frames.py bench_data/c3_ra10_2160p.hmr.gz --frames 0,1 --reps 8 > gpurun_out/${TAG}_plain2.log 2>&1 || { echo "plain run failed"; exit 1; }
for K in mc_kernel intra_kernel resid_kernel sao_kernel; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:^$K -s 2 -c 1 -o gpurun_out/${TAG}_$K -f python tools/replay_frames.py bench_data/c3_ra10_2160p.hmr.gz --frames 0,1 --reps 8 > gpurun_out/${TAG}_ncu_$K.log 2>&1
done
ls -la gpurun_out | grep ${TAG}_ | tail -8
