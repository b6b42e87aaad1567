#!/bin/bash
# Round profile on the GPU box: bench (ours + reference), then the ncu launch list of one stage-timing pass and full captures.
# usage: tools/profile_run.sh <tag>
TAG=${1:-r01}
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -c 2500 gpurun_out/${TAG}_bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err; tail -c 600 gpurun_out/${TAG}_bench_reference.json
CMD="python tools/stage_times.py bench_data/c3_ra10_2160p.hmr.gz --reps 1"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain.log; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
# frames 0,1 = the I picture and the first B picture, 8 times: launch #3 (-s 2) of intra/resid/sao is the I picture of the
# second repetition (the heaviest launch), launch #3 of mc_kernel is the luma launch of the B picture (luma, chroma alternate)
for K in mc_kernel intra_kernel resid_kernel sao_kernel; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:^$K -s 2 -c 1 -o gpurun_out/${TAG}_$K -f python tools/replay_frames.py bench_data/c3_ra10_2160p.hmr.gz --frames 0,1 --reps 8 > gpurun_out/${TAG}_ncu_$K.log 2>&1
done
ls -la gpurun_out | tail -12
