#!/bin/bash
# One GPU visit: parity tests, per-stage timing of the headline workload, optional bench.
# usage: tools/gpu_round.sh <tag> [bench]
TAG=${1:-dev}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/${TAG}_pytest.log
[ -f gpurun_out/parity_failures.log ] && head -20 gpurun_out/parity_failures.log
python tools/stage_times.py bench_data/c3_ra10_2160p.hmr.gz --reps 3 --per-frame --check > gpurun_out/${TAG}_stages.log 2>&1; tail -12 gpurun_out/${TAG}_stages.log; grep -c MISMATCH gpurun_out/${TAG}_stages.log
if [ "$2" = "bench" ]; then python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 3000 gpurun_out/${TAG}_bench.json; fi
