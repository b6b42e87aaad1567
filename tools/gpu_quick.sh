#!/bin/bash
# Quick GPU visit for kernel work: engine parity (C ABI vs HM's MD5s at three stages, all golden streams), per-stage timing of the
# headline workload with an MD5 check of every picture.   usage: tools/gpu_quick.sh <tag> [extra pytest args]
TAG=${1:-dev}; shift
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q "$@" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/${TAG}_pytest.log
[ -f gpurun_out/parity_failures.log ] && head -20 gpurun_out/parity_failures.log
timeout 300 python tools/stage_times.py bench_data/c3_ra10_2160p.hmr.gz --reps 3 --per-frame --check > gpurun_out/${TAG}_stages.log 2>&1; echo "stages rc=$?"; head -3 gpurun_out/${TAG}_stages.log; tail -12 gpurun_out/${TAG}_stages.log; grep -c MISMATCH gpurun_out/${TAG}_stages.log
