#!/bin/bash
# Kernel-work GPU visit: engine parity (all golden streams, three stages), per-stage timing of the headline workload with an MD5
# check of every picture, then the resident bench (value only) for each HMR_INTRA_CTAS given.   usage: tools/gpu_quick2.sh <tag> [ctas ...]
TAG=${1:-dev}; shift
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
[ -f gpurun_out/parity_failures.log ] && head -20 gpurun_out/parity_failures.log
for N in "" "$@"; do
  [ -n "$N" ] && export HMR_INTRA_CTAS=$N
  echo "== HMR_INTRA_CTAS=${N:-default}"
  timeout 300 python tools/stage_times.py bench_data/c3_ra10_2160p.hmr.gz --reps 3 --per-frame --check > gpurun_out/${TAG}_stages${N}.log 2>&1; echo "stages rc=$?"; head -2 gpurun_out/${TAG}_stages${N}.log | cut -c1-200; tail -9 gpurun_out/${TAG}_stages${N}.log; grep -c MISMATCH gpurun_out/${TAG}_stages${N}.log
  timeout 300 python bench.py --no-e2e --no-extra --no-cpu-baseline > gpurun_out/${TAG}_bench${N}.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/${TAG}_bench${N}.json').read().strip().splitlines()[-1]); print('value', d['value'], 'ms/step', d['ms_per_step'], 'single-stream us/picture', d['roofline']['single_stream_us_per_picture'])"
done
