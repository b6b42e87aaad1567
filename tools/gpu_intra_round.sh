#!/bin/bash
# Intra-wavefront GPU visit: cycle accounting of the chain warps (INTRA_PROFILE build, libhm_b200/libhmrecon_prof.so built in the
# container), the saturation probe and the resident bench at other stream counts.   usage: tools/gpu_intra_round.sh <tag> [pytest -k expr]
TAG=${1:-dev}; K=${2:-}
mkdir -p gpurun_out
if [ -n "$K" ]; then timeout 600 python -m pytest tests -m gpu -x -q -k "$K" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_pytest.log; fi
if [ -f libhm_b200/libhmrecon_prof.so ]; then
  cp libhm_b200/libhmrecon.so /tmp/keep.so; cp libhm_b200/libhmrecon_prof.so libhm_b200/libhmrecon.so
  for F in 0 1 2; do timeout 120 python tools/intra_profile.py bench_data/c3_ra10_2160p.hmr.gz --frames $F > gpurun_out/${TAG}_intra_prof_f$F.log 2>&1; done
  cp /tmp/keep.so libhm_b200/libhmrecon.so
  cat gpurun_out/${TAG}_intra_prof_f0.log | head -60
fi
timeout 300 python tools/saturation_probe.py bench_data/c3_ra10_2160p.hmr.gz > gpurun_out/${TAG}_sat8.log 2>&1; cat gpurun_out/${TAG}_sat8.log
for S in 12 16; do timeout 300 python bench.py --streams $S --no-e2e --no-extra --no-cpu-baseline > gpurun_out/${TAG}_bench_s$S.json 2>/dev/null; python -c "
import json,sys; d=json.loads(open('gpurun_out/${TAG}_bench_s$S.json').read().strip().splitlines()[-1]); print('streams $S value', d['value'], 'ms/step', d['ms_per_step'])"; done
