#!/bin/bash
# compute-sanitizer over the engine parity tests of the small golden streams: memcheck (out-of-bounds / misaligned global and shared
# accesses) on every stream of test_engine_matches_hm_and_oracle plus the hash kernels, racecheck (shared-memory hazards) on two
# streams.  usage: tools/gpu_sanitizer.sh <tag>
TAG=${1:-dev}; mkdir -p gpurun_out
SAN=/usr/local/cuda/bin/compute-sanitizer
timeout 420 $SAN --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "test_engine_matches_hm_and_oracle or test_picture_hashes_match_hm or test_device_md5_matches_hm_golden" > gpurun_out/${TAG}_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|misaligned" gpurun_out/${TAG}_memcheck.log | head -12
timeout 240 $SAN --tool racecheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "test_engine_matches_hm_and_oracle and (c1_intra8_240p or s_ra8_odd)" > gpurun_out/${TAG}_racecheck.log 2>&1
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/${TAG}_racecheck.log | head -12
