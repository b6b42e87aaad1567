#!/bin/bash
# Host-side GPU visit: drop-in tests, e2e sweep over thread counts / process splits, sampling profile of the 24-thread harness.
# usage: tools/gpu_host_round.sh <tag>
TAG=${1:-dev}; S=bench_data/c3_ra10_2160p.bin; MT="timeout 120 frontend/_build/hmdec_mt"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_decode.py -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
( for T in 16 24 32; do echo -n "threads $T: "; $MT -b $S --threads $T --repeat 3 | grep -o '"fps": [0-9.]*\|"failures": [0-9]*\|"cpu_user_s": [0-9.]*\|"cpu_sys_s": [0-9.]*' | tr '\n' ' '; echo; done
  echo -n "no prefetch, 24: "; HMDEC_B200_NO_PREFETCH=1 $MT -b $S --threads 24 --repeat 3 | grep -o '"fps": [0-9.]*\|"cpu_user_s": [0-9.]*' | tr '\n' ' '; echo
  echo -n "hm coeff, 24: "; HMDEC_B200_HM_COEFF=1 $MT -b $S --threads 24 --repeat 3 | grep -o '"fps": [0-9.]*\|"cpu_user_s": [0-9.]*' | tr '\n' ' '; echo
  echo -n "flag BS, 24: "; HMDEC_B200_BS_FLAGS=1 $MT -b $S --threads 24 --repeat 3 | grep -o '"fps": [0-9.]*\|"cpu_user_s": [0-9.]*' | tr '\n' ' '; echo
  echo -n "no hash, 24: "; $MT -b $S --threads 24 --repeat 3 --no-hash | grep -o '"fps": [0-9.]*\|"cpu_user_s": [0-9.]*' | tr '\n' ' '; echo
  echo -n "q27, 24: "; $MT -b bench_data/q27_ra10_2160p.bin --threads 24 --repeat 3 | grep -o '"fps": [0-9.]*\|"cpu_user_s": [0-9.]*' | tr '\n' ' '; echo
  bash tools/e2e_split.sh 24 3
) > gpurun_out/${TAG}_e2e.log 2>&1
cat gpurun_out/${TAG}_e2e.log
bash tools/host_profile.sh > gpurun_out/${TAG}_host_profile.log 2>&1; grep -A 32 "hmdec_mt, 24" gpurun_out/${TAG}_host_profile.log | cut -c1-130
