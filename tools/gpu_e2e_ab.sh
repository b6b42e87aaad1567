#!/bin/bash
# A/B of host-side switches on the GPU box, interleaved runs (every run bounded).  usage: tools/gpu_e2e_ab.sh <tag> <ENV=1> [threads]
TAG=${1:-dev}; SW=${2:-HMDEC_B200_NO_PREFETCH=1}; T=${3:-24}; S=bench_data/c3_ra10_2160p.bin; MT="timeout 120 frontend/_build/hmdec_mt"
( for i in 1 2 3; do
    echo -n "base   : "; $MT -b $S --threads $T --repeat 3 | grep -o '"fps": [0-9.]*\|"failures": [0-9]*' | tr '\n' ' '; echo
    echo -n "$SW: "; env $SW $MT -b $S --threads $T --repeat 3 | grep -o '"fps": [0-9.]*\|"failures": [0-9]*' | tr '\n' ' '; echo
  done ) > gpurun_out/${TAG}_ab.log 2>&1
cat gpurun_out/${TAG}_ab.log
