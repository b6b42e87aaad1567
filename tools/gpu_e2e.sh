#!/bin/bash
# e2e sweep on the GPU box (every run bounded).  usage: tools/gpu_e2e.sh <tag>
TAG=${1:-dev}; S=bench_data/c3_ra10_2160p.bin; MT="timeout 120 frontend/_build/hmdec_mt"
( $MT -b $S --threads 1 --repeat 3 --pin 0
  $MT -b $S --threads 16 --repeat 3 --pin 0
  $MT -b $S --threads 16 --repeat 3
  $MT -b $S --threads 16 --repeat 3 --pin 0 --no-hash
  $MT -b $S --threads 24 --repeat 3
  $MT -b $S --threads 32 --repeat 3
  echo "c5 low delay:"; $MT -b bench_data/c5_ld10_2160p_s50.bin --threads 16 --repeat 3 --pin 0
) > gpurun_out/${TAG}_e2e.log 2>&1
cat gpurun_out/${TAG}_e2e.log
