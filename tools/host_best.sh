#!/bin/bash
# Host-side parse + record emission speed without a GPU (null sink): best of N runs of hmdec_cli (the build container is noisy).
# usage: tools/host_best.sh [stream] [runs] [repeat]      extra environment is passed through (e.g. HMDEC_B200_HM_COEFF=1)
S=${1:-bench_data/c3_ra10_2160p.bin}; N=${2:-5}; R=${3:-2}
for i in $(seq 1 $N); do HMDEC_B200_QUIET=1 frontend/_build/hmdec_cli -b $S --dump null --no-hash --repeat $R 2>&1 | tail -1; done | sed 's/.*(\([0-9.]*\) fps).*/\1/' | sort -n | tail -1
