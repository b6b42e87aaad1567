#!/bin/bash
# e2e exploration on the GPU box: drop-in decoder threads vs the reference decoder processes
S=${1:-bench_data/c3_ra10_2160p.bin}
for t in 1 4 8 16; do frontend/_build/hmdec_mt -b $S --threads $t --repeat 2 --pin 0; done
for t in 1 16; do frontend/_build/hmdec_mt -b $S --threads $t --repeat 2 --pin 0 --no-hash; done
frontend/_build/hmdec_mt -b $S --threads 16 --repeat 2 --pin 0 --no-hash --no-planes
echo "reference 1 proc:"; /usr/bin/time -f "%e s" oracle/_ref/TAppDecoderStatic -b $S -d 0 > /dev/null
echo "reference 1 proc no hash:"; /usr/bin/time -f "%e s" oracle/_ref/TAppDecoderStatic -b $S -d 0 --SEIDecodedPictureHash=0 > /dev/null
