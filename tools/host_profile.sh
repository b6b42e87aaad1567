#!/bin/bash
# Sampling profile (SIGPROF, 250 Hz) of the host side of the real drop-in on the GPU box: where a decoder thread's CPU time goes.
gcc -O2 -shared -fPIC -o /tmp/libpcsample.so tools/pcsample.c || exit 1
HMDEC_B200_QUIET=1 PCS_OUT=/tmp/pcsample.txt LD_PRELOAD=/tmp/libpcsample.so frontend/_build/hmdec_cli -b bench_data/c3_ra10_2160p.bin --touch-planes --repeat 4 2>&1 | tail -1
python tools/pcsample_resolve.py /tmp/pcsample.txt 45
