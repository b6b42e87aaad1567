#!/bin/bash
# Sampling profile (SIGPROF on process CPU time, 1 kHz) of the host side of the real drop-in on the GPU box: where the
# decoder threads' CPU time goes.  Part 1: one decoder thread; part 2: the e2e harness (hmdec_mt, 24 threads, steady state).
gcc -O2 -shared -fPIC -o /tmp/libpcsample.so tools/pcsample.c || exit 1
echo "== hmdec_cli, 1 thread, 6 passes (first pass cold)"
HMDEC_B200_QUIET=1 PCS_OUT=/tmp/pcsample.txt LD_PRELOAD=/tmp/libpcsample.so frontend/_build/hmdec_cli -b bench_data/c3_ra10_2160p.bin --touch-planes --repeat 6 2>&1 | tail -1
python tools/pcsample_resolve.py /tmp/pcsample.txt 40
echo "== hmdec_mt, 24 threads x 3 passes + warm-up pass"
PCS_OUT=/tmp/pcsample_mt.txt LD_PRELOAD=/tmp/libpcsample.so frontend/_build/hmdec_mt -b bench_data/c3_ra10_2160p.bin --threads 24 --repeat 3 2>&1 | tail -1
python tools/pcsample_resolve.py /tmp/pcsample_mt.txt 40
