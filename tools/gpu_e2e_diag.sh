#!/bin/bash
# What bounds the end-to-end rate: the decoder threads' CPU, or the D2H copy of every output picture (24.9 MB at 2160p Main10)?
# usage: tools/gpu_e2e_diag.sh <tag>
TAG=${1:-dev}; S=bench_data/c3_ra10_2160p.bin; MT="timeout 120 frontend/_build/hmdec_mt"
mkdir -p gpurun_out
F='"fps": [0-9.]*\|"failures": [0-9]*\|"cpu_user_s": [0-9.]*\|"cpu_sys_s": [0-9.]*\|"seconds": [0-9.]*'
(
nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max,pcie.link.width.max --format=csv
nproc; lscpu | grep -E "Model name|Thread|Core|Socket|MHz" | head -8
python - <<'PY'
import torch, time
d = torch.empty(3840*2160*3//2, dtype=torch.int16, device='cuda')
h = torch.empty_like(d, device='cpu').pin_memory()
for n in (1, 4):
    ss = [torch.cuda.Stream() for _ in range(n)]
    hs = [torch.empty_like(d, device='cpu').pin_memory() for _ in range(n)]
    torch.cuda.synchronize(); t = time.time()
    for it in range(40):
        for s, hh in zip(ss, hs):
            with torch.cuda.stream(s): hh.copy_(d, non_blocking=True)
    torch.cuda.synchronize(); dt = time.time() - t
    print(f"D2H contiguous 24.9 MB x {40*n} on {n} stream(s): {24.9e-3*40*n/dt:.1f} GB/s")
torch.cuda.synchronize(); t = time.time()
for it in range(40): d.copy_(h, non_blocking=True)
torch.cuda.synchronize(); print(f"H2D contiguous: {24.9e-3*40/(time.time()-t):.1f} GB/s")
PY
for T in 16 24; do echo -n "default, $T threads: "; $MT -b $S --threads $T --repeat 3 | grep -o "$F" | tr '\n' ' '; echo; done
echo -n "no planes requested, lazy (no D2H), hash on, 24: "; HMDEC_B200_LAZY_PLANES=1 $MT -b $S --threads 24 --repeat 3 --no-planes | grep -o "$F" | tr '\n' ' '; echo
echo -n "no planes, lazy, no hash, 24: "; HMDEC_B200_LAZY_PLANES=1 $MT -b $S --threads 24 --repeat 3 --no-planes --no-hash | grep -o "$F" | tr '\n' ' '; echo
echo -n "no planes, lazy, no hash, 16: "; HMDEC_B200_LAZY_PLANES=1 $MT -b $S --threads 16 --repeat 3 --no-planes --no-hash | grep -o "$F" | tr '\n' ' '; echo
echo -n "no planes, lazy, no hash, 32: "; HMDEC_B200_LAZY_PLANES=1 $MT -b $S --threads 32 --repeat 3 --no-planes --no-hash | grep -o "$F" | tr '\n' ' '; echo
echo -n "eager D2H but planes never waited for (--no-planes), 24: "; $MT -b $S --threads 24 --repeat 3 --no-planes | grep -o "$F" | tr '\n' ' '; echo
echo -n "null sink (parse + emit only, no GPU), 16 procs x hmdec_cli: "; t0=$(date +%s.%N); for i in $(seq 1 16); do HMDEC_B200_QUIET=1 frontend/_build/hmdec_cli -b $S --dump null --no-hash --repeat 3 > /dev/null 2>&1 & done; wait; t1=$(date +%s.%N); python -c "print(16*99/($t1-$t0), 'fps')"
) > gpurun_out/${TAG}_e2e_diag.log 2>&1
cat gpurun_out/${TAG}_e2e_diag.log
