#!/bin/bash
# e2e sweep on the GPU box (every run bounded).  usage: tools/gpu_e2e.sh <tag>
TAG=${1:-dev}; S=bench_data/c3_ra10_2160p.bin; MT="timeout 150 frontend/_build/hmdec_mt"
( $MT -b $S --threads 1 --repeat 2 --pin 0
  $MT -b $S --threads 1 --repeat 2 --pin 0 --no-hash
  $MT -b $S --threads 16 --repeat 2 --pin 0
  $MT -b $S --threads 16 --repeat 2 --pin 0 --no-hash
  timeout 300 python - <<PY
import sys, os; sys.path.insert(0,'tools')
import e2e_procs, json
for p,t,extra,env in ((4,4,(),{}),(8,2,(),{}),(16,1,(),{}),(4,4,("--no-hash",),{}),(4,4,(),{"HMDEC_B200_HOST_MD5":"1"}),(4,4,(),{"HMDEC_B200_LAZY_PLANES":"1"})):
    os.environ.pop("HMDEC_B200_HOST_MD5",None); os.environ.pop("HMDEC_B200_LAZY_PLANES",None); os.environ.update(env)
    r=e2e_procs.run("$S", p, t, 3, extra, lead=10.0); r["extra"]=list(extra); r["env"]=env
    print(json.dumps(r), flush=True)
PY
) > gpurun_out/${TAG}_e2e.log 2>&1
cat gpurun_out/${TAG}_e2e.log
