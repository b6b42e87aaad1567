#!/bin/bash
# e2e throughput of hmdec_mt as P processes x T/P threads (same total), to expose process-wide serialisation.  usage: tools/e2e_split.sh <total threads> <passes>
T=${1:-24}; R=${2:-3}; S=bench_data/c3_ra10_2160p.bin
for P in 1 2 4; do
  rm -f /tmp/e2esplit_*.json
  START=$(python -c "import time; print(f'{time.time()+25:.3f}')")
  for i in $(seq 1 $P); do
    timeout 300 frontend/_build/hmdec_mt -b $S --threads $((T / P)) --repeat $R --start-at $START > /tmp/e2esplit_${P}_${i}.json 2>/dev/null &
  done
  wait
  python - <<'PY'
import json,glob
rs=[json.loads(open(f).read().strip().splitlines()[-1]) for f in glob.glob('/tmp/e2esplit_*.json')]
t0=min(r['t_start'] for r in rs); t1=max(r['t_end'] for r in rs); pics=sum(r['pictures'] for r in rs)
print(f"{len(rs)} process(es): {pics} pictures in {t1-t0:.3f} s = {pics/(t1-t0):.1f} fps; failures {sum(r['failures'] for r in rs)}; cpu {sum(r['cpu_user_s']+r['cpu_sys_s'] for r in rs):.1f} s")
PY
done
