import subprocess, bisect, sys
maps = []; sites = []
for l in open(sys.argv[1]):
    p = l.split()
    if p[0] == 'M':
        a, b = [int(x, 16) for x in p[1].split('-')]; maps.append((a, b, int(p[3], 16), p[6] if len(p) > 6 else ''))
    else: sites.append((int(p[1], 16), int(p[2]), int(p[3]), int(p[4]), int(p[5])))
syms = {}
def load(path):
    if path not in syms:
        out = subprocess.run(['nm', '-C', '--defined-only', '-n', path], capture_output=True, text=True).stdout or subprocess.run(['nm', '-C', '-D', '--defined-only', '-n', path], capture_output=True, text=True).stdout
        arr = [(int(x.split(' ', 2)[0], 16), x.split(' ', 2)[2]) for x in out.splitlines() if len(x.split(' ', 2)) == 3 and x.split(' ', 2)[1] in 'TtWw']
        arr.sort(); syms[path] = (arr, [a for a, _ in arr])
    return syms[path]
tot = sum(s[3] for s in sites)
print('large mallocs', sum(s[1] for s in sites), 'total Mcycles', tot / 1e6)
for pc, n, b, cyc, worst in sorted(sites, key=lambda s: -s[3])[:int(sys.argv[2]) if len(sys.argv) > 2 else 20]:
    name = '?'
    for a, e, off, path in maps:
        if a <= pc < e and path.startswith('/'):
            arr, keys = load(path); i = bisect.bisect_right(keys, pc - a + off) - 1; name = path.split('/')[-1] + ' ' + (arr[i][1] if i >= 0 else '?'); break
    print(f'{cyc/1e6:10.1f} Mcyc {100*cyc/tot:5.1f}%  n={n:8d} avg={cyc/max(n,1):9.0f} cyc worst={worst:10d} {b/1e6:9.1f} MB  {name[:110]}')
