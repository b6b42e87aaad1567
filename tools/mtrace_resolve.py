"""Large-malloc call sites with time per site: tools/mtrace.c (LD_PRELOAD, MTRACE_OUT) -> table.  usage: mtrace_resolve.py out.txt lib.so [lib2.so ...]"""
import subprocess, bisect, sys
maps = []; C = []
for l in open(sys.argv[1]):
    if l[0] == 'M':
        p = l[2:].split(); a, b = [int(x, 16) for x in p[0].split('-')]; maps.append((a, b, int(p[2], 16), p[5] if len(p) > 5 else ''))
    else:
        p = l.split(); C.append((int(p[6]), int(p[1]), int(p[2]), [int(x, 16) if x != '(nil)' else 0 for x in p[3:6]]))
tables = {}
def table(path):
    if path not in tables:
        syms = []
        for l in subprocess.run(['nm', '-C', '--defined-only', path], capture_output=True, text=True).stdout.splitlines():
            p = l.split(' ', 2)
            if len(p) == 3 and p[1] in 'tTwW': syms.append((int(p[0], 16), p[2]))
        syms.sort(); tables[path] = (syms, [a for a, _ in syms])
    return tables[path]
def name(pc):
    for a, b, off, path in maps:
        if a <= pc < b:
            if path.startswith('/') and ('/repo/' in path or 'hmdec' in path):
                syms, addrs = table(path); i = bisect.bisect_right(addrs, pc - a + off) - 1
                return syms[i][1][:60] if i >= 0 else '?'
            return '[' + path.split('/')[-1] + ']'
    return '?'
C.sort(reverse=True)
print("total ticks in mallocs >= 1 KB: %.3f G in %d calls" % (sum(c[0] for c in C) / 1e9, sum(c[1] for c in C)))
for t, n, b, pcs in C[:int(sys.argv[2]) if len(sys.argv) > 2 else 16]:
    print(f'{t / 1e6:9.1f} Mticks {n:7d} calls avg {b // n:8d} B {t // n:8d} ticks/call  ' + ' <- '.join(name(p) for p in pcs))
