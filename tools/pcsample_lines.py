"""Line-level view of a pcsample profile for one function: tools/pcsample_lines.py <samples> <shared object> <function substring> [top]
(the object must have been compiled with -g; addr2line maps the sampled PCs to source lines)."""
import sys, subprocess, collections
maps = []; samples = []
for l in open(sys.argv[1]):
    if l[0] == 'M':
        p = l[2:].split(); a, b = [int(x, 16) for x in p[0].split('-')]
        maps.append((a, b, int(p[2], 16), p[5] if len(p) > 5 else ''))
    else: samples.append(int(l[2:].split()[0], 16))
so = sys.argv[2]; want = sys.argv[3]; top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
vas = collections.Counter()
for pc in samples:
    for a, b, off, path in maps:
        if a <= pc < b and path.endswith(so.split('/')[-1]): vas[pc - a + off] += 1
addrs = sorted(vas)
out = subprocess.run(['addr2line', '-f', '-C', '-i', '-e', so] + [hex(a) for a in addrs], capture_output=True, text=True).stdout.splitlines()
# -i prints chains of (function, file:line); split per address by re-running without -i for the count of lines is awkward: use one call per address batch without -i
out = subprocess.run(['addr2line', '-f', '-C', '-e', so] + [hex(a) for a in addrs], capture_output=True, text=True).stdout.splitlines()
lines = collections.Counter(); total = 0
for i, a in enumerate(addrs):
    fn, loc = out[2 * i], out[2 * i + 1]
    total += vas[a]
    lines[(fn[:60], loc.split('/')[-1])] += vas[a]
sel = [(k, c) for k, c in lines.items() if want in k[0] or want in k[1]]
n = sum(c for _, c in sel)
print(f'{n} of {len(samples)} samples match "{want}"')
for (fn, loc), c in sorted(sel, key=lambda x: -x[1])[:top]: print(f'{c:5d} {100.0 * c / max(n, 1):5.1f}%  {loc:40s} {fn}')
