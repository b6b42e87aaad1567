# usage: python tools/ioctl_trace_resolve.py /tmp/ioctl_trace.txt [rows]
import sys,subprocess,bisect
maps=[];rows=[]
for l in open(sys.argv[1]):
    if l[0]=='M':
        p=l[2:].split(); a,b=[int(x,16) for x in p[0].split('-')]; maps.append((a,b,int(p[2],16),p[5] if len(p)>5 else ''))
    else:
        q=l.split(); rows.append((int(q[1],16),int(q[2]),int(q[3]),int(q[4],16),int(q[5],16)))
syms={}
def load(path):
    if path not in syms:
        out=subprocess.run(['nm','-C','--defined-only','-n',path],capture_output=True,text=True).stdout
        if not out.strip(): out=subprocess.run(['nm','-C','-D','--defined-only','-n',path],capture_output=True,text=True).stdout
        arr=sorted((int(p[0],16),p[2]) for p in (l.split(' ',2) for l in out.splitlines()) if len(p)==3 and p[1] in 'TtWw')
        syms[path]=(arr,[a for a,_ in arr])
    return syms[path]
def sym_of(pc):
    if not pc: return '-'
    for a,b,off,path in maps:
        if a<=pc<b and path.startswith('/'):
            arr,keys=load(path); i=bisect.bisect_right(keys,pc-a+off)-1
            return (arr[i][1] if i>=0 else '?')[:60]
    return '?'
tot=sum(r[2] for r in rows); n=sum(r[1] for r in rows)
print(f'ioctl calls {n}, {tot/1e9:.3f} Gticks')
agg={}
for req,c,t,a,b in rows:
    k=(req&0xffff,sym_of(a),sym_of(b)); x=agg.setdefault(k,[0,0]); x[0]+=c; x[1]+=t
for (req,a,b),(c,t) in sorted(agg.items(),key=lambda kv:-kv[1][1])[:int(sys.argv[2]) if len(sys.argv)>2 else 30]:
    print(f'{100*t/max(tot,1):5.1f}% {t/1e6:9.1f} Mticks {c:8d} calls {t/max(c,1):9.0f} ticks/call  req {req:#06x}  {a} <- {b}')
