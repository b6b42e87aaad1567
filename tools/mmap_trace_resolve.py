# usage: python tools/mmap_trace_resolve.py /tmp/mmap_trace.txt [rows]
import sys,subprocess,bisect
maps=[];rows=[]
for l in open(sys.argv[1]):
    if l[0]=='M':
        p=l[2:].split(); a,b=[int(x,16) for x in p[0].split('-')]; maps.append((a,b,int(p[2],16),p[5] if len(p)>5 else ''))
    else:
        q=l.split(); rows.append((int(q[1]),int(q[2]),int(q[3]),[int(x,16) if x!='(nil)' else 0 for x in q[4:8]]))
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
            return (arr[i][1] if i>=0 else '?')[:48]
    return '?'
kinds=['mmap','munmap','madvise']
for kind,n,b,pcs in sorted(rows,key=lambda r:-r[2])[:int(sys.argv[2]) if len(sys.argv)>2 else 30]:
    print(f'{kinds[kind]:8s} {n:8d} calls {b/1e6:10.1f} MB  '+' <- '.join(sym_of(p) for p in pcs))
