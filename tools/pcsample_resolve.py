import sys,subprocess,bisect,collections
maps=[];samples=[];tids=[];rets=[]
for l in open(sys.argv[1]):
    if l[0]=='M':
        p=l[2:].split()
        a,b=[int(x,16) for x in p[0].split('-')]; off=int(p[2],16); path=p[5] if len(p)>5 else ''
        maps.append((a,b,off,path))
    else:
        q=l[2:].split(); samples.append(int(q[0],16)); tids.append(int(q[1]) if len(q)>1 else 0); rets.append(int(q[2],16) if len(q)>2 else 0)
syms={}
def load(path):
    if path in syms: return syms[path]
    try:
        out=subprocess.run(['nm','-C','--defined-only','-n',path],capture_output=True,text=True).stdout
        if not out.strip(): out=subprocess.run(['nm','-C','-D','--defined-only','-n',path],capture_output=True,text=True).stdout
    except Exception: out=''
    arr=[]
    for l in out.splitlines():
        p=l.split(' ',2)
        if len(p)==3 and p[1] in 'TtWw':
            arr.append((int(p[0],16),p[2]))
    arr.sort(); syms[path]=(arr,[a for a,_ in arr]); return syms[path]
cnt=collections.Counter(); mod=collections.Counter(); raw=collections.Counter()
bytid=collections.Counter(tids); polltid=collections.Counter()
for pc,tid in zip(samples,tids):
    for a,b,off,path in maps:
        if a<=pc<b:
            arr,keys=load(path) if path.startswith('/') else ([],[])
            va=pc-a+off
            # for PIE/shared: symbol addresses are file vaddrs; assume vaddr==file offset mapping for text
            i=bisect.bisect_right(keys,va)-1
            name=arr[i][1] if i>=0 else '?'
            cnt[(path.split('/')[-1],name)]+=1; mod[path.split('/')[-1]]+=1
            if 'libc.so' in path: raw[va & ~0x3f]+=1
            if name.startswith('poll') or name.startswith('ioctl'): polltid[tid]+=1
            break
    else: cnt[('?','?')]+=1
tot=len(samples); print('samples',tot)
for m,c in mod.most_common(8): print(f'  module {m}: {100*c/tot:.1f}%')
for (m,nm),c in cnt.most_common(int(sys.argv[2]) if len(sys.argv)>2 else 45): print(f'{100*c/tot:5.1f}% {c:6d} {m[:22]:22s} {nm[:100]}')

print('libc hot 64-byte blocks (file vaddr):', ' '.join(f'{a:#x}:{c}' for a,c in raw.most_common(12)))
print('samples per thread:', dict(bytid.most_common(8)), ' poll/ioctl samples per thread:', dict(polltid.most_common(8)))

# callers of the libc samples: the word at the stack pointer is the return address while a leaf routine (memset / memcpy) runs
def sym_of(pc):
    for a,b,off,path in maps:
        if a<=pc<b:
            arr,keys=load(path) if path.startswith('/') else ([],[])
            i=bisect.bisect_right(keys,pc-a+off)-1
            return path.split('/')[-1], (arr[i][1] if i>=0 else '?')
    return '?','?'
callers=collections.Counter(); nlibc=0
for pc,ret in zip(samples,rets):
    m,_=sym_of(pc)
    if 'libc.so' in m and ret:
        nlibc+=1
        cm,cn=sym_of(ret)
        callers[(cm,cn) if cm not in ('?',) and 'libc.so' not in cm else ('(inside libc: malloc / free internals)','')]+=1
if nlibc:
    print('callers of the libc samples (leaf routines: memset / memcpy):')
    for (m,nm),c in callers.most_common(14): print(f'{100*c/tot:5.1f}% {c:6d} {m[:22]:22s} {nm[:100]}')
