# Resolver for pcsample_stack.c: for every sample whose PC is inside libc or libcuda, the first two stack words that point into
# one of OUR text mappings (not libc / libcuda / libstdc++) name the callers.  usage: python tools/pcsample_stack_resolve.py /tmp/pcsample_stack.txt [rows]
import sys,subprocess,bisect,collections
maps=[];S=[]
for l in open(sys.argv[1]):
    if l[0]=='M':
        p=l[2:].split(); a,b=[int(x,16) for x in p[0].split('-')]; maps.append((a,b,int(p[2],16),p[5] if len(p)>5 else ''))
    else:
        q=[int(x,16) for x in l[2:].split()]; S.append((q[0],q[1:]))
syms={}
def load(path):
    if path not in syms:
        out=subprocess.run(['nm','-C','--defined-only','-n',path],capture_output=True,text=True).stdout
        if not out.strip(): out=subprocess.run(['nm','-C','-D','--defined-only','-n',path],capture_output=True,text=True).stdout
        arr=sorted((int(p[0],16),p[2]) for p in (l.split(' ',2) for l in out.splitlines()) if len(p)==3 and p[1] in 'TtWw')
        syms[path]=(arr,[a for a,_ in arr])
    return syms[path]
def sym_of(pc):
    for a,b,off,path in maps:
        if a<=pc<b and path.startswith('/'):
            arr,keys=load(path); i=bisect.bisect_right(keys,pc-a+off)-1
            return path.split('/')[-1], (arr[i][1] if i>=0 else '?')
    return None,None
tot=len(S); c1=collections.Counter(); nl=0
for pc,st in S:
    m,n=sym_of(pc)
    if not m or ('libc.so' not in m and 'libcuda' not in m): continue
    nl+=1; chain=[]
    for w in st:
        cm,cn=sym_of(w)
        if cm and 'libc.so' not in cm and 'pcs' not in cm and 'libcuda' not in cm and 'libpthread' not in cm and 'libstdc++' not in cm:
            if not chain or chain[-1]!=cn[:70]: chain.append(cn[:70])
            if len(chain)==2: break
    c1[(n[:28],' <- '.join(chain))]+=1
print('samples',tot,'in libc',nl)
for (n,ch),c in c1.most_common(int(sys.argv[2]) if len(sys.argv)>2 else 30): print(f'{100*c/tot:5.1f}% {c:5d} {n:28s} {ch}')
