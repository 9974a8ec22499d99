#!/usr/bin/env python
"""developer helper: per-region execution counts / active lanes from an `ncu --page source --print-source sass --csv` dump"""
import csv, sys, collections
rows=list(csv.reader(open(sys.argv[1])))
# several kernels may be concatenated: take the first
out=[]; hdr=None
for r in rows:
    if r and r[0]=="Address":
        if hdr is not None: break
        hdr=r; continue
    if hdr is not None and len(r)==len(hdr): out.append(r)
ix={h:i for i,h in enumerate(hdr)}
data=out
I=lambda r,k:int(float(r[ix[k]] or 0))
tot=sum(I(r,'Instructions Executed') for r in data)
print('total warp instr',tot,'sass lines',len(data))
b=collections.Counter()
for r in data:
    n=I(r,'Instructions Executed')
    if n==0: continue
    a=float(r[ix['Avg. Threads Executed']])
    b[int(a//4)*4]+=n
for k in sorted(b): print('active lanes %2d-%2d: %10d %.1f%%'%(k,k+3,b[k],100*b[k]/tot))
prev=None; start=0; acc=0; regions=[]
for i,r in enumerate(data):
    n=I(r,'Instructions Executed')
    if prev is None or abs(n-prev)>0.02*max(n,prev,1):
        if prev is not None: regions.append((start,i-1,prev,acc))
        start=i; acc=0
    prev=n; acc+=n
regions.append((start,len(data)-1,prev,acc))
thr=float(sys.argv[2]) if len(sys.argv)>2 else 0.004
for s,e,n,a in regions:
    if a<thr*tot: continue
    ops=collections.Counter(data[k][ix['Source']].split()[0] if not data[k][ix['Source']].strip().startswith('@') else data[k][ix['Source']].split()[1] for k in range(s,e+1))
    top=' '.join('%s:%d'%(o,c) for o,c in ops.most_common(5))
    st=sum(I(data[k],'# Samples') for k in range(s,e+1))
    print('%5d-%5d len %4d execs %8d share %5.1f%% lanes %5s samples %6d | %s'%(s,e,e-s+1,n,100*a/tot,data[s][ix['Avg. Threads Executed']],st,top))
