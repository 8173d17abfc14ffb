#!/usr/bin/env python
"""Stall reasons per function of the fused kernel (same inputs as ncu_regions.py)."""
import re,csv,collections,sys
csv.field_size_limit(10**9)
dis,sasscsv,srcf=sys.argv[1],sys.argv[2],sys.argv[3]
lines=open(dis).read().split('\n')
start=[i for i,l in enumerate(lines) if l.startswith('.text._ZN7b200aug20fused_augment_kernelILb0E')][0]
cur=None; ins=[]
for l in lines[start+1:]:
    if (l.startswith('.text.') or l.startswith('//-----')) and ins: break
    m=re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)',l)
    if m: cur=(m.group(1).split('/')[-1],int(m.group(2))); continue
    m=re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);',l)
    if m: ins.append((cur,m.group(2)))
rows=list(csv.reader(open(sasscsv)))
hdr=None;sec=[];n=0
for r in rows:
    if r and r[0]=="Kernel Name":
        n+=1
        if n>1: break
        continue
    if r and r[0]=="Address": hdr=r; continue
    if len(r)>6: sec.append(r)
assert len(ins)==len(sec)
src=open(srcf).read().split('\n')
marks=[]
for i,l in enumerate(src,1):
    if l.startswith(('__device__','__global__','static ','extern "C"')) and '(' in l:
        head=l.split('(')[0].split(); nm=head[-1].split('<')[0]
        if nm in ('__launch_bounds__',): nm=l.split(')')[1].split('(')[0].split()[-1]
        marks.append((i,nm))
def region(f,ln):
    if f!='b200aug_fused.cu': return f
    name='top'
    for i,nm in marks:
        if i<=ln: name=nm
        else: break
    return name
cols=[h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
idx={c:hdr.index(c) for c in cols}
agg=collections.defaultdict(lambda: collections.Counter())
for (cur,_),r in zip(ins,sec):
    k=region(*cur) if cur else 'none'
    for c in cols:
        try: agg[k][c]+=int(r[idx[c]])
        except: pass
tot=sum(sum(v.values()) for v in agg.values())
for k,v in sorted(agg.items(), key=lambda kv:-sum(kv[1].values()))[:10]:
    s=sum(v.values())
    print(f"{k:28s} {100*s/tot:5.1f}% :", ', '.join(f"{c[6:]} {100*x/s:.0f}%" for c,x in v.most_common(6)))
