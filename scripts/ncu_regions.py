#!/usr/bin/env python
"""Instruction shares per function of the fused kernel: joins the SASS page of an ncu report (`ncu -i X.ncu-rep --page source --csv --print-source sass > sass.csv`) with `nvdisasm -g -c` line info of the same build (`nvcc ... -cubin`).  usage: python scripts/ncu_regions.py kernel.dis sass.csv b200aug_fused.cu"""
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
sec=[];n=0
for r in rows:
    if r and r[0]=="Kernel Name":
        n+=1
        if n>1: break
        continue
    if r and r[0]=="Address": continue
    if len(r)>6: sec.append(r)
assert len(ins)==len(sec),(len(ins),len(sec))
src=open(srcf).read().split('\n')
# regions by function: find function start lines
marks=[]
for i,l in enumerate(src,1):
    if l.startswith(('__device__','__global__','static ','extern "C"')) and '(' in l:
        head=l.split('(')[0].split()
        nm=head[-1].split('<')[0]
        if nm in ('__launch_bounds__',): nm=l.split(')')[1].split('(')[0].split()[-1]
        marks.append((i,nm))
# kernel sub-regions by "// ----" comments inside kernel
kstart=[i for i,nm in marks if nm=='fused_augment_kernel'][0]
kend=[i for i,nm in marks if i>kstart][0]
sub=[(i,'k:'+src[i-1].strip()[:50]) for i in range(kstart,kend) if src[i-1].strip().startswith('// ----')]
allm=sorted(marks+sub)
def region(f,ln):
    if f!='b200aug_fused.cu': return f
    name='top'
    for i,nm in allm:
        if i<=ln: name=nm
        else: break
    return name
agg=collections.Counter(); smp=collections.Counter()
for (cur,_),r in zip(ins,sec):
    k=region(*cur) if cur else 'none'
    agg[k]+=int(r[5]); smp[k]+=int(r[4])
tot=sum(agg.values()); ts=sum(smp.values())
print('total',tot,'samples',ts)
for k,v in agg.most_common(40): print(f"{k:60s} inst {100*v/tot:5.1f}% ({v/1e6:5.2f}M) samp {100*smp[k]/ts:5.1f}%")
