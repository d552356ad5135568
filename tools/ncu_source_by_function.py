#!/usr/bin/env python3
"""`ncu --page source --csv --print-source cuda,sass` export -> warp instructions, lanes per instruction and stall samples per enclosing
device function (by source line ranges of csrc/*): python tools/ncu_source_by_function.py file.csv.gz [x = also list Trav::step lines]; SRC_REV=<commit> reads the sources of that commit"""
import collections, csv, gzip, io, os, re, subprocess, sys
rows = list(csv.reader(io.TextIOWrapper(gzip.open(sys.argv[1]))))
root='/root/repo/rust-pathtracer_b200/csrc/'
# map line -> enclosing function (crude: last line matching a definition at column 0 or 2)
def fmap(path):
    m={}; cur='?'
    rev=os.environ.get('SRC_REV')  # map against the sources of that commit (the one the profile was taken from)
    text=subprocess.run(['git','show',rev+':'+os.path.relpath(path,'/root/repo')],capture_output=True,text=True,check=True).stdout.splitlines(True) if rev else open(path)
    for i,l in enumerate(text,1):
        mm=re.match(r'^\s{0,2}(?:template\s*<[^>]*>\s*)?(?:static\s+)?(?:__device__|__global__|__host__)[^;(]*?(\w+)\s*\(', l)
        if mm: cur=mm.group(1)
        m[i]=cur
    return m
maps={}
cur_file=None; agg=collections.Counter(); agt=collections.Counter(); smp=collections.Counter()
lineagg=collections.Counter(); lineagt=collections.Counter()
for r in rows:
    if len(r)>=2 and r[0]=='File Path':
        cur_file=r[1].split('/')[-1]
        if cur_file not in maps:
            try: maps[cur_file]=fmap(root+cur_file)
            except Exception: 
                try: maps[cur_file]=fmap('/root/repo/include/'+cur_file)
                except Exception: maps[cur_file]={}
        continue
    if len(r)>=9 and r[0].isdigit():
        try: wi=int(r[7]); ti=int(r[8]); s=int(r[6])
        except ValueError: continue
        fn=maps[cur_file].get(int(r[0]),'?')
        key=(cur_file,fn)
        agg[key]+=wi; agt[key]+=ti; smp[key]+=s
        lineagg[(cur_file,int(r[0]))]+=wi; lineagt[(cur_file,int(r[0]))]+=ti
tot=sum(agg.values()); tt=sum(agt.values())
print('total warp inst',tot,'thread inst',tt,'lanes',tt/tot)
for k,v in agg.most_common(40):
    print(f'{100*v/tot:5.1f}% warp-inst  lanes {agt[k]/max(v,1):5.1f}  stall {100*smp[k]/sum(smp.values()):5.1f}%  {k[0]}:{k[1]}')
if len(sys.argv)>2:
    print('--- Trav::step lines')
    for (f,l),v in sorted(lineagg.items()):
        if f=='rpt_device.cuh' and 583<=l<=702 and v: print(l, f'{100*v/tot:5.1f}%', f'lanes {lineagt[(f,l)]/v:5.1f}')
