#!/bin/bash
# Per-level Schur kernel times (ncu launch list) for the kernel variants selected by the environment.
# usage: tools/schur_levels.sh tag [ENV=VAL ...]
tag=$1; shift
env "$@" timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_schur -c 24 --csv \
    --log-file gpurun_out/schur_levels_$tag.csv python tools/prof_one.py 3499 2 > /dev/null 2>&1
python - <<PY
import csv,re
rows=[r for r in csv.reader(open('gpurun_out/schur_levels_$tag.csv')) if len(r)>5]
hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
H=rows[hdr]; kn=H.index('Kernel Name'); mv=H.index('Metric Value')
v=[float(r[mv].replace(',',''))/1e6 for r in rows[hdr+1:]][-12:]
n=[re.sub(r'.*::k_schur_(\w+)<(\d+), (\d+), (\d+).*',r'\1\4',r[kn]) for r in rows[hdr+1:]][-12:]
print('$tag', ' '.join('%s:%.3f'%(a,b) for a,b in zip(n,v)), 'sum %.3f'%sum(v))
PY
