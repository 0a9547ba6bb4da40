#!/bin/bash
# ncu launch list of one eager pretraining step (tools/train_profile.py): tools/train_launches.sh <tag>
tag=${1:-tl}; mkdir -p gpurun_out
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_train_${tag}.csv python tools/train_profile.py > gpurun_out/launches_train_${tag}.log 2>&1
tail -2 gpurun_out/launches_train_${tag}.log | cut -c1-300
python - <<PY
import csv,collections,re
rows=list(csv.reader(open('gpurun_out/launches_train_${tag}.csv')))
hdr=None
agg=collections.defaultdict(lambda:[0,0.0])
for r in rows:
    if 'Kernel Name' in r: hdr=r; continue
    if hdr is None or len(r)<len(hdr): continue
    d=dict(zip(hdr,r))
    if d.get('Metric Name')!='gpu__time_duration.sum': continue
    n=re.sub(r'\(.*','',d['Kernel Name'])[:60]
    v=float(d['Metric Value'].replace(',',''))
    u=d['Metric Unit']
    if u=='ns': v/=1000
    elif u=='ms': v*=1000
    agg[n][0]+=1; agg[n][1]+=v
    if 'wgrad' in n: print(d['ID'], d['Grid Size'], round(v,1))
tot=sum(v[1] for v in agg.values())
for n,(c,t) in sorted(agg.items(), key=lambda x:-x[1][1])[:24]: print(f"{t:9.1f} us {c:4d} {t/c:7.1f} {n}")
print('total', tot, 'launches', sum(v[0] for v in agg.values()))
PY
