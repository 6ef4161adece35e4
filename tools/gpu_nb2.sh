#!/bin/bash
tag=${1:-nb2}
mkdir -p gpurun_out
MCMCB200_DEBUG=1 timeout 300 python tools/prof_c4.py 4096 200 200 2>&1 | tee gpurun_out/c4_$tag.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 30000 --launch-count 120 --csv --log-file gpurun_out/launches_$tag.csv python tools/prof_c4.py 4096 60 60 > gpurun_out/ncu_$tag.log 2>&1
tail -3 gpurun_out/ncu_$tag.log
python - <<'PY'
import csv,sys,collections
rows=[r for r in csv.reader(open('gpurun_out/launches_%s.csv' % (sys.argv[1] if len(sys.argv)>1 else 'nb2'))) if len(r)>5]
agg=collections.defaultdict(list)
hdr=None
for r in rows:
    if r[0]=='ID': hdr=r; continue
    if hdr is None: continue
    name=r[hdr.index('Kernel Name')][:60]; v=float(r[hdr.index('Metric Value')].replace(',','')); u=r[hdr.index('Metric Unit')]
    agg[name].append(v if u in ('us','usecond') else (v/1000 if u in ('ns','nsecond') else v*1000))
for k,v in agg.items(): print(k, len(v), 'avg us %.2f' % (sum(v)/len(v)), 'min %.2f max %.2f' % (min(v), max(v)))
PY
