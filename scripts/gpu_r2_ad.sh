#!/bin/bash
# round 2, call ad: launch list of a k = 120 step (where do the 70 us outside the sweep go?)
mkdir -p gpurun_out
NS=1250000 CASES=256x120 REPS=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ad_k120.csv python scripts/probe_step.py > /dev/null 2>&1
python - <<'PY'
import csv, re
rows=list(csv.reader(open('gpurun_out/ad_k120.csv')))
hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
H=rows[hdr]; ki=H.index('Kernel Name'); vi=H.index('Metric Value'); gi=H.index('Grid Size')
seq=[(re.sub(r'\(.*','',re.sub(r'sg::<unnamed>::|sg::|<unnamed>::|void ','',r[ki])), r[gi], float(r[vi])/1e3) for r in rows[hdr+1:] if len(r)>vi]
for k,g,v in seq[-12:]: print('%-40s %-14s %8.1f us'%(k[:40],g,v))
PY
