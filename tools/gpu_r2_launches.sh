#!/bin/bash
# per-launch durations (ncu, gpu__time_duration only) of tools/time_k1.py for one config
CFG=${1:-P80k}
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$CFG.csv python tools/time_k1.py $CFG 0x0 > gpurun_out/ncu_launches_$CFG.log 2>&1
python - <<PY
import csv
lines=open('gpurun_out/launches_$CFG.csv').read().splitlines()
h=[i for i,l in enumerate(lines) if l.startswith('"ID"')][0]
rows=list(csv.reader(lines[h:])); hdr=rows[0]; ix={x:i for i,x in enumerate(hdr)}
data=[r for r in rows[1:] if len(r)==len(hdr)]
for r in data[-14:]:
    n=r[ix['Kernel Name']].replace('tvbf::','')[:70]
    print(f"{float(r[ix['Metric Value']])/1e3:10.1f} us {r[ix['Grid Size']]:>14s} {r[ix['Block Size']]:>12s} {n}")
PY
