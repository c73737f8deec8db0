#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/time_k1.py P80k 0x0 0x10000000 0x0 > gpurun_out/time_k1_p80k.txt 2>&1; cat gpurun_out/time_k1_p80k.txt
timeout 600 python tools/time_k1.py C3 0x0 > gpurun_out/time_k1_c3.txt 2>&1; cat gpurun_out/time_k1_c3.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_c3_b.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity-check --no-dense-probe --extra '' > gpurun_out/ncu_launches.log 2>&1
python - <<'PY'
import csv
lines=open('gpurun_out/r2_launches_c3_b.csv').read().splitlines()
h=[i for i,l in enumerate(lines) if l.startswith('"ID"')][0]
rows=list(csv.reader(lines[h:])); hdr=rows[0]; ix={x:i for i,x in enumerate(hdr)}
data=[r for r in rows[1:] if len(r)==len(hdr)]
for r in data[-16:]:
    n=r[ix['Kernel Name']].replace('tvbf::','').replace('<unnamed>::','')[:60]
    print(f"{float(r[ix['Metric Value']])/1e3:10.1f} us {r[ix['Grid Size']]:>14s} {n}")
PY
