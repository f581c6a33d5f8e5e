#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:matching -c 24 --csv --log-file gpurun_out/c1_launches.csv python bench.py --workload c1 --steps 5 --warmup 3 --warm-start-iters 60 --no-cpu > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/c1_launches.csv")) if len(r)>5]
hdr=None; agg=collections.defaultdict(list)
for r in rows:
    if r[0]=="ID": hdr=r; continue
    if hdr is None: continue
    d=dict(zip(hdr,r))
    try: agg[d["Kernel Name"][:70]].append(float(d["Metric Value"].replace(",","")))
    except Exception: pass
for k,v in agg.items(): print("  %-72s n=%3d mean %.1f us (last 4: %s)"%(k,len(v),sum(v)/len(v)/1e3, [round(x/1e3,1) for x in v[-4:]]))
PY
timeout 900 python -m pytest tests/test_gpu_midcols.py tests/test_gpu_parity.py -m gpu -q 2>&1 | grep -v "Warning\|sparse_csc" | tail -5
