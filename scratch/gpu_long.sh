#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_midcols.py -m gpu -x -q 2>&1 | grep -v "Warning\|sparse_csc" | tail -25
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
grep -v "Warning\|sparse_csc" gpurun_out/pytest.log | tail -6
timeout 900 python bench.py --workload c1 --steps 100 --warmup 10 --no-cpu > gpurun_out/c1_n1.json 2> gpurun_out/c1_n1.err; echo "c1 rc=$?"
python - <<'PY'
import json
for f in ["c1_n1"]:
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "it/s %.1f ms/step %.4f kernel_ms %.4f (min %.4f max %.4f) frac %.3f e2e %.1f launches %s"%(d["value"],d["ms_per_step"],d["roofline"]["kernel_ms"],d["roofline"]["kernel_ms_min"],d["roofline"]["kernel_ms_max"],d["roofline"]["frac"], d["e2e"]["value"], d["gpu_launches"]))
    except Exception as e:
        print(f, "ERR", e); print(open(f"gpurun_out/{f}.err").read()[-800:])
PY
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
