#!/bin/bash
# what the driver runs at round end: smoke, the default bench line, the reference arm
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
( time timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err ) 2>&1 | grep real; echo "bench rc=$?"
( time timeout 900 python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err ) 2>&1 | grep real; echo "ref rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_default.json").read().strip().splitlines()[-1])
print("native: it/s %.1f ms/step %.4f frac %.3f kernel_ms %.4f (min %.4f max %.4f) e2e %.1f launches %s clocks %s"%(d["value"],d["ms_per_step"],d["roofline"]["frac"],d["roofline"]["kernel_ms"],d["roofline"]["kernel_ms_min"],d["roofline"]["kernel_ms_max"],d["e2e"]["value"],d["gpu_launches"],d["clocks"]))
print("   cpu_baseline:", d["cpu_baseline"]["kind"], d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"], "port", d["cpu_baseline"].get("port",{}).get("value"))
print("   keys:", sorted(d.keys()))
r=json.loads(open("gpurun_out/bench_reference.json").read().strip().splitlines()[-1])
print("reference:", {k:r[k] for k in ("impl","value","ms_per_step","steps","warmup")}, r["cpu_baseline"]["kind"], r["cpu_baseline"]["cores"])
PY
