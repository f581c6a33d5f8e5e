#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --workload c1 --steps 100 --warmup 10 > gpurun_out/c1_n1.json 2> gpurun_out/c1_n1.err; echo "c1 rc=$?"; tail -c 600 gpurun_out/c1_n1.err
timeout 900 python bench.py --workload c5 --steps 3000 --warmup 50 > gpurun_out/c5_n1.json 2> gpurun_out/c5_n1.err; echo "c5 rc=$?"; tail -c 600 gpurun_out/c5_n1.err
python - <<'PY'
import json
for f in ["c1_n1","c5_n1"]:
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "it/s %.1f ms/step %.4f kernel_ms %.4f frac %.3f e2e %.1f launches %s cpu %s obj %s"%(d["value"],d["ms_per_step"],d["roofline"]["kernel_ms"],d["roofline"]["frac"], d["e2e"]["value"], d["gpu_launches"], d.get("cpu_baseline",{}).get("value"), d["final_dual_objective"]))
        print("   ", d["config"].get("column_lengths"), d.get("setup",{}).get("plan"))
    except Exception as e:
        print(f, "ERR", e)
PY
