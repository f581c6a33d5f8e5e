#!/bin/bash
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 900 python bench.py --workload c1 --steps 100 --warmup 10 --no-cpu --no-e2e > gpurun_out/c1_$tag.json 2> gpurun_out/c1_$tag.err; echo "$tag rc=$?"; }
run box_team DUALIP_C1_PROJ=box
run box_inner DUALIP_C1_PROJ=box DUALIP_MID_KERNEL=0
python - <<'PY'
import json
for f in ["box_team","box_inner"]:
    try:
        d=json.loads(open(f"gpurun_out/c1_{f}.json").read().strip().splitlines()[-1])
        print(f, "it/s %.1f ms/step %.4f kernel_ms %.4f (min %.4f max %.4f) frac %.3f launches %s fixed %s"%(d["value"],d["ms_per_step"],d["roofline"]["kernel_ms"],d["roofline"]["kernel_ms_min"],d["roofline"]["kernel_ms_max"],d["roofline"]["frac"], d["gpu_launches"], d["setup"]["plan"]["fixed_point"]))
    except Exception as e:
        print(f, "ERR", e); print(open(f"gpurun_out/c1_{f}.err").read()[-800:])
PY
