#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | grep -v Warning | tail -15
timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu > gpurun_out/b_c3.json 2> gpurun_out/b_c3.err; echo "bench rc=$?"
timeout 300 python bench.py --workload c2 --steps 300 --warmup 20 --no-cpu > gpurun_out/b_c2.json 2> gpurun_out/b_c2.err; echo "c2 rc=$?"
DUALIP_ONE_LAUNCH=0 timeout 300 python bench.py --workload c2 --steps 300 --warmup 20 --no-cpu --no-e2e > gpurun_out/b_c2_two.json 2>/dev/null
python - <<'PY'
import json
for f in ["b_c3","b_c2","b_c2_two"]:
    d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    print(f, "it/s %.1f ms/step %.4f kernel_ms %.4f frac %.3f e2e %s launches %s"%(d["value"],d["ms_per_step"],d["roofline"]["kernel_ms"],d["roofline"]["frac"], d.get("e2e",{}).get("value"), d["gpu_launches"]))
PY
