#!/bin/bash
# session 2, step E: all-CTA tail (default for m >= 16384), full suite, c1 / c2 lines
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
grep -v "Warning\|sparse_csc" gpurun_out/pytest.log | tail -8
timeout 900 python bench.py --workload c1 --steps 100 --warmup 10 > gpurun_out/c1_n1.json 2> gpurun_out/c1_n1.err; echo "c1 rc=$?"
timeout 600 python bench.py --workload c2 --steps 2000 --warmup 50 --no-cpu > gpurun_out/c2_n1.json 2> gpurun_out/c2_n1.err; echo "c2 rc=$?"
python - <<'PY'
import json
for f in ["c1_n1","c2_n1"]:
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "it/s %.1f ms/step %.4f kernel_ms %.4f frac %.3f e2e %.1f launches %s cpu %s"%(d["value"],d["ms_per_step"],d["roofline"]["kernel_ms"],d["roofline"]["frac"], d["e2e"]["value"], d["gpu_launches"], d.get("cpu_baseline",{}).get("value")))
    except Exception as e:
        print(f, "ERR", e); print(open(f"gpurun_out/{f}.err").read()[-800:])
PY
