#!/bin/bash
# session 2, step B: graph replay + row equilibration: new tests, full GPU suite, C3 / C2 bench lines
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_graph.py -m gpu -x -q 2>&1 | grep -v Warning | tail -15
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
grep -v Warning gpurun_out/pytest.log | tail -6
timeout 900 python bench.py --no-cpu > gpurun_out/c3_n1.json 2> gpurun_out/c3_n1.err; echo "c3 rc=$?"
timeout 600 python bench.py --workload c2 --steps 2000 --warmup 50 --no-cpu > gpurun_out/c2_n1.json 2> gpurun_out/c2_n1.err; echo "c2 rc=$?"
timeout 600 python bench.py --workload c2 --steps 2000 --warmup 50 --no-cpu --graph off --no-e2e > gpurun_out/c2_n1_nograph.json 2> gpurun_out/c2_n1_nograph.err; echo "c2 rc=$?"
DUALIP_ROW_SCALE=0 timeout 600 python bench.py --workload c2 --steps 2000 --warmup 50 --no-cpu --graph off --no-e2e > gpurun_out/c2_n1_f32acc.json 2> gpurun_out/c2_n1_f32acc.err; echo "c2 rc=$?"
python - <<'PY'
import json
for f in ["c3_n1","c2_n1","c2_n1_nograph","c2_n1_f32acc"]:
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "it/s %.1f ms/step %.4f kernel_ms %.4f frac %.3f e2e %s launches %s fixed %s rowscaled %s graph %s"%(d["value"],d["ms_per_step"],d["roofline"]["kernel_ms"],d["roofline"]["frac"], d.get("e2e",{}).get("value"), d["gpu_launches"], d["setup"]["plan"]["fixed_point"], d["setup"]["plan"].get("row_scaled"), d["config"].get("cuda_graph")))
    except Exception as e:
        print(f, "ERR", e); print(open(f"gpurun_out/{f}.err").read()[-1500:])
PY
