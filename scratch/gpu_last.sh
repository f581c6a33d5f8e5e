#!/bin/bash
# last GPU call of round 2: full GPU suite, smoke, the default bench line (what the driver runs)
mkdir -p gpurun_out
timeout 150 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_last.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_last.log
grep -v Warning gpurun_out/pytest_last.log | tail -6
timeout 60 python __graft_entry__.py smoke 2>&1 | tail -1
( time timeout 100 python bench.py > gpurun_out/bench_last.json 2> gpurun_out/bench_last.err ) 2>&1 | grep real; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_last.json").read().strip().splitlines()[-1])
print("native: it/s %.1f ms/step %.4f frac %.3f kernel_ms %.4f e2e %.1f launches %s clocks %s"%(d["value"],d["ms_per_step"],d["roofline"]["frac"],d["roofline"]["kernel_ms"],d["e2e"]["value"],d["gpu_launches"],d["clocks"]))
PY
