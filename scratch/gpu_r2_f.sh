#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
grep -v "Warning\|sparse_csc" $O/pytest.log | tail -6
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 7 --log-file $O/sanitizer2_synccheck.log python -m pytest tests/test_gpu_graph.py tests/test_gpu_midcols.py -m gpu -q > $O/sanitizer2_synccheck_pytest.log 2>&1; echo "synccheck rc=$?"
tail -2 $O/sanitizer2_synccheck.log; tail -1 $O/sanitizer2_synccheck_pytest.log
timeout 900 python bench.py --no-cpu > $O/c3_n1.json 2> $O/c3_n1.err; echo "c3 rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/c3_n1.json").read().strip().splitlines()[-1]); r=d["roofline"]
print("c3 it/s %.1f kernel_ms %.4f (min %.4f max %.4f) frac %.3f e2e %.1f obj %.9f"%(d["value"], r["kernel_ms"], r["kernel_ms_min"], r["kernel_ms_max"], r["frac"], d["e2e"]["value"], d["final_dual_objective"]))
PY
