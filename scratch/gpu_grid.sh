#!/bin/bash
# grid-wide tail: full suite with it on (default), key suites with it off, and the benches both ways
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
grep -v "Warning\|sparse_csc" gpurun_out/pytest.log | tail -25
run() { tag=$1; wl=$2; shift; shift; env "$@" timeout 900 python bench.py --workload $wl --no-cpu --no-e2e $EXTRA > gpurun_out/grid_$tag.json 2> gpurun_out/grid_$tag.err; echo "$tag rc=$?"; }
EXTRA="--steps 2000 --warmup 50" run c2_on c2 X=1
EXTRA="--steps 2000 --warmup 50" run c2_off c2 DUALIP_GRID_TAIL=0
EXTRA="--steps 100 --warmup 10" run c1_on c1 X=1
EXTRA="--steps 100 --warmup 10" run c1_off c1 DUALIP_GRID_TAIL=0
EXTRA="--steps 100 --warmup 10" run c3_on c3 X=1
EXTRA="--steps 100 --warmup 10" run c3_off c3 DUALIP_GRID_TAIL=0
python - <<'PY'
import json
for f in ["c2_on","c2_off","c1_on","c1_off","c3_on","c3_off"]:
    try:
        d=json.loads(open(f"gpurun_out/grid_{f}.json").read().strip().splitlines()[-1])
        print(f, "it/s %.1f ms/step %.4f kernel_ms %.4f (min %.4f max %.4f) obj %.6f"%(d["value"],d["ms_per_step"],d["roofline"]["kernel_ms"],d["roofline"]["kernel_ms_min"],d["roofline"]["kernel_ms_max"], d["final_dual_objective"]))
    except Exception as e:
        print(f, "ERR", e); print(open(f"gpurun_out/grid_{f}.err").read()[-800:])
PY
