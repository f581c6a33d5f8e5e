#!/bin/bash
# same box, back to back: bench.py of commit 67dc627 (before the bisection / mid-column / grid-tail work) and of HEAD
show() { python - "$1" <<'PY'
import json, sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]
print(sys.argv[1], "it/s %.1f kernel_ms %.4f (min %.4f max %.4f) frac %.3f clocks %s"%(d["value"], r["kernel_ms"], r["kernel_ms_min"], r["kernel_ms_max"], r["frac"], d["clocks"]["sm_mhz"]))
PY
}
for rep in 1 2; do
  (cd scratch/old_tree && timeout 300 python bench.py --no-cpu --no-e2e > /tmp/old_$rep.json 2>/dev/null); show /tmp/old_$rep.json
  timeout 300 python bench.py --no-cpu --no-e2e > /tmp/head_$rep.json 2>/dev/null; show /tmp/head_$rep.json
done
DUALIP_REBALANCE=0 timeout 300 python bench.py --no-cpu --no-e2e > /tmp/head_norebal.json 2>/dev/null; show /tmp/head_norebal.json
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv,noheader
