#!/bin/bash
# final lines of round 2 at HEAD: default bench (what the driver runs), C2 and MovieLens-shaped lines without the CPU legs
mkdir -p gpurun_out
( time timeout 90 python bench.py > gpurun_out/final_c3.json 2> gpurun_out/final_c3.err ) 2>&1 | grep real; echo "c3 rc=$?"
( time timeout 60 python bench.py --workload c2 --no-cpu > gpurun_out/final_c2.json 2> gpurun_out/final_c2.err ) 2>&1 | grep real
( time timeout 60 python bench.py --workload c1 --no-cpu > gpurun_out/final_c1.json 2> gpurun_out/final_c1.err ) 2>&1 | grep real
python - <<'PY'
import json
for f in ["final_c3","final_c2","final_c1"]:
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "it/s %.1f ms/step %.4f frac %.3f kernel_ms %.4f e2e %s launches %s traffic %s"%(d["value"],d["ms_per_step"],d["roofline"]["frac"],d["roofline"]["kernel_ms"],d.get("e2e",{}).get("value"),d["gpu_launches"],d["roofline"]["traffic"]))
        print("   note:", d["roofline"].get("note","")[-110:], "| clocks", d["clocks"])
    except Exception as e:
        print(f, "ERR", e); print(open(f"gpurun_out/{f}.err").read()[-800:])
PY
