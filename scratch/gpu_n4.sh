#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 4 --steps 200 --warmup 10 --no-cpu > gpurun_out/n4.json 2> gpurun_out/n4.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/n4.json").read().strip().splitlines()[-1])
print("n4 it/s %.1f ms/step %.4f kernel_ms %.4f (min %.4f max %.4f) e2e %s launches %s replicas %s"%(d["value"],d["ms_per_step"],d["roofline"]["kernel_ms"],d["roofline"]["kernel_ms_min"],d["roofline"]["kernel_ms_max"], d.get("e2e",{}).get("value"), d["gpu_launches"], d.get("replicas")))
PY
