#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_distributed.py -m gpu -x -q 2>&1 | grep -v Warning | tail -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 10 > gpurun_out/n2.json 2> gpurun_out/n2.err; echo "n2 rc=$?"
DUALIP_ONE_LAUNCH=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 100 --warmup 10 --no-e2e > gpurun_out/n2_two.json 2> gpurun_out/n2_two.err; echo "n2two rc=$?"
python - <<'PY'
import json
for f in ["n2","n2_two"]:
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "it/s %.1f ms/step %.4f kernel_ms %.4f frac %.3f e2e %s launches %s replicas %s"%(d["value"],d["ms_per_step"],d["roofline"]["kernel_ms"],d["roofline"]["frac"], d.get("e2e",{}).get("value"), d["gpu_launches"], d.get("replicas")))
    except Exception as e:
        print(f, "ERR", e); print(open(f"gpurun_out/{f}.err").read()[-1500:])
PY
