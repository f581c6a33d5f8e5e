#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_distributed.py tests/test_gpu_peer.py -m gpu -x -q 2>&1 | grep -v Warning | tail -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 10 --no-cpu > gpurun_out/n2.json 2> gpurun_out/n2.err; echo "n2 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 200 --warmup 10 --no-e2e --no-cpu --graph on > gpurun_out/n2_graph.json 2> gpurun_out/n2_graph.err; echo "n2graph rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --workload c2 --steps 2000 --warmup 50 --no-cpu > gpurun_out/n2_c2.json 2> gpurun_out/n2_c2.err; echo "n2c2 rc=$?"
python - <<'PY'
import json
for f in ["n2","n2_graph","n2_c2"]:
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "it/s %.1f ms/step %.4f kernel_ms %.4f frac %.3f e2e %s launches %s replicas %s graph %s"%(d["value"],d["ms_per_step"],d["roofline"]["kernel_ms"],d["roofline"]["frac"], d.get("e2e",{}).get("value"), d["gpu_launches"], d.get("replicas"), d["config"].get("cuda_graph")))
    except Exception as e:
        print(f, "ERR", e); print(open(f"gpurun_out/{f}.err").read()[-1500:])
PY
