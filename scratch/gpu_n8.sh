#!/bin/bash
mkdir -p gpurun_out
N=${NG:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 200 --warmup 10 --kernel-series --no-cpu > gpurun_out/n${N}.json 2> gpurun_out/n${N}.err; echo "rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 200 --warmup 10 --no-e2e --no-cpu --graph on > gpurun_out/n${N}_graph.json 2> gpurun_out/n${N}_graph.err; echo "rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 200 --warmup 10 --no-e2e --no-cpu --exchange nccl > gpurun_out/n${N}_nccl.json 2> gpurun_out/n${N}_nccl.err; echo "rc=$?"
python - <<PY
import json
for f in ["n$N","n${N}_graph","n${N}_nccl"]:
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "it/s %.1f ms/step %.4f kernel_ms %.4f (min %.4f max %.4f) frac %.3f e2e %s launches %s replicas %s"%(d["value"],d["ms_per_step"],d["roofline"]["kernel_ms"],d["roofline"]["kernel_ms_min"],d["roofline"]["kernel_ms_max"],d["roofline"]["frac"], d.get("e2e",{}).get("value"), d["gpu_launches"], d.get("replicas")))
    except Exception as e:
        print(f, "ERR", e); print(open(f"gpurun_out/{f}.err").read()[-1500:])
PY
