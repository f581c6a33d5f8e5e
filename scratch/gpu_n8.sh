#!/bin/bash
mkdir -p gpurun_out
N=${NG:-8}
run() { tag=$1; shift; env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N --steps 200 --warmup 10 --no-cpu $EXTRA > gpurun_out/n${N}_$tag.json 2> gpurun_out/n${N}_$tag.err; echo "$tag rc=$?"; }
EXTRA="--kernel-series" run push X=1
EXTRA="--no-e2e" run pull DUALIP_PEER_PUSH=0
EXTRA="--no-e2e" run two DUALIP_ONE_LAUNCH=0
EXTRA="--exchange nccl" run nccl X=1
python - <<PY
import json
for f in ["push","pull","two","nccl"]:
    try:
        d=json.loads(open(f"gpurun_out/n${N}_{f}.json").read().strip().splitlines()[-1])
        print(f, "it/s %.1f ms/step %.4f kernel_ms %.4f (min %.4f max %.4f) e2e %s launches %s replicas %s"%(d["value"],d["ms_per_step"],d["roofline"]["kernel_ms"],d["roofline"]["kernel_ms_min"],d["roofline"]["kernel_ms_max"], d.get("e2e",{}).get("value"), d["gpu_launches"], d.get("replicas")))
    except Exception as e:
        print(f, "ERR", e); print(open(f"gpurun_out/n${N}_{f}.err").read()[-1500:])
PY
