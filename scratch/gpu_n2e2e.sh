#!/bin/bash
mkdir -p gpurun_out
N=${NG:-2}
timeout 900 python -m pytest tests/test_gpu_distributed.py -m gpu -x -q 2>&1 | grep -v Warning | tail -3
run() { tag=$1; shift; env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N --steps 100 --warmup 10 --no-cpu $EXTRA > gpurun_out/n${N}_$tag.json 2> gpurun_out/n${N}_$tag.err; echo "$tag rc=$?"; }
run e2e_peer X=1
run e2e_nccl DUALIP_PEER_EXCHANGE=0
python - <<PY
import json
for f in ["e2e_peer","e2e_nccl"]:
    try:
        d=json.loads(open(f"gpurun_out/n${N}_{f}.json").read().strip().splitlines()[-1])
        print(f, "it/s %.1f ms/step %.4f kernel_ms %.4f e2e %.1f launches %s replicas %s"%(d["value"],d["ms_per_step"],d["roofline"]["kernel_ms"], d["e2e"]["value"], d["gpu_launches"], d.get("replicas")))
    except Exception as e:
        print(f, "ERR", e); print(open(f"gpurun_out/n${N}_{f}.err").read()[-1500:])
PY
