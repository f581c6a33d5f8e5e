#!/bin/bash
mkdir -p gpurun_out
N=${NG:-2}
timeout 900 python -m pytest tests/test_gpu_distributed.py tests/test_gpu_peer.py -m gpu -x -q 2>&1 | grep -v Warning | tail -3
run() { tag=$1; shift; env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N --steps 200 --warmup 10 --no-cpu --no-e2e > gpurun_out/n${N}_$tag.json 2> gpurun_out/n${N}_$tag.err; echo "$tag rc=$?"; }
run grid_on X=1
run grid_off DUALIP_GRID_TAIL=0
run grid_on2 X=1
run grid_off2 DUALIP_GRID_TAIL=0
python - <<PY
import json
for f in ["grid_on","grid_off","grid_on2","grid_off2"]:
    try:
        d=json.loads(open(f"gpurun_out/n${N}_{f}.json").read().strip().splitlines()[-1])
        print(f, "it/s %.1f ms/step %.4f kernel_ms %.4f (min %.4f max %.4f) replicas %s obj %.6f"%(d["value"],d["ms_per_step"],d["roofline"]["kernel_ms"],d["roofline"]["kernel_ms_min"],d["roofline"]["kernel_ms_max"], d.get("replicas"), d["final_dual_objective"]))
    except Exception as e:
        print(f, "ERR", e); print(open(f"gpurun_out/n${N}_{f}.err").read()[-1500:])
PY
