#!/bin/bash
mkdir -p gpurun_out
N=8
run() { tag=$1; shift; env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N --steps 300 --warmup 10 --no-cpu --no-e2e > gpurun_out/n${N}_$tag.json 2> gpurun_out/n${N}_$tag.err; echo "$tag rc=$?"; }
run gridoff X=1
run gridon DUALIP_GRID_TAIL=1
run gridoff2 X=1
run gridon2 DUALIP_GRID_TAIL=1
python - <<PY
import json
for f in ["gridoff","gridon","gridoff2","gridon2"]:
    try:
        d=json.loads(open(f"gpurun_out/n8_{f}.json").read().strip().splitlines()[-1])
        print(f, "it/s %.1f ms/step %.4f kernel_ms %.4f (min %.4f max %.4f) replicas %s obj %.6f"%(d["value"],d["ms_per_step"],d["roofline"]["kernel_ms"],d["roofline"]["kernel_ms_min"],d["roofline"]["kernel_ms_max"], d.get("replicas"), d["final_dual_objective"]))
    except Exception as e:
        print(f, "ERR", e); print(open(f"gpurun_out/n8_{f}.err").read()[-1500:])
PY
