#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_distributed.py -m gpu -x -q 2>&1 | grep -v Warning | tail -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/dist_worker.py > /tmp/w.out 2> /tmp/w.err; echo rc=$?
grep -v "Warning\|warn" /tmp/w.err | grep -A8 "Traceback" | head -30; tail -1 /tmp/w.out
