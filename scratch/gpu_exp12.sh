#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py tests/test_gpu_configs.py tests/test_gpu_peer.py -m gpu -x -q 2>&1 | grep -v Warning | tail -4
out=gpurun_out/exp12.jsonl; : > $out
run() { env "$@" timeout 300 python scratch/kbench.py 2>/dev/null | grep '^{' | tail -1 >> $out; }
run N=100000000
run N=100000000 DUALIP_REBALANCE=0
run N=100000000 KIND=simplex
run N=12500000
cat $out
REPS=10 DUALIP_TIMELINE=1 timeout 600 python scratch/kbench_shard.py 2>/dev/null | tail -6
REPS=10 WARM=3 DUALIP_REBALANCE=0 timeout 600 python scratch/kbench_shard.py 2>/dev/null | tail -1
N=100000000 REPS=5 DUALIP_TIMELINE=1 timeout 300 python scratch/kbench.py 2>/dev/null | tail -7
