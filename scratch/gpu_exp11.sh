#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py tests/test_gpu_configs.py -m gpu -x -q 2>&1 | grep -v Warning | tail -4
REPS=10 DUALIP_TIMELINE=1 DUMP_LAYOUT=gpurun_out/layout_shard_real2.json timeout 600 python scratch/kbench_shard.py 2>/dev/null | tail -7
