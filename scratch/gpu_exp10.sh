#!/bin/bash
mkdir -p gpurun_out
REPS=10 DUALIP_TIMELINE=1 DUMP_LAYOUT=gpurun_out/layout_shard_real.json timeout 600 python scratch/kbench_shard.py 2>/dev/null | tail -7
DUALIP_B200_LIB=$PWD/scratch/variants/old.so timeout 600 python scratch/kbench_shard.py 2>/dev/null | tail -1
