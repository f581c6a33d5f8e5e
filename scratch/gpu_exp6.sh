#!/bin/bash
mkdir -p gpurun_out
for N in 100000000 12500000; do N=$N REPS=5 DUALIP_TIMELINE=1 DUMP_LAYOUT=gpurun_out/layout_$N.json timeout 300 python scratch/kbench.py 2>/dev/null | grep '^{' | tail -1; done
