#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py tests/test_gpu_configs.py tests/test_gpu_custom_projection.py -m gpu -x -q 2>&1 | grep -v Warning | tail -15
out=gpurun_out/exp5.jsonl; : > $out
run() { env "$@" timeout 300 python scratch/kbench.py 2>/dev/null | grep '^{' | tail -1 >> $out; }
for N in 100000000 12500000; do
  run N=$N
  run N=$N DUALIP_SLAB_BETA=0
  run N=$N DUALIP_SLAB_BETA=6
  run N=$N DUALIP_STAGE=0
done
cat $out
for N in 100000000 12500000; do N=$N REPS=5 DUALIP_TIMELINE=1 timeout 300 python scratch/kbench.py 2>/dev/null | tail -7; done
