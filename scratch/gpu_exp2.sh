#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py tests/test_gpu_configs.py -m gpu -x -q 2>&1 | grep -v Warning | tail -5
out=gpurun_out/exp2.jsonl; : > $out
run() { env "$@" timeout 300 python scratch/kbench.py 2>/dev/null | tail -1 >> $out; }
for N in 100000000 12500000; do
  run N=$N
  run N=$N DUALIP_B200_LIB=$PWD/scratch/variants/sortfirst.so
  run N=$N DUALIP_STAGE=0
  run N=$N DUALIP_STAGE_REGION=6656
  run N=$N DUALIP_STAGE_REGION=4608
done
cat $out
