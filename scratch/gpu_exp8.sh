#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/exp8.jsonl; : > $out
run() { env "$@" timeout 300 python scratch/kbench.py 2>/dev/null | grep '^{' | tail -1 >> $out; }
for N in 100000000 12500000; do
  run N=$N
  run N=$N DUALIP_COST_SCALE_GENERIC=1.5
  run N=$N KIND=simplex
  run N=$N KIND=box
done
cat $out
for N in 100000000 12500000; do N=$N REPS=5 DUALIP_TIMELINE=1 DUMP_LAYOUT=gpurun_out/layout3_$N.json timeout 300 python scratch/kbench.py 2>/dev/null | tail -7; done
