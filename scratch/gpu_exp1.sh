#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/exp1.jsonl; : > $out
run() { env "$@" timeout 300 python scratch/kbench.py 2>/dev/null | tail -1 >> $out; }
for N in 100000000 12500000; do
  run N=$N
  run N=$N DUALIP_B200_LIB=$PWD/scratch/variants/sortfirst.so
done
run N=12500000 DUALIP_STAGE=20
run N=12500000 DUALIP_STAGE=20 DUALIP_B200_LIB=$PWD/scratch/variants/sortfirst.so
run N=12500000 DUALIP_STAGE=12
cat $out
# source-level profile of the baseline at full size, late iterate
N=100000000 REPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:matching_slab -s 153 -c 1 -f -o gpurun_out/prof_full_late python scratch/kbench.py > gpurun_out/ncu_full.log 2>&1
ncu -i gpurun_out/prof_full_late.ncu-rep --page source --csv > /tmp/src.csv 2>/dev/null && python tools/ncu_source_summary.py /tmp/src.csv 60 > gpurun_out/prof_full_late_source_summary.txt
ncu -i gpurun_out/prof_full_late.ncu-rep --page raw --csv > gpurun_out/prof_full_late_raw.csv 2>/dev/null
ls -la gpurun_out/*.ncu-rep
