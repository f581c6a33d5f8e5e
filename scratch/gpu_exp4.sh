#!/bin/bash
mkdir -p gpurun_out
for S in 20 0; do
N=12500000 REPS=5 DUALIP_TIMELINE=1 DUALIP_STAGE=$S timeout 300 python scratch/kbench.py 2>/dev/null | tail -8
done
N=12500000 REPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:matching_slab -s 153 -c 1 -f -o gpurun_out/prof_shard_late python scratch/kbench.py > gpurun_out/ncu_shard.log 2>&1
ncu -i gpurun_out/prof_shard_late.ncu-rep --page source --csv > /tmp/src.csv 2>/dev/null && python tools/ncu_source_summary.py /tmp/src.csv 50 > gpurun_out/prof_shard_late_source_summary.txt
ncu -i gpurun_out/prof_shard_late.ncu-rep --page raw --csv > gpurun_out/prof_shard_late_raw.csv 2>/dev/null
N=100000000 REPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:matching_slab -s 153 -c 1 -f -o gpurun_out/prof_full_late_v14 python scratch/kbench.py > gpurun_out/ncu_full_v14.log 2>&1
ncu -i gpurun_out/prof_full_late_v14.ncu-rep --page source --csv > /tmp/src2.csv 2>/dev/null && python tools/ncu_source_summary.py /tmp/src2.csv 50 > gpurun_out/prof_full_late_v14_source_summary.txt
ncu -i gpurun_out/prof_full_late_v14.ncu-rep --page raw --csv > gpurun_out/prof_full_late_v14_raw.csv 2>/dev/null
head -5 gpurun_out/prof_shard_late_source_summary.txt
