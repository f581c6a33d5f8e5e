#!/bin/bash
DUALIP_TIMELINE=1 PRE=2000 timeout 300 python scratch/kbench_c2.py 2>&1 | grep -v "Warn\|sparse_csc" | tail -14
