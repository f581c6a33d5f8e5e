#!/bin/bash
# state check of HEAD: GPU suite, smoke, N=1 bench lines (C3, C2) and the reference arm
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
grep -v Warning gpurun_out/pytest.log | tail -8
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/c3_n1.json 2> gpurun_out/c3_n1.err; echo "c3 rc=$?"; tail -c 2500 gpurun_out/c3_n1.json
timeout 600 python bench.py --workload c2 --steps 2000 --warmup 50 > gpurun_out/c2_n1.json 2> gpurun_out/c2_n1.err; echo "c2 rc=$?"; tail -c 1500 gpurun_out/c2_n1.json
