#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_configs.py -m gpu -x -q -k "bisection" 2>&1 | grep -v "Warning\|sparse_csc" | tail -30
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
grep -v "Warning\|sparse_csc" gpurun_out/pytest.log | tail -12
