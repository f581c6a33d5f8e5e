#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_operators.py -m gpu -x -q 2>&1 | grep -v "Warning\|sparse_csc" | tail -30
