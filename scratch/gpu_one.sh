#!/bin/bash
timeout 600 python -m pytest "tests/test_gpu_peer.py::test_peer_step_matches_summed_partials" -m gpu -x -q 2>&1 | grep -v Warning | grep -E "^E|passed|failed" | head -30
