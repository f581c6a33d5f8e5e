#!/bin/bash
# round 2, call A: GPU tests, warm-started bench, real-reference arms
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/a_gpu.txt 2>&1
nproc >> gpurun_out/a_gpu.txt; free -g >> gpurun_out/a_gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a_pytest.log
tail -5 gpurun_out/a_pytest.log
timeout 600 python bench.py --steps 100 --warmup 10 > gpurun_out/a_bench_c3.json 2> gpurun_out/a_bench_c3.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/a_bench_ref.json 2> gpurun_out/a_bench_ref.err; echo "ref rc=$?"
timeout 300 python bench.py --workload c2 --steps 200 --warmup 20 --no-cpu > gpurun_out/a_bench_c2.json 2> gpurun_out/a_bench_c2.err; echo "c2 rc=$?"
timeout 1500 python benchmark/reference_gpu_baseline.py --out gpurun_out/a_ref_gpu.jsonl > gpurun_out/a_ref_gpu.log 2>&1; echo "refgpu rc=$?"
tail -3 gpurun_out/a_bench_c3.err
