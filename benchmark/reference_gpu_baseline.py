#!/usr/bin/env python
"""Reference GPU baseline: the UNMODIFIED reference's own PyTorch-CUDA path on this box's GPU.

    python benchmark/reference_gpu_baseline.py [--out gpurun_out/ref_gpu.jsonl] [--sizes c2,10000000,25000000,100000000]

For each size and batching in {False, True} (reference benchmark/config.py:22 defaults to False): W warm-up + K timed
iterations of `AcceleratedGradientDescent.maximize` on `MatchingSolverDualObjectiveFunction` with CUDA tensors
(reference benchmark/run_matching_benchmark.py:81-107), no-op callback, wall clock between synchronisations.  One JSON
line per run.  Sizes other than c2 are the first n entities of the C3 workload (100M entities x 10k duals, mixed map,
Jacobi); a run that exhausts device memory is recorded as such and the sweep continues.
"""
from __future__ import annotations

import argparse
import gc
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--sizes", default="c2,10000000,25000000,100000000")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--budget-s", type=float, default=150.0, help="skip larger sizes once one run took longer than this")
    args = ap.parse_args()
    import torch

    from benchmark import reference_arm as R

    dev = "cuda:0"
    out = open(args.out, "a") if args.out else None

    def emit(rec):
        line = json.dumps(rec)
        print(line, flush=True)
        if out:
            out.write(line + "\n")
            out.flush()

    slow = False
    for size in args.sizes.split(","):
        if size == "c2":
            n_total, m, sp, mixed, jac, n_s, name = 1_000_000, 1_000, 1e-2, False, False, 1_000_000, "c2"
        else:
            n_total, m, sp, mixed, jac, n_s, name = 100_000_000, 10_000, 1e-3, True, True, int(size), f"c3[:{int(size)}]"
        if slow:
            emit({"workload": name, "skipped": "previous size exceeded the time budget"})
            continue
        for batching in (False, True):
            rec = {"impl": "reference-cuda", "workload": name, "entities": n_s, "duals": m, "batching": batching,
                   "steps": args.steps, "warmup": args.warmup, "gpu": torch.cuda.get_device_name(0)}
            t_all = time.perf_counter()
            try:
                torch.cuda.reset_peak_memory_stats()
                input_args, nnz = R.build_reference_problem(n_total, m, sp, 42, n_s, mixed, jac, dev, dev)
                dt, t_obj, result = R.time_reference_maximize(input_args, args.steps, args.warmup, batching, dev)
                rec.update({"nnz": nnz, "iterations_per_s": args.steps / dt, "ms_per_iteration": 1e3 * dt / args.steps,
                            "nnz_per_s": nnz * args.steps / dt, "objective_build_s": t_obj,
                            "peak_device_gb": torch.cuda.max_memory_allocated() / 1e9,
                            "final_dual_objective": float(result.dual_objective)})
                del input_args, result
            except torch.cuda.OutOfMemoryError as e:
                rec["error"] = "CUDA out of memory: " + str(e).split("\n")[0][:200]
            except Exception as e:  # keep the sweep going, record what happened
                rec["error"] = repr(e)[:300]
            gc.collect()
            torch.cuda.empty_cache()
            rec["wall_s"] = time.perf_counter() - t_all
            emit(rec)
            if rec["wall_s"] > args.budget_s:
                slow = True


if __name__ == "__main__":
    main()
