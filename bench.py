#!/usr/bin/env python
"""Benchmark of the dual-ascent hot path.  One JSON line on stdout (rank 0).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repository's CUDA path
    python bench.py --impl reference [--steps K] [--warmup W]      # CPU arm: the oracle port on all host threads

Metric (BASELINE.json): dual-ascent iterations/sec on the synthetic matching LP with 100M entities x 10k duals
(sparsity 1e-3, ~1e9 nonzeros), simplex(z=1) on even entities and box[0,1] on odd ones, Jacobi row preconditioning,
Nesterov AGD with initial_step_size 1e-3 / max_step_size 1e-1, gamma 1e-3.  A "step" is one full iteration:
evaluate the dual at lambda (fused kernel, + one all-reduce when sharded), then the accelerated update.
With N GPUs the SAME 100M-entity problem is sharded by contiguous entity ranges (strong scaling).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (entities, duals, sparsity, mixed projection map?, jacobi?)
    "c3": (100_000_000, 10_000, 1e-3, True, True),
    "c2": (1_000_000, 1_000, 1e-2, False, False),
    "c1": (138_493, 26_744, None, False, False),  # MovieLens-shaped (benchmark/extra_workloads.py)
    "c5": (150, 7_822, None, False, False),       # MIPLIB-2017 example LP (benchmark/extra_workloads.py)
    "c3_small": (10_000_000, 10_000, 1e-3, True, True),
    "tiny": (200_000, 1_000, 1e-2, True, True),
}
GAMMA = 1e-3
INITIAL_STEP, MAX_STEP = 1e-3, 1e-1  # reference benchmark/config.py:17-18
SEED = 42  # reference benchmark/config.py:12
L2_BYTES = 126 * 1024 * 1024


def log(*a):
    print(*a, file=sys.stderr, flush=True)


class ClockSampler:
    """Samples SM clocks / throttle reasons with nvidia-smi while the timed region runs."""

    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu_index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return

        def pump():
            for line in self.proc.stdout:
                self.rows.append(line.strip())

        self.thread = threading.Thread(target=pump, daemon=True)
        self.thread.start()

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for r in self.rows:
            f = [t.strip() for t in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


NCU_TRAFFIC_FILES = ("profiles/r2/ncu_traffic.json", "profiles/r1_ncu_traffic.json")  # newest capture first


def ncu_traffic_bytes(workload_id: str):
    """(dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, file it came from) from the committed
    `ncu --set full` capture of this workload, or (None, None) if there is none."""
    for rel in NCU_TRAFFIC_FILES:
        try:
            with open(os.path.join(ROOT, rel)) as fh:
                entry = json.load(fh)[workload_id]
            return float(entry["traffic_bytes_per_launch"]), entry.get("source", rel)
        except Exception:
            continue
    return None, None


def mixed_projection_map(n_local: int, col_start: int, device):
    """simplex(z=1) on even global entities, box[0,1] on odd ones (our choice of mixed map)."""
    import torch

    from dualip_b200.projections import create_projection_map

    first_even = (col_start % 2)  # local index of the first even global column
    even = torch.arange(first_even, n_local, 2, device=device)
    odd = torch.arange(1 - first_even, n_local, 2, device=device)
    pm = {}
    pm.update(create_projection_map("simplex", {"z": 1.0}, n_local, indices=even))
    pm.update(create_projection_map("box", {"lower": 0.0, "upper": 1.0}, n_local, indices=odd))
    return pm


def build_problem(args, rank, world, device):
    """Generates this rank's entity shard on its GPU and returns (objective, b, info)."""
    import torch
    import torch.distributed as dist

    from benchmark.synthetic import capacity_vector, generate_shard
    from dualip_b200.objectives.matching import (
        MatchingInputArgs,
        MatchingSolverDualObjectiveFunction,
        MatchingSolverDualObjectiveFunctionDistributed,
    )
    from dualip_b200.preprocessing.precondition import jacobi_precondition
    from dualip_b200.projections import create_projection_map
    from dualip_b200.utils.dist_utils import shard_sizes

    n, m, sparsity, mixed, jacobi = args.entities, args.duals, args.sparsity, args.mixed, args.jacobi
    sizes = shard_sizes(n, world)
    col_start = sum(sizes[:rank])
    col_end = col_start + sizes[rank]
    t0 = time.time()
    shard = generate_shard(n, m, sparsity, SEED, device, col_start, col_end)
    load = shard.greedy_load
    if world > 1:
        dist.all_reduce(load)
    b = capacity_vector(load, m, sparsity, SEED, device)
    n_local = col_end - col_start
    A = torch.sparse_csc_tensor(shard.ccol, shard.row, shard.a, size=(m, n_local))
    C = torch.sparse_csc_tensor(shard.ccol, shard.row, shard.c, size=(m, n_local))
    if jacobi:
        jacobi_precondition(A, b, sharded=True)  # global row norms (one all-reduce of m doubles when sharded)
    pm = mixed_projection_map(n_local, col_start, device) if mixed else create_projection_map("simplex", {"z": 1.0}, n_local)
    torch.cuda.synchronize(device)
    t_gen = time.time() - t0
    t0 = time.time()
    if world == 1:
        obj = MatchingSolverDualObjectiveFunction(MatchingInputArgs(A, C, pm, b), gamma=GAMMA)
        local = obj
    else:
        obj = MatchingSolverDualObjectiveFunctionDistributed(MatchingInputArgs(A, C, pm, None), b, GAMMA, host_device=device)
        local = obj.local_objective
    torch.cuda.synchronize(device)
    t_plan = time.time() - t0
    nnz_local = local.nnz
    nnz_total = torch.tensor([nnz_local], dtype=torch.int64, device=device)
    if world > 1:
        dist.all_reduce(nnz_total)
    info = dict(nnz_local=nnz_local, nnz_total=int(nnz_total.item()), n_local=n_local, gen_s=round(t_gen, 2),
                plan_s=round(t_plan, 2), plan=local.plan_info(), col_start=col_start, col_end=col_end)
    return obj, local, b, shard, info


def cpu_baseline_port(shard, b, args, n_sample_cols, threads, repeats=3):
    """Times the C restatement (oracle/matching_oracle.c, OpenMP) on a column sample of this workload."""
    import numpy as np
    import torch

    from oracle import c_oracle

    n_s = min(n_sample_cols, shard.ccol.numel() - 1)
    e_s = int(shard.ccol[n_s].item())
    ccol = shard.ccol[: n_s + 1].cpu().numpy()
    row = shard.row[:e_s].cpu().numpy()
    a = shard.a[:e_s].cpu().numpy()
    c = shard.c[:e_s].cpu().numpy()
    m = args.duals
    if args.mixed:
        classes = [c_oracle.make_class("simplex", {"z": 1.0}), c_oracle.make_class("box", {"lower": 0.0, "upper": 1.0})]
        col_class = ((np.arange(n_s) + shard.col_start) % 2).astype(np.uint8)
    else:
        classes, col_class = [c_oracle.make_class("simplex", {"z": 1.0})], None
    lam = np.zeros(m, dtype=np.float32)
    bh = b.cpu().numpy()
    c_oracle.calculate(ccol, row, a, c, m, classes, lam, GAMMA, bh, col_class, want_x=False, want_diag=False, threads=threads)
    times = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        r = c_oracle.calculate(ccol, row, a, c, m, classes, lam, GAMMA, bh, col_class, want_x=False, want_diag=False, threads=threads)
        times.append(time.perf_counter() - t0)
        lam = np.maximum(lam + np.float32(INITIAL_STEP) * r["grad"], 0).astype(np.float32)
    return min(times), e_s, n_s


def run_native(args):
    import torch
    import torch.distributed as dist

    from dualip_b200.optimizers.agd import AcceleratedGradientDescent, FusedAscentLoop, no_iteration_callback

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL announces its version on stdout when the first communicator comes up; stdout carries the JSON line only
        sys.stdout.flush()
        try:
            saved = os.dup(1)
            os.dup2(2, 1)
        except OSError:
            saved = None
        try:
            dist.init_process_group("nccl", device_id=device)
            dist.barrier()
            torch.cuda.synchronize(device)
        finally:
            if saved is not None:
                sys.stdout.flush()
                os.dup2(saved, 1)
                os.close(saved)
    if args.gpus != world:
        log(f"[bench] note: --gpus {args.gpus} but WORLD_SIZE={world}; using WORLD_SIZE")

    os.environ["DUALIP_PEER_EXCHANGE"] = "1" if args.exchange == "peer" else "0"
    obj, local, b, shard, info = build_problem(args, rank, world, device)
    if rank == 0:
        log(f"[bench] problem ready: {info}")
    K, W = args.steps, args.warmup
    m = args.duals

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    def max_over_ranks(v: float) -> float:
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- warm start ("lambda_0 loaded from a saved C3 run"): an untimed run from zero produces the
    #      saved dual; the measured run starts from it, so the timed iterations see a late iterate whatever --steps is
    #      (from zero the first iterations are dominated by the cheap top-2 shortcut) ----
    lam0 = torch.zeros(m, dtype=torch.float32, device=device)
    if args.warm_start_iters > 0:
        pre = AcceleratedGradientDescent(max_iter=args.warm_start_iters, gamma=GAMMA, initial_step_size=INITIAL_STEP,
                                         max_step_size=MAX_STEP, iteration_callback=no_iteration_callback)
        lam0 = pre.maximize(obj, lam0, rank=rank).dual_val.clone()
        barrier()

    # ---- device-resident loop: W warm-up + K timed iterations ----
    solver = AcceleratedGradientDescent(max_iter=W + K, gamma=GAMMA, initial_step_size=INITIAL_STEP, max_step_size=MAX_STEP,
                                        iteration_callback=no_iteration_callback)
    loop = FusedAscentLoop(solver, obj, lam0, rank)
    b_alg = local.algorithmic_bytes()
    use_graph = loop.graphable and (args.graph == "on" or (args.graph == "auto" and b_alg <= 2 * L2_BYTES))
    if use_graph:
        while not local.plan_settled():  # a graph is captured once the plan has stopped re-cutting its ranges (128 launches)
            local.calculate(lam0)
        loop.graph_chunk = max(1, min(loop.graph_chunk, K))
    for i in range(1, W + 1):
        loop.step(i)
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    if not use_graph:
        loop.kernel_events = kev  # FusedAscentLoop brackets the objective's kernel(s) of step W+j with kev[j]
        loop.kernel_events_base = W + 1
    torch.cuda.nvtx.range_push("dualip_timed_region")  # ncu --nvtx --nvtx-include "dualip_timed_region/" lists exactly these launches
    ev0.record()
    if use_graph:
        loop.run(W + 1, W + K)  # whole chunks as graph replays, the remainder one by one
    else:
        for i in range(W + 1, W + K + 1):
            loop.step(i)
    ev1.record()
    torch.cuda.nvtx.range_pop()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    result = loop.finish()
    lam_now = loop.current_dual()
    loop.close()
    # every rank updates its own replica of the dual; they must never drift apart: compare bit patterns across ranks
    replicas = None
    if world > 1:
        bits = lam_now.view(torch.int32).to(torch.int64)
        hi, lo = bits.clone(), bits.clone()
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        replicas = {"ranks": world, "dual_bit_identical_on_all_ranks": bool((hi == lo).all().item())}
    # evaluation and accelerated step share ONE launch (the kernel's last CTA steps); the two-launch paths add the update kernel
    launches_per_step = info["plan"]["launches_per_calc"] + (0 if loop.one_launch and (world == 1 or loop.peer is not None) else 1)
    exchange = "none (single GPU)" if world == 1 else (
        "peer memory: the shard kernel's last CTA publishes its sums, reads the peers' over NVLink and steps; no collective call, one launch per iteration" if loop.peer is not None
        else "one NCCL all_reduce of m+2 floats per iteration")

    # ---- dominant kernel, for the roofline: CUDA events around every launch of the timed region ----
    graph_info = None
    if use_graph:
        # events cannot bracket kernels inside a graph: the same K iterations again, launched one by one, only for kernel_ms
        graph_info = {"chunk": loop.graph_chunk, "replays": loop.graph_launches,
                      "kernel_ms_from": "a second pass of the same K iterations launched one by one with CUDA events"}
        solver2 = AcceleratedGradientDescent(max_iter=W + K, gamma=GAMMA, initial_step_size=INITIAL_STEP, max_step_size=MAX_STEP,
                                             iteration_callback=no_iteration_callback)
        loop2 = FusedAscentLoop(solver2, obj, lam0, rank)
        loop2.kernel_events, loop2.kernel_events_base = kev, W + 1
        for i in range(1, W + K + 1):
            loop2.step(i)
        loop2.finish()
        loop2.close()
    torch.cuda.synchronize(device)
    kernel_times = [a.elapsed_time(b) for a, b in kev]
    kernel_ms = sum(kernel_times) / len(kernel_times)
    peak, peak_src = measured_peak_gbs()
    achieved = b_alg / (kernel_ms * 1e-3) / 1e9

    # ---- end to end through the public API with host buffers: same schedule (W warm-up + K timed iterations) ----
    e2e = None
    if not args.no_e2e:
        Ke = K if args.e2e_steps <= 0 else min(K, args.e2e_steps)
        lam_host = lam0.cpu().pin_memory()

        def host_run(iters):
            # the call a user makes: maximize() with a host-resident dual and no per-iteration reporting; every rank runs the
            # host loop on its own copy of lambda
            hs = AcceleratedGradientDescent(max_iter=iters, gamma=GAMMA, initial_step_size=INITIAL_STEP, max_step_size=MAX_STEP,
                                            iteration_callback=no_iteration_callback)
            barrier()
            t0 = time.perf_counter()
            hs.maximize(obj, lam_host, rank=rank)
            torch.cuda.synchronize(device)
            return max_over_ranks(time.perf_counter() - t0)

        host_run(W)  # first use: windows, pinned buffers
        t_w = host_run(W)
        t_wk = host_run(W + Ke)
        dt = max(t_wk - t_w, 1e-9)  # iterations W+1 .. W+Ke: the same iterations as `value`
        h2d, d2h = obj.host_io_bytes()
        e2e = {"value": Ke / dt, "unit": "iterations/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "steps": Ke, "path": "AcceleratedGradientDescent.maximize with a pinned host dual vector: per iteration lambda "
               "host->device, fused kernel(s), grad+scalars device->host, host-side update (one native call per iteration); "
               "wall clock of a W+K-iteration run minus a W-iteration run, max over ranks"}

    # ---- CPU baseline (rank 0, N=1 only): the unmodified reference (oracle/_ref) on a bounded sample, the C port beside it ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from benchmark import reference_arm as R
        from oracle import make_ref

        threads = R.host_threads()
        t_s, e_s, n_s = cpu_baseline_port(shard, b, args, args.cpu_sample_cols, threads)
        port = {"value": 1.0 / (t_s * info["nnz_total"] / max(e_s, 1)), "unit": "iterations/s", "cores": threads, "kind": "port",
                "sample": f"first {n_s} entities ({e_s} nnz), best of 3 evaluations of the dual with the C/OpenMP restatement: "
                          f"{e_s / t_s / 1e6:.1f} M nnz/s, scaled by nnz", "nnz_per_s": e_s / t_s}
        cpu = port
        if make_ref.available():
            del shard
            torch.set_num_threads(threads)
            n_r = min(args.ref_sample_cols, args.entities)
            ref_args, e_r = R.build_reference_problem(args.entities, m, args.sparsity, SEED, n_r, args.mixed, args.jacobi, "cpu", device)
            steps_r = 8
            dt, _, _ = R.time_reference_maximize(ref_args, steps_r, 2, args.ref_batching, "cpu")
            nnz_s = e_r * steps_r / dt
            cpu = {"value": nnz_s / info["nnz_total"], "unit": "iterations/s", "cores": threads, "kind": "reference",
                   "sample": f"UNMODIFIED reference (oracle/_ref) maximize() on torch CPU ({threads} threads), first {n_r} entities "
                             f"({e_r} nnz), {steps_r} iterations after 2 warm-up, batching={args.ref_batching}: {nnz_s / 1e6:.1f} M nnz/s, "
                             f"scaled by nnz to the full problem", "nnz_per_s": nnz_s, "port": port}

    if rank == 0:
        it_per_s = K / (ms_total * 1e-3)
        if args.ncu_traffic_bytes is not None:
            traffic, traffic_src = args.ncu_traffic_bytes, "--ncu-traffic-bytes"
        elif world == 1 and args.entities == WORKLOADS[args.workload][0]:
            traffic, traffic_src = ncu_traffic_bytes(args.workload)
        else:
            traffic, traffic_src = None, None
        line = {
            "metric": "dual-ascent iterations/sec", "value": it_per_s, "unit": "iterations/s", "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"synthetic matching LP {args.entities} entities x {args.duals} duals, sparsity {args.sparsity}, "
                                   f"{'simplex(z=1) even / box[0,1] odd' if args.mixed else 'simplex(z=1)'}, "
                                   f"{'Jacobi precond, ' if args.jacobi else ''}Nesterov AGD, gamma={GAMMA}",
                       "workload_id": args.workload, "entities": args.entities, "duals": args.duals, "nnz": info["nnz_total"],
                       "parallelism": f"entity-sharded x{world}" if world > 1 else "single GPU", "exchange": exchange,
                       "l2": "inputs (%.2f GB per GPU) exceed the 126 MB L2; no flush between iterations" % (b_alg / 1e9)
                       if b_alg > 2 * L2_BYTES else "inputs fit in L2: numbers are L2-resident, not a roofline claim",
                       "index_dtype": "int64 inputs, uint16 row ids in the plan",
                       "warm_start": (f"lambda after {args.warm_start_iters} untimed iterations from zero (BASELINE.md C4: saved-dual "
                                      f"warm start); fresh optimizer state") if args.warm_start_iters > 0 else "none (lambda_0 = 0)"},
            "entities_x_constraints_per_s": args.entities * args.duals * it_per_s,
            "nnz_per_s": info["nnz_total"] * it_per_s,
            "gpu_launches": launches_per_step * K,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic,
                         "kernel": "matching_slab_kernel", "kernel_ms": kernel_ms,
                         "kernel_ms_min": min(kernel_times), "kernel_ms_max": max(kernel_times),
                         "algorithmic_bytes": b_alg, "peak_source": peak_src,
                         "note": ("rank-0 shard; " if world > 1 else "") + "average over the K launches of the timed region "
                                 "(the kernel's cost depends on the iterate: how many simplex columns take the sorted scan); "
                                 + (f"traffic: ncu --set full capture of this workload ({traffic_src})" if traffic is not None
                                    else "traffic: no ncu capture of this configuration")},
            "final_dual_objective": result.dual_objective,
            "setup": {"generate_s": info["gen_s"], "plan_s": info["plan_s"], "plan": info["plan"]},
        }
        if graph_info is not None:
            line["config"]["cuda_graph"] = graph_info
        if args.kernel_series:
            line["roofline"]["kernel_ms_series"] = [round(t, 4) for t in kernel_times]
        if e2e is not None:
            line["e2e"] = e2e
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if replicas is not None:
            line["replicas"] = replicas
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_reference(args):
    """Reference arm: the UNMODIFIED reference (oracle/_ref = `pip install --target` of linkedin/DuaLip v5.0.1, see
    oracle/make_ref.py) on the host cores of this box: its own MatchingSolverDualObjectiveFunction +
    AcceleratedGradientDescent.maximize on torch CPU tensors, all hardware threads (set explicitly: torchrun exports
    OMP_NUM_THREADS=1).  Each step is one dual-ascent iteration on a bounded entity sample of the same workload; `value`
    is scaled by nnz to the full problem.  The C/OpenMP port of the oracle is timed beside it (`port`)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from benchmark import reference_arm as R

    threads = R.host_threads()
    os.environ["OMP_NUM_THREADS"] = str(threads)  # before torch / libgomp start their pools
    os.environ["MKL_NUM_THREADS"] = str(threads)
    import numpy as np
    import torch

    torch.set_num_threads(threads)
    from oracle import make_ref

    n, m = args.entities, args.duals
    K, W = args.steps, args.warmup
    config = {"workload": f"synthetic matching LP {args.entities} entities x {args.duals} duals, sparsity {args.sparsity}, "
                          f"{'simplex(z=1) even / box[0,1] odd' if args.mixed else 'simplex(z=1)'}, "
                          f"{'Jacobi precond, ' if args.jacobi else ''}Nesterov AGD, gamma={GAMMA}",
              "workload_id": args.workload, "entities": n, "duals": m, "parallelism": "host threads (CPU arm)"}
    line = {"impl": "reference", "metric": "dual-ascent iterations/sec", "unit": "iterations/s", "n_gpus": args.gpus,
            "steps": K, "warmup": W, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": config}

    # ---- the C/OpenMP port of the oracle on a larger sample (kept as context: 10-20x faster than the reference) ----
    port = None
    try:
        port = _time_port(args, threads, min(args.cpu_sample_cols, n))
    except Exception as e:  # the port is context, never the headline
        log(f"[bench] port timing failed: {e!r}")

    if make_ref.available():
        n_s = min(args.ref_sample_cols, n)
        gen_device = "cuda" if torch.cuda.is_available() else "cpu"
        input_args, e_s = R.build_reference_problem(n, m, args.sparsity, SEED, n_s, args.mixed, args.jacobi, "cpu", gen_device)
        e_full = e_s * (n / n_s)
        dt, t_obj, result = R.time_reference_maximize(input_args, K, W, args.ref_batching, "cpu")
        it_s_sample = K / dt
        it_s_full = it_s_sample * (e_s / e_full)
        config.update({"nnz": int(round(e_full)), "sample_entities": n_s, "batching": args.ref_batching})
        line.update({
            "value": it_s_full, "ms_per_step": 1e3 / it_s_full,
            "cpu_baseline": {"value": it_s_full, "unit": "iterations/s", "cores": threads, "kind": "reference",
                             "torch_threads": torch.get_num_threads(),
                             "sample": f"UNMODIFIED reference (oracle/_ref, {open(os.path.join(make_ref.OUT, 'HOW')).read().strip()}) "
                                       f"maximize() on torch CPU, first {n_s} of {n} entities ({e_s} nnz), batching={args.ref_batching}: "
                                       f"{it_s_sample:.3f} it/s on the sample = {e_s * it_s_sample / 1e6:.1f} M nnz/s, scaled by nnz to "
                                       f"the full problem" + ("; note: the reference corrupts mixed projection maps "
                                       "(utils/sparse_utils.py:177,220), its timing is unaffected" if args.mixed else ""),
                             "nnz_per_s": e_s * it_s_sample, "objective_build_s": t_obj, "port": port},
            "e2e": {"value": it_s_full, "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        })
    elif port is not None:
        config.update({"nnz": port["nnz_full"], "sample_entities": port["sample_entities"]})
        line.update({
            "value": port["value"], "ms_per_step": 1e3 / port["value"],
            "cpu_baseline": {"value": port["value"], "unit": "iterations/s", "cores": threads, "kind": "port",
                             "sample": port["sample"] + " (oracle/_ref absent: the reference itself was not available)"},
            "e2e": {"value": port["value"], "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        })
    else:
        line = {"impl": "reference", "unavailable": "neither oracle/_ref nor the C port could be run"}
    print(json.dumps(line), flush=True)


def _time_port(args, threads, n_s, iters=6, warm=2):
    """oracle/matching_oracle.c (OpenMP) driving the numpy AGD restatement: iterations/s scaled to the full problem."""
    import numpy as np
    import torch

    from benchmark.synthetic import capacity_vector, generate_shard
    from oracle import c_oracle
    from oracle import dualip_oracle as O

    n, m = args.entities, args.duals
    gen_device = "cuda" if torch.cuda.is_available() else "cpu"
    shard = generate_shard(n, m, args.sparsity, SEED, gen_device, 0, n_s)
    b = capacity_vector(shard.greedy_load * (n / n_s), m, args.sparsity, SEED, gen_device).cpu().numpy()
    ccol, row = shard.ccol.cpu().numpy(), shard.row.cpu().numpy()
    a, c = shard.a.cpu().numpy(), shard.c.cpu().numpy()
    if args.jacobi:
        a, b, _ = O.jacobi_precondition(a, row, b, m)
        a = a.astype(np.float32)
        b = b.astype(np.float32)
    e_s = row.size
    e_full = e_s * (n / n_s)
    if args.mixed:
        classes = [c_oracle.make_class("simplex", {"z": 1.0}), c_oracle.make_class("box", {"lower": 0.0, "upper": 1.0})]
        col_class = (np.arange(n_s) % 2).astype(np.uint8)
    else:
        classes, col_class = [c_oracle.make_class("simplex", {"z": 1.0})], None
    state = {"i": 0, "t0": None}

    def timed_calc(lam, gamma):
        if state["i"] == warm:
            state["t0"] = time.perf_counter()
        state["i"] += 1
        r = c_oracle.calculate(ccol, row, a, c, m, classes, lam, gamma, b, col_class, want_x=False, want_diag=False, threads=threads)
        return r["grad"], r["scal"][0]

    O.agd_maximize(timed_calc, np.zeros(m, dtype=np.float32), warm + iters, GAMMA, INITIAL_STEP, MAX_STEP)
    it_s_sample = iters / (time.perf_counter() - state["t0"])
    return {"value": it_s_sample * (e_s / e_full), "unit": "iterations/s", "cores": threads, "kind": "port",
            "nnz_full": int(round(e_full)), "sample_entities": n_s,
            "sample": f"C/OpenMP restatement (oracle/matching_oracle.c), first {n_s} entities ({e_s} nnz), {iters} iterations: "
                      f"{e_s * it_s_sample / 1e6:.1f} M nnz/s, scaled by nnz"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", choices=["native", "reference"], default="native")
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default="c3")
    ap.add_argument("--entities", type=int, default=None)
    ap.add_argument("--duals", type=int, default=None)
    ap.add_argument("--sparsity", type=float, default=None)
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = same K as the device loop")
    ap.add_argument("--ncu-traffic-bytes", type=float, default=None,
                    help="dram__bytes_read.sum + dram__bytes_write.sum per launch from an ncu --set full capture of this workload")
    ap.add_argument("--cpu-sample-cols", type=int, default=4_000_000)
    ap.add_argument("--ref-sample-cols", type=int, default=1_000_000,
                    help="--impl reference: entities of the workload the unmodified reference is timed on")
    ap.add_argument("--ref-batching", action="store_true",
                    help="--impl reference: batching=True (the reference benchmark's default is False, config.py:22)")
    ap.add_argument("--warm-start-iters", type=int, default=200,
                    help="untimed iterations from zero that produce the dual the measured run starts from (0: start at zero)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--exchange", choices=["peer", "nccl"], default="peer",
                    help="sharded runs: read the partial sums from peer memory inside the update kernel (default), or NCCL all_reduce")
    ap.add_argument("--graph", choices=["auto", "on", "off"], default="auto",
                    help="timed region as CUDA-graph replays of the one-launch iteration (what maximize() does once the plan has "
                         "settled). auto: on for workloads whose inputs fit in L2 (launch-latency-bound, e.g. c2), off otherwise so "
                         "that CUDA events can bracket every kernel of the timed region")
    ap.add_argument("--kernel-series", action="store_true", help="add the per-iteration kernel times (ms) to the JSON line")
    args = ap.parse_args()
    n, m, sp, mixed, jac = WORKLOADS[args.workload]
    args.entities = args.entities or n
    args.duals = args.duals or m
    args.sparsity = args.sparsity or sp
    args.mixed, args.jacobi = mixed, jac
    args.warmup = max(args.warmup, 3)
    if args.workload in ("c1", "c5"):
        # single-GPU side configurations (BASELINE.json configs[0] and configs[4]); their line carries the reference's CPU run
        if int(os.environ.get("WORLD_SIZE", "1")) > 1 or args.impl == "reference":
            raise SystemExit("--workload c1 / c5 run on one GPU with the native arm (cpu_baseline holds the reference's CPU run)")
        import sys as _sys

        from benchmark import extra_workloads as X

        (X.run_c1 if args.workload == "c1" else X.run_c5)(args, _sys.modules[__name__])
        return
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
