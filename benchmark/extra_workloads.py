"""bench.py --workload c1 / c5: the two BASELINE.json configurations that are not synthetic matching LPs of the benchmark
generator.  Same JSON line as the main arm (metric, value, e2e, roofline, cpu_baseline, clocks); single GPU.

c1  configs[0]: MovieLens-shaped matching (examples/movielens_matching): 138,493 users x 26,744 movies, ~20M ratings,
    a == 1, c = -rating, simplex z = 1 per user (the example's map), gamma = 0.1.  Every column is longer than the register
    path handles and m needs more shared memory than lambda + accumulator fit in: this is the generic-path / long-column
    workload.
c5  configs[4]: the shipped MIPLIB-2017 instance v150d30-2hopcds (7822 x 150, 103,991 nnz; tests/golden/lp_miplib.npz holds
    the matrices the reference's MPS parser produced) through MIPLIB2017ObjectiveFunction with a step gamma schedule.
"""
from __future__ import annotations

import json
import os
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _clocks(sampler_cls, idx, fn):
    s = sampler_cls(idx)
    s.start()
    out = fn()
    return out, s.stop()


def run_c1(args, B):
    import torch

    from benchmark.synthetic import generate_movielens_shaped
    from dualip_b200.objectives.matching import MatchingInputArgs, MatchingSolverDualObjectiveFunction
    from dualip_b200.optimizers.agd import AcceleratedGradientDescent, FusedAscentLoop, no_iteration_callback
    from dualip_b200.projections import create_projection_map

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    n, m = args.entities or 138_493, args.duals or 26_744
    gamma, K, W = 0.1, args.steps, args.warmup
    t0 = time.time()
    shard, b = generate_movielens_shaped(n, m, B.SEED, dev)
    A = torch.sparse_csc_tensor(shard.ccol, shard.row, shard.a, size=(m, n))
    C = torch.sparse_csc_tensor(shard.ccol, shard.row, shard.c, size=(m, n))
    proj = os.environ.get("DUALIP_C1_PROJ", "simplex")  # the example's own map is simplex z=1 (:163); BASELINE.json words it as box
    pm = create_projection_map("simplex", {"z": 1}, n) if proj == "simplex" else create_projection_map("box", {"lower": 0.0, "upper": 1.0}, n)
    torch.cuda.synchronize()
    t_gen = time.time() - t0
    t0 = time.time()
    obj = MatchingSolverDualObjectiveFunction(MatchingInputArgs(A, C, pm, b), gamma=gamma)
    torch.cuda.synchronize()
    t_plan = time.time() - t0
    info = obj.plan_info()
    lens = (shard.ccol[1:] - shard.ccol[:-1])
    B.log(f"[bench] c1 ready: nnz {info['nnz']}, column lengths min {int(lens.min())} mean {float(lens.float().mean()):.1f} max {int(lens.max())}, plan {info}")
    kw = dict(gamma=gamma, initial_step_size=B.INITIAL_STEP, max_step_size=B.MAX_STEP, iteration_callback=no_iteration_callback)
    lam0 = torch.zeros(m, device=dev)
    if args.warm_start_iters > 0:
        lam0 = AcceleratedGradientDescent(max_iter=args.warm_start_iters, **kw).maximize(obj, lam0).dual_val.clone()
    solver = AcceleratedGradientDescent(max_iter=W + K, **kw)
    loop = FusedAscentLoop(solver, obj, lam0)
    for i in range(1, W + 1):
        loop.step(i)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    loop.kernel_events, loop.kernel_events_base = kev, W + 1
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = B.ClockSampler(0)
    torch.cuda.synchronize()
    sampler.start()
    ev0.record()
    for i in range(W + 1, W + K + 1):
        loop.step(i)
    ev1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    ms_total = ev0.elapsed_time(ev1)
    result = loop.finish()
    loop.close()
    kt = [a.elapsed_time(e) for a, e in kev]
    kernel_ms = sum(kt) / len(kt)
    b_alg = obj.algorithmic_bytes()
    peak, peak_src = B.measured_peak_gbs()
    # end to end with host buffers: wall clock of a W+K-iteration maximize() minus a W-iteration one
    lam_host = lam0.cpu().pin_memory()

    def host_run(iters):
        hs = AcceleratedGradientDescent(max_iter=iters, **kw)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        hs.maximize(obj, lam_host)
        torch.cuda.synchronize()
        return time.perf_counter() - t0

    host_run(W)
    dt = max(host_run(W + K) - host_run(W), 1e-9)
    h2d, d2h = obj.host_io_bytes()
    # the unmodified reference on the same problem, CPU
    cpu = None
    if not args.no_cpu:
        cpu = _reference_cpu_matching(shard, b, m, n, gamma, B, batching=True, proj=proj)
    it_s = K / (ms_total * 1e-3)
    line = {
        "metric": "dual-ascent iterations/sec", "value": it_s, "unit": "iterations/s", "n_gpus": 1, "steps": K, "warmup": W,
        "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": f"MovieLens-shaped matching LP (configs[0]): {n} users x {m} movies, a=1, c=-rating, {'simplex z=1' if proj == 'simplex' else 'box [0,1]'}, "
                               f"gamma={gamma}; ratings drawn (ml-20m is not available offline)",
                   "workload_id": "c1", "entities": n, "duals": m, "nnz": info["nnz"], "parallelism": "single GPU",
                   "column_lengths": {"min": int(lens.min()), "mean": float(lens.float().mean()), "max": int(lens.max())},
                   "l2": "plan (%.0f MB) is about the size of the 126 MB L2: partly L2-resident" % (info["owned_bytes"] / 1e6),
                   "warm_start": f"lambda after {args.warm_start_iters} untimed iterations from zero"},
        "nnz_per_s": info["nnz"] * it_s, "gpu_launches": (info["launches_per_calc"] + (0 if loop.one_launch else 1)) * K,
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": b_alg / (kernel_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": b_alg / (kernel_ms * 1e-3) / 1e9 / peak, "traffic": None, "kernel": "matching_slab_kernel (generic path)",
                     "kernel_ms": kernel_ms, "kernel_ms_min": min(kt), "kernel_ms_max": max(kt), "algorithmic_bytes": b_alg,
                     "peak_source": peak_src},
        "final_dual_objective": result.dual_objective,
        "setup": {"generate_s": round(t_gen, 2), "plan_s": round(t_plan, 2), "plan": info},
        "e2e": {"value": K / dt, "unit": "iterations/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": K,
                "path": "AcceleratedGradientDescent.maximize with a pinned host dual vector"},
    }
    if cpu is not None:
        line["cpu_baseline"] = cpu
    print(json.dumps(line), flush=True)


def _reference_cpu_matching(shard, b, m, n, gamma, B, batching, steps=4, warm=1, proj="simplex"):
    import torch

    from benchmark import reference_arm as R
    from oracle import make_ref

    if not make_ref.available():
        return None
    threads = R.host_threads()
    torch.set_num_threads(threads)
    make_ref.import_reference()
    from dualip.objectives.matching import MatchingInputArgs, MatchingSolverDualObjectiveFunction
    from dualip.optimizers.agd import AcceleratedGradientDescent
    from dualip.projections.base import create_projection_map

    ccol, row = shard.ccol.cpu(), shard.row.cpu()
    A = torch.sparse_csc_tensor(ccol, row, shard.a.cpu(), size=(m, n))
    C = torch.sparse_csc_tensor(ccol, row, shard.c.cpu(), size=(m, n))
    pm = create_projection_map("simplex", {"z": 1}, n) if proj == "simplex" else create_projection_map("box", {"lower": 0.0, "upper": 1.0}, n)
    args = MatchingInputArgs(A=A, c=C, projection_map=pm, b_vec=b.cpu(), equality_mask=None)
    objective = MatchingSolverDualObjectiveFunction(matching_input_args=args, gamma=gamma, batching=batching)
    marks = {}

    def cb(i, r):
        if i == warm:
            marks["t0"] = time.perf_counter()

    solver = AcceleratedGradientDescent(max_iter=warm + steps, gamma=gamma, initial_step_size=B.INITIAL_STEP, max_step_size=B.MAX_STEP,
                                        iteration_callback=cb)
    solver.maximize(objective, torch.zeros(m))
    dt = time.perf_counter() - marks["t0"]
    return {"value": steps / dt, "unit": "iterations/s", "cores": threads, "kind": "reference",
            "sample": f"UNMODIFIED reference (oracle/_ref) maximize() on torch CPU ({threads} threads), the whole problem "
                      f"({int(row.numel())} nnz), {steps} iterations after {warm} warm-up, batching={batching}",
            "nnz_per_s": int(row.numel()) * steps / dt}


def run_c5(args, B):
    import numpy as np
    import torch

    from dualip_b200.objectives.miplib import MIPLIB2017ObjectiveFunction, MIPLIBInputArgs
    from dualip_b200.optimizers.agd import AcceleratedGradientDescent, FusedAscentLoop, no_iteration_callback
    from dualip_b200.projections import create_projection_map

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    d = np.load(os.path.join(ROOT, "tests", "golden", "lp_miplib.npz"))
    m, n = d["A"].shape
    K, W = args.steps, args.warmup
    gamma = 1e-3
    decay = {"decay_steps": max(35, (W + K) // 12), "decay_factor": 0.7}  # benchmark values 35 / 0.7, stretched to the run length

    def projection_map(npz):
        pm, groups = {}, {}
        for j in range(npz["lower"].size):
            groups.setdefault((float(npz["lower"][j]), float(npz["upper"][j])), []).append(j)
        for k, ((lo, hi), idx) in enumerate(groups.items()):
            if np.isinf(lo) and np.isinf(hi):
                continue
            if np.isinf(hi):
                pm.update(create_projection_map("cone", {"lower": lo}, n, indices=idx, key_prefix=f"g{k}_"))
            elif np.isinf(lo):
                pm.update(create_projection_map("cone", {"upper": hi}, n, indices=idx, key_prefix=f"g{k}_"))
            else:
                pm.update(create_projection_map("box", {"lower": lo, "upper": hi}, n, indices=idx, key_prefix=f"g{k}_"))
        return pm

    A_host = torch.from_numpy(d["A"]).to_sparse()
    host_args = MIPLIBInputArgs(A=A_host, c=torch.from_numpy(d["c"]), projection_map=projection_map(d), b_vec=torch.from_numpy(d["b"]))
    obj = MIPLIB2017ObjectiveFunction(MIPLIBInputArgs(A=A_host.to(dev), c=host_args.c.to(dev), projection_map=host_args.projection_map,
                                                      b_vec=host_args.b_vec.to(dev)))
    kw = dict(gamma=gamma, initial_step_size=1e-5, max_step_size=0.1, gamma_decay_type="step", gamma_decay_params=decay,
              iteration_callback=no_iteration_callback)
    solver = AcceleratedGradientDescent(max_iter=W + K, **kw)
    loop = FusedAscentLoop(solver, obj, torch.zeros(m, device=dev))
    for i in range(1, W + 1):
        loop.step(i)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = B.ClockSampler(0)
    torch.cuda.synchronize()
    sampler.start()
    ev0.record()
    for i in range(W + 1, W + K + 1):
        loop.step(i)
    ev1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    ms_total = ev0.elapsed_time(ev1)
    result = loop.finish()
    loop.close()
    # end to end: host tensors in, run through the public entry, result read back (problem upload included)
    from dualip_b200.run_solver import build_objective, transfer_tensors_to_device
    from dualip_b200.types import ComputeArgs, ObjectiveArgs, SolverArgs

    t0 = time.perf_counter()
    on_dev = transfer_tensors_to_device(host_args, "cuda:0")
    f = build_objective(on_dev, SolverArgs(gamma=gamma), ComputeArgs(host_device="cuda:0"), ObjectiveArgs(objective_type="miplib2017"))
    res = AcceleratedGradientDescent(max_iter=W + K, **kw).maximize(f, torch.zeros(m, device=dev))
    dual_host = res.dual_val.cpu()
    dt = time.perf_counter() - t0
    h2d = int(A_host._nnz() * 20 + 4 * (m + n))
    nnz = int(A_host._nnz())
    b_alg = 2 * nnz * 8 + 4 * (3 * m + 3 * n)  # A read twice (CSR for Ax, CSC for A^T lambda): value + index per entry
    peak, peak_src = B.measured_peak_gbs()
    it_s = K / (ms_total * 1e-3)
    cpu = None
    if not args.no_cpu:
        cpu = _reference_cpu_lp(d, projection_map, gamma, decay, B)
    line = {
        "metric": "dual-ascent iterations/sec", "value": it_s, "unit": "iterations/s", "n_gpus": 1, "steps": K, "warmup": W,
        "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "MIPLIB-2017 v150d30-2hopcds (matrices parsed by the reference's MPS reader, tests/golden/lp_miplib.npz)",
        "config": {"workload": f"MIPLIB-2017 example LP (configs[4]) {m} constraints x {n} variables, {nnz} nnz, box bounds, "
                               f"gamma={gamma} with step decay {decay}", "workload_id": "c5", "parallelism": "single GPU",
                   "l2": "the whole problem (1.7 MB) lives in L2: latency-bound, not a roofline claim"},
        "gpu_launches": 4 * K, "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": b_alg / (ms_total / K * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": b_alg / (ms_total / K * 1e-3) / 1e9 / peak, "traffic": None,
                     "kernel": "lp_primal_kernel + lp_dual_kernel + lp_finalize_kernel + agd_step_kernel (whole step)",
                     "kernel_ms": ms_total / K, "algorithmic_bytes": b_alg, "peak_source": peak_src,
                     "note": "four dependent launches of microsecond kernels: launch latency, not bandwidth, bounds the step"},
        "final_dual_objective": result.dual_objective,
        "e2e": {"value": (W + K) / dt, "unit": "iterations/s", "h2d_bytes_per_step": h2d / (W + K), "d2h_bytes_per_step": 4 * m / (W + K),
                "steps": W + K, "path": "host tensors -> transfer_tensors_to_device -> build_objective -> maximize -> dual back on the "
                "host; problem upload and plan construction inside the timed region", "final_dual_objective": res.dual_objective,
                "dual_on_host_norm": float(dual_host.norm())},
    }
    if cpu is not None:
        line["cpu_baseline"] = cpu
    print(json.dumps(line), flush=True)


def _reference_cpu_lp(d, projection_map, gamma, decay, B, steps=300, warm=20):
    import torch

    from benchmark import reference_arm as R
    from oracle import make_ref

    if not make_ref.available():
        return None
    threads = R.host_threads()
    torch.set_num_threads(threads)
    make_ref.import_reference()
    from dualip.objectives.miplib import MIPLIB2017ObjectiveFunction, MIPLIBInputArgs
    from dualip.optimizers.agd import AcceleratedGradientDescent

    m = d["A"].shape[0]
    # the reference's own key names for the bounds (its objective reads lower/upper from box entries)
    args = MIPLIBInputArgs(A=torch.from_numpy(d["A"]).to_sparse(), c=torch.from_numpy(d["c"]), projection_map=projection_map(d),
                           b_vec=torch.from_numpy(d["b"]), equality_mask=None)
    objective = MIPLIB2017ObjectiveFunction(miplib_input_args=args)
    marks = {}

    def cb(i, r):
        if i == warm:
            marks["t0"] = time.perf_counter()

    solver = AcceleratedGradientDescent(max_iter=warm + steps, gamma=gamma, initial_step_size=1e-5, max_step_size=0.1,
                                        gamma_decay_type="step", gamma_decay_params=decay, iteration_callback=cb)
    solver.maximize(objective, torch.zeros(m))
    dt = time.perf_counter() - marks["t0"]
    return {"value": steps / dt, "unit": "iterations/s", "cores": threads, "kind": "reference",
            "sample": f"UNMODIFIED reference (oracle/_ref) MIPLIB2017ObjectiveFunction + maximize() on torch CPU ({threads} threads), "
                      f"{steps} iterations after {warm} warm-up"}
