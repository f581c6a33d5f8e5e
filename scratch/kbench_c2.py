"""Kernel timeline at C2 (1M entities x 1k duals, simplex, no Jacobi): where do the 67 us of a 10M-nnz launch go?

    DUALIP_TIMELINE=1 python scratch/kbench_c2.py          env: N (entities), M (duals), SP (sparsity), PRE, REPS
"""
import os, sys, json, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from benchmark.synthetic import generate_shard, capacity_vector
from dualip_b200 import _native
from dualip_b200.objectives.matching import MatchingInputArgs, MatchingSolverDualObjectiveFunction
from dualip_b200.optimizers.agd import AcceleratedGradientDescent, no_iteration_callback
from dualip_b200.projections import create_projection_map

dev = torch.device("cuda:0")
n, m, sp = int(os.environ.get("N", 1_000_000)), int(os.environ.get("M", 1000)), float(os.environ.get("SP", 1e-2))
sh = generate_shard(n, m, sp, 42, dev, 0, n)
b = capacity_vector(sh.greedy_load, m, sp, 42, dev)
A = torch.sparse_csc_tensor(sh.ccol, sh.row, sh.a, size=(m, n)); C = torch.sparse_csc_tensor(sh.ccol, sh.row, sh.c, size=(m, n))
obj = MatchingSolverDualObjectiveFunction(MatchingInputArgs(A, C, create_projection_map("simplex", {"z": 1.0}, n), b), gamma=1e-3)
pre = int(os.environ.get("PRE", 200))
solver = AcceleratedGradientDescent(max_iter=pre, gamma=1e-3, initial_step_size=1e-3, max_step_size=1e-1, iteration_callback=no_iteration_callback)
lam = solver.maximize(obj, torch.zeros(m, device=dev)).dual_val.clone()
grad = torch.empty(m, device=dev); scal = torch.zeros(8, dtype=torch.float64, device=dev)
reps = int(os.environ.get("REPS", 50))
for _ in range(5): obj.launch_calc(lam.data_ptr(), 1e-3, grad.data_ptr(), scal.data_ptr())
torch.cuda.synchronize()
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
for a, e in ev:
    a.record(); obj.launch_calc(lam.data_ptr(), 1e-3, grad.data_ptr(), scal.data_ptr()); e.record()
torch.cuda.synchronize()
ts = sorted(a.elapsed_time(e) for a, e in ev)
info = obj.plan_info()
# back-to-back launches without events in between (launch overhead hidden)
t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
t0.record()
for _ in range(200): obj.launch_calc(lam.data_ptr(), 1e-3, grad.data_ptr(), scal.data_ptr())
t1.record(); torch.cuda.synchronize()
print(json.dumps({"n": n, "m": m, "nnz": info["nnz"], "ms_min": round(ts[0], 4), "ms_med": round(ts[len(ts)//2], 4),
                  "back_to_back_ms": round(t0.elapsed_time(t1) / 200, 4), "fixed": info["fixed_point"], "row_scaled": info["row_scaled"],
                  "ctas": info["n_ctas"], "obj": float(scal[0])}), flush=True)
if os.environ.get("DUALIP_TIMELINE"):
    nc = info["n_ctas"]
    buf = (ctypes.c_uint64 * (10 * nc))()
    fn = _native.lib().dualip_debug_timeline; fn.restype = ctypes.c_int; fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
    assert fn(obj._plan, buf, nc) == 0
    tl = np.frombuffer(buf, dtype=np.uint64).reshape(nc, 5, 2).astype(np.int64)
    g = tl[:, :, 1]
    rel = g - g[:, 0].min()
    names = ["start", "lambda staged", "main loop end", "flush end", "cta end"]
    for i, nm in enumerate(names):
        v = rel[:, i][g[:, i] > 0]
        print(f"  {nm:14s} ns after first CTA start: min {v.min():8d} mean {int(v.mean()):8d} max {v.max():8d} (n={v.size})")
    main = rel[:, 2] - rel[:, 1]
    print("  main loop ns: min %d mean %d max %d" % (main.min(), main.mean(), main.max()))
    order = np.argsort(rel[:, 2])
    print("  slowest CTAs (main loop end):", [(int(c), int(rel[c, 2])) for c in order[-6:]], "fastest:", [(int(c), int(rel[c, 2])) for c in order[:4]])
    # column-length histogram of the plan
    lib = _native.lib()
    fn2 = lib.dualip_debug_layout; fn2.restype = ctypes.c_int
    fn2.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
    gbuf = (ctypes.c_int64 * (6 * 4096))(); rbuf = (ctypes.c_int64 * (nc + 1))()
    G = fn2(obj._plan, gbuf, 4096, rbuf, nc + 1)
    groups = np.frombuffer(gbuf, dtype=np.int64)[: 6 * G].reshape(G, 6)
    print("  groups (d: slabs):", {int(r[2]): int(r[1]) for r in groups})
# the most popular rows: how concentrated is the scatter?
cnt = torch.bincount(sh.row, minlength=m).float()
top = torch.sort(cnt, descending=True).values
print("  row popularity: top1 %.3f%% top10 %.3f%% of nnz; max/mean %.1f" % (100 * top[0] / cnt.sum(), 100 * top[:10].sum() / cnt.sum(), top[0] / cnt.mean()))
# fused evaluation + step (one launch per iteration) from the same dual, 200 iterations back to back
from dualip_b200.optimizers.agd import FusedAscentLoop
for one in ("1", "0"):
    os.environ["DUALIP_ONE_LAUNCH"] = one
    os.environ["DUALIP_GRAPH"] = "0"
    s2 = AcceleratedGradientDescent(max_iter=400, gamma=1e-3, initial_step_size=1e-3, max_step_size=1e-1, iteration_callback=no_iteration_callback)
    loop = FusedAscentLoop(s2, obj, lam)
    for i in range(1, 101): loop.step(i)
    torch.cuda.synchronize()
    t0.record()
    for i in range(101, 401): loop.step(i)
    t1.record(); torch.cuda.synchronize()
    loop.finish(); loop.close()
    print("  fused loop one_launch=%s: %.2f us per iteration" % (one, 1000 * t0.elapsed_time(t1) / 300))
