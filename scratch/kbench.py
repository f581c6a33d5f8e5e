"""Kernel A/B harness (not product code): times dualip_matching_calc at a LATE dual on the C3 workload (or a shard of it).

    N=100000000 REPS=20 python scratch/kbench.py [tag]      env: DUALIP_B200_LIB, DUALIP_STAGE, DUALIP_CTAS, PRE=150
"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from benchmark.synthetic import generate_shard, capacity_vector
from dualip_b200.objectives.matching import MatchingInputArgs, MatchingSolverDualObjectiveFunction
from dualip_b200.optimizers.agd import AcceleratedGradientDescent, no_iteration_callback
from dualip_b200.preprocessing.precondition import jacobi_precondition
from dualip_b200.projections import create_projection_map

dev = torch.device("cuda:0")
n_total, m, sp = 100_000_000, 10_000, 1e-3
n = int(os.environ.get("N", n_total))
kind = os.environ.get("KIND", "mixed")
tag = sys.argv[1] if len(sys.argv) > 1 else "run"
sh = generate_shard(n_total, m, sp, 42, dev, 0, n)
b = capacity_vector(sh.greedy_load * (n_total / n), m, sp, 42, dev)
A = torch.sparse_csc_tensor(sh.ccol, sh.row, sh.a, size=(m, n)); C = torch.sparse_csc_tensor(sh.ccol, sh.row, sh.c, size=(m, n))
jacobi_precondition(A, b)
pm = {"mixed": lambda: bench.mixed_projection_map(n, 0, dev), "simplex": lambda: create_projection_map("simplex", {"z": 1.0}, n),
      "box": lambda: create_projection_map("box", {"lower": 0.0, "upper": 1.0}, n)}[kind]()
obj = MatchingSolverDualObjectiveFunction(MatchingInputArgs(A, C, pm, b), gamma=1e-3)
del sh, A, C
pre = int(os.environ.get("PRE", 150))
solver = AcceleratedGradientDescent(max_iter=pre, gamma=1e-3, initial_step_size=1e-3, max_step_size=1e-1, iteration_callback=no_iteration_callback)
lam = solver.maximize(obj, torch.zeros(m, device=dev)).dual_val.clone()
grad = torch.empty(m, device=dev); scal = torch.zeros(8, dtype=torch.float64, device=dev)
reps = int(os.environ.get("REPS", 20))
for _ in range(3): obj.launch_calc(lam.data_ptr(), 1e-3, grad.data_ptr(), scal.data_ptr())
torch.cuda.synchronize()
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
for a, e in ev:
    a.record(); obj.launch_calc(lam.data_ptr(), 1e-3, grad.data_ptr(), scal.data_ptr()); e.record()
torch.cuda.synchronize()
ts = sorted(a.elapsed_time(e) for a, e in ev)
ms = sum(ts) / len(ts)
balg = obj.algorithmic_bytes()
info = obj.plan_info()
print(json.dumps({"tag": tag, "n": n, "kind": kind, "ms_mean": round(ms, 4), "ms_min": round(ts[0], 4), "ms_med": round(ts[len(ts)//2], 4),
                  "frac": round(balg / ms / 1e6 / 6552.6, 4), "stage": info["staged_degree"], "obj": float(scal[0]),
                  "lib": os.path.basename(os.environ.get("DUALIP_B200_LIB", "default"))}), flush=True)
if os.environ.get("DUALIP_TIMELINE"):
    import ctypes
    import numpy as np
    from dualip_b200 import _native
    nc = info["n_ctas"]
    buf = (ctypes.c_uint64 * (10 * nc))()
    fn = _native.lib().dualip_debug_timeline; fn.restype = ctypes.c_int; fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
    assert fn(obj._plan, buf, nc) == 0
    tl = np.frombuffer(buf, dtype=np.uint64).reshape(nc, 5, 2).astype(np.int64)
    g = tl[:, :, 1]
    g0 = g[:, 0].min()
    rel = g - g0
    names = ["start", "lambda staged", "main loop end", "flush end", "cta end"]
    for i, nm in enumerate(names):
        v = rel[:, i][g[:, i] > 0]
        print(f"  {nm:14s} ns after first CTA start: min {v.min():8d} mean {int(v.mean()):8d} max {v.max():8d} (n={v.size})")
    order = np.argsort(rel[:, 2])
    print("  slowest CTAs (main loop end):", [(int(c), int(rel[c, 2])) for c in order[-6:]], "fastest:", [(int(c), int(rel[c, 2])) for c in order[:4]])
if os.environ.get("DUMP_LAYOUT") and os.environ.get("DUALIP_TIMELINE"):
    lib = _native.lib()
    fn2 = lib.dualip_debug_layout; fn2.restype = ctypes.c_int
    fn2.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
    gbuf = (ctypes.c_int64 * (6 * 4096))(); rbuf = (ctypes.c_int64 * (nc + 1))()
    G = fn2(obj._plan, gbuf, 4096, rbuf, nc + 1)
    groups = np.frombuffer(gbuf, dtype=np.int64)[: 6 * G].reshape(G, 6).tolist()
    ranges = list(rbuf)
    kinds = {}
    json.dump({"n": n, "groups": groups, "ranges": ranges, "main_ns": (rel[:, 2] - rel[:, 1]).tolist()},
              open(os.environ["DUMP_LAYOUT"], "w"))
