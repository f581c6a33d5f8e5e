"""A/B harness (not product code): the kernel on shard 0 of an 8-way split of C3 at the dual the FULL problem reaches after
PRE iterations (what rank 0 of an 8-GPU run computes)."""
import os, sys, json, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from benchmark.synthetic import generate_shard, capacity_vector
from dualip_b200 import _native
from dualip_b200.objectives.matching import MatchingInputArgs, MatchingSolverDualObjectiveFunction
from dualip_b200.optimizers.agd import AcceleratedGradientDescent, no_iteration_callback
from dualip_b200.preprocessing.precondition import jacobi_precondition
dev = torch.device("cuda:0")
n_total, m, sp = 100_000_000, 10_000, 1e-3
world = int(os.environ.get("WORLD", 8))
sh = generate_shard(n_total, m, sp, 42, dev)
b = capacity_vector(sh.greedy_load, m, sp, 42, dev)
A = torch.sparse_csc_tensor(sh.ccol, sh.row, sh.a, size=(m, n_total)); C = torch.sparse_csc_tensor(sh.ccol, sh.row, sh.c, size=(m, n_total))
jacobi_precondition(A, b)
obj = MatchingSolverDualObjectiveFunction(MatchingInputArgs(A, C, bench.mixed_projection_map(n_total, 0, dev), b), gamma=1e-3)
pre = int(os.environ.get("PRE", 200))
solver = AcceleratedGradientDescent(max_iter=pre, gamma=1e-3, initial_step_size=1e-3, max_step_size=1e-1, iteration_callback=no_iteration_callback)
lam = solver.maximize(obj, torch.zeros(m, device=dev)).dual_val.clone()
del obj
n = n_total // world
e1 = int(sh.ccol[n].item())
A2 = torch.sparse_csc_tensor(sh.ccol[: n + 1].clone(), sh.row[:e1].clone(), A.values()[:e1].clone(), size=(m, n))
C2 = torch.sparse_csc_tensor(sh.ccol[: n + 1].clone(), sh.row[:e1].clone(), sh.c[:e1].clone(), size=(m, n))
del A, C, sh
torch.cuda.empty_cache()
o2 = MatchingSolverDualObjectiveFunction(MatchingInputArgs(A2, C2, bench.mixed_projection_map(n, 0, dev), None), gamma=1e-3)
part = torch.empty(m + 2, device=dev)
reps = int(os.environ.get("REPS", 20))
for _ in range(int(os.environ.get('WARM', 140))):
    o2.launch_partial(lam.data_ptr(), 1e-3, part.data_ptr()); o2.launched()
torch.cuda.synchronize()
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
for a, e in ev:
    a.record(); o2.launch_partial(lam.data_ptr(), 1e-3, part.data_ptr()); e.record()
torch.cuda.synchronize()
ts = sorted(a.elapsed_time(e) for a, e in ev)
ms = sum(ts) / len(ts)
print(json.dumps({"shard": f"0/{world}", "ms_mean": round(ms, 4), "ms_min": round(ts[0], 4), "frac": round(o2.algorithmic_bytes() / ms / 1e6 / 6552.6, 4),
                  "lib": os.path.basename(os.environ.get("DUALIP_B200_LIB", "default")), "stage": o2.plan_info()["staged_degree"]}), flush=True)
if os.environ.get("DUALIP_TIMELINE"):
    nc = o2.plan_info()["n_ctas"]
    buf = (ctypes.c_uint64 * (10 * nc))()
    fn = _native.lib().dualip_debug_timeline; fn.restype = ctypes.c_int; fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
    assert fn(o2._plan, buf, nc) == 0
    g = np.frombuffer(buf, dtype=np.uint64).reshape(nc, 5, 2).astype(np.int64)[:, :, 1]
    rel = g - g[:, 0].min()
    for i, nm in enumerate(["start", "lambda staged", "main loop end", "flush end", "cta end"]):
        v = rel[:, i][g[:, i] > 0]
        print(f"  {nm:14s} ns: min {v.min():8d} mean {int(v.mean()):8d} max {v.max():8d} (n={v.size})")
    if os.environ.get("DUMP_LAYOUT"):
        fn2 = _native.lib().dualip_debug_layout; fn2.restype = ctypes.c_int
        fn2.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
        gbuf = (ctypes.c_int64 * (6 * 4096))(); rbuf = (ctypes.c_int64 * (nc + 1))()
        G = fn2(o2._plan, gbuf, 4096, rbuf, nc + 1)
        json.dump({"groups": np.frombuffer(gbuf, dtype=np.int64)[: 6 * G].reshape(G, 6).tolist(), "ranges": list(rbuf),
                   "main_ns": (rel[:, 2] - rel[:, 1]).tolist()}, open(os.environ["DUMP_LAYOUT"], "w"))
