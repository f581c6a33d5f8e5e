"""Experiment (not product code): is the CTA-level imbalance of the slab kernel systematic?  Per-CTA main-loop durations of
several launches of the same plan (DUALIP_TIMELINE=1) and their correlation across launches."""
import ctypes, os, sys
os.environ["DUALIP_TIMELINE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from benchmark.synthetic import generate_shard, capacity_vector
from dualip_b200 import _native
from dualip_b200.objectives.matching import MatchingInputArgs, MatchingSolverDualObjectiveFunction
from dualip_b200.preprocessing.precondition import jacobi_precondition
import bench
dev = torch.device('cuda:0')
n, m, sp = int(os.environ.get('N', 10_000_000)), 10_000, 1e-3
sh = generate_shard(n, m, sp, 42, dev); b = capacity_vector(sh.greedy_load, m, sp, 42, dev)
A = torch.sparse_csc_tensor(sh.ccol, sh.row, sh.a, size=(m, n)); C = torch.sparse_csc_tensor(sh.ccol, sh.row, sh.c, size=(m, n))
jacobi_precondition(A, b)
root = os.path.dirname(os.path.abspath(__file__))
lam = torch.from_numpy(np.load(os.path.join(root, 'lams_c3small.npz'))['lam100']).to(dev)
obj = MatchingSolverDualObjectiveFunction(MatchingInputArgs(A, C, bench.mixed_projection_map(n, 0, dev), b), gamma=1e-3)
grad = torch.empty(m, device=dev); scal = torch.zeros(8, dtype=torch.float64, device=dev)
nc = obj.plan_info()["n_ctas"]
fn = _native.lib().dualip_debug_timeline; fn.restype = ctypes.c_int; fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
runs = []
for r in range(6):
    obj.launch_calc(lam.data_ptr(), 1e-3, grad.data_ptr(), scal.data_ptr())
    torch.cuda.synchronize()
    buf = (ctypes.c_uint64 * (10 * nc))()
    assert fn(obj._plan, buf, nc) == 0
    t = np.frombuffer(buf, dtype=np.uint64).reshape(nc, 5, 2).astype(np.int64)
    g = t[:, :, 1]
    runs.append((g[:, 2] - g[:, 1]) / 1e3)  # main loop, microseconds (globaltimer)
R = np.array(runs[2:])  # skip warm-up launches
print("main loop us per CTA: mean %.1f  min %.1f  max %.1f  (per launch max-mean: %s)" % (R.mean(), R.min(), R.max(), np.round(R.max(1) - R.mean(1), 1)))
cc = np.corrcoef(R)
print("correlation of per-CTA durations between launches:", np.round(cc[np.triu_indices(len(R), 1)], 2))
avg = R.mean(0)
order = np.argsort(-avg)
print("slowest CTAs (blockIdx: mean us):", [(int(i), round(float(avg[i]), 1)) for i in order[:12]])
print("fastest CTAs:", [(int(i), round(float(avg[i]), 1)) for i in order[-8:]])
print("std across CTAs of the mean duration %.2f us; mean within-CTA std across launches %.2f us" % (avg.std(), R.std(0).mean()))
