"""Experiment: per-CTA phase timeline of one launch (DUALIP_TIMELINE=1).  Not product code."""
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
for _ in range(4): obj.launch_calc(lam.data_ptr(), 1e-3, grad.data_ptr(), scal.data_ptr())
torch.cuda.synchronize()
nc = obj.plan_info()["n_ctas"]
buf = (ctypes.c_uint64 * (10 * nc))()
fn = _native.lib().dualip_debug_timeline; fn.restype = ctypes.c_int; fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
assert fn(obj._plan, buf, nc) == 0
t = np.frombuffer(buf, dtype=np.uint64).reshape(nc, 5, 2).astype(np.int64)
clk, g = t[:, :, 0], t[:, :, 1]
g0 = g[:, 0].min()
names = ["prologue", "main loop", "flush", "ticket..end"]
for i, nm in enumerate(names):
    d = (clk[:, i + 1] - clk[:, i]); d = d[clk[:, i + 1] > 0]
    print(f"{nm:12s} cycles: mean {d.mean():9.0f} min {d.min():9.0f} max {d.max():9.0f}  (n={d.size})")
print("globaltimer ns rel. to first CTA start: start max %d | main-loop end min %d mean %d max %d | flush end max %d | kernel end %d" % (
    (g[:, 0] - g0).max(), (g[:, 2] - g0).min(), (g[:, 2] - g0).mean(), (g[:, 2] - g0).max(), (g[:, 3] - g0).max(), (g[:, 4].max() - g0)))

big = (ctypes.c_uint64 * (12 * 4096))()
assert fn(obj._plan, big, -1) == 0
arr = np.frombuffer(big, dtype=np.uint64).astype(np.int64)
tr = arr[10 * 4096: 10 * 4096 + 1024]; hy = arr[11 * 4096: 11 * 4096 + 1024]
ph = arr[10 * 4096 + 1024: 10 * 4096 + 1024 + 4096].reshape(1024, 4)
nz = tr > 0
tr, hy, ph = tr[nz], hy[nz], ph[nz]
tot = np.diff(tr); dd = (hy & 0xffff)[:-1]; cls = ((hy >> 16) & 0xff)[:-1]
wait = (ph[:, 0] - tr)[:-1]; proj = (ph[:, 1] - ph[:, 0])[:-1]; emit = tr[1:] - ph[:-1, 1]
print("slab trace (warp 0 of CTA 0): d:total(wait/project/emit) x count")
for c in np.unique(cls):
    line = f"class {c}: "
    for dv in np.unique(dd[cls == c]):
        sel = (cls == c) & (dd == dv)
        line += f"d{dv}:{int(tot[sel].mean())}({int(wait[sel].mean())}/{int(proj[sel].mean())}/{int(emit[sel].mean())})x{sel.sum()} "
    print(line)
