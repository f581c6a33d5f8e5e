"""GPU debug: cfg1 (MovieLens-shaped, simplex) -- x of the CUDA path vs the C oracle at the first iterates of the ascent."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import c_oracle, dualip_oracle as O
from dualip_b200.objectives.matching import MatchingInputArgs, MatchingSolverDualObjectiveFunction
from dualip_b200.projections import create_projection_map
d = np.load(os.path.join(ROOT, "tests/golden/cfg1_movielens_shaped.npz"))
n, m, gamma = d["ccol"].size - 1, int(d["n_rows"]), float(d["gamma"])
dev = "cuda:0"
ccol, row = torch.from_numpy(d["ccol"]), torch.from_numpy(d["row"])
A = torch.sparse_csc_tensor(ccol, row, torch.from_numpy(d["a"]), size=(m, n)).to(dev)
C = torch.sparse_csc_tensor(ccol, row, torch.from_numpy(d["c"]), size=(m, n)).to(dev)
obj = MatchingSolverDualObjectiveFunction(MatchingInputArgs(A, C, create_projection_map("simplex", {"z": 1.0}, n), torch.from_numpy(d["b"]).to(dev)), gamma=gamma)
cls = [c_oracle.make_class("simplex", {"z": 1.0})]
lens = np.diff(d["ccol"])
col_of = np.repeat(np.arange(n), lens)
lams = []
def calc(lam, g):
    lams.append(lam.copy())
    r = c_oracle.calculate(d["ccol"], d["row"], d["a"], d["c"], m, cls, lam, g, d["b"])
    return r["grad"], np.float32(r["scal"][0])
O.agd_maximize(calc, np.zeros(m, np.float32), 8, gamma, 1e-3, 1e-1)
for it, lam in enumerate(lams):
    ref = c_oracle.calculate(d["ccol"], d["row"], d["a"], d["c"], m, cls, lam, gamma, d["b"])
    r = obj.calculate(torch.from_numpy(lam).to(dev), save_primal=True, diagnostics=True)
    x = r.primal_var.cpu().numpy()
    bad = np.nonzero(x != ref["x"])[0]
    cols = np.unique(col_of[bad])
    print(f"iter {it}: obj gpu {float(r.scalars64[0]):.6f} ref {ref['scal'][0]:.6f}; differing entries {bad.size} in {cols.size} columns; "
          f"grad max abs diff {np.abs(r.dual_gradient.cpu().numpy() - ref['grad']).max():.3e} (max |g| {np.abs(ref['grad']).max():.3e})")
    diag = r.projection_diag.cpu().numpy()
    for j in cols[:6]:
        e0, e1 = d["ccol"][j], d["ccol"][j + 1]
        dj = diag[e0]
        print(f"   col {j} len {lens[j]} gpu branch {dj & 3} rho {dj >> 2} | oracle branch {ref['diag'][j] & 3} rho {ref['diag'][j] >> 2}"
              f" | sum gpu {x[e0:e1].sum():.7f} ref {ref['x'][e0:e1].sum():.7f} nnz gpu {(x[e0:e1] > 0).sum()} ref {(ref['x'][e0:e1] > 0).sum()} max diff {np.abs(x[e0:e1] - ref['x'][e0:e1]).max():.3e}")
