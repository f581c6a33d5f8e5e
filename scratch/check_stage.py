"""GPU check (not product code): staged vs plain loads bit-identity for the installed library, at 2M entities."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from benchmark.synthetic import generate_shard, capacity_vector
from dualip_b200.objectives.matching import MatchingInputArgs, MatchingSolverDualObjectiveFunction
from dualip_b200.preprocessing.precondition import jacobi_precondition
import bench
dev = torch.device("cuda:0")
n, m, sp = 2_000_000, 10_000, 1e-3
sh = generate_shard(n, m, sp, 42, dev); b = capacity_vector(sh.greedy_load, m, sp, 42, dev)
A = torch.sparse_csc_tensor(sh.ccol, sh.row, sh.a, size=(m, n)); C = torch.sparse_csc_tensor(sh.ccol, sh.row, sh.c, size=(m, n))
jacobi_precondition(A, b)
lam = torch.rand(m, device=dev) * 40
def obj():
    return MatchingSolverDualObjectiveFunction(MatchingInputArgs(A, C, bench.mixed_projection_map(n, 0, dev), b), gamma=1e-3)
os.environ["DUALIP_STAGE"] = "0"
r0 = obj().calculate(lam, save_primal=True)
for st in ("20", "9", "3"):
    os.environ["DUALIP_STAGE"] = st
    o = obj(); info = o.plan_info()
    ok = True
    for _ in range(3):
        r1 = o.calculate(lam, save_primal=True)
        ok = ok and torch.equal(r1.primal_var, r0.primal_var) and torch.equal(r1.dual_gradient, r0.dual_gradient)
    print("DUALIP_STAGE", st, "staged_degree", info["staged_degree"], "smem", info["smem_bytes"], "bit-identical", ok)
