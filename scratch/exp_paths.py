"""Experiment: kernel time by projection class and iterate; branch histogram (not product code)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from benchmark.synthetic import generate_shard, capacity_vector
from dualip_b200.objectives.matching import MatchingInputArgs, MatchingSolverDualObjectiveFunction
from dualip_b200.optimizers.agd import AcceleratedGradientDescent, FusedAscentLoop
from dualip_b200.preprocessing.precondition import jacobi_precondition
from dualip_b200.projections import create_projection_map
import bench

dev = torch.device('cuda:0')
n, m, sp = int(os.environ.get('N', 10_000_000)), 10_000, 1e-3
sh = generate_shard(n, m, sp, 42, dev)
b = capacity_vector(sh.greedy_load, m, sp, 42, dev)
A = torch.sparse_csc_tensor(sh.ccol, sh.row, sh.a, size=(m, n)); C = torch.sparse_csc_tensor(sh.ccol, sh.row, sh.c, size=(m, n))
jacobi_precondition(A, b)
maps = {'mixed': bench.mixed_projection_map(n, 0, dev), 'simplex': create_projection_map('simplex', {'z': 1.0}, n), 'box': create_projection_map('box', {'lower': 0.0, 'upper': 1.0}, n)}
def timeit(obj, lam, reps=20):
    grad = torch.empty(m, device=dev); scal = torch.zeros(8, dtype=torch.float64, device=dev)
    for _ in range(3): obj.launch_calc(lam.data_ptr(), 1e-3, grad.data_ptr(), scal.data_ptr())
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): obj.launch_calc(lam.data_ptr(), 1e-3, grad.data_ptr(), scal.data_ptr())
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
lams = {}
objs = {k: MatchingSolverDualObjectiveFunction(MatchingInputArgs(A, C, pm, b), gamma=1e-3) for k, pm in maps.items()}
solver = AcceleratedGradientDescent(max_iter=200, gamma=1e-3, initial_step_size=1e-3, max_step_size=1e-1, iteration_callback=lambda i, r: None)
loop = FusedAscentLoop(solver, objs['mixed'], torch.zeros(m, device=dev))
lams[0] = torch.zeros(m, device=dev)
for i in range(1, 201):
    loop.step(i)
    if i in (10, 30, 100, 200): lams[i] = loop.current_dual().clone()
torch.cuda.synchronize()
balg = objs['mixed'].algorithmic_bytes()
for it, lam in lams.items():
    line = f'iter {it:3d} |lam|max {float(lam.max()):.3g}: '
    for k, o in objs.items():
        ms = timeit(o, lam); line += f'{k} {ms:.3f} ms ({balg/ms/1e6/6552.6*100:.0f}%)  '
    r = objs['simplex'].calculate(lam, diagnostics=True, save_primal=True)
    dg = r.projection_diag; first = sh.ccol[:-1][(sh.ccol[1:] - sh.ccol[:-1]) > 0]
    dv = dg[first]; br = (dv & 3); rho = (dv >> 2)
    x = r.primal_var
    line += f'| branch feas/short/duchi {[int((br==k).sum()) for k in range(3)]} rho>=3 {int(((br==2)&(rho>=3)).sum())} nnz(x)/E {float((x!=0).float().mean()):.3f}'
    rb = objs['box'].calculate(lam, save_primal=True); line += f' box nnz(x)/E {float((rb.primal_var!=0).float().mean()):.3f}'
    print(line, flush=True)
