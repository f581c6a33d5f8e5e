"""Experiment helper: save iterates of the c3_small trajectory / time one class at a saved iterate (not product code)."""
import os, sys
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
mode = sys.argv[1]
sh = generate_shard(n, m, sp, 42, dev); b = capacity_vector(sh.greedy_load, m, sp, 42, dev)
A = torch.sparse_csc_tensor(sh.ccol, sh.row, sh.a, size=(m, n)); C = torch.sparse_csc_tensor(sh.ccol, sh.row, sh.c, size=(m, n))
jacobi_precondition(A, b)
root = os.path.dirname(os.path.abspath(__file__))
if mode == 'save':
    obj = MatchingSolverDualObjectiveFunction(MatchingInputArgs(A, C, bench.mixed_projection_map(n, 0, dev), b), gamma=1e-3)
    solver = AcceleratedGradientDescent(max_iter=100, gamma=1e-3, initial_step_size=1e-3, max_step_size=1e-1, iteration_callback=lambda i, r: None)
    loop = FusedAscentLoop(solver, obj, torch.zeros(m, device=dev))
    out = {}
    for i in range(1, 101):
        loop.step(i)
        if i in (30, 100): out[f'lam{i}'] = loop.current_dual().cpu().numpy()
    np.savez(os.path.join(root, '..', 'gpurun_out', 'lams_c3small.npz'), **out)
else:
    kind, it = mode, int(sys.argv[2])
    lam = torch.from_numpy(np.load(os.path.join(root, 'lams_c3small.npz'))[f'lam{it}']).to(dev)
    pm = {'mixed': lambda: bench.mixed_projection_map(n, 0, dev), 'simplex': lambda: create_projection_map('simplex', {'z': 1.0}, n), 'box': lambda: create_projection_map('box', {'lower': 0.0, 'upper': 1.0}, n)}[kind]()
    obj = MatchingSolverDualObjectiveFunction(MatchingInputArgs(A, C, pm, b), gamma=1e-3)
    grad = torch.empty(m, device=dev); scal = torch.zeros(8, dtype=torch.float64, device=dev)
    reps = int(os.environ.get('REPS', 5))
    for _ in range(2): obj.launch_calc(lam.data_ptr(), 1e-3, grad.data_ptr(), scal.data_ptr())
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): obj.launch_calc(lam.data_ptr(), 1e-3, grad.data_ptr(), scal.data_ptr())
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f'{kind} iter {it}: {ms:.3f} ms  {obj.algorithmic_bytes()/ms/1e6:.0f} GB/s ({obj.algorithmic_bytes()/ms/1e6/6552.6*100:.0f}%)', flush=True)
