import glob, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import dualip_oracle as O
from dualip_b200.objectives.matching import MatchingInputArgs, MatchingSolverDualObjectiveFunction
from dualip_b200.projections import create_projection_map
from dualip_b200.optimizers.agd import AcceleratedGradientDescent

dev = torch.device('cuda:0')
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for f in sorted(glob.glob(os.path.join(root, 'tests/golden/case_*.npz'))):
    d = np.load(f)
    n = d['ccol'].size - 1; m = int(d['n_rows'])
    params = {str(k): float(v) for k, v in zip(d['proj_keys'], d['proj_vals'])}
    ptype = str(d['proj_type'])
    A = torch.sparse_csc_tensor(torch.from_numpy(d['ccol']), torch.from_numpy(d['row']), torch.from_numpy(d['a']), size=(m, n)).to(dev)
    Cm = torch.sparse_csc_tensor(torch.from_numpy(d['ccol']), torch.from_numpy(d['row']), torch.from_numpy(d['c']), size=(m, n)).to(dev)
    pm = create_projection_map(ptype, params, n)
    obj = MatchingSolverDualObjectiveFunction(MatchingInputArgs(A, Cm, pm, torch.from_numpy(d['b']).to(dev)), gamma=float(d['gamma']))
    print(os.path.basename(f), obj.plan_info())
    lam = torch.from_numpy(d['lam']).to(dev)
    for rep in range(3):
        r = obj.calculate(lam, save_primal=True, diagnostics=True)
    torch.cuda.synchronize()
    x = r.primal_var.cpu().numpy(); xr = d['x_b1']
    orc = O.matching_calculate(d['ccol'], d['row'], d['a'], d['c'], m, {'k': O.ProjEntry(ptype, params, np.arange(n))}, d['lam'], float(d['gamma']), d['b'])
    print('  x exact', np.array_equal(x, xr), 'maxabs', float(np.abs(x - xr).max()), 'support mismatch', int(((x != 0) != (xr != 0)).sum()))
    g = r.dual_gradient.cpu().numpy()
    print('  grad max rel', float(np.max(np.abs(g - d['grad_b1']) / np.maximum(np.abs(d['grad_b1']), 1e-6))), 'obj', float(r.dual_objective), d['scal_b1'][0],
          'rel', abs(float(r.scalars64[0]) - d['scal_b1'][0]) / abs(d['scal_b1'][0]))
    print('  scal', r.scalars64.cpu().numpy()[:6], d['scal_b1'])
    if ptype.startswith('simplex'):
        diag = r.projection_diag.cpu().numpy()
        first = d['ccol'][:-1][np.diff(d['ccol']) > 0]
        dv = diag[first]
        ob = orc.branch[np.diff(d['ccol']) > 0]; orho = orc.rho[np.diff(d['ccol']) > 0]
        print('  branch exact', np.array_equal(dv & 3, ob), 'rho exact', np.array_equal((dv >> 2)[ob > 0], np.minimum(orho[ob > 0], 63)), np.bincount(dv & 3, minlength=3))

# 5x5 golden through fused maximize
a = torch.tensor([[0.307766110869125,0.483770735096186,0.624996477039531,0.669021712383255,0.535811153938994],
 [0.257672501029447,0.812402617651969,0.882165518123657,0.204612161964178,0.710803845431656],
 [0.552322433330119,0.370320537127554,0.28035383997485,0.357524853432551,0.538348698290065],
 [0.0563831503968686,0.546558595029637,0.398487901547924,0.359475114848465,0.74897222686559],
 [0.468549283919856,0.170262051047757,0.76255108229816,0.690290528349578,0.420101450523362]])
A = a.T.to_sparse_csc().to(dev); Cm = (-a).T.to_sparse_csc().to(dev)
b = torch.full((5,), 0.7, device=dev)
obj = MatchingSolverDualObjectiveFunction(MatchingInputArgs(A, Cm, create_projection_map('simplex', {'z': 1}, 5), b), gamma=1e-3)
solver = AcceleratedGradientDescent(max_iter=30, gamma=1e-3, iteration_callback=lambda i, r: None)
res = solver.maximize(obj, 0.1 * torch.ones(5, device=dev))
for i, tv in [(2, -3.6010155991401818), (16, -3.60842718733725), (23, -3.5080258013053136), (29, -3.4868496294227143)]:
    print('iter', i, res.dual_objective_log[i - 1], tv, abs(res.dual_objective_log[i - 1] - tv) < 1e-5)
