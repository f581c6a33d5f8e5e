"""Generates tests/golden/lp_*.npz by running the UNMODIFIED reference's generic-LP objective
(MIPLIB2017ObjectiveFunction, /root/reference/src/dualip/objectives/miplib.py) in the build container:

* lp_miplib: the shipped MIPLIB-2017 instance examples/miplib_2017/v150d30-2hopcds.mps.gz read with the reference's own
  MPS parser (A 7822 x 150, `UP 1` bounds, no equality rows); calculate() at two dual points and a 60-iteration
  AcceleratedGradientDescent trace with the gamma step-decay schedule of config 5;
* lp_eq_cone: a derived small LP (ours) with equality rows, one-sided (cone) and box bounds and Jacobi row scaling on a
  dense A, the combination the shipped instance does not exercise.

    python tests/golden/make_golden_lp.py         # needs /root/reference; writes next to this file
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import REF, _import_reference  # noqa: E402


def trace(obj, m, eq_mask, gamma, iters, decay):
    import torch
    from dualip.optimizers.agd import AcceleratedGradientDescent

    kw = dict(gamma_decay_type="step", gamma_decay_params=decay) if decay else {}
    solver = AcceleratedGradientDescent(max_iter=iters, gamma=gamma, initial_step_size=1e-3, max_step_size=0.1, **kw)
    solver.iteration_callback = lambda i, r: None
    res = solver.maximize(obj, torch.zeros(m))
    return (np.array(res.dual_objective_log, dtype=np.float64), np.array(res.step_size_log, dtype=np.float64),
            res.dual_val.numpy().copy())


def case(A_dense, c, b, lower, upper, eq_mask, lams, gamma, jacobi, iters, decay, pm):
    import torch
    from dualip.objectives.miplib import MIPLIB2017ObjectiveFunction, MIPLIBInputArgs

    At = torch.from_numpy(A_dense)
    args = MIPLIBInputArgs(A=At if jacobi else At.to_sparse(), c=torch.from_numpy(c), projection_map=pm,
                           b_vec=torch.from_numpy(b), equality_mask=torch.from_numpy(eq_mask) if eq_mask is not None else None)
    obj = MIPLIB2017ObjectiveFunction(args, use_jacobi_precondition=jacobi)
    out = dict(A=A_dense, c=c, b=b, lower=lower, upper=upper, gamma=np.float64(gamma), jacobi=np.int32(jacobi),
               eq_mask=eq_mask if eq_mask is not None else np.zeros(0, dtype=bool), lams=np.stack(lams))
    for k, lam in enumerate(lams):
        r = obj.calculate(torch.from_numpy(lam), gamma=gamma, save_primal=True)
        out[f"grad{k}"] = r.dual_gradient.numpy().copy()
        out[f"x{k}"] = r.primal_var.numpy().copy()
        out[f"scal{k}"] = np.array([float(r.dual_objective), float(r.reg_penalty), float(r.primal_objective)])
    obj_log, step_log, dual = trace(obj, b.size, eq_mask, gamma, iters, decay)
    out.update(obj_log=obj_log, step_log=step_log, dual=dual, iters=np.int32(iters),
               decay=np.array([decay["decay_steps"], decay["decay_factor"]] if decay else [0, 1.0], dtype=np.float64))
    return out


def main():
    _import_reference()
    import torch
    from dualip.projections.base import create_projection_map

    rng = np.random.default_rng(20260118)
    # ---- the shipped instance, through the reference's own parser ----
    sys.path.insert(0, os.path.join(REF, "examples", "miplib_2017"))
    from read_mps_data import read_mps_file

    data = read_mps_file(os.path.join(REF, "examples", "miplib_2017", "v150d30-2hopcds.mps.gz")).to_dualip_format()
    A = data.A.to_dense().numpy().astype(np.float32) if data.A.layout != torch.strided else data.A.numpy().astype(np.float32)
    c, b = data.C.numpy().astype(np.float32), data.b_vec.numpy().astype(np.float32)
    n = c.size
    lower, upper = np.full(n, -np.inf, dtype=np.float32), np.full(n, np.inf, dtype=np.float32)
    for item in data.projection_map.values():
        idx = np.asarray(item.indices, dtype=np.int64)
        p = item.proj_params
        if item.proj_type == "box":
            lower[idx] = p.get("lower", 0.0)
            upper[idx] = p.get("upper", 1.0)
        else:
            if p.get("lower") is not None:
                lower[idx] = p["lower"]
            if p.get("upper") is not None:
                upper[idx] = p["upper"]
    eq = data.equality_mask.numpy() if data.equality_mask is not None else None
    print("miplib instance:", A.shape, "nnz", int((A != 0).sum()), "equalities", 0 if eq is None else int(eq.sum()),
          "proj types", sorted({v.proj_type for v in data.projection_map.values()}))
    lams = [np.zeros(b.size, dtype=np.float32), (rng.random(b.size) * 0.02).astype(np.float32)]
    d = case(A, c, b, lower, upper, eq, lams, 1e-3, False, 60, {"decay_steps": 35, "decay_factor": 0.7}, data.projection_map)
    np.savez_compressed(os.path.join(HERE, "lp_miplib.npz"), **d)
    print("lp_miplib: obj log head/tail", d["obj_log"][:2], d["obj_log"][-2:])

    # ---- derived LP: equalities + cone + box + Jacobi (dense A) ----
    m, n = 40, 25
    A = (rng.standard_normal((m, n)) * (rng.random((m, n)) < 0.3)).astype(np.float32)
    A[np.abs(A).sum(1) == 0, 0] = 1.0
    c = rng.standard_normal(n).astype(np.float32)
    b = rng.standard_normal(m).astype(np.float32)
    eq = np.zeros(m, dtype=bool)
    eq[::5] = True
    pm = {}
    pm.update(create_projection_map("cone", {"lower": 0.0}, n, indices=list(range(0, 10)), key_prefix="a_"))
    pm.update(create_projection_map("cone", {"upper": 2.0}, n, indices=list(range(10, 15)), key_prefix="b_"))
    pm.update(create_projection_map("box", {"lower": -1.0, "upper": 3.0}, n, indices=list(range(15, 22)), key_prefix="c_"))
    lower, upper = np.full(n, -np.inf, dtype=np.float32), np.full(n, np.inf, dtype=np.float32)
    lower[0:10] = 0.0
    upper[10:15] = 2.0
    lower[15:22], upper[15:22] = -1.0, 3.0
    lams = [np.zeros(m, dtype=np.float32), (rng.standard_normal(m) * 0.3).astype(np.float32)]
    d = case(A, c, b, lower, upper, eq, lams, 5e-2, True, 40, {"decay_steps": 10, "decay_factor": 0.5}, pm)
    np.savez_compressed(os.path.join(HERE, "lp_eq_cone.npz"), **d)
    print("lp_eq_cone: obj log head/tail", d["obj_log"][:2], d["obj_log"][-2:])


if __name__ == "__main__":
    main()
