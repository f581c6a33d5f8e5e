"""Fixtures for the fairness-row objective and the public CSC operators, produced by the UNMODIFIED reference operators
(oracle/_ref or /root/reference) in the build container; committed because the GPU box has no reference tree.

    python tests/golden/make_golden_fair.py

The reference has no fairness class in src/: docs/demo/matching_complex.rst shows how a user builds one by subclassing
MatchingSolverDualObjectiveFunction and composing left_multiply_sparse / elementwise_csc / apply_F_to_columns /
row_sums_csc / calc_grad.  `DemoFairnessObjective` below is that recipe (rst:46-64 and rst:86-167) executed with the
reference's own operators on CPU tensors; fair_*.npz hold its inputs and outputs.
"""
import os
import sys
from operator import add, mul

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    from oracle import make_ref

    make_ref.make()
    make_ref.import_reference()
    from dualip.objectives.matching import MatchingInputArgs, MatchingSolverDualObjectiveFunction, calc_grad
    from dualip.projections.base import create_projection_map, project
    from dualip.utils.sparse_utils import (apply_F_to_columns, elementwise_csc, hstack_csc, left_multiply_sparse, row_sums_csc,
                                           split_csc_by_cols)
    from conftest import random_problem

    class DemoFairnessObjective(MatchingSolverDualObjectiveFunction):
        def __init__(self, args, gamma, batching, group_ratio):
            super().__init__(args, gamma, batching)
            self.group_ratio = group_ratio
            self.A_fairness = self._build_fairness_constraints()

        def _build_fairness_constraints(self):  # rst:46-64
            num_cols = self.A.size(1)
            g1 = max(0, min(int(num_cols * self.group_ratio), num_cols))
            g2 = num_cols - g1
            A1, A2 = split_csc_by_cols(self.A, [g1, g2])
            return hstack_csc([1 / g1 * A1, -1 / g2 * A2])

        def calculate(self, dual_val, gamma=None, save_primal=False):  # rst:86-167
            grad = torch.zeros_like(dual_val)
            if gamma is not None:
                self.gamma = gamma
            scaled = -1.0 / self.gamma * dual_val
            left_multiply_sparse(scaled[:-2], self.A, output_tensor=self.intermediate)
            elementwise_csc(self.intermediate, scaled[-2] * self.A_fairness, add, output_tensor=self.intermediate)
            elementwise_csc(self.intermediate, -1 * scaled[-1] * self.A_fairness, add, output_tensor=self.intermediate)
            elementwise_csc(self.intermediate, self.c_rescaled, add, output_tensor=self.intermediate)
            for _, (buckets, proj_type, proj_params) in self.buckets.items():
                apply_F_to_columns(self.intermediate, project(proj_type, **proj_params), buckets, output_tensor=self.intermediate)
            grad[:-2] = row_sums_csc(elementwise_csc(self.A, self.intermediate, mul))
            grad[-2] = elementwise_csc(self.A_fairness, self.intermediate, mul).values().sum()
            grad[-1] = elementwise_csc(-1 * self.A_fairness, self.intermediate, mul).values().sum()
            vals = self.intermediate.values()
            reg = (self.gamma / 2) * torch.norm(vals) ** 2
            dual_obj = torch.dot(self.c.values(), vals)
            primal_obj = dual_obj.clone()
            grad, dual_obj = calc_grad(grad, dual_obj, dual_val, self.b_vec, reg)
            return dict(grad=grad, dual_obj=dual_obj, reg=reg, primal_obj=primal_obj, lam_grad=torch.dot(dual_val, grad),
                        max_pos=max(torch.max(grad), 0), sum_pos=torch.relu(grad).sum(), x=vals.clone())

    for name, ptype, params, seed in (("simplex", "simplex", {"z": 1.0}, 21), ("box", "box", {"lower": 0.0, "upper": 0.4}, 22),
                                      ("simplex_eq", "simplex_eq", {"z": 0.8}, 23)):
        p = random_problem(seed, 4000, 48, 6.0, scale_c={"simplex": 12.0, "box": 1.5, "simplex_eq": 2.5}[name], lam_scale=0.03)
        n, m = p["ccol"].size - 1, 48
        rng = np.random.default_rng(seed)
        # large fairness duals: scaled[-2] * f is of the order of the other terms although f carries 1/|group|
        lam = np.concatenate([p["lam"], (rng.random(2) * 30).astype(np.float32)]).astype(np.float32)
        b = np.concatenate([p["b"], np.float32([0.01, 0.01])]).astype(np.float32)
        gamma, ratio = 0.05, 0.35
        ccol, row = torch.from_numpy(p["ccol"]), torch.from_numpy(p["row"])
        A = torch.sparse_csc_tensor(ccol, row, torch.from_numpy(p["a"]), size=(m, n))
        C = torch.sparse_csc_tensor(ccol, row, torch.from_numpy(p["c"]), size=(m, n))
        out = {}
        for batching in (True, False):
            obj = DemoFairnessObjective(MatchingInputArgs(A, C, create_projection_map(ptype, params, n), torch.from_numpy(b)),
                                        gamma, batching, ratio)
            r = obj.calculate(torch.from_numpy(lam), save_primal=True)
            tag = "b1" if batching else "b0"
            out[f"x_{tag}"] = r["x"].numpy()
            out[f"grad_{tag}"] = r["grad"].numpy()
            out[f"scal_{tag}"] = np.array([float(r["dual_obj"]), float(r["primal_obj"]), float(r["reg"]), float(r["lam_grad"]),
                                           float(r["max_pos"]), float(r["sum_pos"])])
            out["f"] = obj.A_fairness.values().numpy()
        np.savez_compressed(os.path.join(HERE, f"fair_{name}.npz"), ccol=p["ccol"], row=p["row"], a=p["a"], c=p["c"], b=b, lam=lam,
                            n_rows=m, gamma=gamma, group_ratio=ratio, proj_type=ptype, proj_keys=np.array(list(params)),
                            proj_vals=np.array(list(params.values()), dtype=np.float64), **out)
        nz = (out["x_b1"] != 0).mean()
        xs = out["x_b1"]
        print("   share of x strictly inside (0, 0.4): %.3f; at 0: %.3f" % (((xs > 0) & (xs < 0.4)).mean(), (xs == 0).mean()))
        print(name, "nnz", p["row"].size, "nonzero x share %.3f" % nz, "dual_obj", out["scal_b1"][0], "fair grads", out["grad_b1"][-2:],
              "b1 == b0:", np.array_equal(out["x_b1"], out["x_b0"]))

    # operator fixtures: the reference's operators on a ragged random matrix
    p = random_problem(31, 500, 30, 5.0)
    n, m = 500, 30
    ccol, row = torch.from_numpy(p["ccol"]), torch.from_numpy(p["row"])
    M = torch.sparse_csc_tensor(ccol, row, torch.from_numpy(p["a"]) - 0.01, size=(m, n))
    v = torch.from_numpy(np.random.default_rng(1).standard_normal(m).astype(np.float32))
    lm = left_multiply_sparse(v, M).values().numpy()
    rs = row_sums_csc(M).numpy()
    cols_a = torch.arange(0, n, 2)
    cols_b = torch.arange(1, n, 2)
    ap = apply_F_to_columns(M, project("simplex", z=0.05), [torch.arange(n)]).values().numpy()
    ap2 = M.values().clone().numpy()
    tmp = apply_F_to_columns(M, lambda blk: blk * 2 + (blk != 0) * 1.0, [cols_a]).values().numpy()  # only even columns are defined
    np.savez_compressed(os.path.join(HERE, "ops_reference.npz"), ccol=p["ccol"], row=p["row"], vals=M.values().numpy(), v=v.numpy(),
                        n_rows=m, left_multiply=lm, row_sums=rs, apply_simplex_all=ap, apply_affine_even=tmp,
                        even_cols=cols_a.numpy())
    print("ops fixture written")


if __name__ == "__main__":
    main()
