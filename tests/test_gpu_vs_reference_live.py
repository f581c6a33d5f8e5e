"""-m gpu: the fused kernel against the UNMODIFIED reference running side by side on the same inputs.

oracle/_ref (the pip-installed linkedin/DuaLip v5.0.1, populated by oracle/make_ref.py in the build container; it travels to
the GPU box like the built .so) evaluates `MatchingSolverDualObjectiveFunction.calculate` on CPU tensors; the same problem --
drawn by the benchmark's generator on the device, at sizes far beyond the committed fixtures -- goes through the CUDA path.
Primal x must be bit-identical, objective and gradient within 1e-5 (north_star).  Skipped where oracle/_ref is absent.

The reference runs in a SUBPROCESS: it is a package called `dualip`, and so is this package's alias."""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest
import torch

from oracle import make_ref

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not make_ref.available(), reason="oracle/_ref (the unmodified reference) is not present")]
DEV = "cuda:0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_REF_SCRIPT = r"""
import json, sys
import numpy as np, torch
sys.path.insert(0, sys.argv[1])
from oracle import make_ref
make_ref.import_reference()
from dualip.objectives.matching import MatchingInputArgs, MatchingSolverDualObjectiveFunction
from dualip.projections.base import create_projection_map
d = np.load(sys.argv[2])
spec = json.loads(sys.argv[3])
m, n = int(d["m"]), d["ccol"].size - 1
ccol, row = torch.from_numpy(d["ccol"]), torch.from_numpy(d["row"])
out = {}
for tag, (ptype, params, cols) in spec.items():
    cols = np.asarray(eval(cols)) if isinstance(cols, str) else None
    # the reference corrupts maps with several entries (utils/sparse_utils.py:177,220): one entry per run, on that entry's columns
    sel = np.arange(n) if cols is None else cols
    lens = np.diff(d["ccol"])[sel]
    sub_ccol = np.zeros(sel.size + 1, dtype=np.int64); np.cumsum(lens, out=sub_ccol[1:])
    pos = np.repeat(d["ccol"][sel], lens) + (np.arange(lens.sum()) - np.repeat(sub_ccol[:-1], lens))
    A = torch.sparse_csc_tensor(torch.from_numpy(sub_ccol), torch.from_numpy(d["row"][pos]), torch.from_numpy(d["a"][pos]), size=(m, sel.size))
    C = torch.sparse_csc_tensor(torch.from_numpy(sub_ccol), torch.from_numpy(d["row"][pos]), torch.from_numpy(d["c"][pos]), size=(m, sel.size))
    for batching in (True, False):
        obj = MatchingSolverDualObjectiveFunction(MatchingInputArgs(A, C, create_projection_map(ptype, params, sel.size), None), gamma=float(d["gamma"]), batching=batching)
        r = obj.calculate(torch.from_numpy(d["lam"]), save_primal=True)
        out[f"{tag}_x_{int(batching)}"] = r.primal_var.numpy().copy()
        out[f"{tag}_pos"] = pos
        out[f"{tag}_grad_{int(batching)}"] = r.dual_gradient.numpy().copy()
        out[f"{tag}_cx_{int(batching)}"] = np.float64(r.dual_objective)
        out[f"{tag}_reg_{int(batching)}"] = np.float64(r.reg_penalty)
np.savez(sys.argv[4], **out)
"""


def _reference(problem, spec):
    with tempfile.TemporaryDirectory() as tmp:
        inp, outp, script = os.path.join(tmp, "in.npz"), os.path.join(tmp, "out.npz"), os.path.join(tmp, "ref.py")
        np.savez(inp, **problem)
        open(script, "w").write(_REF_SCRIPT)
        env = dict(os.environ, OMP_NUM_THREADS=str(max(1, len(os.sched_getaffinity(0)))), CUDA_VISIBLE_DEVICES="")
        res = subprocess.run([sys.executable, script, ROOT, inp, json.dumps(spec), outp], capture_output=True, text=True, timeout=900, env=env)
        assert res.returncode == 0, res.stderr[-3000:]
        return dict(np.load(outp))


@pytest.mark.parametrize("jacobi", [False, True])
def test_generator_problem_side_by_side_with_the_reference(jacobi):
    """300k entities x 2000 duals from the benchmark generator (3M nnz, column lengths 0..30), late-ish dual: simplex on the
    even entities, box on the odd ones (configs[2]'s map), plus simplex_eq on all.  The reference evaluates each entry on that
    entry's column sub-matrix in local-shard mode (b_vec=None: raw sums); the CUDA path evaluates the whole map at once."""
    from benchmark.synthetic import capacity_vector, generate_shard
    from dualip_b200.objectives.matching import MatchingInputArgs, MatchingSolverDualObjectiveFunction
    from dualip_b200.optimizers.agd import AcceleratedGradientDescent, no_iteration_callback
    from dualip_b200.preprocessing.precondition import jacobi_precondition
    from dualip_b200.projections import create_projection_map

    n, m, gamma = 300_000, 2000, 1e-3
    sh = generate_shard(n, m, 5e-3, 7, DEV)
    b = capacity_vector(sh.greedy_load, m, 5e-3, 7, DEV)
    A = torch.sparse_csc_tensor(sh.ccol, sh.row, sh.a, size=(m, n))
    C = torch.sparse_csc_tensor(sh.ccol, sh.row, sh.c, size=(m, n))
    if jacobi:
        jacobi_precondition(A, b)
    pm = {}
    pm.update(create_projection_map("simplex", {"z": 1.0}, n, indices=torch.arange(0, n, 2, device=DEV)))
    pm.update(create_projection_map("box", {"lower": 0.0, "upper": 1.0}, n, indices=torch.arange(1, n, 2, device=DEV)))
    mixed = MatchingSolverDualObjectiveFunction(MatchingInputArgs(A, C, pm, b), gamma=gamma)
    lam = AcceleratedGradientDescent(max_iter=60, gamma=gamma, initial_step_size=1e-3, max_step_size=0.1,
                                     iteration_callback=no_iteration_callback).maximize(mixed, torch.zeros(m, device=DEV)).dual_val.clone()
    problem = dict(ccol=sh.ccol.cpu().numpy(), row=sh.row.cpu().numpy(), a=A.values().cpu().numpy(), c=sh.c.cpu().numpy(),
                   lam=lam.cpu().numpy(), gamma=gamma, m=m)
    spec = {"sx": ("simplex", {"z": 1.0}, f"np.arange(0, {n}, 2)"), "bx": ("box", {"lower": 0.0, "upper": 1.0}, f"np.arange(1, {n}, 2)"),
            "eq": ("simplex_eq", {"z": 1.0}, None)}
    ref = _reference(problem, spec)
    # mixed map: x of both entries interleaved, sums added
    r = mixed.calculate(lam, save_primal=True)
    x = r.primal_var.cpu().numpy()
    for batching in (1, 0):
        for tag in ("sx", "bx"):
            assert np.array_equal(x[ref[f"{tag}_pos"]], ref[f"{tag}_x_{batching}"]), f"{tag}: x differs from the reference (batching={batching})"
        grad_ref = ref[f"sx_grad_{batching}"].astype(np.float64) + ref[f"bx_grad_{batching}"].astype(np.float64) - b.cpu().numpy()
        g = r.dual_gradient.cpu().numpy()
        assert np.allclose(g, grad_ref, rtol=1e-5, atol=1e-5 * max(1.0, np.abs(grad_ref).max()))
        cx_ref = float(ref[f"sx_cx_{batching}"]) + float(ref[f"bx_cx_{batching}"])
        assert abs(float(r.scalars64[1]) - cx_ref) <= 2e-5 * abs(cx_ref)
    # simplex_eq on all columns: the padded length of the reference's buckets is part of the result (batching on / off)
    for batching in (True, False):
        eq = MatchingSolverDualObjectiveFunction(MatchingInputArgs(A, C, create_projection_map("simplex_eq", {"z": 1.0}, n), b), gamma=gamma,
                                                 batching=batching)
        xe = eq.calculate(lam, save_primal=True).primal_var.cpu().numpy()
        assert np.array_equal(xe, ref[f"eq_x_{int(batching)}"]), f"simplex_eq: x differs from the reference (batching={batching})"
    assert not np.array_equal(ref["eq_x_1"], ref["eq_x_0"])
