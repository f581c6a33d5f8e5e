import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch

        has_cuda = torch.cuda.is_available()
    except Exception:
        has_cuda = False
    if has_cuda:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def golden_cases():
    return sorted(f[5:-4] for f in os.listdir(GOLDEN) if f.startswith("case_") and f.endswith(".npz"))


def load_case(name):
    d = np.load(os.path.join(GOLDEN, f"case_{name}.npz"))
    params = {str(k): float(v) for k, v in zip(d["proj_keys"], d["proj_vals"])}
    return d, str(d["proj_type"]), params


def random_csc(rng, n_cols, n_rows, mean_deg, max_deg=None, empty_frac=0.05, long_cols=()):
    """Random ragged CSC pattern: Poisson column lengths, some empty columns, sorted distinct rows per column."""
    deg = rng.poisson(mean_deg, size=n_cols)
    deg = np.minimum(deg, n_rows if max_deg is None else min(max_deg, n_rows))
    deg[rng.random(n_cols) < empty_frac] = 0
    for j, d in long_cols:
        deg[j] = min(d, n_rows)
    ccol = np.zeros(n_cols + 1, dtype=np.int64)
    np.cumsum(deg, out=ccol[1:])
    parts = [np.sort(rng.choice(n_rows, size=d, replace=False)) for d in deg]
    row = np.concatenate(parts + [np.zeros(0, dtype=np.int64)]).astype(np.int64)
    return ccol, row


def random_problem(seed, n_cols, n_rows, mean_deg, scale_c=1.0, lam_scale=1.0, **kw):
    rng = np.random.default_rng(seed)
    ccol, row = random_csc(rng, n_cols, n_rows, mean_deg, **kw)
    E = row.size
    c = (-np.minimum(rng.lognormal(-4.0, 0.75, E) * rng.lognormal(0, 0.7, E), 0.5) * scale_c).astype(np.float32)
    a = (rng.lognormal(0, 1, E) * (-c)).astype(np.float32)
    b = (rng.uniform(0.5, 1.0, n_rows) * 0.05 * n_cols / n_rows).astype(np.float32)
    lam = (rng.random(n_rows) * lam_scale).astype(np.float32)
    return dict(ccol=ccol, row=row, a=a, c=c, b=b, lam=lam, n_rows=n_rows, n_cols=n_cols)
