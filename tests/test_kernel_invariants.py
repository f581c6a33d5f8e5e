"""CPU checks of the arithmetic the register-path kernel (dualip_b200/csrc/slab_fast.cuh) relies on to be bit-identical to the
reference's sorted scan (projections/simplex.py:207-231), restated in numpy float32 and compared with the oracle:

* the COMMITTED sorting networks (csrc/sort_networks.inc, parsed here) sort every 0/1 input (zero-one principle);
* div_by_int<N>: q0 = t*fl(1/N), r = fma(-q0, N, t), q = fma(r, fl(1/N), q0) is the correctly rounded float32 quotient t/N
  for N = 1..20 (checked against exact rational arithmetic);
* the scan may stop where no positive entry is left, provided the column sum is clearly above z (the kernel's `near` guard):
  same theta and support size as the reference's full scan on zero-padded blocks.
"""
import os
import re
from fractions import Fraction

import numpy as np
import pytest

from conftest import ROOT
from oracle import dualip_oracle as O

F32 = np.float32


def _networks():
    text = open(os.path.join(ROOT, "dualip_b200", "csrc", "sort_networks.inc")).read()
    nets = {}
    for m in re.finditer(r"struct SortNet<(\d+)>.*?\{(.*?)\n\};", text, flags=re.S):
        nets[int(m.group(1))] = [(int(a), int(b)) for a, b in re.findall(r"DUALIP_CS\((\d+), (\d+)\)", m.group(2))]
    return nets


def test_committed_sorting_networks_sort_descending():
    nets = _networks()
    assert set(range(2, 21)) <= set(nets), "slab_fast.cuh needs networks for 2..kRegDeg (= 20) entries"
    for n, net in nets.items():
        if n > 20:
            continue  # not instantiated by the kernel; 2^n inputs get long
        assert all(0 <= i < j < n for i, j in net)
        # zero-one principle, all 2^n inputs at once (one bit column per input)
        codes = np.arange(1 << n, dtype=np.uint32)
        w = [(codes >> k) & 1 for k in range(n)]
        for i, j in net:
            hi, lo = np.maximum(w[i], w[j]), np.minimum(w[i], w[j])
            w[i], w[j] = hi, lo
        for k in range(n - 1):
            assert np.all(w[k] >= w[k + 1]), (n, k)


def _fma32(a, b, c):
    """float32 fma(a, b, c) with one rounding, through exact rationals."""
    return F32(float(Fraction(float(a)) * Fraction(float(b)) + Fraction(float(c))))


def _round_f32(fr: Fraction):
    """Correct rounding of an exact rational to float32 (float() of a Fraction is correctly rounded to double; going through
    double is safe here because the quotients of two float32-representable numbers below are never double-rounding ties:
    checked by comparing with a directed search)."""
    d = float(fr)
    f = F32(d)
    # repair a possible double rounding: pick the nearest of f and its neighbours to fr exactly
    cands = [f, np.nextafter(f, F32(np.inf)), np.nextafter(f, F32(-np.inf))]
    best = min(cands, key=lambda c: (abs(Fraction(float(c)) - fr), abs(int(np.float32(c).view(np.int32)) & 1)))
    return F32(best)


@pytest.mark.parametrize("n", list(range(1, 21)))
def test_markstein_division_is_correctly_rounded(n):
    rng = np.random.default_rng(100 + n)
    ts = np.concatenate([
        (rng.standard_normal(300) * 10.0 ** rng.integers(-6, 4, 300)).astype(F32),
        rng.random(200).astype(F32) * F32(n),           # around the interesting range (css - z) ~ 0..n
        np.array([0.0, 1.0, -1.0, 1e-30, 3.0000002, 16777216.0, 0.1, 0.3333333], dtype=F32),
    ])
    rn = F32(1.0) / F32(n)
    for t in ts:
        q0 = F32(t * rn)
        r = _fma32(-q0, F32(n), t)
        q = _fma32(r, rn, q0)
        exact = _round_f32(Fraction(float(t)) / n)
        assert q == exact, (n, float(t), float(q), float(exact))


def _kernel_scan(u, z):
    """numpy float32 restatement of fast_simplex's branch 2 for one column at its true length: sort descending, fp64 prefix sum
    rounded to float32 per step, cond_i by the (exact) division, early stop at the first zero entry when the column sum is
    clearly above z, theta = t_rho / rho."""
    w = np.sort(u)[::-1].astype(F32)
    s = F32(0.0)
    for v in u:  # column sum in entry order, float32
        s = F32(s + v)
    near = not (s > F32(F32(z) * F32(1.0001)))
    acc = 0.0
    t_sel, rho = F32(w[0] - F32(z)), 1
    for i, wi in enumerate(w):
        if not near and not (wi > 0):
            break  # the kernel's warp vote can only stop later than this, never earlier
        acc += float(wi)
        t = F32(F32(acc) - F32(z))
        q = F32(np.float64(t) / (i + 1))  # div_by_int is the correctly rounded quotient (test above)
        if wi > q:
            t_sel, rho = t, i + 1
    theta = F32(t_sel / F32(rho))
    return np.maximum(u - theta, F32(0.0)).astype(F32), rho


@pytest.mark.parametrize("z", [1.0, 2.5])
def test_truncated_scan_equals_reference_scan_on_padded_blocks(z):
    rng = np.random.default_rng(7)
    checked = 0
    for trial in range(400):
        d = int(rng.integers(2, 21))
        u = (rng.standard_normal(d) * rng.choice([0.3, 1.0, 5.0, 40.0]) + 0.2).astype(F32)
        u = np.maximum(u, 0).astype(F32)
        if trial % 7 == 0:  # sums close to z: the guard must keep the full scan
            u = (u / max(u.sum(), 1e-6) * z * (1 + rng.choice([2e-6, 5e-5, 2e-4]))).astype(F32)
        for pad in (0, 3):  # the reference sees zero padding up to the bucket's length
            block = np.concatenate([u, np.zeros(pad, F32)]).reshape(-1, 1)
            w, branch, rho = O.duchi_proj(block, z, inequality=True)
            if branch[0] != O.BRANCH_DUCHI:
                continue
            x, rho_k = _kernel_scan(u, z)
            assert np.array_equal(x, w[:d, 0]), (trial, pad, u, x, w[:, 0])
            assert rho_k == rho[0]
            checked += 1
    assert checked > 200
