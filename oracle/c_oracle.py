"""ctypes wrapper of oracle/matching_oracle.c (TEST INFRASTRUCTURE / CPU baseline only)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "libmatching_oracle.so")


class OracleClass(C.Structure):
    _fields_ = [("kind", C.c_int32), ("lo", C.c_float), ("hi", C.c_float), ("z", C.c_float), ("z_thr", C.c_float),
                ("flags", C.c_uint32)]


def build() -> str:
    res = subprocess.run(["make", "-C", _HERE], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("building the C oracle failed:\n" + res.stdout + res.stderr)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        _lib = C.CDLL(LIB_PATH)
        _lib.oracle_matching_calculate.restype = C.c_int
        _lib.oracle_matching_calculate.argtypes = [C.c_int64, C.c_int64, C.c_int32] + [C.c_void_p] * 8 + [C.c_double] + [C.c_void_p] * 4 + [C.c_int, C.c_void_p]
        _lib.oracle_max_threads.restype = C.c_int
    return _lib


def make_class(proj_type: str, params: dict, d1_unpadded: bool = False) -> OracleClass:
    flags = 1 if d1_unpadded else 0
    if proj_type == "box":
        return OracleClass(0, float(params.get("lower", 0.0)), float(params.get("upper", 1.0)), 1.0, 1.0, 0)
    if proj_type == "cone":
        lo, hi = params.get("lower"), params.get("upper")
        if lo is not None and hi is not None:
            raise ValueError("Only one of 'lower' or 'upper' should be specified, not both.")
        return OracleClass(0, -np.inf if lo is None else float(lo), np.inf if hi is None else float(hi), 1.0, 1.0, 0)
    if proj_type in ("simplex", "simplex_eq"):
        z = float(params.get("z", 1.0))
        return OracleClass(1 if proj_type == "simplex" else 2, 0.0, 0.0, float(np.float32(z)), float(np.float32(z + 1e-6)), flags)
    if proj_type == "identity":
        return OracleClass(0, -np.inf, np.inf, 1.0, 1.0, 0)
    raise ValueError(f"Unknown projection operator '{proj_type}'")


def calculate(ccol, row, a, c, n_rows, classes, lam, gamma, b=None, col_class=None, want_x=True, want_diag=True, threads=0,
              pad_len=None):
    """Returns dict(grad, scal[7], x, diag).  classes: list[OracleClass]; col_class: uint8 per column or None;
    pad_len: int32 [n_classes, 32] padded block lengths (dualip_oracle.pad_table) or None."""
    ccol = np.ascontiguousarray(ccol, dtype=np.int64)
    row = np.ascontiguousarray(row, dtype=np.int64)
    a = np.ascontiguousarray(a, dtype=np.float32)
    c = np.ascontiguousarray(c, dtype=np.float32)
    lam = np.ascontiguousarray(lam, dtype=np.float32)
    n, nnz = ccol.size - 1, row.size
    cls_arr = (OracleClass * len(classes))(*classes)
    grad = np.empty(n_rows, dtype=np.float32)
    scal = np.zeros(7, dtype=np.float64)
    x = np.empty(nnz, dtype=np.float32) if want_x else None
    diag = np.empty(n, dtype=np.uint8) if want_diag else None
    bb = np.ascontiguousarray(b, dtype=np.float32) if b is not None else None
    cc = np.ascontiguousarray(col_class, dtype=np.uint8) if col_class is not None else None
    pl = np.ascontiguousarray(pad_len, dtype=np.int32) if pad_len is not None else None
    assert pl is None or pl.shape == (len(classes), 32)

    def p(arr):
        return arr.ctypes.data_as(C.c_void_p) if arr is not None else None

    rc = lib().oracle_matching_calculate(n, nnz, int(n_rows), p(ccol), p(row), p(a), p(c), p(cc), C.cast(cls_arr, C.c_void_p),
                                         p(lam), p(bb), float(gamma), p(grad), p(scal), p(x), p(diag), int(threads), p(pl))
    if rc != 0:
        raise RuntimeError("oracle_matching_calculate failed")
    return dict(grad=grad, scal=scal, x=x, diag=diag)


def max_threads() -> int:
    return int(lib().oracle_max_threads())
