"""ctypes binding of the C-ABI in include/dualip_b200.h (libdualip_b200.so, built in-tree by `make -C dualip_b200/csrc`).

There is no CPU fallback: `lib()` raises if the shared library is missing, and every entry point needs a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
# DUALIP_B200_LIB: an alternative build of the same library (kernel A/B experiments); never a different implementation
LIB_PATH = os.environ.get("DUALIP_B200_LIB") or os.path.join(_HERE, "_lib", "libdualip_b200.so")
CSRC_DIR = os.path.join(_HERE, "csrc")

OK, EINVAL, ECUDA, ENOMEM, ERANGE = 0, -1, -2, -3, -4
PROJ_CLAMP, PROJ_SIMPLEX, PROJ_SIMPLEX_EQ, PROJ_SIMPLEX_BISECT, PROJ_SIMPLEX_EQ_BISECT = 0, 1, 2, 3, 4
MAX_BISECT_COLUMN = 1024  # longest column the fused kernel projects with a bisection class
PROJ_FLAG_D1_UNPADDED = 1
PEER_HANDLE_BYTES, PEER_MAX_WORLD = 64, 16


class ProjClass(C.Structure):
    _fields_ = [("kind", C.c_int32), ("lo", C.c_float), ("hi", C.c_float), ("z", C.c_float), ("z_thr", C.c_float),
                ("flags", C.c_uint32)]


class CscDesc(C.Structure):
    _fields_ = [("n_cols", C.c_int64), ("nnz", C.c_int64), ("n_rows", C.c_int32), ("index_bits", C.c_int32),
                ("ccol_dev", C.c_void_p), ("row_dev", C.c_void_p), ("a_dev", C.c_void_p), ("c_dev", C.c_void_p),
                ("col_class_dev", C.c_void_p), ("classes", C.POINTER(ProjClass)), ("n_classes", C.c_int32),
                ("device", C.c_int32), ("pad_len", C.POINTER(C.c_int32))]


PAD_BUCKETS = 32


class LpDesc(C.Structure):
    _fields_ = [("m", C.c_int32), ("n", C.c_int32), ("nnz", C.c_int64),
                ("csr_rowptr_dev", C.c_void_p), ("csr_col_dev", C.c_void_p), ("csr_val_dev", C.c_void_p),
                ("csc_colptr_dev", C.c_void_p), ("csc_row_dev", C.c_void_p), ("csc_val_dev", C.c_void_p),
                ("c_dev", C.c_void_p), ("b_dev", C.c_void_p), ("lo_dev", C.c_void_p), ("hi_dev", C.c_void_p),
                ("row_scale_dev", C.c_void_p), ("device", C.c_int32)]


class Scalars(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("dual_objective", "primal_objective", "reg_penalty", "dual_val_times_grad",
                                          "max_pos_slack", "sum_pos_slack", "x_sq_norm", "grad_sq_norm")]


SCALAR_FIELDS = [n for n, _ in Scalars._fields_]

# name -> (restype, argtypes); must list every symbol declared in include/dualip_b200.h
SIGNATURES = {
    "dualip_abi_version": (C.c_int, []),
    "dualip_last_error": (C.c_char_p, []),
    "dualip_plan_create": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(CscDesc)]),
    "dualip_plan_destroy": (None, [C.c_void_p]),
    "dualip_plan_rebalance": (C.c_int, [C.c_void_p, C.c_void_p]),
    "dualip_plan_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.c_int]),
    "dualip_matching_calc": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]),
    "dualip_matching_partial": (C.c_int, [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_uint32, C.c_void_p]),
    "dualip_matching_epilogue": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p,
                                           C.c_void_p, C.c_void_p]),
    "dualip_lp_calc": (C.c_int, [C.POINTER(LpDesc), C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_void_p]),
    "dualip_matching_calc_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p,
                                            C.c_void_p]),
    "dualip_agd_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_double,
                                    C.c_double, C.c_int32]),
    "dualip_agd_destroy": (None, [C.c_void_p]),
    "dualip_agd_x": (C.c_void_p, [C.c_void_p]),
    "dualip_agd_y": (C.c_void_p, [C.c_void_p]),
    "dualip_agd_get": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dualip_agd_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int32, C.c_double, C.c_int32,
                                  C.c_void_p]),
    "dualip_agd_step_sharded": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_float,
                                          C.c_int32, C.c_double, C.c_int32, C.c_void_p]),
    "dualip_peer_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "dualip_peer_destroy": (None, [C.c_void_p]),
    "dualip_peer_export": (C.c_int, [C.c_void_p, C.c_void_p]),
    "dualip_peer_connect_ipc": (C.c_int, [C.c_void_p, C.c_void_p]),
    "dualip_peer_connect_ptrs": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "dualip_peer_window": (C.c_void_p, [C.c_void_p]),
    "dualip_peer_next_slot": (C.c_void_p, [C.c_void_p]),
    "dualip_peer_status": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32), C.c_void_p]),
    "dualip_agd_step_peer": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_float,
                                       C.c_int32, C.c_double, C.c_int32, C.c_void_p]),
    "dualip_peer_status_nowait": (C.c_int, [C.c_void_p]),
    "dualip_matching_ascent_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p,
                                              C.c_float, C.c_int32, C.c_double, C.c_int32, C.c_void_p]),
    "dualip_matching_ascent_step_peer": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p,
                                                   C.c_void_p, C.c_float, C.c_int32, C.c_double, C.c_int32, C.c_void_p]),
    "dualip_matching_calc_peer": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p,
                                            C.c_void_p]),
    "dualip_matching_calc_peer_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p,
                                                 C.c_void_p]),
    "dualip_agd_set_schedule": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double]),
    "dualip_agd_steps_launched": (C.c_longlong, [C.c_void_p]),
    "dualip_matching_ascent_step_scheduled": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                        C.c_void_p]),
    "dualip_ascent_graph_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                             C.c_void_p, C.c_int32]),
    "dualip_ascent_graph_launch": (C.c_int, [C.c_void_p, C.c_void_p]),
    "dualip_ascent_graph_destroy": (None, [C.c_void_p]),
    "dualip_agd_host_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int32, C.c_void_p, C.c_void_p, C.c_double, C.c_double,
                                         C.c_int32]),
    "dualip_agd_host_destroy": (None, [C.c_void_p]),
    "dualip_agd_host_x": (C.c_void_p, [C.c_void_p]),
    "dualip_agd_host_y": (C.c_void_p, [C.c_void_p]),
    "dualip_agd_host_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_float, C.c_int32, C.c_double, C.POINTER(C.c_double)]),
    "dualip_matching_step_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_float, C.c_int32, C.c_double,
                                            C.c_void_p, C.c_void_p, C.POINTER(C.c_double), C.c_void_p]),
    "dualip_agd_read_log": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dualip_agd_reserve_log": (C.c_int, [C.c_void_p, C.c_int32]),
    "dualip_row_sq_norms": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_int32, C.c_void_p, C.c_int32,
                                      C.c_void_p]),
    "dualip_scale_rows": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_int32, C.c_void_p]),
    "dualip_project_block": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.POINTER(ProjClass), C.c_void_p]),
    "dualip_csc_left_multiply": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_int32,
                                           C.c_void_p]),
    "dualip_csc_row_sums": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]),
    "dualip_csc_gather_block": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p,
                                          C.c_int32, C.c_void_p]),
    "dualip_csc_scatter_block": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p,
                                           C.c_int32, C.c_void_p]),
    "dualip_fair_calc": (C.c_int, [C.POINTER(CscDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_void_p]),
    "dualip_fair_work_bytes": (C.c_int64, [C.c_int32, C.c_int32]),
    "dualip_jacobi_precondition": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_int32,
                                             C.c_void_p, C.c_int32, C.c_void_p]),
}

_lib = None
_lock = threading.Lock()


def _source_digest() -> str:
    """sha256 over the CUDA sources, the public header and the Makefile (names and contents)."""
    import hashlib

    h = hashlib.sha256()
    files = sorted(os.path.join(CSRC_DIR, n) for n in os.listdir(CSRC_DIR)
                   if n.endswith((".cu", ".cuh", ".inc")) or n == "Makefile")
    files.append(os.path.join(os.path.dirname(CSRC_DIR), "..", "include", "dualip_b200.h"))
    for path in files:
        h.update(os.path.basename(path).encode())
        with open(path, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def build(verbose: bool = False) -> str:
    """Compile the CUDA sources for sm_100a into dualip_b200/_lib/ (nvcc cross-compiles without a GPU).  The library is up to
    date when the digest of the sources recorded beside it matches: a copied tree (file times lost, e.g. the snapshot on a GPU
    box) is not recompiled -- calc.cu alone takes minutes."""
    digest = _source_digest()
    stamp = os.path.join(os.path.dirname(LIB_PATH), "SOURCE_DIGEST")
    if os.path.exists(LIB_PATH) and os.path.exists(stamp) and open(stamp).read().strip() == digest:
        return LIB_PATH
    res = subprocess.run(["make", "-j4", "-C", CSRC_DIR], capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout)
        print(res.stderr)
    if res.returncode != 0:
        raise RuntimeError("building libdualip_b200.so failed")
    with open(stamp, "w") as fh:
        fh.write(digest + "\n")
    return LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(
                    f"{LIB_PATH} is missing: build it with `make -C {CSRC_DIR}` (or __graft_entry__.build()). "
                    "dualip_b200 has no CPU fallback."
                )
            handle = C.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(handle, name)
                fn.restype = res
                fn.argtypes = args
            _lib = handle
    return _lib


class NativeError(RuntimeError):
    pass


def check(rc: int, what: str = "") -> None:
    """Maps C error codes to the exception types the reference raises for the same conditions."""
    if rc == OK:
        return
    msg = lib().dualip_last_error().decode("utf-8", "replace")
    text = f"{what}: {msg}" if what else msg
    if rc in (EINVAL, ERANGE):
        raise ValueError(text)
    if rc == ENOMEM:
        raise MemoryError(text)
    raise NativeError(text)
