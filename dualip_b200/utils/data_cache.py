"""Shard-direct reader (and writer) for the on-disk cache of the reference's synthetic generator.

Layout (reference benchmark/generate_synthetic_data.py:172-342): in `cache_dir`, files named
`s{sources}_d{destinations}_sp{sparsity}_{dtype}_seed{seed}_{A_ccol,A_row,A_vals,c_vals,b_vec}.dat` (raw numpy memmaps)
plus `..._meta.json` holding the key, the array shapes and the numpy dtypes.  `c_vals` is stored positive; the generator
negates it when it builds the tensors (:448), and so does this reader.

The reference loads every array on rank 0, builds the full problem, splits it with one `.item()` per shard and pickles
the shards to the other ranks (benchmark/run_matching_benchmark_dist.py:44-100).  Here every rank maps the files and copies
only its own column range [col_start, col_end) to its device: ccol is read for the range, rebased to 0
(utils/sparse_utils.py:281), and the three nnz-length arrays are read for [ccol[col_start], ccol[col_end]).
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from dualip_b200.utils.dist_utils import shard_sizes

ARRAYS = ("A_ccol", "A_row", "A_vals", "c_vals", "b_vec")


def cache_prefix(num_sources: int, num_destinations: int, target_sparsity: float, dtype=torch.float32, seed: int = 42) -> str:
    """File-name prefix of a cache entry (generate_synthetic_data.py:183-187)."""
    return (f"s{int(num_sources)}_d{int(num_destinations)}_sp{float(target_sparsity)}_"
            f"{str(dtype).replace('torch.', '')}_seed{int(seed)}")


@dataclass
class CachedShard:
    ccol: torch.Tensor  # (n_local + 1) rebased to 0, dtype as stored (int64 in the reference)
    row: torch.Tensor
    a: torch.Tensor  # float32
    c: torch.Tensor  # float32, negated (minimisation convention)
    b: torch.Tensor  # float32 (m), the full vector: replicated on every rank
    n_rows: int
    n_cols_total: int
    col_start: int
    col_end: int

    def csc(self):
        """(A, c) as torch.sparse_csc tensors of shape (m, n_local) sharing one pattern (MatchingInputArgs layout)."""
        size = (self.n_rows, self.col_end - self.col_start)
        return (torch.sparse_csc_tensor(self.ccol, self.row, self.a, size=size),
                torch.sparse_csc_tensor(self.ccol, self.row, self.c, size=size))


def read_meta(cache_dir: str, prefix: str) -> dict:
    with open(os.path.join(cache_dir, f"{prefix}_meta.json")) as fh:
        meta = json.load(fh)
    for name in ARRAYS:
        if name not in meta.get("shapes", {}) or name not in meta.get("array_dtypes", {}):
            raise ValueError(f"cache metadata of {prefix} lacks '{name}'")
    return meta


def _memmap(cache_dir: str, prefix: str, meta: dict, name: str) -> np.memmap:
    return np.memmap(os.path.join(cache_dir, f"{prefix}_{name}.dat"), dtype=np.dtype(meta["array_dtypes"][name]), mode="r",
                     shape=tuple(meta["shapes"][name]))


def load_shard(cache_dir: str, prefix: str, rank: int = 0, world: int = 1, device="cpu",
               col_range: Optional[tuple] = None) -> CachedShard:
    """This rank's contiguous column shard (sizes as reference utils/dist_utils.py:53-57) straight onto `device`."""
    meta = read_meta(cache_dir, prefix)
    n, m = int(meta["num_sources"]), int(meta["num_destinations"])
    ccol_mm = _memmap(cache_dir, prefix, meta, "A_ccol")
    if ccol_mm.shape[0] != n + 1:
        raise ValueError(f"A_ccol has {ccol_mm.shape[0]} entries, expected {n + 1}")
    if col_range is None:
        sizes = shard_sizes(n, world)
        col_start = sum(sizes[:rank])
        col_end = col_start + sizes[rank]
    else:
        col_start, col_end = col_range
    if not (0 <= col_start <= col_end <= n):
        raise ValueError(f"column range [{col_start}, {col_end}) outside [0, {n}]")
    ccol = np.array(ccol_mm[col_start: col_end + 1])  # copies only this range out of the mapping
    e0, e1 = int(ccol[0]), int(ccol[-1])
    ccol -= ccol[0]
    row = np.array(_memmap(cache_dir, prefix, meta, "A_row")[e0:e1])
    a = np.array(_memmap(cache_dir, prefix, meta, "A_vals")[e0:e1], dtype=np.float32)
    c = np.array(_memmap(cache_dir, prefix, meta, "c_vals")[e0:e1], dtype=np.float32)
    np.negative(c, out=c)
    b = np.array(_memmap(cache_dir, prefix, meta, "b_vec"), dtype=np.float32)
    dev = torch.device(device)

    def put(x):
        t = torch.from_numpy(x)
        return t.to(dev, non_blocking=False) if dev.type != "cpu" else t

    return CachedShard(put(ccol), put(row), put(a), put(c), put(b), m, n, col_start, col_end)


def save_cache(cache_dir: str, prefix_args: dict, ccol: np.ndarray, row: np.ndarray, a: np.ndarray, c_positive: np.ndarray,
               b: np.ndarray) -> str:
    """Writes a cache entry the reference's loader accepts (generate_synthetic_data.py:289-342).  `c_positive` is the
    un-negated reward, as the reference stores it.  Returns the prefix."""
    prefix = cache_prefix(**prefix_args)
    os.makedirs(cache_dir, exist_ok=True)
    arrays = dict(A_ccol=ccol, A_row=row, A_vals=a, c_vals=c_positive, b_vec=b)
    for name, arr in arrays.items():
        mm = np.memmap(os.path.join(cache_dir, f"{prefix}_{name}.dat"), dtype=arr.dtype, mode="w+", shape=arr.shape)
        mm[...] = arr
        mm.flush()
        del mm
    meta = {
        "num_sources": int(prefix_args["num_sources"]), "num_destinations": int(prefix_args["num_destinations"]),
        "target_sparsity": float(prefix_args["target_sparsity"]), "dtype": str(prefix_args.get("dtype", torch.float32)),
        "seed": int(prefix_args.get("seed", 42)),
        "shapes": {k: list(v.shape) for k, v in arrays.items()},
        "array_dtypes": {k: str(v.dtype) for k, v in arrays.items()},
    }
    with open(os.path.join(cache_dir, f"{prefix}_meta.json"), "w") as fh:
        json.dump(meta, fh, indent=2)
    return prefix
