"""Entity sharding helpers with the reference's names and return shapes (src/dualip/utils/dist_utils.py:9-71)."""
from __future__ import annotations

from typing import Sequence

import numpy as np
import torch

from dualip_b200.projections.base import ProjectionEntry
from dualip_b200.utils.sparse_utils import split_csc_by_cols


def shard_sizes(num_cols: int, num_shards: int) -> list[int]:
    """n//W columns per shard, the first n%W shards take one more (reference dist_utils.py:53-57)."""
    base, rem = divmod(num_cols, num_shards)
    return [base + (1 if r < rem else 0) for r in range(num_shards)]


def global_to_local_projection_map(global_map: dict[str, ProjectionEntry], local_cols) -> dict[str, ProjectionEntry]:
    """Restrict a projection map to a shard and renumber its columns from 0.

    `local_cols` is the shard's list (or range) of global column ids, as returned by split_tensors_to_devices.  For a
    contiguous shard and `range`/array indices this is an interval intersection; no per-column dictionary is built
    (the reference's dict over every column is infeasible at 10^8 entities).  Entries without local columns are
    dropped, and list inputs give list outputs, like the reference (dist_utils.py:9-25)."""
    if isinstance(local_cols, range) and local_cols.step == 1:
        lo, hi, contiguous = local_cols.start, local_cols.stop, True
    else:
        arr = np.asarray(local_cols, dtype=np.int64)
        contiguous = arr.size > 0 and bool(np.all(np.diff(arr) == 1))
        lo, hi = (int(arr[0]), int(arr[-1]) + 1) if contiguous else (0, 0)
    local_map: dict[str, ProjectionEntry] = {}
    for key, entry in global_map.items():
        ind = entry.indices
        if contiguous and isinstance(ind, range) and ind.step == 1:
            a, b = max(ind.start, lo), min(ind.stop, hi)
            local = range(a - lo, b - lo) if b > a else range(0)
        elif contiguous:
            g = ind.cpu().numpy() if isinstance(ind, torch.Tensor) else np.asarray(ind, dtype=np.int64)
            sel = g[(g >= lo) & (g < hi)] - lo
            local = sel.tolist() if isinstance(ind, list) else sel
        else:
            position = {int(g): k for k, g in enumerate(np.asarray(local_cols).tolist())}
            local = [position[int(g)] for g in (ind.tolist() if hasattr(ind, "tolist") else ind) if int(g) in position]
        if len(local):
            local_map[key] = ProjectionEntry(proj_type=entry.proj_type, proj_params=entry.proj_params, indices=local)
    return local_map


def split_tensors_to_devices(a_mat: torch.Tensor, c_mat: torch.Tensor, compute_devices: Sequence) -> tuple:
    """Balanced contiguous column split of A and c across devices; returns (A shards, c shards, per-shard global
    column ids) like the reference (dist_utils.py:28-71)."""
    if a_mat.layout != torch.sparse_csc or c_mat.layout != torch.sparse_csc:
        raise ValueError("Both A and B must be CSC-format sparse tensors")
    num_cols = a_mat.size(1)
    if not compute_devices:
        return [a_mat], [c_mat], list(range(num_cols))
    sizes = shard_sizes(num_cols, len(compute_devices))
    index_map, start = [], 0
    for size in sizes:
        index_map.append(list(range(start, start + size)))
        start += size
    a_parts = [blk.to(dev) for blk, dev in zip(split_csc_by_cols(a_mat, sizes), compute_devices)]
    c_parts = [blk.to(dev) for blk, dev in zip(split_csc_by_cols(c_mat, sizes), compute_devices)]
    return a_parts, c_parts, index_map
