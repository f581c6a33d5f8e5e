"""ProjectionMap: registry, ProjectionEntry and create_projection_map with the reference's names, signatures and key
format (reference src/dualip/projections/base.py:8-97).

Operators are thin descriptors: `native_class()` gives the row of the C-ABI class table that the fused kernel
uses, and `__call__` runs the same projection on a zero-padded [L x K] CUDA block through dualip_project_block.
"""
from __future__ import annotations

import ctypes
from abc import ABC, abstractmethod
from dataclasses import dataclass, field
from typing import Dict, List, Union

import numpy as np
import torch

from dualip_b200 import _native


@dataclass
class ProjectionEntry:
    proj_type: str = ""
    proj_params: dict[str, float] = field(default_factory=dict)
    indices: list[int] = field(default_factory=list)  # also accepted: range, numpy array, torch tensor


class ProjectionOperator(ABC):
    """Base class for projection operators (reference projections/base.py:15-36).

    Built-in operators describe themselves to the fused kernel through `native_class()`.  A user-registered operator
    written for the reference (it overrides `__init__` and `__call__` on a zero-padded [L x K] block and knows nothing
    about `native_class`) keeps working: its columns are routed through padded blocks on the device, the reference's
    apply_F_to_columns scheme (utils/sparse_utils.py:133-220), see objectives/matching.py."""

    @abstractmethod
    def __init__(self, **params):
        pass

    def native_class(self):
        """Row of the C-ABI class table (include/dualip_b200.h: dualip_proj_class), or None for an operator the fused
        kernel does not implement."""
        return None

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        """Project the columns of a zero-padded [L x K] block; does not modify x."""
        if not isinstance(x, torch.Tensor):
            raise TypeError("expected a torch.Tensor")
        if not x.is_cuda:
            raise RuntimeError("dualip_b200 projections run on CUDA tensors only (no CPU fallback)")
        if x.dtype != torch.float32:
            raise TypeError("dualip_b200 projections are float32-only")
        cls = self.native_class()
        if cls is None:
            raise NotImplementedError(f"{type(self).__name__} must override __call__ (it has no native class)")
        squeeze = x.ndim == 1
        if squeeze:
            x = x.unsqueeze(1)
        if x.ndim != 2:
            raise ValueError("expected a 1-D vector or an [L x K] block")
        xc = x.contiguous()
        out = torch.empty_like(xc)
        with torch.cuda.device(x.device):
            rc = _native.lib().dualip_project_block(xc.data_ptr(), out.data_ptr(), xc.shape[0], xc.shape[1],
                                                    ctypes.byref(cls), torch.cuda.current_stream().cuda_stream)
        _native.check(rc, "dualip_project_block")
        return out


_registry: dict[str, type] = {}


def register(name):
    def decorator(cls):
        _registry[name] = cls
        return cls

    return decorator


def project(name: str, **params) -> ProjectionOperator:
    """Instantiate a projection operator by name (reference projections/base.py:51-57)."""
    if name not in _registry:
        raise ValueError(f"Unknown projection operator '{name}'")
    return _registry[name](**params)


def create_projection_map(
    proj_type: str,
    proj_params: Dict[str, float],
    num_indices: int,
    indices: Union[List[int], range, np.ndarray, torch.Tensor, None] = None,
    key_prefix: str = "",
) -> Dict[str, ProjectionEntry]:
    """Same key format as the reference (projections/base.py:60-97).  `indices=None` means all of
    [0, num_indices); it is kept as a `range` instead of the reference's materialised list so that 10^8
    entities do not cost gigabytes of Python ints."""
    if indices is None:
        indices = range(num_indices)
    param_str = "_".join(f"{k}_{v}" for k, v in sorted(proj_params.items()))
    key = f"{key_prefix}{proj_type}_{param_str}" if key_prefix else f"{proj_type}_{param_str}"
    return {key: ProjectionEntry(proj_type=proj_type, proj_params=proj_params, indices=indices)}
