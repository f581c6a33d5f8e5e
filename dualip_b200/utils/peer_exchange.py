"""Exchange windows of the sharded path: the packed partial sums are read from the peers' memory over NVLink inside the
update kernel (include/dualip_b200.h, dualip_peer_* / dualip_agd_step_peer) instead of going through a collective call.

The reference reduces with three dist.reduce + a barrier and broadcasts the iterate twice per iteration
(src/dualip/objectives/matching.py:272-277, optimizers/agd.py:204-206).  Here torch.distributed only carries the 64-byte
CUDA IPC handles at setup time; per iteration there is no collective call at all.
"""
from __future__ import annotations

import ctypes
import os
import socket
from typing import Optional, Sequence

import torch
import torch.distributed as dist

from dualip_b200 import _native


class PeerExchange:
    """This rank's exchange window plus the mapped windows of its peers."""

    def __init__(self, m: int, rank: int, world: int, device: torch.device):
        self.m, self.rank, self.world, self.device = m, rank, world, torch.device(device)
        self.lib = _native.lib()
        self.handle = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _native.check(self.lib.dualip_peer_create(ctypes.byref(self.handle), m, rank, world, self.device.index),
                          "dualip_peer_create")

    # -- wiring -------------------------------------------------------------------------------------------
    def export_handle(self) -> bytes:
        buf = (ctypes.c_uint8 * _native.PEER_HANDLE_BYTES)()
        _native.check(self.lib.dualip_peer_export(self.handle, buf), "dualip_peer_export")
        return bytes(buf)

    def connect_ipc(self, handles: Sequence[bytes]) -> None:
        blob = b"".join(handles)
        assert len(blob) == self.world * _native.PEER_HANDLE_BYTES
        buf = (ctypes.c_uint8 * len(blob)).from_buffer_copy(blob)
        with torch.cuda.device(self.device):
            _native.check(self.lib.dualip_peer_connect_ipc(self.handle, buf), "dualip_peer_connect_ipc")

    @property
    def window(self) -> int:
        return self.lib.dualip_peer_window(self.handle)

    @staticmethod
    def connect_local(exchanges: Sequence["PeerExchange"]) -> None:
        """Wires windows that live in ONE process (several shards per process; tests on a single GPU)."""
        ptrs = (ctypes.c_void_p * len(exchanges))(*[e.window for e in exchanges])
        for e in exchanges:
            _native.check(e.lib.dualip_peer_connect_ptrs(e.handle, ptrs), "dualip_peer_connect_ptrs")

    @classmethod
    def over_process_group(cls, m: int, device: torch.device, group=None, enabled: bool = True) -> Optional["PeerExchange"]:
        """Collective over `group`: allocates a window per rank and maps every peer's.  Returns None (on EVERY rank) when the
        ranks are not all on one host, the world is larger than the window's flag array, DUALIP_PEER_EXCHANGE=0, or any
        rank fails to map a peer window; the caller then keeps the NCCL all-reduce."""
        if not (dist.is_available() and dist.is_initialized()):
            return None
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        if world < 2 or world > _native.PEER_MAX_WORLD:
            return None  # the same on every rank: no need to agree on it
        want = enabled and os.environ.get("DUALIP_PEER_EXCHANGE", "1") != "0"
        device = torch.device(device)
        # setup-time exchange of small Python objects: works over any backend (NCCL, or gloo when several ranks share a GPU)
        with torch.cuda.device(device):
            every = [None] * world
            dist.all_gather_object(every, (bool(want), socket.gethostname()), group=group)
            if not all(w for w, _ in every) or len({h for _, h in every}) != 1:
                return None
            ex, ok, handle = None, True, b""
            try:
                ex = cls(m, rank, world, device)
                handle = ex.export_handle()
            except Exception:
                ok = False
            handles = [None] * world
            dist.all_gather_object(handles, (ok, handle), group=group)
            ok = all(o for o, _ in handles)
            if ok:
                try:
                    ex.connect_ipc([h for _, h in handles])
                except Exception:
                    ok = False
            oks = [None] * world
            dist.all_gather_object(oks, ok, group=group)
            if not all(oks):
                if ex is not None:
                    ex.close()
                return None
        return ex

    # -- per step -----------------------------------------------------------------------------------------
    def next_slot(self) -> int:
        """Device address the partial sums of the upcoming step must be written to."""
        return self.lib.dualip_peer_next_slot(self.handle)

    def status(self) -> int:
        out = ctypes.c_int32(0)
        with torch.cuda.device(self.device):
            _native.check(self.lib.dualip_peer_status(self.handle, ctypes.byref(out),
                                                      torch.cuda.current_stream(self.device).cuda_stream), "dualip_peer_status")
        return int(out.value)

    def status_nowait(self) -> int:
        """The time-out flag as the host sees it right now (mapped host memory; no synchronisation)."""
        return int(self.lib.dualip_peer_status_nowait(self.handle))

    def close(self) -> None:
        if self.handle:
            self.lib.dualip_peer_destroy(self.handle)
            self.handle = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
