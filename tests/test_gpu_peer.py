"""-m gpu: the peer-memory exchange (dualip_peer_* / dualip_agd_step_peer) on ONE GPU: two shards of a problem, two
optimizer states and two exchange windows live in one process, wired by pointer (dualip_peer_connect_ptrs); each "rank"
runs on its own stream, so the two update kernels wait for each other exactly as two processes on two GPUs would.  The
two-process version over CUDA IPC runs in tests/dist_worker.py (needs 2 GPUs)."""
import ctypes
import os

import numpy as np
import pytest
import torch

from conftest import random_problem
from dualip_b200 import _native
from dualip_b200.objectives.matching import MatchingInputArgs, MatchingSolverDualObjectiveFunction
from dualip_b200.optimizers.agd import AcceleratedGradientDescent
from dualip_b200.projections import create_projection_map
from dualip_b200.utils.dist_utils import global_to_local_projection_map, split_tensors_to_devices
from dualip_b200.utils.peer_exchange import PeerExchange
from test_gpu_parity import DEV, _csc

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("CUDA_LAUNCH_BLOCKING", "0") not in ("", "0"),
                                 reason="the ranks' update kernels must run concurrently (they wait for each other)")]
N_SCAL = len(_native.SCALAR_FIELDS)


class _Rank:
    def __init__(self, objective, m, initial_step, max_step):
        self.obj = objective
        self.stream = torch.cuda.Stream(device=DEV)
        self.agd = ctypes.c_void_p()
        lib = _native.lib()
        _native.check(lib.dualip_agd_create(ctypes.byref(self.agd), m, 0, None, None, initial_step, max_step, 15))
        _native.check(lib.dualip_agd_reserve_log(self.agd, 64))
        self.x_ptr = lib.dualip_agd_x(self.agd)
        self.grad = torch.empty(m, device=DEV)
        self.scal = torch.zeros(N_SCAL, dtype=torch.float64, device=DEV)
        self.partial = torch.empty(m + 2, device=DEV)

    def logs(self, n):
        o, s = (ctypes.c_double * n)(), (ctypes.c_double * n)()
        _native.check(_native.lib().dualip_agd_read_log(self.agd, n, o, s, torch.cuda.current_stream().cuda_stream))
        return np.array(o[:n]), np.array(s[:n])

    def dual(self):
        y = torch.empty_like(self.grad)
        _native.check(_native.lib().dualip_agd_get(self.agd, None, y.data_ptr(), torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        return y


@pytest.mark.parametrize("world,n_rows", [(1, 96), (2, 96), (3, 96), (2, 9001), (3, 4095)])
def test_peer_step_matches_summed_partials(world, n_rows, monkeypatch):
    """n_rows 9001 / 4095: the update kernel runs as 3 / 2 CTAs (one float4 per thread and peer), the last one to finish
    takes the step; 96: a single CTA."""
    monkeypatch.setenv("DUALIP_PEER_TIMEOUT_MS", "3000")
    lib = _native.lib()
    p = random_problem(23, 6001 if n_rows == 96 else 30011, n_rows, 8.0, scale_c=10.0, lam_scale=0.5)
    n, m, gamma, iters = p["n_cols"], p["n_rows"], 2e-2, 40
    A, C = _csc(p)
    b = torch.from_numpy(p["b"]).to(DEV)
    pm = create_projection_map("simplex", {"z": 1.0}, n)
    a_s, c_s, index_map = split_tensors_to_devices(A, C, [DEV] * world)
    beta = AcceleratedGradientDescent(max_iter=iters, gamma=gamma).beta_seq.tolist()

    def shards():
        return [MatchingSolverDualObjectiveFunction(
            MatchingInputArgs(a_s[k], c_s[k], global_to_local_projection_map(pm, index_map[k]), None), gamma) for k in range(world)]

    # (a) peer path: every rank on its own stream, no host-side reduction
    ranks = [_Rank(o, m, 1e-3, 0.1) for o in shards()]
    ex = [PeerExchange(m, k, world, torch.device(DEV)) for k in range(world)]
    PeerExchange.connect_local(ex)
    torch.cuda.synchronize()
    for i in range(iters):
        decay = 1 if (i + 1) % 9 == 0 else 0
        g_i = gamma * 0.5 ** (i // 9)
        for k, r in enumerate(ranks):  # all shard kernels first: they never wait, the update kernels do
            with torch.cuda.stream(r.stream):
                r.obj.launch_partial(r.x_ptr, g_i, ex[k].next_slot())
        for k, r in enumerate(ranks):
            with torch.cuda.stream(r.stream):
                _native.check(lib.dualip_agd_step_peer(r.agd, ex[k].handle, b.data_ptr(), g_i, r.grad.data_ptr(), r.scal.data_ptr(),
                                                       beta[i], decay, 0.5, i, r.stream.cuda_stream))
    torch.cuda.synchronize()
    assert [e.status() for e in ex] == [0] * world
    peer_logs = [r.logs(iters) for r in ranks]
    peer_duals = [r.dual() for r in ranks]
    for k in range(1, world):  # replicas add the same numbers in the same order: bit-identical state
        assert np.array_equal(peer_logs[k][0], peer_logs[0][0]) and np.array_equal(peer_logs[k][1], peer_logs[0][1])
        assert torch.equal(peer_duals[k], peer_duals[0])

    # (a') the same exchange fused into the shard kernel: ONE launch per rank and iteration (dualip_matching_ascent_step_peer)
    fused = [_Rank(o, m, 1e-3, 0.1) for o in shards()]
    ex2 = [PeerExchange(m, k, world, torch.device(DEV)) for k in range(world)]
    PeerExchange.connect_local(ex2)
    torch.cuda.synchronize()
    for i in range(iters):
        decay = 1 if (i + 1) % 9 == 0 else 0
        g_i = gamma * 0.5 ** (i // 9)
        for k, r in enumerate(fused):
            with torch.cuda.stream(r.stream):
                r.obj.launch_ascent_step_peer(r.agd, ex2[k].handle, b.data_ptr(), g_i, r.grad.data_ptr(), r.scal.data_ptr(),
                                              beta[i], decay, 0.5, i)
    torch.cuda.synchronize()
    assert [e.status() for e in ex2] == [0] * world and [e.status_nowait() for e in ex2] == [0] * world
    fused_logs = [r.logs(iters) for r in fused]
    fused_duals = [r.dual() for r in fused]
    for k in range(1, world):
        assert np.array_equal(fused_logs[k][0], fused_logs[0][0]) and np.array_equal(fused_logs[k][1], fused_logs[0][1])
        assert torch.equal(fused_duals[k], fused_duals[0])
    # against the two-launch path: the same sums in the same order; only the block reductions of the step run with a different
    # thread count (double precision partial sums in another order)
    # (a last-bit difference in a Lipschitz ratio changes a step by one ulp; the ascent carries it along)
    assert np.allclose(fused_logs[0][0], peer_logs[0][0], rtol=1e-6) and np.allclose(fused_logs[0][1], peer_logs[0][1], rtol=1e-5)
    assert torch.allclose(fused_duals[0], peer_duals[0], rtol=1e-4, atol=1e-6)
    for r in fused:
        lib.dualip_agd_destroy(r.agd)
    for e in ex2:
        e.close()

    # (b) what a collective would do: partial sums added in rank order on the host side, then dualip_agd_step_sharded
    ref = _Rank(None, m, 1e-3, 0.1)
    objs = shards()
    parts = [torch.empty(m + 2, device=DEV) for _ in range(world)]
    st = torch.cuda.current_stream().cuda_stream
    for i in range(iters):
        decay = 1 if (i + 1) % 9 == 0 else 0
        g_i = gamma * 0.5 ** (i // 9)
        for k, o in enumerate(objs):
            o.launch_partial(ref.x_ptr, g_i, parts[k].data_ptr())
        total = parts[0].clone()
        for k in range(1, world):
            total += parts[k]
        _native.check(lib.dualip_agd_step_sharded(ref.agd, total.data_ptr(), b.data_ptr(), g_i, ref.grad.data_ptr(),
                                                  ref.scal.data_ptr(), beta[i], decay, 0.5, i, st))
    torch.cuda.synchronize()
    ref_obj, ref_step = ref.logs(iters)
    if all(o.plan_info()["fixed_point"] == 1 for o in objs):
        # fixed-point shard sums are order-independent, and both paths add the shards in rank order: same bits
        assert np.array_equal(peer_logs[0][0], ref_obj) and np.array_equal(peer_logs[0][1], ref_step)
        assert torch.equal(peer_duals[0], ref.dual())
    else:  # fp32 atomics inside a shard: summation order varies from launch to launch
        assert np.allclose(peer_logs[0][0], ref_obj, rtol=1e-5) and np.allclose(peer_logs[0][1], ref_step, rtol=3e-2)
        assert torch.allclose(peer_duals[0], ref.dual(), rtol=1e-3, atol=1e-3)
    for r in ranks + [ref]:
        lib.dualip_agd_destroy(r.agd)
    for e in ex:
        e.close()


def test_peer_wait_is_bounded(monkeypatch):
    """A rank whose peer never arrives sets the status word after the time-out instead of hanging."""
    monkeypatch.setenv("DUALIP_PEER_TIMEOUT_MS", "200")
    lib = _native.lib()
    m = 64
    ex = [PeerExchange(m, k, 2, torch.device(DEV)) for k in range(2)]
    PeerExchange.connect_local(ex)
    r = _Rank(None, m, 1e-3, 0.1)
    b = torch.zeros(m, device=DEV)
    torch.cuda.synchronize()
    _native.check(lib.dualip_agd_step_peer(r.agd, ex[0].handle, b.data_ptr(), 1e-2, r.grad.data_ptr(), r.scal.data_ptr(), 0.0, 0, 1.0, 0,
                                           torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert ex[0].status() == 1 and ex[1].status() == 0
    with pytest.raises(ValueError):
        PeerExchange(m, 0, _native.PEER_MAX_WORLD + 1, torch.device(DEV))
    lib.dualip_agd_destroy(r.agd)


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("push,grid", [("1", "0"), ("0", "0"), ("1", "1")])
def test_sharded_evaluation_through_peer_memory_without_a_step(world, push, grid, monkeypatch):
    """dualip_matching_calc_peer: shard kernel + exchange + m-length tail in one launch, no optimizer step (the host-buffer
    path of the sharded objective, reference matching.py:247-307).  Every rank obtains the same bits, and they equal the
    unsharded evaluation up to the order of the shard sums; a second call reuses the other slot parity."""
    monkeypatch.setenv("DUALIP_PEER_TIMEOUT_MS", "3000")
    monkeypatch.setenv("DUALIP_PEER_PUSH", push)
    monkeypatch.setenv("DUALIP_GRID_TAIL", grid)  # "1": every CTA takes a slice of the exchange and of the tail (grid_tail.cuh)
    lib = _native.lib()
    p = random_problem(31, 5003, 200, 8.0, scale_c=10.0, lam_scale=0.5)
    n, m, gamma = p["n_cols"], p["n_rows"], 2e-2
    A, C = _csc(p)
    b = torch.from_numpy(p["b"]).to(DEV)
    pm = create_projection_map("simplex", {"z": 1.0}, n)
    whole = MatchingSolverDualObjectiveFunction(MatchingInputArgs(A, C, pm, b), gamma)
    a_s, c_s, index_map = split_tensors_to_devices(A, C, [DEV] * world)
    objs = [MatchingSolverDualObjectiveFunction(MatchingInputArgs(a_s[k], c_s[k], global_to_local_projection_map(pm, index_map[k]), None), gamma)
            for k in range(world)]
    ex = [PeerExchange(m, k, world, torch.device(DEV)) for k in range(world)]
    PeerExchange.connect_local(ex)
    streams = [torch.cuda.Stream(device=DEV) for _ in range(world)]
    grads = [torch.empty(m, device=DEV) for _ in range(world)]
    scals = [torch.zeros(N_SCAL, dtype=torch.float64, device=DEV) for _ in range(world)]
    rng = np.random.default_rng(0)
    for call in range(3):
        lam = torch.from_numpy((rng.random(m) * 0.5).astype(np.float32)).to(DEV)
        torch.cuda.synchronize()
        for k in range(world):
            with torch.cuda.stream(streams[k]):
                _native.check(lib.dualip_matching_calc_peer(objs[k]._plan, ex[k].handle, lam.data_ptr(), b.data_ptr(), gamma,
                                                            grads[k].data_ptr(), scals[k].data_ptr(), streams[k].cuda_stream))
        torch.cuda.synchronize()
        assert [e.status() for e in ex] == [0] * world
        ref = whole.calculate(lam)
        for k in range(world):
            assert torch.equal(grads[k], grads[0]) and torch.equal(scals[k], scals[0]), "ranks must obtain identical bits"
        assert torch.allclose(grads[0], ref.dual_gradient, rtol=1e-5, atol=1e-5)
        assert abs(float(scals[0][0]) - float(ref.scalars64[0])) <= 1e-6 * abs(float(ref.scalars64[0]))
    for e in ex:
        e.close()


@pytest.mark.parametrize("world,shared_tail", [(2, 0), (4, 1)])
def test_sharded_launches_of_four_or_more_ranks_share_the_tail(world, shared_tail, monkeypatch):
    """Size rule of the all-CTA tail in sharded launches (csrc/calc.cu launch_eval): from m = 8192 on, launches of 4 or more
    ranks let every CTA store a slice of the sums into the peers' windows and add a slice of the W slots (measured on 8 GPUs:
    2935 -> 3135 it/s); 2 ranks keep the last-CTA tail.  Both for the evaluation-only launch of the host-buffer path and for the
    fused step, against the unsharded evaluation and against the same run with the rule switched off.  DUALIP_CTAS=32 keeps
    the `world` grids co-resident on the one GPU of this test (the grid-wide barriers need that; one process per GPU has it
    by construction)."""
    monkeypatch.setenv("DUALIP_PEER_TIMEOUT_MS", "3000")
    monkeypatch.setenv("DUALIP_CTAS", "32")
    monkeypatch.setenv("DUALIP_REBALANCE", "0")
    monkeypatch.delenv("DUALIP_GRID_TAIL", raising=False)
    lib = _native.lib()
    p = random_problem(37, 6000, 8200, 8.0, scale_c=10.0, lam_scale=0.5)
    n, m, gamma, iters = p["n_cols"], p["n_rows"], 2e-2, 12
    A, C = _csc(p)
    b = torch.from_numpy(p["b"]).to(DEV)
    pm = create_projection_map("simplex", {"z": 1.0}, n)
    whole = MatchingSolverDualObjectiveFunction(MatchingInputArgs(A, C, pm, b), gamma)
    a_s, c_s, index_map = split_tensors_to_devices(A, C, [DEV] * world)
    beta = AcceleratedGradientDescent(max_iter=iters, gamma=gamma).beta_seq.tolist()

    def shards():
        return [MatchingSolverDualObjectiveFunction(
            MatchingInputArgs(a_s[k], c_s[k], global_to_local_projection_map(pm, index_map[k]), None), gamma) for k in range(world)]

    def exchanges():
        ex = [PeerExchange(m, k, world, torch.device(DEV)) for k in range(world)]
        PeerExchange.connect_local(ex)
        return ex

    # evaluation only: dualip_matching_calc_peer
    objs, ex = shards(), exchanges()
    assert all(o.plan_info()["n_ctas"] <= 32 and o.plan_info()["grid_tail"] == 0 for o in objs)
    streams = [torch.cuda.Stream(device=DEV) for _ in range(world)]
    grads = [torch.empty(m, device=DEV) for _ in range(world)]
    scals = [torch.zeros(N_SCAL, dtype=torch.float64, device=DEV) for _ in range(world)]
    lam = torch.from_numpy(p["lam"]).to(DEV)
    torch.cuda.synchronize()
    for _ in range(2):
        for k in range(world):
            with torch.cuda.stream(streams[k]):
                _native.check(lib.dualip_matching_calc_peer(objs[k]._plan, ex[k].handle, lam.data_ptr(), b.data_ptr(), gamma,
                                                            grads[k].data_ptr(), scals[k].data_ptr(), streams[k].cuda_stream))
        torch.cuda.synchronize()
    assert [e.status() for e in ex] == [0] * world
    assert [o.plan_info()["last_launch_grid_tail"] for o in objs] == [shared_tail] * world
    ref = whole.calculate(lam)
    assert whole.plan_info()["last_launch_grid_tail"] == 0
    for k in range(world):
        assert torch.equal(grads[k], grads[0]) and torch.equal(scals[k], scals[0]), "ranks must obtain identical bits"
    assert torch.allclose(grads[0], ref.dual_gradient, rtol=1e-5, atol=1e-5)
    assert abs(float(scals[0][0]) - float(ref.scalars64[0])) <= 1e-6 * abs(float(ref.scalars64[0]))
    for e in ex:
        e.close()

    # fused step: dualip_matching_ascent_step_peer, rule on (default) and off
    runs = {}
    for tag in ("rule", "off"):
        if tag == "off":
            monkeypatch.setenv("DUALIP_GRID_TAIL", "0")
        ranks, ex = [_Rank(o, m, 1e-3, 0.1) for o in shards()], exchanges()
        torch.cuda.synchronize()
        for i in range(iters):
            for k, r in enumerate(ranks):
                with torch.cuda.stream(r.stream):
                    r.obj.launch_ascent_step_peer(r.agd, ex[k].handle, b.data_ptr(), gamma, r.grad.data_ptr(), r.scal.data_ptr(),
                                                  beta[i], 0, 1.0, i)
        torch.cuda.synchronize()
        assert [e.status() for e in ex] == [0] * world
        assert [r.obj.plan_info()["last_launch_grid_tail"] for r in ranks] == [shared_tail if tag == "rule" else 0] * world
        logs, duals = [r.logs(iters) for r in ranks], [r.dual() for r in ranks]
        for k in range(1, world):
            assert np.array_equal(logs[k][0], logs[0][0]) and torch.equal(duals[k], duals[0]), "replicas must stay bit-identical"
        runs[tag] = (logs[0], duals[0])
        for r in ranks:
            lib.dualip_agd_destroy(r.agd)
        for e in ex:
            e.close()
    assert np.allclose(runs["rule"][0][0], runs["off"][0][0], rtol=1e-9) and np.allclose(runs["rule"][0][1], runs["off"][0][1], rtol=1e-6)
    assert torch.allclose(runs["rule"][1], runs["off"][1], rtol=1e-5, atol=1e-7)
