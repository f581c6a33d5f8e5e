"""World-size-2 test of the sharded path's HOST logic on CPU (gloo): contiguous column shards, local projection maps,
the single packed all-reduce (dualip_b200.objectives.matching.reduce_partials) and the replicated Maximizer update.
The per-shard arithmetic is supplied by the oracle here (the CUDA kernels are exercised by the -m gpu tests)."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import random_problem
from oracle import dualip_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


class _ShardedStandIn:
    """Same control flow as MatchingSolverDualObjectiveFunctionDistributed.calculate: local partial -> packed
    [grad | c.x | ||x||^2] -> ONE all-reduce -> m-length tail on every rank."""

    result_on_all_ranks = True
    equality_mask = None

    def __init__(self, prob, c0, c1, pm_local, gamma):
        e0, e1 = prob["ccol"][c0], prob["ccol"][c1]
        self.ccol = prob["ccol"][c0:c1 + 1] - e0
        self.row, self.a, self.c = prob["row"][e0:e1], prob["a"][e0:e1], prob["c"][e0:e1]
        self.b, self.m, self.pm, self.gamma = prob["b"], prob["n_rows"], pm_local, gamma
        self.collectives = 0

    def calculate(self, dual_val, gamma=None, save_primal=False, rank=0):
        from dualip_b200.objectives.matching import reduce_partials
        from dualip_b200.types import ObjectiveResult

        if gamma is not None:
            self.gamma = gamma
        lam = dual_val.numpy()
        pm = {k: O.ProjEntry(e.proj_type, e.proj_params, np.asarray(list(e.indices))) for k, e in self.pm.items()}
        loc = O.matching_calculate(self.ccol, self.row, self.a, self.c, self.m, pm, lam, self.gamma, None)
        xx = 2.0 * loc.reg_penalty / self.gamma
        packed = torch.from_numpy(np.concatenate([loc.dual_gradient, np.float32([loc.primal_objective, xx])]).astype(np.float32))
        reduce_partials(packed)
        self.collectives += 1
        s = packed.numpy()
        grad = s[: self.m] - self.b
        reg = 0.5 * self.gamma * float(s[self.m + 1])
        lg = float(np.dot(lam.astype(np.float64), grad.astype(np.float64)))
        return ObjectiveResult(dual_gradient=torch.from_numpy(grad.copy()), dual_objective=torch.tensor(float(s[self.m]) + reg + lg),
                               reg_penalty=torch.tensor(reg))


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from dualip_b200.optimizers.agd import AcceleratedGradientDescent
        from dualip_b200.projections import create_projection_map
        from dualip_b200.utils.dist_utils import global_to_local_projection_map, shard_sizes

        prob = random_problem(3, 501, 16, 5.0, scale_c=10.0)
        n = prob["n_cols"]
        pm = {**create_projection_map("simplex", {"z": 1.0}, n, indices=list(range(0, n, 2))),
              **create_projection_map("box", {"lower": 0.0, "upper": 1.0}, n, indices=list(range(1, n, 2)))}
        sizes = shard_sizes(n, world)
        c0 = sum(sizes[:rank])
        pm_local = global_to_local_projection_map(pm, range(c0, c0 + sizes[rank]))
        f = _ShardedStandIn(prob, c0, c0 + sizes[rank], pm_local, 2e-2)
        solver = AcceleratedGradientDescent(max_iter=25, gamma=2e-2, initial_step_size=1e-3, iteration_callback=lambda i, r: None)
        res = solver.maximize(f, torch.zeros(prob["n_rows"]), rank=rank)
        q.put((rank, res.dual_val.numpy(), res.dual_objective_log, f.collectives, sizes))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharded_solve_matches_unsplit_oracle():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    out.sort(key=lambda t: t[0])
    assert out[0][4] == [251, 250]  # reference split sizes: n//W (+1 for the first n%W ranks)
    # every rank holds the full result and took exactly one collective per iteration
    assert np.array_equal(out[0][1], out[1][1]) and out[0][2] == out[1][2]
    assert out[0][3] == 25 and out[1][3] == 25
    prob = random_problem(3, 501, 16, 5.0, scale_c=10.0)
    n = prob["n_cols"]
    pm = {"s": O.ProjEntry("simplex", {"z": 1.0}, np.arange(0, n, 2)), "b": O.ProjEntry("box", {"lower": 0.0, "upper": 1.0}, np.arange(1, n, 2))}

    def calc(lam, gamma):
        r = O.matching_calculate(prob["ccol"], prob["row"], prob["a"], prob["c"], prob["n_rows"], pm, lam, gamma, prob["b"])
        return r.dual_gradient, np.float32(r.dual_objective)

    y, obj_log, _, _ = O.agd_maximize(calc, np.zeros(prob["n_rows"], np.float32), 25, 2e-2, 1e-3, 0.1)
    assert np.allclose(out[0][2], obj_log, rtol=2e-5)
    assert np.allclose(out[0][1], y, rtol=1e-3, atol=1e-4)
