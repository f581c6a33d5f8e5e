"""Worker for tests/test_gpu_distributed.py: run under torchrun, one process per GPU (NCCL).

Reference counterpart: tests/distributed/test_matching_distributed.py:116-195 (same 5x5 problem and golden values, sharded
across world_size GPUs), plus a random problem checked against the C oracle on every rank."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    rank, world, local_rank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    one_gpu = os.environ.get("DUALIP_TEST_ONE_GPU") == "1"
    if one_gpu:
        # every rank on cuda:0 (a single-GPU box): NCCL refuses two ranks per device, so the process group is gloo; the
        # per-iteration exchange still goes through CUDA-IPC windows and the in-kernel flags, the path under test
        local_rank = 0
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if one_gpu:
        dist.init_process_group("gloo")
    else:
        dist.init_process_group("nccl", device_id=dev)

    def all_gather(t):
        if one_gpu:  # gloo gathers host tensors
            out = [torch.empty_like(t, device="cpu") for _ in range(world)]
            dist.all_gather(out, t.cpu())
            return out
        out = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return out

    from conftest import random_problem
    from dualip_b200.objectives.matching import MatchingInputArgs, MatchingSolverDualObjectiveFunctionDistributed
    from dualip_b200.optimizers.agd import AcceleratedGradientDescent
    from dualip_b200.projections import create_projection_map
    from dualip_b200.run_solver import run_solver
    from dualip_b200.types import ComputeArgs, ObjectiveArgs, SolverArgs
    from dualip_b200.utils.dist_utils import global_to_local_projection_map, split_tensors_to_devices
    from oracle import c_oracle
    from test_oracle_golden import _scala_5x5

    # 1) the reference's distributed known-answer test
    ccol, row, a, c, b = _scala_5x5()
    A = torch.sparse_csc_tensor(torch.from_numpy(ccol), torch.from_numpy(row), torch.from_numpy(a), size=(5, 5))
    C = torch.sparse_csc_tensor(torch.from_numpy(ccol), torch.from_numpy(row), torch.from_numpy(c), size=(5, 5))
    pm = create_projection_map("simplex", {"z": 1}, 5)
    a_s, c_s, index_map = split_tensors_to_devices(A, C, ["cpu"] * world)
    local = MatchingInputArgs(a_s[rank].to(dev), c_s[rank].to(dev), global_to_local_projection_map(pm, index_map[rank]), None, None)
    f = MatchingSolverDualObjectiveFunctionDistributed(local_matching_input_args=local, b_vec=torch.from_numpy(b), gamma=1e-3, host_device=dev)
    solver = AcceleratedGradientDescent(max_iter=30, gamma=1e-3, iteration_callback=lambda i, r: None)
    res = solver.maximize(f, 0.1 * torch.ones(5, device=dev), rank=rank)
    for i, true_val in [(2, -3.6010155991401818), (16, -3.60842718733725), (23, -3.5080258013053136), (29, -3.4868496294227143)]:
        assert abs(res.dual_objective_log[i - 1] - true_val) < 1e-5, (rank, i, res.dual_objective_log[i - 1])

    # 2) random problem through run_solver(compute_device_num=world): every rank gets the full result
    p = random_problem(77, 4001, 64, 8.0, scale_c=10.0)
    n, m, gamma = p["n_cols"], p["n_rows"], 2e-2
    A = torch.sparse_csc_tensor(torch.from_numpy(p["ccol"]), torch.from_numpy(p["row"]), torch.from_numpy(p["a"]), size=(m, n))
    C = torch.sparse_csc_tensor(torch.from_numpy(p["ccol"]), torch.from_numpy(p["row"]), torch.from_numpy(p["c"]), size=(m, n))
    args = MatchingInputArgs(A, C, create_projection_map("simplex", {"z": 1.0}, n), torch.from_numpy(p["b"]))
    out = run_solver(args, SolverArgs(max_iter=15, gamma=gamma, initial_step_size=1e-3), ComputeArgs(f"cuda:{local_rank}", world),
                     ObjectiveArgs("matching"))
    lam = out.dual_val.clone()
    ref_first = c_oracle.calculate(p["ccol"], p["row"], p["a"], p["c"], m, [c_oracle.make_class("simplex", {"z": 1.0})],
                                   np.zeros(m, np.float32), gamma, p["b"])
    assert abs(out.dual_objective_log[0] - ref_first["scal"][0]) <= 1e-5 * abs(ref_first["scal"][0])
    gathered = all_gather(lam)
    assert all(torch.equal(g, gathered[0]) for g in gathered), "replicated optimizer state diverged across ranks"
    # 3) the same shard through both exchange paths: partial sums read from peer memory inside the update kernel (no
    #    per-iteration callback) vs the NCCL all-reduce (DUALIP_PEER_EXCHANGE=0)
    from dualip_b200.optimizers.agd import no_iteration_callback

    a_s, c_s, index_map = split_tensors_to_devices(A, C, ["cpu"] * world)
    runs = {}
    for tag, env in (("peer", "1"), ("nccl", "0")):
        os.environ["DUALIP_PEER_EXCHANGE"] = env
        local = MatchingInputArgs(a_s[rank].to(dev), c_s[rank].to(dev), global_to_local_projection_map(args.projection_map, index_map[rank]), None, None)
        f = MatchingSolverDualObjectiveFunctionDistributed(local_matching_input_args=local, b_vec=args.b_vec, gamma=gamma, host_device=dev)
        solver = AcceleratedGradientDescent(max_iter=40, gamma=gamma, initial_step_size=1e-3, max_step_size=0.1,
                                            gamma_decay_type="step", gamma_decay_params={"decay_steps": 9, "decay_factor": 0.5},
                                            iteration_callback=no_iteration_callback)
        runs[tag] = solver.maximize(f, torch.zeros(m, device=dev), rank=rank)
        assert (f._peer is not None) == (tag == "peer"), f"exchange path {tag}: peer windows {'missing' if tag == 'peer' else 'unexpected'}"
        again = solver.__class__(max_iter=5, gamma=gamma, iteration_callback=no_iteration_callback).maximize(f, runs[tag].dual_val, rank=rank)
        assert len(again.dual_objective_log) == 5  # a second loop on the same objective keeps the windows' step counter
    assert np.allclose(runs["peer"].dual_objective_log, runs["nccl"].dual_objective_log, rtol=1e-6)
    assert np.allclose(runs["peer"].step_size_log, runs["nccl"].step_size_log, rtol=1e-3)
    assert torch.allclose(runs["peer"].dual_val, runs["nccl"].dual_val, rtol=1e-4, atol=1e-5)
    lam = runs["peer"].dual_val.clone()
    gathered = all_gather(lam)
    assert all(torch.equal(g, gathered[0]) for g in gathered), "peer path: replicas must be bit-identical"
    # 4) host-buffer evaluation of the sharded objective (lambda on the host): shard kernel + exchange through peer memory +
    #    tail in one native call (dualip_matching_calc_peer_host) against the device-resident evaluation over NCCL
    os.environ["DUALIP_PEER_EXCHANGE"] = "1"
    local = MatchingInputArgs(a_s[rank].to(dev), c_s[rank].to(dev), global_to_local_projection_map(args.projection_map, index_map[rank]), None, None)
    f = MatchingSolverDualObjectiveFunctionDistributed(local_matching_input_args=local, b_vec=args.b_vec, gamma=gamma, host_device=dev)
    lam_host = runs["peer"].dual_val.cpu()
    r_host = f.calculate(lam_host)
    assert f._peer is not None and r_host.dual_gradient.device.type == "cpu"
    r_dev = f.calculate(runs["peer"].dual_val)  # device tensors: partial kernel, all-reduce, epilogue kernel
    assert torch.allclose(r_host.dual_gradient, r_dev.dual_gradient.cpu(), rtol=1e-5, atol=1e-5)
    assert abs(float(r_host.dual_objective) - float(r_dev.dual_objective)) <= 1e-5 * abs(float(r_dev.dual_objective))
    g_all = all_gather(r_host.dual_gradient.to(dev))
    assert all(torch.equal(g, g_all[0]) for g in g_all), "host-buffer path: every rank must obtain the same bits"
    host_run = AcceleratedGradientDescent(max_iter=12, gamma=gamma, initial_step_size=1e-3, max_step_size=0.1,
                                          gamma_decay_type="step", gamma_decay_params={"decay_steps": 9, "decay_factor": 0.5},
                                          iteration_callback=lambda i, r: None).maximize(f, torch.zeros(m), rank=rank)
    assert np.allclose(host_run.dual_objective_log, runs["peer"].dual_objective_log[:12], rtol=1e-5)
    # 5) sharded CUDA-graph replay: chunks of 8 iterations (exchange inside the captured launches) against single launches
    os.environ["DUALIP_REBALANCE"] = "0"
    outs = {}
    for tag, graph in (("graph", "1"), ("single", "0")):
        os.environ["DUALIP_GRAPH"], os.environ["DUALIP_GRAPH_CHUNK"] = graph, "8"
        local = MatchingInputArgs(a_s[rank].to(dev), c_s[rank].to(dev), global_to_local_projection_map(args.projection_map, index_map[rank]), None, None)
        f = MatchingSolverDualObjectiveFunctionDistributed(local_matching_input_args=local, b_vec=args.b_vec, gamma=gamma, host_device=dev)
        outs[tag] = AcceleratedGradientDescent(max_iter=37, gamma=gamma, initial_step_size=1e-3, max_step_size=0.1,
                                               gamma_decay_type="step", gamma_decay_params={"decay_steps": 9, "decay_factor": 0.5},
                                               iteration_callback=no_iteration_callback).maximize(f, torch.zeros(m, device=dev), rank=rank)
    assert outs["graph"].dual_objective_log == outs["single"].dual_objective_log and torch.equal(outs["graph"].dual_val, outs["single"].dual_val)
    dist.barrier()
    if rank == 0:
        print("DIST_WORKER_OK", world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
