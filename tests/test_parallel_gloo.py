"""CPU, world_size = 2 over gloo: the sharded residual (sums weighted by 1/global counts + ONE packed
all-reduce) reproduces the single-process result; the pmap-style dict mean of the trainer averages ranks."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import model as o_model, problems as o_prob, residuals as o_res, taylor as o_tay


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    from pde_inverse_problem_b200 import parallel
    shard = parallel.init_from_env("gloo")
    assert (shard.rank, shard.world) == (rank, world)
    d, T, gamma = 3, 2.0, 1.0
    g = torch.Generator().manual_seed(0)
    params = o_model.init_mlp_params(d, 32, 2)
    W, b = o_tay.unpack(params)
    data = {k: torch.randn(n, 2 * d, generator=g, dtype=torch.float64)
            for k, n in (("0T", 101), ("terminal", 37), ("initial", 64))}
    # each rank evaluates its shard with weights 1/global count (what the CUDA accumulate does per rank)
    loss = torch.zeros((), dtype=torch.float64)
    flat = torch.zeros(o_model.flatten_params(params).numel(), dtype=torch.float64)
    for name, al, be, cg in (("0T", -2.0, 2 * gamma, 1.0), ("terminal", 0.0, 2 / T, 0.0), ("initial", 0.0, -2 / T, 0.0)):
        lo, hi = shard.bounds(data[name].shape[0])
        z = data[name][lo:hi]
        val, dW, db, _, _ = o_tay.point_set(W, b, z[:, :d], [(z[:, d:], al, be)], 0.0, cg, 1.0 / data[name].shape[0])
        loss = loss + val
        flat = flat + torch.cat([torch.cat([w_.reshape(-1), b_.reshape(-1)]) for w_, b_ in zip(dW, db)])
    loss_r, flat_r = parallel.allreduce_sum_packed([loss.float(), flat.float()])
    mean = parallel.allreduce_mean_dict({"x": torch.tensor([float(rank)]), "g": torch.full((5,), 2.0 * rank)}, ["x", "g"])
    if rank == 0:
        q.put((loss_r.double(), flat_r.double(), mean["x"], mean["g"]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_residual_allreduce_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    loss_r, flat_r, mx, mg = q.get(timeout=240)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    # single-process truth from the autodiff oracle
    d, T = 3, 2.0
    g = torch.Generator().manual_seed(0)
    params = o_model.init_mlp_params(d, 32, 2)
    data = {k: torch.randn(n, 2 * d, generator=g, dtype=torch.float64)
            for k, n in (("0T", 101), ("terminal", 37), ("initial", 64))}
    pde = o_prob.KineticOUProblem(d, T=T)
    ref = o_res.kfp_value_and_grad_fn(o_model.mlp_apply, params, data, pde)
    gt = data["0T"][:, :d] @ pde.initial_configuration["tilde_F"].T
    ref_loss = ref["loss"] - (gt ** 2).sum(-1).mean()  # the workers leave out the constant |gV_true|^2 term
    assert abs(loss_r.item() - ref_loss.item()) < 1e-4 * abs(ref_loss.item())
    rg = o_model.flatten_params(ref["grad"])
    assert ((flat_r - rg).abs().max() / rg.abs().max()).item() < 1e-5
    assert mx.item() == pytest.approx(0.5) and torch.allclose(mg, torch.full((5,), 1.0))


def _meanfield_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import numpy as np
    from oracle import philox as o_philox
    from pde_inverse_problem_b200 import parallel
    shard = parallel.init_from_env("gloo")
    n, d, S, seed = 600, 3, 12, 99
    g = torch.Generator().manual_seed(1)
    z0 = torch.randn(n, 2 * d, generator=g, dtype=torch.float64) + 0.5
    lo, hi = shard.bounds(n)
    ids = np.arange(lo, hi)
    # what pdeip_meanfield_noise_sums accumulates on each rank: per-step sums of the Philox normals keyed by the GLOBAL
    # particle id, then sum q0, sum p0 — float64, reduced with ONE all-reduce (pipeline.HotPath.meanfield_table)
    sums = torch.zeros((S + 1) * d + 2 * d, dtype=torch.float64)
    for s in range(S + 1):
        sums[s * d:(s + 1) * d] = torch.as_tensor(o_philox.normals(seed, ids, s, d).sum(0))
    sums[(S + 1) * d:(S + 1) * d + d] = z0[lo:hi, :d].sum(0)
    sums[(S + 1) * d + d:] = z0[lo:hi, d:].sum(0)
    dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    if rank == 0:
        q.put(sums)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_meanfield_noise_sums_shard_and_close_world2():
    """The mean-field drift needs no per-step exchange: per-rank noise sums + ONE float64 all-reduce + the closed mean
    recursion reproduce the per-step empirical mean of the full interacting ensemble (rank-count invariant)."""
    import numpy as np
    from oracle import integrator as o_int, philox as o_philox
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_meanfield_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    sums = q.get(timeout=240)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    n, d, S, seed, dt, gamma = 600, 3, 12, 99, 0.05, 1.0
    g = torch.Generator().manual_seed(1)
    z0 = torch.randn(n, 2 * d, generator=g, dtype=torch.float64) + 0.5
    noise = torch.as_tensor(np.stack([o_philox.normals(seed, np.arange(n), s, d) for s in range(S + 1)], 1))
    A = torch.tensor([[2.0, 0.3, 0.0], [0.3, 1.5, 0.2], [0.0, 0.2, 1.0]], dtype=torch.float64)
    _, _, xbars, _ = o_int.interacting_langevin_scan(z0, S, dt, noise, A, gamma)
    xb = o_int.meanfield_mean_recursion(sums[: (S + 1) * d].view(S + 1, d), sums[(S + 1) * d:(S + 1) * d + d],
                                        sums[(S + 1) * d + d:], n, dt, gamma)
    assert (xb - xbars).abs().max().item() < 1e-12
