"""world_size-2 gloo tests (CPU) of the multi-GPU host logic (socialways_b200/distributed.py):
scene sharding, global-batch loss scaling and the flat-buffer gradient all-reduce.  The compute inside
each rank is the CPU oracle (test infrastructure); what is under test is that the sharded, summed
gradients and the post-Adam weights equal the single-process ones (SURVEY.md §8e parity rule)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from golden_data import synthetic_scenes


def test_shard_scenes_partitions_every_scene_once():
    from socialways_b200.distributed import shard_scenes
    rng = np.random.RandomState(0)
    for world in (1, 2, 3, 4, 8):
        sizes = rng.randint(1, 9, size=rng.randint(1, 40))
        offs = np.concatenate([[0], np.cumsum(sizes)]) + 17              # not rebased on purpose
        sb = np.stack([offs[:-1], offs[1:]], 1)
        seen, covered = 0, []
        for r in range(world):
            lo, hi, loc = shard_scenes(sb, world, r)
            if len(loc):
                assert loc[0, 0] == 0 and loc[-1, 1] == hi - lo
                assert np.all(loc[1:, 0] == loc[:-1, 1])
                covered.append((lo, hi))
            seen += len(loc)
        assert seen == len(sb)
        covered.sort()
        assert covered[0][0] == offs[0] and covered[-1][1] == offs[-1]
        assert all(a[1] == b[0] for a, b in zip(covered[:-1], covered[1:]))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _losses(so, P, obsv, pred, noise, zeros, ones, scenes, mse):
    o4, p4 = so.traj_4d(obsv, pred)
    with torch.no_grad():
        fake = so.predict(P, obsv, noise, 12, scenes, True, "closed")
    fl, fc = so.discriminator(P, o4, fake)
    rl, _ = so.discriminator(P, o4, p4)
    d_loss = mse(fl, zeros) + mse(rl, ones) + 0.5 * mse(fc, noise[:, :2])
    gen = so.predict(P, obsv, noise, 12, scenes, True, "closed")
    gl, gc = so.discriminator(P, o4, gen)
    g_loss = mse(gl, ones) + 0.5 * mse(gc, noise[:, :2])
    return d_loss, g_loss


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    from oracle import socialways_oracle as so
    from socialways_b200 import distributed as swdist
    W = so.init_weights(seed=1)
    P = {k: v.clone().requires_grad_(True) for k, v in W.items()}
    data = synthetic_scenes([3, 5, 2, 7, 1, 4], seed=2)
    sc = so.IsoScale(data["obsvs"], data["preds"])
    obsv, pred = torch.from_numpy(sc.normalize(data["obsvs"])), torch.from_numpy(sc.normalize(data["preds"]))
    bs = obsv.shape[0]
    torch.manual_seed(0)
    noise = torch.rand(bs, 32)
    zeros, ones = torch.full((bs, 1), 0.03), torch.full((bs, 1), 0.97)
    lo, hi, loc = swdist.shard_scenes(data["batches"], world, rank)
    mse = lambda a, b: swdist.global_mse(a, b, bs * (a.numel() // a.shape[0]))
    d_keys = [k for k in P if k.startswith("D.")]
    g_keys = [k for k in P if not k.startswith("D.")]
    opt_d = torch.optim.Adam([P[k] for k in d_keys], lr=1e-3)
    opt_g = torch.optim.Adam([P[k] for k in g_keys], lr=1e-4)
    d_loss, g_loss = _losses(so, P, obsv[lo:hi], pred[lo:hi], noise[lo:hi], zeros[lo:hi], ones[lo:hi], loc, mse)
    gd = torch.autograd.grad(d_loss, [P[k] for k in d_keys], allow_unused=True)
    gg = torch.autograd.grad(g_loss, [P[k] for k in g_keys], allow_unused=True)
    for k, g in zip(d_keys, gd):
        P[k].grad = g
    for k, g in zip(g_keys, gg):
        P[k].grad = g
    swdist.allreduce_grads([P[k] for k in d_keys], world)
    swdist.allreduce_grads([P[k] for k in g_keys], world)
    opt_d.step()
    opt_g.step()
    tot = swdist.allreduce_scalars([d_loss.item(), g_loss.item()], "cpu", world)
    if rank == 0:
        q.put(({k: P[k].grad.numpy().copy() for k in P}, {k: P[k].detach().numpy().copy() for k in P}, tot))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_step_equals_single_process():
    from oracle import socialways_oracle as so
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    grads2, weights2, tot2 = q.get(timeout=300)
    grads2 = {k: torch.from_numpy(v) for k, v in grads2.items()}
    weights2 = {k: torch.from_numpy(v) for k, v in weights2.items()}
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single process, plain nn.MSELoss semantics
    torch.set_num_threads(1)
    W = so.init_weights(seed=1)
    P = {k: v.clone().requires_grad_(True) for k, v in W.items()}
    data = synthetic_scenes([3, 5, 2, 7, 1, 4], seed=2)
    sc = so.IsoScale(data["obsvs"], data["preds"])
    obsv, pred = torch.from_numpy(sc.normalize(data["obsvs"])), torch.from_numpy(sc.normalize(data["preds"]))
    bs = obsv.shape[0]
    torch.manual_seed(0)
    noise = torch.rand(bs, 32)
    zeros, ones = torch.full((bs, 1), 0.03), torch.full((bs, 1), 0.97)
    d_loss, g_loss = _losses(so, P, obsv, pred, noise, zeros, ones, data["batches"], so.mse)
    d_keys = [k for k in P if k.startswith("D.")]
    g_keys = [k for k in P if not k.startswith("D.")]
    gd = torch.autograd.grad(d_loss, [P[k] for k in d_keys], allow_unused=True)
    gg = torch.autograd.grad(g_loss, [P[k] for k in g_keys], allow_unused=True)
    assert abs(tot2[0] - float(d_loss)) < 1e-6 and abs(tot2[1] - float(g_loss)) < 1e-6
    for k, g in list(zip(d_keys, gd)) + list(zip(g_keys, gg)):
        g = torch.zeros_like(P[k]) if g is None else g
        scale = max(1e-3, g.abs().max().item())
        assert (grads2[k] - g).abs().max().item() <= 2e-5 * scale + 1e-8, k
    for k, g in list(zip(d_keys, gd)) + list(zip(g_keys, gg)):
        P[k].grad = torch.zeros_like(P[k]) if g is None else g
    torch.optim.Adam([P[k] for k in d_keys], lr=1e-3).step()
    torch.optim.Adam([P[k] for k in g_keys], lr=1e-4).step()
    for k in P:
        # Adam's first step is lr * sign-like: a gradient flip near zero could move a weight by 2*lr
        assert (weights2[k] - P[k].detach()).abs().max().item() < 2.1e-3, k
        assert (weights2[k] - P[k].detach()).abs().mean().item() < 2e-5, k
