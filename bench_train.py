#!/usr/bin/env python
"""Secondary benchmark: the GAN TRAINING step (train(), reference train.py:439-560) on the toy set scaled
up (BASELINE.json configs[3]: 65 536 trajectories, n_per_batch = n_conditions so scenes keep 6 agents,
SURVEY.md D6), obs 2 / pred 2, use_social=True, unroll 1.  Reports agents/s of one epoch per batch size,
next to the CPU oracle port on a bounded sample.  Not the driver's contract (that is bench.py)."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def toy(n):
    from socialways_b200.toy import create_samples, pack_scenes
    np.random.seed(30)
    samples, ts = create_samples(n, 6, 3, n_per_batch=6)
    o, p, t, b = pack_scenes(samples, ts)
    return dict(obsvs=o, preds=p, times=t, batches=b)


def _iter_groups(tr):
    """The mini-batch grouping of train() (train.py:452-459): yields once per iteration of an epoch."""
    count = 0
    for ii, b in enumerate(tr.train_batches):
        count += int(b[1] - b[0])
        if ii >= tr.train_size - 1 or count + (tr.the_batches[ii + 1][1] - tr.the_batches[ii + 1][0]) > tr.batch_size:
            yield ii
            count = 0


def main_distributed(args):
    """torchrun: BASELINE config 4 -- the toy set scaled to 65 532 trajectories, scenes sharded over the ranks, one
    flat-buffer NCCL all-reduce per optimiser step (3 per iteration with unroll 1).  Rank 0 prints one JSON line."""
    import torch.distributed as dist
    from socialways_b200.trainer import SocialWaysTrainer
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    world, rank = dist.get_world_size(), dist.get_rank()
    data = toy(args.n)
    res = []
    modes = [("nccl_eager", dict()), ("fused_eager", dict(fused_adam=True)), ("fused_graph", dict(fused_adam=True, cuda_graph=True))]
    for bs in [int(x) for x in args.batch_sizes.split(",")]:
        rec = {"global_batch_size": bs}
        for name, kw in modes:
            tr = SocialWaysTrainer(data, batch_size=bs, use_social=True, n_unrolling_steps=1, device=f"cuda:{local}", **kw)
            step = tr.train_graphed if kw.get("cuda_graph") else tr.train
            np.random.seed(0)
            torch.manual_seed(0)
            try:
                for _ in range(2 if kw.get("cuda_graph") else 1):      # warm-up (graph mode: eager pass, then capture)
                    step(verbose=False)
                torch.cuda.synchronize()
                dist.barrier()
                t0 = time.perf_counter()
                for _ in range(args.epochs):
                    ade, fde = step(verbose=False)
                torch.cuda.synchronize()
                dist.barrier()
                dt = torch.tensor([(time.perf_counter() - t0) / args.epochs], device="cuda", dtype=torch.float64)
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
                iters = sum(1 for _ in _iter_groups(tr))
                rec[name] = {"epoch_s": dt.item(), "agents_per_s": tr.n_train_samples / dt.item(), "iterations_per_epoch": iters,
                             "ms_per_iteration": 1e3 * dt.item() / iters, "train_ade": ade, "train_fde": fde}
            except Exception as e:                                    # report, do not hide
                rec[name] = {"error": repr(e)[:300]}
            del tr
        res.append(rec)
    if rank == 0:
        print(json.dumps({"metric": "train_agents_per_sec", "n_gpus": world, "n_trajectories": args.n,
                          "parallelism": f"scenes of every mini-batch sharded x{world}; nccl_eager = flat-buffer NCCL all-reduce + "
                                         "torch Adam per optimiser step, fused_* = one peer-memory all-reduce+Adam kernel per "
                                         "optimiser step (csrc/flat_adam.cu), fused_graph = whole iteration replayed from a CUDA graph",
                          "results": res}))
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=65536 // 6 * 6)
    ap.add_argument("--batch-sizes", default="256,4096,49152")
    ap.add_argument("--epochs", type=int, default=2)
    ap.add_argument("--cpu-n", type=int, default=216)
    ap.add_argument("--graph", action="store_true", help="also time train_graphed() (CUDA-graph replay per batch shape)")
    args = ap.parse_args()
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        return main_distributed(args)
    from socialways_b200.trainer import SocialWaysTrainer
    data = toy(args.n)
    out = {"metric": "train_agents_per_sec", "n_trajectories": args.n, "results": []}
    for bs in [int(x) for x in args.batch_sizes.split(",")]:
        tr = SocialWaysTrainer(data, batch_size=bs, use_social=True, n_unrolling_steps=1)
        np.random.seed(0)
        torch.manual_seed(0)
        tr.train(verbose=False)                       # warm-up epoch
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.epochs):
            ade, fde = tr.train(verbose=False)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / args.epochs
        iters = len(tr.loss_log) // (args.epochs + 1)
        rec = {"batch_size": bs, "epoch_s": dt, "agents_per_s": tr.n_train_samples / dt,
               "iterations_per_epoch": iters, "ms_per_iteration": 1e3 * dt / iters, "train_ade": ade, "train_fde": fde}
        if args.graph:
            for key, kw in (("cuda_graph", dict()), ("cuda_graph_fused_adam", dict(fused_adam=True))):
                try:
                    tg = SocialWaysTrainer(data, batch_size=bs, use_social=True, n_unrolling_steps=1, cuda_graph=True, **kw)
                    np.random.seed(0)
                    torch.manual_seed(0)
                    tg.train_graphed(verbose=False)           # eager pass + capture
                    tg.train_graphed(verbose=False)
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    for _ in range(args.epochs):
                        gade, gfde = tg.train_graphed(verbose=False)
                    torch.cuda.synchronize()
                    gdt = (time.perf_counter() - t0) / args.epochs
                    rec[key] = {"epoch_s": gdt, "agents_per_s": tg.n_train_samples / gdt,
                                "ms_per_iteration": 1e3 * gdt / iters, "train_ade": gade, "train_fde": gfde,
                                "graphs": len(tg._graphs)}
                    del tg
                except Exception as e:                        # report, do not hide
                    rec[key] = {"error": repr(e)[:300]}
        out["results"].append(rec)
    # CPU oracle port of the same loop on a bounded sample
    from oracle import socialways_oracle as so
    torch.set_num_threads(os.cpu_count() or 1)
    cpu = so.OracleTrainer(so.init_weights(n_next=2), toy(args.cpu_n), batch_size=64, use_social=True, pool="loop")
    t0 = time.perf_counter()
    cpu.train_epoch()
    dt = time.perf_counter() - t0
    out["cpu_baseline"] = {"agents_per_s": cpu.n_train / dt, "cores": torch.get_num_threads(), "kind": "port",
                           "sample": f"toy {args.cpu_n}/6, batch 64, one epoch ({dt:.2f} s)"}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
