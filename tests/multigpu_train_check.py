"""Launched by torchrun (one rank per GPU): one epoch of the sharded training step on N ranks must
leave every rank with the weights a single-GPU run produces.  Rank 0 prints PASS/FAIL.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29517 tests/multigpu_train_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    from golden_data import synthetic_scenes
    from oracle import socialways_oracle as so
    from socialways_b200.trainer import SocialWaysTrainer
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    world, rank = dist.get_world_size(), dist.get_rank()
    rng = np.random.RandomState(3)
    data = synthetic_scenes(list(rng.randint(1, 9, size=60)), seed=8)
    W = so.init_weights(seed=2)

    # nccl | fused (eager, peer-memory all-reduce + Adam) | fused_graph | native (own-kernel iteration, graph replay)
    mode = os.environ.get("SW_CHECK_MODE", "nccl")
    fused, graph_n, native = mode != "nccl", mode == "fused_graph", mode == "native"
    epochs = 3 if (graph_n or native) else 1              # graph modes: epoch 1 eager, epoch 2 captures, epoch 3 replays

    def run(w, graph=False, fused=False, native=False):
        tr = SocialWaysTrainer(data, batch_size=64, use_social=True, n_unrolling_steps=1, weights=W,
                               device=f"cuda:{local}", world=w, cuda_graph=graph, fused_adam=fused)
        np.random.seed(5)
        torch.manual_seed(5)
        for _ in range(epochs):
            ade, fde = (tr.train_native if native else tr.train_graphed if graph else tr.train)(verbose=False)
        return tr.reference_weights(), ade, fde

    w_n, ade_n, fde_n = run((world, rank), graph=graph_n, fused=fused, native=native)
    ok = True
    # all ranks hold identical weights
    for k, v in w_n.items():
        ref = v.clone()
        dist.broadcast(ref, src=0)
        if not torch.equal(ref, v):
            ok = False
            print(f"rank {rank}: {k} differs from rank 0")
    if rank == 0:
        w_1, ade_1, fde_1 = run((1, 0))                   # single process, torch.optim.Adam, eager
        worst = max((w_n[k] - w_1[k]).abs().max().item() for k in w_1)
        mean = max((w_n[k] - w_1[k]).abs().mean().item() for k in w_1)
        print(f"mode {mode}, world {world}: max |w_N - w_1| = {worst:.3e}, max mean = {mean:.3e}, "
              f"ADE {ade_n:.6f} vs {ade_1:.6f}, FDE {fde_n:.6f} vs {fde_1:.6f}")
        ok = ok and mean < 2e-5 * epochs and abs(ade_n - ade_1) < 1e-4 and abs(fde_n - fde_1) < 1e-4
    flag = torch.tensor([1.0 if ok else 0.0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MULTIGPU_TRAIN_CHECK", "PASS" if flag.item() > 0 else "FAIL")
    dist.destroy_process_group()
    sys.exit(0 if flag.item() > 0 else 1)


if __name__ == "__main__":
    main()
