"""Probe of the native training iteration: eager launches (for an ncu launch list) or graph replay timings.

    python scripts/train_native_probe.py --batch 4096 --iters 6 [--graph]
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--n", type=int, default=65532)
    ap.add_argument("--epochs", type=int, default=2)
    ap.add_argument("--graph", action="store_true")
    ap.add_argument("--shape", default="toy", choices=["toy", "eth"])
    ap.add_argument("--force-tc", action="store_true", help="tcgen05 contractions even for tiny problems")
    ap.add_argument("--ffma-contract", action="store_true", help="weight-gradient contractions on the FFMA kernel instead of tcgen05")
    args = ap.parse_args()
    from bench import toy_dataset
    from socialways_b200.trainer import SocialWaysTrainer
    if args.shape == "toy":
        data = toy_dataset(args.n)
    else:
        from golden_data import synthetic_scenes
        data = synthetic_scenes([8] * (args.n // 8), seed=1)
    tr = SocialWaysTrainer(data, batch_size=args.batch, use_social=True, n_unrolling_steps=1, fused_adam=True)
    tr.native_tensor_cores = False if args.ffma_contract else ("force" if args.force_tc else True)
    iters = sum(1 for _ in tr._minibatches())
    np.random.seed(0)
    torch.manual_seed(0)
    for _ in range(3 if args.graph else 1):
        tr.train_native(verbose=False, use_graph=args.graph)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.epochs):
        tr.train_native(verbose=False, use_graph=args.graph)
    e1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    if args.graph:      # pure GPU time of one iteration: replay the largest captured graph back to back (no host RNG / copies between)
        ent = max(tr._native_steps.values(), key=lambda e: e["step"].bs)
        if ent["graph"] is not None:
            for _ in range(3):
                ent["graph"].replay()
            torch.cuda.synchronize()
            r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            r0.record()
            for _ in range(20):
                ent["graph"].replay()
            r1.record()
            torch.cuda.synchronize()
            print(f"   graph replay only (bs {ent['step'].bs}): {r0.elapsed_time(r1) / 20:.4f} ms/iteration")
    print(f"batch {args.batch} shape {args.shape} graph {args.graph} contraction {'ffma' if args.ffma_contract else 'tcgen05'}: {iters} iterations/epoch, "
          f"{e0.elapsed_time(e1) / args.epochs / iters:.4f} ms/iteration (device), {1e3 * wall / args.epochs / iters:.4f} ms (wall)")


if __name__ == "__main__":
    main()
