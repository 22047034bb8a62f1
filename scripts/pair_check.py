"""Bring-up check of the CTA-pair decode kernel (sw_decode_fwd_pair) on a GPU box: parity vs the fp32 oracle on small ragged
cases, agreement with the one-tile-per-SM kernel on the bench workload, and the CUDA-event time of both."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import socialways_b200 as sw
from golden_data import synthetic_scenes
from oracle import socialways_oracle as so


def small_cases():
    P = so.init_weights(seed=6)
    gen = sw.Generator(use_social=True)
    gen.load_state_dict({key: v for key, v in P.items() if not key.startswith("D.")})
    gen = gen.cuda().requires_grad_(False)
    for sizes, k in (([8] * 16, 1), ([5, 1, 32, 2, 9], 3), ([8] * 40, 7), ([6] * 36, 2), ([8] * 100, 20)):
        data = synthetic_scenes(sizes, seed=13)
        sc = so.IsoScale(data["obsvs"], data["preds"])
        obsv = torch.from_numpy(sc.normalize(data["obsvs"]))
        n = obsv.shape[0]
        torch.manual_seed(4)
        noise = torch.rand(k, n, 32)
        for social in (True, False):
            gen.use_social = social
            got = gen.predict_k(obsv.cuda(), noise.cuda(), 12, data["batches"], precision="fp16x2")
            old = gen.predict_k(obsv.cuda(), noise.cuda(), 12, data["batches"], precision="fp16x2s")
            torch.cuda.synchronize()
            if k * n <= 2000:
                want = torch.stack([so.predict(P, obsv, noise[i], 12, data["batches"], social, "closed") for i in range(k)])
                err = (got.cpu() - want).abs().max().item()
            else:
                err = float("nan")
            d = (got - old).abs().max().item()
            print(f"rows {k * n:6d} social={social}: |pair - oracle| = {err:.2e}  |pair - tcx| = {d:.2e}  overflow={gen.fp16_overflowed()}",
                  flush=True)


def bench_case(scenes=16384, agents=8, k=20, reps=5):
    P = so.init_weights(seed=0)
    gen = sw.Generator(use_social=True)
    gen.load_state_dict({key: v for key, v in P.items() if not key.startswith("D.")})
    gen = gen.cuda().requires_grad_(False)
    n = scenes * agents
    g = torch.Generator(device="cuda").manual_seed(1)
    h = torch.randn(n, 64, device="cuda", generator=g) * 0.3
    c = torch.randn(n, 64, device="cuda", generator=g) * 0.3
    pooled = torch.randn(n, 64, device="cuda", generator=g) * 0.3
    x_last = torch.randn(n, 4, device="cuda", generator=g) * 0.1
    noise = torch.rand(k, n, 32, device="cuda", generator=g)
    pk = gen.packs()
    from socialways_b200 import ops
    outs = {}
    for name, fn in (("tcx", lambda o: ops.decode_tcx(*pk["tcx"], h, c, pooled, noise, x_last, 12, out=o)),
                     ("pair", lambda o: ops.decode_pair(*pk["pair"], h, c, pooled, noise, x_last, 12, out=o))):
        out = torch.empty(k, n, 12, 4, device="cuda")
        fn(out)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
        ev[0].record()
        for i in range(reps):
            fn(out)
            ev[i + 1].record()
        torch.cuda.synchronize()
        ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(reps)]
        outs[name] = out
        print(f"{name}: {min(ms):.3f} ms (median {sorted(ms)[len(ms) // 2]:.3f})  -> {k * n / min(ms) / 1e3:.1f} M traj/s", flush=True)
    print("max |pair - tcx| on the bench workload:", (outs["pair"] - outs["tcx"]).abs().max().item(),
          " finite:", bool(torch.isfinite(outs["pair"]).all()))


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("small", "all"):
        small_cases()
    if what in ("bench", "all"):
        bench_case()
