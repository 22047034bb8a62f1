"""Is the in-step decode time (bench.py) the sustained time of the kernel?  40 back-to-back launches of the pair decode on the bench
workload, per-launch times; then the same with a pass over the output (best-of-K metrics) between the launches, as in the bench step."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import socialways_b200 as sw
from socialways_b200 import ops
from oracle import socialways_oracle as so      # weight init only

P = so.init_weights(seed=0)
gen = sw.Generator(use_social=True)
gen.load_state_dict({k: v for k, v in P.items() if not k.startswith("D.")})
gen = gen.cuda().requires_grad_(False)
pk = gen.packs()
n, k, T = 131072, 20, 12
g = torch.Generator(device="cuda").manual_seed(0)
h = torch.randn(n, 64, device="cuda", generator=g) * 0.3
c = torch.randn(n, 64, device="cuda", generator=g) * 0.3
pooled = torch.randn(n, 64, device="cuda", generator=g) * 0.3
noise = torch.rand(k, n, 32, device="cuda", generator=g)
xl = torch.rand(n, 4, device="cuda", generator=g)
gt = torch.rand(n, T, 2, device="cuda", generator=g)
out = torch.empty(k, n, T, 4, device="cuda")
MODES = ("decode only", "decode + best-of-K pass", "decode only, 3 ms idle between launches")
for mode in (MODES[:1] if len(sys.argv) > 1 and sys.argv[1] == "short" else MODES):
    ev = []
    for i in range(40):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ops.decode_pair(*pk["pair"], h, c, pooled, noise, xl, T, out=out)
        b.record()
        ev.append((a, b))
        if mode.startswith("decode +"):
            ops.bestofk_metrics(out, gt, 1.0)
        if "idle" in mode:
            torch.cuda._sleep(int(3e-3 * 1.9e9))
    torch.cuda.synchronize()
    ms = [a.elapsed_time(b) for a, b in ev]
    print(mode, " ".join(f"{m:.2f}" for m in ms), flush=True)
