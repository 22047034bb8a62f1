"""Split the decode kernel time into per-tile overhead (prologue + hoist) and per-step cost: T(n_next) = P + n_next * S.
    python scripts/tcx_step_cost.py [tcx|pair]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import socialways_b200 as sw
from socialways_b200 import ops
from oracle import socialways_oracle as so

P = so.init_weights(seed=0)
gen = sw.Generator(use_social=True)
gen.load_state_dict({k: v for k, v in P.items() if not k.startswith("D.")})
gen = gen.cuda().requires_grad_(False)
pk = gen.packs()
n, k = 131072, 20
h = torch.randn(n, 64, device="cuda") * 0.3
c = torch.randn(n, 64, device="cuda") * 0.3
pooled = torch.randn(n, 64, device="cuda") * 0.3
noise = torch.rand(k, n, 32, device="cuda")
xl = torch.rand(n, 4, device="cuda")
which = sys.argv[1] if len(sys.argv) > 1 else "pair"
run = (lambda T, out: ops.decode_pair(*pk["pair"], h, c, pooled, noise, xl, T, out=out)) if which == "pair" else \
      (lambda T, out: ops.decode_tcx(*pk["tcx"], h, c, pooled, noise, xl, T, out=out))
res = {}
for T in (1, 2, 6, 12, 24):
    out = torch.empty(k, n, T, 4, device="cuda")
    for _ in range(2):
        run(T, out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        run(T, out)
    e1.record()
    torch.cuda.synchronize()
    res[T] = e0.elapsed_time(e1) / 5
    print(f"n_next={T:3d}: {res[T]:.3f} ms")
S = (res[24] - res[12]) / 12
print(f"per-step {S:.3f} ms, per-tile overhead {res[12] - 12 * S:.3f} ms ({100 * (res[12] - 12 * S) / res[12]:.1f} % of the 12-step kernel)")
