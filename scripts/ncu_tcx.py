"""One launch of the tcx decode kernel on the profile workload (655 360 trajectories, 12 steps) -- the target of
`ncu --set full --import-source on -k regex:decode_fwd_ --launch-skip 1 -c 1`."""
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
n, k, T = 32768, 20, 12
g = torch.Generator(device="cuda").manual_seed(0)
h = torch.randn(n, 64, device="cuda", generator=g) * 0.3
c = torch.randn(n, 64, device="cuda", generator=g) * 0.3
pooled = torch.randn(n, 64, device="cuda", generator=g) * 0.3
noise = torch.rand(k, n, 32, device="cuda", generator=g)
xl = torch.rand(n, 4, device="cuda", generator=g)
out = torch.empty(k, n, T, 4, device="cuda")
which = sys.argv[1] if len(sys.argv) > 1 else "tcx"       # "tcx" (one tile per SM), "pair" (CTA pairs, two tiles per SM) or "bf16p" (pair kernel, bf16 operands)
for _ in range(3):
    if which == "bf16p":
        ops.decode_pair(*pk["pair_bf16"], h, c, pooled, noise, xl, T, out=out, bf16=True)
    elif which == "pair":
        ops.decode_pair(*pk["pair"], h, c, pooled, noise, xl, T, out=out)
    else:
        ops.decode_tcx(*pk["tcx"], h, c, pooled, noise, xl, T, out=out)
torch.cuda.synchronize()
print("ok", float(out.abs().max()))
