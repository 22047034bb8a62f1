"""Small invocation of every kernel added or changed in this round, for compute-sanitizer:
    compute-sanitizer --tool memcheck python scripts/sanitize_target.py
tcx decode (coalesced prologue, x-feedback K block), tensor-core pooling, tensor-core encoder, statistics kernels
(1-NN, EMD cost, assignment), flat Adam."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import socialways_b200 as sw
from socialways_b200 import statistics as st
from socialways_b200.fused_optim import FlatAdam
from golden_data import synthetic_scenes

data = synthetic_scenes([8, 1, 5, 33, 2, 64, 7, 3], seed=0)
obsv = torch.from_numpy(data["obsvs"]).cuda() * 0.1
n = obsv.shape[0]
gen = sw.Generator(use_social=True).cuda().requires_grad_(False)
noise = torch.rand(3, n, 32, device="cuda")
out = gen.predict_k(obsv, noise, 12, data["batches"], precision="fp16x2")
assert torch.isfinite(out).all()
rng = np.random.RandomState(0)
reals = rng.normal(0, 1, size=(9, 7, 14, 2)).astype(np.float32)
fakes = (reals[::-1] + rng.normal(0, 0.3, size=reals.shape)).astype(np.float32)
print(st.compute_1nn(reals, fakes), st.compute_wasserstein(reals, fakes))
ps = [torch.nn.Parameter(torch.randn(s, device="cuda")) for s in [(64, 4), (7,), (33, 3)]]
opt = FlatAdam(ps, lr=1e-3)
for _ in range(3):
    opt.zero_grad()
    for p in ps:
        p.grad.add_(torch.randn_like(p))
    opt.step()
torch.cuda.synchronize()
print("sanitize target ok", float(out.abs().max()), float(ps[0].abs().max()))
