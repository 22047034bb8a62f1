"""Small invocation of every kernel added or changed in rounds 1-2, for compute-sanitizer:
    compute-sanitizer --tool memcheck python scripts/sanitize_target.py
tcx decode (TMA weights / noise tile, x-feedback K block, range guard), tensor-core pooling (unit table, scenes > 64),
tensor-core encoder, device noise, statistics kernels (1-NN, EMD cost, assignment), flat Adam, and one native training
iteration (pack kernels, fused D heads, tcgen05 + FFMA weight-gradient contractions, rows_linear, stats)."""
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

data = synthetic_scenes([8, 1, 5, 33, 2, 64, 7, 3, 70, 130], seed=0)
obsv = torch.from_numpy(data["obsvs"]).cuda() * 0.1
n = obsv.shape[0]
gen = sw.Generator(use_social=True).cuda().requires_grad_(False)
noise = torch.rand(3, n, 32, device="cuda")
out = gen.predict_k(obsv, noise, 12, data["batches"], precision="fp16x2")
assert torch.isfinite(out).all()
out2 = gen.predict_k(obsv, None, 12, data["batches"], precision="fp16x2", seed=3, k=2)
assert torch.isfinite(out2).all() and not gen.fp16_overflowed()
out3 = gen.predict_k(obsv, noise, 12, data["batches"], precision="bf16p")       # the bf16 build of the pair kernel
assert torch.isfinite(out3).all() and (out3 - out).abs().max().item() < 5e-2
from socialways_b200.trainer import SocialWaysTrainer
small = synthetic_scenes([3, 8, 1, 5, 6, 2, 7, 4, 9, 3, 2, 6], seed=1)
for tc in (True, False):
    tr = SocialWaysTrainer(small, batch_size=32, use_social=True, fused_adam=True)
    tr.native_tensor_cores = "force" if tc else False
    np.random.seed(0); torch.manual_seed(0)
    print("native epoch", tc, tr.train_native(verbose=False, use_graph=False))
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
