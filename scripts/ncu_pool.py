"""One launch of the tensor-core pooling kernel on the bench workload (16 384 scenes x 8 agents) -- the target of
`ncu --set full --import-source on -k regex:pool_fwd_tcx --launch-skip 1 -c 1`."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import socialways_b200 as sw
from socialways_b200 import ops

a, s = int(os.environ.get("SW_A", "8")), int(os.environ.get("SW_SCENES", "16384"))
n = a * s
gen = sw.Generator(use_social=True).cuda().requires_grad_(False)
pk = gen.packs()
g = torch.Generator(device="cuda").manual_seed(0)
h = torch.randn(n, 64, device="cuda", generator=g) * 0.3
xl = torch.rand(n, 4, device="cuda", generator=g)
scenes = gen.scene_index([(i * a, (i + 1) * a) for i in range(s)], n, torch.device("cuda"))
ub = torch.addmm(pk["pool_m0"], h, pk["pool_m"])
for _ in range(3):
    out = ops.pool_tcx(pk["pool"], pk["pool_tcx"], xl, h, ub, scenes)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    out = ops.pool_tcx(pk["pool"], pk["pool_tcx"], xl, h, ub, scenes)
e1.record()
torch.cuda.synchronize()
print(f"pool_tcx: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per launch, {n} agents x {a}", float(out.abs().max()))
