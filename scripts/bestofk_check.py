"""sw_bestofk_metrics: the wide-load kernel (32-byte aligned rows, n_next % 4 == 0) against the generic one (forced by a 16-byte
offset of the same data): bit-identical; timing on the bench shape (K = 20, 131 072 agents, 12 steps)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from socialways_b200 import ops
g = torch.Generator(device="cuda").manual_seed(0)
for k, n, t in ((20, 131072, 12), (3, 1000, 8), (5, 77, 16), (2, 50, 4), (4, 33, 6)):
    buf = torch.randn(k * n * t * 4 + 4, device="cuda", generator=g)
    pred_off = buf[4:].view(k, n, t, 4)                       # 16-byte offset -> generic kernel
    pred = pred_off.clone()                                   # aligned -> wide kernel (when t % 4 == 0)
    gt = torch.randn(n, t, 2, device="cuda", generator=g)
    a, b = ops.bestofk_metrics(pred, gt, 1.7), ops.bestofk_metrics(pred_off, gt, 1.7)
    print(k, n, t, "aligned ptr:", pred.data_ptr() % 32 == 0, pred_off.data_ptr() % 32, " bit-identical:", bool(torch.equal(a, b)))
k, n, t = 20, 131072, 12
pred = torch.randn(k, n, t, 4, device="cuda", generator=g); gt = torch.randn(n, t, 2, device="cuda", generator=g)
for name, p in (("wide", pred), ("generic", torch.randn(k * n * t * 4 + 4, device="cuda", generator=g)[4:].view(k, n, t, 4))):
    for _ in range(3): ops.bestofk_metrics(p, gt, 1.0)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(11)]
    ev[0].record()
    for i in range(10):
        ops.bestofk_metrics(p, gt, 1.0); ev[i + 1].record()
    torch.cuda.synchronize()
    print(name, min(ev[i].elapsed_time(ev[i + 1]) for i in range(10)) * 1e3, "us")
