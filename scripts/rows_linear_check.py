"""sw_rows_linear: the 4-rows-per-thread path (>= 8192 rows) against the 1-row path on the same rows (bit-identical: same order of the
sum over k) and a float64 matmul; timing on the bench shape (131 072 x 64 -> 65)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from socialways_b200 import ops
g = torch.Generator(device="cuda").manual_seed(0)
for n, k, m in ((8192 + 77, 64, 65), (131072, 64, 65), (9000, 80, 80), (8192, 3, 1)):
    x = torch.randn(n, k, device="cuda", generator=g)
    w = torch.randn(k, m, device="cuda", generator=g)
    b = torch.randn(m, device="cuda", generator=g)
    a1 = torch.randn(n, m, device="cuda", generator=g)
    big = ops.rows_linear(x, w, b, a1)
    small = torch.cat([ops.rows_linear(x[i:i + 4096].contiguous(), w, b, a1[i:i + 4096].contiguous()) for i in range(0, n, 4096)])
    ref = (x.double() @ w.double() + b.double() + a1.double())
    print(n, k, m, "bit-identical to the 1-row path:", bool(torch.equal(big, small)), " max err vs f64:", float((big - ref).abs().max()))
x = torch.randn(131072, 64, device="cuda", generator=g); w = torch.randn(64, 65, device="cuda", generator=g); b = torch.randn(65, device="cuda", generator=g)
out = torch.empty(131072, 65, device="cuda")
for _ in range(3): ops.rows_linear(x, w, b, out=out)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(11)]
ev[0].record()
for i in range(10):
    ops.rows_linear(x, w, b, out=out); ev[i + 1].record()
torch.cuda.synchronize()
print("131072 x 64 -> 65:", min(ev[i].elapsed_time(ev[i + 1]) for i in range(10)) * 1e3, "us")
