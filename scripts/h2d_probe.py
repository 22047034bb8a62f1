"""Host->device bandwidth of pinned buffers on this box (explains the e2e line of bench.py: 128 B of caller-supplied
noise per trajectory cross PCIe every step)."""
import time, torch
dev = torch.device("cuda")
for mb in (16, 64, 335):
    n = mb * 1024 * 1024 // 4
    h = torch.empty(n, dtype=torch.float32).pin_memory()
    d = torch.empty(n, dtype=torch.float32, device=dev)
    for _ in range(3):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        d.copy_(h, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"H2D pinned {mb:4d} MB: {ms:.3f} ms  {mb * 1.048576 / ms:.1f} GB/s")
    e0.record()
    for _ in range(10):
        h.copy_(d, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"D2H pinned {mb:4d} MB: {ms:.3f} ms  {mb * 1.048576 / ms:.1f} GB/s")
