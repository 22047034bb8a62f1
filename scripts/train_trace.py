"""Per-iteration trace of the native training step under torchrun (diagnostic for the multi-GPU train_step numbers):
    SW_TRACE_TRAIN=1 python -m torch.distributed.run --nproc-per-node N ... scripts/train_trace.py [global_batch]"""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch.distributed as dist
import bench
from socialways_b200.trainer import SocialWaysTrainer

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
bench.bind_to_gpu_numa(local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
gbs = int(sys.argv[1]) if len(sys.argv) > 1 else 49152
data = bench.toy_dataset(65536 // 6 * 6)
np.random.seed(0); torch.manual_seed(0)
tr = SocialWaysTrainer(data, batch_size=gbs, use_social=True, n_unrolling_steps=1, device=str(dev), fused_adam=True)
os.environ.pop("SW_TRACE_TRAIN", None)
for _ in range(3):
    tr.train_native(verbose=False, device_noise_seed=1)
os.environ["SW_TRACE_TRAIN"] = "1"
for ep in range(4):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    tr.train_native(verbose=False, device_noise_seed=1)
    torch.cuda.synchronize()
    if rank == 0:
        print(f"epoch {ep}: wall {1e3 * (time.perf_counter() - t0):.2f} ms", flush=True)
if world > 1:
    dist.destroy_process_group()
