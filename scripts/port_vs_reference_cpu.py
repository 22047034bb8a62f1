"""Build-container only (needs /root/reference): time the UNMODIFIED reference test(K) loop (train.py:563-616, lifted by
tests/golden/reference_harness.py) and the oracle port that bench.py's CPU arm runs, on the same ETH-shaped scenes and the
same host threads -- shows that the `kind: "port"` baseline is not slower than the reference it stands in for."""
import contextlib, io, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import reference_harness as rh
from oracle import socialways_oracle as so

n_scenes, A, K = int(os.environ.get("SCENES", "96")), 8, 20
torch.set_num_threads(os.cpu_count() or 1)
data = rh.synthetic_scenes([A] * n_scenes, seed=1234)
# the reference keeps 4/5 of the scenes for training: hand it 5/4 of the scenes so that test() sees n_scenes... simpler: time
# its test() over whatever it holds out and normalise by the trajectories it actually predicted
ref = rh.Reference(data, batch_size=64, use_social=True, weight_seed=0)
n_test_agents = sum(int(b - a) for a, b in ref.test_batches)
torch.manual_seed(0)
with contextlib.redirect_stdout(io.StringIO()):
    ref.test(2)                                             # warm-up
    t0 = time.perf_counter()
    ref.test(K)
    t_ref = time.perf_counter() - t0
ref_rate = n_test_agents * K / t_ref

P = so.init_weights(seed=0)
sc = so.IsoScale(data["obsvs"], data["preds"])
obsv = torch.from_numpy(sc.normalize(data["obsvs"]))
pred = torch.from_numpy(sc.normalize(data["preds"]))
batches = [(int(a), int(b)) for a, b in ref.test_batches]


def port():
    with torch.no_grad():
        for a, b in batches:
            errs = []
            for k in range(K):
                noise = torch.rand(b - a, 32)
                hat = so.predict(P, obsv[a:b], noise, 12, None, use_social=True, pool="loop")
                errs.append((((hat[:, :, :2] - pred[a:b]) / sc.sx) ** 2).sum(dim=2).sqrt())
            e = torch.stack(errs)
            _ = (e.mean(2).min(0)[0].sum().item(), e[:, :, -1].min(0)[0].sum().item())


port()
t0 = time.perf_counter()
port()
t_port = time.perf_counter() - t0
print(f"threads {torch.get_num_threads()}: reference test({K}) {ref_rate:.0f} traj/s ({t_ref:.2f} s, {n_test_agents} agents), "
      f"oracle port {n_test_agents * K / t_port:.0f} traj/s ({t_port:.2f} s) -> port/reference = {t_ref / t_port:.2f}x")
