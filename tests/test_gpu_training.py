"""GPU parity of the TRAINING path: forward+backward kernels behind torch.autograd vs (a) gradients
recorded from the unmodified reference (tests/golden/ops_*.npz) and (b) whole train()/test() runs of
the reference from the same seeds (tests/golden/train_*.npz): every mse_loss value it evaluated, its
printed ADE/FDE lines, post-training weights, K-sample test metrics (the 1e-4 bar of north_star)."""
import numpy as np
import pytest
import torch

from conftest import golden_weights, load_golden
from golden_data import case_data

pytestmark = pytest.mark.gpu


def _modules(P, use_social):
    import socialways_b200 as sw
    gen = sw.Generator(use_social=use_social)
    gen.load_state_dict({k: v for k, v in P.items() if not k.startswith("D.")})
    D = sw.Discriminator(12, 64, 2)
    D.load_state_dict({k[2:]: v for k, v in P.items() if k.startswith("D.")})
    return gen.cuda(), D.cuda()


def _check_grad(name, got, want, rel=2e-4):
    scale = max(1e-3, float(np.abs(want).max()))
    err = float(np.abs(got - want).max())
    assert err <= rel * scale + 1e-7, f"{name}: max abs err {err:.3e} vs scale {scale:.3e}"


@pytest.mark.parametrize("case", ["ops_ragged.npz", "ops_zara.npz"])
@pytest.mark.parametrize("social", [True, False])
def test_gradients_vs_reference_golden(case, social):
    import socialways_b200 as sw
    g = load_golden(case)
    tag = "soc" if social else "nos"
    gen, D = _modules(golden_weights(g), social)
    obsv, pred, noise = (torch.from_numpy(g[k]).cuda() for k in ("obsv", "pred", "noise"))
    n = obsv.shape[0]
    o4, p4 = sw.get_traj_4d(obsv, pred)
    mse = torch.nn.MSELoss()
    hat = gen.predict(obsv, noise, 12, g["scenes"])
    assert hat.requires_grad
    np.testing.assert_allclose(hat.detach().cpu().numpy(), g[f"{tag}.pred_hat"], atol=2e-5, rtol=0)
    lab, code = D(o4, hat)
    g_loss = mse(lab, torch.full((n, 1), 0.95, device="cuda")) + 0.5 * mse(code.squeeze(), noise[:, :2])
    assert abs(g_loss.item() - float(g[f"{tag}.g_loss"])) < 5e-6
    g_loss.backward()
    for k, p in gen.named_parameters():
        want = g[f"{tag}.ggrad.{k}"]
        got = p.grad.cpu().numpy() if p.grad is not None else np.zeros_like(want)
        _check_grad(k, got, want)
    D.zero_grad()
    fl, fc = D(o4, hat.detach())
    rl, _ = D(o4, p4)
    d_loss = mse(fl, torch.full((n, 1), 0.05, device="cuda")) + mse(rl, torch.full((n, 1), 0.95, device="cuda")) + \
        0.5 * mse(fc.squeeze(), noise[:, :2])
    assert abs(d_loss.item() - float(g[f"{tag}.d_loss"])) < 5e-6
    d_loss.backward()
    for k, p in D.named_parameters():
        _check_grad("D." + k, p.grad.cpu().numpy(), g[f"{tag}.dgrad.D.{k}"])


def test_gradients_vs_oracle_multi_tile_toy_shapes():
    """Toy shapes (obs 2 / pred 2), 216 agents = 7 tiles with a ragged tail, scenes of 6."""
    import socialways_b200 as sw
    from oracle import socialways_oracle as so
    W = so.init_weights(seed=3, n_next=2)
    d = so.toy_samples(216, 6)
    sc = so.IsoScale(d["obsvs"], d["preds"])
    obsv = torch.from_numpy(sc.normalize(d["obsvs"]))
    torch.manual_seed(5)
    noise = torch.rand(216, 32)
    P = {k: v.clone().requires_grad_(True) for k, v in W.items()}
    hat = so.predict(P, obsv, noise, 2, d["batches"], True, "closed")
    lab, code = so.discriminator(P, so.traj_4d(obsv), hat)
    loss = so.mse(lab, torch.full((216, 1), 0.9)) + 0.5 * so.mse(code.squeeze(), noise[:, :2])
    loss.backward()
    gen = sw.Generator(use_social=True)
    gen.load_state_dict({k: v for k, v in W.items() if not k.startswith("D.")})
    D = sw.Discriminator(2, 64, 2)
    D.load_state_dict({k[2:]: v for k, v in W.items() if k.startswith("D.")})
    gen, D = gen.cuda(), D.cuda()
    hat2 = gen.predict(obsv.cuda(), noise.cuda(), 2, d["batches"])
    lab2, code2 = D(sw.get_traj_4d(obsv.cuda(), []), hat2)
    mse = torch.nn.MSELoss()
    loss2 = mse(lab2, torch.full((216, 1), 0.9, device="cuda")) + 0.5 * mse(code2.squeeze(), noise[:, :2].cuda())
    assert abs(loss2.item() - loss.item()) < 5e-6
    loss2.backward()
    for k, p in gen.named_parameters():
        want = P[k].grad.numpy() if P[k].grad is not None else np.zeros(tuple(p.shape), np.float32)
        _check_grad(k, p.grad.cpu().numpy(), want)


@pytest.mark.parametrize("case", ["train_toy_216.npz", "train_ragged.npz", "train_unroll0.npz"])
@pytest.mark.parametrize("social", [True, False])
def test_training_epochs_vs_reference_golden(case, social, capsys):
    from oracle import socialways_oracle as so
    from socialways_b200.trainer import SocialWaysTrainer
    g = load_golden(case)
    tag = "soc" if social else "nos"
    data = so.toy_samples(216, 6) if case == "train_toy_216.npz" else case_data(case)
    seed = int(g["seed"][0])
    tr = SocialWaysTrainer(data, batch_size=int(g["batch_size"]), use_social=social,
                           n_unrolling_steps=int(g["unroll"]), weights=golden_weights(g, "w0."))
    np.random.seed(seed)
    torch.manual_seed(seed)
    for ep in range(1, int(g["epochs"]) + 1):
        tr.epoch = ep
        tr.train()
    rng_state = torch.get_rng_state()
    tr.generator.inference_precision = "fp32"          # FFMA decode: the printed line matches character for character
    tr.test(int(g["k_test"]))
    lines = [l for l in capsys.readouterr().out.splitlines() if l.strip()]
    ref_lines = [str(s) for s in g[f"{tag}.stdout"]]
    assert [r.split(" | time")[0] for r in ref_lines[:-1]] == [l.split(" | time")[0] for l in lines[:-1]]
    assert ref_lines[-1] == lines[-1]
    # default tensor-core decode (fp16 hi/lo split): same noise stream, every printed number within one unit of
    # the last printed digit (a 1e-6 arithmetic difference may flip the %.3f rounding)
    import re
    torch.set_rng_state(rng_state)
    tr.generator.inference_precision = "fp16x2"
    tr.test(int(g["k_test"]))
    tc_line = [l for l in capsys.readouterr().out.splitlines() if l.strip()][-1]
    nums = lambda t: [float(x) for x in re.findall(r"-?\d+\.\d+", t)]
    assert len(nums(tc_line)) == len(nums(ref_lines[-1])) == 4
    assert all(abs(a - b) <= 1.001e-3 for a, b in zip(nums(tc_line), nums(ref_lines[-1])))
    ref_mse = g[f"{tag}.mse_values"]
    per_iter = 3 * (int(g["unroll"]) + 1) + 3
    assert len(ref_mse) == per_iter * len(tr.loss_log)
    for it, rec in enumerate(tr.loss_log):
        row = ref_mse[it * per_iter:(it + 1) * per_iter]
        assert abs(rec["d_fake"] - row[-6]) < 1e-5 and abs(rec["d_info"] - row[-5]) < 1e-5
        assert abs(rec["d_real"] - row[-4]) < 1e-5
        assert abs(rec["g_fool"] - row[-2]) < 1e-5 and abs(rec["g_info"] - row[-1]) < 1e-5
    W = tr.reference_weights()
    names = [str(s) for s in g[f"{tag}.w1_names"]]
    for i, k in enumerate(names):
        assert abs(W[k].double().sum().item() - g[f"{tag}.w1_sum"][i]) < 5e-4, k
        assert abs(W[k].double().norm().item() - g[f"{tag}.w1_l2"][i]) < 5e-4, k
    if social and "w1.encoder.embed.weight" in g:
        for k, v in golden_weights(g, "w1.").items():
            assert (W[k].cpu() - v).abs().max().item() < 5e-5, k
    torch.manual_seed(seed + 1)
    m = tr.test(int(g["k_test"]), verbose=False)
    got = np.array([m["ade_avg"], m["fde_avg"], m["ade_min"], m["fde_min"]])
    np.testing.assert_allclose(got, g[f"{tag}.test_metrics"], atol=1e-4, rtol=0)


def test_checkpoint_keys_match_reference():
    from oracle import socialways_oracle as so
    from socialways_b200.trainer import SocialWaysTrainer
    tr = SocialWaysTrainer(so.toy_samples(216, 6), batch_size=64, use_social=True)
    st = tr.state()
    assert set(st) == {'epoch', 'attentioner_dict', 'feature_embedder_dict', 'encoder_dict', 'decoder_dict',
                       'pred_optimizer', 'D_dict', 'D_optimizer'}                          # train.py:653-663
    ref = so.init_weights(n_next=2)
    for tag, key in (("attention", 'attentioner_dict'), ("feature_embedder", 'feature_embedder_dict'),
                     ("encoder", 'encoder_dict'), ("decoder", 'decoder_dict'), ("D", 'D_dict')):
        want = {k[len(tag) + 1:]: tuple(v.shape) for k, v in ref.items() if k.startswith(tag + ".")}
        assert {k: tuple(v.shape) for k, v in st[key].items()} == want


@pytest.mark.parametrize("mode", ["nccl", "native"])
def test_sharded_training_matches_single_gpu(mode):
    """N-rank run of the sharded step vs one GPU (needs >= 2 visible GPUs; `gpurun --gpus 2`): NCCL all-reduce + torch Adam,
    and the native iteration with the all-reduce inside the Adam kernel.  (On a single-GPU box the same comparison runs
    inside `bench.py --gpus N` as train_step.train_parity_max_abs.)"""
    import os
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29517",
                        os.path.join(here, "multigpu_train_check.py")], capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, SW_CHECK_MODE=mode))
    assert "MULTIGPU_TRAIN_CHECK PASS" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_cuda_graph_training_matches_eager():
    """train_graphed(): capture + replay of the whole GAN iteration vs the eager train() from the same seeds.
    (capturable Adam computes its bias corrections on the device: weights agree to fp32 rounding, not bit for bit.)"""
    from oracle import socialways_oracle as so
    from socialways_b200.trainer import SocialWaysTrainer
    data = so.toy_samples(216, 6)
    W = so.init_weights(seed=9, n_next=2)
    runs = {}
    for graph in (False, True):
        tr = SocialWaysTrainer(data, batch_size=64, use_social=True, n_unrolling_steps=1, weights=W, cuda_graph=graph)
        np.random.seed(3)
        torch.manual_seed(3)
        res = [(tr.train_graphed if graph else tr.train)(verbose=False) for _ in range(3)]
        runs[graph] = (res, tr.reference_weights())
        if graph:
            assert any(e["graph"] is not None for e in tr._graphs.values()), "no batch shape was captured"
    for (a0, f0), (a1, f1) in zip(runs[False][0], runs[True][0]):
        assert abs(a0 - a1) < 1e-4 and abs(f0 - f1) < 1e-4
    for k, v in runs[False][1].items():
        assert (v - runs[True][1][k]).abs().max().item() < 5e-5, k
