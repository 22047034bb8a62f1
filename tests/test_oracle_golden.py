"""Pin the CPU oracle (oracle/socialways_oracle.py) to vectors produced by the unmodified reference
(tests/golden/make_golden.py).  CPU only; this is the gate before the oracle is trusted as checker."""
import numpy as np
import pytest
import torch

from conftest import golden_weights, load_golden
from oracle import socialways_oracle as so

torch.set_num_threads(1)


@pytest.mark.parametrize("n,c", [(216, 6), (768, 6), (768, 8)])
def test_toy_generator_matches_reference(n, c):
    g = load_golden(f"toy_{n}_{c}.npz")
    d = so.toy_samples(n, c)
    for k in ("obsvs", "preds", "times", "batches"):
        assert d[k].shape == g[k].shape, k
        assert np.array_equal(d[k], g[k]), k            # bit-exact: same RNG stream, same float ops


def test_toy_scene_structure():
    d6 = so.toy_samples(216, 6)
    assert np.all(d6["batches"][:, 1] - d6["batches"][:, 0] == 6) and len(d6["batches"]) == 36
    d8 = so.toy_samples(768, 8)                          # README's 8 conditions: ragged 1..3 agents
    sizes = d8["batches"][:, 1] - d8["batches"][:, 0]
    assert len(d8["batches"]) == 512 and sizes.min() == 1 and sizes.max() <= 3


@pytest.mark.parametrize("case", ["ops_ragged.npz", "ops_zara.npz"])
def test_operator_forward(case):
    g = load_golden(case)
    P = golden_weights(g)
    obsv, pred = torch.from_numpy(g["obsv"]), torch.from_numpy(g["pred"])
    o4, p4 = so.traj_4d(obsv, pred)
    assert np.array_equal(o4.numpy(), g["obsv_4d"]) and np.array_equal(p4.numpy(), g["pred_4d"])
    n = obsv.shape[0]
    h, c = so.encoder_steps(P, o4, torch.zeros(n, 64), torch.zeros(n, 64))
    np.testing.assert_allclose(h.numpy(), g["enc_h"], atol=2e-6, rtol=0)
    np.testing.assert_allclose(c.numpy(), g["enc_c"], atol=2e-6, rtol=0)
    f = so.social_features(o4[:, -1])
    np.testing.assert_allclose(f.numpy(), g["social_features"], atol=1e-6, rtol=1e-6)
    emb = so.embed_features(P, f)
    np.testing.assert_allclose(emb.numpy(), g["social_emb"], atol=1e-5, rtol=1e-5)
    scenes = g["scenes"]
    s_loop = so.attention_pool_loop(P, emb, torch.from_numpy(g["enc_h"]), scenes)
    s_closed = so.attention_pool_closed(P, o4[:, -1], torch.from_numpy(g["enc_h"]), scenes)
    np.testing.assert_allclose(s_loop.numpy(), g["pooled"], atol=2e-6, rtol=0)
    np.testing.assert_allclose(s_closed.numpy(), g["pooled"], atol=2e-6, rtol=0)
    np.testing.assert_allclose(so.constant_velocity(obsv, 12).numpy(), g["cv"], atol=1e-6, rtol=0)
    noise = torch.from_numpy(g["noise"])
    for tag, social in (("soc", True), ("nos", False)):
        for pool in ("loop", "closed"):
            hat = so.predict(P, obsv, noise, 12, scenes, use_social=social, pool=pool)
            np.testing.assert_allclose(hat.numpy(), g[f"{tag}.pred_hat"], atol=5e-6, rtol=0)
        lab, code = so.discriminator(P, o4, torch.from_numpy(g[f"{tag}.pred_hat"]))
        np.testing.assert_allclose(lab.numpy(), g[f"{tag}.gen_label"], atol=2e-6, rtol=0)
        np.testing.assert_allclose(code.numpy(), g[f"{tag}.gen_code"], atol=2e-6, rtol=0)
        lab, code = so.discriminator(P, o4, p4)
        np.testing.assert_allclose(lab.numpy(), g[f"{tag}.real_label"], atol=2e-6, rtol=0)
        np.testing.assert_allclose(code.numpy(), g[f"{tag}.real_code"], atol=2e-6, rtol=0)


@pytest.mark.parametrize("case", ["ops_ragged.npz", "ops_zara.npz"])
@pytest.mark.parametrize("social", [True, False])
def test_operator_gradients(case, social):
    g = load_golden(case)
    tag = "soc" if social else "nos"
    P = {k: v.requires_grad_(True) for k, v in golden_weights(g).items()}
    obsv, pred, noise = (torch.from_numpy(g[k]) for k in ("obsv", "pred", "noise"))
    o4, p4 = so.traj_4d(obsv, pred)
    n = obsv.shape[0]
    hat = so.predict(P, obsv, noise, 12, g["scenes"], use_social=social, pool="closed")
    lab, code = so.discriminator(P, o4, hat)
    g_loss = so.mse(lab, torch.full((n, 1), 0.95)) + 0.5 * so.mse(code.squeeze(), noise[:, :2])
    assert abs(g_loss.item() - float(g[f"{tag}.g_loss"])) < 2e-6
    g_loss.backward()
    for k, p in P.items():
        if k.startswith("D."):
            continue
        want = g[f"{tag}.ggrad.{k}"]
        got = p.grad.numpy() if p.grad is not None else np.zeros_like(want)
        scale = max(1e-3, np.abs(want).max())
        assert np.abs(got - want).max() <= 2e-5 * scale + 1e-8, k
    for p in P.values():
        p.grad = None
    fl, fc = so.discriminator(P, o4, hat.detach())
    rl, _ = so.discriminator(P, o4, p4)
    d_loss = so.mse(fl, torch.full((n, 1), 0.05)) + so.mse(rl, torch.full((n, 1), 0.95)) + \
        0.5 * so.mse(fc.squeeze(), noise[:, :2])
    assert abs(d_loss.item() - float(g[f"{tag}.d_loss"])) < 2e-6
    d_loss.backward()
    for k, p in P.items():
        if k.startswith("D."):
            want = g[f"{tag}.dgrad.{k}"]
            scale = max(1e-3, np.abs(want).max())
            assert np.abs(p.grad.numpy() - want).max() <= 2e-5 * scale + 1e-8, k


@pytest.mark.parametrize("case,data_fn", [
    ("train_toy_216.npz", lambda: so.toy_samples(216, 6)),
    ("train_ragged.npz", None),
    ("train_unroll0.npz", None),
])
@pytest.mark.parametrize("social", [True, False])
def test_training_epochs_and_test_metrics(case, data_fn, social):
    """train() x epochs + test(K) from the same seeds: every mse_loss value the reference evaluated,
    its printed ADE/FDE lines, the post-training weights and a noise-pinned K-sample evaluation."""
    from golden_data import case_data
    g = load_golden(case)
    tag = "soc" if social else "nos"
    data = data_fn() if data_fn else case_data(case)
    seed = int(g["seed"][0])
    tr = so.OracleTrainer(golden_weights(g, "w0."), data, batch_size=int(g["batch_size"]), use_social=social,
                          unroll=int(g["unroll"]), pool="loop")
    np.random.seed(seed)
    torch.manual_seed(seed)
    lines = []
    for ep in range(1, int(g["epochs"]) + 1):
        ade, fde = tr.train_epoch()
        lines.append(" Epc=%4d, Train ADE,FDE = (%.3f, %.3f)" % (ep, ade, fde))
    m = tr.test_epoch(int(g["k_test"]))
    lines.append('Avg ADE,FDE (12)= (%.3f, %.3f) | Min(20) ADE,FDE (12)= (%.3f, %.3f)'
                 % (m["ade_avg"], m["fde_avg"], m["ade_min"], m["fde_min"]))
    ref_lines = [str(s) for s in g[f"{tag}.stdout"]]
    assert [r.split(" | time")[0] for r in ref_lines[:-1]] == lines[:-1]
    assert ref_lines[-1] == lines[-1]
    # loss trace: D step = (fake, info, real) per unroll pass, G step = (l2, fooling, info)
    ref_mse = g[f"{tag}.mse_values"]
    per_iter = 3 * (int(g["unroll"]) + 1) + 3
    assert len(ref_mse) == per_iter * len(tr.log)
    for it, rec in enumerate(tr.log):
        row = ref_mse[it * per_iter:(it + 1) * per_iter]
        d_last = row[-6] + row[-4] + 0.5 * row[-5]
        assert abs(rec["d_loss"] - d_last) < 5e-6
        assert abs(rec["g_fool"] - row[-2]) < 5e-6 and abs(rec["g_info"] - row[-1]) < 5e-6
    W = tr.weights()
    names = [str(s) for s in g[f"{tag}.w1_names"]]
    for i, k in enumerate(names):
        assert abs(W[k].double().sum().item() - g[f"{tag}.w1_sum"][i]) < 2e-4, k
        assert abs(W[k].double().norm().item() - g[f"{tag}.w1_l2"][i]) < 2e-4, k
    if social and "w1.encoder.embed.weight" in g:
        for k, v in golden_weights(g, "w1.").items():
            assert (W[k] - v).abs().max().item() < 2e-5, k
    torch.manual_seed(seed + 1)
    m2 = tr.test_epoch(int(g["k_test"]))
    want = g[f"{tag}.test_metrics"]
    got = np.array([m2["ade_avg"], m2["fde_avg"], m2["ade_min"], m2["fde_min"]])
    np.testing.assert_allclose(got, want, atol=1e-4, rtol=0)       # the 1e-4 ADE/FDE bar of north_star


@pytest.mark.parametrize("n,c", [(216, 6), (768, 8)])
def test_product_toy_generator_matches_reference(n, c):
    """socialways_b200.toy (what the repo's create_toy.py entry point runs) vs the reference's dataset."""
    from socialways_b200.toy import create_samples, pack_scenes
    g = load_golden(f"toy_{n}_{c}.npz")
    np.random.seed(30)
    samples, ts = create_samples(n, c, 3, n_per_batch=6)
    obsvs, preds, times, batches = pack_scenes(samples, ts)
    assert np.array_equal(obsvs, g["obsvs"]) and np.array_equal(preds, g["preds"])
    assert np.array_equal(times, g["times"]) and np.array_equal(batches, g["batches"])


def test_scale_matches_oracle():
    from socialways_b200.scale import Scale
    d = so.toy_samples(216, 6)
    ref = so.IsoScale(d["obsvs"], d["preds"])
    s = Scale()
    s.min_x, s.max_x = min(d["obsvs"][..., 0].min(), d["preds"][..., 0].min()), max(d["obsvs"][..., 0].max(), d["preds"][..., 0].max())
    s.min_y, s.max_y = min(d["obsvs"][..., 1].min(), d["preds"][..., 1].min()), max(d["obsvs"][..., 1].max(), d["preds"][..., 1].max())
    s.calc_scale(keep_ratio=True)
    assert s.sx == ref.sx == s.sy
    a = d["obsvs"].copy()
    n1 = s.normalize(a.copy())
    assert np.array_equal(n1, ref.normalize(a))
    assert np.allclose(s.denormalize(n1), a, atol=1e-6)
