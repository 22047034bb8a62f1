"""GPU parity of the forward (K-sample inference) path: CUDA kernels through the C-ABI vs the golden
vectors of the unmodified reference and vs the CPU oracle on the same seeded inputs.

Tolerances (fp32 path; the contract is 1e-4 on ADE/FDE in metres): operator outputs are compared in
normalised coordinates with atol 2e-5 (observed ~1e-6); metrics with atol 1e-4.
"""
import numpy as np
import pytest
import torch

from conftest import golden_weights, load_golden
from golden_data import synthetic_scenes

pytestmark = pytest.mark.gpu

ATOL = 2e-5
PRECISION = "fp32"      # set per test by the `precision` fixture: every test runs on all three fp32-faithful decode kernels
                        # (FFMA, CTA-pair tcgen05 = the default, one-tile-per-SM tcgen05)


@pytest.fixture(autouse=True, params=["fp32", "fp16x2", "fp16x2s"])
def precision(request):
    global PRECISION
    PRECISION = request.param
    yield request.param


def _generator(P, use_social=True):
    import socialways_b200 as sw
    g = sw.Generator(use_social=use_social)
    sd = {k: v for k, v in P.items() if not k.startswith("D.")}
    g.load_state_dict(sd, strict=True)
    g.inference_precision = PRECISION
    return g.cuda().requires_grad_(False)      # inference path; the autograd path has its own tests


@pytest.mark.parametrize("case", ["ops_ragged.npz", "ops_zara.npz"])
def test_operators_vs_reference_golden(case):
    from socialways_b200 import ops
    g = load_golden(case)
    gen = _generator(golden_weights(g))
    obsv = torch.from_numpy(g["obsv"]).cuda()
    pk = gen.packs()
    enc = ops.lstm_seq(pk["enc"], obsv, want_x_last=True)
    np.testing.assert_allclose(enc["h"].cpu().numpy(), g["enc_h"], atol=ATOL, rtol=0)
    np.testing.assert_allclose(enc["c"].cpu().numpy(), g["enc_c"], atol=ATOL, rtol=0)
    np.testing.assert_array_equal(enc["x_last"].cpu().numpy(), g["obsv_4d"][:, -1])
    # 4-d input form gives the same state
    enc4 = ops.lstm_seq(pk["enc"], torch.from_numpy(g["obsv_4d"]).cuda())
    assert torch.equal(enc4["h"], enc["h"]) and torch.equal(enc4["c"], enc["c"])
    scenes = ops.SceneIndex(g["scenes"], obsv.shape[0], obsv.device)
    h_ref = torch.from_numpy(g["enc_h"]).cuda()
    ub = torch.addmm(pk["pool_m0"], h_ref, pk["pool_m"])
    pooled, attn = ops.pool(pk["pool"], enc["x_last"], h_ref, ub, scenes, want_attn=True)
    np.testing.assert_allclose(pooled.cpu().numpy(), g["pooled"], atol=ATOL, rtol=0)
    sizes = g["scenes"][:, 1] - g["scenes"][:, 0]
    rowsum = attn.sum(1).cpu().numpy()
    expect = np.repeat((sizes > 1).astype(np.float32), sizes)
    np.testing.assert_allclose(rowsum, expect, atol=1e-5)
    noise = torch.from_numpy(g["noise"]).cuda()
    for tag, social in (("soc", True), ("nos", False)):
        gen.use_social = social
        hat = gen.predict(obsv, noise, 12, g["scenes"])
        assert hat.shape == (obsv.shape[0], 12, 4)
        np.testing.assert_allclose(hat.cpu().numpy(), g[f"{tag}.pred_hat"], atol=ATOL, rtol=0)


def test_predict_default_sub_batches_is_one_scene():
    """predict(obsv, noise, n_next) with no sub_batches pools over the whole batch (train.py:405-406)."""
    from oracle import socialways_oracle as so
    g = load_golden("ops_ragged.npz")
    P = golden_weights(g)
    gen = _generator(P)
    obsv, noise = torch.from_numpy(g["obsv"]), torch.from_numpy(g["noise"])
    want = so.predict(P, obsv, noise, 12, None, use_social=True, pool="closed")
    got = gen.predict(obsv.cuda(), noise.cuda(), 12)
    np.testing.assert_allclose(got.cpu().numpy(), want.numpy(), atol=ATOL, rtol=0)


@pytest.mark.parametrize("sizes,k", [([1], 3), ([2], 2), ([1, 1, 2, 31, 33, 1], 4), ([6] * 11, 20), ([40, 7, 64], 3)])
def test_predict_k_vs_oracle_ragged(sizes, k):
    from oracle import socialways_oracle as so
    P = so.init_weights(seed=5)
    data = synthetic_scenes(sizes, seed=21)
    sc = so.IsoScale(data["obsvs"], data["preds"])
    obsv = torch.from_numpy(sc.normalize(data["obsvs"]))
    pred = torch.from_numpy(sc.normalize(data["preds"]))
    n = obsv.shape[0]
    torch.manual_seed(3)
    noise = torch.rand(k, n, 32)
    want = torch.stack([so.predict(P, obsv, noise[i], 12, data["batches"], True, "closed") for i in range(k)])
    gen = _generator(P)
    got = gen.predict_k(obsv.cuda(), noise.cuda(), 12, data["batches"])
    np.testing.assert_allclose(got.cpu().numpy(), want.numpy(), atol=ATOL, rtol=0)
    # best-of-K metrics (train.py:587,602-607)
    from socialways_b200 import ops
    m = ops.bestofk_metrics(got, pred.cuda(), sc.sx).cpu().numpy()
    e = (((want[..., :2] - pred) / sc.sx) ** 2).sum(-1).sqrt()          # [K, N, T]
    ref = torch.stack([e.mean(2).mean(0), e[:, :, -1].mean(0), e.mean(2).min(0)[0], e[:, :, -1].min(0)[0]], 1)
    np.testing.assert_allclose(m, ref.numpy(), atol=1e-4, rtol=1e-5)


def test_toy_shapes_obs2_pred2():
    """Config 1: toy set, 2 observed + 2 predicted points, 6 agents per scene."""
    from oracle import socialways_oracle as so
    P = so.init_weights(seed=0)
    d = so.toy_samples(216, 6)
    sc = so.IsoScale(d["obsvs"], d["preds"])
    obsv = torch.from_numpy(sc.normalize(d["obsvs"]))
    torch.manual_seed(0)
    noise = torch.rand(216, 32)
    for social in (True, False):
        want = so.predict(P, obsv, noise, 2, d["batches"], social, "closed")
        gen = _generator(P, use_social=social)
        got = gen.predict(obsv.cuda(), noise.cuda(), 2, d["batches"])
        np.testing.assert_allclose(got.cpu().numpy(), want.numpy(), atol=ATOL, rtol=0)


@pytest.mark.parametrize("a", [256, 400])
def test_dense_scene_pooling(a):
    """Dense-crowd config (A=256, staged in shared memory) and a span beyond the staging cap (A=400,
    read through L1/L2): pooled vector vs the closed-form oracle."""
    from oracle import socialways_oracle as so
    from socialways_b200 import ops
    P = so.init_weights(seed=2)
    data = synthetic_scenes([a, 3], seed=9)
    sc = so.IsoScale(data["obsvs"], data["preds"])
    obsv = torch.from_numpy(sc.normalize(data["obsvs"]))
    x4 = so.traj_4d(obsv)
    n = obsv.shape[0]
    h, _ = so.encoder_steps(P, x4, torch.zeros(n, 64), torch.zeros(n, 64))
    want = so.attention_pool_closed(P, x4[:, -1], h, data["batches"])
    gen = _generator(P)
    pk = gen.packs()
    hc = h.cuda()
    ub = torch.addmm(pk["pool_m0"], hc, pk["pool_m"])
    got = ops.pool(pk["pool"], x4[:, -1].contiguous().cuda(), hc, ub, ops.SceneIndex(data["batches"], n, hc.device))
    np.testing.assert_allclose(got.cpu().numpy(), want.numpy(), atol=ATOL, rtol=0)


def test_encoder_module_carries_state_like_reference():
    """EncoderLstm.forward: whole sequence, then single steps from the carried state (train.py:268,430)."""
    from oracle import socialways_oracle as so
    P = so.init_weights(seed=4)
    gen = _generator(P)
    torch.manual_seed(1)
    x = torch.rand(37, 8, 4)
    step = torch.rand(37, 4)
    h, c = so.encoder_steps(P, x, torch.zeros(37, 64), torch.zeros(37, 64))
    h2, c2 = so.encoder_steps(P, step, h, c)
    enc = gen.encoder
    enc.init_lstm(torch.zeros(1, 37, 64).cuda(), torch.zeros(1, 37, 64).cuda())
    y = enc(x.cuda())
    assert y.shape == (37, 8, 64)
    np.testing.assert_allclose(enc.lstm_h[0][0].cpu().numpy(), h.numpy(), atol=ATOL, rtol=0)
    np.testing.assert_allclose(y[:, -1].cpu().numpy(), h.numpy(), atol=ATOL, rtol=0)
    y1 = enc(step.cuda())
    assert y1.shape == (37, 1, 64)
    np.testing.assert_allclose(enc.lstm_h[0][0].cpu().numpy(), h2.numpy(), atol=ATOL, rtol=0)
    np.testing.assert_allclose(enc.lstm_h[1][0].cpu().numpy(), c2.numpy(), atol=ATOL, rtol=0)
    # DecoderFC module forward = one decode step
    torch.manual_seed(2)
    hh, ss_, zz = torch.rand(37, 64), torch.rand(37, 64), torch.rand(37, 32)
    want = so.decoder_fc(P, hh, ss_, zz)
    got = gen.decoder(hh.cuda(), ss_.cuda(), zz.cuda())
    np.testing.assert_allclose(got.cpu().numpy(), want.numpy(), atol=ATOL, rtol=0)


def test_full_size_properties_eth_config():
    """BASELINE config 2 at bench size (A=8, K=20, obs 8 / pred 12): properties that need no oracle run
    at full size, plus an oracle spot check on a slice of scenes.
      * sample-independence: identical noise for two k gives bit-identical rows
      * scene-independence: a scene decoded alone equals the same scene inside the big batch
      * integration invariant: p_t - p_{t-1} == v_t exactly as emitted (train.py:423-425)"""
    from oracle import socialways_oracle as so
    n_scenes, a, k = 4096, 8, 20
    P = so.init_weights(seed=7)
    data = synthetic_scenes([a] * n_scenes, seed=1)
    sc = so.IsoScale(data["obsvs"], data["preds"])
    obsv = torch.from_numpy(sc.normalize(data["obsvs"]))
    n = obsv.shape[0]
    gen_cpu = torch.Generator().manual_seed(11)
    noise = torch.rand(k, n, 32, generator=gen_cpu)
    noise[5] = noise[2]
    gen = _generator(P)
    out = gen.predict_k(obsv.cuda(), noise.cuda(), 12, data["batches"])
    assert out.shape == (k, n, 12, 4) and torch.isfinite(out).all()
    assert torch.equal(out[5], out[2])
    last_obs = obsv[:, -1, :].cuda()
    first = torch.cat([last_obs[None, :, None, :].expand(k, -1, 1, -1), out[:, :, :-1, :2]], dim=2)
    step = out[..., :2] - first
    assert (step - out[..., 2:]).abs().max().item() < 1e-6
    lo, hi = 8 * 1000, 8 * 1003
    alone = gen.predict_k(obsv[lo:hi].cuda(), noise[:, lo:hi].contiguous().cuda(), 12, data["batches"][1000:1003] - lo)
    assert (alone - out[:, lo:hi]).abs().max().item() < 1e-6
    want = torch.stack([so.predict(P, obsv[lo:hi], noise[i, lo:hi], 12, data["batches"][1000:1003] - lo, True, "closed")
                        for i in (0, 7, 19)])
    np.testing.assert_allclose(out[[0, 7, 19], lo:hi].cpu().numpy(), want.numpy(), atol=ATOL, rtol=0)


def test_config5_dense_crowd_k128():
    """BASELINE config 5 shape: 256 agents per scene, K = 128 samples (65 536 decode rows for 2 scenes).
    Full-size run through the public API; oracle comparison on 3 of the 128 samples; sample-independence and
    finiteness on all of them."""
    from oracle import socialways_oracle as so
    P = so.init_weights(seed=8)
    data = synthetic_scenes([256, 256], seed=17)
    sc = so.IsoScale(data["obsvs"], data["preds"])
    obsv = torch.from_numpy(sc.normalize(data["obsvs"]))
    n, k = obsv.shape[0], 128
    noise = torch.rand(k, n, 32, generator=torch.Generator().manual_seed(5))
    noise[100] = noise[3]
    gen = _generator(P)
    out = gen.predict_k(obsv.cuda(), noise.cuda(), 12, data["batches"])
    assert out.shape == (k, n, 12, 4) and torch.isfinite(out).all()
    assert torch.equal(out[100], out[3])
    for i in (0, 64, 127):
        want = so.predict(P, obsv, noise[i], 12, data["batches"], True, "closed")
        np.testing.assert_allclose(out[i].cpu().numpy(), want.numpy(), atol=ATOL, rtol=0)
