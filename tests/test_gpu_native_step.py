"""GPU parity of the NATIVE training iteration (socialways_b200/native_step.py: ~30 launches of this library's kernels,
no autograd / cuBLAS / ATen) against (a) plain torch evaluations of the same contractions / folds, (b) the autograd path of
the package (itself pinned to the reference's gradients, tests/test_gpu_training.py) and (c) whole train()/test() runs of
the unmodified reference (tests/golden/train_*.npz): every mse_loss value, the printed ADE/FDE lines, post-training weights."""
import ctypes

import numpy as np
import pytest
import torch

from conftest import golden_weights, load_golden
from golden_data import case_data, synthetic_scenes

pytestmark = pytest.mark.gpu


def _images(rows_major):
    """[rows, k] -> tile images [ceil(rows/32), k, 32] (zero padded)."""
    n, k = rows_major.shape
    t = (n + 31) // 32
    pad = torch.zeros(t * 32, k, device=rows_major.device)
    pad[:n] = rows_major
    return pad.view(t, 32, k).transpose(1, 2).contiguous()


@pytest.mark.parametrize("tensor_cores", [False, True])
def test_contract_kinds_permutation_and_split(tensor_cores):
    """sw_contract (FFMA) and sw_contract_tc (tcgen05, tf32-split operands) against float64 matmuls: image and row-record
    operands, the ones row (bias gradients), segments, the LSTM gate permutation, jobs split over CTAs, operands whose
    magnitudes span 1e-9 .. 1 (gradient records)."""
    from socialways_b200.native_step import ContractPlan, _job, seg, ROWS
    g = torch.Generator(device="cuda").manual_seed(1)
    rows = 32 * 700 - 5                      # many images: the jobs are split over CTAs (partials + fixed-order reduce)
    A = torch.randn(rows, 68, device="cuda", generator=g)
    B = torch.randn(rows, 256, device="cuda", generator=g) * torch.logspace(-9, 0, 256, device="cuda")   # gradient-like range
    Ai, Bi = _images(A), _images(B)
    n_img = Ai.shape[0]
    w_ih, w_hh = torch.zeros(256, 4, device="cuda"), torch.zeros(256, 64, device="cuda")
    b_ih, b_hh = torch.zeros(256, device="cuda"), torch.zeros(256, device="cuda")
    out_plain = torch.zeros(69, 256, device="cuda")
    small_a = torch.randn(777, 7, device="cuda", generator=g)
    small_b = torch.randn(777, 130, device="cuda", generator=g) * 1e-4
    out_rows, out_rows_b = torch.zeros(100, 3, device="cuda"), torch.zeros(100, device="cuda")
    wide_a = torch.randn(32 * 9, 160, device="cuda", generator=g)                                    # K > 128: two slabs
    wide_b = torch.randn(32 * 9, 80, device="cuda", generator=g)
    out_wide, out_wide_b = torch.zeros(80, 160, device="cuda"), torch.zeros(80, device="cuda")
    jobs = [
        _job(Ai, Bi, 68, 256, n_img, [seg(w_ih, 0, 4, 1, 4), seg(w_hh, 4, 64, 1, 64), seg(b_ih, 68, 1, 0, 1, out2=b_hh)],
             68 * 32, 256 * 32, ones=True, perm=1),
        _job(Ai, Bi, 68, 256, n_img, [seg(out_plain, 0, 69, 256, 1)], 68 * 32, 256 * 32, ones=True),
        _job(small_a, small_b, 3, 100, (777 + 31) // 32, [seg(out_rows, 0, 3, 1, 3), seg(out_rows_b, 3, 1, 0, 1)], 7, 130,
             a_k0=2, b_n0=30, a_kind=ROWS, b_kind=ROWS, ones=True, n_rows=777),
        _job(_images(wide_a), _images(wide_b), 160, 80, 9, [seg(out_wide, 0, 160, 1, 160), seg(out_wide_b, 160, 1, 0, 1)],
             160 * 32, 80 * 32, ones=True),
    ]
    plan = ContractPlan(jobs, torch.device("cuda"), tensor_cores)
    assert plan.ws.numel() > 4, "the large jobs must be split over several CTAs"
    for _ in range(2):                                   # second launch: the slab counters were restored
        plan.run()
    torch.cuda.synchronize()
    tol = 3e-6 if tensor_cores else 1e-6                 # relative to sum |a||b| of the element (tf32 split: ~2^-21 per product)

    def check(got, want, bound):
        assert ((got.double() - want).abs() <= tol * bound + 1e-30).all(), float(((got.double() - want).abs() / (bound + 1e-30)).max())

    ref = A.double().t() @ B.double()                    # [68, 256], columns n' = 4*unit + gate
    bound = A.double().abs().t() @ B.double().abs()
    bsum, bbound = B.double().sum(0), B.double().abs().sum(0)
    perm = torch.tensor([(n & 3) * 64 + (n >> 2) for n in range(256)], device="cuda")
    scat = lambda m: torch.zeros(256, m.shape[0], device="cuda", dtype=torch.float64).index_copy_(0, perm, m.t().contiguous())
    check(w_ih, scat(ref[:4]), scat(bound[:4]))
    check(w_hh, scat(ref[4:68]), scat(bound[4:68]))
    check(b_ih, scat(bsum[None])[:, 0], scat(bbound[None])[:, 0])
    assert torch.equal(b_ih, b_hh)
    check(out_plain, torch.cat([ref, bsum[None]]), torch.cat([bound, bbound[None]]))
    sa, sb = small_a[:, 2:5].double(), small_b[:, 30:130].double()
    check(out_rows, (sa.t() @ sb).t(), (sa.abs().t() @ sb.abs()).t())
    check(out_rows_b, sb.sum(0), sb.abs().sum(0))
    check(out_wide, (wide_a.double().t() @ wide_b.double()).t(), (wide_a.double().abs().t() @ wide_b.double().abs()).t())
    check(out_wide_b, wide_b.double().sum(0), wide_b.double().abs().sum(0))
    first = out_plain.clone()
    plan.run()
    torch.cuda.synchronize()
    assert torch.equal(first, out_plain), "fixed summation order: bit-identical across launches"


def test_rows_linear():
    from socialways_b200 import _lib
    from socialways_b200.ops import _stream
    g = torch.Generator(device="cuda").manual_seed(2)
    for n, k, m in ((77, 64, 65), (300, 65, 64), (5, 3, 80)):
        x = torch.randn(n, k, device="cuda", generator=g)
        w = torch.randn(k, m, device="cuda", generator=g)
        b = torch.randn(m, device="cuda", generator=g)
        a1 = torch.randn(n, m, device="cuda", generator=g)
        a2 = torch.randn(n, m, device="cuda", generator=g)
        out = torch.empty(n, m, device="cuda")
        _lib.check(_lib.lib().sw_rows_linear(x.data_ptr(), k, w.data_ptr(), b.data_ptr(), a1.data_ptr(), a2.data_ptr(),
                                             out.data_ptr(), m, n, k, m, _stream()), "sw_rows_linear")
        want = (x.double() @ w.double() + b.double() + a1.double() + a2.double())
        assert (out.double() - want).abs().max().item() < 1e-4


def _trainer(data, social=True, unroll=1, weights=None, n_next=12, batch=64, tensor_cores=True, **kw):
    from oracle import socialways_oracle as so
    from socialways_b200.trainer import SocialWaysTrainer
    W = weights if weights is not None else so.init_weights(seed=4, n_next=n_next)
    tr = SocialWaysTrainer(data, batch_size=batch, use_social=social, n_unrolling_steps=unroll, weights=W, fused_adam=True, **kw)
    tr.native_tensor_cores = "force" if tensor_cores else False      # small test batches: force the tcgen05 kernel
    return tr


def test_pack_kernels_match_packing_py():
    from socialways_b200 import packing
    from socialways_b200.native_step import NativePacks
    rng = np.random.RandomState(0)
    tr = _trainer(synthetic_scenes(list(rng.randint(1, 9, size=12)), seed=3))
    pk = NativePacks(tr)
    pk.pack_generator()
    pk.pack_discriminator()
    torch.cuda.synchronize()
    gen, D = tr.generator, tr.D
    with torch.no_grad():
        enc = gen.encoder.packed()
        dec = gen.decoder.packed()
        fe, att = gen.feature_embedder.fc, gen.attention.W
        m, m0 = packing.pool_agent_matrix(att.weight, att.bias, fe[4].weight, fe[4].bias)
        pool = packing.pack_pool(fe[0].weight, fe[0].bias, fe[2].weight, fe[2].bias)
        dl = D.packed_lstm()
    close = lambda a, b, tol=2e-6: (a.reshape(-1) - b.reshape(-1)).abs().max().item() <= tol * max(1.0, b.abs().max().item())
    assert close(pk.enc, enc) and close(pk.enc_t, enc[:68].t().contiguous())
    assert close(pk.dec, dec)
    w1 = dec[:25600].view(160, 160)
    w2 = dec[25760:38560].view(160, 80)
    assert close(pk.dec_t, torch.cat([w1[:64].t().reshape(-1), w2.t().reshape(-1), dec[-162:-2]]))
    assert close(pk.pool, pool)
    assert close(pk.pool_m, torch.cat([m.reshape(-1), m0])) and close(pk.pool_mt, m.t().contiguous())
    assert close(pk.d_lstm, dl) and close(pk.d_lstm_t, dl[:68].t().contiguous())


@pytest.mark.parametrize("social", [True, False])
@pytest.mark.parametrize("shape", ["ragged_8_12", "toy_2_2"])
@pytest.mark.parametrize("tensor_cores", [True, False])
def test_native_gradients_match_autograd_path(social, shape, tensor_cores):
    """One D pass and one G pass: every parameter gradient of the native launch sequence vs the package's autograd path
    (which tests/test_gpu_training.py pins to the unmodified reference's gradients)."""
    from oracle import socialways_oracle as so
    from socialways_b200.native_step import NativePacks, NativeStep
    from socialways_b200.reference_api import get_traj_4d
    if shape == "toy_2_2":
        data, n_next = so.toy_samples(216, 6), 2
    else:
        rng = np.random.RandomState(1)
        data, n_next = synthetic_scenes(list(rng.randint(1, 9, size=14)), seed=5), 12
    tr = _trainer(data, social=social, n_next=n_next, tensor_cores=tensor_cores)
    lo, hi, sub = next(iter(tr._minibatches()))
    bs = hi - lo
    step = NativeStep(tr, NativePacks(tr), bs, tr.generator.scene_index(sub, bs, tr.device), bs)
    torch.manual_seed(11)
    step.obsv.copy_(tr.dataset_obsv[lo:hi])
    step.pred.copy_(tr.dataset_pred[lo:hi])
    step.noise.copy_(torch.rand(bs, 32))
    step.targets.copy_(torch.tensor([0.07, 0.93]))
    gparams, dparams = step.pk.gen_params, step.pk.disc_params

    # ---- native ----
    with torch.no_grad():
        step.generator_forward()
        step.discriminator_grads()
        d_native = [p.grad.clone() for p in dparams]
        step.generator_grads()
        g_native = [p.grad.clone() for p in gparams]
        hat_native = step.out[0].clone()
        step._disc_step(0)                                     # loss partial sums of the D pass again (stats below)
    import socialways_b200._lib as _lib
    from socialways_b200.ops import _stream, sm_count
    _lib.check(_lib.lib().sw_train_stats(step.out.data_ptr(), step.pred.data_ptr(), bs, step.Tp, float(tr.ss), step.loss_d.data_ptr(),
                                         step.t16, step.loss_g.data_ptr(), step.t32, step.inv_n, step.info_w,
                                         step.stats_partial.data_ptr(), step.stats_counter.data_ptr(), step.stats.data_ptr(),
                                         sm_count(tr.device), _stream()), "sw_train_stats")
    stats = step.stats.tolist()

    # ---- autograd path of the package, same inputs ----
    tr.D_optimizer.zero_grad()
    tr.predictor_optimizer.zero_grad()
    obsv, pred, noise = step.obsv, step.pred, step.noise
    zeros = torch.full((bs, 1), 0.07, device="cuda")
    ones = torch.full((bs, 1), 0.93, device="cuda")
    obsv_4d, pred_4d = get_traj_4d(obsv, pred)
    mse = torch.nn.MSELoss()
    with torch.no_grad():
        hat = tr.predict(obsv, noise, tr.n_next, sub)
    assert (hat - hat_native).abs().max().item() < 1e-6
    oh = tr.D.encode_obsv(obsv_4d)
    fl, code = tr.D.heads(oh, hat)
    rl, _ = tr.D.heads(oh, pred_4d)
    d_fake, d_real, d_info = mse(fl, zeros), mse(rl, ones), mse(code, noise[:, :2])
    (d_fake + d_real + 0.5 * d_info).backward()
    d_ref = [p.grad.clone() for p in dparams]
    tr.D_optimizer.zero_grad()
    hat2 = tr.predict(obsv, noise, tr.n_next, sub)
    with torch.no_grad():
        oh = tr.D.encode_obsv(obsv_4d)
    gl, code = tr.D.heads(oh, hat2)
    g_fool, g_info = mse(gl, ones), mse(code, noise[:, :2])
    (g_fool + 0.5 * g_info).backward()
    g_ref = [p.grad.clone() for p in gparams]

    def check(name, got, want):
        scale = max(1e-4, want.abs().max().item())
        err = (got - want).abs().max().item()
        assert err <= 2e-4 * scale + 1e-8, f"{name}: max abs err {err:.3e} vs scale {scale:.3e}"

    for i, (a, b) in enumerate(zip(d_native, d_ref)):
        check(f"D[{i}] {tuple(b.shape)}", a, b)
    for i, (a, b) in enumerate(zip(g_native, g_ref)):
        check(f"G[{i}] {tuple(b.shape)}", a, b)
    want = [d_fake.item() + d_real.item() + 0.5 * d_info.item(), d_fake.item(), d_real.item(), d_info.item(),
            g_fool.item(), g_info.item()]
    for a, b in zip(stats[2:], want):
        assert abs(a - b) < 2e-6 * max(1.0, abs(b)), (stats, want)
    err = ((hat[:, :, :2] - pred) / tr.ss).pow(2).sum(2).sqrt()
    assert abs(stats[0] - err.sum().item() / tr.n_next) < 1e-3 * max(1.0, stats[0]) * 1e-2
    assert abs(stats[1] - err[:, -1].sum().item()) < 1e-5 * max(1.0, stats[1])


@pytest.mark.parametrize("case", ["train_toy_216.npz", "train_ragged.npz", "train_unroll0.npz"])
@pytest.mark.parametrize("social", [True, False])
@pytest.mark.parametrize("graph", [False, True])
def test_native_training_epochs_vs_reference_golden(case, social, graph, capsys):
    """train_native() x epochs from the golden seeds vs the unmodified reference: every mse_loss value, the printed
    train lines, the weights after training, the K-sample test metrics (same bars as the autograd path)."""
    from oracle import socialways_oracle as so
    from socialways_b200.trainer import SocialWaysTrainer
    g = load_golden(case)
    tag = "soc" if social else "nos"
    data = so.toy_samples(216, 6) if case == "train_toy_216.npz" else case_data(case)
    seed = int(g["seed"][0])
    tr = SocialWaysTrainer(data, batch_size=int(g["batch_size"]), use_social=social, n_unrolling_steps=int(g["unroll"]),
                           weights=golden_weights(g, "w0."), fused_adam=True)
    tr.native_tensor_cores = "force" if graph else True      # both contraction kernels see the golden runs
    np.random.seed(seed)
    torch.manual_seed(seed)
    for ep in range(1, int(g["epochs"]) + 1):
        tr.epoch = ep
        tr.train_native(use_graph=graph, log_losses=True)
    lines = [l for l in capsys.readouterr().out.splitlines() if l.strip()]
    ref_lines = [str(s) for s in g[f"{tag}.stdout"]]
    import re
    nums = lambda t: [float(x) for x in re.findall(r"-?\d+\.\d+", t.split(" | time")[0])]
    for r, l in zip(ref_lines[:-1], lines):
        assert len(nums(r)) == len(nums(l)) and all(abs(a - b) <= 1.001e-3 for a, b in zip(nums(r), nums(l))), (r, l)
    ref_mse = g[f"{tag}.mse_values"]
    per_iter = 3 * (int(g["unroll"]) + 1) + 3
    assert len(ref_mse) == per_iter * len(tr.loss_log)
    for it, rec in enumerate(tr.loss_log):
        row = ref_mse[it * per_iter:(it + 1) * per_iter]
        assert abs(rec["d_fake"] - row[-6]) < 1e-5 and abs(rec["d_info"] - row[-5]) < 1e-5, (it, rec, row[-6:])
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


def test_graphed_training_followed_by_test_uses_current_weights():
    """graph replays rewrite the parameters without bumping their version counters; test() after
    every epoch must see the CURRENT weights (packed-weight cache invalidated), i.e. equal the eager trainer's metrics."""
    from oracle import socialways_oracle as so
    data = so.toy_samples(216, 6)
    W = so.init_weights(seed=9, n_next=2)
    runs = {}
    for mode in ("eager", "graphed", "native"):
        tr = _trainer(data, weights=W, n_next=2, cuda_graph=(mode == "graphed"))
        np.random.seed(3)
        torch.manual_seed(3)
        res = []
        for ep in range(5):
            {"eager": tr.train, "graphed": tr.train_graphed, "native": tr.train_native}[mode](verbose=False)
            m = tr.test(5, verbose=False)
            res.append([m["ade_avg"], m["fde_avg"], m["ade_min"], m["fde_min"]])
        runs[mode] = np.array(res)
    # a stale cache would repeat an earlier epoch's metrics: those move by ~2e-2 per epoch, an order of magnitude above the
    # tolerance (which only has to absorb fp32 rounding differences between the three paths, amplified by 5 GAN epochs)
    assert np.abs(runs["eager"][1:] - runs["eager"][:-1]).min(axis=0).max() > 1e-2, "metrics must move between epochs"
    np.testing.assert_allclose(runs["graphed"], runs["eager"], atol=2e-3, rtol=0)
    np.testing.assert_allclose(runs["native"], runs["eager"], atol=2e-3, rtol=0)
