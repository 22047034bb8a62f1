"""Device-side latent noise (sw_noise_uniform): bit-exact against a numpy restatement of Philox4x32-10, and usable as
the noise source of predict_k (same result as passing the generated tensor explicitly)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def philox4x32_10(counter, key):
    """counter [n, 4] uint32, key [2] uint32 -> [n, 4] uint32 (Salmon et al. 2011, the cuRAND / torch constants)."""
    M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
    c = counter.astype(np.uint64)
    k0, k1 = int(key[0]), int(key[1])
    for _ in range(10):
        p0, p1 = M0 * c[:, 0], M1 * c[:, 2]
        hi0, lo0, hi1, lo1 = p0 >> 32, p0 & 0xFFFFFFFF, p1 >> 32, p1 & 0xFFFFFFFF
        c = np.stack([hi1 ^ c[:, 1] ^ k0, lo1, hi0 ^ c[:, 3] ^ k1, lo0], axis=1)
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return c.astype(np.uint32)


@pytest.mark.parametrize("n", [4, 1023, 40000])
def test_noise_matches_numpy_philox(n):
    from socialways_b200 import ops
    seed, offset = 0x1234_5678_9ABC_DEF1, 0x7_0000_0003
    got = ops.noise_uniform((n,), torch.device("cuda"), seed, offset).cpu().numpy()
    g = np.arange((n + 3) // 4, dtype=np.uint64)
    ctr = np.stack([g & 0xFFFFFFFF, g >> 32, np.full_like(g, offset & 0xFFFFFFFF), np.full_like(g, offset >> 32)], axis=1)
    bits = philox4x32_10(ctr, (seed & 0xFFFFFFFF, seed >> 32)).reshape(-1)[:n]
    want = (bits >> 8).astype(np.float32) * np.float32(2.0 ** -24)
    assert np.array_equal(got, want)
    assert 0.0 <= got.min() and got.max() < 1.0 and (n < 10000 or abs(got.mean() - 0.5) < 0.01)


def test_noise_slices_of_one_stream():
    """first_element: a rank draws exactly its rows of the iteration's logical noise tensor (sharded training)."""
    from socialways_b200 import ops
    dev = torch.device("cuda")
    full = ops.noise_uniform((4096, 32), dev, 99, offset=7)
    part = ops.noise_uniform((1000, 32), dev, 99, offset=7, first_element=517 * 32)
    assert torch.equal(part, full[517:1517])
    assert not torch.equal(ops.noise_uniform((4096, 32), dev, 99, offset=8), full)


def test_native_training_with_device_noise_runs_and_learns():
    import numpy as np
    from oracle import socialways_oracle as so
    from socialways_b200.trainer import SocialWaysTrainer
    tr = SocialWaysTrainer(so.toy_samples(216, 6), batch_size=64, use_social=True, weights=so.init_weights(seed=1, n_next=2),
                           fused_adam=True)
    np.random.seed(0)
    res = [tr.train_native(verbose=False, device_noise_seed=5) for _ in range(4)]
    assert all(np.isfinite(r).all() for r in res) and res[0] != res[-1]
    assert all(np.isfinite(v) for v in tr.last_epoch_mean_losses.values())


def test_predict_k_with_device_noise():
    import socialways_b200 as sw
    from socialways_b200 import ops
    from golden_data import synthetic_scenes
    d = synthetic_scenes([3, 8, 1, 5], seed=2)
    gen = sw.Generator(use_social=True).cuda()
    obsv = torch.from_numpy(d["obsvs"]).cuda() * 0.05
    a = gen.predict_k(obsv, None, 12, d["batches"], seed=(77, 5), k=6)
    z = ops.noise_uniform((6, obsv.shape[0], 32), obsv.device, 77, 5)
    b = gen.predict_k(obsv, z, 12, d["batches"])
    assert a.shape == (6, obsv.shape[0], 12, 4) and torch.equal(a, b)
    with pytest.raises(ValueError):
        gen.predict_k(obsv, None, 12, d["batches"])
