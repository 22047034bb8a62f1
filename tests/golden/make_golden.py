"""Generate tests/golden/*.npz by running the UNMODIFIED reference on CPU (build container only).

    python tests/golden/make_golden.py

Every array written here comes out of /root/reference code executed through
reference_harness.py; seeds are recorded in the files.  The committed .npz files are what the
CPU tests pin the oracle to and what the GPU tests compare the CUDA path with.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import reference_harness as rh  # noqa: E402

torch.set_num_threads(1)          # deterministic reduction order for the recorded values


def save(name, **arrays):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **arrays)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


def toy_sets():
    for n, c in ((216, 6), (768, 6), (768, 8)):
        d = rh.reference_toy(n, c)
        save(f"toy_{n}_{c}.npz", **d)


class LossTap:
    """Wraps the reference's `mse_loss` global: records every value it returns, changes nothing."""

    def __init__(self):
        self.inner = torch.nn.MSELoss()
        self.values = []

    def __call__(self, a, b):
        out = self.inner(a, b)
        self.values.append(float(out))
        return out


def ops_case(name, scene_sizes, weight_seed, data_seed, noise_seed):
    """Operator-level vectors: forward of every op on the path, and the gradients of one D loss and
    one G loss (train.py:484-494, :512-523) w.r.t. every parameter, all from reference code."""
    data = rh.synthetic_scenes(scene_sizes, seed=data_seed)
    out = {}
    for social in (True, False):
        ref = rh.Reference(data, batch_size=10 ** 6, use_social=social, weight_seed=weight_seed)
        tag = "soc" if social else "nos"
        n = ref.dataset_obsv.shape[0]
        scenes = np.array(data["batches"])
        torch.manual_seed(noise_seed)
        noise = torch.rand(n, 32)
        obsv, pred = ref.dataset_obsv, ref.dataset_pred
        obsv_4d, pred_4d = ref.get_traj_4d(obsv, pred)
        if social:
            out.update({f"w.{k}": v for k, v in ref.state().items()})
            out.update(obsv=obsv.numpy(), pred=pred.numpy(), scenes=scenes, noise=noise.numpy(),
                       obsv_4d=obsv_4d.numpy(), pred_4d=pred_4d.numpy(), ss=np.float64(ref.ss),
                       seeds=np.array([weight_seed, data_seed, noise_seed]))
            with torch.no_grad():
                feats = ref.SocialFeatures(obsv_4d, scenes)
                emb = ref.feature_embedder(feats, scenes)
                out.update(social_features=feats.numpy(), social_emb=emb.numpy())
        # generator forward with grad, then the G loss through D (train.py:507-523)
        for m in ref.generator_modules() + [ref.D]:
            m.zero_grad()
        hat = ref.predict(obsv, noise, 12, scenes)
        if social:
            out["enc_h_after"] = ref.encoder.lstm_h[0].detach().squeeze(0).numpy()   # state after the extra step (:430)
        lab, code = ref.D(obsv_4d, hat)
        ones = torch.ones(n, 1) * 0.95
        zeros = torch.zeros(n, 1) + 0.05
        g_loss = ref.mse_loss(lab, ones) + 0.5 * ref.mse_loss(code.squeeze(), noise[:, :2])
        g_loss.backward()
        out[f"{tag}.pred_hat"] = hat.detach().numpy()
        out[f"{tag}.gen_label"] = lab.detach().numpy()
        out[f"{tag}.gen_code"] = code.detach().numpy()
        out[f"{tag}.g_loss"] = np.float64(g_loss.item())
        for mod_tag in ("attention", "feature_embedder", "encoder", "decoder"):
            for k, p in ref.ns[mod_tag].named_parameters():
                g = p.grad if p.grad is not None else torch.zeros_like(p)
                out[f"{tag}.ggrad.{mod_tag}.{k}"] = g.numpy().copy()
        # D loss (train.py:482-494) on detached fake + real
        ref.D.zero_grad()
        fake_lab, fake_code = ref.D(obsv_4d, hat.detach())
        real_lab, real_code = ref.D(obsv_4d, pred_4d)
        d_loss = ref.mse_loss(fake_lab, zeros) + ref.mse_loss(real_lab, ones) + \
            0.5 * ref.mse_loss(fake_code.squeeze(), noise[:, :2])
        d_loss.backward()
        out[f"{tag}.real_label"] = real_lab.detach().numpy()
        out[f"{tag}.real_code"] = real_code.detach().numpy()
        out[f"{tag}.d_loss"] = np.float64(d_loss.item())
        for k, p in ref.D.named_parameters():
            out[f"{tag}.dgrad.D.{k}"] = p.grad.numpy().copy()
        if social:
            # encoder state right after the observation, and the pooled vector, via the ref modules
            with torch.no_grad():
                ref.encoder.init_lstm(torch.zeros(1, n, 64), torch.zeros(1, n, 64))
                ref.encoder(obsv_4d)
                h = ref.encoder.lstm_h[0].squeeze(0)
                out["enc_h"] = h.numpy().copy()
                out["enc_c"] = ref.encoder.lstm_h[1].squeeze(0).numpy().copy()
                out["pooled"] = ref.attention(emb, h, scenes).numpy()
                out["cv"] = ref.predict_cv(obsv, 12).numpy()
    save(name, **out)


def train_case(name, data, batch_size, epochs, k_test, seed, weight_seed=0, unroll=1, full_weights=True):
    """Run the reference train() for `epochs` epochs then test(k_test); record everything observable."""
    out = dict(seed=np.array([seed, weight_seed]), epochs=np.int64(epochs), batch_size=np.int64(batch_size),
               k_test=np.int64(k_test), unroll=np.int64(unroll))
    for social in (True, False):
        tag = "soc" if social else "nos"
        ref = rh.Reference(data, batch_size=batch_size, use_social=social, weight_seed=weight_seed, unroll=unroll)
        if social:
            out.update({f"w0.{k}": v for k, v in ref.state().items()})
        tap = LossTap()
        ref.ns["mse_loss"] = tap
        np.random.seed(seed)
        torch.manual_seed(seed)
        import io
        import contextlib
        lines = io.StringIO()
        with contextlib.redirect_stdout(lines):
            for ep in range(1, epochs + 1):
                ref.ns["epoch"] = ep
                ref.train()
            ref.test(k_test)
        txt = [l for l in lines.getvalue().splitlines() if l.strip()]
        out[f"{tag}.stdout"] = np.array(txt)
        out[f"{tag}.mse_values"] = np.array(tap.values)
        st = ref.state()
        if full_weights and social:
            out.update({f"w1.{k}": v for k, v in st.items()})
        out[f"{tag}.w1_sum"] = np.array([float(v.astype(np.float64).sum()) for v in st.values()])
        out[f"{tag}.w1_l2"] = np.array([float(np.sqrt((v.astype(np.float64) ** 2).sum())) for v in st.values()])
        out[f"{tag}.w1_names"] = np.array(list(st.keys()))
        # exact train / test metrics (un-rounded) via a second identical run of test() internals:
        # the print lines carry 3 decimals only, so also record a noise-pinned K-sample run
        torch.manual_seed(seed + 1)
        rows = []
        preds_first = None
        with torch.no_grad():
            for a, b in ref.test_batches:
                a, b = int(a), int(b)
                obsv, pred = ref.dataset_obsv[a:b], ref.dataset_pred[a:b]
                errs = []
                for k in range(k_test):
                    noise = torch.rand(b - a, 32)
                    hat = ref.predict(obsv, noise, ref.n_next)
                    errs.append((((hat[:, :, :2] - pred) / ref.ss) ** 2).sum(dim=2).sqrt())
                    if preds_first is None:
                        preds_first = hat.numpy().copy()
                e = torch.stack(errs)
                rows.append([e.mean(2).mean(0).sum().item(), e[:, :, -1].mean(0).sum().item(),
                             e.mean(2).min(0)[0].sum().item(), e[:, :, -1].min(0)[0].sum().item()])
        out[f"{tag}.test_metrics"] = np.sum(np.array(rows), axis=0) / ref.n_test_samples
        out[f"{tag}.test_first_pred"] = preds_first
    save(name, **out)


if __name__ == "__main__":
    toy_sets()
    ops_case("ops_ragged.npz", [1, 2, 6, 8, 5, 3], weight_seed=0, data_seed=3, noise_seed=7)
    ops_case("ops_zara.npz", [32, 33], weight_seed=1, data_seed=4, noise_seed=8)
    train_case("train_toy_216.npz", rh.reference_toy(216, 6), batch_size=64, epochs=2, k_test=20, seed=11)
    train_case("train_ragged.npz", rh.synthetic_scenes([1, 2, 6, 8, 5, 3, 4, 7, 2, 9], seed=5), batch_size=16,
               epochs=2, k_test=5, seed=12, full_weights=False)
    train_case("train_unroll0.npz", rh.synthetic_scenes([4, 4, 4, 4, 4], seed=6), batch_size=8,
               epochs=1, k_test=3, seed=13, unroll=0, full_weights=False)
