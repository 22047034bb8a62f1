"""CPU ORACLE for the Social Ways hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may
import this file, and only as the checker / the CPU arm.  The product path (socialways_b200/)
never imports it and has no CPU fallback.

What it is: a functional, torch-CPU, fp32 restatement of the reference algorithm
(/root/reference/train.py and helpers, commit 0b13f2d), written from the maths, each function
citing the reference file:line it follows.  Weights are a flat {name: tensor} dict keyed exactly
like the reference modules' state_dicts, prefixed with the module-global name that train.py uses
(`encoder.`, `feature_embedder.`, `attention.`, `decoder.`, `D.`), so reference checkpoints map 1:1.

Third-party arithmetic: the reference calls torch (`nn.LSTM`, `nn.Linear`, `softmax`, `Adam`,
version un-pinned, README.md:83).  The LSTM cell, the MLPs, the social features and the attention
pooling are restated here explicitly with matmul/elementwise ops; Adam stays `torch.optim.Adam`
(the very code the reference calls, train.py:381,385).

PARITY PIN: the reference ships no tests, golden vectors or fixtures, so the pin is
the reference ITSELF, executed unmodified in the build container by tests/golden/reference_harness.py
(AST-lifted train.py / create_toy.py) with recorded seeds; its outputs are committed as
tests/golden/*.npz by tests/golden/make_golden.py and tests/test_oracle_golden.py checks every
function below against them (forward values, gradients, post-Adam weights, ADE/FDE print values).
"""
import copy
import math

import numpy as np
import torch

GEN_PREFIXES = ("attention.", "feature_embedder.", "encoder.", "decoder.")   # train.py:379-380 order


# --------------------------------------------------------------------------------------------
# data helpers
# --------------------------------------------------------------------------------------------
def toy_samples(n_samples, n_conditions, n_modes=3, n_per_batch=6, seed=30):
    """create_toy.py:11-54 (sample generator) + :162-187 (scene packing), numpy-2 safe.

    One uniform draw per turn angle, two per sample, in sample order, from np.random seeded with 30
    (create_toy.py:145).  `n_conditions / n_per_batch` is a true division (create_toy.py:18), which
    for 8 conditions gives fractional time stamps -- reproduced, not fixed.
    """
    rs = np.random.RandomState(seed)
    per_cond = n_samples // n_conditions
    pts = np.zeros((n_samples, 4, 2), dtype=np.float64)
    t_first = []
    for s in range(n_samples):
        way = (s * n_conditions) // n_samples
        t0 = s % per_cond + (way % (n_conditions / n_per_batch)) * per_cond
        ang = way * (2.0 * np.pi / n_conditions)
        turn = ((s % n_modes) - n_modes // 2) * 16 * np.pi / 180
        d2 = (rs.rand() - 0.5) * 4 * np.pi / 180
        d3 = (rs.rand() - 0.5) * 6 * np.pi / 180
        for q, (radius, a) in enumerate(((4, ang), (3, ang), (2, ang + turn + d2), (1, ang + turn + d2 + d3))):
            pts[s, q, 0] = np.cos(a) * radius
            pts[s, q, 1] = np.sin(a) * radius
        t_first.append(t0 * 4)
    pts = pts / 4
    groups = {}
    for s, t in enumerate(t_first):                        # dict keeps first-seen order (:162-166)
        groups.setdefault(t, []).append(s)
    order, batches, start = [], [], 0
    for members in groups.values():
        batches.append([start, start + len(members)])
        start += len(members)
        order.extend(members)
    order = np.array(order)
    return dict(obsvs=pts[order, :2].astype(np.float32), preds=pts[order, 2:].astype(np.float32),
                times=np.array([t_first[s] for s in order]).astype(np.int32), batches=np.array(batches))


class IsoScale:
    """utils/parse_utils.py:11-76 with keep_ratio=True: one isotropic scale, per-axis shift."""

    def __init__(self, obsv, pred):                        # train.py:113-118
        self.min_x = min(obsv[..., 0].min(), pred[..., 0].min())
        self.max_x = max(obsv[..., 0].max(), pred[..., 0].max())
        self.min_y = min(obsv[..., 1].min(), pred[..., 1].min())
        self.max_y = max(obsv[..., 1].max(), pred[..., 1].max())
        self.sx = self.sy = min(1 / (self.max_x - self.min_x), 1 / (self.max_y - self.min_y))

    def normalize(self, a):
        out = np.array(a, copy=True)
        out[..., 0] = (a[..., 0] - self.min_x) * self.sx
        out[..., 1] = (a[..., 1] - self.min_y) * self.sy
        return out

    def denormalize(self, a):
        out = np.array(a, copy=True)
        out[..., 0] = a[..., 0] / self.sx + self.min_x
        out[..., 1] = a[..., 1] / self.sy + self.min_y
        return out


def traj_4d(obsv_p, pred_p=None):
    """train.py:130-138: append finite-difference velocities; v_0 := v_1 for the observed part,
    and the first predicted velocity is measured from the last observed position."""
    dv = obsv_p[:, 1:] - obsv_p[:, :-1]
    obsv_4d = torch.cat([obsv_p, torch.cat([dv[:, :1], dv], dim=1)], dim=2)
    if pred_p is None:
        return obsv_4d
    prev = torch.cat([obsv_p[:, -1:], pred_p[:, :-1]], dim=1)
    return obsv_4d, torch.cat([pred_p, pred_p - prev], dim=2)


def constant_velocity(obsv, n_next):
    """utils/linear_models.py:9-20: every future step adds the SAME velocity to the newest point."""
    vel = (obsv[:, -1] - obsv[:, -3]) / 2.0 if obsv.shape[1] > 2 else obsv[:, -1] - obsv[:, -2]
    steps = torch.arange(1, n_next + 1, dtype=obsv.dtype).view(1, -1, 1)
    return obsv[:, -1:].clone() + steps * vel.unsqueeze(1)


# --------------------------------------------------------------------------------------------
# weights
# --------------------------------------------------------------------------------------------
def init_weights(hidden=64, n_next=12, n_latent=2, seed=0):
    """Same shapes, names and init RNG stream as train.py:370-384 (encoder, feature_embedder,
    attention, decoder, D constructed in that order after torch.manual_seed(seed))."""
    import torch.nn as nn
    torch.manual_seed(seed)
    mods = {}
    enc = nn.Module()
    enc.embed = nn.Linear(4, hidden)
    enc.lstm = nn.LSTM(hidden, hidden, num_layers=1, batch_first=True)
    mods["encoder"] = enc
    fe = nn.Module()
    fe.fc = nn.Sequential(nn.Linear(3, 32), nn.ReLU(), nn.Linear(32, 64), nn.ReLU(), nn.Linear(64, hidden))
    mods["feature_embedder"] = fe
    att = nn.Module()
    att.W = nn.Linear(hidden, hidden)
    mods["attention"] = att
    dec = nn.Module()
    d = hidden + hidden + hidden // 2
    dec.fc1 = nn.Sequential(nn.Linear(d, d), nn.LeakyReLU(0.2), nn.Linear(d, d // 2), nn.LeakyReLU(0.2),
                            nn.Linear(d // 2, d // 4), nn.Linear(d // 4, 2))
    mods["decoder"] = dec
    D = nn.Module()
    D.obsv_encoder_lstm = nn.LSTM(4, hidden, batch_first=True)
    h2 = hidden // 2
    D.obsv_encoder_fc = nn.Sequential(nn.Linear(hidden, h2), nn.LeakyReLU(0.2), nn.Linear(h2, h2))
    D.pred_encoder = nn.Sequential(nn.Linear(n_next * 4, h2), nn.LeakyReLU(0.2), nn.Linear(h2, h2))
    D.classifier = nn.Sequential(nn.Linear(hidden, h2), nn.LeakyReLU(0.2), nn.Linear(h2, 1))
    D.latent_decoder = nn.Sequential(nn.Linear(hidden, h2), nn.LeakyReLU(0.2), nn.Linear(h2, n_latent))
    mods["D"] = D
    out = {}
    for tag, m in mods.items():
        for k, v in m.state_dict().items():
            out[f"{tag}.{k}"] = v.detach().clone()
    return out


def _lin(P, name, x):
    return x @ P[name + ".weight"].t() + P[name + ".bias"]


def _lrelu(x, slope=0.2):
    return torch.where(x > 0, x, x * slope)


# --------------------------------------------------------------------------------------------
# operators
# --------------------------------------------------------------------------------------------
def lstm_cell(x, h, c, w_ih, w_hh, b_ih, b_hh):
    """One step of torch.nn.LSTM (the op behind train.py:254,268 and :278,299); gate rows i,f,g,o."""
    gates = x @ w_ih.t() + b_ih + h @ w_hh.t() + b_hh
    i, f, g, o = gates.chunk(4, dim=1)
    c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
    h = torch.sigmoid(o) * torch.tanh(c)
    return h, c


def encoder_steps(P, x4_seq, h, c):
    """EncoderLstm.forward (train.py:262-269): Linear(4,H) embedding then the LSTM, for a whole
    [N,T,4] sequence or a single [N,4] step; returns the carried state (train.py:268)."""
    if x4_seq.dim() == 2:
        x4_seq = x4_seq.unsqueeze(1)
    for t in range(x4_seq.shape[1]):
        e = _lin(P, "encoder.embed", x4_seq[:, t])
        h, c = lstm_cell(e, h, c, P["encoder.lstm.weight_ih_l0"], P["encoder.lstm.weight_hh_l0"],
                         P["encoder.lstm.bias_ih_l0"], P["encoder.lstm.bias_hh_l0"])
    return h, c


def social_features(x_last):
    """SocialFeatures / BearingMTX / DCA_MTX (train.py:208-241) on the last observed state
    x=(px,py,vx,vy) of ALL agents of the mini-batch: out[i,j] = (distance, bearing cosine of j
    seen from i's heading, distance at closest approach), D[i,j] = x_i - x_j (train.py:232-234)."""
    d = x_last.unsqueeze(1) - x_last.unsqueeze(0)
    dp, dv = d[..., :2], d[..., 2:]
    dist = dp.norm(dim=2)
    v_i = x_last[:, 2:].unsqueeze(1).expand(-1, x_last.shape[0], -1)
    bearing = (dp * v_i).sum(-1) / (dist * v_i.norm(dim=2) + 1e-6)               # :224-225
    ttca = -((dp * dv).sum(-1) / ((dv * dv).sum(-1) + 1e-6))                     # :211-213
    dca = (dp + ttca.unsqueeze(-1) * dv).norm(dim=2)                             # :214-217
    return torch.stack([dist, bearing, dca], dim=2)


def embed_features(P, f):
    """EmbedSocialFeatures.fc (train.py:183-188): 3 -> 32 ReLU -> 64 ReLU -> H on every pair."""
    a = torch.relu(_lin(P, "feature_embedder.fc.0", f))
    a = torch.relu(_lin(P, "feature_embedder.fc.2", a))
    return _lin(P, "feature_embedder.fc.4", a)


def attention_pool_loop(P, emb, h, scenes):
    """AttentionPooling.forward (train.py:160-175) in the reference's own evaluation order: one
    agent at a time, scores via a batched dot product, self score forced to -1000, softmax over
    the scene, weighted sum of the RAW hidden states.  This is the form the CPU baseline times."""
    wh = _lin(P, "attention.W", h)
    pooled = torch.zeros_like(h)
    for a, b in scenes:
        a, b = int(a), int(b)
        if b - a == 1:
            continue
        for i in range(a, b):
            score = torch.bmm(emb[i, a:b].unsqueeze(1), wh[a:b].unsqueeze(2)).view(-1).clone()
            score[i - a] = -1000
            w = torch.softmax(score, dim=0)
            pooled[i] = w.view(1, -1) @ h[a:b]
    return pooled


def attention_pool_closed(P, x_last, h, scenes):
    """Second oracle: per-scene closed form of train.py:208-241,183-188,160-175;
    never forms the cross-scene blocks the reference computes and discards."""
    wh = _lin(P, "attention.W", h)
    pooled = torch.zeros_like(h)
    for a, b in scenes:
        a, b = int(a), int(b)
        if b - a == 1:
            continue
        emb = embed_features(P, social_features(x_last[a:b]))
        score = (emb * wh[a:b].unsqueeze(0)).sum(-1)
        score = score.masked_fill(torch.eye(b - a, dtype=torch.bool), -1000.0)
        pooled[a:b] = torch.softmax(score, dim=1) @ h[a:b]
    return pooled


def decoder_fc(P, h, s, z):
    """DecoderFC (train.py:320-335): cat[h,s,z] -> 160 LReLU -> 80 LReLU -> 40 -> 2 (the last two
    Linear layers have no activation between them, train.py:327-328)."""
    a = _lrelu(_lin(P, "decoder.fc1.0", torch.cat([h, s, z], dim=1)))
    a = _lrelu(_lin(P, "decoder.fc1.2", a))
    return _lin(P, "decoder.fc1.5", _lin(P, "decoder.fc1.4", a))


def predict(P, obsv_p, noise, n_next, scenes=None, use_social=True, pool="loop"):
    """predict() (train.py:392-432).  Encode the observation from a zero state, pool once on the
    post-observation state, then n_next x {decode velocity, integrate, feed (p,v) back through one
    encoder step}.  The reference's extra encoder step after the last prediction (:430) has no
    observable effect and is omitted.  Returns [N, n_next, 4] = (p, v) per step."""
    n = obsv_p.shape[0]
    hdim = P["attention.W.weight"].shape[1]
    x4 = traj_4d(obsv_p)
    h, c = encoder_steps(P, x4, torch.zeros(n, hdim), torch.zeros(n, hdim))
    if scenes is None or len(scenes) == 0:
        scenes = [[0, n]]                                                        # :405-406
    if use_social:
        if pool == "loop":
            emb = embed_features(P, social_features(x4[:, -1]))                 # N x N, cross-scene too (:229-241)
            s = attention_pool_loop(P, emb, h, scenes)
        else:
            s = attention_pool_closed(P, x4[:, -1], h, scenes)
    else:
        s = torch.zeros_like(h)                                                  # :413
    last = x4[:, -1]
    out = []
    for t in range(n_next):
        v = decoder_fc(P, h, s, noise)
        last = torch.cat([v + last[:, :2], v], dim=1)                            # :423-425
        out.append(last)
        if t + 1 < n_next:
            h, c = encoder_steps(P, last, h, c)
    return torch.stack(out, dim=1)


def discriminator(P, obsv_4d, pred_4d):
    """Discriminator.forward (train.py:294-309): LSTM(4->H) over the observation from a zero state,
    FC on its last output; FC on the flattened prediction; classifier and InfoGAN code heads."""
    n = obsv_4d.shape[0]
    hdim = P["D.obsv_encoder_lstm.weight_hh_l0"].shape[1]
    h, c = torch.zeros(n, hdim), torch.zeros(n, hdim)
    for t in range(obsv_4d.shape[1]):
        h, c = lstm_cell(obsv_4d[:, t], h, c, P["D.obsv_encoder_lstm.weight_ih_l0"],
                         P["D.obsv_encoder_lstm.weight_hh_l0"], P["D.obsv_encoder_lstm.bias_ih_l0"],
                         P["D.obsv_encoder_lstm.bias_hh_l0"])
    oc = _lin(P, "D.obsv_encoder_fc.2", _lrelu(_lin(P, "D.obsv_encoder_fc.0", h)))
    pc = _lin(P, "D.pred_encoder.2", _lrelu(_lin(P, "D.pred_encoder.0", pred_4d.reshape(n, -1))))
    both = torch.cat([oc, pc], dim=1)
    label = _lin(P, "D.classifier.2", _lrelu(_lin(P, "D.classifier.0", both)))
    code = _lin(P, "D.latent_decoder.2", _lrelu(_lin(P, "D.latent_decoder.0", both)))
    return label, code


def mse(a, b):
    return ((a - b) ** 2).mean()                                                 # nn.MSELoss, train.py:386


# --------------------------------------------------------------------------------------------
# training step / evaluation
# --------------------------------------------------------------------------------------------
class OracleTrainer:
    """train() (train.py:439-560) and test() (train.py:563-616) over a weight dict."""

    def __init__(self, weights, data, batch_size=64, use_social=True, unroll=1, lr_g=1e-4, lr_d=1e-3,
                 pool="loop"):
        self.P = {k: v.clone().requires_grad_(True) for k, v in weights.items()}
        self.use_social, self.unroll, self.batch_size, self.pool = use_social, unroll, batch_size, pool
        obsv, pred = np.asarray(data["obsvs"], np.float32), np.asarray(data["preds"], np.float32)
        self.scale = IsoScale(obsv, pred)
        self.ss = self.scale.sx
        self.obsv = torch.from_numpy(self.scale.normalize(obsv))
        self.pred = torch.from_numpy(self.scale.normalize(pred))
        self.times = data["times"]
        scenes = np.asarray(data["batches"])
        self.train_size = max(1, (len(scenes) * 4) // 5)                         # :95
        self.n_train = int(scenes[self.train_size - 1][1])
        self.n_test = self.obsv.shape[0] - self.n_train
        if self.n_test == 0:                                                     # :107-109
            self.n_test = 1
            scenes = np.array([scenes[0], scenes[0]])
        self.scenes = scenes
        self.n_next = self.pred.shape[1]
        self.noise_len = self.P["attention.W.weight"].shape[1] // 2
        gen = [self.P[k] for pre in GEN_PREFIXES for k in self.P if k.startswith(pre)]
        self.opt_g = torch.optim.Adam(gen, lr=lr_g, betas=(0.9, 0.999))          # :381
        self.d_keys = [k for k in self.P if k.startswith("D.")]
        self.opt_d = torch.optim.Adam([self.P[k] for k in self.d_keys], lr=lr_d, betas=(0.9, 0.999))  # :385
        self.log = []

    def minibatches(self):
        """train.py:446-461: pack whole scenes until the next one would overflow batch_size agents."""
        group, count = [], 0
        for i in range(self.train_size):
            a, b = self.scenes[i]
            group.append((int(a), int(b)))
            count += int(b - a)
            if i >= self.train_size - 1 or count + int(self.scenes[i + 1][1] - self.scenes[i + 1][0]) > self.batch_size:
                lo = group[0][0]
                yield lo, group[-1][1], [(a - lo, b - lo) for a, b in group]
                group, count = [], 0

    def _predict(self, obsv, noise, scenes):
        return predict(self.P, obsv, noise, self.n_next, scenes, self.use_social, self.pool)

    def train_epoch(self):
        ade = fde = 0.0
        for lo, hi, scenes in self.minibatches():
            obsv, pred = self.obsv[lo:hi], self.pred[lo:hi]
            bs = hi - lo
            obsv_4d, pred_4d = traj_4d(obsv, pred)
            zeros = torch.zeros(bs, 1) + np.random.uniform(0, 0.1)               # :471
            ones = torch.ones(bs, 1) * np.random.uniform(0.9, 1.0)               # :472
            noise = torch.rand(bs, self.noise_len)                              # :473
            backup = None
            for u in range(self.unroll + 1):                                     # :476-499
                self.opt_d.zero_grad(set_to_none=True)
                with torch.no_grad():
                    fake = self._predict(obsv, noise, scenes)
                fake_lab, code = discriminator(self.P, obsv_4d, fake)
                d_loss = mse(fake_lab, zeros)
                d_info = mse(code.squeeze(), noise[:, :2])
                real_lab, _ = discriminator(self.P, obsv_4d, pred_4d)
                d_loss = d_loss + mse(real_lab, ones) + 0.5 * d_info
                d_loss.backward()
                self.opt_d.step()
                if u == 0 and self.unroll > 0:
                    backup = {k: self.P[k].detach().clone() for k in self.d_keys}
            self.opt_d.zero_grad(set_to_none=True)                               # :503-505
            self.opt_g.zero_grad(set_to_none=True)
            gen = self._predict(obsv, noise, scenes)
            gen_lab, code = discriminator(self.P, obsv_4d, gen)
            g_fool = mse(gen_lab, ones)
            g_info = mse(code.squeeze(), noise[:, :2])
            g_loss = g_fool + 0.5 * g_info                                       # :520-523
            g_loss.backward()
            self.opt_g.step()
            if backup is not None:                                               # :541-543 + :311-316
                with torch.no_grad():
                    for k in self.d_keys:
                        if "lstm" not in k:                                      # Linear layers only
                            self.P[k].copy_(backup[k])
            with torch.no_grad():                                                # :546-551
                err = (((gen[:, :, :2] - pred) / self.ss) ** 2).sum(dim=2).sqrt()
                ade += err.sum().item() / self.n_next
                fde += err[:, -1].sum().item()
            self.log.append(dict(d_loss=d_loss.item(), g_fool=g_fool.item(), g_info=g_info.item()))
        return ade / self.n_train, fde / self.n_train

    def test_epoch(self, n_gen_samples=20, just_one=False, noise_fn=None, collect=False):
        """test() (train.py:563-616): one scene at a time, K serial predict() calls each with a fresh
        torch.rand noise, errors in metres (divided by ss); avg-over-K and min-over-K ADE/FDE."""
        acc = np.zeros(4)
        dumps = []
        test_scenes = self.scenes[self.train_size:]
        with torch.no_grad():
            for a, b in test_scenes:
                a, b = int(a), int(b)
                obsv, pred = self.obsv[a:b], self.pred[a:b]
                errs, preds = [], []
                for k in range(n_gen_samples):
                    noise = torch.rand(b - a, self.noise_len) if noise_fn is None else noise_fn(a, b, k)
                    hat = self._predict(obsv, noise, None)
                    errs.append((((hat[:, :, :2] - pred) / self.ss) ** 2).sum(dim=2).sqrt())
                    preds.append(hat)
                e = torch.stack(errs)                                            # [K, A, T]
                acc += [e.mean(2).mean(0).sum().item(), e[:, :, -1].mean(0).sum().item(),
                        e.mean(2).min(0)[0].sum().item(), e[:, :, -1].min(0)[0].sum().item()]
                if collect:
                    dumps.append(dict(timestamp=self.times[a], obsvs=obsv, preds_our=torch.stack(preds)[..., :2],
                                      preds_gtt=pred, preds_lnr=constant_velocity(obsv, self.n_next)))
                if just_one:
                    break
        ade_avg, fde_avg, ade_min, fde_min = acc / self.n_test
        out = dict(ade_avg=ade_avg, fde_avg=fde_avg, ade_min=ade_min, fde_min=fde_min)
        return (out, dumps) if collect else out

    def weights(self):
        return {k: v.detach().clone() for k, v in self.P.items()}
