"""Host-side mirror of the reference's operator interface for the hot path (reference train.py).

Same names, constructor arguments, state_dict keys and call semantics as the reference classes, so
a reference checkpoint (train.py:653-663) loads unchanged and train.py's call sites read the same;
the arithmetic runs in the sm_100a kernels behind the C-ABI (include/socialways_b200.h).  There is
no CPU path: every forward raises if its inputs are not CUDA tensors or the library is missing.

Reference surface mirrored here:
    get_traj_4d            train.py:130-138        EncoderLstm           train.py:245-269
    SocialFeatures (+MTX)  train.py:208-241        EmbedSocialFeatures   train.py:178-189
    AttentionPooling       train.py:153-175        DecoderFC             train.py:320-335
    DecoderLstm (dead)     train.py:339-366        Discriminator         train.py:272-316
    predict                train.py:392-432        predict_cv            utils/linear_models.py:9-20
The reference has no `Generator` class: `Generator` below owns the four generator
modules under the names the reference uses for its module-level globals, so state_dict keys are
`encoder.*`, `feature_embedder.*`, `attention.*`, `decoder.*`.
"""
import torch
import torch.nn as nn

from . import ops, packing
from ._lib import SocialWaysCudaError

HIDDEN = 64


def _require_path_sizes(hidden):
    if hidden != HIDDEN:
        raise SocialWaysCudaError(
            f"the sm_100a kernels are built for hidden_size={HIDDEN} (train.py:43-45 default); got {hidden}")


def get_traj_4d(obsv_p, pred_p):
    """train.py:130-138.  Plain tensor ops (4-6 tiny launches, not on the fused path: the kernels
    form the velocities on the fly); kept for call-site parity with train()/test()."""
    obsv_v = obsv_p[:, 1:] - obsv_p[:, :-1]
    obsv_4d = torch.cat([obsv_p, torch.cat([obsv_v[:, :1], obsv_v], dim=1)], dim=2)
    if len(pred_p) == 0:
        return obsv_4d
    prev = torch.cat([obsv_p[:, -1:], pred_p[:, :-1]], dim=1)
    return obsv_4d, torch.cat([pred_p, pred_p - prev], dim=2)


def predict_cv(obsv, n_next):
    """utils/linear_models.py:9-20 constant-velocity baseline used by test() (train.py:577)."""
    vel = (obsv[:, -1] - obsv[:, -3]) / 2.0 if obsv.shape[1] > 2 else obsv[:, -1] - obsv[:, -2]
    steps = torch.arange(1, n_next + 1, dtype=obsv.dtype, device=obsv.device).view(1, -1, 1)
    return obsv[:, -1:] + steps * vel.unsqueeze(1)


def SocialFeatures(x, sub_batches=None):
    """train.py:229-241: the dense [N,N,3] feature matrix over the whole mini-batch.  Compatibility
    helper only -- predict() never materialises it (fused into sw_pool_fwd)."""
    xl = x[:, -1]
    d = xl.unsqueeze(1) - xl.unsqueeze(0)
    dp, dv = d[..., :2], d[..., 2:]
    dist = dp.norm(dim=2)
    vi = xl[:, 2:].unsqueeze(1).expand(-1, xl.shape[0], -1)
    bearing = (dp * vi).sum(-1) / (dist * vi.norm(dim=2) + 1e-6)
    ttca = -((dp * dv).sum(-1) / ((dv * dv).sum(-1) + 1e-6))
    dca = (dp + ttca.unsqueeze(-1) * dv).norm(dim=2)
    return torch.stack([dist, bearing, dca], dim=2)


class EncoderLstm(nn.Module):
    """train.py:245-269.  forward() runs sw_lstm_seq_fwd from the carried state (train.py:268)."""

    def __init__(self, hidden_size, n_layers=2):
        super().__init__()
        self.hidden_size = hidden_size
        self.embed = nn.Linear(4, hidden_size)
        self.lstm = nn.LSTM(hidden_size, hidden_size, num_layers=n_layers, batch_first=True)
        self.lstm_h = []
        self.n_layers = n_layers

    def init_lstm(self, h, c):
        self.lstm_h = (h, c)

    def packed(self):
        _require_path_sizes(self.hidden_size)
        if self.n_layers != 1:
            raise SocialWaysCudaError("the path uses n_lstm_layers = 1 (train.py:82)")
        return packing.pack_encoder(self.embed.weight, self.embed.bias, self.lstm.weight_ih_l0,
                                    self.lstm.weight_hh_l0, self.lstm.bias_ih_l0, self.lstm.bias_hh_l0)

    def forward(self, obsv):
        bs = obsv.shape[0]
        x = obsv.reshape(bs, -1, 4)
        h_in = c_in = None
        if len(self.lstm_h) == 2:
            h_in, c_in = self.lstm_h[0].reshape(bs, -1), self.lstm_h[1].reshape(bs, -1)
        with torch.no_grad():
            r = ops.lstm_seq(self.packed(), x, h_in=h_in, c_in=c_in, want_y=True)
        self.lstm_h = (r["h"].unsqueeze(0), r["c"].unsqueeze(0))
        return r["y"]


class EmbedSocialFeatures(nn.Module):
    """train.py:178-189.  forward() on a materialised feature tensor is a compatibility path."""

    def __init__(self, input_size, hidden_size):
        super().__init__()
        self.input_size = input_size
        self.hidden_size = hidden_size
        self.fc = nn.Sequential(nn.Linear(input_size, 32), nn.ReLU(), nn.Linear(32, 64), nn.ReLU(),
                                nn.Linear(64, hidden_size))

    def forward(self, ftr_list, sub_batches=None):
        return self.fc(ftr_list)


class AttentionPooling(nn.Module):
    """train.py:153-175.  forward() on materialised embeddings is a compatibility path (per-scene
    closed form); predict() uses the fused sw_pool_fwd kernel instead."""

    def __init__(self, h_dim, f_dim):
        super().__init__()
        self.f_dim = f_dim
        self.h_dim = h_dim
        self.W = nn.Linear(h_dim, f_dim, bias=True)

    def forward(self, f, h, sub_batches):
        wh = self.W(h)
        out = torch.zeros_like(h)
        for sb in sub_batches:
            a, b = int(sb[0]), int(sb[1])
            if b - a == 1:
                continue
            sig = (f[a:b, a:b] * wh[a:b].unsqueeze(0)).sum(-1)
            sig = sig.masked_fill(torch.eye(b - a, dtype=torch.bool, device=h.device), -1000.0)
            out[a:b] = torch.softmax(sig, dim=1) @ h[a:b]
        return out


class DecoderFC(nn.Module):
    """train.py:320-335.  forward() = one decode step through sw_decode_fwd without LSTM feedback."""

    def __init__(self, hidden_dim):
        super().__init__()
        self.fc1 = nn.Sequential(nn.Linear(hidden_dim, hidden_dim), nn.LeakyReLU(0.2),
                                 nn.Linear(hidden_dim, hidden_dim // 2), nn.LeakyReLU(0.2),
                                 nn.Linear(hidden_dim // 2, hidden_dim // 4),
                                 nn.Linear(hidden_dim // 4, 2))

    def packed(self):
        f = self.fc1
        if f[0].weight.shape != (160, 160):
            raise SocialWaysCudaError("DecoderFC kernel is built for hidden+social+noise = 160 (train.py:375)")
        return packing.pack_decoder(f[0].weight, f[0].bias, f[2].weight, f[2].bias, f[4].weight, f[4].bias,
                                    f[5].weight, f[5].bias)

    def forward(self, h, s, z):
        # one-step decode: with zero "last position" the first emitted p equals v (train.py:421)
        n = h.shape[0]
        with torch.no_grad():
            zero_lstm = torch.zeros(69, 256, device=h.device)
            out = ops.decode(zero_lstm, self.packed(), h, torch.zeros_like(h), s, z.unsqueeze(0),
                             torch.zeros(n, 4, device=h.device), 1)
        return out[0, :, 0, 2:4]


class DecoderLstm(nn.Module):
    """train.py:339-366.  Dead surface in the reference (its construction is commented out,
    train.py:376); kept so `from ... import DecoderLstm` keeps working.  Not on the path."""

    def __init__(self, input_size, hidden_size):
        super().__init__()
        self.lstm = nn.LSTM(input_size, hidden_size, num_layers=1, batch_first=True)
        self.fc = nn.Sequential(nn.Linear(hidden_size, 64), nn.Sigmoid(), nn.Linear(64, 64), nn.LeakyReLU(0.2),
                                nn.Linear(64, 32), nn.LeakyReLU(0.2), nn.Linear(32, 2))
        self.lstm_h = []

    def init_lstm(self, h, c):
        self.lstm_h = (h, c)

    def forward(self, h, s, z):
        raise SocialWaysCudaError("DecoderLstm is not on the Social Ways path (train.py:375-376 uses DecoderFC)")


class Discriminator(nn.Module):
    """train.py:272-316.  The observation LSTM runs in sw_lstm_seq_fwd / sw_lstm_seq_bwd, the four FC heads
    (64->32->32, n_next*4->32->32, 64->32->1, 64->32->2) in sw_disc_heads_fwd / sw_disc_heads_bwd."""

    def __init__(self, n_next, hidden_dim, n_latent_code):
        super().__init__()
        self.lstm_dim = hidden_dim
        self.n_next = n_next
        self.obsv_encoder_lstm = nn.LSTM(4, hidden_dim, batch_first=True)
        self.obsv_encoder_fc = nn.Sequential(nn.Linear(hidden_dim, hidden_dim // 2), nn.LeakyReLU(0.2),
                                             nn.Linear(hidden_dim // 2, hidden_dim // 2))
        self.pred_encoder = nn.Sequential(nn.Linear(n_next * 4, hidden_dim // 2), nn.LeakyReLU(0.2),
                                          nn.Linear(hidden_dim // 2, hidden_dim // 2))
        self.classifier = nn.Sequential(nn.Linear(hidden_dim, hidden_dim // 2), nn.LeakyReLU(0.2),
                                        nn.Linear(hidden_dim // 2, 1))
        self.latent_decoder = nn.Sequential(nn.Linear(hidden_dim, hidden_dim // 2), nn.LeakyReLU(0.2),
                                            nn.Linear(self.lstm_dim // 2, n_latent_code))

    def packed_lstm(self):
        _require_path_sizes(self.lstm_dim)
        l = self.obsv_encoder_lstm
        return packing.pack_disc_lstm(l.weight_ih_l0, l.weight_hh_l0, l.bias_ih_l0, l.bias_hh_l0)

    def encode_obsv(self, obsv):
        """Last hidden state of the observation LSTM from a zero state (train.py:296-301)."""
        if not obsv.is_cuda:
            raise SocialWaysCudaError("Discriminator runs on CUDA tensors only (no CPU fallback)")
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.obsv_encoder_lstm.parameters()):
            from .autograd_path import LstmSeqFn
            return LstmSeqFn.apply(self.packed_lstm(), obsv)[0]
        with torch.no_grad():
            return ops.lstm_seq(self.packed_lstm(), obsv)["h"]

    def packed_heads(self):
        """Flat parameter pack of the 8 Linear layers, torch [out][in] layouts (csrc/disc_heads.cu)."""
        mods = (self.obsv_encoder_fc[0], self.obsv_encoder_fc[2], self.pred_encoder[0], self.pred_encoder[2],
                self.classifier[0], self.classifier[2], self.latent_decoder[0], self.latent_decoder[2])
        return torch.cat([t.reshape(-1) for m in mods for t in (m.weight, m.bias)])

    def heads(self, obsv_h, pred):
        """FC heads (train.py:301-309) in the fused kernels; label [N,1], code_hat [N,n_latent]."""
        pred2 = pred.reshape(-1, self.n_next * 4)
        n_latent = self.latent_decoder[2].out_features
        if torch.is_grad_enabled() and (pred2.requires_grad or obsv_h.requires_grad or
                                        any(p.requires_grad for p in self.parameters())):
            from .autograd_path import DiscHeadsFn
            return DiscHeadsFn.apply(self.packed_heads(), obsv_h.contiguous(), pred2.contiguous(), n_latent)
        with torch.no_grad():
            label, code, _ = ops.disc_heads_fwd(self.packed_heads(), obsv_h, pred2, n_latent)
        return label, code

    def forward(self, obsv, pred):
        return self.heads(self.encode_obsv(obsv), pred)

    def load(self, backup):
        """train.py:311-316: restores the nn.Linear layers ONLY (the LSTM is not rolled back)."""
        for m_from, m_to in zip(backup.modules(), self.modules()):
            if isinstance(m_to, nn.Linear):
                # same values as the reference's `.data = clone()` rebinding, written in place so that parameters that
                # live inside a flat optimiser buffer (fused_optim.FlatAdam) keep their storage
                m_to.weight.data.copy_(m_from.weight.data)
                if m_to.bias is not None:
                    m_to.bias.data.copy_(m_from.bias.data)


class Generator(nn.Module):
    """The four generator modules of train.py:370-376 + predict() (train.py:392-432)."""

    def __init__(self, hidden_size=64, n_lstm_layers=1, num_social_features=3, social_feature_size=None,
                 noise_len=None, use_social=False):
        super().__init__()
        social_feature_size = hidden_size if social_feature_size is None else social_feature_size
        self.noise_len = hidden_size // 2 if noise_len is None else noise_len
        self.encoder = EncoderLstm(hidden_size, n_lstm_layers)                                  # :370
        self.feature_embedder = EmbedSocialFeatures(num_social_features, social_feature_size)   # :371
        self.attention = AttentionPooling(hidden_size, social_feature_size)                     # :372
        self.decoder = DecoderFC(hidden_size + social_feature_size + self.noise_len)            # :375
        self.use_social = use_social                                                            # :83 default False
        # decode kernel of the no-grad path: "fp16x2" (tensor cores, fp32-faithful), "fp32" (FFMA) or "bf16" (fast)
        self.inference_precision = "fp16x2"
        self._pack_cache = None
        self._scene_cache = {}
        self._fp16_status = None        # device int32[1]: set by the fp16-split kernels when an operand left fp16's range

    def optimizer_parameters(self):
        """Parameter order of train.py:379-380 (attention, feature_embedder, encoder, decoder)."""
        from itertools import chain
        return chain(self.attention.parameters(), self.feature_embedder.parameters(),
                     self.encoder.parameters(), self.decoder.parameters())

    # ---- packed weights, cached on the parameters' version counters ----
    def packs(self):
        key = tuple((p.data_ptr(), p._version) for p in self.parameters())
        if self._pack_cache is None or self._pack_cache[0] != key:
            with torch.no_grad():
                fe, att = self.feature_embedder.fc, self.attention.W
                m, m0 = packing.pool_agent_matrix(att.weight, att.bias, fe[4].weight, fe[4].bias)
                packs = dict(enc=self.encoder.packed(), dec=self.decoder.packed(),
                             pool=packing.pack_pool(fe[0].weight, fe[0].bias, fe[2].weight, fe[2].bias),
                             pool_m=m.contiguous(), pool_m0=m0.contiguous())
                packs["tc_w16"], packs["tc_f32"] = packing.pack_decoder_tc(packs["enc"], packs["dec"])
                packs["tcx"] = packing.pack_decoder_tcx(packs["enc"], packs["dec"])
                packs["pair"] = packing.pack_decoder_pair(packs["enc"], packs["dec"])
                packs["pair_bf16"] = packing.pack_decoder_pair(packs["enc"], packs["dec"], bf16=True)
                packs["enc_tcx"] = packing.pack_encoder_tcx(packs["enc"])
                packs["pool_tcx"] = packing.pack_pool_tcx(fe[2].weight)
                # weights beyond fp16's range (|w| > 65 504 -> inf in the hi part) poison the split: raise the status word now
                bad = torch.stack([torch.isinf(t).any() for t in (*packs["tcx"][:2], packs["enc_tcx"][0], packs["pool_tcx"])]).any()
                self._status_word(packs["enc"].device).bitwise_or_(bad.to(torch.int32))
            self._pack_cache = (key, packs)
        return self._pack_cache[1]

    def _status_word(self, device):
        if self._fp16_status is None or self._fp16_status.device != device:
            self._fp16_status = torch.zeros(1, dtype=torch.int32, device=device)
        return self._fp16_status

    def fp16_overflowed(self, reset=True):
        """True when a tensor-core (fp16 hi/lo split) kernel saw a weight, state or activation outside fp16's exponent
        range since the last reset: its output for those inputs is not trustworthy -- use precision="fp32" (the FFMA
        kernels; trainer.test() does that automatically).  Synchronises."""
        if self._fp16_status is None:
            return False
        flag = bool(self._fp16_status.item())
        if reset and flag:
            self._fp16_status.zero_()
            self._pack_cache = None          # a weight-range flag is raised again when the packs are rebuilt
        return flag

    def invalidate_packs(self):
        """Drop the packed-weight cache: call after anything that rewrites the parameters without advancing their
        version counters (CUDA-graph replay of an optimiser step, raw-pointer kernels)."""
        self._pack_cache = None

    def scene_index(self, sub_batches, n_agents, device):
        if isinstance(sub_batches, ops.SceneIndex):
            return sub_batches
        import numpy as np
        sb = np.asarray(sub_batches, dtype=np.int64)
        key = (n_agents, sb.tobytes())
        if key not in self._scene_cache:
            if len(self._scene_cache) > 64:
                self._scene_cache.clear()
            self._scene_cache[key] = ops.SceneIndex(sb, n_agents, device)
        return self._scene_cache[key]

    @torch.no_grad()
    def predict_k(self, obsv_p, noise, n_next, sub_batches=(), out=None, precision=None, seed=None, k=None, noise_buf=None):
        """K-sample predict(): noise [K,N,32] -> [K,N,n_next,4].  The observation encoding and the
        pooled social vector do not depend on the sample and are computed once.
        precision: "fp32" = FFMA decode kernel; "fp16x2" = tcgen05 kernel on fp16 hi/lo split operands
        (fp32-faithful, ~1e-6); "bf16" / "bf16p" = tcgen05 kernels on bf16 operands (fast modes, ~3e-3; "bf16p" = the
        CTA-pair kernel of "fp16x2" with one MMA per product).
        noise=None, seed=s, k=K: the K x N x 32 latent noise is drawn ON THE DEVICE (Philox4x32-10, sw_noise_uniform) instead
        of being supplied by the caller -- same distribution as the reference's torch.rand (train.py:584), a different
        stream, and no 128 B / trajectory upload; `seed` may be an int or (seed, offset)."""
        if not obsv_p.is_cuda:
            raise SocialWaysCudaError("predict() runs on CUDA tensors only (no CPU fallback)")
        precision = precision or self.inference_precision
        pk = self.packs()
        n = obsv_p.shape[0]
        if noise is None:
            if seed is None or k is None:
                raise ValueError("predict_k: give `noise`, or `seed` and `k` for device-side noise")
            sd, off = (seed if isinstance(seed, tuple) else (seed, 0))
            noise = ops.noise_uniform((k, n, self.noise_len), obsv_p.device, sd, off, out=noise_buf)
        # "fp16x2" decodes with two tiles in flight per SM (CTA pairs, csrc/decode_fwd_pair.cu); "fp16x2s" = the same arithmetic
        # on the one-tile-per-SM kernel (csrc/decode_fwd_tcx.cu)
        # "bf16p" = the CTA-pair kernel on single bf16 operands (encoder and pooling as in "fp16x2"); "bf16" = the older one-tile
        # bf16 kernel with the FFMA encoder / pooling
        single, pair_bf16 = precision == "fp16x2s", precision == "bf16p"
        if single or pair_bf16:
            precision = "fp16x2"
        if precision == "fp16x2":           # both recurrent kernels on the tensor cores (fp16 hi/lo split operands)
            enc = ops.lstm_seq_tcx(*pk["enc_tcx"], obsv_p)
        else:
            enc = ops.lstm_seq(pk["enc"], obsv_p, want_x_last=True)
        pooled = None
        if self.use_social:                                                    # train.py:408-413
            scenes = self.scene_index(sub_batches, n, obsv_p.device)
            ub = ops.rows_linear(enc["h"], pk["pool_m"], pk["pool_m0"])        # (u | beta) = h . M + m0, own kernel
            if precision == "fp16x2" and scenes.max_scene <= ops.pool_tcx_max_scene():
                pooled = ops.pool_tcx(pk["pool"], pk["pool_tcx"], enc["x_last"], enc["h"], ub, scenes)   # layer 2 on tcgen05
            else:
                pooled = ops.pool(pk["pool"], enc["x_last"], enc["h"], ub, scenes)
        if precision == "bf16":
            return ops.decode_tc(pk["tc_w16"], pk["tc_f32"], enc["h"], enc["c"], pooled, noise, enc["x_last"], n_next, out=out)
        if single:
            return ops.decode_tcx(*pk["tcx"], enc["h"], enc["c"], pooled, noise, enc["x_last"], n_next, out=out,
                                  status=self._status_word(obsv_p.device))
        if pair_bf16:
            return ops.decode_pair(*pk["pair_bf16"], enc["h"], enc["c"], pooled, noise, enc["x_last"], n_next, out=out, bf16=True)
        if precision == "fp16x2":
            return ops.decode_pair(*pk["pair"], enc["h"], enc["c"], pooled, noise, enc["x_last"], n_next, out=out,
                                   status=self._status_word(obsv_p.device))
        if precision != "fp32":
            raise ValueError("precision must be 'fp32', 'fp16x2', 'fp16x2s', 'bf16' or 'bf16p'")
        return ops.decode(pk["enc"], pk["dec"], enc["h"], enc["c"], pooled, noise, enc["x_last"], n_next, out=out)

    def predict(self, obsv_p, noise, n_next, sub_batches=()):
        """predict(obsv_p, noise, n_next, sub_batches=[]) -> [N, n_next, 4]  (train.py:392-432)."""
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            from .autograd_path import predict_with_grad
            return predict_with_grad(self, obsv_p, noise, n_next, sub_batches)
        return self.predict_k(obsv_p, noise.unsqueeze(0), n_next, sub_batches)[0]

    forward = predict
