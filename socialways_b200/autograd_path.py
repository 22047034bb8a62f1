"""Training path: torch.autograd.Function wrappers around the forward+backward kernels.

The kernels compute every quantity that has a sequential or pairwise dependency (the LSTM / decode
recurrences, their reverse-time data gradients, the per-pair attention gradients).  What is left on
the host are plain dense contractions without any dependency -- the weight gradients
`dW = sum_rows X^T dY` over stash images the kernels wrote, and the tiny weight folds of
packing.py -- which are library GEMMs (cuBLAS through torch.matmul, fp32, TF32 off).

Tile-image layout used by the stashes: [..., tiles, k, 32] = the kernels' shared-memory operand
(32 rows contiguous), tiles = ceil(rows / 32).
"""
import torch

from . import ops


def _contract(a_img, b_img):
    """sum over (leading dims, rows) of a[..., k, r] * b[..., n, r]  ->  [k, n]."""
    k, n = a_img.shape[-2], b_img.shape[-2]
    a2 = a_img.movedim(-2, 0).reshape(k, -1)
    b2 = b_img.movedim(-2, 0).reshape(n, -1)
    return a2 @ b2.t()


def _image_to_rows(img, n_rows):
    """[tiles, k, 32] -> [n_rows, k]."""
    return img.transpose(-1, -2).reshape(-1, img.shape[-2])[:n_rows]


def _lstm_pack_grad(stash_xh, g_gates):
    return torch.cat([_contract(stash_xh, g_gates), g_gates.sum(dim=(0, 1, 3)).unsqueeze(0)], dim=0)


class LstmSeqFn(torch.autograd.Function):
    """(lstm_pack [69,256], x [N,T,2|4]) -> (h_T, c_T, x_last) from a zero initial state."""

    @staticmethod
    def forward(ctx, pack, x):
        r = ops.lstm_seq(pack, x, want_x_last=True, stash=True)
        ctx.save_for_backward(pack, r["stash_gates"], r["stash_xh"])
        ctx.n_rows = x.shape[0]
        ctx.mark_non_differentiable(r["x_last"])
        return r["h"], r["c"], r["x_last"]

    @staticmethod
    def backward(ctx, dh, dc, _dx_last):
        pack, stash_gates, stash_xh = ctx.saved_tensors
        pack_t = pack[:68].t().contiguous()
        g_gates = ops.lstm_seq_bwd(pack_t, stash_gates, dh, dc, ctx.n_rows)
        return _lstm_pack_grad(stash_xh, g_gates), None


class PoolFn(torch.autograd.Function):
    """(pool_pack, x_last, h, ub [N,65]) -> pooled [N,64]."""

    @staticmethod
    def forward(ctx, pool_pack, x_last, h, ub, scenes):
        pooled, attn = ops.pool(pool_pack, x_last, h, ub, scenes, want_attn=True)
        ctx.save_for_backward(pool_pack, x_last, h, ub, attn, pooled)
        ctx.scenes = scenes
        return pooled

    @staticmethod
    def backward(ctx, d_pooled):
        pool_pack, x_last, h, ub, attn, pooled = ctx.saved_tensors
        d_pooled = d_pooled.contiguous()
        tdot = (d_pooled * pooled).sum(dim=1)
        dub, dh, st_a1, st_g2, st_g1, st_f = ops.pool_bwd(pool_pack, x_last, h, ub, d_pooled, tdot, attn, ctx.scenes)
        d_pack = torch.cat([(st_g1.t() @ st_f).reshape(-1), (st_g2.t() @ st_a1).reshape(-1), st_g2.sum(dim=0)])
        return d_pack, None, dh, dub, None


class DecodeFn(torch.autograd.Function):
    """(enc_pack, dec_pack, h0, c0, pooled|None, noise [N,32], x_last) -> [N, n_next, 4]."""

    @staticmethod
    def forward(ctx, enc_pack, dec_pack, h0, c0, pooled, noise, x_last, n_next):
        out, st = ops.decode(enc_pack, dec_pack, h0, c0, pooled, noise.unsqueeze(0), x_last, n_next, stash=True)
        ctx.save_for_backward(enc_pack, dec_pack, c0, noise, st["xh"], st["gates"], st["a1"], st["a2"],
                              pooled if pooled is not None else noise.new_zeros(0))
        ctx.has_pooled = pooled is not None
        ctx.n_next = n_next
        return out[0]

    @staticmethod
    def backward(ctx, d_out):
        enc_pack, dec_pack, c0, noise, s_xh, s_gates, s_a1, s_a2, pooled = ctx.saved_tensors
        n, t = noise.shape[0], ctx.n_next
        w1 = dec_pack[:160 * 160].view(160, 160)                    # k-major: [in, out], rows {h, S, z}
        w2 = dec_pack[160 * 160 + 160:160 * 160 + 160 + 160 * 80].view(160, 80)
        w34 = dec_pack[-162:-2]
        dec_pack_t = torch.cat([w1[:64].t().reshape(-1), w2.t().reshape(-1), w34])
        pack_t = enc_pack[:68].t().contiguous()
        g = ops.decode_bwd(pack_t, dec_pack_t, c0, dict(gates=s_gates, a1=s_a1, a2=s_a2), d_out.contiguous(), n, 1)
        if t > 1:
            d_enc = _lstm_pack_grad(s_xh[:t - 1], g["gates"])
        else:
            d_enc = torch.zeros_like(enc_pack)
        sum_a1 = _image_to_rows(g["a1"].sum(dim=0), n)             # [N, 160] = sum_t da1pre
        d_w1h = _contract(s_xh[:, :, 4:68], g["a1"])               # [64, 160]
        d_w1s = pooled.t() @ sum_a1 if ctx.has_pooled else sum_a1.new_zeros(64, 160)
        d_w1z = noise.t() @ sum_a1
        d_dec = torch.cat([torch.cat([d_w1h, d_w1s, d_w1z], dim=0).reshape(-1), sum_a1.sum(dim=0),
                           _contract(s_a1, g["a2"]).reshape(-1), g["a2"].sum(dim=(0, 1, 3)),
                           _contract(s_a2, g["v"]).reshape(-1), g["v"].sum(dim=(0, 1, 3))])
        d_pooled = sum_a1 @ w1[64:128].t() if ctx.has_pooled else None
        return d_enc, d_dec, g["dh0"], g["dc0"], d_pooled, None, None, None


class DiscHeadsFn(torch.autograd.Function):
    """(heads pack, h [N,64], pred [N,P]) -> (label [N,1], code [N,L]): the four FC blocks of the Discriminator
    (train.py:281-292,300-309) in one forward and one backward launch; every parameter gradient is a block of
    ONE GEMM  X^T . G  over the per-row records the kernels write."""

    @staticmethod
    def forward(ctx, pack, h, pred, n_latent):
        label, code, xrec = ops.disc_heads_fwd(pack, h, pred, n_latent, record=True)
        ctx.save_for_backward(pack, xrec)
        ctx.p, ctx.n_latent = pred.shape[1], n_latent
        ctx.need = (h.requires_grad, pred.requires_grad)
        return label, code

    @staticmethod
    def backward(ctx, d_label, d_code):
        pack, xrec = ctx.saved_tensors
        p, L = ctx.p, ctx.n_latent
        d_h, d_pred, g = ops.disc_heads_bwd(pack, xrec, p, d_label.contiguous(), d_code.contiguous(), L,
                                            want_dh=ctx.need[0], want_dpred=ctx.need[1])
        m = xrec.t() @ g                                     # [257 + P, 193 + L]
        one = 256 + p
        blocks = [(0, 64, 0, 32), (64, 96, 32, 64), (96, 96 + p, 64, 96), (96 + p, 128 + p, 96, 128),
                  (128 + p, 192 + p, 128, 160), (192 + p, 224 + p, 160, 161), (128 + p, 192 + p, 161, 193),
                  (224 + p, 256 + p, 193, 193 + L)]
        parts = []
        for r0, r1, c0, c1 in blocks:
            parts.append(m[r0:r1, c0:c1].t().reshape(-1))   # dW [out][in]
            parts.append(m[one, c0:c1])                      # db
        return torch.cat(parts), d_h, d_pred, None


def predict_with_grad(gen, obsv_p, noise, n_next, sub_batches=()):
    """predict() (train.py:392-432) with autograd: same three kernels as inference plus their stashes."""
    from . import packing
    if not obsv_p.is_cuda:
        from ._lib import SocialWaysCudaError
        raise SocialWaysCudaError("predict() runs on CUDA tensors only (no CPU fallback)")
    n = obsv_p.shape[0]
    enc_pack = gen.encoder.packed()
    dec_pack = gen.decoder.packed()
    h, c, x_last = LstmSeqFn.apply(enc_pack, obsv_p)
    pooled = None
    if gen.use_social:
        fe, att = gen.feature_embedder.fc, gen.attention.W
        scenes = gen.scene_index(sub_batches, n, obsv_p.device)
        m, m0 = packing.pool_agent_matrix(att.weight, att.bias, fe[4].weight, fe[4].bias)
        ub = torch.addmm(m0, h, m)
        pool_pack = packing.pack_pool(fe[0].weight, fe[0].bias, fe[2].weight, fe[2].bias)
        pooled = PoolFn.apply(pool_pack, x_last, h, ub, scenes)
    return DecodeFn.apply(enc_pack, dec_pack, h, c, pooled, noise, x_last, n_next)
