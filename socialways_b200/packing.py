"""Weight folding / packing for the sm_100a kernels (differentiable torch ops, tiny).

All folds are exact algebra on the reference parameters (fp32 re-association only):

* ``pack_encoder``: EncoderLstm.embed (Linear 4->64, train.py:251) is folded into the LSTM input
  projection:  gates = W_ih (W_e x + b_e) + b_ih + W_hh h + b_hh = (W_ih W_e) x + W_hh h + b.
* ``pack_decoder``: DecoderFC.fc1 (train.py:324-328): layer-1 weight stored k-major with the
  rows ordered {h, S, z}; the last two Linear layers (no activation between) become one 80->2.
* ``pack_pool``: EmbedSocialFeatures.fc.{0,2} as-is; fc.4 and AttentionPooling.W folded into the
  per-agent vectors  u_j = W3^T (W h_j + b_W),  beta_j = b3 . (W h_j + b_W)  (``pool_agent_terms``), so
  that sigma_ij = a2_ij . u_j + beta_j  (train.py:169 with emb = fc(features)).

lstm_pack layout [69][256]: rows 0..3 Wx (k-major), 4..67 Whh (k-major), 68 bias; column
n' = 4*unit + gate with torch's gate order (i, f, g, o).
"""
import torch


def _interleave(m):
    """[256, K] with rows gate*64+unit  ->  [K, 256] k-major with columns unit*4+gate."""
    k = m.shape[1]
    return m.view(4, 64, k).permute(2, 1, 0).reshape(k, 256)


def pack_lstm(w_x, w_hh, bias):
    return torch.cat([_interleave(w_x), _interleave(w_hh), _interleave(bias.view(256, 1))], dim=0).contiguous()


def pack_encoder(embed_w, embed_b, w_ih, w_hh, b_ih, b_hh):
    return pack_lstm(w_ih @ embed_w, w_hh, w_ih @ embed_b + b_ih + b_hh)


def pack_disc_lstm(w_ih, w_hh, b_ih, b_hh):
    return pack_lstm(w_ih, w_hh, b_ih + b_hh)


def pack_decoder(w1, b1, w2, b2, w3, b3, w4, b4):
    w34 = w4 @ w3                       # [2, 80]
    b34 = w4 @ b3 + b4
    return torch.cat([w1.t().reshape(-1), b1, w2.t().reshape(-1), b2, w34.t().reshape(-1), b34]).contiguous()


def pack_pool(fc0_w, fc0_b, fc2_w, fc2_b):
    return torch.cat([torch.cat([fc0_w, fc0_b.unsqueeze(1)], dim=1).reshape(-1), fc2_w.reshape(-1), fc2_b]).contiguous()


def pool_agent_matrix(att_w, att_b, fc4_w, fc4_b):
    """[64, 65] matrix M and [65] offset m0 with  (u | beta) = h @ M + m0."""
    tail = torch.cat([fc4_w, fc4_b.unsqueeze(1)], dim=1)    # [64(n), 65]: Wh @ tail = (u | beta)
    return att_w.t() @ tail, att_b @ tail


def _canonical_kmajor(w_nk):
    """[N, K] (K % 8 == 0) -> canonical no-swizzle K-major UMMA operand: [K/8][N][8] (8-row x 16-byte core
    matrices, SBO = 128 B, LBO = N * 16 B)."""
    n, k = w_nk.shape
    return w_nk.reshape(n, k // 8, 8).permute(1, 0, 2).contiguous().reshape(-1)


def pack_decoder_tc(lstm_pack, dec_pack):
    """Operands of the tcgen05 decode kernel (csrc/decode_fwd_tc.cu) from the fp32 packs:
    bf16 section  W1 [160 n][160 k] | W2 [80][160] | W34 [16 (2 real)][80] | Whh [256 n'][64], each canonical;
    fp32 section  Wx [4][256] | bL [256] | b1 [160] | b2 [80] | b34 [2] | pad -> 1536 floats."""
    w1 = dec_pack[:25600].view(160, 160).t()
    b1 = dec_pack[25600:25760]
    w2 = dec_pack[25760:38560].view(160, 80).t()
    b2 = dec_pack[38560:38640]
    w34 = dec_pack[38640:38800].view(80, 2).t()
    b34 = dec_pack[38800:38802]
    w34p = torch.zeros(16, 80, device=dec_pack.device, dtype=dec_pack.dtype)
    w34p[:2] = w34
    whh = lstm_pack[4:68].t()
    w16 = torch.cat([_canonical_kmajor(w1), _canonical_kmajor(w2), _canonical_kmajor(w34p), _canonical_kmajor(whh)])
    f32 = torch.cat([lstm_pack[0:4].reshape(-1), lstm_pack[68], b1, b2, b34, b34.new_zeros(14)])
    return w16.to(torch.bfloat16).contiguous(), f32.contiguous()


def _split_f16(w):
    hi = w.to(torch.float16)
    lo = (w - hi.to(torch.float32)).to(torch.float16)
    return hi, lo


def _split_bf16_bits(w):
    """hi / lo split in bf16, returned as the BIT PATTERNS in fp16-typed tensors (the packs are typed fp16; the bf16 kernels read
    the same 16-bit lanes with the bf16 operand format)."""
    hi = w.to(torch.bfloat16)
    lo = (w - hi.to(torch.float32)).to(torch.bfloat16)
    return hi.view(torch.float16), lo.view(torch.float16)


LOG2E = 1.4426950408889634


def _gate_prescale(device, dtype):
    """[256] scale of the gate-interleaved rows n' = 4*unit + gate: -log2(e) for i, f, o and -2 log2(e) for g, so that the
    gate accumulator is directly the ex2 argument of  sigma(x) = 1/(1 + 2^(-log2e x)),  tanh(x) = 2/(1 + 2^(-2 log2e x)) - 1."""
    s = torch.full((64, 4), -LOG2E, device=device, dtype=dtype)      # device-side fills only: legal inside a CUDA-graph capture
    s[:, 2] = -2.0 * LOG2E
    return s.reshape(256)


def pack_decoder_tcx(lstm_pack, dec_pack):
    """Operands of the fp16-split tcgen05 decode kernel (csrc/decode_fwd_tcx.cu):
    w16   fp16 [82944]: W1h [160 n][64 k] and Whh [256 n'][64] as canonical hi block then canonical lo block (x = hi + lo);
          W2 [80][160]: ONE canonical block of 160 rows = hi rows then lo rows; then the x-feedback K block of the gate MMA
          [2 chunks][256 n'][8]: k 0..3 Wx_hi, 4..7 Wx_hi, 8..11 Wx_lo, 12 b_hi, 13 b_lo, 14..15 zero (the kernel's A row is
          [x_hi | x_lo | x_hi | 1 | 1 | 0 0]).  The gate rows (Whh, Wx, b) carry the ex2 prescale (`_gate_prescale`).
    wsz16 fp16 [3][2][4][160][8]: the hoisted rows of W1 (S: k 0..63, z: 64..95) in three K = 32 chunks, hi | lo
    f32   [416]: b1 [160] | b2 [80] | b34 [2] | pad | W34 [80 k][2] (fp32: the folded 80 -> 2 output layer runs as FMAs
          inside the layer-2 epilogue)"""
    w1 = dec_pack[:25600].view(160, 160).t()                  # [n, k], k order {h, S, z}
    b1 = dec_pack[25600:25760]
    w2 = dec_pack[25760:38560].view(160, 80).t()
    b2 = dec_pack[38560:38640]
    b34 = dec_pack[38800:38802]
    scale = _gate_prescale(lstm_pack.device, lstm_pack.dtype)
    whh = lstm_pack[4:68].t() * scale[:, None]                # [256 n', 64]
    wx = lstm_pack[0:4].t() * scale[:, None]                  # [256 n', 4]
    bl = lstm_pack[68] * scale                                # [256]
    parts = []
    for name, m in (("w1h", w1[:, :64]), ("w2", w2), ("whh", whh)):
        hi, lo = _split_f16(m.contiguous())
        if name == "w2":        # hi and lo rows stacked along N: one N = 160 MMA pass covers a1.W2_hi and a1.W2_lo
            parts.append(_canonical_kmajor(torch.cat([hi, lo], dim=0)))
        else:
            parts += [_canonical_kmajor(hi), _canonical_kmajor(lo)]
    wx_hi, wx_lo = _split_f16(wx.contiguous())
    bl_hi, bl_lo = _split_f16(bl.contiguous())
    zero = torch.zeros(256, 2, device=wx.device, dtype=torch.float16)
    parts.append(_canonical_kmajor(torch.cat([wx_hi, wx_hi, wx_lo, bl_hi[:, None], bl_lo[:, None], zero], dim=1)))
    w16 = torch.cat(parts).contiguous()
    chunks = []
    for ch in range(3):
        hi, lo = _split_f16(w1[:, 64 + 32 * ch:64 + 32 * (ch + 1)].contiguous())
        chunks += [_canonical_kmajor(hi), _canonical_kmajor(lo)]
    wsz16 = torch.cat(chunks).contiguous()
    f32 = torch.cat([b1, b2, b34, b34.new_zeros(14), dec_pack[38640:38800]]).contiguous()
    return w16, wsz16, f32


def pack_decoder_pair(lstm_pack, dec_pack, bf16=False):
    """Operands of the CTA-pair decode kernel (csrc/decode_fwd_pair.cu, tcgen05 cta_group::2): the same hi/lo split matrices
    as `pack_decoder_tcx`, but every B matrix [N][K] is cut into the two N halves the two CTAs of a pair supply (rank 0: rows
    [0, N/2), rank 1: [N/2, N)), each canonical K-major.
    w16  fp16 [2 ranks][56832]: W1h hi | lo [8][80][8]; W2 hi | lo [20][40][8] (not stacked: the pair kernel accumulates the
         three products into ONE 80-column region); Whh per gate half g (n' in [128 g, 128 g + 128), the rank supplies 64 of
         them) hi | lo [8][64][8]; the x-feedback K block per gate half [2][64][8] (k 0..3 Wx_hi, 4..7 Wx_hi, 8..11 Wx_lo,
         12 b_hi, 13 b_lo, 14..15 zero); the hoisted rows of W1 (S: k 0..63, z: 64..95) hi | lo [12][80][8] -- resident in the
         pair kernel.  Gate rows carry the ex2 prescale.
    f32  [416]: b1 [160] | b2 [80] | b34 [2] | pad | W34 [80 k][2]  (as pack_decoder_tcx)
    bf16=True: the same image with bf16 bit patterns, for sw_decode_fwd_pair_bf16 (which reads the hi parts only, except in the
    x-feedback block)."""
    _split_f16 = _split_bf16_bits if bf16 else globals()["_split_f16"]
    w1 = dec_pack[:25600].view(160, 160).t()                  # [n, k], k order {h, S, z}
    b1 = dec_pack[25600:25760]
    w2 = dec_pack[25760:38560].view(160, 80).t()
    b2 = dec_pack[38560:38640]
    b34 = dec_pack[38800:38802]
    scale = _gate_prescale(lstm_pack.device, lstm_pack.dtype)
    whh = lstm_pack[4:68].t() * scale[:, None]                # [256 n', 64]
    wx = lstm_pack[0:4].t() * scale[:, None]                  # [256 n', 4]
    bl = lstm_pack[68] * scale                                # [256]
    w1h_hi, w1h_lo = _split_f16(w1[:, :64].contiguous())
    w2_hi, w2_lo = _split_f16(w2.contiguous())
    whh_hi, whh_lo = _split_f16(whh.contiguous())
    wsz_hi, wsz_lo = _split_f16(w1[:, 64:].contiguous())
    wx_hi, wx_lo = _split_f16(wx.contiguous())
    bl_hi, bl_lo = _split_f16(bl.contiguous())
    zero = torch.zeros(256, 2, device=wx.device, dtype=torch.float16)
    xkb = torch.cat([wx_hi, wx_hi, wx_lo, bl_hi[:, None], bl_lo[:, None], zero], dim=1)      # [256, 16]
    images = []
    for rank in (0, 1):
        parts = [_canonical_kmajor(m[80 * rank:80 * rank + 80]) for m in (w1h_hi, w1h_lo)]
        parts += [_canonical_kmajor(m[40 * rank:40 * rank + 40]) for m in (w2_hi, w2_lo)]
        for g in (0, 1):
            r0 = 128 * g + 64 * rank
            parts += [_canonical_kmajor(whh_hi[r0:r0 + 64]), _canonical_kmajor(whh_lo[r0:r0 + 64])]
        for g in (0, 1):
            r0 = 128 * g + 64 * rank
            parts.append(_canonical_kmajor(xkb[r0:r0 + 64]))
        parts += [_canonical_kmajor(m[80 * rank:80 * rank + 80]) for m in (wsz_hi, wsz_lo)]
        images.append(torch.cat(parts))
    w16 = torch.cat(images).contiguous()
    f32 = torch.cat([b1, b2, b34, b34.new_zeros(14), dec_pack[38640:38800]]).contiguous()
    return w16, f32


def pack_pool_tcx(fc2_w):
    """Layer-2 operand of the tensor-core pooling kernel (csrc/pool_fwd_tcx.cu): EmbedSocialFeatures.fc.2.weight [64 n][32 k]
    as canonical hi block then canonical lo block ([4][64][8] each), fp16 [4096]."""
    hi, lo = _split_f16(fc2_w.contiguous())
    return torch.cat([_canonical_kmajor(hi), _canonical_kmajor(lo)]).contiguous()


def pack_encoder_tcx(lstm_pack):
    """Operand of the tensor-core observation encoder (csrc/lstm_seq_fwd_tcx.cu), a 1-tuple (w16,): Whh [256 n'][64 k] as
    canonical hi | lo fp16 blocks (x = hi + lo), then the x-feedback K block [2][256][8] with rows
    [Wx_hi(4) | Wx_hi(4) | Wx_lo(4) | b_hi | b_lo | 0 0] -- the three products of the hi/lo split of  Wx . x4 + b  in ONE K block
    against the A row [x_hi | x_lo | x_hi | 1 | 1 | 0 0].  Gate rows carry the ex2 prescale (-log2 e for i, f, o; -2 log2 e for g)."""
    scale = _gate_prescale(lstm_pack.device, lstm_pack.dtype)
    hi, lo = _split_f16((lstm_pack[4:68].t() * scale[:, None]).contiguous())
    wx_hi, wx_lo = _split_f16((lstm_pack[0:4].t() * scale[:, None]).contiguous())
    bl_hi, bl_lo = _split_f16((lstm_pack[68] * scale).contiguous())
    zero = torch.zeros(256, 2, device=lstm_pack.device, dtype=torch.float16)
    xkb = torch.cat([wx_hi, wx_hi, wx_lo, bl_hi[:, None], bl_lo[:, None], zero], dim=1)      # [256, 16]
    w16 = torch.cat([_canonical_kmajor(hi), _canonical_kmajor(lo), _canonical_kmajor(xkb)]).contiguous()
    return (w16,)
