"""Host-side operand packing of the tensor-core kernels (socialways_b200/packing.py), checked on the CPU: layouts are
inverted and the kernels' arithmetic is emulated from the PACKED operands (fp16 hi/lo three-product split, x-feedback K
block, ex2 prescale, shared-reciprocal cell) against the oracle's fp32 LSTM cell.  No GPU involved."""
import ctypes

import numpy as np
import torch

from socialways_b200 import packing

LOG2E = 1.4426950408889634


def uncanon(flat, n, k):
    """inverse of packing._canonical_kmajor: [K/8][N][8] -> [N, K]"""
    return flat.reshape(k // 8, n, 8).permute(1, 0, 2).reshape(n, k)


def make_packs(seed=0):
    from oracle import socialways_oracle as so
    P = so.init_weights(seed=seed)
    enc = packing.pack_encoder(P["encoder.embed.weight"], P["encoder.embed.bias"], P["encoder.lstm.weight_ih_l0"],
                               P["encoder.lstm.weight_hh_l0"], P["encoder.lstm.bias_ih_l0"], P["encoder.lstm.bias_hh_l0"])
    f = lambda i, t: P[f"decoder.fc1.{i}.{t}"]
    dec = packing.pack_decoder(f(0, "weight"), f(0, "bias"), f(2, "weight"), f(2, "bias"), f(4, "weight"), f(4, "bias"),
                               f(5, "weight"), f(5, "bias"))
    return P, enc, dec


def test_tcx_pack_sizes_match_the_kernel():
    from socialways_b200 import _lib
    _, enc, dec = make_packs()
    w16, wsz16, f32 = packing.pack_decoder_tcx(enc, dec)
    a, b, c = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    assert _lib.lib().sw_decode_tcx_pack_sizes(ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)) == 0
    assert (w16.numel(), wsz16.numel(), f32.numel()) == (a.value, b.value, c.value)
    assert w16.dtype == torch.float16 and wsz16.dtype == torch.float16 and f32.dtype == torch.float32
    assert packing.pack_pool_tcx(torch.randn(64, 32)).shape == (4096,)


def test_tcx_pack_layout_and_split_precision():
    _, enc, dec = make_packs(1)
    w16, wsz16, f32 = packing.pack_decoder_tcx(enc, dec)
    w16 = w16.float()
    w1 = dec[:25600].view(160, 160).t()
    w2 = dec[25760:38560].view(160, 80).t()
    scale = torch.tensor([-LOG2E, -LOG2E, -2 * LOG2E, -LOG2E]).repeat(64)
    # W1[h rows]: canonical hi block then lo block
    w1h = uncanon(w16[0:10240], 160, 64) + uncanon(w16[10240:20480], 160, 64)
    assert (w1h - w1[:, :64]).abs().max() <= 2.0 ** -21 * w1[:, :64].abs().max()
    # W2: one block of 160 rows = hi rows then lo rows
    cat = uncanon(w16[20480:46080], 160, 160)
    assert (cat[:80] + cat[80:] - w2).abs().max() <= 2.0 ** -21 * w2.abs().max()
    assert torch.equal(cat[:80], w2.half().float())
    # Whh with the ex2 prescale on the gate-interleaved rows n' = 4 unit + gate
    whh = uncanon(w16[46080:62464], 256, 64) + uncanon(w16[62464:78848], 256, 64)
    want = enc[4:68].t() * scale[:, None]
    assert (whh - want).abs().max() <= 2.0 ** -20 * want.abs().max()
    # x-feedback K block [2][256][8]: Wx_hi | Wx_hi | Wx_lo | b_hi | b_lo | 0 0
    xk = uncanon(w16[78848:82944], 256, 16)
    wx, bl = enc[0:4].t() * scale[:, None], enc[68] * scale
    assert torch.equal(xk[:, 0:4], xk[:, 4:8]) and torch.equal(xk[:, 0:4], wx.half().float())
    assert (xk[:, 0:4] + xk[:, 8:12] - wx).abs().max() <= 2.0 ** -20 * wx.abs().max()
    assert (xk[:, 12] + xk[:, 13] - bl).abs().max() <= 2.0 ** -20 * bl.abs().max()
    assert xk[:, 14:].abs().max() == 0
    # hoisted rows of W1 ([S ; z], K = 96) in three K = 32 chunks, hi | lo each
    wsz = wsz16.float().view(3, 2, -1)
    for ch in range(3):
        rec = uncanon(wsz[ch, 0], 160, 32) + uncanon(wsz[ch, 1], 160, 32)
        want = w1[:, 64 + 32 * ch:96 + 32 * ch]
        assert (rec - want).abs().max() <= 2.0 ** -21 * want.abs().max()
    # fp32 section: b1 | b2 | b34 | pad | W34 [80][2]
    assert torch.equal(f32[0:160], dec[25600:25760]) and torch.equal(f32[160:240], dec[38560:38640])
    assert torch.equal(f32[240:242], dec[38800:38802]) and torch.equal(f32[256:416], dec[38640:38800])


def split(x):
    hi = x.half().float()
    return hi, (x - hi).half().float()


def test_emulated_gate_mma_and_cell_match_the_fp32_lstm_cell():
    """One encoder step on the decode path (train.py:430), computed the way decode_fwd_tcx_kernel does it from the packed
    operands, vs the oracle's explicit fp32 LSTM cell on the reference parameters."""
    from oracle import socialways_oracle as so
    P, enc, dec = make_packs(2)
    w16 = packing.pack_decoder_tcx(enc, dec)[0].float()
    whh_hi, whh_lo = uncanon(w16[46080:62464], 256, 64), uncanon(w16[62464:78848], 256, 64)
    xk = uncanon(w16[78848:82944], 256, 16)
    g = torch.Generator().manual_seed(0)
    n = 64
    h, c = torch.randn(n, 64, generator=g) * 0.5, torch.randn(n, 64, generator=g) * 0.7
    x4 = torch.cat([torch.rand(n, 2, generator=g), torch.randn(n, 2, generator=g) * 0.05], 1)
    # the MMA: three products for h, one K block [x_hi | x_lo | x_hi | 1 | 1 | 0 0] for the feedback and the bias
    h_hi, h_lo = split(h)
    x_hi, x_lo = split(x4)
    a_blk = torch.cat([x_hi, x_lo, x_hi, torch.ones(n, 2), torch.zeros(n, 2)], 1)
    e = h_hi @ whh_hi.t() + h_hi @ whh_lo.t() + h_lo @ whh_hi.t() + a_blk @ xk.t()          # [n, 256], ex2 arguments
    e = e.view(n, 64, 4)
    # lstm_cell_pair_prescaled (csrc/sw_umma.cuh)
    ai, af, ag, ao = (1 + torch.exp2(torch.clamp(e[..., q], max=30.0)) for q in range(4))
    r = 1.0 / (ai * ag * af * ao)
    r_ig, r_fo = r * (af * ao), r * (ai * ag)
    c_new = (ao * r_fo) * c + (ag * r_ig) * (2 * ai * r_ig - 1)
    a_c = 1 + torch.exp2(torch.clamp(-2 * LOG2E * c_new, max=60.0))
    h_new = (af * r_fo) * (2 / a_c - 1)
    # reference arithmetic: embed (Linear 4 -> 64) + LSTM cell with the un-folded parameters
    emb = x4 @ P["encoder.embed.weight"].t() + P["encoder.embed.bias"]
    h_ref, c_ref = so.lstm_cell(emb, h, c, P["encoder.lstm.weight_ih_l0"], P["encoder.lstm.weight_hh_l0"],
                                P["encoder.lstm.bias_ih_l0"], P["encoder.lstm.bias_hh_l0"])
    assert (h_new - h_ref).abs().max() < 2e-6 and (c_new - c_ref).abs().max() < 2e-6


def test_pool_tcx_pack_reconstructs_layer_two():
    w = torch.randn(64, 32, generator=torch.Generator().manual_seed(3)) * 0.3
    p = packing.pack_pool_tcx(w).float()
    hi, lo = uncanon(p[:2048], 64, 32), uncanon(p[2048:], 64, 32)
    assert torch.equal(hi, w.half().float()) and (hi + lo - w).abs().max() <= 2.0 ** -21 * w.abs().max()
    a1 = torch.relu(torch.randn(128, 32, generator=torch.Generator().manual_seed(4)))
    a_hi, a_lo = split(a1)
    got = a_hi @ hi.t() + a_hi @ lo.t() + a_lo @ hi.t()
    assert (got - a1 @ w.t()).abs().max() < 5e-6


def test_pair_pack_bf16_has_the_fp16_layout_with_bf16_bit_patterns():
    """pack_decoder_pair(bf16=True) (operands of sw_decode_fwd_pair_bf16): same sizes and element order as the fp16 hi/lo image;
    every 16-bit lane read as bf16 is the bf16 split of the same weight -- hi parts agree with the fp16 hi parts to bf16 rounding,
    hi + lo reproduces the weight to 2^-16 relative."""
    _, enc, dec = make_packs(5)
    w16, f32 = packing.pack_decoder_pair(enc, dec)
    b16, bf32 = packing.pack_decoder_pair(enc, dec, bf16=True)
    assert b16.shape == w16.shape and b16.dtype == torch.float16 and torch.equal(f32, bf32)
    as_bf = b16.view(torch.bfloat16).float()
    as_f = w16.float()
    # W1[h rows] of rank 0: hi block [0, 5120), lo block [5120, 10240)
    hi_f, lo_f = as_f[0:5120], as_f[5120:10240]
    hi_b, lo_b = as_bf[0:5120], as_bf[5120:10240]
    w = hi_f + lo_f                                           # the fp32 weight to 2^-22
    assert (hi_b - w).abs().max() <= (w.abs() * 2.0 ** -8).max()
    assert ((hi_b + lo_b) - w).abs().max() <= (w.abs().max() * 2.0 ** -15)
    # the ones of the x-feedback block live in the A operand (kernel side); its bias columns are bf16 splits of the same bias
    assert torch.isfinite(as_bf).all()


def test_pair_pack_halves_reassemble_the_one_tile_operands():
    """pack_decoder_pair cuts every B matrix [N][K] into the two N halves the two CTAs of a pair supply; the halves of both ranks,
    put back together, must be the matrices of pack_decoder_tcx (same split, same prescale), and the sizes must match the
    kernel's (sw_decode_pair_pack_sizes)."""
    from socialways_b200 import _lib
    _, enc, dec = make_packs(3)
    w16, f32 = packing.pack_decoder_pair(enc, dec)
    a, b = ctypes.c_int(), ctypes.c_int()
    assert _lib.lib().sw_decode_pair_pack_sizes(ctypes.byref(a), ctypes.byref(b)) == 0
    assert (w16.numel(), f32.numel()) == (a.value, b.value) and w16.dtype == torch.float16
    per_rank = w16.numel() // 2
    one = packing.pack_decoder_tcx(enc, dec)
    t16, tsz, tf32 = one[0].float(), one[1].float(), one[2]
    assert torch.equal(f32, tf32)
    ranks = [w16[i * per_rank:(i + 1) * per_rank].float() for i in range(2)]

    def both(off, rows, k):           # [2 ranks][K/8][rows][8] at `off` -> [2 * rows, K]
        return torch.cat([uncanon(r[off:off + rows * k], rows, k) for r in ranks])

    # W1[h rows] hi | lo: [8][80][8] per rank
    assert torch.equal(both(0, 80, 64), uncanon(t16[0:10240], 160, 64))
    assert torch.equal(both(5120, 80, 64), uncanon(t16[10240:20480], 160, 64))
    # W2 hi | lo (not stacked in the pair pack): [20][40][8] per rank vs the stacked [20][hi 80 | lo 80][8]
    cat = uncanon(t16[20480:46080], 160, 160)
    assert torch.equal(both(10240, 40, 160), cat[:80]) and torch.equal(both(10240 + 6400, 40, 160), cat[80:])
    # Whh per gate half g (rows [128 g, 128 g + 128), 64 per rank), hi | lo; x-feedback K block per gate half
    whh_hi, whh_lo = uncanon(t16[46080:62464], 256, 64), uncanon(t16[62464:78848], 256, 64)
    xk = uncanon(t16[78848:82944], 256, 16)
    base = 10240 + 2 * 6400
    for g in range(2):
        o = base + g * 8192
        assert torch.equal(both(o, 64, 64), whh_hi[128 * g:128 * g + 128])
        assert torch.equal(both(o + 4096, 64, 64), whh_lo[128 * g:128 * g + 128])
        assert torch.equal(both(base + 16384 + g * 1024, 64, 16), xk[128 * g:128 * g + 128])
    # hoisted rows of W1 (S, z): hi | lo [12][80][8] per rank vs three K = 32 chunks of [hi 160 | lo 160] rows
    wsz_hi = torch.cat([uncanon(tsz[ch * 10240:ch * 10240 + 5120], 160, 32) for ch in range(3)], 1)
    wsz_lo = torch.cat([uncanon(tsz[ch * 10240 + 5120:(ch + 1) * 10240], 160, 32) for ch in range(3)], 1)
    o = base + 16384 + 2048
    assert torch.equal(both(o, 80, 96), wsz_hi) and torch.equal(both(o + 7680, 80, 96), wsz_lo)


def test_emulated_tensor_core_encoder_matches_the_fp32_lstm():
    """The observation encoder the way lstm_seq_fwd_tcx_kernel computes it from pack_encoder_tcx -- zero state, per step ONE K
    block [x_hi | x_lo | x_hi | 1 | 1 | 0 0] for Wx . x4 + b, three products for h . Whh^T from step 1 on, pre-scaled gates, the
    shared-reciprocal cell -- over 8 steps, vs the oracle's fp32 embed + LSTM cell on the reference parameters."""
    from oracle import socialways_oracle as so
    P, enc, _ = make_packs(4)
    (w16,) = packing.pack_encoder_tcx(enc)
    assert w16.numel() == 2 * 64 * 256 + 256 * 16 and w16.dtype == torch.float16
    w16 = w16.float()
    whh_hi, whh_lo, xk = uncanon(w16[0:16384], 256, 64), uncanon(w16[16384:32768], 256, 64), uncanon(w16[32768:], 256, 16)
    g = torch.Generator().manual_seed(1)
    n, T = 48, 8
    pos = torch.cumsum(torch.randn(n, T, 2, generator=g) * 0.05, 1) + torch.rand(n, 1, 2, generator=g)
    vel = torch.cat([pos[:, 1:2] - pos[:, 0:1], pos[:, 1:] - pos[:, :-1]], 1)            # v_0 := v_1 (train.py:131-133)
    x4 = torch.cat([pos, vel], 2)
    h = c = torch.zeros(n, 64)
    h_ref, c_ref = torch.zeros(n, 64), torch.zeros(n, 64)
    for t in range(T):
        x_hi, x_lo = split(x4[:, t])
        a_blk = torch.cat([x_hi, x_lo, x_hi, torch.ones(n, 2), torch.zeros(n, 2)], 1)
        e = a_blk @ xk.t()
        if t > 0:
            h_hi, h_lo = split(h)
            e = e + h_hi @ whh_hi.t() + h_hi @ whh_lo.t() + h_lo @ whh_hi.t()
        e = e.view(n, 64, 4)
        ai, af, ag, ao = (1 + torch.exp2(torch.clamp(e[..., q], max=30.0)) for q in range(4))
        r = 1.0 / (ai * ag * af * ao)
        r_ig, r_fo = r * (af * ao), r * (ai * ag)
        c = (ao * r_fo) * c + (ag * r_ig) * (2 * ai * r_ig - 1)
        h = (af * r_fo) * (2 / (1 + torch.exp2(torch.clamp(-2 * LOG2E * c, max=60.0))) - 1)
        emb = x4[:, t] @ P["encoder.embed.weight"].t() + P["encoder.embed.bias"]
        h_ref, c_ref = so.lstm_cell(emb, h_ref, c_ref, P["encoder.lstm.weight_ih_l0"], P["encoder.lstm.weight_hh_l0"],
                                    P["encoder.lstm.bias_ih_l0"], P["encoder.lstm.bias_hh_l0"])
    assert (h - h_ref).abs().max() < 3e-6 and (c - c_ref).abs().max() < 3e-6


def test_newton_reciprocal_of_the_pair_kernel_cell():
    """rcp_newton (csrc/sw_umma.cuh): magic-constant seed + three Newton steps, over the range the cell update feeds it
    (products of four (1 + 2^e) terms, e <= 30): relative error below 1e-7."""
    x = torch.exp2(torch.linspace(0.0, 120.0, 200001)).float()
    seed = (torch.tensor(0x7EF311C7, dtype=torch.int32) - x.view(torch.int32)).view(torch.float32)
    y = seed
    for _ in range(3):
        y = y + y * (1.0 - x * y)
    rel = ((y.double() * x.double()) - 1.0).abs().max().item()
    assert rel < 1.2e-7, rel


def test_packed_epilogue_identities_hold_bit_for_bit_in_fp32():
    """The packed fp32x2 epilogues of the decode kernels rely on two rewrites being exact in IEEE fp32 (sw_umma.cuh / decode_pair.cuh):
    lrelu as max(y, 0.2 y) instead of the select form, and -2 log2e . c instead of -log2e . (2 c)."""
    import numpy as np
    rng = np.random.default_rng(0)
    y = np.concatenate([rng.standard_normal(200000).astype(np.float32) * np.float32(10.0) ** rng.integers(-30, 30, 200000).astype(np.float32),
                        np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 1e-45, -1e-45, 3e38, -3e38], np.float32)])
    with np.errstate(all="ignore"):
        select = np.where(y > 0, y, np.float32(0.2) * y).astype(np.float32)
        viamax = np.fmax(y, np.float32(0.2) * y).astype(np.float32)      # fmaxf: the non-NaN operand wins, NaN only if both are
    same = (select.view(np.uint32) == viamax.view(np.uint32)) | (np.isnan(select) & np.isnan(viamax))
    assert same.all()
    c = y[np.isfinite(y) & (np.abs(y) < 1e37) & (np.abs(y) > 1e-30)]
    a = np.float32(-1.4426950408889634) * (np.float32(2.0) * c)
    b = np.float32(-2.8853900817779268) * c
    assert np.float32(-2.8853900817779268) == np.float32(2.0) * np.float32(-1.4426950408889634)
    assert (a.view(np.uint32) == b.view(np.uint32)).all()
