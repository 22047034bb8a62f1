"""Thin tensor-level wrappers over the C-ABI (include/socialways_b200.h): allocate outputs with torch,
pass raw device pointers + the current CUDA stream.  No arithmetic happens here."""
import os

import numpy as np
import torch

from . import _lib

H = 64
Z = 32


def _stream():
    return torch.cuda.current_stream().cuda_stream


_SM_COUNT = {}


def sm_count(device):
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx not in _SM_COUNT:
        _SM_COUNT[idx] = torch.cuda.get_device_properties(idx).multi_processor_count
    return _SM_COUNT[idx]


def _f32(t):
    if t.dtype != torch.float32:
        raise _lib.SocialWaysCudaError(f"fp32 tensor expected, got {t.dtype}")
    return t.contiguous()


class SceneIndex:
    """Device-side form of the reference's `sub_batches` (train.py:461): ascending, contiguous
    [start, end) agent ranges.  Agents not covered by any range are treated as 1-agent scenes
    (their pooled vector stays zero, exactly what the reference's zero-initialised S gives them)."""

    def __init__(self, sub_batches, n_agents, device):
        sb = np.asarray(sub_batches, dtype=np.int64).reshape(-1, 2) if len(sub_batches) else np.array([[0, n_agents]])
        offs = [0]
        for a, b in sb:
            a, b = int(a), int(b)
            if a < offs[-1] or b <= a or b > n_agents:
                raise ValueError("sub_batches must be ascending, non-overlapping [start,end) ranges inside the batch")
            offs.extend(range(offs[-1] + 1, a + 1))          # uncovered agents -> singleton scenes
            offs.append(b)
        offs.extend(range(offs[-1] + 1, n_agents + 1))
        offs = np.asarray(offs, dtype=np.int32)
        sizes = np.diff(offs)
        self.n_agents = n_agents
        self.n_scenes = len(sizes)
        self.max_scene = int(sizes.max())
        self.offsets = torch.from_numpy(offs).to(device)
        self.agent_scene = torch.from_numpy(np.repeat(np.arange(self.n_scenes, dtype=np.int32), sizes)).to(device)
        self.sizes = sizes
        pairs = np.concatenate([[0], np.cumsum(sizes.astype(np.int64) ** 2)])
        self.n_pairs = int(pairs[-1])
        self._pairs_np = pairs
        self._pair_offsets = None
        self._pool_units = None
        self.device = device

    def pool_units(self):
        """Work units of sw_pool_fwd_tcx: (device int32 [n_units, 2] = (first row, rows), max span, max ordered pairs).
        Whole scenes are packed into units of up to 64 rows; a scene with more than 64 agents is cut into units of up to
        32 rows (16 beyond 256 agents) of that scene alone, so a unit's span is the unit itself or one scene."""
        if self._pool_units is None:
            big_rows = 32 if self.max_scene <= 256 else 16
            cap = int(os.environ.get("SW_POOL_UNIT_ROWS", "64"))     # A/B knob; must not exceed the kernel's SW_PT_ROWS
            units, span, pairs = [], 1, 1
            start, rows, npairs = 0, 0, 0
            offs = np.concatenate([[0], np.cumsum(self.sizes)])
            for s_i, a in enumerate(self.sizes.tolist()):
                if a > 64:
                    if rows:
                        units.append((start, rows)); span, pairs = max(span, rows), max(pairs, npairs)
                        rows, npairs = 0, 0
                    for r0 in range(0, a, big_rows):
                        r = min(big_rows, a - r0)
                        units.append((int(offs[s_i]) + r0, r)); span, pairs = max(span, a), max(pairs, r * a)
                    start = int(offs[s_i + 1])
                    continue
                if rows + a > cap:
                    units.append((start, rows)); span, pairs = max(span, rows), max(pairs, npairs)
                    start, rows, npairs = int(offs[s_i]), 0, 0
                if rows == 0:
                    start = int(offs[s_i])
                rows += a
                npairs += a * a
            if rows:
                units.append((start, rows)); span, pairs = max(span, rows), max(pairs, npairs)
            self._pool_units = (torch.tensor(units, dtype=torch.int32, device=self.device).contiguous(), span, pairs)
        return self._pool_units

    @property
    def pair_offsets(self):
        """[n_scenes+1] int64 offsets of every scene's A x A block of ordered pairs (backward pass only)."""
        if self._pair_offsets is None:
            self._pair_offsets = torch.from_numpy(self._pairs_np).to(self.device)
        return self._pair_offsets


def n_tiles(n_rows):
    return (n_rows + 31) // 32


def lstm_seq(lstm_pack, x, h_in=None, c_in=None, want_y=False, want_x_last=False, stash=False):
    """sw_lstm_seq_fwd.  x: [N,T,2] positions or [N,T,4] states."""
    x = _f32(x)
    n, t, d = x.shape
    dev = x.device
    h = torch.empty(n, H, device=dev)
    c = torch.empty(n, H, device=dev)
    y = torch.empty(n, t, H, device=dev) if want_y else None
    xl = torch.empty(n, 4, device=dev) if want_x_last else None
    sg = sx = None
    if stash:
        sg = torch.empty(t, n_tiles(n), 5, H, 32, device=dev)
        sx = torch.empty(t, n_tiles(n), 68, 32, device=dev)
    code = _lib.lib().sw_lstm_seq_fwd(_lib.ptr(_f32(lstm_pack)), _lib.ptr(x), d, n, t,
                                      _lib.ptr(None if h_in is None else _f32(h_in)),
                                      _lib.ptr(None if c_in is None else _f32(c_in)),
                                      _lib.ptr(y), _lib.ptr(h), _lib.ptr(c), _lib.ptr(xl),
                                      _lib.ptr(sg), _lib.ptr(sx), sm_count(dev), _stream())
    _lib.check(code, "sw_lstm_seq_fwd")
    return dict(h=h, c=c, y=y, x_last=xl, stash_gates=sg, stash_xh=sx)


def lstm_seq_tcx(enc_w16, x):
    """sw_lstm_seq_fwd_tcx: tensor-core observation encoder (zero initial state, inference only)."""
    x = _f32(x)
    n, t, d = x.shape
    dev = x.device
    if enc_w16.dtype != torch.float16 or not enc_w16.is_contiguous():
        raise ValueError("lstm_seq_tcx: fp16 contiguous weight pack expected (packing.pack_encoder_tcx)")
    h, c, xl = torch.empty(n, H, device=dev), torch.empty(n, H, device=dev), torch.empty(n, 4, device=dev)
    if enc_w16.numel() != 2 * 64 * 256 + 256 * 16:
        raise ValueError("lstm_seq_tcx: pack size does not match the kernel's (packing.pack_encoder_tcx)")
    code = _lib.lib().sw_lstm_seq_fwd_tcx(enc_w16.data_ptr(), _lib.ptr(x), d, n, t, _lib.ptr(h),
                                          _lib.ptr(c), _lib.ptr(xl), sm_count(dev), _stream())
    _lib.check(code, "sw_lstm_seq_fwd_tcx")
    return dict(h=h, c=c, x_last=xl)


def lstm_seq_bwd(lstm_pack_t, stash_gates, dh_last, dc_last, n_rows, want_dx=False):
    """sw_lstm_seq_bwd -> gate gradients [T, tiles, 256, 32] (and dL/dx [N,T,4] if asked)."""
    t = stash_gates.shape[0]
    dev = stash_gates.device
    g = torch.empty(t, n_tiles(n_rows), 256, 32, device=dev)
    dx = torch.empty(n_rows, t, 4, device=dev) if want_dx else None
    code = _lib.lib().sw_lstm_seq_bwd(_lib.ptr(_f32(lstm_pack_t)), _lib.ptr(stash_gates),
                                      _lib.ptr(None if dh_last is None else _f32(dh_last)),
                                      _lib.ptr(None if dc_last is None else _f32(dc_last)),
                                      _lib.ptr(g), _lib.ptr(dx), n_rows, t, sm_count(dev), _stream())
    _lib.check(code, "sw_lstm_seq_bwd")
    return (g, dx) if want_dx else g


def pool(pool_pack, x_last, h, ub, scenes, want_attn=False):
    """sw_pool_fwd -> pooled [N,64] (and the softmax weights [N, round4(max_scene)] if asked)."""
    n = h.shape[0]
    pooled = torch.empty(n, H, device=h.device)
    attn = torch.zeros(n, (scenes.max_scene + 3) // 4 * 4, device=h.device) if want_attn else None
    code = _lib.lib().sw_pool_fwd(_lib.ptr(_f32(pool_pack)), _lib.ptr(_f32(x_last)), _lib.ptr(_f32(h)),
                                  _lib.ptr(_f32(ub)), _lib.ptr(scenes.offsets), _lib.ptr(scenes.agent_scene),
                                  _lib.ptr(pooled), _lib.ptr(attn), n, scenes.max_scene, _stream())
    _lib.check(code, "sw_pool_fwd")
    return (pooled, attn) if want_attn else pooled


def pool_tcx_max_scene():
    return _lib.lib().sw_pool_tcx_max_scene()


def pool_tcx(pool_pack, pool_w16, x_last, h, ub, scenes):
    """sw_pool_fwd_tcx: the pooling kernel with layer 2 of the pair MLP on tcgen05 (fp16 hi/lo split operands,
    fp32-faithful); inference only, scenes of up to pool_tcx_max_scene() agents."""
    n = h.shape[0]
    if pool_w16.dtype != torch.float16 or not pool_w16.is_contiguous() or pool_w16.numel() != 4096:
        raise ValueError("pool_tcx: pool_w16 must be the contiguous fp16 [4096] pack of packing.pack_pool_tcx")
    pooled = torch.empty(n, H, device=h.device)
    units, span, pairs = scenes.pool_units()
    code = _lib.lib().sw_pool_fwd_tcx(_lib.ptr(_f32(pool_pack)), pool_w16.data_ptr(), _lib.ptr(_f32(x_last)), _lib.ptr(_f32(h)),
                                      _lib.ptr(_f32(ub)), _lib.ptr(scenes.offsets), _lib.ptr(scenes.agent_scene),
                                      _lib.ptr(pooled), units.data_ptr(), units.shape[0], span, pairs, n, scenes.max_scene,
                                      _stream())
    _lib.check(code, "sw_pool_fwd_tcx")
    return pooled


def pool_bwd(pool_pack, x_last, h, ub, d_pooled, tdot, attn, scenes):
    """sw_pool_bwd -> (dub [N,65], dh_direct [N,64], per-pair stashes A1 [P,32], G2 [P,64], G1 [P,32], F [P,4])."""
    n, dev, p = h.shape[0], h.device, scenes.n_pairs
    dub = torch.empty(n, 65, device=dev)
    dh = torch.empty(n, H, device=dev)
    st_a1, st_g2 = torch.zeros(p, 32, device=dev), torch.zeros(p, 64, device=dev)
    st_g1, st_f = torch.zeros(p, 32, device=dev), torch.zeros(p, 4, device=dev)
    code = _lib.lib().sw_pool_bwd(_lib.ptr(_f32(pool_pack)), _lib.ptr(_f32(x_last)), _lib.ptr(_f32(h)), _lib.ptr(_f32(ub)),
                                  _lib.ptr(_f32(d_pooled)), _lib.ptr(_f32(tdot)), None, _lib.ptr(_f32(attn)),
                                  _lib.ptr(scenes.offsets), _lib.ptr(scenes.agent_scene), _lib.ptr(scenes.pair_offsets),
                                  _lib.ptr(dub), _lib.ptr(dh), _lib.ptr(st_a1), _lib.ptr(st_g2), _lib.ptr(st_g1),
                                  _lib.ptr(st_f), n, scenes.max_scene, _stream())
    _lib.check(code, "sw_pool_bwd")
    return dub, dh, st_a1, st_g2, st_g1, st_f


def decode(lstm_pack, dec_pack, h0, c0, pooled, noise, x_last, n_next, out=None, stash=False, stash_bufs=None):
    """sw_decode_fwd.  noise [K,N,32] -> out [K,N,n_next,4] (plus the backward stash if asked)."""
    noise = _f32(noise)
    k, n, z = noise.shape
    if z != Z or h0.shape != (n, H):
        raise ValueError("decode: noise must be [K, N, 32] and h0 [N, 64]")
    dev = noise.device
    if out is None:
        out = torch.empty(k, n, n_next, 4, device=dev)
    st = [None] * 5
    if stash_bufs is not None:          # caller-owned stash (xh, gates, a1, a2, sz): the native training step
        st = list(stash_bufs)
    elif stash:
        tl = n_tiles(k * n)
        st = [torch.empty(n_next, tl, 68, 32, device=dev), torch.empty(max(n_next - 1, 1), tl, 5, H, 32, device=dev),
              torch.empty(n_next, tl, 160, 32, device=dev), torch.empty(n_next, tl, 80, 32, device=dev), None]
    code = _lib.lib().sw_decode_fwd(_lib.ptr(_f32(lstm_pack)), _lib.ptr(_f32(dec_pack)), _lib.ptr(_f32(h0)),
                                    _lib.ptr(_f32(c0)), _lib.ptr(None if pooled is None else _f32(pooled)),
                                    _lib.ptr(noise), _lib.ptr(_f32(x_last)), _lib.ptr(out),
                                    _lib.ptr(st[0]), _lib.ptr(st[1]), _lib.ptr(st[2]), _lib.ptr(st[3]), _lib.ptr(st[4]),
                                    n, k, n_next, sm_count(dev), _stream())
    _lib.check(code, "sw_decode_fwd")
    if stash:
        return out, dict(xh=st[0], gates=st[1][:n_next - 1], a1=st[2], a2=st[3])
    return out


def decode_tc(tc_w16, tc_f32, h0, c0, pooled, noise, x_last, n_next, out=None):
    """sw_decode_fwd_tc: the tcgen05 / TMEM decode kernel (bf16 operands, fp32 accumulate)."""
    noise = _f32(noise)
    k, n, z = noise.shape
    if z != Z or h0.shape != (n, H):
        raise ValueError("decode_tc: noise must be [K, N, 32] and h0 [N, 64]")
    if tc_w16.dtype != torch.bfloat16 or not tc_w16.is_contiguous():
        raise ValueError("decode_tc: tc_w16 must be a contiguous bf16 tensor (packing.pack_decoder_tc)")
    if out is None:
        out = torch.empty(k, n, n_next, 4, device=noise.device)
    code = _lib.lib().sw_decode_fwd_tc(tc_w16.data_ptr(), _lib.ptr(_f32(tc_f32)), _lib.ptr(_f32(h0)), _lib.ptr(_f32(c0)),
                                       _lib.ptr(None if pooled is None else _f32(pooled)), _lib.ptr(noise),
                                       _lib.ptr(_f32(x_last)), _lib.ptr(out), n, k, n_next, sm_count(noise.device),
                                       _stream())
    _lib.check(code, "sw_decode_fwd_tc")
    return out


_TCX_SIZES = None


def _tcx_pack_sizes():
    global _TCX_SIZES
    if _TCX_SIZES is None:
        import ctypes
        a, b, c = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        _lib.check(_lib.lib().sw_decode_tcx_pack_sizes(ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)),
                   "sw_decode_tcx_pack_sizes")
        _TCX_SIZES = (a.value, b.value, c.value)
    return _TCX_SIZES


def decode_tcx(w16, wsz16, f32, h0, c0, pooled, noise, x_last, n_next, out=None, status=None):
    """sw_decode_fwd_tcx: tcgen05 decode kernel on fp16 hi/lo split operands (fp32-faithful).  status: device int32[1] that
    receives bit 0 when an operand left fp16's exponent range (then the result is not trustworthy)."""
    noise = _f32(noise)
    k, n, z = noise.shape
    if z != Z or h0.shape != (n, H):
        raise ValueError("decode_tcx: noise must be [K, N, 32] and h0 [N, 64]")
    for t in (w16, wsz16):
        if t.dtype != torch.float16 or not t.is_contiguous():
            raise ValueError("decode_tcx: fp16 contiguous operand packs expected (packing.pack_decoder_tcx)")
    if (w16.numel(), wsz16.numel(), f32.numel()) != _tcx_pack_sizes():
        raise ValueError(f"decode_tcx: pack sizes {(w16.numel(), wsz16.numel(), f32.numel())} do not match the kernel's "
                         f"{_tcx_pack_sizes()} (packing.pack_decoder_tcx / sw_decode_tcx_pack_sizes)")
    if out is None:
        out = torch.empty(k, n, n_next, 4, device=noise.device)
    code = _lib.lib().sw_decode_fwd_tcx(w16.data_ptr(), wsz16.data_ptr(), _lib.ptr(_f32(f32)), _lib.ptr(_f32(h0)),
                                        _lib.ptr(_f32(c0)), _lib.ptr(None if pooled is None else _f32(pooled)),
                                        _lib.ptr(noise), _lib.ptr(_f32(x_last)), _lib.ptr(out),
                                        None if status is None else status.data_ptr(), n, k, n_next,
                                        sm_count(noise.device), _stream())
    _lib.check(code, "sw_decode_fwd_tcx")
    return out


_PAIR_SIZES = None
_PAIR_SCRATCH = {}


def _pair_pack_sizes():
    global _PAIR_SIZES
    if _PAIR_SIZES is None:
        import ctypes
        a, b = ctypes.c_int(), ctypes.c_int()
        _lib.check(_lib.lib().sw_decode_pair_pack_sizes(ctypes.byref(a), ctypes.byref(b)), "sw_decode_pair_pack_sizes")
        _PAIR_SIZES = (a.value, b.value)
    return _PAIR_SIZES


def decode_pair_scratch(device):
    """Per-device scratch of sw_decode_fwd_pair (the hoisted layer-1 term of the tiles in flight: 160 KB per CTA, 23.7 MB for
    148 SMs; written and re-read inside one launch, so it stays in L2).  One buffer per (device, stream)."""
    key = (device.index if device.index is not None else torch.cuda.current_device(), torch.cuda.current_stream(device).cuda_stream)
    if key not in _PAIR_SCRATCH:
        n = int(_lib.lib().sw_decode_pair_scratch_bytes(sm_count(device)))
        _PAIR_SCRATCH[key] = torch.empty(n, dtype=torch.uint8, device=device)
    return _PAIR_SCRATCH[key]


def decode_pair(w16, f32, h0, c0, pooled, noise, x_last, n_next, out=None, status=None, scratch=None, bf16=False):
    """sw_decode_fwd_pair: the fp16 hi/lo split tcgen05 decode kernel with two tiles in flight per SM (CTA pairs, cta_group::2,
    epilogue warps in ping-pong over two tile slots, dedicated issuing warp); same inputs, outputs and arithmetic as decode_tcx.
    Packs from packing.pack_decoder_pair.  bf16=True: sw_decode_fwd_pair_bf16, the same kernel on single bf16 operands (fast mode,
    ~3e-3; pack with pack_decoder_pair(..., bf16=True))."""
    noise = _f32(noise)
    k, n, z = noise.shape
    if z != Z or h0.shape != (n, H):
        raise ValueError("decode_pair: noise must be [K, N, 32] and h0 [N, 64]")
    if w16.dtype != torch.float16 or not w16.is_contiguous():
        raise ValueError("decode_pair: fp16 contiguous operand pack expected (packing.pack_decoder_pair)")
    if (w16.numel(), f32.numel()) != _pair_pack_sizes():
        raise ValueError(f"decode_pair: pack sizes {(w16.numel(), f32.numel())} do not match the kernel's {_pair_pack_sizes()}")
    if out is None:
        out = torch.empty(k, n, n_next, 4, device=noise.device)
    if scratch is None:
        scratch = decode_pair_scratch(noise.device)
    h0, c0 = _f32(h0), _f32(c0)
    if h0.data_ptr() % 32:      # the kernel reads the state rows in 32-byte pieces (a view at an odd element offset of a larger buffer)
        h0 = h0.clone()
    if c0.data_ptr() % 32:
        c0 = c0.clone()
    entry = _lib.lib().sw_decode_fwd_pair_bf16 if bf16 else _lib.lib().sw_decode_fwd_pair
    code = entry(w16.data_ptr(), _lib.ptr(_f32(f32)), _lib.ptr(h0), _lib.ptr(c0),
                 _lib.ptr(None if pooled is None else _f32(pooled)), _lib.ptr(noise),
                 _lib.ptr(_f32(x_last)), _lib.ptr(out), scratch.data_ptr(), scratch.numel(),
                 None if status is None else status.data_ptr(), n, k, n_next,
                 sm_count(noise.device), _stream())
    _lib.check(code, "sw_decode_fwd_pair_bf16" if bf16 else "sw_decode_fwd_pair")
    return out


def decode_bwd(lstm_pack_t, dec_pack_t, c0, stash, d_out, n_agents, n_samples):
    """sw_decode_bwd.  d_out [K*N, T, 4] -> dict of gradient images + dh0/dc0 [K*N, 64]."""
    d_out = _f32(d_out)
    t = d_out.shape[-2]
    rows = n_agents * n_samples
    dev = d_out.device
    tl = n_tiles(rows)
    g_gates = torch.empty(max(t - 1, 1), tl, 256, 32, device=dev)
    g_a1 = torch.empty(t, tl, 160, 32, device=dev)
    g_a2 = torch.empty(t, tl, 80, 32, device=dev)
    g_v = torch.empty(t, tl, 2, 32, device=dev)
    dh0 = torch.empty(rows, H, device=dev)
    dc0 = torch.empty(rows, H, device=dev)
    gates = stash["gates"] if t > 1 else None
    code = _lib.lib().sw_decode_bwd(_lib.ptr(_f32(lstm_pack_t)), _lib.ptr(_f32(dec_pack_t)), _lib.ptr(_f32(c0)),
                                    _lib.ptr(gates), _lib.ptr(stash["a1"]), _lib.ptr(stash["a2"]), _lib.ptr(d_out),
                                    _lib.ptr(g_gates), _lib.ptr(g_a1), _lib.ptr(g_a2), _lib.ptr(g_v), _lib.ptr(dh0),
                                    _lib.ptr(dc0), None, None, None, n_agents, n_samples, t, sm_count(dev), _stream())
    _lib.check(code, "sw_decode_bwd")
    return dict(gates=g_gates[:t - 1], a1=g_a1, a2=g_a2, v=g_v, dh0=dh0, dc0=dc0)


def disc_heads_fwd(pack, h, pred, n_latent=2, record=False):
    """sw_disc_heads_fwd: h [N,64], pred [N,P] -> label [N,1], code [N,n_latent] (+ the activation record)."""
    h, pred = _f32(h), _f32(pred)
    n, p = pred.shape
    dev = h.device
    label = torch.empty(n, 1, device=dev)
    code = torch.empty(n, n_latent, device=dev)
    xrec = torch.empty(n, 257 + p, device=dev) if record else None
    rc = _lib.lib().sw_disc_heads_fwd(_lib.ptr(_f32(pack)), _lib.ptr(h), _lib.ptr(pred), p, n_latent, n, _lib.ptr(label),
                                      _lib.ptr(code), _lib.ptr(xrec), _stream())
    _lib.check(rc, "sw_disc_heads_fwd")
    return label, code, xrec


def disc_heads_bwd(pack, xrec, pred_dim, d_label, d_code, n_latent=2, want_dh=True, want_dpred=True):
    """sw_disc_heads_bwd -> (d_h [N,64] | None, d_pred [N,P] | None, gradient record G [N, 193 + n_latent])."""
    n, dev = xrec.shape[0], xrec.device
    d_h = torch.empty(n, H, device=dev) if want_dh else None
    d_pred = torch.empty(n, pred_dim, device=dev) if want_dpred else None
    grec = torch.empty(n, 193 + n_latent, device=dev)
    rc = _lib.lib().sw_disc_heads_bwd(_lib.ptr(_f32(pack)), _lib.ptr(xrec), pred_dim, n_latent, n,
                                      _lib.ptr(None if d_label is None else _f32(d_label)),
                                      _lib.ptr(None if d_code is None else _f32(d_code)), _lib.ptr(d_h), _lib.ptr(d_pred),
                                      _lib.ptr(grec), _stream())
    _lib.check(rc, "sw_disc_heads_bwd")
    return d_h, d_pred, grec


def rows_linear(x, w, bias=None, add1=None, add2=None, out=None):
    """sw_rows_linear: out[row] = bias + add1[row] + add2[row] + x[row] @ w   (x [N,K], w [K,M] row-major, K, M <= 80)."""
    x, w = _f32(x), _f32(w)
    n, k = x.shape
    m = w.shape[1]
    if out is None:
        out = torch.empty(n, m, device=x.device)
    code = _lib.lib().sw_rows_linear(_lib.ptr(x), k, _lib.ptr(w), _lib.ptr(None if bias is None else _f32(bias)),
                                     _lib.ptr(None if add1 is None else _f32(add1)), _lib.ptr(None if add2 is None else _f32(add2)),
                                     _lib.ptr(out), m, n, k, m, _stream())
    _lib.check(code, "sw_rows_linear")
    return out


def noise_uniform(shape, device, seed, offset=0, out=None, first_element=0):
    """sw_noise_uniform: uniform [0, 1) fp32 noise drawn on the device (Philox4x32-10 keyed by seed, counter offset).
    first_element (multiple of 4): the output is the slice [first_element, first_element + n) of the logical stream."""
    if out is None:
        out = torch.empty(*shape, device=device)
    if first_element % 4:
        raise ValueError("noise_uniform: first_element must be a multiple of 4")
    code = _lib.lib().sw_noise_uniform(_lib.ptr(out), out.numel(), int(seed) & (2 ** 64 - 1), int(offset) & (2 ** 64 - 1),
                                       first_element // 4, sm_count(out.device), _stream())
    _lib.check(code, "sw_noise_uniform")
    return out


def bestofk_metrics(pred, gt, ss):
    """sw_bestofk_metrics.  pred [K,N,T,4], gt [N,T,2] -> [N,4] (avg ADE, avg FDE, min ADE, min FDE)."""
    pred, gt = _f32(pred), _f32(gt)
    k, n, t, _ = pred.shape
    out = torch.empty(n, 4, device=pred.device)
    code = _lib.lib().sw_bestofk_metrics(_lib.ptr(pred), _lib.ptr(gt), float(ss), n, k, t, _lib.ptr(out), _stream())
    _lib.check(code, "sw_bestofk_metrics")
    return out
