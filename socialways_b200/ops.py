"""Thin tensor-level wrappers over the C-ABI (include/socialways_b200.h): allocate outputs with torch,
pass raw device pointers + the current CUDA stream.  No arithmetic happens here."""
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


def lstm_seq(lstm_pack, x, h_in=None, c_in=None, want_y=False, want_x_last=False, stash=False):
    """sw_lstm_seq_fwd.  x: [N,T,2] positions or [N,T,4] states."""
    x = _f32(x)
    n, t, d = x.shape
    dev = x.device
    h = torch.empty(n, H, device=dev)
    c = torch.empty(n, H, device=dev)
    y = torch.empty(n, t, H, device=dev) if want_y else None
    xl = torch.empty(n, 4, device=dev) if want_x_last else None
    sg = sh = sx = None
    if stash:
        sg = torch.empty(t, n, H, 5, device=dev)
        sh = torch.empty(t, n, H, device=dev)
        sx = torch.empty(t, n, 4, device=dev)
    code = _lib.lib().sw_lstm_seq_fwd(_lib.ptr(_f32(lstm_pack)), _lib.ptr(x), d, n, t,
                                      _lib.ptr(None if h_in is None else _f32(h_in)),
                                      _lib.ptr(None if c_in is None else _f32(c_in)),
                                      _lib.ptr(y), _lib.ptr(h), _lib.ptr(c), _lib.ptr(xl),
                                      _lib.ptr(sg), _lib.ptr(sh), _lib.ptr(sx), sm_count(dev), _stream())
    _lib.check(code, "sw_lstm_seq_fwd")
    return dict(h=h, c=c, y=y, x_last=xl, stash_gates=sg, stash_h=sh, stash_x4=sx)


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


def decode(lstm_pack, dec_pack, h0, c0, pooled, noise, x_last, n_next, out=None):
    """sw_decode_fwd.  noise [K,N,32] -> out [K,N,n_next,4]."""
    noise = _f32(noise)
    k, n, z = noise.shape
    if z != Z or h0.shape != (n, H):
        raise ValueError("decode: noise must be [K, N, 32] and h0 [N, 64]")
    if out is None:
        out = torch.empty(k, n, n_next, 4, device=noise.device)
    code = _lib.lib().sw_decode_fwd(_lib.ptr(_f32(lstm_pack)), _lib.ptr(_f32(dec_pack)), _lib.ptr(_f32(h0)),
                                    _lib.ptr(_f32(c0)), _lib.ptr(None if pooled is None else _f32(pooled)),
                                    _lib.ptr(noise), _lib.ptr(_f32(x_last)), _lib.ptr(out), n, k, n_next,
                                    sm_count(noise.device), _stream())
    _lib.check(code, "sw_decode_fwd")
    return out


def bestofk_metrics(pred, gt, ss):
    """sw_bestofk_metrics.  pred [K,N,T,4], gt [N,T,2] -> [N,4] (avg ADE, avg FDE, min ADE, min FDE)."""
    pred, gt = _f32(pred), _f32(gt)
    k, n, t, _ = pred.shape
    out = torch.empty(n, 4, device=pred.device)
    code = _lib.lib().sw_bestofk_metrics(_lib.ptr(pred), _lib.ptr(gt), float(ss), n, k, t, _lib.ptr(out), _stream())
    _lib.check(code, "sw_bestofk_metrics")
    return out
