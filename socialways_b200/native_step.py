"""One iteration of train() (reference train.py:470-551) as ~30 kernel launches behind the C-ABI, no autograd.

The autograd path (autograd_path.py + trainer.train()) leaves everything around the forward / backward kernels to
torch: weight folds re-traced every iteration, ~100 small cat / copy / add kernels, cuBLAS for every weight gradient,
ATen reductions for the bias gradients and the losses -- ~580 launches per iteration.  Here the iteration is a fixed
sequence of this library's own kernels on preallocated buffers:

    generator forward (ONCE per iteration -- the reference evaluates predict() three times per iteration, train.py:480,
        507, with identical weights, noise and therefore identical results; the D passes and the G pass share one):
        sw_gen_pack -> sw_lstm_seq_fwd -> [sw_rows_linear -> sw_pool_fwd] -> sw_decode_fwd
    D pass x (n_unrolling_steps + 1)  (train.py:476-499):
        sw_disc_pack -> sw_lstm_seq_fwd -> sw_disc_step(mode 0: heads fwd + losses + heads bwd) -> sw_lstm_seq_bwd
        -> sw_contract (all 20 parameter gradients) -> Adam (FlatAdam: one kernel, carries the all-reduce under torchrun)
    G pass  (train.py:501-543):
        sw_disc_pack -> sw_lstm_seq_fwd -> sw_disc_step(mode 1) -> sw_decode_bwd -> [sw_pool_bwd -> sw_rows_linear]
        -> sw_lstm_seq_bwd -> sw_contract -> sw_gen_pack_bwd -> Adam -> D.load(backup) -> sw_train_stats

The sequence contains no host synchronisation and no allocation, so it is captured into ONE CUDA graph per mini-batch
shape.  Same quirks as the reference: the unrolled-D rollback restores nn.Linear layers only (train.py:311-316), the
same noise feeds every pass, the D gradients of g_loss.backward() are discarded (D.zero_grad(), train.py:478).
"""
import ctypes

import numpy as np
import torch

from . import _lib
from .ops import _stream, sm_count

H = 64


class ContractSeg(ctypes.Structure):
    """sw_contract_seg (include/sw_contract.h)."""
    _fields_ = [("out", ctypes.c_void_p), ("out2", ctypes.c_void_p), ("k_begin", ctypes.c_int), ("k_count", ctypes.c_int),
                ("out_sk", ctypes.c_int), ("out_sn", ctypes.c_int)]


class ContractJob(ctypes.Structure):
    """sw_contract_job (include/sw_contract.h)."""
    _fields_ = [("a", ctypes.c_void_p), ("b", ctypes.c_void_p), ("a_stride", ctypes.c_longlong), ("b_stride", ctypes.c_longlong),
                ("a_k0", ctypes.c_int), ("K", ctypes.c_int), ("b_n0", ctypes.c_int), ("N", ctypes.c_int),
                ("n_images", ctypes.c_int), ("n_rows", ctypes.c_int), ("a_kind", ctypes.c_int), ("b_kind", ctypes.c_int),
                ("ones_row", ctypes.c_int), ("n_perm", ctypes.c_int), ("n_segs", ctypes.c_int), ("reserved", ctypes.c_int),
                ("seg", ContractSeg * 3)]


IMAGE, ROWS = 0, 1


def _ptr(t):
    return None if t is None else (t if isinstance(t, int) else t.data_ptr())


def seg(out, k_begin, k_count, out_sk, out_sn, out2=None):
    """Rows [k_begin, k_begin + k_count) of the job's result go to out[(k - k_begin) * out_sk + n * out_sn]."""
    return ContractSeg(_ptr(out), _ptr(out2), k_begin, k_count, out_sk, out_sn)


def _job(a, b, K, N, n_images, segs, a_stride=0, b_stride=0, a_k0=0, b_n0=0, a_kind=IMAGE, b_kind=IMAGE, ones=False, perm=0,
         n_rows=None):
    """G[k][n] = sum_rows A[row][a_k0 + k] B[row][b_n0 + n] (+ the all-ones row K when `ones`), scattered by `segs`."""
    arr = (ContractSeg * 3)(*segs)
    return ContractJob(_ptr(a), _ptr(b), a_stride, b_stride, a_k0, K, b_n0, N, n_images,
                       n_images * 32 if n_rows is None else n_rows, a_kind, b_kind, 1 if ones else 0, perm, len(segs), 0, arr)


class ContractPlan:
    """A job list + its workspace, ready to launch (sw_contract_tc, or the FFMA sw_contract)."""

    # below this many 32-row images in the largest job the launch is latency-bound and the FFMA kernel (3 small CTAs per SM,
    # no TMEM set-up) is quicker: measured crossover between 16 images (batch 256: 0.40 vs 0.43 ms per iteration) and 256
    # images (batch 4 096: 0.62 vs 0.56 ms), profiles/r2_train_probe.txt
    TC_MIN_IMAGES = 96

    def __init__(self, jobs, device, tensor_cores=True):
        self.n = len(jobs)
        self.tc = tensor_cores == "force" or (bool(tensor_cores) and max(j.n_images for j in jobs) >= self.TC_MIN_IMAGES)
        self.jobs = (ContractJob * self.n)(*jobs)
        ws, nc = ctypes.c_longlong(), ctypes.c_int()
        _lib.check(_lib.lib().sw_contract_plan(self.jobs, self.n, sm_count(device), 1 if self.tc else 0, ctypes.byref(ws),
                                               ctypes.byref(nc)), "sw_contract_plan")
        self.ws = torch.empty(max(int(ws.value), 4), device=device)
        self.counters = torch.zeros(max(int(nc.value), 1), dtype=torch.int32, device=device)
        self.device = device

    def run(self):
        fn = _lib.lib().sw_contract_tc if self.tc else _lib.lib().sw_contract
        _lib.check(fn(self.jobs, self.n, self.ws.data_ptr(), self.ws.numel(), self.counters.data_ptr(), self.counters.numel(),
                      sm_count(self.device), _stream()), "sw_contract_tc" if self.tc else "sw_contract")


def _ptr_array(tensors):
    return (ctypes.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


class NativePacks:
    """Packed-weight buffers of the generator and the discriminator (one set per trainer)."""

    def __init__(self, trainer):
        lib, dev = _lib.lib(), trainer.device
        self.gen_params = list(trainer.generator.optimizer_parameters())
        self.disc_params = list(trainer.D.parameters())
        if len(self.gen_params) != 22 or len(self.disc_params) != 20:
            raise _lib.SocialWaysCudaError("native step: unexpected parameter lists (22 generator / 20 discriminator tensors)")
        sz = [ctypes.c_int() for _ in range(7)]
        _lib.check(lib.sw_gen_pack_sizes(*[ctypes.byref(s) for s in sz]), "sw_gen_pack_sizes")
        names = ("enc", "enc_t", "dec", "dec_t", "pool", "pool_m", "pool_mt")
        for name, s in zip(names, sz):
            setattr(self, name, torch.empty(s.value, device=dev))
        self.pred_dim = trainer.n_next * 4
        sd = [ctypes.c_int() for _ in range(3)]
        _lib.check(lib.sw_disc_pack_sizes(self.pred_dim, *[ctypes.byref(s) for s in sd]), "sw_disc_pack_sizes")
        self.d_lstm, self.d_lstm_t, self.d_heads = (torch.empty(s.value, device=dev) for s in sd)
        self.device = dev
        # D.load(backup) (train.py:311-316): the nn.Linear parameters = every discriminator tensor after the LSTM's four
        lin = self.disc_params[4:]
        flat = trainer.D_optimizer.flat_p
        off0 = (lin[0].data_ptr() - flat.data_ptr()) // 4
        off1 = (lin[-1].data_ptr() - flat.data_ptr()) // 4 + lin[-1].numel()
        if off1 - off0 != sum(p.numel() for p in lin):
            raise _lib.SocialWaysCudaError("native step: the discriminator's Linear parameters are not contiguous in the flat buffer")
        self.d_linear = flat[off0:off1]

    def pack_generator(self):
        _lib.check(_lib.lib().sw_gen_pack(_ptr_array(self.gen_params), self.enc.data_ptr(), self.enc_t.data_ptr(),
                                          self.dec.data_ptr(), self.dec_t.data_ptr(), self.pool.data_ptr(),
                                          self.pool_m.data_ptr(), self.pool_mt.data_ptr(), _stream()), "sw_gen_pack")

    def pack_discriminator(self):
        _lib.check(_lib.lib().sw_disc_pack(_ptr_array(self.disc_params), self.pred_dim, self.d_lstm.data_ptr(),
                                           self.d_lstm_t.data_ptr(), self.d_heads.data_ptr(), _stream()), "sw_disc_pack")


class NativeStep:
    """Buffers + launch sequence of one mini-batch shape (bs rows of this rank, `scenes` = ops.SceneIndex)."""

    def __init__(self, trainer, packs, bs, scenes, global_bs):
        self.tr, self.pk, self.bs, self.scenes = trainer, packs, bs, scenes
        dev = self.dev = trainer.device
        To, Tp = trainer.n_past, trainer.n_next
        self.To, self.Tp, self.P = To, Tp, Tp * 4
        t32, t16 = (bs + 31) // 32, (bs + 15) // 16
        self.t32, self.t16 = t32, t16
        self.inv_n = 1.0 / float(global_bs)
        self.info_w = float(trainer.loss_info_w) if trainer.use_info_loss else 0.0
        self.social = bool(trainer.generator.use_social)
        f = lambda *shape: torch.empty(*shape, device=dev)
        # inputs (filled by the caller before every launch / replay)
        self.obsv, self.pred = f(bs, To, 2), f(bs, Tp, 2)
        self.noise, self.targets = f(bs, trainer.noise_len), f(2)
        self._targets_pin = [torch.empty(2).pin_memory() for _ in range(4)]     # rotating: an upload may still be in flight
        self._targets_i = 0
        # generator forward + stash; observation pass and decode steps share one image sequence so that the LSTM weight
        # gradient is ONE contraction job
        self.h, self.c, self.x_last = f(bs, H), f(bs, H), f(bs, 4)
        self.xh_all = f(To + Tp, t32, 68, 32)
        self.gates_all = f(To + max(Tp - 1, 1), t32, 5, H, 32)
        self.g_gates_all = f(To + max(Tp - 1, 1), t32, 256, 32)
        self.ub, self.pooled = f(bs, 65), f(bs, H)
        self.attn = torch.zeros(bs, (scenes.max_scene + 3) // 4 * 4, device=dev)
        self.out = f(1, bs, Tp, 4)
        self.s_a1, self.s_a2, self.s_sz = f(Tp, t32, 160, 32), f(Tp, t32, 80, 32), f(t32, 96, 32)
        # discriminator pass
        xr, gr = ctypes.c_int(), ctypes.c_int()
        _lib.check(_lib.lib().sw_disc_step_image_rows(self.P, ctypes.byref(xr), ctypes.byref(gr)), "sw_disc_step_image_rows")
        self.xr, self.gr = xr.value, gr.value
        self.obsv_h, self.c_tmp, self.d_h = f(bs, H), f(bs, H), f(bs, H)
        self.d_gates, self.d_xh, self.d_g_gates = f(To, t32, 5, H, 32), f(To, t32, 68, 32), f(To, t32, 256, 32)
        self.x_img, self.g_img = f(t16, self.xr, 32), f(t16, self.gr, 32)
        self.loss_d, self.loss_g = torch.zeros(t16, 4, device=dev), torch.zeros(t32, 4, device=dev)
        # generator backward
        self.d_pred = f(bs, self.P)
        self.g_a1, self.g_a2, self.g_v = f(Tp, t32, 160, 32), f(Tp, t32, 80, 32), f(Tp, t32, 2, 32)
        self.g_a1sum = f(t32, 160, 32)
        self.dh0, self.dc0, self.d_pooled, self.dh_total = f(bs, H), f(bs, H), f(bs, H), f(bs, H)
        npairs = max(scenes.n_pairs, 1)
        self.n_pairs = scenes.n_pairs
        self.dub, self.dh_direct = f(bs, 65), f(bs, H)
        # pairs of one-agent scenes are never written by sw_pool_bwd: zero once, stays zero
        self.st_a1, self.st_g2 = torch.zeros(npairs, 32, device=dev), torch.zeros(npairs, 64, device=dev)
        self.st_g1, self.st_f = torch.zeros(npairs, 32, device=dev), torch.zeros(npairs, 4, device=dev)
        self.d_enc, self.d_w34, self.d_m = f(69, 256), f(162), f(65 * 65)
        self.stats = torch.zeros(8, device=dev)
        self.stats_partial = torch.zeros(sm_count(dev), 2, device=dev)
        self.stats_counter = torch.zeros(1, dtype=torch.int32, device=dev)
        if self.social:
            _ = scenes.pair_offsets
        self.d_linear = packs.d_linear
        self.backup = torch.empty_like(self.d_linear)
        self._build_jobs()

    def set_targets(self, zeros_value, ones_value):
        """The smoothed labels of the iteration (train.py:471-472) -> device, without a host synchronisation."""
        t = self._targets_pin[self._targets_i & 3]
        self._targets_i += 1
        t[0], t[1] = zeros_value, ones_value
        self.targets.copy_(t, non_blocking=True)

    # ---------------------------------------------------------------- contraction job lists
    def _build_jobs(self):
        tr, pk, To, Tp, P, t32, t16 = self.tr, self.pk, self.To, self.Tp, self.P, self.t32, self.t16
        g = lambda p: p.grad
        tc = self.tr.native_tensor_cores
        dp = pk.disc_params
        # ---- discriminator: LSTM(4, 64) as ONE job (w_ih rows 0..3, w_hh rows 4..67, the ones row -> b_ih and b_hh) ----
        jobs = [_job(self.d_xh, self.d_g_gates, 68, 256, To * t32,
                     [seg(g(dp[0]), 0, 4, 1, 4), seg(g(dp[1]), 4, 64, 1, 64), seg(g(dp[2]), 68, 1, 0, 1, out2=g(dp[3]))],
                     68 * 32, 256 * 32, ones=True, perm=1)]
        xs, gs = self.xr * 32, self.gr * 32
        # (X-image rows offset, K, G-image rows offset, N, weight index) per Linear layer, rows per csrc/disc_layout.cuh
        heads = [(0, 64, 0, 32, 4), (64, 32, 32, 32, 6), (96, P, 64, 32, 8), (96 + P, 32, 96, 32, 10),
                 (128 + P, 64, 128, 32, 12), (192 + P, 32, 160, 1, 14), (128 + P, 64, 161, 32, 16), (224 + P, 32, 193, 2, 18)]
        for a0, K, b0, N, wi in heads:          # weight [N][K] (torch [out][in]) + bias [N] from the ones row
            jobs.append(_job(self.x_img, self.g_img, K, N, t16, [seg(g(dp[wi]), 0, K, 1, K), seg(g(dp[wi + 1]), K, 1, 0, 1)],
                             xs, gs, a_k0=a0, b_n0=b0, ones=True))
        self.d_plan = ContractPlan(jobs, self.dev, tc)

        # ---- generator ----
        gp = pk.gen_params
        n_all = (To + Tp - 1) * t32
        dec_xh = self.xh_all[To:]
        w1 = g(gp[14]).view(-1)
        jobs = [
            # encoder LSTM (observation pass + decode steps, one image sequence) -> d_enc [69][256] in pack layout (folds: sw_gen_pack_bwd)
            _job(self.xh_all, self.g_gates_all, 68, 256, n_all, [seg(self.d_enc, 0, 69, 256, 1)], 68 * 32, 256 * 32, ones=True),
            # DecoderFC layer 1: h columns per step; [S ; z] columns and b1 against sum_t da1pre
            _job(dec_xh, self.g_a1, 64, 160, Tp * t32, [seg(w1, 0, 64, 1, 160)], 68 * 32, 160 * 32, a_k0=4),
            _job(self.s_sz, self.g_a1sum, 96, 160, t32, [seg(w1[64:], 0, 96, 1, 160), seg(g(gp[15]), 96, 1, 0, 1)],
                 96 * 32, 160 * 32, ones=True),
            _job(self.s_a1, self.g_a2, 160, 80, Tp * t32, [seg(g(gp[16]), 0, 160, 1, 160), seg(g(gp[17]), 160, 1, 0, 1)],
                 160 * 32, 80 * 32, ones=True),
            # folded 80 -> 2 output layer -> d_w34 [80][2] | d_b34 [2] (folds: sw_gen_pack_bwd)
            _job(self.s_a2, self.g_v, 80, 2, Tp * t32, [seg(self.d_w34, 0, 80, 2, 1), seg(self.d_w34[160:], 80, 1, 0, 1)],
                 80 * 32, 2 * 32, ones=True),
        ]
        if self.social and self.n_pairs > 0:
            npi, bs, npairs = (self.n_pairs + 31) // 32, self.bs, self.n_pairs
            jobs += [
                # pair MLP layer 1: records F = (dist, bearing, dca, 1) -> fc.0.weight [32][3] and fc.0.bias from the stored ones
                _job(self.st_f, self.st_g1, 4, 32, npi, [seg(g(gp[2]), 0, 3, 1, 3), seg(g(gp[3]), 3, 1, 0, 1)], 4, 32,
                     a_kind=ROWS, b_kind=ROWS, n_rows=npairs),
                _job(self.st_a1, self.st_g2, 32, 64, npi, [seg(g(gp[4]), 0, 32, 1, 32), seg(g(gp[5]), 32, 1, 0, 1)], 32, 64,
                     a_kind=ROWS, b_kind=ROWS, ones=True, n_rows=npairs),
                # (u | beta) = h . M + m0 -> d_m [64][65] | d_m0 [65] (folds: sw_gen_pack_bwd)
                _job(self.h, self.dub, 64, 65, (bs + 31) // 32, [seg(self.d_m, 0, 65, 65, 1)], 64, 65, a_kind=ROWS, b_kind=ROWS,
                     ones=True, n_rows=bs),
            ]
        self.g_plan = ContractPlan(jobs, self.dev, tc)

    # ---------------------------------------------------------------- launches
    def _lstm_fwd(self, pack, h, c, x_last, st_g, st_x):
        code = _lib.lib().sw_lstm_seq_fwd(pack.data_ptr(), self.obsv.data_ptr(), 2, self.bs, self.To, None, None, None, h.data_ptr(),
                                          c.data_ptr(), None if x_last is None else x_last.data_ptr(),
                                          None if st_g is None else st_g.data_ptr(), None if st_x is None else st_x.data_ptr(),
                                          sm_count(self.dev), _stream())
        _lib.check(code, "sw_lstm_seq_fwd")

    def _lstm_bwd(self, pack_t, st_g, dh, dc, g_gates):
        code = _lib.lib().sw_lstm_seq_bwd(pack_t.data_ptr(), st_g.data_ptr(), dh.data_ptr(), None if dc is None else dc.data_ptr(),
                                          g_gates.data_ptr(), None, self.bs, self.To, sm_count(self.dev), _stream())
        _lib.check(code, "sw_lstm_seq_bwd")

    def generator_forward(self):
        lib, pk, sc, bs, dev = _lib.lib(), self.pk, self.scenes, self.bs, self.dev
        To = self.To
        pk.pack_generator()
        self._lstm_fwd(pk.enc, self.h, self.c, self.x_last, self.gates_all[:To], self.xh_all[:To])
        pooled = None
        if self.social:
            _lib.check(lib.sw_rows_linear(self.h.data_ptr(), H, pk.pool_m.data_ptr(), pk.pool_m[64 * 65:].data_ptr(), None, None,
                                          self.ub.data_ptr(), 65, bs, H, 65, _stream()), "sw_rows_linear")
            _lib.check(lib.sw_pool_fwd(pk.pool.data_ptr(), self.x_last.data_ptr(), self.h.data_ptr(), self.ub.data_ptr(),
                                       sc.offsets.data_ptr(), sc.agent_scene.data_ptr(), self.pooled.data_ptr(),
                                       self.attn.data_ptr(), bs, sc.max_scene, _stream()), "sw_pool_fwd")
            pooled = self.pooled.data_ptr()
        _lib.check(lib.sw_decode_fwd(pk.enc.data_ptr(), pk.dec.data_ptr(), self.h.data_ptr(), self.c.data_ptr(), pooled,
                                     self.noise.data_ptr(), self.x_last.data_ptr(), self.out.data_ptr(),
                                     self.xh_all[To:].data_ptr(), self.gates_all[To:].data_ptr(), self.s_a1.data_ptr(),
                                     self.s_a2.data_ptr(), self.s_sz.data_ptr(), bs, 1, self.Tp, sm_count(dev), _stream()),
                   "sw_decode_fwd")

    def _disc_step(self, mode):
        pk = self.pk
        d = lambda t, on: t.data_ptr() if on else None
        code = _lib.lib().sw_disc_step(pk.d_heads.data_ptr(), self.P, mode, self.obsv_h.data_ptr(), self.out.data_ptr(),
                                       self.pred.data_ptr(), self.obsv.data_ptr(), self.To, self.noise.data_ptr(),
                                       self.noise.shape[1], self.targets.data_ptr(), self.inv_n, self.info_w,
                                       d(self.d_h, mode == 0), d(self.d_pred, mode == 1), d(self.x_img, mode == 0),
                                       d(self.g_img, mode == 0), (self.loss_d if mode == 0 else self.loss_g).data_ptr(),
                                       None, None, self.bs, sm_count(self.dev), _stream())
        _lib.check(code, "sw_disc_step")

    def discriminator_grads(self):
        """d_loss.backward() (train.py:484-495): every discriminator gradient, written (not accumulated)."""
        pk = self.pk
        pk.pack_discriminator()
        self._lstm_fwd(pk.d_lstm, self.obsv_h, self.c_tmp, None, self.d_gates, self.d_xh)
        self._disc_step(0)
        self._lstm_bwd(pk.d_lstm_t, self.d_gates, self.d_h, None, self.d_g_gates)
        self.d_plan.run()

    def discriminator_pass(self, first):
        self.discriminator_grads()
        self.tr.D_optimizer.step()
        if first and self.tr.n_unrolling_steps > 0:
            self.backup.copy_(self.d_linear)

    def generator_grads(self):
        """g_loss.backward() (train.py:514-538) for the generator's parameters, written (not accumulated)."""
        lib, pk, tr, sc, bs, dev = _lib.lib(), self.pk, self.tr, self.scenes, self.bs, self.dev
        To = self.To
        pk.pack_discriminator()
        self._lstm_fwd(pk.d_lstm, self.obsv_h, self.c_tmp, None, None, None)
        self._disc_step(1)
        w1 = pk.gen_params[14]
        _lib.check(lib.sw_decode_bwd(pk.enc_t.data_ptr(), pk.dec_t.data_ptr(), self.c.data_ptr(),
                                     self.gates_all[To:].data_ptr() if self.Tp > 1 else None, self.s_a1.data_ptr(),
                                     self.s_a2.data_ptr(), self.d_pred.data_ptr(),
                                     self.g_gates_all[To:].data_ptr() if self.Tp > 1 else None, self.g_a1.data_ptr(),
                                     self.g_a2.data_ptr(), self.g_v.data_ptr(), self.dh0.data_ptr(), self.dc0.data_ptr(),
                                     self.g_a1sum.data_ptr(), w1.data_ptr(), self.d_pooled.data_ptr() if self.social else None,
                                     bs, 1, self.Tp, sm_count(dev), _stream()), "sw_decode_bwd")
        dh_enc = self.dh0
        if self.social:
            _lib.check(lib.sw_pool_bwd(pk.pool.data_ptr(), self.x_last.data_ptr(), self.h.data_ptr(), self.ub.data_ptr(),
                                       self.d_pooled.data_ptr(), None, self.pooled.data_ptr(), self.attn.data_ptr(),
                                       sc.offsets.data_ptr(), sc.agent_scene.data_ptr(), sc.pair_offsets.data_ptr(),
                                       self.dub.data_ptr(), self.dh_direct.data_ptr(), self.st_a1.data_ptr(),
                                       self.st_g2.data_ptr(), self.st_g1.data_ptr(), self.st_f.data_ptr(), bs, sc.max_scene,
                                       _stream()), "sw_pool_bwd")
            _lib.check(lib.sw_rows_linear(self.dub.data_ptr(), 65, pk.pool_mt.data_ptr(), None, self.dh0.data_ptr(),
                                          self.dh_direct.data_ptr(), self.dh_total.data_ptr(), H, bs, 65, H, _stream()),
                       "sw_rows_linear")
            dh_enc = self.dh_total
        self._lstm_bwd(pk.enc_t, self.gates_all[:To], dh_enc, self.dc0, self.g_gates_all[:To])
        self.g_plan.run()
        grads = [p.grad for p in pk.gen_params]
        _lib.check(lib.sw_gen_pack_bwd(_ptr_array(pk.gen_params), _ptr_array(grads), self.d_enc.data_ptr(), self.d_w34.data_ptr(),
                                       self.d_m.data_ptr(), 1 if (self.social and self.n_pairs > 0) else 0, _stream()),
                   "sw_gen_pack_bwd")

    def generator_pass(self):
        lib, tr, bs, dev = _lib.lib(), self.tr, self.bs, self.dev
        self.generator_grads()
        tr.predictor_optimizer.step()
        if tr.n_unrolling_steps > 0:
            self.d_linear.copy_(self.backup)                                   # D.load(backup): Linear layers only
        _lib.check(lib.sw_train_stats(self.out.data_ptr(), self.pred.data_ptr(), bs, self.Tp, float(tr.ss), self.loss_d.data_ptr(),
                                      self.t16, self.loss_g.data_ptr(), self.t32, self.inv_n, self.info_w,
                                      self.stats_partial.data_ptr(), self.stats_counter.data_ptr(), self.stats.data_ptr(),
                                      sm_count(dev), _stream()), "sw_train_stats")

    def run(self):
        """The whole iteration (train.py:470-551) on the current contents of obsv / pred / noise / targets."""
        with torch.no_grad():
            self.generator_forward()
            for u in range(self.tr.n_unrolling_steps + 1):
                self.discriminator_pass(first=(u == 0))
            self.generator_pass()

    @staticmethod
    def launches_per_iteration(social, unroll):
        """Kernel launches of run() (memcpy nodes of the D backup / rollback included)."""
        fwd = 3 + (2 if social else 0)
        d_pass = 6
        g_pass = 9 + (2 if social else 0)
        return fwd + d_pass * (unroll + 1) + g_pass + (2 if unroll > 0 else 0)
