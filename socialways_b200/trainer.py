"""train() and test() of the reference (train.py:439-560, :563-616) over the CUDA modules.

Same control flow, RNG consumption order (numpy scalars for the smoothed labels, torch CPU RNG for
the noise, train.py:471-473,584), loss definitions, optimiser settings, unrolling + Linear-only
rollback quirk (train.py:498-499,541-543) and printed lines as the reference.  Differences, all
arithmetic-neutral: the K samples of test() are decoded in one launch (the noise is still drawn
per (scene, k) from the torch CPU RNG in the reference's order), and the discriminator's
observation LSTM is evaluated once per D pass for the fake and the real branch (same weights,
same input -- the reference evaluates it twice with identical results, train.py:482,487).
"""
import copy
import os
import time

import numpy as np
import torch
import torch.nn as nn
import torch.optim as opt

from . import distributed as swdist
from . import ops
from .reference_api import Discriminator, Generator, get_traj_4d, predict_cv
from .scale import Scale


class SocialWaysTrainer:
    def __init__(self, data, batch_size=256, hidden_size=64, use_social=False, n_unrolling_steps=1,
                 lr_g=1e-4, lr_d=1e-3, device="cuda", weights=None, n_latent_codes=2,
                 use_info_loss=True, loss_info_w=0.5, world=None, cuda_graph=False, fused_adam=False):
        self.device = torch.device(device)
        self.batch_size, self.n_unrolling_steps = batch_size, n_unrolling_steps
        self.use_info_loss, self.loss_info_w, self.n_latent_codes = use_info_loss, loss_info_w, n_latent_codes
        # ---- train.py:89-124: data, 4/5 split over scenes, Scale ----
        obsv = np.array(data["obsvs"], dtype=np.float32, copy=True)
        pred = np.array(data["preds"], dtype=np.float32, copy=True)
        self.dataset_t = np.asarray(data["times"])
        the_batches = np.asarray(data["batches"])
        self.train_size = max(1, (len(the_batches) * 4) // 5)
        self.n_past, self.n_next = obsv.shape[1], pred.shape[1]
        self.n_train_samples = int(the_batches[self.train_size - 1][1])
        self.n_test_samples = obsv.shape[0] - self.n_train_samples
        # train.py:96-107 slices both tables BEFORE the n_test_samples == 0 substitution: a single-scene dataset is
        # trained on that scene and test() iterates over nothing (reports zeros)
        self.train_batches = the_batches[:self.train_size]
        self.test_batches = the_batches[self.train_size:]
        if self.n_test_samples == 0:
            self.n_test_samples = 1
            the_batches = np.array([the_batches[0], the_batches[0]])
        self.the_batches = the_batches
        self.scale = Scale()
        self.scale.max_x = max(np.max(obsv[:, :, 0]), np.max(pred[:, :, 0]))
        self.scale.min_x = min(np.min(obsv[:, :, 0]), np.min(pred[:, :, 0]))
        self.scale.max_y = max(np.max(obsv[:, :, 1]), np.max(pred[:, :, 1]))
        self.scale.min_y = min(np.min(obsv[:, :, 1]), np.min(pred[:, :, 1]))
        self.scale.calc_scale(keep_ratio=True)
        self.ss = self.scale.sx
        self.dataset_obsv = torch.from_numpy(self.scale.normalize(obsv)).to(self.device)
        self.dataset_pred = torch.from_numpy(self.scale.normalize(pred)).to(self.device)
        # ---- train.py:370-386: modules (same construction order => same init stream), optimisers ----
        self.generator = Generator(hidden_size, 1, 3, hidden_size, hidden_size // 2, use_social=use_social)
        self.D = Discriminator(self.n_next, hidden_size, n_latent_codes)
        if weights is not None:
            self.load_reference_weights(weights)
        self.generator.to(self.device)
        self.D.to(self.device)
        self.noise_len = hidden_size // 2
        self.cuda_graph = cuda_graph
        self._graphs = {}
        self.world_size, self.rank = swdist.world() if world is None else world
        self.fused_adam = fused_adam
        if fused_adam:
            # one flat buffer per optimiser; step() = ONE kernel that also sums the gradients over the ranks through NVLink
            # peer memory (fused_optim.FlatAdam / csrc/flat_adam.cu) -- no NCCL call inside the iteration
            from .fused_optim import FlatAdam
            group = None
            if self.world_size > 1:
                import torch.distributed as dist
                group = dist.group.WORLD
            self.predictor_optimizer = FlatAdam(self.generator.optimizer_parameters(), lr=lr_g, betas=(0.9, 0.999), group=group)
            self.D_optimizer = FlatAdam(self.D.parameters(), lr=lr_d, betas=(0.9, 0.999), group=group)
        else:
            extra = dict(capturable=True) if cuda_graph else {}
            self.predictor_optimizer = opt.Adam(self.generator.optimizer_parameters(), lr=lr_g, betas=(0.9, 0.999), **extra)
            self.D_optimizer = opt.Adam(self.D.parameters(), lr=lr_d, betas=(0.9, 0.999), **extra)
        self.mse_loss = nn.MSELoss()
        # weight-gradient contractions of train_native(): True = tcgen05 (tf32-split operands) when the problem is large enough to
        # pay for it, else the FFMA kernel; "force" = always tcgen05; False = always FFMA
        self.native_tensor_cores = True
        self.epoch = 1
        self.loss_log = []

    # reference module-global names
    encoder = property(lambda self: self.generator.encoder)
    feature_embedder = property(lambda self: self.generator.feature_embedder)
    attention = property(lambda self: self.generator.attention)
    decoder = property(lambda self: self.generator.decoder)

    def load_reference_weights(self, weights):
        """`weights`: {"encoder.embed.weight": ..., "D.classifier.0.bias": ...} (reference state_dict keys)."""
        as_t = lambda v: v if torch.is_tensor(v) else torch.from_numpy(np.asarray(v))
        self.generator.load_state_dict({k: as_t(v) for k, v in weights.items() if not k.startswith("D.")})
        self.D.load_state_dict({k[2:]: as_t(v) for k, v in weights.items() if k.startswith("D.")})

    def predict(self, obsv_p, noise, n_next, sub_batches=()):
        """predict() inside train(): the no-grad passes use the same FFMA decode kernel as the autograd pass,
        so the fake samples D sees and the ones G is trained on come from one arithmetic."""
        if not torch.is_grad_enabled():
            return self.generator.predict_k(obsv_p, noise.unsqueeze(0), n_next, sub_batches, precision="fp32")[0]
        return self.generator.predict(obsv_p, noise, n_next, sub_batches)

    def _zero_grad_D(self, set_to_none=False):
        if self.fused_adam:
            self.D_optimizer.zero_grad()          # one memset; the gradient views into the flat buffer stay
        elif set_to_none:
            self.D.zero_grad(set_to_none=True)
        else:
            self.D.zero_grad()

    # ------------------------------------------------------------------ train.py:439-560
    def train(self, verbose=True):
        tic = time.perf_counter()
        train_ADE, train_FDE = 0, 0
        batch_size_accum = 0
        sub_batches = []
        D, mse_loss, dev = self.D, self.mse_loss, self.device
        for ii, batch_i in enumerate(self.train_batches):
            batch_size_accum += batch_i[1] - batch_i[0]
            sub_batches.append(batch_i)
            if ii >= self.train_size - 1 or \
                    batch_size_accum + (self.the_batches[ii + 1][1] - self.the_batches[ii + 1][0]) > self.batch_size:
                obsv = self.dataset_obsv[sub_batches[0][0]:sub_batches[-1][1]]
                pred = self.dataset_pred[sub_batches[0][0]:sub_batches[-1][1]]
                sub_batches = np.asarray(sub_batches) - sub_batches[0][0]
                bs = int(batch_size_accum)
                obsv_4d, pred_4d = get_traj_4d(obsv, pred)
                zeros = (torch.zeros(bs, 1) + np.random.uniform(0, 0.1)).to(dev)          # :471
                ones = (torch.ones(bs, 1) * np.random.uniform(0.9, 1.0)).to(dev)           # :472
                noise = torch.rand(bs, self.noise_len).to(dev)                             # :473 (CPU RNG)
                if self.world_size > 1:
                    # labels / noise are drawn for the GLOBAL batch on every rank (identical
                    # RNG streams), then each rank keeps the rows of its own block of scenes
                    lo, hi, sub_batches = swdist.shard_scenes(sub_batches, self.world_size, self.rank)
                    obsv, pred, obsv_4d, pred_4d = obsv[lo:hi], pred[lo:hi], obsv_4d[lo:hi], pred_4d[lo:hi]
                    zeros, ones, noise = zeros[lo:hi], ones[lo:hi], noise[lo:hi]
                    mse_loss = lambda a, b, _bs=bs: swdist.global_mse(a, b, _bs * max(1, a.numel() // max(1, a.shape[0])))
                have_rows = obsv.shape[0] > 0
                backup = None
                # ============== Train Discriminator ================ :476-499
                for u in range(self.n_unrolling_steps + 1):
                    self._zero_grad_D()
                    if have_rows:
                        with torch.no_grad():
                            pred_hat_4d = self.predict(obsv, noise, self.n_next, sub_batches)
                        obsv_h = D.encode_obsv(obsv_4d)
                        fake_labels, code_hat = D.heads(obsv_h, pred_hat_4d)
                        d_loss_fake = mse_loss(fake_labels, zeros)
                        d_loss_info = mse_loss(code_hat.squeeze(1) if code_hat.shape[1] == 1 else code_hat,
                                               noise[:, :self.n_latent_codes])
                        real_labels, code_hat = D.heads(obsv_h, pred_4d)
                        d_loss_real = mse_loss(real_labels, ones)
                        d_loss = d_loss_fake + d_loss_real
                        if self.use_info_loss:
                            d_loss = d_loss + self.loss_info_w * d_loss_info
                        d_loss.backward()
                    if not self.fused_adam:
                        swdist.allreduce_grads(D.parameters(), self.world_size)
                    self.D_optimizer.step()
                    if u == 0 and self.n_unrolling_steps > 0:
                        backup = copy.deepcopy(D)
                # =============== Train Generator ================= :501-543
                self._zero_grad_D()
                self.predictor_optimizer.zero_grad()
                if have_rows:
                    pred_hat_4d = self.predict(obsv, noise, self.n_next, sub_batches)
                    with torch.no_grad():          # D's observation code does not depend on the generator
                        obsv_h = D.encode_obsv(obsv_4d)
                    gen_labels, code_hat = D.heads(obsv_h, pred_hat_4d)
                    g_loss_fooling = mse_loss(gen_labels, ones)
                    g_loss_info = mse_loss(code_hat.squeeze(1) if code_hat.shape[1] == 1 else code_hat,
                                           noise[:, :self.n_latent_codes])
                    g_loss = g_loss_fooling
                    if self.use_info_loss:
                        g_loss = g_loss + self.loss_info_w * g_loss_info
                    g_loss.backward()
                if not self.fused_adam:
                    swdist.allreduce_grads(list(self.generator.optimizer_parameters()), self.world_size)
                self.predictor_optimizer.step()
                if self.n_unrolling_steps > 0:
                    D.load(backup)
                    del backup
                if have_rows:
                    with torch.no_grad():                                                  # :546-551
                        err_all = torch.pow((pred_hat_4d[:, :, :2] - pred) / self.ss, 2)
                        err_all = err_all.sum(dim=2).sqrt()
                        e = err_all.sum().item() / self.n_next
                        train_ADE += e
                        train_FDE += err_all[:, -1].sum().item()
                if not have_rows:
                    d_loss = d_loss_fake = d_loss_real = d_loss_info = g_loss_fooling = g_loss_info = torch.zeros(())
                self.loss_log.append(dict(d_loss=d_loss.item(), d_fake=d_loss_fake.item(), d_real=d_loss_real.item(),
                                          d_info=d_loss_info.item(), g_fool=g_loss_fooling.item(),
                                          g_info=g_loss_info.item()))
                batch_size_accum = 0
                sub_batches = []
        train_ADE, train_FDE = swdist.allreduce_scalars([train_ADE, train_FDE], dev, self.world_size)
        train_ADE /= self.n_train_samples
        train_FDE /= self.n_train_samples
        toc = time.perf_counter()
        if verbose and self.rank == 0:
            print(" Epc=%4d, Train ADE,FDE = (%.3f, %.3f) | time = %.1f" % (self.epoch, train_ADE, train_FDE, toc - tic))
        return train_ADE, train_FDE

    # ------------------------------------------------------------------ CUDA-graph variant of train()
    def _graph_body(self, st):
        """One iteration of train() (train.py:470-551) on static tensors, free of host synchronisation so that it
        can be captured into a CUDA graph: no .item(), no host RNG, in-place D backup/rollback instead of
        copy.deepcopy / .data rebinding (same values)."""
        D, nl = self.D, self.n_latent_codes
        obsv, pred, noise, zeros, ones, scenes = st["obsv"], st["pred"], st["noise"], st["zeros"], st["ones"], st["scenes"]
        if self.world_size > 1:       # this rank's share of nn.MSELoss over the GLOBAL mini-batch
            mse_loss = lambda a, b, _bs=st["global_bs"]: swdist.global_mse(a, b, _bs * max(1, a.numel() // max(1, a.shape[0])))
        else:
            mse_loss = self.mse_loss
        obsv_4d, pred_4d = get_traj_4d(obsv, pred)
        lin = [p for m in D.modules() if isinstance(m, nn.Linear) for p in (m.weight, m.bias)]
        for u in range(self.n_unrolling_steps + 1):
            self._zero_grad_D(set_to_none=True)
            with torch.no_grad():
                pred_hat_4d = self.predict(obsv, noise, self.n_next, scenes)
            obsv_h = D.encode_obsv(obsv_4d)
            fake_labels, code_hat = D.heads(obsv_h, pred_hat_4d)
            d_fake = mse_loss(fake_labels, zeros)
            d_info = mse_loss(code_hat, noise[:, :nl])
            real_labels, _ = D.heads(obsv_h, pred_4d)
            d_real = mse_loss(real_labels, ones)
            d_loss = d_fake + d_real + (self.loss_info_w * d_info if self.use_info_loss else 0.0)
            d_loss.backward()
            if not self.fused_adam:
                swdist.allreduce_grads(D.parameters(), self.world_size)
            self.D_optimizer.step()
            if u == 0 and self.n_unrolling_steps > 0:
                with torch.no_grad():
                    for b, p in zip(st["backup"], lin):
                        b.copy_(p)
        self._zero_grad_D(set_to_none=True)
        self.predictor_optimizer.zero_grad(set_to_none=True)
        pred_hat_4d = self.predict(obsv, noise, self.n_next, scenes)
        with torch.no_grad():
            obsv_h = D.encode_obsv(obsv_4d)
        gen_labels, code_hat = D.heads(obsv_h, pred_hat_4d)
        g_fool = mse_loss(gen_labels, ones)
        g_info = mse_loss(code_hat, noise[:, :nl])
        g_loss = g_fool + (self.loss_info_w * g_info if self.use_info_loss else 0.0)
        g_loss.backward()
        if not self.fused_adam:
            swdist.allreduce_grads(list(self.generator.optimizer_parameters()), self.world_size)
        self.predictor_optimizer.step()
        with torch.no_grad():
            if self.n_unrolling_steps > 0:                                   # D.load(backup): Linear layers only
                for b, p in zip(st["backup"], lin):
                    p.copy_(b)
            err_all = torch.pow((pred_hat_4d[:, :, :2] - pred) / self.ss, 2).sum(dim=2).sqrt()
            st["stats"].copy_(torch.stack([err_all.sum() / self.n_next, err_all[:, -1].sum(), d_loss.detach(), d_fake.detach(),
                                           d_real.detach(), d_info.detach(), g_fool.detach(), g_info.detach()]))

    def train_graphed(self, verbose=True):
        """train() with every mini-batch shape captured once into a CUDA graph and replayed (removes the
        ~10 ms/iteration of Python + launch overhead that bounds small batches).  Needs optimisers built with
        capturable=True (constructor flag cuda_graph=True) and world_size 1.  The first occurrence of a batch
        shape runs eagerly (it also warms up the lazily-initialised optimiser state), the second is captured."""
        if not self.cuda_graph or (self.world_size != 1 and not self.fused_adam):
            # capturing the NCCL all-reduces of the sharded step was tried (torch 2.11 / NCCL 2.28) and hung in capture;
            # the sharded step is captured with fused_adam=True instead: its optimiser kernel does the gradient exchange
            # itself over NVLink peer memory (csrc/flat_adam.cu), so the graph holds plain kernels only
            raise RuntimeError("train_graphed() needs cuda_graph=True, and fused_adam=True for multi-GPU runs")
        tic = time.perf_counter()
        dev = self.device
        stats_acc = torch.zeros(8, device=dev, dtype=torch.float64)
        n_iter = 0
        group, count = [], 0
        for ii, batch_i in enumerate(self.train_batches):
            count += int(batch_i[1] - batch_i[0])
            group.append(batch_i)
            if not (ii >= self.train_size - 1 or
                    count + (self.the_batches[ii + 1][1] - self.the_batches[ii + 1][0]) > self.batch_size):
                continue
            lo, hi = int(group[0][0]), int(group[-1][1])
            sub = np.asarray(group) - lo
            global_bs, g_lo = hi - lo, lo
            if self.world_size > 1:       # this rank's contiguous block of scenes (the NCCL all-reduces are captured too)
                s_lo, s_hi, sub = swdist.shard_scenes(sub, self.world_size, self.rank)
                if s_hi <= s_lo:
                    raise RuntimeError("train_graphed(): a rank received no scene of this mini-batch; use train()")
                lo, hi = g_lo + s_lo, g_lo + s_hi
            key = (global_bs, hi - lo, lo - g_lo, sub.tobytes())
            ent = self._graphs.get(key)
            if ent is None:
                bs = hi - lo
                lin_n = [p for m in self.D.modules() if isinstance(m, nn.Linear) for p in (m.weight, m.bias)]
                st = dict(obsv=torch.empty(bs, self.n_past, 2, device=dev), pred=torch.empty(bs, self.n_next, 2, device=dev),
                          noise=torch.empty(bs, self.noise_len, device=dev), zeros=torch.empty(bs, 1, device=dev),
                          ones=torch.empty(bs, 1, device=dev), stats=torch.zeros(8, device=dev),
                          backup=[torch.empty_like(p) for p in lin_n], global_bs=global_bs,
                          scenes=self.generator.scene_index(sub, bs, dev))
                _ = st["scenes"].pair_offsets                                 # build every lazy device index up front
                ent = self._graphs[key] = dict(st=st, graph=None, seen=0)
            st = ent["st"]
            st["obsv"].copy_(self.dataset_obsv[lo:hi])
            st["pred"].copy_(self.dataset_pred[lo:hi])
            st["zeros"].fill_(float(np.random.uniform(0, 0.1)))               # train.py:471
            st["ones"].fill_(float(np.random.uniform(0.9, 1.0)))              # train.py:472
            st["noise"].copy_(torch.rand(global_bs, self.noise_len)[lo - g_lo:hi - g_lo])   # train.py:473 (CPU RNG, global batch)
            if ent["graph"] is None and ent["seen"] >= 1:
                g = torch.cuda.CUDAGraph()
                torch.cuda.synchronize()
                with torch.cuda.graph(g):
                    self._graph_body(st)
                ent["graph"] = g
            if ent["graph"] is not None:
                ent["graph"].replay()
            else:
                self._graph_body(st)
            ent["seen"] += 1
            stats_acc += st["stats"]
            n_iter += 1
            group, count = [], 0
        # graph.replay() rewrites the parameters through raw pointers without advancing their version counters
        # (neither capturable Adam nor FlatAdam's python-side increment_version runs on replay): the packed-weight cache
        # Generator.packs() keys on (data_ptr, _version) would go stale, so test() would evaluate old weights
        self.generator.invalidate_packs()
        if self.world_size > 1:
            import torch.distributed as dist
            dist.all_reduce(stats_acc, op=dist.ReduceOp.SUM)
        vals = stats_acc.tolist()
        train_ADE, train_FDE = vals[0] / self.n_train_samples, vals[1] / self.n_train_samples
        self.last_epoch_mean_losses = dict(zip(("d_loss", "d_fake", "d_real", "d_info", "g_fool", "g_info"),
                                               [v / max(1, n_iter) for v in vals[2:]]))
        toc = time.perf_counter()
        if verbose and self.rank == 0:
            print(" Epc=%4d, Train ADE,FDE = (%.3f, %.3f) | time = %.1f" % (self.epoch, train_ADE, train_FDE, toc - tic))
        return train_ADE, train_FDE

    # ------------------------------------------------------------------ native iteration (native_step.py)
    def _minibatches(self):
        """Mini-batch grouping of train() (train.py:446-461): yields (global lo, global hi, scene table rebased to lo)."""
        group, count = [], 0
        for ii, batch_i in enumerate(self.train_batches):
            count += int(batch_i[1] - batch_i[0])
            group.append(batch_i)
            if ii >= self.train_size - 1 or count + (self.the_batches[ii + 1][1] - self.the_batches[ii + 1][0]) > self.batch_size:
                lo, hi = int(group[0][0]), int(group[-1][1])
                yield lo, hi, np.asarray(group) - lo
                group, count = [], 0

    def _native_empty_iteration(self):
        """This rank holds no scene of the mini-batch: take part in the optimiser steps (their kernels carry the gradient
        exchange) with zero gradients, including the Linear-only rollback of the unrolled discriminator."""
        d_lin, backup = self._native_packs.d_linear, None
        for u in range(self.n_unrolling_steps + 1):
            self.D_optimizer.zero_grad()
            self.D_optimizer.step()
            if u == 0 and self.n_unrolling_steps > 0:
                backup = d_lin.clone()
        self.predictor_optimizer.zero_grad()
        self.predictor_optimizer.step()
        if backup is not None:
            d_lin.copy_(backup)

    def train_native(self, verbose=True, use_graph=True, log_losses=False, device_noise_seed=None):
        """train() with the iteration executed by native_step.NativeStep: ~30 launches of this library's kernels per
        iteration (no autograd, no cuBLAS, no ATen reductions), one CUDA graph per mini-batch shape.  Needs
        fused_adam=True (flat parameter / gradient buffers).  Same RNG consumption and quirks as train().
        device_noise_seed=s: the per-iteration latent noise (train.py:473: torch.rand on the CPU, ~1 ms per 4 096 agents --
        more than the GPU needs for the whole iteration) is drawn on the device instead (Philox4x32-10, sw_noise_uniform): same
        distribution, a different stream, identical for any number of ranks.  The two label scalars stay on numpy's RNG."""
        if not self.fused_adam:
            raise RuntimeError("train_native() needs fused_adam=True (flat parameter / gradient buffers)")
        from .native_step import NativePacks, NativeStep
        tic = time.perf_counter()
        dev = self.device
        if not hasattr(self, "_native_packs"):
            self._native_packs, self._native_steps = NativePacks(self), {}
            self._pin = {}
        stats_acc = torch.zeros(8, device=dev, dtype=torch.float64)
        n_iter = 0
        if not hasattr(self, "_native_plan"):
            # the mini-batch grouping, this rank's shard of every mini-batch and its launch sequence are the same every
            # epoch: walk the scene table once (the reference re-walks it in Python every epoch -- ~1 us per scene, more than
            # the GPU needs for the whole epoch at large batches)
            self._native_plan = []
            for g_lo, g_hi, sub in self._minibatches():
                lo, hi = g_lo, g_hi
                if self.world_size > 1:       # this rank's contiguous block of scenes
                    s_lo, s_hi, sub = swdist.shard_scenes(sub, self.world_size, self.rank)
                    lo, hi = g_lo + s_lo, g_lo + s_hi
                ent = None
                if hi > lo:
                    key = (g_hi - g_lo, hi - lo, lo - g_lo, sub.tobytes())
                    ent = self._native_steps.get(key)
                    if ent is None:
                        step = NativeStep(self, self._native_packs, hi - lo, self.generator.scene_index(sub, hi - lo, dev), g_hi - g_lo)
                        ent = self._native_steps[key] = dict(step=step, graph=None, seen=0)
                self._native_plan.append((g_lo, g_hi, lo, hi, ent))
        trace = [] if os.environ.get("SW_TRACE_TRAIN") else None    # per-iteration (host ms, GPU ms) of this rank, printed per epoch
        for g_lo, g_hi, lo, hi, ent in self._native_plan:
            if trace is not None:
                e0 = torch.cuda.Event(enable_timing=True); e0.record()
                trace.append([time.perf_counter(), e0, None, hi - lo])
            self._native_iteration = getattr(self, "_native_iteration", 0) + 1
            global_bs = g_hi - g_lo
            # train.py:471-473: two numpy scalars, then the noise of the GLOBAL mini-batch from torch's CPU generator
            t01 = (float(np.random.uniform(0, 0.1)), float(np.random.uniform(0.9, 1.0)))
            pin = None
            if device_noise_seed is None:
                pins = self._pin.get(global_bs)
                if pins is None:
                    pins = self._pin[global_bs] = [dict(noise=torch.empty(global_bs, self.noise_len).pin_memory(),
                                                        ev=torch.cuda.Event()) for _ in range(2)]
                pin = pins[n_iter & 1]
                pin["ev"].synchronize()                       # the upload that last read this pinned buffer has finished
                torch.rand(global_bs, self.noise_len, out=pin["noise"])     # same stream as torch.rand(bs, noise_len)
            if ent is None:               # more ranks than scenes: contribute zero gradients to the three all-reduces
                self._native_empty_iteration()
                n_iter += 1
                continue
            step = ent["step"]
            step.obsv.copy_(self.dataset_obsv[lo:hi])
            step.pred.copy_(self.dataset_pred[lo:hi])
            if pin is not None:
                step.noise.copy_(pin["noise"][lo - g_lo:hi - g_lo], non_blocking=True)
                pin["ev"].record()
            else:                                             # rows [lo - g_lo, hi - g_lo) of the iteration's logical [global_bs, 32] noise
                ops.noise_uniform(None, dev, device_noise_seed, offset=self._native_iteration, out=step.noise,
                                  first_element=(lo - g_lo) * self.noise_len)
            step.set_targets(*t01)
            if use_graph and ent["graph"] is None and ent["seen"] >= 1:
                g = torch.cuda.CUDAGraph()
                torch.cuda.synchronize()
                with torch.cuda.graph(g):
                    step.run()
                ent["graph"] = g
            if ent["graph"] is not None:
                ent["graph"].replay()
            else:
                step.run()
            ent["seen"] += 1
            stats_acc += step.stats
            if trace is not None:
                e1 = torch.cuda.Event(enable_timing=True); e1.record()
                trace[-1][2] = e1
                trace[-1][0] = (time.perf_counter() - trace[-1][0]) * 1e3
            if log_losses:                                    # tests: one host read per iteration
                v = step.stats.tolist()
                self.loss_log.append(dict(d_loss=v[2], d_fake=v[3], d_real=v[4], d_info=v[5], g_fool=v[6], g_info=v[7]))
            n_iter += 1
        self.generator.invalidate_packs()                     # parameters were rewritten through raw pointers
        if self.world_size > 1:
            import torch.distributed as dist
            dist.all_reduce(stats_acc, op=dist.ReduceOp.SUM)
        vals = stats_acc.tolist()
        for o in (self.predictor_optimizer, self.D_optimizer):
            o.check_status()
        if trace is not None:
            torch.cuda.synchronize()
            print(f"[rank {self.rank}] train_native epoch: total {1e3 * (time.perf_counter() - tic):.2f} ms; per iteration (rows, host ms, gpu ms): "
                  + ", ".join(f"({r}, {h:.2f}, {a.elapsed_time(b):.2f})" for h, a, b, r in trace if b is not None), flush=True)
        train_ADE, train_FDE = vals[0] / self.n_train_samples, vals[1] / self.n_train_samples
        self.last_epoch_mean_losses = dict(zip(("d_loss", "d_fake", "d_real", "d_info", "g_fool", "g_info"),
                                               [v / max(1, n_iter) for v in vals[2:]]))
        toc = time.perf_counter()
        if verbose and self.rank == 0:
            print(" Epc=%4d, Train ADE,FDE = (%.3f, %.3f) | time = %.1f" % (self.epoch, train_ADE, train_FDE, toc - tic))
        return train_ADE, train_FDE

    # ------------------------------------------------------------------ train.py:563-616
    def _test_noise(self, scenes, n_gen_samples):
        """Noise of test() for `scenes` ([start, end) rows of the dataset), drawn from torch's CPU RNG in the reference's
        order -- per scene, K consecutive torch.rand(bs, noise_len) calls (train.py:583-584).  Consecutive torch.rand calls
        read consecutive values of ONE stream, so a single torch.rand(total) followed by a per-scene reshape yields
        bit-identical values (tests/test_entry_points_cpu.py pins this) without scenes x K host calls.
        Returns a pinned [K, rows, noise_len] tensor, rows relative to scenes[0][0]."""
        K, nl = n_gen_samples, self.noise_len
        sizes = np.asarray([int(b[1] - b[0]) for b in scenes], dtype=np.int64)
        rows = int(sizes.sum())
        flat = torch.rand(K * rows * nl)
        noise = torch.empty(K, rows, nl).pin_memory() if self.device.type == "cuda" else torch.empty(K, rows, nl)
        if len(sizes) and (sizes == sizes[0]).all():           # equal scenes: one permute
            bs = int(sizes[0])
            noise.copy_(flat.view(len(sizes), K, bs, nl).permute(1, 0, 2, 3).reshape(K, rows, nl))
        else:
            off, a = 0, 0
            for bs in sizes.tolist():
                noise[:, a:a + bs] = flat[off:off + K * bs * nl].view(K, bs, nl)
                off += K * bs * nl
                a += bs
        return noise

    def skip_test_rng(self, n_gen_samples=20, linear=False, write_to_file=None, just_one=False):
        """Consume exactly the CPU random numbers test() would (ranks that do not evaluate call this, so every rank's
        noise stream stays identical to rank 0's and to a single-GPU run)."""
        if linear and not write_to_file:
            return
        scenes = self.test_batches[:1] if just_one else self.test_batches
        rows = int(sum(int(b[1] - b[0]) for b in scenes))
        if rows:
            torch.rand(n_gen_samples * rows * self.noise_len)

    def test(self, n_gen_samples=20, linear=False, write_to_file=None, just_one=False, verbose=True):
        """test() of the reference with ALL test scenes decoded by one predict_k call (the scenes are contiguous rows of the
        dataset; pooling never crosses `sub_batches`), one best-of-K metrics kernel and ONE device->host read per call.  The
        reference's loop order survives where it is observable: the noise stream (see _test_noise) and the dump files."""
        scenes = self.test_batches[:1] if just_one else self.test_batches
        sums = np.zeros(4)
        if len(scenes):
            lo, hi = int(scenes[0][0]), int(scenes[-1][1])
            obsv, pred = self.dataset_obsv[lo:hi], self.dataset_pred[lo:hi]
            with torch.no_grad():
                linear_preds = predict_cv(obsv, self.n_next)
                if linear and not write_to_file:
                    all_preds = torch.cat([linear_preds, torch.zeros_like(linear_preds)], dim=2).unsqueeze(0)
                else:
                    noise = self._test_noise(scenes, n_gen_samples).to(self.device, non_blocking=True)
                    sub = np.asarray(scenes, dtype=np.int64) - lo
                    all_preds = self.generator.predict_k(obsv, noise, self.n_next, sub)      # [K, rows, T, 4]
                m = ops.bestofk_metrics(all_preds.contiguous(), pred.contiguous(), self.ss)   # :587,602-607, per agent
                sums = m.double().sum(dim=0).cpu().numpy()                                    # the call's one sync
                if not (linear and not write_to_file) and self.generator.fp16_overflowed():
                    # weights / states left fp16's exponent range: the fp16-split tensor-core kernels are not valid for this
                    # checkpoint -- same call on the fp32 FFMA kernels (no silent inf / NaN)
                    print("socialways_b200: operands outside fp16's range, test() falls back to the fp32 kernels")
                    all_preds = self.generator.predict_k(obsv, noise, self.n_next, sub, precision="fp32")
                    m = ops.bestofk_metrics(all_preds.contiguous(), pred.contiguous(), self.ss)
                    sums = m.double().sum(dim=0).cpu().numpy()
                if write_to_file:
                    ours = self.scale.denormalize(all_preds[:, :, :, :2].cpu().numpy())
                    o_np = self.scale.denormalize(obsv[:, :, :2].cpu().numpy())
                    p_np = self.scale.denormalize(pred[:, :, :2].cpu().numpy())
                    l_np = self.scale.denormalize(linear_preds[:, :, :2].cpu().numpy())
                    for batch_i in scenes:
                        a, b = int(batch_i[0]) - lo, int(batch_i[1]) - lo
                        current_t = self.dataset_t[batch_i[0]]
                        file_name = os.path.join(write_to_file, str(self.epoch) + '-' + str(current_t) + '.npz')
                        print('saving to ', file_name)
                        np.savez(file_name, timestamp=current_t, obsvs=o_np[a:b], preds_our=ours[:, a:b],
                                 preds_gtt=p_np[a:b], preds_lnr=l_np[a:b])
        ade_avg_12, fde_avg_12, ade_min_12, fde_min_12 = (float(v) / self.n_test_samples for v in sums)
        if verbose:
            print('Avg ADE,FDE (12)= (%.3f, %.3f) | Min(20) ADE,FDE (12)= (%.3f, %.3f)'
                  % (ade_avg_12, fde_avg_12, ade_min_12, fde_min_12))
        return dict(ade_avg=ade_avg_12, fde_avg=fde_avg_12, ade_min=ade_min_12, fde_min=fde_min_12)

    # ------------------------------------------------------------------ train.py:622-663
    def state(self):
        return {'epoch': self.epoch,
                'attentioner_dict': self.attention.state_dict(),
                'feature_embedder_dict': self.feature_embedder.state_dict(),
                'encoder_dict': self.encoder.state_dict(),
                'decoder_dict': self.decoder.state_dict(),
                'pred_optimizer': self.predictor_optimizer.state_dict(),
                'D_dict': self.D.state_dict(),
                'D_optimizer': self.D_optimizer.state_dict()}

    def load_state(self, checkpoint):
        self.attention.load_state_dict(checkpoint['attentioner_dict'])
        self.feature_embedder.load_state_dict(checkpoint['feature_embedder_dict'])
        self.encoder.load_state_dict(checkpoint['encoder_dict'])
        self.decoder.load_state_dict(checkpoint['decoder_dict'])
        self.predictor_optimizer.load_state_dict(checkpoint['pred_optimizer'])
        self.D.load_state_dict(checkpoint['D_dict'])
        self.D_optimizer.load_state_dict(checkpoint['D_optimizer'])
        if self.cuda_graph and not self.fused_adam:
            # Optimizer.load_state_dict() takes `capturable` (and the host-side `step`) from the checkpoint; reference and
            # eager checkpoints store capturable=False, which fails torch's check at graph capture.  Restore what this
            # trainer was built with: capturable groups, step counters as fp32 device tensors.
            for o in (self.predictor_optimizer, self.D_optimizer):
                for g in o.param_groups:
                    g['capturable'] = True
                for st in o.state.values():
                    if 'step' in st:
                        st['step'] = torch.as_tensor(st['step'], dtype=torch.float32).to(self.device)
        self.generator.invalidate_packs()
        return checkpoint['epoch'] + 1

    def reference_weights(self):
        out = {k: v.detach().clone() for k, v in self.generator.state_dict().items()}
        out.update({"D." + k: v.detach().clone() for k, v in self.D.state_dict().items()})
        return out
