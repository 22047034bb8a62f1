#!/usr/bin/env python
"""Headline benchmark: predicted trajectories / second of the Social Ways K-sample inference path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[1]): ETH-shaped synthetic scenes, 8 agents per scene, obs 8 / pred 12,
K = 20 samples per agent, hidden 64, use_social = True, fp32.  One "step" = one pass of the hot path
(observation encode -> pairwise social attention pooling -> K-sample 12-step decode -> best-of-K
ADE/FDE) over one batch of `--scenes` scenes per GPU.  One predicted trajectory = one (agent, sample)
pair x 12 steps.

`value`   : whole-job trajectories/s with the inputs already resident in HBM (device-timed, CUDA events,
            max over ranks).
`e2e`     : same metric through the public API with HOST (pinned) inputs: per step the observations,
            ground truth and noise are copied host->device and the per-agent metrics device->host,
            all inside the timed region (copies of step i+1 overlap compute of step i on a second stream).
`roofline`: the decode kernel (dominant), algorithmic FLOPs per trajectory over its
            CUDA-event time on the launching stream, against the measured bf16 tensor peak
            (MEASURED_PEAKS.json) -- the kernel is compute-bound.
`cpu_baseline`: the CPU oracle port of the reference's test() loop (one scene at a time, K serial
            predict() calls, per-agent attention loop) on the host cores, bounded sample.
`--impl reference`: the same CPU arm as its own JSON line (the reference is Python over torch and
            cannot travel to the GPU box; the oracle port restates it).
`train_step`: the OTHER half of north_star -- the GAN training iteration (train(), reference train.py:439-560) on
            BASELINE configs[3] (toy set scaled to 65 532 trajectories, 6-agent scenes, obs 2 / pred 2, use_social, unroll 1):
            scenes of every mini-batch sharded over the N ranks, the iteration = ~30 launches of this library's kernels
            replayed from one CUDA graph (socialways_b200/native_step.py), the gradient all-reduce done by peer loads over
            NVLink inside the Adam kernel (csrc/flat_adam.cu).  Reported per global batch size: ms per iteration (device clock,
            max over ranks), agents/s, launches per iteration, and `train_parity_max_abs` = the largest weight difference
            between the N-rank native run and a single-GPU run of the autograd trainer from the same seeds.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

A_PER_SCENE, N_PAST, N_NEXT, K_SAMPLES = 8, 8, 12, 20
FLOPS_PER_TRAJ = 1_726_848          # 12 x DecoderFC (83 360) + 11 x encoder step (66 048)
FLOPS_PER_TRAJ_EXECUTED = 12 * 2 * (64 * 160 + 160 * 80 + 80 * 2) + 11 * 2 * 68 * 256 + 2 * 96 * 160
# dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of each decode kernel, divided by the
# trajectories of that launch (655 360): profiles/r2_decode_pair.txt, r1_decode_tcx.txt, r1_decode_fp32_ffma.txt.
# Algorithmic: 128 B noise in + 192 B (p, v) x 12 out = 320 B per trajectory.
NCU_DRAM_BYTES_PER_TRAJ = {"fp16x2": (158.654464e6 + 122.619136e6) / 655360, "fp16x2s": (115.264768e6 + 92.413184e6) / 655360,
                           "fp32": (114.652672e6 + 91.719424e6) / 655360, "bf16": None, "bf16p": None}


KERNEL_OF = {"fp32": "decode_fwd_kernel", "fp16x2": "decode_fwd_pair_kernel", "fp16x2s": "decode_fwd_tcx_kernel",
             "bf16": "decode_fwd_tc_kernel", "bf16p": "decode_fwd_pair_bf16_kernel"}
DTYPE_OF = {"fp32": "f32", "fp16x2": "f32 (fp16 hi/lo split operands on tcgen05, fp32 accumulate)",
            "fp16x2s": "f32 (fp16 hi/lo split operands on tcgen05, fp32 accumulate)", "bf16": "bf16", "bf16p": "bf16"}
NOTE_OF = {"fp32": "fp32 FFMA path; CUDA-core fp32 peak is ~74 TFLOP/s",
           "fp16x2": "tcgen05: 3 MMAs per product on fp16 hi/lo split operands, fp32 accumulate in TMEM; matches the fp32 "
                     "oracle to ~1e-6 (tests/test_gpu_tensorcore.py), i.e. inside the 1e-4 ADE/FDE parity bar; executed "
                     "tensor FLOPs are 3x the algorithmic ones",
           "fp16x2s": "the same arithmetic as fp16x2 on the one-tile-per-SM kernel (decode_fwd_tcx.cu)",
           "bf16": "tcgen05 bf16 operands / fp32 accumulate in TMEM (fast mode, outside the 1e-4 parity bar)",
           "bf16p": "the CTA-pair kernel on single bf16 operands (one MMA per product; fast mode, outside the 1e-4 parity bar)"}


def make_scenes(n_scenes, seed):
    from golden_data import synthetic_scenes
    return synthetic_scenes([A_PER_SCENE] * n_scenes, n_past=N_PAST, n_next=N_NEXT, seed=seed)


def normalised(data):
    """Scale-normalised observations / ground truth with the package's own Scale (train.py:113-121)."""
    from socialways_b200.scale import Scale
    o, p = np.array(data["obsvs"], dtype=np.float32), np.array(data["preds"], dtype=np.float32)
    sc = Scale()
    sc.max_x, sc.min_x = max(o[..., 0].max(), p[..., 0].max()), min(o[..., 0].min(), p[..., 0].min())
    sc.max_y, sc.min_y = max(o[..., 1].max(), p[..., 1].max()), min(o[..., 1].min(), p[..., 1].min())
    sc.calc_scale(keep_ratio=True)
    return sc.normalize(o), sc.normalize(p), float(sc.sx)


def bind_to_gpu_numa(index):
    """Pin this rank's host threads (and so its pinned staging buffers, first touch) to the CPUs local to its GPU."""
    try:
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(index))
        return sorted(os.sched_getaffinity(0))
    except Exception:                                   # best effort: missing NVML / no permission changes nothing else
        return None


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm_gbs=p["hbm_gbs"], bf16=p["bf16_tflops"], bf16_sustained=p["bf16_tflops_sustained"], src="measured")
    return dict(hbm_gbs=6650.0, bf16=1590.0, bf16_sustained=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.25)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 6 and r[2 + i] == "Active" for r in self.rows)]
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=reasons, samples=len(sm))


def cpu_arm(n_scenes, repeats=1):
    """CPU oracle port of test() (train.py:563-616): per scene, K serial predict() calls with the
    reference's per-agent attention loop, all host threads.  Returns (traj/s, threads, seconds)."""
    from oracle import socialways_oracle as so
    torch.set_num_threads(os.cpu_count() or 1)
    P = so.init_weights(seed=0)
    data = make_scenes(n_scenes, seed=1234)
    sc = so.IsoScale(data["obsvs"], data["preds"])
    obsv = torch.from_numpy(sc.normalize(data["obsvs"]))
    pred = torch.from_numpy(sc.normalize(data["preds"]))
    best = None
    for _ in range(repeats):
        torch.manual_seed(0)
        t0 = time.perf_counter()
        with torch.no_grad():
            for a, b in data["batches"]:
                errs = []
                for k in range(K_SAMPLES):
                    noise = torch.rand(b - a, 32)
                    hat = so.predict(P, obsv[a:b], noise, N_NEXT, None, use_social=True, pool="loop")
                    errs.append((((hat[:, :, :2] - pred[a:b]) / sc.sx) ** 2).sum(dim=2).sqrt())
                e = torch.stack(errs)
                _ = (e.mean(2).min(0)[0].sum().item(), e[:, :, -1].min(0)[0].sum().item())
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    traj = n_scenes * A_PER_SCENE * K_SAMPLES
    return traj / best, torch.get_num_threads(), best


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    for _ in range(args.warmup):
        cpu_arm(max(1, args.cpu_scenes // 8))
    t_total, traj_total = 0.0, 0
    for _ in range(args.steps):
        v, cores, dt = cpu_arm(args.cpu_scenes)
        t_total += dt
        traj_total += args.cpu_scenes * A_PER_SCENE * K_SAMPLES
    value = traj_total / t_total
    sample = f"{args.cpu_scenes} scenes x {A_PER_SCENE} agents x K={K_SAMPLES} per step"
    print(json.dumps({
        "impl": "reference", "metric": "predicted_trajectories_per_sec", "value": value, "unit": "traj/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, args.cpu_scenes),
        "cpu_baseline": {"value": value, "unit": "traj/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "traj/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def workload_config(args, scenes):
    shape = {8: "ETH-shaped", 32: "Zara-shaped", 256: "dense-crowd"}.get(A_PER_SCENE, "synthetic")
    return {"workload": f"{shape} synthetic K-sample inference: {scenes} scenes/GPU/step x {A_PER_SCENE} agents, "
                        f"obs {N_PAST} pred {N_NEXT}, K={K_SAMPLES}, hidden 64, use_social=True",
            "scenes_per_gpu_per_step": scenes, "agents_per_scene": A_PER_SCENE, "K": K_SAMPLES,
            "obs_len": N_PAST, "pred_len": N_NEXT, "parallelism": f"scenes sharded x{args.gpus}, no collective",
            "l2_policy": "per-step inputs+outputs (noise 128 B + pred 192 B per trajectory) exceed the 126 MB L2"}


# ------------------------------------------------------------------------------------------------------------------
# train_step: the GAN training iteration (reference train(), train.py:439-560) on BASELINE configs[3]
# ------------------------------------------------------------------------------------------------------------------
TOY_TO, TOY_TP, TOY_A = 2, 2, 6
# algorithmic FLOPs per agent and iteration, SURVEY.md 8(d): 5 x predict-forward equivalents (3 fwd + 1 bwd ~ 2 fwd) +
# 15 x discriminator-forward equivalents (5 fwd + 5 bwd), at the toy shapes (obs 2 / pred 2, 6-agent scenes)
_PREDICT_FWD = TOY_TO * 66_048 + 8_192 + TOY_A * 12_736 + TOY_TP * 83_360 + (TOY_TP - 1) * 66_048
_D_FWD = TOY_TO * 2 * (4 + 64) * 256 + 2 * (64 * 32 + 32 * 32 + 4 * TOY_TP * 32 + 32 * 32 + 64 * 32 + 32 + 64 * 32 + 64)
TRAIN_FLOPS_PER_AGENT = 5 * _PREDICT_FWD + 15 * _D_FWD


def toy_dataset(n):
    from socialways_b200.toy import create_samples, pack_scenes
    state = np.random.get_state()
    np.random.seed(30)                                   # create_toy.py:145
    samples, ts = create_samples(n, TOY_A, 3, n_per_batch=TOY_A)     # n_per_batch = n_conditions: 6-agent scenes
    np.random.set_state(state)
    o, p, t, b = pack_scenes(samples, ts)
    return dict(obsvs=o, preds=p, times=t, batches=b)


def train_cpu_baseline():
    """The oracle port of the reference's train() on the host cores, bounded sample (toy 216 / batch 64, one epoch)."""
    from oracle import socialways_oracle as so
    torch.set_num_threads(os.cpu_count() or 1)
    cpu = so.OracleTrainer(so.init_weights(n_next=2), toy_dataset(216), batch_size=64, use_social=True, pool="loop")
    t0 = time.perf_counter()
    cpu.train_epoch()
    dt = time.perf_counter() - t0
    return {"value": cpu.n_train / dt, "unit": "agents/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"toy 216 trajectories / 6-agent scenes, batch 64, one epoch of train() ({dt:.2f} s)"}


def train_parity(world, rank, dev):
    """N-rank native training vs ONE GPU running the autograd trainer (torch.optim.Adam, eager) from the same seeds on a
    small ragged data set: largest weight difference, and whether every rank holds bit-identical weights."""
    import torch.distributed as dist
    from golden_data import synthetic_scenes
    from socialways_b200.trainer import SocialWaysTrainer
    rng = np.random.RandomState(3)
    data = synthetic_scenes(list(rng.randint(1, 9, size=200)), seed=8)

    def run(native):
        torch.manual_seed(2)                                      # module construction draws the initial weights
        tr = SocialWaysTrainer(data, batch_size=256, use_social=True, n_unrolling_steps=1, device=str(dev),
                               world=None if native else (1, 0), fused_adam=native)
        np.random.seed(5)
        torch.manual_seed(5)
        for _ in range(3):                                        # epoch 1 eager, epoch 2 captures the graphs, epoch 3 replays
            (tr.train_native if native else tr.train)(verbose=False)
        return tr.reference_weights()

    w_n = run(True)
    identical = True
    if world > 1:
        for k, v in w_n.items():
            ref = v.clone()
            dist.broadcast(ref, src=0)
            identical = identical and bool(torch.equal(ref, v))
        flag = torch.tensor([1.0 if identical else 0.0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        identical = flag.item() > 0
    worst = None
    if rank == 0:
        w_1 = run(False)
        worst = max((w_n[k] - w_1[k]).abs().max().item() for k in w_1)
    return worst, identical


def train_block(args, world, rank, dev):
    import torch.distributed as dist
    from socialways_b200.native_step import NativeStep
    from socialways_b200.trainer import SocialWaysTrainer
    parity, identical = train_parity(world, rank, dev)
    data = toy_dataset(args.train_n)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    results = []
    for gbs in [int(x) for x in args.train_batches.split(",")]:
        np.random.seed(0)
        torch.manual_seed(0)
        tr = SocialWaysTrainer(data, batch_size=gbs, use_social=True, n_unrolling_steps=1, device=str(dev), fused_adam=True)
        iters = sum(1 for _ in tr._minibatches())
        for _ in range(3):                                        # eager pass, graph capture, one replayed epoch (warm-up)
            tr.train_native(verbose=False)

        def timed_epochs(**kw):
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            e0, e1 = ev(), ev()
            w0 = time.perf_counter()
            e0.record()
            for _ in range(args.train_epochs):
                res = tr.train_native(verbose=False, **kw)
            e1.record()
            torch.cuda.synchronize()
            wall = time.perf_counter() - w0
            if world > 1:
                dist.barrier()
            t = torch.tensor([e0.elapsed_time(e1) * 1e-3, wall], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            d, w = (float(x) / args.train_epochs for x in t.cpu())
            return d, w, res

        host_s, host_wall, (ade, fde) = timed_epochs()                        # the reference's RNG contract: CPU torch.rand noise
        dev_s, dev_wall, _ = timed_epochs(device_noise_seed=1234)             # noise drawn on the GPU (Philox), same otherwise
        # the GPU's own time for one iteration: the largest captured graph replayed back to back, nothing on the host between
        ent = max(tr._native_steps.values(), key=lambda e: e["step"].bs)
        r0, r1 = ev(), ev()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        r0.record()
        for _ in range(20):
            ent["graph"].replay()
        r1.record()
        torch.cuda.synchronize()
        t = torch.tensor([r0.elapsed_time(r1) / 20], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        replay_ms = float(t.item())
        for o in (tr.predictor_optimizer, tr.D_optimizer):
            o.check_status()
        results.append({"global_batch": gbs, "iterations_per_epoch": iters,
                        "ms_per_iteration": 1e3 * dev_s / iters, "agents_per_s": tr.n_train_samples / dev_s,
                        "noise": "device (Philox4x32-10, train_native(device_noise_seed=)): same distribution as train.py:473, different stream",
                        "host_noise": {"ms_per_iteration": 1e3 * host_s / iters, "agents_per_s": tr.n_train_samples / host_s,
                                       "note": "reference RNG contract: every rank draws torch.rand(global_batch, 32) on the CPU per "
                                               "iteration (train.py:473) -- the host generator, not the GPU, bounds this variant"},
                        "gpu_ms_per_iteration_graph_replay": replay_ms, "rows_per_rank_in_largest_graph": ent["step"].bs,
                        "epoch_device_s": dev_s, "epoch_wall_s": dev_wall,
                        "achieved_tflops_fp32": tr.n_train_samples * TRAIN_FLOPS_PER_AGENT / dev_s / 1e12,
                        "cuda_graphs": sum(1 for e in tr._native_steps.values() if e["graph"] is not None),
                        "train_ade": ade, "train_fde": fde})
        del tr
    out = None
    if rank == 0:
        out = {"workload": f"BASELINE configs[3]: toy set scaled to {args.train_n} trajectories (6 conditions x 3 modes, 6-agent scenes, "
                           "obs 2 / pred 2), use_social=True, unroll 1 (2 D updates + 1 G update per iteration), 4/5 of the scenes train",
               "mode": "native step (socialways_b200/native_step.py): own kernels only, one CUDA graph per mini-batch shape",
               "n_gpus": world, "parallelism": f"scenes of every mini-batch sharded x{world}",
               "collective": "none (1 GPU)" if world == 1 else
                             "3 per iteration, each = peer loads over NVLink (symmetric memory) inside the Adam kernel, rank order; no NCCL call",
               "launches_per_iteration": NativeStep.launches_per_iteration(True, 1),
               "results": results,
               "flops_per_agent_iteration_algorithmic": TRAIN_FLOPS_PER_AGENT,
               "fp32_cuda_core_peak_tflops_nominal": 74.4,
               "train_parity_max_abs": parity, "replicas_bit_identical": identical,
               "train_parity_note": "largest |w_N-rank native - w_1-GPU autograd trainer| after 3 epochs of a 200-scene ragged set (batch 256) "
                                    "(obs 8 / pred 12), same seeds; Adam amplifies fp32 rounding differences of the gradients"}
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = train_cpu_baseline()
    return out


def run_ours(args):
    import torch.distributed as dist
    import socialways_b200 as sw
    from socialways_b200 import ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    cpus = bind_to_gpu_numa(local)          # before any pinned allocation: staging buffers land on the GPU's NUMA node
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    torch.manual_seed(0)                    # random-init weights of the architecture (torch's default module init,
    gen = sw.Generator(use_social=True).to(dev)     # the stream the reference's own constructors draw from, train.py:370-376)

    n_scenes = args.scenes
    data = make_scenes(n_scenes, seed=100 + rank)            # every rank its own shard of scenes
    obsv_n, pred_n, ss = normalised(data)
    obsv_h = torch.from_numpy(obsv_n).pin_memory()
    pred_h = torch.from_numpy(pred_n).pin_memory()
    n = obsv_h.shape[0]
    g = torch.Generator().manual_seed(7 + rank)
    noise_h = torch.rand(K_SAMPLES, n, 32, generator=g).pin_memory()
    scenes = ops.SceneIndex(data["batches"], n, dev)
    traj_per_step = n * K_SAMPLES

    obsv_d, pred_d, noise_d = obsv_h.to(dev), pred_h.to(dev), noise_h.to(dev)
    out = torch.empty(K_SAMPLES, n, N_NEXT, 4, device=dev)
    pk = gen.packs()
    ev = lambda: torch.cuda.Event(enable_timing=True)
    dec_events = []

    stage_events = []

    def step_resident(timed, precision=None, events=None):
        precision = precision or args.precision
        headline = events is None
        events = dec_events if events is None else events
        if timed and headline:
            s0, s1, s2, s3 = ev(), ev(), ev(), ev()
            s0.record()
        split = precision in ("fp16x2", "fp16x2s", "bf16p")
        if split:
            enc = ops.lstm_seq_tcx(*pk["enc_tcx"], obsv_d)
        else:
            enc = ops.lstm_seq(pk["enc"], obsv_d, want_x_last=True)
        if timed and headline:
            s1.record()
        ub = ops.rows_linear(enc["h"], pk["pool_m"], pk["pool_m0"])
        if timed and headline:
            s2.record()
        if split and A_PER_SCENE <= ops.pool_tcx_max_scene():
            pooled = ops.pool_tcx(pk["pool"], pk["pool_tcx"], enc["x_last"], enc["h"], ub, scenes)
        else:
            pooled = ops.pool(pk["pool"], enc["x_last"], enc["h"], ub, scenes)
        if timed and headline:
            s3.record()
            stage_events.append((s0, s1, s2, s3))
        if timed:
            e0, e1 = ev(), ev()
            e0.record()
        if precision == "bf16":
            ops.decode_tc(pk["tc_w16"], pk["tc_f32"], enc["h"], enc["c"], pooled, noise_d, enc["x_last"], N_NEXT, out=out)
        elif precision == "fp16x2":
            ops.decode_pair(*pk["pair"], enc["h"], enc["c"], pooled, noise_d, enc["x_last"], N_NEXT, out=out)
        elif precision == "fp16x2s":
            ops.decode_tcx(*pk["tcx"], enc["h"], enc["c"], pooled, noise_d, enc["x_last"], N_NEXT, out=out)
        elif precision == "bf16p":
            ops.decode_pair(*pk["pair_bf16"], enc["h"], enc["c"], pooled, noise_d, enc["x_last"], N_NEXT, out=out, bf16=True)
        else:
            ops.decode(pk["enc"], pk["dec"], enc["h"], enc["c"], pooled, noise_d, enc["x_last"], N_NEXT, out=out)
        if timed:
            e1.record()
            events.append((e0, e1))
        return ops.bestofk_metrics(out, pred_d, ss)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident timing ----------------
    for _ in range(args.warmup):
        step_resident(False)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    t0, t1 = ev(), ev()
    t0.record()
    for _ in range(args.steps):
        m = step_resident(True)
    t1.record()
    barrier()
    ms = t0.elapsed_time(t1)
    dec_ms = sum(a.elapsed_time(b) for a, b in dec_events) / len(dec_events)
    enc_ms = sum(a.elapsed_time(b) for a, b, _, _ in stage_events) / len(stage_events)
    pool_ms = sum(c.elapsed_time(d) for _, _, c, d in stage_events) / len(stage_events)
    metrics_sum = m.sum(0).cpu().numpy() / n

    # ---------------- the other decode kernels, same inputs, reported beside the headline ----------------
    ref_out = out.clone()
    others = {}
    for other in [p for p in ("fp32", "fp16x2", "fp16x2s", "bf16", "bf16p") if p != args.precision]:
        other_events = []
        for _ in range(args.warmup):
            step_resident(False, other)
        barrier()
        o0, o1 = ev(), ev()
        o0.record()
        for _ in range(args.steps):
            m_other = step_resident(True, other, other_events)
        o1.record()
        barrier()
        others[other] = dict(ms=o0.elapsed_time(o1), dec_ms=sum(a.elapsed_time(b) for a, b in other_events) / len(other_events),
                             dev=(out - ref_out).abs().max().item(), metrics=m_other.sum(0).cpu().numpy() / n)

    # ---------------- end-to-end: host buffers in, metrics out, copies inside the timed region ----------------
    # Two variants of the public call Generator.predict_k:
    #   device noise (headline `e2e`): predict_k(obsv, None, ..., seed=, k=) -- the K x N x 32 latent noise is drawn on the GPU
    #       (Philox4x32-10, csrc/noise.cu); per step the host sends observations + ground truth and reads the metrics back.
    #   host noise (`e2e_host_noise`): the reference's contract -- the caller draws the noise on the CPU (train.py:584) and
    #       uploads it: 128 B per trajectory over PCIe, 94 % of the step's host->device bytes.
    copy_stream = torch.cuda.Stream()
    bufs = [dict(obsv=torch.empty_like(obsv_d), pred=torch.empty_like(pred_d), noise=torch.empty_like(noise_d),
                 ready=torch.cuda.Event(), free=torch.cuda.Event()) for _ in range(2)]
    res_h = [torch.empty(n, 4).pin_memory() for _ in range(2)]
    out2 = out

    def upload(i, with_noise):
        b = bufs[i % 2]
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(b["free"])
            b["obsv"].copy_(obsv_h, non_blocking=True)
            b["pred"].copy_(pred_h, non_blocking=True)
            if with_noise:
                b["noise"].copy_(noise_h, non_blocking=True)
            b["ready"].record(copy_stream)

    def e2e_steps(count, host_noise):
        main = torch.cuda.current_stream()
        for b in bufs:
            b["free"].record(main)
        upload(0, host_noise)
        for i in range(count):
            b = bufs[i % 2]
            if i + 1 < count:
                upload(i + 1, host_noise)
            main.wait_event(b["ready"])
            if host_noise:
                hat = gen.predict_k(b["obsv"], b["noise"], N_NEXT, scenes, out=out2, precision=args.precision)
            else:
                hat = gen.predict_k(b["obsv"], None, N_NEXT, scenes, out=out2, precision=args.precision, seed=(1234 + rank, i),
                                    k=K_SAMPLES, noise_buf=b["noise"])
            met = ops.bestofk_metrics(hat, b["pred"], ss)
            res_h[i % 2].copy_(met, non_blocking=True)
            b["free"].record(main)
        main.synchronize()

    e2e_res = {}
    for host_noise in (False, True):
        e2e_steps(max(2, min(args.warmup, 3)), host_noise)
        barrier()
        w0 = time.perf_counter()
        x0, x1 = ev(), ev()
        x0.record()
        e2e_steps(args.steps, host_noise)
        x1.record()
        barrier()
        e2e_res[host_noise] = dict(wall=time.perf_counter() - w0, dev=x0.elapsed_time(x1) * 1e-3)   # device clock; wall must agree
    e2e_s, e2e_wall = e2e_res[False]["dev"], e2e_res[False]["wall"]
    e2e_host_s = e2e_res[True]["dev"]
    clocks = sampler.stop() if rank == 0 else None

    train = train_block(args, world, rank, dev) if not args.no_train else None

    if world > 1:
        names = sorted(others)
        t = torch.tensor([ms, e2e_s, dec_ms, e2e_host_s] + [others[k]["ms"] for k in names] + [others[k]["dec_ms"] for k in names],
                         device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        vals = [float(x) for x in t.cpu()]
        ms, e2e_s, dec_ms, e2e_host_s = vals[:4]
        for i, k in enumerate(names):
            others[k]["ms"], others[k]["dec_ms"] = vals[4 + i], vals[4 + len(names) + i]

    if rank == 0:
        pk_ = peaks()
        value = world * traj_per_step * args.steps / (ms * 1e-3)
        e2e_value = world * traj_per_step * args.steps / e2e_s
        ach = traj_per_step * FLOPS_PER_TRAJ / (dec_ms * 1e-3) / 1e12
        peak = pk_["bf16_sustained"]
        h2d_dev = obsv_h.numel() * 4 + pred_h.numel() * 4
        h2d_host = h2d_dev + noise_h.numel() * 4
        line = {
            "metric": "predicted_trajectories_per_sec", "value": value, "unit": "traj/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": DTYPE_OF[args.precision],
            "data": "synthetic",
            "config": workload_config(args, n_scenes),
            "e2e": {"value": e2e_value, "unit": "traj/s", "h2d_bytes_per_step": h2d_dev, "d2h_bytes_per_step": n * 16,
                    "wall_s": e2e_wall, "device_s": e2e_s,
                    "noise": "drawn on the device (Philox4x32-10, predict_k(noise=None, seed=, k=)); same distribution as the "
                             "reference's torch.rand, different stream"},
            "e2e_host_noise": {"value": world * traj_per_step * args.steps / e2e_host_s, "unit": "traj/s",
                               "h2d_bytes_per_step": h2d_host, "d2h_bytes_per_step": n * 16, "device_s": e2e_host_s,
                               "h2d_gbs_per_gpu": h2d_host * args.steps / e2e_host_s / 1e9,
                               "noise": "caller-supplied host tensor (the reference's contract, train.py:584): 128 B per "
                                        "trajectory over PCIe -- the limiter of this variant when several GPUs share the "
                                        "host's memory system (profiles/r1_h2d_probe.txt: 55 GB/s for one GPU alone)"},
            "gpu_launches": 5 * args.steps,          # encoder, (u | beta) map, pooling, decode, best-of-K metrics: all own kernels
            "host_cpu_affinity": None if cpus is None else f"{len(cpus)} cpus ({cpus[0]}-{cpus[-1]}), GPU-local (NVML)",
            "train_step": train,
            "clocks": clocks,
            "roofline": {"kernel": KERNEL_OF[args.precision],
                         "bound": "tensor", "achieved": ach, "peak": peak,
                         "unit": "TFLOP/s", "frac": ach / peak,
                         "traffic": (NCU_DRAM_BYTES_PER_TRAJ[args.precision] * traj_per_step
                                     if NCU_DRAM_BYTES_PER_TRAJ[args.precision] else None),
                         "traffic_note": "bytes per launch = ncu DRAM bytes per trajectory (profiles/) x trajectories of this "
                                         "launch; algorithmic 320 B/trajectory",
                         "peak_source": f"bf16_tflops_sustained ({pk_['src']})",
                         "kernel_ms": dec_ms, "kernel_share_of_step": dec_ms / (ms / args.steps),
                         "flops_per_traj_algorithmic": FLOPS_PER_TRAJ,
                         "flops_per_traj_executed": FLOPS_PER_TRAJ_EXECUTED,
                         "note": NOTE_OF[args.precision]},
            # the other kernels of the step, live CUDA-event times (north_star asks for the pairwise kernel's HBM figure;
            # it is compute-bound -- SURVEY.md D9 -- so the fraction is small by construction)
            "secondary_kernels": {
                "pool_fwd_tcx_kernel" if args.precision in ("fp16x2", "fp16x2s", "bf16p") and A_PER_SCENE <= ops.pool_tcx_max_scene() else "pool_fwd_kernel": {
                                    "kernel_ms": pool_ms, "bound": "hbm (as asked; actually compute-bound)",
                                    "algorithmic_bytes": n * 788, "achieved_gbs": n * 788 / (pool_ms * 1e-3) / 1e9,
                                    "peak_gbs": pk_["hbm_gbs"], "frac_of_hbm": n * 788 / (pool_ms * 1e-3) / 1e9 / pk_["hbm_gbs"],
                                    "achieved_tflops_fp32": n * A_PER_SCENE * 4544 / (pool_ms * 1e-3) / 1e12,
                                    "note": "788 B/agent = x_last 16 + h 256 + (u|beta) 260 read, S 256 written"},
                "encoder": {"kernel": "lstm_seq_fwd_tcx_kernel" if args.precision in ("fp16x2", "fp16x2s", "bf16p") else "lstm_seq_fwd_kernel",
                            "kernel_ms": enc_ms,
                            "achieved_tflops": n * 528384 / (enc_ms * 1e-3) / 1e12,
                            "frac_of_tensor_peak": n * 528384 / (enc_ms * 1e-3) / 1e12 / peak}},
            "other_precisions": {
                k: {"kernel": KERNEL_OF[k], "value": world * traj_per_step * args.steps / (o["ms"] * 1e-3), "unit": "traj/s",
                    "ms_per_step": o["ms"] / args.steps, "kernel_ms": o["dec_ms"],
                    "roofline_achieved_tflops": traj_per_step * FLOPS_PER_TRAJ / (o["dec_ms"] * 1e-3) / 1e12,
                    "roofline_frac": traj_per_step * FLOPS_PER_TRAJ / (o["dec_ms"] * 1e-3) / 1e12 / peak,
                    "max_abs_dev_vs_headline_normalised": o["dev"],
                    "ade_avg": float(o["metrics"][0]), "fde_avg": float(o["metrics"][1]),
                    "ade_min": float(o["metrics"][2]), "fde_min": float(o["metrics"][3])}
                for k, o in others.items()},
            "accuracy": {"ade_avg": float(metrics_sum[0]), "fde_avg": float(metrics_sum[1]),
                         "ade_min": float(metrics_sum[2]), "fde_min": float(metrics_sum[3]),
                         "note": "random-init weights, synthetic data"},
        }
        if world == 1 and not args.no_cpu_baseline:
            v, cores, dt = cpu_arm(args.cpu_scenes)
            line["cpu_baseline"] = {"value": v, "unit": "traj/s", "cores": cores, "kind": "port",
                                    "sample": f"{args.cpu_scenes} scenes x {A_PER_SCENE} agents x K={K_SAMPLES} "
                                              f"({dt:.1f} s of CPU work, same generator)"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scenes", type=int, default=16384, help="scenes per GPU per step")
    ap.add_argument("--cpu-scenes", type=int, default=384, help="scenes in the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the train_step block (GAN training iteration, configs[3])")
    ap.add_argument("--train-n", type=int, default=65536 // 6 * 6, help="trajectories of the scaled toy set (configs[3])")
    ap.add_argument("--train-batches", default="4096,49152", help="global mini-batch sizes of the train_step block")
    ap.add_argument("--train-epochs", type=int, default=3, help="timed epochs per batch size")
    ap.add_argument("--precision", default="fp16x2", choices=["fp32", "fp16x2", "fp16x2s", "bf16", "bf16p"],
                    help="decode kernel of the headline line: fp16x2 = tcgen05 on fp16 hi/lo split operands "
                         "(fp32-faithful, default), fp32 = CUDA-core FFMA, bf16 = tcgen05 on bf16 operands (fast mode, one tile per SM), bf16p = the CTA-pair kernel on bf16 operands")
    ap.add_argument("--agents-per-scene", type=int, default=8,
                    help="8 = BASELINE configs[1] (ETH-shaped, the headline); 32 = Zara-shaped; 256 with --k 128 = the dense-crowd "
                         "stress of configs[4] (give --scenes so that scenes x agents x K fits the GPU)")
    ap.add_argument("--k", type=int, default=20, help="samples per agent (K)")
    args = ap.parse_args()
    global A_PER_SCENE, K_SAMPLES
    A_PER_SCENE, K_SAMPLES = args.agents_per_scene, args.k
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
