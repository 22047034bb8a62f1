"""Run the UNMODIFIED reference (/root/reference) on CPU to produce golden vectors.

This file is test infrastructure and only works inside the build container,
where /root/reference is mounted.  Nothing under tests/ -m gpu, smoke() or
bench.py imports it: the GPU box has no /root/reference.  The vectors it
produces are committed under tests/golden/*.npz by make_golden.py.

Method: train.py cannot run as a script (imports matplotlib,
loads ../hotel-8-12.npz, calls .cuda() and time.clock()).  So the top-level
Import / FunctionDef / ClassDef nodes of /root/reference/train.py are lifted
by `ast` and exec'd, unmodified, into a namespace; the module-level globals
the functions read (train.py:61-83, 92-124, 370-386) are injected by us.

Shims (none of them changes arithmetic):
  1. `matplotlib`, `matplotlib.pyplot`, `matplotlib.animation` stubbed in sys.modules
  2. `torch.Tensor.cuda` / `nn.Module.cuda` -> identity (CPU run)
  3. `time.clock` -> `time.perf_counter`
  4. `create_toy.create_samples`: numpy>=1.24 rejects the ragged
     `np.array([[x0, y0], ..., [x2(shape (1,)), ...]])`; `np.random.rand(1)` is
     wrapped to return a python float (ONE draw per call, same RNG stream).
"""
import ast
import copy
import os
import sys
import time
import types

import numpy as np
import torch
import torch.nn as nn

REF = os.environ.get("SOCIALWAYS_REFERENCE", "/root/reference")


def _install_shims():
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        plt.close = lambda *a, **k: None
        anim = types.ModuleType("matplotlib.animation")
        mpl.pyplot = plt
        mpl.animation = anim
        mpl.rc = lambda *a, **k: None
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt
        sys.modules["matplotlib.animation"] = anim
    if not getattr(torch.Tensor.cuda, "_sw_identity", False):
        def _tensor_cuda(self, *a, **k):
            return self
        _tensor_cuda._sw_identity = True
        torch.Tensor.cuda = _tensor_cuda

        def _module_cuda(self, *a, **k):
            return self
        nn.Module.cuda = _module_cuda
    if not hasattr(time, "clock"):
        time.clock = time.perf_counter
    if REF not in sys.path:
        sys.path.insert(0, REF)


def _lift(path, keep=(ast.Import, ast.ImportFrom, ast.FunctionDef, ast.ClassDef)):
    with open(path) as f:
        tree = ast.parse(f.read(), filename=path)
    tree.body = [n for n in tree.body if isinstance(n, keep)]
    return compile(tree, path, "exec")


class _ScalarRand:
    """np.random proxy whose rand(1) returns a float (shim 4)."""

    def __getattr__(self, name):
        return getattr(np.random, name)

    @staticmethod
    def rand(*shape):
        out = np.random.rand(*shape)
        return float(out[0]) if shape == (1,) else out


class _NumpyProxy:
    random = _ScalarRand()

    def __getattr__(self, name):
        return getattr(np, name)


def reference_toy(n_samples, n_conditions, n_modes=3, n_per_batch=6, seed=30):
    """create_toy.py: create_samples (:11-54) + the __main__ packing (:162-187), run as shipped."""
    _install_shims()
    ns = {"__name__": "create_toy_lifted"}
    exec(_lift(os.path.join(REF, "create_toy.py")), ns)
    ns["np"] = _NumpyProxy()
    np.random.seed(seed)                                   # create_toy.py:145
    samples, time_stamps = ns["create_samples"](n_samples, n_conditions, n_modes, n_per_batch=n_per_batch)
    # packing: the statements of create_toy.py:162-179 lifted verbatim from the __main__ block
    with open(os.path.join(REF, "create_toy.py")) as f:
        tree = ast.parse(f.read())
    main_if = [n for n in tree.body if isinstance(n, ast.If)][-1]
    wanted = []
    for node in main_if.body:
        src = ast.unparse(node)
        if src.startswith(("t_dict", "for ii in range(args.n_samples)", "obsvs = []", "preds = []",
                           "times = []", "batches = []", "for (key, values)", "for key, values",
                           "obsvs = np", "preds = np", "times = np")):
            wanted.append(node)
    mod = ast.Module(body=wanted, type_ignores=[])
    env = {"np": np, "samples": samples, "time_stamps": time_stamps,
           "args": types.SimpleNamespace(n_samples=n_samples)}
    exec(compile(mod, "create_toy_main", "exec"), env)
    return dict(obsvs=env["obsvs"], preds=env["preds"], times=env["times"],
                batches=np.array(env["batches"]))


class Reference:
    """The reference train.py namespace with injected module-level state."""

    def __init__(self, data, batch_size=64, hidden_size=64, use_social=True, unroll=1,
                 lr_g=1e-4, lr_d=1e-3, weight_seed=0):
        _install_shims()
        ns = {"__name__": "train_lifted"}
        exec(_lift(os.path.join(REF, "train.py")), ns)
        self.ns = ns
        g = ns
        # train.py:61-83
        g.update(n_unrolling_steps=unroll, use_info_loss=True, loss_info_w=0.5, n_latent_codes=2,
                 use_l2_loss=False, use_variety_loss=False, loss_l2_w=0.5, lr_g=lr_g, lr_d=lr_d,
                 batch_size=batch_size, hidden_size=hidden_size, num_social_features=3,
                 social_feature_size=hidden_size, noise_len=hidden_size // 2, n_lstm_layers=1,
                 use_social=use_social, epoch=1)
        # train.py:92-124
        obsv, pred = data["obsvs"].astype(np.float32).copy(), data["preds"].astype(np.float32).copy()
        the_batches = np.array(data["batches"])
        train_size = max(1, (len(the_batches) * 4) // 5)
        n_train = the_batches[train_size - 1][1]
        n_test = obsv.shape[0] - n_train
        if n_test == 0:
            n_test = 1
            the_batches = np.array([the_batches[0], the_batches[0]])
        from utils.parse_utils import Scale
        scale = Scale()
        scale.max_x = max(np.max(obsv[:, :, 0]), np.max(pred[:, :, 0]))
        scale.min_x = min(np.min(obsv[:, :, 0]), np.min(pred[:, :, 0]))
        scale.max_y = max(np.max(obsv[:, :, 1]), np.max(pred[:, :, 1]))
        scale.min_y = min(np.min(obsv[:, :, 1]), np.min(pred[:, :, 1]))
        scale.calc_scale(keep_ratio=True)
        obsv = scale.normalize(obsv)
        pred = scale.normalize(pred)
        g.update(dataset_obsv=torch.FloatTensor(obsv), dataset_pred=torch.FloatTensor(pred),
                 dataset_t=data["times"], the_batches=the_batches, train_size=train_size,
                 train_batches=the_batches[:train_size], test_batches=the_batches[train_size:],
                 n_past=obsv.shape[1], n_next=pred.shape[1], n_train_samples=n_train,
                 n_test_samples=n_test, scale=scale, ss=scale.sx)
        # train.py:370-386, construction order preserved (it fixes the RNG stream of the init)
        torch.manual_seed(weight_seed)
        import torch.optim as opt
        from itertools import chain
        g["encoder"] = g["EncoderLstm"](hidden_size, 1)
        g["feature_embedder"] = g["EmbedSocialFeatures"](3, hidden_size)
        g["attention"] = g["AttentionPooling"](hidden_size, hidden_size)
        g["decoder"] = g["DecoderFC"](hidden_size + hidden_size + hidden_size // 2)
        params = chain(g["attention"].parameters(), g["feature_embedder"].parameters(),
                       g["encoder"].parameters(), g["decoder"].parameters())
        g["predictor_optimizer"] = opt.Adam(params, lr=lr_g, betas=(0.9, 0.999))
        g["D"] = g["Discriminator"](pred.shape[1], hidden_size, 2)
        g["D_optimizer"] = opt.Adam(g["D"].parameters(), lr=lr_d, betas=(0.9, 0.999))
        g["mse_loss"] = nn.MSELoss()
        g["bce_loss"] = nn.BCELoss()

    def __getattr__(self, name):
        return self.ns[name]

    def state(self):
        out = {}
        for tag in ("encoder", "feature_embedder", "attention", "decoder", "D"):
            for k, v in self.ns[tag].state_dict().items():
                out[f"{tag}.{k}"] = v.detach().clone().numpy()
        return out

    def generator_modules(self):
        return [self.ns[t] for t in ("attention", "feature_embedder", "encoder", "decoder")]


def synthetic_scenes(scene_sizes, n_past=8, n_next=12, seed=0):
    """ETH/Zara-like synthetic scenes: p0~U(-5,5)^2, v~N(0,0.4^2),
    cumulative per-step jitter N(0,0.05^2).  Returned un-normalised, dataset-npz shaped."""
    rng = np.random.RandomState(seed)
    n = int(np.sum(scene_sizes))
    T = n_past + n_next
    p0 = rng.uniform(-5, 5, size=(n, 1, 2))
    v = rng.normal(0, 0.4, size=(n, 1, 2))
    jit = np.cumsum(rng.normal(0, 0.05, size=(n, T, 2)), axis=1)
    traj = (p0 + v * np.arange(T)[None, :, None] + jit).astype(np.float32)
    offs = np.concatenate([[0], np.cumsum(scene_sizes)])
    batches = np.stack([offs[:-1], offs[1:]], axis=1).astype(np.int64)
    return dict(obsvs=traj[:, :n_past], preds=traj[:, n_past:],
                times=np.repeat(np.arange(len(scene_sizes)), scene_sizes).astype(np.int32), batches=batches)
