"""Input data of the golden train cases, regenerated without the reference (synthetic generator is
ours; the toy generator is the oracle's, itself pinned bit-exactly to the reference's)."""
import numpy as np

CASES = {
    "train_ragged.npz": ([1, 2, 6, 8, 5, 3, 4, 7, 2, 9], 5),
    "train_unroll0.npz": ([4, 4, 4, 4, 4], 6),
}


def synthetic_scenes(scene_sizes, n_past=8, n_next=12, seed=0):
    """ETH/Zara-like synthetic scenes: p0~U(-5,5)^2, v~N(0,0.4^2) per step,
    cumulative jitter N(0,0.05^2); dataset-npz layout (create_toy.py:181-187)."""
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


def case_data(case):
    sizes, seed = CASES[case]
    return synthetic_scenes(sizes, seed=seed)
