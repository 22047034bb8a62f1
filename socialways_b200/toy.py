"""Toy dataset of the reference (create_toy.py:11-54 sample generator, :162-187 scene packing).

`create_samples` keeps the reference signature and return value (samples [n,4,2] float64 scaled by
1/4, list of per-sample time-stamp arrays) and consumes the numpy global RNG exactly like the
reference (two uniform draws per sample), so `np.random.seed(30)` reproduces its dataset bit for bit.
"""
import numpy as np


def create_samples(n_samples, n_conditions, n_modes, n_per_batch=2):
    per_cond = n_samples // n_conditions
    samples = np.zeros((n_samples, 4, 2))
    time_stamps = []
    for ii in range(n_samples):
        way = (ii * n_conditions) // n_samples
        t0 = ii % per_cond + (way % (n_conditions / n_per_batch)) * per_cond        # true division, as shipped
        angle = way * (2.0 * np.pi / n_conditions)
        turn = ((ii % n_modes) - n_modes // 2) * 16 * np.pi / 180
        d2 = (float(np.random.rand(1)[0]) - 0.5) * 4 * np.pi / 180
        d3 = (float(np.random.rand(1)[0]) - 0.5) * 6 * np.pi / 180
        for q, (radius, a) in enumerate(((4, angle), (3, angle), (2, angle + turn + d2), (1, angle + turn + d2 + d3))):
            samples[ii, q] = (np.cos(a) * radius, np.sin(a) * radius)
        time_stamps.append(np.array([t0 * 4, t0 * 4 + 1, t0 * 4 + 2, t0 * 4 + 3]))
    return samples / 4, time_stamps


def pack_scenes(samples, time_stamps):
    """Group samples by first time stamp into scenes (create_toy.py:162-179); first 2 points observed."""
    groups = {}
    for ii, ts in enumerate(time_stamps):
        groups.setdefault(ts[0], []).append(ii)
    obsvs, preds, times, batches = [], [], [], []
    for members in groups.values():
        batches.append([len(obsvs), len(obsvs) + len(members)])
        for m in members:
            obsvs.append(samples[m][:2])
            preds.append(samples[m][2:])
            times.append(time_stamps[m][0])
    return (np.array(obsvs).astype(np.float32), np.array(preds).astype(np.float32),
            np.array(times).astype(np.int32), np.array(batches))


def write_to_file(real_samples, timesteps, filename):
    with open(filename, 'w+') as f:
        for ii, sample in enumerate(real_samples):
            for tt, val in enumerate(np.reshape(sample, (-1, 2))):
                f.write("%.1f %.1f %.3f %.3f\n" % (timesteps[ii][tt], ii + 1, val[0], val[1]))
    print('writing to ' + filename)
