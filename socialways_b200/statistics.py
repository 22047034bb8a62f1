"""Sample-set statistics of the reference's calc_statistics.py on the GPU.

Mirrors the reference functions by name and argument meaning:

    compute_1nn(reals, fakes, obsv_len=2)          calc_statistics.py:7-46
    compute_wasserstein(reals, fakes, obsv_len=2)  calc_statistics.py:49-66
    calc_and_store_stats(main_dir, real_samples, n_past, n_next, stats_file)   calc_statistics.py:70-120

`reals` / `fakes` are [K, nPed, T, 2] arrays (numpy or torch; fp32 or fp64 -- the dtype decides the arithmetic,
as it does in numpy).  Every pedestrian is an independent problem, so a whole directory of dumps is evaluated by
concatenating along the pedestrian axis: one launch for the 1-NN test, two for the EMD (cost matrices, then one
warp per assignment problem).  No CPU fallback: CUDA tensors are created from the inputs and the C-ABI library
must be present.
"""
import os

import numpy as np
import torch

from . import _lib
from .ops import _stream


def _device_samples(x, device):
    t = torch.as_tensor(x)
    if t.dtype not in (torch.float32, torch.float64):
        t = t.to(torch.float64)                        # numpy would compute integer inputs in float64
    if t.dim() != 4 or t.shape[-1] != 2:
        raise ValueError("samples must be [K, nPed, T, 2]")
    return t.to(device).contiguous()


def _pair(reals, fakes, device):
    device = torch.device(device if device is not None else "cuda")
    r, f = _device_samples(reals, device), _device_samples(fakes, device)
    if r.dtype != f.dtype:                             # numpy promotes mixed operands
        r, f = r.to(torch.float64), f.to(torch.float64)
    if r.shape[1:] != f.shape[1:]:
        raise ValueError("reals and fakes must agree on [nPed, T, 2]")
    return r, f


def nn1_counts(reals, fakes, obsv_len=2, device=None):
    """(Real_pos, Real_neg, Fake_pos, Fake_neg) of compute_1nn, summed over pedestrians (int64 tensor on the device)."""
    r, f = _pair(reals, fakes, device)
    counts = torch.empty(4, dtype=torch.int32, device=r.device)
    with torch.cuda.device(r.device):
        code = _lib.lib().sw_traj_nn1_counts(r.data_ptr(), f.data_ptr(), r.element_size(), r.shape[0], f.shape[0], r.shape[1],
                                             r.shape[2], int(obsv_len), counts.data_ptr(), _stream())
    _lib.check(code, "sw_traj_nn1_counts")
    return counts


def compute_1nn(reals, fakes, obsv_len=2, device=None):
    """calc_statistics.py:7-46 -> np.array([accuracy, real accuracy, fake accuracy])."""
    n_reals, n_fakes, n_ped = int(np.shape(reals)[0]), int(np.shape(fakes)[0]), int(np.shape(reals)[1])
    real_pos, _, fake_pos, _ = (int(v) for v in nn1_counts(reals, fakes, obsv_len, device).tolist())
    n_mixed = n_reals + n_fakes
    return np.array([(real_pos + fake_pos) / (n_mixed * n_ped), real_pos / (n_reals * n_ped), fake_pos / (n_fakes * n_ped)])


def emd_cost_matrices(reals, fakes, obsv_len=2, device=None):
    """[nPed, n, n] fp64 cost matrices exactly as compute_wasserstein's loop leaves them (calc_statistics.py:53-58)."""
    r, f = _pair(reals, fakes, device)
    if r.shape[0] != f.shape[0]:
        # the reference's mirrored write D[jj, ii] indexes out of bounds for non-square problems
        raise IndexError("compute_wasserstein needs n_reals == n_fakes (calc_statistics.py:58 writes D[jj, ii])")
    n, n_ped = r.shape[0], r.shape[1]
    cost = torch.empty(n_ped, n, n, dtype=torch.float64, device=r.device)
    with torch.cuda.device(r.device):
        code = _lib.lib().sw_traj_emd_cost(r.data_ptr(), f.data_ptr(), r.element_size(), n, n_ped, r.shape[2], int(obsv_len),
                                           cost.data_ptr(), _stream())
    _lib.check(code, "sw_traj_emd_cost")
    return cost


def linear_sum_assignment(cost):
    """Batched square assignment on the device: cost [P, n, n] fp64 CUDA tensor -> col_ind [P, n] int32 (row_ind is
    arange(n), as scipy returns for square problems)."""
    if not cost.is_cuda or cost.dtype != torch.float64 or cost.dim() != 3 or cost.shape[1] != cost.shape[2]:
        raise ValueError("cost must be a [P, n, n] float64 CUDA tensor")
    cost = cost.contiguous()
    p, n = cost.shape[0], cost.shape[1]
    col = torch.empty(p, n, dtype=torch.int32, device=cost.device)
    status = torch.empty(1, dtype=torch.int32, device=cost.device)
    with torch.cuda.device(cost.device):
        code = _lib.lib().sw_lsap_solve(cost.data_ptr(), n, p, col.data_ptr(), status.data_ptr(), _stream())
    _lib.check(code, "sw_lsap_solve")
    if int(status.item()) != 0:
        raise ValueError("cost matrix is infeasible")       # scipy's message for the same condition
    return col


def compute_wasserstein(reals, fakes, obsv_len=2, device=None):
    """calc_statistics.py:49-66: mean optimal-assignment cost per (sample, pedestrian)."""
    n_reals, n_ped = int(np.shape(reals)[0]), int(np.shape(reals)[1])
    cost = emd_cost_matrices(reals, fakes, obsv_len, device)
    col = linear_sum_assignment(cost)
    picked = torch.gather(cost, 2, col.long().unsqueeze(-1)).squeeze(-1).cpu().numpy()      # D[row_ind, col_ind]
    total = 0
    for kk in range(n_ped):
        total += picked[kk].sum()                       # the reference's per-pedestrian numpy sum, then python adds
    return total / (n_reals * n_ped)


def dump_samples(path, k, n_past, n_next):
    """One prediction dump of test(write_to_file=...) -> fake samples [K, nPed, n_past + n_next, 2]
    (calc_statistics.py:85-97), or None when the file has fewer than 6 pedestrians (:92-93)."""
    fake = np.load(path)
    fake_obsvs, fake_preds = fake['obsvs'], fake['preds_our']
    n_ped = fake_obsvs.shape[0]
    if n_ped < 6:
        return None
    fake_obsvs = np.concatenate([fake_obsvs.reshape((1, n_ped, n_past, 2)) for _ in range(k)], axis=0)
    return np.concatenate((fake_obsvs.reshape(-1, n_past, 2), fake_preds[:k].reshape(-1, n_next, 2)), axis=1) \
        .reshape(k, n_ped, n_past + n_next, 2)


def calc_and_store_stats(main_dir, real_samples, n_past, n_next, stats_file=None, device=None, verbose=True):
    """calc_statistics.py:70-120: walk `main_dir/<epoch>/*.npz`, average the 1-NN accuracy and the EMD over the files of
    each epoch, store the two lists.  `real_samples` is the [K, nPed, n_past+n_next, 2] array the reference builds at
    :200-208.  Every pedestrian of a file is one problem of ONE launch per statistic (distance matrices, 1-NN counts, assignment);
    the files of an epoch are evaluated one after another, with one host read per statistic and file."""
    stats_1nn, stats_wst = {}, {}
    k = real_samples.shape[0]
    for dirpath, dirnames, filenames in sorted(os.walk(main_dir)):
        cur_dir = dirpath[dirpath.rfind('/') + 1:]
        if not cur_dir.isdigit():
            continue
        epoch = int(cur_dir)
        stat_1nn_i, stat_wst_i, n_files = 0, 0, 0
        for f in sorted(filenames):
            if 'npz' not in f:
                continue
            fake_samples = dump_samples(os.path.join(dirpath, f), k, n_past, n_next)
            if fake_samples is None:
                continue
            n_ped = fake_samples.shape[1]
            real = real_samples.reshape(k, n_ped, n_past + n_next, 2)
            stat_1nn_i += compute_1nn(real, fake_samples, device=device)[0]
            stat_wst_i += compute_wasserstein(real, fake_samples, device=device)
            n_files += 1
        if verbose:
            print(main_dir, 'epoch = %d, EMD = %.5f, 1nn = %.5f' % (epoch, stat_wst_i / n_files, stat_1nn_i / n_files))
        stats_1nn[epoch] = stat_1nn_i / n_files
        stats_wst[epoch] = stat_wst_i / n_files
    stats_wst_list = [stats_wst[key] for key in sorted(stats_wst.keys())]
    stats_1nn_list = [stats_1nn[key] for key in sorted(stats_1nn.keys())]
    if stats_file is not None:
        np.savez(stats_file, stats_1nn=stats_1nn_list, stats_wst=stats_wst_list)
    return stats_1nn_list, stats_wst_list
