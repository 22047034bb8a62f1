"""CPU restatement of calc_statistics.py's two sample-set statistics -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench scripts' CPU arms may import this module; the product path
(socialways_b200/statistics.py) never does.

Follows /root/reference/calc_statistics.py:
  traj_dist_matrix     :29-32 / :54-57   d = np.mean(np.sqrt(np.sum(np.power(diff, 2), 1))) over t >= obsv_len
  compute_1nn          :7-46             mixed set, D = 1000 on the diagonal, np.argmin per row, 4 counters
  compute_wasserstein  :49-66            including the mirrored write `D[ii, jj], D[jj, ii] = dij, dij` (:58): both
                                         loops run over the full index ranges, so each cell is written twice and the
                                         later write wins -> the solver sees C[max(a,b)][min(a,b)]
The assignment itself is scipy.optimize.linear_sum_assignment -- a third-party dependency of the reference
(calc_statistics.py:3, version unpinned; scipy is present in this image), called here exactly as the reference does.
Pinned against the lifted reference functions by tests/golden/stats_*.npz (tests/test_statistics.py).
"""
import numpy as np
import scipy.optimize as op


def traj_dist_matrix(a, b, obsv_len=2):
    """a [na, T, 2], b [nb, T, 2] -> [na, nb], each entry computed like the reference's scalar expression (same dtype,
    same operation order: the per-pair np.mean runs over a contiguous vector, so do it per pair here too)."""
    na, nb = a.shape[0], b.shape[0]
    out = np.empty((na, nb), dtype=np.float64)
    for i in range(na):
        diff = a[i, None, obsv_len:] - b[:, obsv_len:]                  # [nb, Tp, 2]
        d = np.sqrt(np.sum(np.power(diff, 2), 2))                      # [nb, Tp]
        for j in range(nb):
            out[i, j] = np.mean(np.ascontiguousarray(d[j]))
    return out


def nn1_counts(reals, fakes, obsv_len=2):
    n_reals, n_fakes, n_ped = reals.shape[0], fakes.shape[0], reals.shape[1]
    counts = np.zeros(4, dtype=np.int64)
    for kk in range(n_ped):
        mixed = np.concatenate([reals[:, kk], fakes[:, kk]], axis=0)
        D = traj_dist_matrix(mixed, mixed, obsv_len)
        np.fill_diagonal(D, 1000.0)
        nn = np.argmin(D, axis=1)
        is_real = np.arange(n_reals + n_fakes) < n_reals
        nn_real = nn < n_reals
        counts += np.array([np.sum(is_real & nn_real), np.sum(is_real & ~nn_real),
                            np.sum(~is_real & ~nn_real), np.sum(~is_real & nn_real)])
    return counts


def compute_1nn(reals, fakes, obsv_len=2):
    n_reals, n_fakes, n_ped = reals.shape[0], fakes.shape[0], reals.shape[1]
    c = nn1_counts(reals, fakes, obsv_len)
    n_mixed = n_reals + n_fakes
    return np.array([(int(c[0]) + int(c[2])) / (n_mixed * n_ped), int(c[0]) / (n_reals * n_ped), int(c[2]) / (n_fakes * n_ped)])


def emd_cost_matrix(reals_k, fakes_k, obsv_len=2):
    """[n, n] matrix of one pedestrian as the reference's loop leaves it."""
    C = traj_dist_matrix(reals_k, fakes_k, obsv_len)
    lower = np.tril(C)
    return lower + np.tril(C, -1).T


def compute_wasserstein(reals, fakes, obsv_len=2):
    n_reals, n_fakes, n_ped = reals.shape[0], fakes.shape[0], reals.shape[1]
    if n_reals != n_fakes:
        raise IndexError("mirrored write out of bounds")
    cost = 0
    for kk in range(n_ped):
        D = emd_cost_matrix(reals[:, kk], fakes[:, kk], obsv_len)
        row_ind, col_ind = op.linear_sum_assignment(D)
        cost += D[row_ind, col_ind].sum()
    return cost / (n_reals * n_ped)
