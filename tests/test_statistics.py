"""calc_statistics.py row: oracle vs the lifted reference (CPU), CUDA path vs both (GPU)."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

GOLD = np.load(os.path.join(ROOT, "tests", "golden", "stats_cases.npz"))
CASES = sorted({k.split(".")[0] for k in GOLD.files})


def load(name):
    return GOLD[f"{name}.reals"], GOLD[f"{name}.fakes"], int(GOLD[f"{name}.obsv_len"]), GOLD[f"{name}.nn1"], float(GOLD[f"{name}.emd"])


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_golden(name):
    from oracle import statistics_oracle as so
    reals, fakes, obsv_len, nn1, emd = load(name)
    assert np.array_equal(so.compute_1nn(reals, fakes, obsv_len), nn1)          # counts / integers: exact
    assert so.compute_wasserstein(reals, fakes, obsv_len) == emd                # same matrices, same solver, same sum


def test_oracle_rejects_non_square_like_the_reference():
    from oracle import statistics_oracle as so
    reals, fakes, obsv_len, _, _ = load("toy_f32")
    with pytest.raises(IndexError):
        so.compute_wasserstein(reals, fakes[:-1], obsv_len)


@pytest.mark.parametrize("kind", ["generic", "ties", "half_steps"])
def test_lsap_restatement_matches_scipy_assignments(kind):
    """oracle/lsap_ref.c (the sequential form of the kernel's algorithm, tie rule included) returns scipy's ASSIGNMENT --
    scipy.optimize.linear_sum_assignment is what the reference calls (calc_statistics.py:60) -- also when the optimum is
    not unique."""
    import scipy.optimize as op
    from oracle import build_lsap
    for seed in range(120):
        rng = np.random.RandomState(seed)
        n = int(rng.randint(1, 48))
        cost = {"generic": rng.rand(n, n) * 10 - 3, "ties": rng.randint(0, 4, size=(n, n)).astype(np.float64),
                "half_steps": np.round(rng.rand(n, n) * 5) / 2}[kind]
        assert np.array_equal(build_lsap.solve(cost), op.linear_sum_assignment(cost)[1]), (kind, seed)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_cuda_statistics_match_reference_golden(name):
    from socialways_b200 import statistics as st
    from oracle import statistics_oracle as so
    reals, fakes, obsv_len, nn1, emd = load(name)
    counts = st.nn1_counts(reals, fakes, obsv_len).cpu().numpy()
    assert np.array_equal(counts, so.nn1_counts(reals, fakes, obsv_len))
    assert np.array_equal(st.compute_1nn(reals, fakes, obsv_len), nn1)
    cost = st.emd_cost_matrices(reals, fakes, obsv_len).cpu().numpy()
    want = np.stack([so.emd_cost_matrix(reals[:, k], fakes[:, k], obsv_len) for k in range(reals.shape[1])])
    assert np.array_equal(cost, want)                                           # numpy's arithmetic order: bit-identical
    assert st.compute_wasserstein(reals, fakes, obsv_len) == emd


@pytest.mark.gpu
@pytest.mark.parametrize("n,p,seed,ties", [(1, 3, 0, False), (2, 5, 1, False), (20, 64, 2, False), (33, 40, 3, False),
                                           (97, 16, 4, False), (256, 4, 5, False), (24, 50, 6, True), (64, 10, 7, True)])
def test_lsap_kernel_matches_scipy(n, p, seed, ties):
    """One warp per problem vs scipy.optimize.linear_sum_assignment (what the reference calls): same ASSIGNMENT on
    generic costs; on tie-heavy integer costs the optimum is not unique, so the optimal COST must match exactly and
    the result must be a permutation."""
    import scipy.optimize as op
    from socialways_b200 import statistics as st
    rng = np.random.RandomState(seed)
    cost = rng.randint(0, 6, size=(p, n, n)).astype(np.float64) if ties else rng.rand(p, n, n) * 10 - 3
    col = st.linear_sum_assignment(torch.from_numpy(cost).cuda()).cpu().numpy()
    for k in range(p):
        row_ref, col_ref = op.linear_sum_assignment(cost[k])
        assert sorted(col[k].tolist()) == list(range(n))
        if ties:
            assert cost[k][np.arange(n), col[k]].sum() == cost[k][row_ref, col_ref].sum()
        else:
            assert np.array_equal(col[k], col_ref)


@pytest.mark.gpu
def test_lsap_kernel_reproduces_scipy_tie_breaking():
    """Tie-heavy integer costs: the warp-parallel scan must pick the column the sequential scan picks (lowest value; among
    equals the last unassigned column in scan order, else the first) -- same assignment as scipy and as oracle/lsap_ref.c."""
    import scipy.optimize as op
    from oracle import build_lsap
    from socialways_b200 import statistics as st
    for n, p, seed in [(7, 40, 0), (24, 40, 1), (33, 30, 2), (64, 12, 3), (100, 6, 4)]:
        rng = np.random.RandomState(seed)
        cost = rng.randint(0, 4, size=(p, n, n)).astype(np.float64)
        col = st.linear_sum_assignment(torch.from_numpy(cost).cuda()).cpu().numpy()
        for k in range(p):
            want = op.linear_sum_assignment(cost[k])[1]
            assert np.array_equal(build_lsap.solve(cost[k]), want)
            assert np.array_equal(col[k], want), (n, k)


@pytest.mark.gpu
def test_statistics_batch_over_files_and_dump_walk(tmp_path):
    """calc_and_store_stats over a directory of test(write_to_file=...) dumps == the oracle applied file by file;
    pedestrians are independent problems, so concatenating files along the pedestrian axis gives the summed counts."""
    from socialways_b200 import statistics as st
    from oracle import statistics_oracle as so
    reals, fakes, obsv_len, _, _ = load("toy_f32")
    k, n_ped = reals.shape[0], reals.shape[1]
    rng = np.random.RandomState(3)
    want_1nn, want_wst = {}, {}
    for epoch in (5, 10):
        d = tmp_path / str(epoch)
        d.mkdir()
        acc1, accw = 0, 0
        for i in range(3):
            preds = (fakes[:, :, 2:] + rng.normal(0, 0.02 * epoch, fakes[:, :, 2:].shape)).astype(np.float32)
            obsvs = reals[0, :, :2]
            np.savez(d / f"{epoch}-{i}.npz", obsvs=obsvs, preds_our=preds, preds_gtt=reals[0, :, 2:], preds_lnr=reals[0, :, 2:],
                     timestamp=i)
            fake_samples = np.concatenate([np.broadcast_to(obsvs, (k, n_ped, 2, 2)), preds], axis=2)
            acc1 += so.compute_1nn(reals, fake_samples)[0]
            accw += so.compute_wasserstein(reals, fake_samples)
        np.savez(d / "small.npz", obsvs=reals[0, :3, :2], preds_our=fakes[:, :3, 2:])      # < 6 pedestrians: skipped (:92)
        want_1nn[epoch], want_wst[epoch] = acc1 / 3, accw / 3
    got_1nn, got_wst = st.calc_and_store_stats(str(tmp_path), reals, 2, 2, stats_file=str(tmp_path / "stats20.npz"), verbose=False)
    assert got_1nn == [want_1nn[5], want_1nn[10]] and got_wst == [want_wst[5], want_wst[10]]
    saved = np.load(tmp_path / "stats20.npz")
    assert np.array_equal(saved["stats_1nn"], got_1nn) and np.array_equal(saved["stats_wst"], got_wst)
    # batching along the pedestrian axis
    both_r, both_f = np.concatenate([reals, reals], 1), np.concatenate([fakes, fakes[::-1]], 1)
    c = st.nn1_counts(both_r, both_f).cpu().numpy()
    assert np.array_equal(c, so.nn1_counts(reals, fakes) + so.nn1_counts(reals, fakes[::-1]))


def test_statistics_have_no_cpu_path():
    from socialways_b200 import statistics as st
    from socialways_b200._lib import SocialWaysCudaError
    reals, fakes, obsv_len, _, _ = load("toy_f32")
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    with pytest.raises((SocialWaysCudaError, RuntimeError, AssertionError)):
        st.compute_1nn(reals, fakes, obsv_len)
