"""Prediction-dump path: test(n, write_to_file=dir) (train.py:563-599) writes the npz files that
visualize.py (:173-183) and calc_statistics.py (:86-101) read.  Golden = the files the unmodified reference wrote from
the same weights and torch seed (tests/golden/dump_toy_216.npz, make_golden_dump.py)."""
import os

import numpy as np
import pytest
import torch

from conftest import golden_weights, load_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("social", [True, False])
def test_dump_files_match_the_reference(tmp_path, capsys, social):
    from socialways_b200.trainer import SocialWaysTrainer
    g = load_golden("dump_toy_216.npz")
    data = load_golden("toy_216_6.npz")
    tag = "soc" if social else "nos"
    tr = SocialWaysTrainer(data, batch_size=64, use_social=social, weights=golden_weights(g))
    tr.epoch = 35
    torch.manual_seed(21)
    res = tr.test(4, write_to_file=str(tmp_path))
    files = sorted(os.listdir(tmp_path))
    assert files == list(g[f"{tag}.files"])                                   # '<epoch>-<timestamp>.npz'
    for f in files:
        z = np.load(tmp_path / f)
        assert sorted(z.files) == ["obsvs", "preds_gtt", "preds_lnr", "preds_our", "timestamp"]
        for key in z.files:
            want = g[f"{tag}.{f}.{key}"]
            assert z[key].shape == want.shape and str(z[key].dtype) == str(g[f"{tag}.{f}.{key}.dtype"]), (f, key)
            if key == "preds_our":      # K samples, denormalised: the decode kernel vs torch CPU (toy coordinates are O(1..10))
                np.testing.assert_allclose(z[key], want, atol=5e-5, rtol=0)
            else:                       # copies of the data / the constant-velocity baseline
                np.testing.assert_allclose(z[key], want, atol=1e-6, rtol=0)
    line = 'Avg ADE,FDE (12)= (%.3f, %.3f) | Min(20) ADE,FDE (12)= (%.3f, %.3f)' % (
        res["ade_avg"], res["fde_avg"], res["ade_min"], res["fde_min"])
    assert line == str(g[f"{tag}.stdout_last"])
    assert line in capsys.readouterr().out


def test_dump_feeds_the_statistics(tmp_path):
    """The dump written here is what calc_statistics reads (obsvs [A,2,2], preds_our [K,A,2,2])."""
    from socialways_b200 import statistics as st
    from socialways_b200.trainer import SocialWaysTrainer
    g = load_golden("dump_toy_216.npz")
    data = load_golden("toy_216_6.npz")
    tr = SocialWaysTrainer(data, batch_size=64, use_social=True, weights=golden_weights(g))
    tr.epoch = 5
    d = tmp_path / "5"
    d.mkdir()
    torch.manual_seed(3)
    tr.test(20, write_to_file=str(d), verbose=False)
    real = np.concatenate((data["obsvs"], data["preds"]), axis=1).reshape((-1, 6, 4, 2))[:20]
    s1, sw = st.calc_and_store_stats(str(tmp_path), real, 2, 2, verbose=False)
    assert len(s1) == 1 and 0.0 <= s1[0] <= 1.0 and np.isfinite(sw[0]) and sw[0] > 0
