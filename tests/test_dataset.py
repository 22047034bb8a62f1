"""Dataset windowing row: socialways_b200.dataset vs the arrays the unmodified reference produced
(tests/golden/dataset_biwi.npz, make_golden_dataset.py).  Integer / index work: bit-exact."""
import os

import numpy as np
import pytest

from conftest import load_golden

G = load_golden("dataset_biwi.npz")
CASES = sorted({k.split(".")[0] for k in G})


def _parse(tmp_path, name):
    from socialways_b200.dataset import BIWIParser
    path = tmp_path / "obsmat.txt"
    path.write_bytes(G[f"{name}.text"].tobytes())
    parser = BIWIParser()
    parser.load(str(path))
    return parser


@pytest.mark.parametrize("name", CASES)
def test_parser_and_windowing_match_the_reference(tmp_path, name):
    from socialways_b200.dataset import create_dataset
    parser = _parse(tmp_path, name)
    assert parser.interval == int(G[f"{name}.interval"]) and len(parser.p_data) == int(G[f"{name}.n_tracks"])
    sc = parser.scale
    assert np.array_equal(np.array([sc.min_x, sc.max_x, sc.min_y, sc.max_y, sc.sx]), G[f"{name}.scale"])
    n_past, n_next = (int(v) for v in G[f"{name}.cfg"])
    t_range = range(parser.t_data[0][0], parser.t_data[-1][-1], parser.interval)
    obsvs, preds, times, batches = create_dataset(parser.p_data, parser.t_data, t_range, n_past, n_next)
    assert obsvs.dtype == np.float32 and preds.dtype == np.float32 and batches.dtype == np.int16
    assert np.array_equal(obsvs, G[f"{name}.obsvs"]) and np.array_equal(preds, G[f"{name}.preds"])
    assert np.array_equal(np.array(times), G[f"{name}.times"]) and np.array_equal(batches, G[f"{name}.batches"])


def test_unit_interval_drops_samples_like_the_reference():
    # t == last_included_t + 1 neither opens nor joins a scene (utils/parse_utils.py:482-487): fewer samples than time stamps
    assert len(G["unit_3_2.times"]) > len(G["unit_3_2.obsvs"])


def test_gap_inside_a_window_fails_like_the_reference():
    from socialways_b200.dataset import create_dataset
    t = np.array([0, 1, 2, 4, 5, 6, 7], dtype=np.int32)            # stamp 3 missing: the window 0..6 has 7 rows instead of 8... ragged
    p = np.stack([t, t], 1).astype(np.float64)
    full = np.arange(8, dtype=np.int32)
    with pytest.raises(ValueError):
        create_dataset([np.stack([full, full], 1).astype(np.float64), p], [full, t], range(0, 8, 1), 4, 3)


def test_int64_index_option_fixes_the_int16_wrap():
    from socialways_b200.dataset import create_dataset
    n_tracks, length = 300, 125                                     # 300 x 113 windows = 33 900 samples > 32 767
    t = np.arange(length, dtype=np.int32) * 2
    p_data = [np.stack([np.arange(length) + i, np.arange(length) - i], 1).astype(np.float64) for i in range(n_tracks)]
    t_data = [t for _ in range(n_tracks)]
    with np.errstate(over="ignore"), pytest.raises(ValueError):
        # the reference's int16 scene table wraps negative beyond 32 767 samples; its slices come back empty and its
        # own np.concatenate raises -- reproduced by the default index dtype
        create_dataset(p_data, t_data, range(0, 2 * length, 2), 8, 5)
    o64, _, times, b64 = create_dataset(p_data, t_data, range(0, 2 * length, 2), 8, 5, index_dtype=np.int64)
    assert len(times) == 33900 and b64[-1, 1] == 33900 and np.all(b64[:, 1] - b64[:, 0] == n_tracks)
    assert o64.shape == (33900, 8, 2)
