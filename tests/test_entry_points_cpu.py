"""The kept entry points parse their command lines on a machine without a GPU (no kernel is touched by --help)."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT


@pytest.mark.parametrize("script,needle", [("train.py", "--fused-adam"), ("create_dataset.py", "--int64-batches"),
                                           ("calc_statistics.py", "--num-samples"), ("create_toy.py", "--"),
                                           ("bench.py", "--agents-per-scene"), ("bench_train.py", "--graph"),
                                           ("bench_stats.py", "--n-ped")])
def test_entry_point_help(script, needle):
    r = subprocess.run([sys.executable, os.path.join(ROOT, script), "--help"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                       text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-500:]
    assert needle in r.stdout


def test_create_dataset_entry_point_round_trip(tmp_path):
    """create_dataset.py on a synthetic obsmat.txt -> npz with the reference's four keys (create_dataset.py:13)."""
    import numpy as np
    from conftest import load_golden
    g = load_golden("dataset_biwi.npz")
    src, dst = tmp_path / "obsmat.txt", tmp_path / "data-8-12.npz"
    src.write_bytes(g["eth_8_12.text"].tobytes())
    r = subprocess.run([sys.executable, os.path.join(ROOT, "create_dataset.py"), str(src), str(dst)], stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-500:]
    z = np.load(dst)
    assert sorted(z.files) == ["batches", "obsvs", "preds", "times"]
    assert np.array_equal(z["obsvs"], g["eth_8_12.obsvs"]) and np.array_equal(z["batches"], g["eth_8_12.batches"])
