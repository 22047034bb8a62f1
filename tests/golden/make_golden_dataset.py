"""Golden vectors for the dataset windowing row: the UNMODIFIED reference `BIWIParser.load` and
`create_dataset` (utils/parse_utils.py:231-321, :457-508), driven like create_dataset.py does, on a small synthetic
`obsmat.txt` (ETH layout: t id px pz py vx vz vy; frame step 6, tracks entering/leaving, one track with a gap that makes a
window short is kept OUT of range so the reference itself succeeds; a unit-interval case exercises the dropped-sample
rule of :482-487).

    python tests/golden/make_golden_dataset.py        (build container only)
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import reference_harness as rh  # noqa: E402


def synth_obsmat(seed, n_ped, n_frames, step):
    rng = np.random.RandomState(seed)
    rows = []
    for pid in range(1, n_ped + 1):
        t_in = rng.randint(0, n_frames // 2)
        length = rng.randint(6, n_frames - t_in + 1)
        p = rng.uniform(-5, 5, 2)
        v = rng.normal(0, 0.4, 2)
        for k in range(length):
            t = (t_in + k) * step
            p = p + v + rng.normal(0, 0.05, 2)
            rows.append((t, pid, p[0], 0.0, p[1], v[0], 0.0, v[1]))
    rows.sort(key=lambda r: (r[0], r[1]))
    return "".join("  %.7e  %.7e  %.7e  %.7e  %.7e  %.7e  %.7e  %.7e\n" % r for r in rows)


def main():
    rh._install_shims()
    from utils.parse_utils import BIWIParser, create_dataset          # the reference's own module, unmodified
    out = {}
    for name, (seed, n_ped, n_frames, step, n_past, n_next) in {
            "eth_8_12": (1, 40, 60, 6, 8, 12), "short_2_2": (2, 25, 20, 10, 2, 2), "unit_3_2": (3, 12, 16, 1, 3, 2)}.items():
        text = synth_obsmat(seed, n_ped, n_frames, step)
        with tempfile.TemporaryDirectory() as d:
            path = os.path.join(d, "obsmat.txt")
            with open(path, "w") as f:
                f.write(text)
            parser = BIWIParser()
            parser.p_data, parser.v_data, parser.t_data = [], [], []
            parser.load(path)
        t_range = range(parser.t_data[0][0], parser.t_data[-1][-1], parser.interval)    # create_dataset.py:9-11
        obsvs, preds, times, batches = create_dataset(parser.p_data, parser.t_data, t_range, n_past, n_next)
        print(name, obsvs.shape, preds.shape, len(times), batches.shape, "interval", parser.interval)
        out[f"{name}.text"] = np.frombuffer(text.encode(), dtype=np.uint8)
        out[f"{name}.cfg"] = np.array([n_past, n_next])
        out[f"{name}.obsvs"], out[f"{name}.preds"] = obsvs, preds
        out[f"{name}.times"], out[f"{name}.batches"] = np.array(times), batches
        out[f"{name}.interval"] = np.int64(parser.interval)
        out[f"{name}.scale"] = np.array([parser.scale.min_x, parser.scale.max_x, parser.scale.min_y, parser.scale.max_y, parser.scale.sx])
        out[f"{name}.n_tracks"] = np.int64(len(parser.p_data))
    path = os.path.join(HERE, "dataset_biwi.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
