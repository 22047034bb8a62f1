"""Golden prediction dumps: the UNMODIFIED reference test(n, write_to_file=dir) (train.py:563-599)
run on the toy set, its npz files re-packed into one fixture.

    python tests/golden/make_golden_dump.py        (build container only)
"""
import contextlib
import glob
import io
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import reference_harness as rh  # noqa: E402

torch.set_num_threads(1)


def main():
    out = {}
    data = {k: v for k, v in np.load(os.path.join(HERE, "toy_216_6.npz")).items()}
    for social in (True, False):
        tag = "soc" if social else "nos"
        ref = rh.Reference(data, batch_size=64, use_social=social, weight_seed=2)
        if social:
            out.update({f"w.{k}": v for k, v in ref.state().items()})
        ref.ns["epoch"] = 35
        with tempfile.TemporaryDirectory() as d:
            torch.manual_seed(21)
            with contextlib.redirect_stdout(io.StringIO()) as txt:
                ref.test(4, write_to_file=d)
            files = sorted(os.path.basename(f) for f in glob.glob(os.path.join(d, "*.npz")))
            out[f"{tag}.files"] = np.array(files)
            out[f"{tag}.stdout_last"] = np.array(txt.getvalue().strip().splitlines()[-1])
            for f in files:
                z = np.load(os.path.join(d, f))
                for key in z.files:
                    out[f"{tag}.{f}.{key}"] = z[key]
                    out[f"{tag}.{f}.{key}.dtype"] = np.array(str(z[key].dtype))
    path = os.path.join(HERE, "dump_toy_216.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB;", len(files), "files per run;", out["soc.stdout_last"])


if __name__ == "__main__":
    main()
