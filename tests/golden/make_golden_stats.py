"""Golden vectors for the sample-set statistics, from the UNMODIFIED reference functions.

    python tests/golden/make_golden_stats.py          (build container only: needs /root/reference)

calc_statistics.py runs plotting code at import, so -- as for train.py -- only its top-level Import / FunctionDef nodes
are lifted with `ast` and exec'd (matplotlib stubbed); compute_1nn and compute_wasserstein then run as shipped.
Cases: the toy layout the script itself uses (real = toy-768 reshaped (-1, 6, 4, 2)[:20], calc_statistics.py:200-208;
fake = perturbed copies), in fp32 and fp64, an 8/12-shaped case (T_pred = 12 exercises numpy's 8-accumulator summation),
and a tie case (duplicated samples: first-minimum argmin, assignment ties).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import reference_harness as rh  # noqa: E402


def lifted():
    rh._install_shims()
    ns = {"__name__": "calc_statistics_lifted"}
    exec(rh._lift(os.path.join(rh.REF, "calc_statistics.py")), ns)
    return ns


def case(ns, reals, fakes, obsv_len):
    return dict(reals=reals, fakes=fakes, obsv_len=np.int64(obsv_len),
                nn1=ns["compute_1nn"](reals, fakes, obsv_len), emd=np.float64(ns["compute_wasserstein"](reals, fakes, obsv_len)))


def main():
    ns = lifted()
    toy = np.load(os.path.join(HERE, "toy_768_6.npz"))
    real = np.concatenate((toy["obsvs"], toy["preds"]), axis=1)           # calc_statistics.py:198-201
    real = real.reshape((-1, 6, 4, 2))[:20]                               # :206
    rng = np.random.RandomState(5)
    out = {}
    cases = {
        "toy_f64": (real.astype(np.float64), (real + rng.normal(0, 0.05, real.shape)).astype(np.float64), 2),
        "toy_f32": (real.astype(np.float32), (real + rng.normal(0, 0.05, real.shape)).astype(np.float32), 2),
    }
    syn = rh.synthetic_scenes([6] * 12, seed=9)
    traj = np.concatenate([syn["obsvs"], syn["preds"]], axis=1).reshape(12, 6, 20, 2)
    base = traj[:1]                                                       # 12 real / 12 fake samples around one 6-agent scene
    cases["eth_f32"] = ((base + rng.normal(0, 0.3, traj.shape)).astype(np.float32),
                        (base + rng.normal(0, 0.5, traj.shape)).astype(np.float32), 8)
    cases["eth_f64_obs0"] = ((base + rng.normal(0, 0.3, traj.shape)).astype(np.float64),
                             (base + rng.normal(0.1, 0.3, traj.shape)).astype(np.float64), 0)
    dup_r = real[:8].astype(np.float32).copy()
    dup_f = dup_r.copy()
    dup_f[1::2] += 0.25                                                   # half of the fakes coincide with reals: exact ties
    dup_r[5] = dup_r[2]
    cases["ties_f32"] = (dup_r, dup_f, 2)
    for name, (r, f, o) in cases.items():
        res = case(ns, r, f, o)
        print(name, res["nn1"], res["emd"])
        out.update({f"{name}.{k}": v for k, v in res.items()})
    path = os.path.join(HERE, "stats_cases.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
