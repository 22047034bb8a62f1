#!/usr/bin/env python
"""Secondary benchmark: the calc_statistics row -- 1-NN two-sample test + earth mover's distance of
K real vs K generated samples per pedestrian.  One "problem" = one pedestrian (a 2K x 2K distance matrix + argmins, a
K x K cost matrix + one linear assignment).  GPU: socialways_b200.statistics (host numpy in, results out: the H2D copy
and the D2H of the assignment are inside the timed region).  CPU: the oracle port of the reference loops on a bounded
sample.  Not the driver's contract (that is bench.py)."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def samples(k, n_ped, t_len, seed):
    rng = np.random.RandomState(seed)
    base = rng.uniform(-1, 1, size=(1, n_ped, 1, 2)) + np.cumsum(rng.normal(0, 0.1, size=(1, n_ped, t_len, 2)), axis=2)
    return ((base + rng.normal(0, 0.05, size=(k, n_ped, t_len, 2))).astype(np.float32),
            (base + rng.normal(0, 0.08, size=(k, n_ped, t_len, 2))).astype(np.float32))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--k", type=int, default=20)
    ap.add_argument("--n-ped", type=int, default=6 * 2000, help="pedestrians (= 6 per dump file x files), one launch")
    ap.add_argument("--t-len", type=int, default=4)
    ap.add_argument("--obsv-len", type=int, default=2)
    ap.add_argument("--cpu-ped", type=int, default=60)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    from socialways_b200 import statistics as st
    from oracle import statistics_oracle as so
    reals, fakes = samples(args.k, args.n_ped, args.t_len, 0)
    for _ in range(2):
        nn1, emd = st.compute_1nn(reals, fakes, args.obsv_len), st.compute_wasserstein(reals, fakes, args.obsv_len)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.reps):
        nn1, emd = st.compute_1nn(reals, fakes, args.obsv_len), st.compute_wasserstein(reals, fakes, args.obsv_len)
    torch.cuda.synchronize()
    gpu_s = (time.perf_counter() - t0) / args.reps
    rc, fc = reals[:, :args.cpu_ped], fakes[:, :args.cpu_ped]
    t0 = time.perf_counter()
    nn1_c, emd_c = so.compute_1nn(rc, fc, args.obsv_len), so.compute_wasserstein(rc, fc, args.obsv_len)
    cpu_s = time.perf_counter() - t0
    same = bool(np.array_equal(st.compute_1nn(rc, fc, args.obsv_len), nn1_c) and st.compute_wasserstein(rc, fc, args.obsv_len) == emd_c)
    print(json.dumps({"metric": "statistics_problems_per_sec", "unit": "pedestrian problems/s (1-NN + EMD)",
                      "config": {"K": args.k, "n_ped": args.n_ped, "t_len": args.t_len, "obsv_len": args.obsv_len, "dtype": "f32"},
                      "value": args.n_ped / gpu_s, "ms_per_call": 1e3 * gpu_s, "nn1": nn1.tolist(), "emd": float(emd),
                      "cpu_baseline": {"value": args.cpu_ped / cpu_s, "kind": "port", "cores": 1,
                                       "sample": f"{args.cpu_ped} pedestrians ({cpu_s:.2f} s, numpy + scipy.optimize.linear_sum_assignment)"},
                      "gpu_equals_cpu_on_sample": same}))


if __name__ == "__main__":
    main()
