#!/usr/bin/env python
"""1-NN accuracy and earth mover's distance of prediction dumps -- the reference's calc_statistics.py entry point
(calc_statistics.py:188-226) over the GPU statistics of socialways_b200.statistics.

    python calc_statistics.py --dataset ../data/toy/toy-768.npz --n-per-batch 6 --num-samples 20 DIR [DIR ...]

Each DIR holds `<epoch>/<epoch>-<t>.npz` dumps written by train.py's test(write_to_file=...).  The reference hard-codes
the dataset path, the directory list and the plot; here they are arguments and the plot is left to the caller
(the stored `stats<K>.npz` has the reference's keys `stats_1nn`, `stats_wst`).
"""
import argparse
import os

import numpy as np

from socialways_b200 import statistics


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("dirs", nargs="+")
    ap.add_argument("--dataset", required=True, help="npz with obsvs/preds (create_toy.py output)")
    ap.add_argument("--n-per-batch", type=int, default=6, help="pedestrians per scene (calc_statistics.py:206)")
    ap.add_argument("--num-samples", type=int, default=20, help="K real samples (calc_statistics.py:205)")
    args = ap.parse_args()
    real = np.load(args.dataset)
    real_obsv, real_pred = real['obsvs'], real['preds']
    n_past, n_next = real_obsv.shape[1], real_pred.shape[1]
    real_samples = np.concatenate((real_obsv, real_pred), axis=1)
    real_samples = real_samples.reshape((-1, args.n_per_batch, n_past + n_next, 2))[:args.num_samples]
    for main_dir in args.dirs:
        stats_file = os.path.join(main_dir, 'stats' + str(args.num_samples) + '.npz')
        if not os.path.exists(stats_file):
            statistics.calc_and_store_stats(main_dir, real_samples, n_past, n_next, stats_file)
        data = np.load(stats_file)
        print(main_dir, 'EMD', data['stats_wst'], '1NN', data['stats_1nn'])


if __name__ == "__main__":
    main()
