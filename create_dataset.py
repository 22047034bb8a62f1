#!/usr/bin/env python
"""Build an obsvs/preds/times/batches npz from an ETH/UCY `obsmat.txt` -- the reference's create_dataset.py, with its
two hard-coded paths as arguments.

    python create_dataset.py path-to-dataset/obsmat.txt ../data-8-12.npz [--n-past 8 --n-next 12] [--int64-batches]
"""
import argparse

import numpy as np

from socialways_b200.dataset import BIWIParser, create_dataset


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("annot_file")
    ap.add_argument("npz_out_file")
    ap.add_argument("--n-past", type=int, default=8)
    ap.add_argument("--n-next", type=int, default=12)
    ap.add_argument("--int64-batches", action="store_true",
                    help="store scene ranges as int64 (the reference's int16 wraps beyond 32 767 samples)")
    args = ap.parse_args()
    parser = BIWIParser()
    parser.load(args.annot_file)
    obsvs, preds, times, batches = create_dataset(parser.p_data, parser.t_data,
                                                  range(parser.t_data[0][0], parser.t_data[-1][-1], parser.interval),
                                                  args.n_past, args.n_next,
                                                  index_dtype=np.int64 if args.int64_batches else np.int16)
    np.savez(args.npz_out_file, obsvs=obsvs, preds=preds, times=times, batches=batches)
    print('dataset was created successfully and stored in:', args.npz_out_file)


if __name__ == "__main__":
    main()
