#!/usr/bin/env python3
"""Toy multi-modal dataset generator -- same CLI and npz schema as the reference create_toy.py
(:145-187: --npz --txt --n_conditions --n_modes --n_samples, seed 30, n_per_batch = 6).  The sample
generator is socialways_b200.toy.create_samples, bit-identical to the reference's under numpy 2
(the reference itself crashes there).  The matplotlib animation (--anim) is not ported."""
import argparse

import numpy as np

from socialways_b200.toy import create_samples, pack_scenes, write_to_file

if __name__ == '__main__':
    parser = argparse.ArgumentParser()
    np.random.seed(30)
    parser.add_argument('--txt', type=str)
    parser.add_argument('--npz', type=str)
    parser.add_argument('--n_conditions', default=6, type=int)
    parser.add_argument('--n_modes', default=3, type=int)
    parser.add_argument('--n_samples', default=3 * 6 * 12, type=int)
    args = parser.parse_args()
    samples, time_stamps = create_samples(args.n_samples, args.n_conditions, args.n_modes, n_per_batch=6)
    if args.txt is not None:
        write_to_file(samples, time_stamps, args.txt)
    obsvs, preds, times, batches = pack_scenes(samples, time_stamps)
    if args.npz is not None:
        print('writing to ' + args.npz)
        np.savez(args.npz, obsvs=obsvs, preds=preds, times=times, batches=batches)
