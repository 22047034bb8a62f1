#!/usr/bin/env python3
"""Render the prediction dumps written by test(write_to_file=...) -- thin entry point kept for the
reference's second script (visualize.py).  It consumes only the npz schema of train.py:598-599
(`timestamp, obsvs [A,To,2], preds_our [K,A,Tp,2], preds_gtt [A,Tp,2], preds_lnr [A,Tp,2]`, metres) and
draws one PNG per dump with OpenCV: observation (blue), ground truth (green), constant-velocity
baseline (grey) and the K generated samples (red, alpha-blended).  The reference's seaborn heat-map
styling is out of scope (CPU post-processing, not on the hot path)."""
import argparse
import os

import numpy as np


def to_pixels(p, lo, scale, size, margin=20):
    q = (np.asarray(p, dtype=np.float64) - lo) * scale + margin
    q[..., 1] = size - q[..., 1]
    return np.round(q).astype(np.int32)


def draw(data, size=480):
    import cv2
    obs, gt, ours, lnr = data['obsvs'], data['preds_gtt'], data['preds_our'], data['preds_lnr']
    pts = np.concatenate([obs.reshape(-1, 2), gt.reshape(-1, 2), ours.reshape(-1, 2), lnr.reshape(-1, 2)])
    lo, hi = pts.min(0), pts.max(0)
    scale = (size - 40) / max(float((hi - lo).max()), 1e-9)
    im = np.full((size, size, 3), 255, np.uint8)
    px = lambda p: to_pixels(p, lo, scale, size)
    overlay = im.copy()
    for k in range(ours.shape[0]):
        for a in range(ours.shape[1]):
            line = np.concatenate([px(obs[a, -1:]), px(ours[k, a])])
            cv2.polylines(overlay, [line.reshape(-1, 1, 2)], False, (0, 0, 220), 1, cv2.LINE_AA)
    im = cv2.addWeighted(overlay, 0.35, im, 0.65, 0)
    for a in range(obs.shape[0]):
        cv2.polylines(im, [np.concatenate([px(obs[a, -1:]), px(lnr[a])]).reshape(-1, 1, 2)], False, (150, 150, 150), 1, cv2.LINE_AA)
        cv2.polylines(im, [np.concatenate([px(obs[a, -1:]), px(gt[a])]).reshape(-1, 1, 2)], False, (0, 160, 0), 2, cv2.LINE_AA)
        cv2.polylines(im, [px(obs[a]).reshape(-1, 1, 2)], False, (200, 60, 0), 2, cv2.LINE_AA)
        cv2.circle(im, tuple(int(v) for v in px(obs[a, 0])), 3, (200, 60, 0), -1)
    return im


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--preds-dir', default='medium/toy/socialWays')      # visualize.py:109
    ap.add_argument('--out-dir', default='medium/figs/socialWays/')      # visualize.py:110
    args = ap.parse_args()
    import cv2
    os.makedirs(args.out_dir, exist_ok=True)
    for dirpath, _, filenames in sorted(os.walk(args.preds_dir)):
        for f in sorted(filenames):
            if 'stats' in f or 'npz' not in f:
                continue
            data = np.load(os.path.join(dirpath, f))
            if data['obsvs'].shape[0] < 2:
                continue
            out = os.path.join(args.out_dir, os.path.basename(dirpath) + '-' + f.replace('.npz', '.png'))
            print('[INF] Plotting results from ' + os.path.join(dirpath, f))
            cv2.imwrite(out, draw(data))


if __name__ == '__main__':
    main()
