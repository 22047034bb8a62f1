"""Dataset windowing of the reference: `create_dataset` (utils/parse_utils.py:457-508) and the BIWI
`obsmat.txt` reader it is fed by (`BIWIParser`, utils/parse_utils.py:231-321; entry point create_dataset.py).

This is host-side data preparation (numpy in the reference, numpy here); what changes is the algorithm: the
reference scans every track with three `np.where` calls for every integer time stamp of the recording
(O(T_range x n_tracks x track_len) interpreter work), here every track is windowed at once with `searchsorted`.
Outputs are identical arrays, including the reference's behaviours:

* time stamps are visited in ascending order, tracks in list order inside one time stamp (:461-462);
* a window needs the three stamps t - step*n_past, t, t + step*(n_next-1) to be PRESENT in the track; the rows in
  between are taken as they are, so a track with a gap inside the window yields a short sample and the final
  `np.concatenate` fails exactly as in the reference (ValueError);
* scene grouping (:478-489): a sample opens a new scene when its t exceeds the last scene's t by MORE than 1, joins
  it when equal, and is silently DROPPED when t == last + 1 (unit-interval recordings);
* `sub_batches` are int16 like the reference -- whose scene table wraps negative beyond 32 767 samples, after which its
  own `np.concatenate` raises.  `index_dtype=np.int64` is the fix; the default reproduces the
  reference, failure included.
"""
import os
import sys

import numpy as np

from .scale import Scale


def _first_index(track_t, values):
    """Index of the FIRST occurrence of each value in track_t (np.where(...)[0][0]), -1 where absent.  Tracks are not
    assumed sorted (a stable argsort makes them so)."""
    order = np.argsort(track_t, kind="stable")
    sorted_t = track_t[order]
    pos = np.searchsorted(sorted_t, values, side="left")
    ok = (pos < len(sorted_t))
    ok[ok] &= sorted_t[pos[ok]] == values[ok]
    return np.where(ok, order[np.minimum(pos, len(sorted_t) - 1)], -1)


def create_dataset(p_data, t_data, t_range, n_past=8, n_next=12, index_dtype=np.int16):
    step = t_range.step
    cand = []                                                   # (t, track, tP_ind, t0_ind, tF_ind)
    for i, tt in enumerate(t_data):
        tt = np.asarray(tt)
        if len(tt) == 0:
            continue
        t0_vals = np.unique(tt)
        t0_vals = t0_vals[(t0_vals >= t_range.start) & (t0_vals < t_range.stop)]        # `for t in range(start, stop, 1)`
        if len(t0_vals) == 0:
            continue
        i0 = _first_index(tt, t0_vals)
        iP = _first_index(tt, t0_vals - step * n_past)
        iF = _first_index(tt, t0_vals + step * (n_next - 1))
        keep = (iP >= 0) & (iF >= 0)
        for t, a, b, c in zip(t0_vals[keep], iP[keep], i0[keep], iF[keep]):
            cand.append((int(t), i, int(a), int(b), int(c)))
    cand.sort(key=lambda c: (c[0], c[1]))                       # time-major, track order inside a time stamp
    times = np.array([c[0] for c in cand], dtype=np.int64)
    xs = [p_data[c[1]][c[2]:c[3]] for c in cand]
    ys = [p_data[c[1]][c[3]:c[4] + 1] for c in cand]
    if len(times) == 0:
        raise ValueError("need at least one array to concatenate")      # what the reference's np.concatenate([]) raises

    # ---- scene grouping (:478-489) as a run-length pass over the sorted time stamps.  The reference walks the samples
    # with `last_included_t`: a stamp opens a scene when it exceeds the last INCLUDED stamp by more than 1, its equal stamps
    # join, and stamps at exactly last + 1 are skipped.  On the sorted unique stamps u this is: inside every run of
    # consecutive integers the stamps at even positions of the run are included, the odd ones dropped.
    u, first, count = np.unique(times, return_index=True, return_counts=True)
    run_start = np.concatenate([[True], np.diff(u) != 1])
    pos_in_run = np.arange(len(u)) - np.maximum.accumulate(np.where(run_start, np.arange(len(u)), 0))
    included = (pos_in_run % 2) == 0
    table = np.stack([first[included], first[included] + count[included]], axis=1)     # [start, end) in the unfiltered list
    if table.max() > np.iinfo(index_dtype).max:
        # the reference stores this table as int16 BEFORE it slices with it (:490): beyond 32 767 samples the wrapped
        # (negative) bounds select wrong / empty ranges and its np.concatenate fails -- same table, same slicing, same error
        wrapped = table.astype(index_dtype)
        np.concatenate([np.array(xs[a:b]) for a, b in wrapped])
        np.concatenate([np.array(ys[a:b]) for a, b in wrapped])
    keep = np.concatenate([np.arange(a, b) for a, b in table])
    widths = {len(xs[i]) for i in keep} | {-len(ys[i]) for i in keep}
    if len(widths) != 2:                                        # a gap inside a window: ragged samples, the reference's concatenate raises
        raise ValueError("all the input array dimensions except for the concatenation axis must match exactly")
    dataset_x = np.stack([xs[i] for i in keep]).astype(np.float32)
    dataset_y = np.stack([ys[i] for i in keep]).astype(np.float32)
    ends = np.cumsum(count[included])
    sub_batches = np.stack([ends - count[included], ends], axis=1).astype(index_dtype)
    return dataset_x, dataset_y, times.tolist(), sub_batches


class BIWIParser:
    """utils/parse_utils.py:231-321: ETH/UCY `obsmat.txt` rows `t id px pz py vx vz vy` -> per-pedestrian tracks.
    The reference appends row by row (np.hstack per sample: quadratic); here a file is parsed into one array and split
    into tracks with a stable sort.  Same outputs: tracks in first-appearance order, a pedestrian that reappears in a later
    file REPLACES its earlier track in place (the per-file `id_list`, :262,277-281), int32 time stamps, Scale over all tracks."""

    def __init__(self):
        self.scale = Scale()
        self.all_ids = list()
        self.delimit = ' '
        self.p_data, self.v_data, self.t_data = [], [], []
        self.min_t, self.max_t, self.interval = int(sys.maxsize), -1, -1

    def _rows(self, path, down_sample):
        """[n, 8] float64 table of the usable rows of one file (>= 8 non-empty fields, t % down_sample == 0)."""
        with open(path, 'r') as f:
            fields = [[c for c in line.split(self.delimit) if c != ''] for line in f]
        rows = np.array([[float(c) for c in r[:8]] for r in fields if len(r) >= 8], dtype=np.float64).reshape(-1, 8)
        return rows[rows[:, 0] % down_sample == 0]

    def load(self, filename, down_sample=1):
        self.all_ids.clear()
        if 'zara' in filename:
            self.delimit = '\t'
        if '*' in filename:
            folder, extension = filename[:filename.index('*')], filename[filename.index('*') + 1:]
            paths = [folder + f for f in os.listdir(folder) if f.endswith(extension)]
        else:
            paths = [filename]
        tracks = {}                                             # pedestrian id -> (positions, velocities, times); insertion-ordered
        for path in paths:
            if not os.path.exists(path):
                raise ValueError("No such file or directory:", path)
            rows = self._rows(path, down_sample)
            if len(rows) == 0:
                continue
            self.min_t, self.max_t = min(self.min_t, rows[:, 0].min()), max(self.max_t, rows[:, 0].max())
            ids = np.round(rows[:, 1]).astype(np.int64)
            order = np.argsort(ids, kind="stable")              # rows of one pedestrian stay in file order
            uniq, first, count = np.unique(ids[order], return_index=True, return_counts=True)
            by_appearance = np.argsort(order[first], kind="stable")
            for k in by_appearance:
                sel = order[first[k]:first[k] + count[k]]
                tracks[int(uniq[k])] = (rows[sel][:, [2, 4]], rows[sel][:, [5, 7]], rows[sel, 0])
            self.all_ids += [int(uniq[k]) for k in by_appearance]
        for _, _, ts in tracks.values():
            if len(ts) > 1 and int(round(ts[1] - ts[0])) > 0:
                self.interval = int(round(ts[1] - ts[0]))
                break
        for pos, vel, ts in tracks.values():
            self.p_data.append(np.ascontiguousarray(pos))
            self.v_data.append(np.ascontiguousarray(vel))
            self.t_data.append(ts.astype(np.int32))
        if self.p_data:
            allp = np.concatenate(self.p_data)
            self.scale.min_x, self.scale.max_x = min(self.scale.min_x, allp[:, 0].min()), max(self.scale.max_x, allp[:, 0].max())
            self.scale.min_y, self.scale.max_y = min(self.scale.min_y, allp[:, 1].min()), max(self.scale.max_y, allp[:, 1].max())
        self.scale.calc_scale()
