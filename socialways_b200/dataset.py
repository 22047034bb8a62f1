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
    dataset_t0 = [c[0] for c in cand]
    dataset_x = [p_data[c[1]][c[2]:c[3]] for c in cand]
    dataset_y = [p_data[c[1]][c[3]:c[4] + 1] for c in cand]

    sub_batches = []
    last_included_t = -1000
    min_interval = 1
    for i, t in enumerate(dataset_t0):
        if t > last_included_t + min_interval:
            sub_batches.append([i, i + 1])
            last_included_t = t
        if t == last_included_t:
            sub_batches[-1][1] = i + 1
    sub_batches = np.array(sub_batches).astype(index_dtype)
    dataset_x_, dataset_y_ = [], []
    last_ind = 0
    for sb in sub_batches:
        dataset_x_.append(dataset_x[sb[0]:sb[1]])
        dataset_y_.append(dataset_y[sb[0]:sb[1]])
        sb[1] = sb[1] - sb[0] + last_ind
        sb[0] = last_ind
        last_ind = sb[1]
    dataset_x = np.concatenate(dataset_x_)
    dataset_y = np.concatenate(dataset_y_)
    sub_batches = np.array(sub_batches).astype(index_dtype)
    return np.array(dataset_x).astype(np.float32), np.array(dataset_y).astype(np.float32), dataset_t0, sub_batches


class BIWIParser:
    """utils/parse_utils.py:231-321: ETH/UCY `obsmat.txt` rows `t id px pz py vx vz vy` -> per-pedestrian tracks."""

    def __init__(self):
        self.scale = Scale()
        self.all_ids = list()
        self.delimit = ' '
        self.p_data, self.v_data, self.t_data = [], [], []
        self.min_t, self.max_t, self.interval = int(sys.maxsize), -1, -1

    def load(self, filename, down_sample=1):
        self.all_ids.clear()
        if 'zara' in filename:
            self.delimit = '\t'
        file_names = []
        if '*' in filename:
            files_path, extension = filename[:filename.index('*')], filename[filename.index('*') + 1:]
            file_names = [files_path + f for f in os.listdir(files_path) if f.endswith(extension)]
        else:
            file_names.append(filename)
        pos, vel, tim = {}, {}, {}
        for file in file_names:
            if not os.path.exists(file):
                raise ValueError("No such file or directory:", file)
            id_list = []
            with open(file, 'r') as data_file:
                for row in data_file.readlines():
                    row = [c for c in row.split(self.delimit) if c != '']
                    if len(row) < 8:
                        continue
                    ts, pid = float(row[0]), round(float(row[1]))
                    if ts % down_sample != 0:
                        continue
                    self.min_t, self.max_t = min(self.min_t, ts), max(self.max_t, ts)
                    if pid not in id_list:
                        id_list.append(pid)
                        pos[pid], vel[pid], tim[pid] = [], [], []
                    pos[pid].append([float(row[2]), float(row[4])])
                    vel[pid].append([float(row[5]), float(row[7])])
                    tim[pid].append(ts)
            self.all_ids += id_list
        for ped_t in tim.values():
            if len(ped_t) > 1:
                interval = int(round(ped_t[1] - ped_t[0]))
                if interval > 0:
                    self.interval = interval
                    break
        for key in pos:
            self.p_data.append(np.array(pos[key]))
            self.v_data.append(np.array(vel[key]))
            self.t_data.append(np.array(tim[key]).astype(np.int32))
        for poss_i in self.p_data:
            self.scale.min_x = min(self.scale.min_x, min(poss_i[:, 0]))
            self.scale.max_x = max(self.scale.max_x, max(poss_i[:, 0]))
            self.scale.min_y = min(self.scale.min_y, min(poss_i[:, 1]))
            self.scale.max_y = max(self.scale.max_y, max(poss_i[:, 1]))
        self.scale.calc_scale()
