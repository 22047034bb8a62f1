"""Scale (reference utils/parse_utils.py:11-76): isotropic min/max normalisation of positions.

Same attributes and method signatures as the reference class (min_x/max_x/min_y/max_y, sx, sy,
calc_scale(keep_ratio), normalize(data, shift, inPlace), denormalize(data, shift, inPlace)); the
per-ndim branches of the reference are expressed once over the last axis.
"""
import math

import numpy as np


class Scale(object):
    def __init__(self):
        self.min_x = +math.inf
        self.max_x = -math.inf
        self.min_y = +math.inf
        self.max_y = -math.inf
        self.sx, self.sy = 1, 1

    def calc_scale(self, keep_ratio=True):
        self.sx = 1 / (self.max_x - self.min_x)
        self.sy = 1 / (self.max_y - self.min_y)
        if keep_ratio:
            self.sx = self.sy = min(self.sx, self.sy)

    def normalize(self, data, shift=True, inPlace=True):
        if not 1 <= data.ndim <= 4:
            return False
        out = data if inPlace else np.copy(data)
        out[..., 0] = (data[..., 0] - self.min_x * shift) * self.sx
        out[..., 1] = (data[..., 1] - self.min_y * shift) * self.sy
        return out

    def denormalize(self, data, shift=True, inPlace=False):
        if not 1 <= data.ndim <= 4:
            return False
        out = data if inPlace else np.copy(data)
        out[..., 0] = data[..., 0] / self.sx + self.min_x * shift
        out[..., 1] = data[..., 1] / self.sy + self.min_y * shift
        return out
