"""Running statistics used across frames by the analysis plugins.

Host-side mirror of pybilt/common/running_stats.py:17-112 (``RunningStats``,
``gen_running_average``): Welford's recurrence with NumPy-scalar arithmetic, so float32
inputs promote exactly as in the reference (the first pushed value keeps its dtype).
"""
from __future__ import annotations

import numpy as np


class RunningStats(object):
    def __init__(self):
        self.n = 0
        self._mean = np.float64(0.0)
        self._s = np.float64(0.0)

    def push(self, val):
        self.n += 1
        if self.n == 1:
            self._mean = np.asarray([val])[0]
            self._s = np.float64(0.0)
            return
        count = np.float64(self.n)
        delta = val - self._mean
        new_mean = self._mean + delta / count
        self._s = self._s + delta * (val - new_mean)
        self._mean = new_mean

    def mean(self):
        return self._mean if self.n > 0 else 0.0

    def variance(self):
        if self.n > 1:
            return self._s / (np.float64(self.n) - np.float64(1.0))
        return 0.0

    def deviation(self):
        return np.sqrt(self.variance())

    def reset(self):
        self.n = 0


def gen_running_average(onednparray):
    """[i, 0] running mean, [i, 1] running sample deviation (running_stats.py:90-112)."""
    stats = RunningStats()
    out = np.zeros((len(onednparray), 2))
    for i, v in enumerate(onednparray):
        stats.push(v)
        out[i, 0] = stats.mean()
        out[i, 1] = stats.deviation()
    return out


def running_rows(times, values):
    """[F,4] rows (time, value, running mean, running deviation) the per-frame analyses
    emit (e.g. analysis_protocols.py:1083-1088, :2251-2256)."""
    stats = RunningStats()
    rows = np.zeros((len(values), 4))
    for f, v in enumerate(values):
        stats.push(v)
        rows[f] = (times[f], v, stats.mean(), stats.deviation())
    return rows, stats


class BlockAverager(object):
    """Block averages of a series: consecutive blocks of ``points_per_block`` values, each reduced by a
    RunningStats; blocks with fewer than ``min_points_in_block`` values are left out of the estimate
    (pybilt/common/running_stats.py:114-300; the drivers only use the running form)."""

    def __init__(self, points_per_block=1000, min_points_in_block=500, store_data=False):
        if store_data:
            raise RuntimeError("BlockAverager(store_data=True) is not provided")
        self._blocks = [RunningStats()]
        self.n_blocks = 1
        self._points_per_block = points_per_block
        self._min_points_in_block = points_per_block - 1 if min_points_in_block > points_per_block \
            else min_points_in_block

    def push_single(self, datum):
        self._blocks[self.n_blocks - 1].push(datum)
        if self._blocks[self.n_blocks - 1].n >= self._points_per_block:
            self._blocks.append(RunningStats())
            self.n_blocks += 1

    def push_container(self, data):
        for datum in data:
            self.push_single(datum)

    def _means(self):
        return np.array([b.mean() for b in self._blocks if b.n >= self._min_points_in_block])

    def get(self):
        means = self._means()
        if len(means) > 1:
            return means.mean(), means.std() / np.sqrt(len(means))
        if len(means) == 1:
            return means[0], 0.0
        return 0.0, 0.0

    def averages_of_blocks(self):
        return self._means()

    def number_of_blocks(self):
        return self.n_blocks
