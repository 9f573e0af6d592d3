"""Summation plan reproducing NumPy's float64 pairwise `.sum()` over n contiguous elements.

The membrane centre of mass is `(f64(u[:,k]) * masses).sum() / masses.sum()`
(pybilt/bilayer_analyzer/com_frame.py:175-178) and is subtracted from every coordinate
before a float32 rounding (:180), so it has to be bit-identical, i.e. summed in NumPy's
order (numpy/_core/src/umath/loops_utils.h.src, DOUBLE_pairwise_sum):

    n < 8    : sequential from 0
    n <= 128 : eight running accumulators over stride-8 lanes, combined
               ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), then the n % 8 tail sequentially
    n > 128  : split at n2 = n/2 - (n/2) % 8, recurse, add

The plan lists the <=128-element leaves and the binary tree that combines them; the
device evaluates leaves with one warp each and the tree level by level.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class PairwisePlan:
    n: int
    leaf_start: np.ndarray     # int32 [L]
    leaf_len: np.ndarray       # int32 [L]
    node_left: np.ndarray      # int32 [K]  child ids: leaf l -> l, internal node k -> L + k
    node_right: np.ndarray     # int32 [K]
    level_off: np.ndarray      # int32 [H+1] internal nodes are ordered by height
    @property
    def n_leaves(self) -> int:
        return int(self.leaf_start.shape[0])

    @property
    def n_nodes(self) -> int:
        return int(self.node_left.shape[0])

    @property
    def n_levels(self) -> int:
        return int(self.level_off.shape[0]) - 1


def build_plan(n: int) -> PairwisePlan:
    leaves = []          # (start, len)
    nodes = []           # (left_ref, right_ref, height); refs: ('l', i) or ('n', i)

    def rec(start: int, m: int):
        if m <= 128:
            leaves.append((start, m))
            return ("l", len(leaves) - 1), 0
        n2 = m // 2
        n2 -= n2 % 8
        (lref, lh) = rec(start, n2)
        (rref, rh) = rec(start + n2, m - n2)
        h = max(lh, rh) + 1
        nodes.append((lref, rref, h))
        return ("n", len(nodes) - 1), h

    import sys
    old = sys.getrecursionlimit()
    sys.setrecursionlimit(max(old, 10000))
    try:
        rec(0, int(n))
    finally:
        sys.setrecursionlimit(old)
    L, K = len(leaves), len(nodes)
    order = sorted(range(K), key=lambda k: (nodes[k][2], k))       # by height, stable
    new_id = {old_k: new_k for new_k, old_k in enumerate(order)}

    def ref_id(ref):
        return ref[1] if ref[0] == "l" else L + new_id[ref[1]]

    node_left = np.array([ref_id(nodes[k][0]) for k in order], dtype=np.int32)
    node_right = np.array([ref_id(nodes[k][1]) for k in order], dtype=np.int32)
    heights = np.array([nodes[k][2] for k in order], dtype=np.int64)
    H = int(heights.max()) if K else 0
    level_off = np.zeros(H + 1, dtype=np.int32)
    for h in range(1, H + 1):
        level_off[h] = int(np.searchsorted(heights, h, side="right"))
    # the root is the node of greatest height and therefore last: the kernel reads L+K-1
    return PairwisePlan(int(n), np.array([s for s, _ in leaves], dtype=np.int32),
                        np.array([m for _, m in leaves], dtype=np.int32),
                        node_left, node_right, level_off)


def evaluate(plan: PairwisePlan, a: np.ndarray) -> float:
    """Host evaluation of the plan (used by the CPU tests to pin the plan to np.sum)."""
    a = np.asarray(a, dtype=np.float64)
    L, K = plan.n_leaves, plan.n_nodes
    vals = np.zeros(L + K)
    for l in range(L):
        s, m = int(plan.leaf_start[l]), int(plan.leaf_len[l])
        seg = a[s:s + m]
        if m < 8:
            r = 0.0
            for v in seg:
                r += v
        else:
            acc = seg[:8].copy()
            full = m - (m % 8)
            for i in range(8, full, 8):
                acc += seg[i:i + 8]
            r = ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]))
            for v in seg[full:]:
                r += v
        vals[l] = r
    for k in range(K):
        vals[L + k] = vals[plan.node_left[k]] + vals[plan.node_right[k]]
    return float(vals[L + K - 1])
