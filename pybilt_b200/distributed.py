"""Frame sharding of the lipid reduction across GPUs.

One process per GPU, ``torch.distributed`` (NCCL over NVLink on the GPU box, gloo in the CPU
tests).  Analysed frames are split into contiguous shards; every rank holds all lipids.
The reference has no distributed path -- its only parallelism is ``multiprocessing.Pool``
over frame ranges (``pybilt/com_trajectory/COMTraj.py:2139-2232``) -- so this module defines
the exchange steps, none of which sits on the per-frame data path:

1. **unwrap across shards.**  The unwrap is a recurrence over frames
   (``pybilt/mda_tools/mda_unwrap_vectorized.py:12-93``).  Rank r reads a one-frame back halo
   (the last frame of rank r-1), unwraps its frames relative to that halo frame (image count
   0 there) and reports the NET image-count change over its owned frames, ``net_r`` int32
   [M,3].  ``image_offsets`` all-gathers the nets and forms the exclusive prefix
   ``O_r = sum_{q<r} net_q`` = the true image counts of rank r's halo frame.  The rank then
   finishes its reduction from those counts with the exact reference recurrence, so every
   unwrapped coordinate is ``f32(f64(x) + s*L)`` with the same integer ``s`` a single process
   would have found.  (The image increments of the first pass are decided from raw
   consecutive frames; a jump within rounding distance of half a box is detected --
   ``PBK_ST_UNWRAP_AMBIGUOUS`` -- and reported instead of silently differing.)
2. **single-origin MSD** needs frame 0's unwrapped COMs and ``msd``/``msd_multi`` need the
   first-frame leaflets: rank 0 broadcasts ``com_unwrap[0]`` ([N,3] f64) and ``labels[0]``.
3. **multi-origin MSD**: a shard also reduces a forward halo of ``n_tau - 1`` frames, so
   every block that starts in the shard ends in it.
4. **per-frame outputs** (msd rows, nnf fractions, grid thickness / APL rows): every rank
   fills its own slots of a zero vector over all analysed frames and the vectors are
   all-reduced (x + 0 = x exactly, so this is a gather that needs no size bookkeeping).
5. **RDF**: partial histograms (int64) and the (N_a, N_b, N) sums are all-reduced.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist


def is_distributed() -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


@dataclass
class Shard:
    """Rank-local view of the analysed frames 0..n_total-1."""
    rank: int
    world: int
    n_total: int
    lo: int                 # first owned analysed frame
    hi: int                 # one past the last owned analysed frame
    back: int               # back-halo frames actually read (0 on rank 0, else 1)
    fwd: int                # forward-halo frames actually read

    @property
    def n_owned(self) -> int:
        return self.hi - self.lo

    @property
    def first_local(self) -> int:
        """global index of local row 0"""
        return self.lo - self.back

    @property
    def n_local(self) -> int:
        return self.back + self.n_owned + self.fwd

    def local_of(self, g: int) -> int:
        return g - self.first_local

    def owned_local(self) -> slice:
        return slice(self.back, self.back + self.n_owned)


def shard_bounds(n_total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous near-equal split: the first ``n_total % world`` ranks own one frame more."""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def make_shard(n_total: int, world: int, rank: int, forward_halo: int = 0, align: int = 1) -> Shard:
    """``align`` > 1 (leaflet update interval): the back halo reaches to the last frame whose index is
    a multiple of ``align`` before the first owned one, so that local row 0 is a frame on which the
    reference rebuilds the leaflets (bilayer_analyzer.py:969-981) and every later frame of the shard
    finds the labels it copies inside the shard."""
    lo, hi = shard_bounds(n_total, world, rank)
    back = (((lo - 1) % max(1, int(align))) + 1) if (rank > 0 and hi > lo) else 0
    fwd = max(0, min(forward_halo, n_total - hi)) if hi > lo else 0
    return Shard(rank, world, n_total, lo, hi, back, fwd)


def net_frames_of(shard: Shard, forward_halo: int = 0, align: int = 1) -> int:
    """Local frames after which a rank reports its image counts: up to and including local row 0 of
    the next rank (its last owned frame when the back halo is one frame)."""
    if shard.rank >= shard.world - 1:
        return shard.back + shard.n_owned
    nxt = make_shard(shard.n_total, shard.world, shard.rank + 1, forward_halo, align)
    if nxt.hi <= nxt.lo:
        return shard.back + shard.n_owned
    return nxt.first_local - shard.first_local + 1


def agree_bits(local_bits: int, device, group=None, width: int = 8) -> int:
    """Bitwise OR over the ranks of a small non-negative integer (a MAX over its bits: NCCL has no
    bitwise reductions).  Every rank gets the same answer, so status-dependent control flow -- raising,
    re-running on another path -- stays in step across the ranks."""
    v = torch.tensor([(int(local_bits) >> b) & 1 for b in range(width)], dtype=torch.int32, device=device)
    w = _wire(v, group)
    dist.all_reduce(w, op=dist.ReduceOp.MAX, group=group)
    return sum(int(x) << b for b, x in enumerate(w.tolist()))


def _wire(t: torch.Tensor, group=None) -> torch.Tensor:
    """Tensor as the backend wants it: NCCL takes device tensors; gloo (CPU tests, or several
    ranks sharing one GPU) gets a host copy."""
    if t.is_cuda and dist.get_backend(group) != "nccl":
        return t.cpu()
    return t.contiguous()


_gather_cache = {}


def image_offsets(net: torch.Tensor, group=None) -> torch.Tensor:
    """Exclusive prefix over ranks of the net image counts (int32 tensor of any shape).
    One all-gather into a reused [world, ...] buffer and one reduction over the ranks before
    this one (nothing is allocated per call after the first)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    w = _wire(net, group)
    key = (w.device, w.dtype, tuple(w.shape), world)
    buf = _gather_cache.get(key)
    if buf is None:
        buf = torch.empty(world * w.numel(), dtype=w.dtype, device=w.device)
        _gather_cache.clear()
        _gather_cache[key] = buf
    dist.all_gather_into_tensor(buf, w.reshape(-1), group=group)     # rank-major concatenation
    off = (buf.view(world, -1)[:rank].sum(dim=0, dtype=w.dtype).view(w.shape) if rank > 0
           else torch.zeros_like(w))
    return off.to(net.device)


def combine_slots(local_full: torch.Tensor, group=None) -> torch.Tensor:
    """All-reduce (sum) of vectors that are zero outside the caller's own slots."""
    out = _wire(local_full, group).clone()
    dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
    return out.to(local_full.device)


def broadcast_from_first(t: torch.Tensor, group=None) -> torch.Tensor:
    """In-place broadcast of rank 0's tensor."""
    w = _wire(t, group)
    dist.broadcast(w, src=0, group=group)
    if w is not t:
        t.copy_(w)
    return t


def msd_multi_owner_blocks(finished: dict, shard: Shard) -> List[int]:
    """Blocks (by index) whose ORIGIN frame is owned by this shard."""
    return [b for b in sorted(finished) if shard.lo <= finished[b][0] < shard.hi]
