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


class PeerComm:
    """NVLink peer-memory exchanges of libpbk (include/pbk.h: pbk_comm_*): every rank maps the
    others' exchange windows through CUDA IPC once; a collective is then ONE kernel that pushes the
    contribution into the peers' windows, releases an epoch flag and combines the slots -- no NCCL
    launch on the critical path.  One node, one GPU per rank (NCCL backend)."""

    def __init__(self, device: torch.device, slot_bytes: int, group=None):
        import ctypes
        from . import _native
        self.dev = device
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.slot_bytes = int(slot_bytes)
        self.ctx = _native.Context(device.index)
        self.lib = self.ctx.lib
        h = ctypes.c_void_p()
        rc = self.lib.pbk_comm_create(self.ctx._h, self.rank, self.world, self.slot_bytes, ctypes.byref(h))
        if rc != 0:
            raise RuntimeError("pbk_comm_create failed (%d): %s" % (rc, (self.lib.pbk_last_error(self.ctx._h) or b"").decode()))
        self._h = h
        mine = (ctypes.c_ubyte * 64)()
        if self.lib.pbk_comm_handle(self._h, mine) != 0:
            raise RuntimeError("pbk_comm_handle failed: %s" % (self.lib.pbk_last_error(self.ctx._h) or b"").decode())
        t = torch.tensor(list(mine), dtype=torch.uint8, device=device)
        allh = torch.empty(self.world * 64, dtype=torch.uint8, device=device)
        dist.all_gather_into_tensor(allh, t, group=group)
        buf = (ctypes.c_ubyte * (64 * self.world)).from_buffer_copy(bytes(allh.cpu().numpy().tobytes()))
        ok = self.lib.pbk_comm_connect(self._h, buf) == 0
        # all ranks or none: a rank without peer access must not leave the others in a peer collective
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if int(flag.item()) != 1:
            self.close()
            raise RuntimeError("pbk_comm_connect: no CUDA IPC peer access between the ranks")
        self.status = torch.zeros(1, dtype=torch.int32, device=device)

    def _run(self, name, src: torch.Tensor, out: torch.Tensor):
        assert src.is_cuda and src.is_contiguous() and out.is_contiguous()
        st = torch.cuda.current_stream(self.dev).cuda_stream
        rc = getattr(self.lib, name)(self._h, src.data_ptr(), out.data_ptr(), int(src.numel()), self.status.data_ptr(), st)
        if rc != 0:
            raise RuntimeError("%s failed (%d): %s" % (name, rc, (self.lib.pbk_last_error(self.ctx._h) or b"").decode()))
        return out

    def prefix_i32(self, net: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Sum of ``net`` (int32) over the ranks before this one."""
        return self._run("pbk_comm_prefix_i32", net, out if out is not None else torch.empty_like(net))

    def allreduce_(self, t: torch.Tensor) -> torch.Tensor:
        """Sum over all ranks, in rank order on every rank; int64 / uint64 / float64, in place."""
        name = "pbk_allreduce_f64" if t.dtype == torch.float64 else "pbk_allreduce_u64"
        assert t.dtype in (torch.float64, torch.int64)
        return self._run(name, t.clone(), t)

    def check(self):
        if int(self.status.item()) != 0:
            raise RuntimeError("libpbk peer exchange: a rank never published its contribution")

    def close(self):
        if getattr(self, "_h", None):
            self.lib.pbk_comm_destroy(self._h)
            self._h = None


_peer = {}


def peer_comm(device: torch.device, nbytes: int, group=None, lane: str = "main") -> Optional["PeerComm"]:
    """The process-wide PeerComm for `device` when the group runs over NCCL with one GPU per rank
    (None otherwise, or with PBK_NO_PEER=1: the callers then use NCCL / gloo collectives).
    Collective: every rank must call it with the same arguments the first time.  Collectives of one
    PeerComm must be issued on ONE stream; a second stream takes its own instance (``lane``)."""
    import os
    if os.environ.get("PBK_NO_PEER") or not is_distributed() or dist.get_backend(group) != "nccl":
        return None
    if dist.get_world_size(group) > 16:
        return None
    key = (device.index, dist.get_world_size(group), lane)
    pc = _peer.get(key)
    if pc is not None and pc.slot_bytes < nbytes:
        pc = None
    if pc is None:
        try:
            pc = PeerComm(device, max(int(nbytes), 1 << 20), group)
        except RuntimeError:
            pc = False                          # no peer access on this box: NCCL from here on
        _peer[key] = pc
    return pc or None


def image_offsets(net: torch.Tensor, group=None) -> torch.Tensor:
    """Exclusive prefix over ranks of the net image counts (int32 tensor of any shape).
    Over NVLink peer memory in one libpbk kernel when every rank has its own GPU (PeerComm);
    otherwise one all-gather into a reused [world, ...] buffer and one reduction over the ranks
    before this one (nothing is allocated per call after the first)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if net.is_cuda and net.dtype == torch.int32:
        pc = peer_comm(net.device, net.numel() * 4, group)
        if pc is not None:
            return pc.prefix_i32(net.contiguous())
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
