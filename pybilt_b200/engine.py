"""Device-resident lipid reduction: frames in, per-lipid COMs / leaflets / analyses out.

This is the frame-batched replacement of the body of ``BilayerAnalyzer.__next__``
(pybilt/bilayer_analyzer/bilayer_analyzer.py:901-1031): instead of one Python iteration
per frame, the analysed frames are pushed through libpbk (include/pbk.h) in batches and
the representations every analysis plugin reads -- wrapped and unwrapped lipid COMs,
membrane COM, leaflet labels -- stay resident in HBM for the plugins' kernels.

PyTorch is used for device buffers, streams and copies only; every computation on the hot
path is a libpbk kernel.  No CPU fallback exists: constructing an engine without a CUDA
device raises.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _native
from .pairwise import build_plan


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


@dataclass
class BeadLayout:
    """Which atoms make up each COM bead (lipid): CSR ``seg`` into ``gather``.

    default (``name_dict`` None): one bead per residue over all its atoms, contiguous
    (com_frame.py:163-168); ``name_dict``: atoms named name_dict[resname][0], then [1], ...
    (com_frame.py:63-77); ``multi_bead``: one bead per listed name (com_frame.py:188-231).
    """
    seg: np.ndarray            # int32 [N+1]
    gather: Optional[np.ndarray]   # int32 [T] or None when beads are contiguous atom ranges
    types: List[str]           # resname of every bead
    resids: np.ndarray         # int64 [N]
    bead_mass: np.ndarray      # f64 [N]  = masses[atoms].sum()  (LipidCOM.mass, com_frame.py:95)

    @property
    def n_beads(self) -> int:
        return len(self.types)


def build_bead_layout(masses, names, resindex, resnames, resids, name_dict=None,
                      multi_bead=False) -> BeadLayout:
    masses = np.asarray(masses, dtype=np.float64)
    resindex = np.asarray(resindex, dtype=np.int64)
    n_res = len(resnames)
    counts = np.bincount(resindex, minlength=n_res)
    rseg = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    if name_dict is None:
        seg = rseg.astype(np.int32)
        bead_mass = np.array([masses[rseg[r]:rseg[r + 1]].sum() for r in range(n_res)])
        return BeadLayout(seg, None, list(resnames), np.asarray(resids, dtype=np.int64), bead_mass)
    names = np.asarray(names, dtype=object)
    gather, seg, types, rid = [], [0], [], []
    for r in range(n_res):
        lo, hi = int(rseg[r]), int(rseg[r + 1])
        rn = resnames[r]
        if rn not in name_dict:
            raise RuntimeError("name_dict has no entry for resname '%s'" % rn)
        per_name = [lo + np.nonzero(names[lo:hi] == nm)[0] for nm in name_dict[rn]]
        groups = per_name if multi_bead else [np.concatenate(per_name)]
        for g in groups:
            gather.append(g)
            seg.append(seg[-1] + len(g))
            types.append(rn)
            rid.append(int(resids[r]))
    gather = np.concatenate(gather).astype(np.int32)
    seg = np.asarray(seg, dtype=np.int32)
    bead_mass = np.array([masses[gather[seg[i]:seg[i + 1]]].sum() for i in range(len(types))])
    return BeadLayout(seg, gather, types, np.asarray(rid, dtype=np.int64), bead_mass)


def stream_bead_limit(layout: BeadLayout):
    """Largest number of atoms in a bead when the streamed reduction (pbk_reduce_stream,
    include/pbk.h) can take the layout, else None (an atom in two beads, an empty bead, or a
    bead of more than 800 atoms)."""
    seg = layout.seg.astype(np.int64)
    sizes = np.diff(seg)
    if len(sizes) == 0 or sizes.min() < 1 or sizes.max() > 800:
        return None
    if layout.gather is not None and len(np.unique(layout.gather)) != len(layout.gather):
        return None
    return int(sizes.max())


def rdf_thresholds(edges: np.ndarray) -> np.ndarray:
    """Histogram edges expressed on d^2 so that binning needs no square root.

    ``np.histogram`` puts r = sqrt(d2) (IEEE correctly rounded) into bin k iff
    edges[k] <= r < edges[k+1] (last bin closed).  sqrt is monotone, so
    thr[k] = min{t >= 0 : sqrt(t) >= edges[k]} and thr[n] = min{t : sqrt(t) > edges[n]} give
    thr[k] <= d2 < thr[k+1]  <=>  r in bin k, exactly.
    """
    edges = np.asarray(edges, dtype=np.float64)
    n = len(edges) - 1
    thr = np.zeros(n + 1)

    def least(e, strict):
        ok = (lambda t: np.sqrt(t) > e) if strict else (lambda t: np.sqrt(t) >= e)
        if e < 0 or (e == 0 and not strict):
            return 0.0
        t = np.float64(e) * np.float64(e)
        while not ok(t):
            t = np.nextafter(t, np.inf)
        while t > 0:
            prev = np.nextafter(t, -np.inf)
            if ok(prev):
                t = prev
            else:
                break
        return float(t)

    for k in range(n):
        thr[k] = least(edges[k], False)
    thr[n] = least(edges[n], True)
    return thr


class FusedPathRejected(RuntimeError):
    """The fused small-system path met a jump it cannot decide chunk-locally (or an image
    count outside int8); the caller re-runs the frames with ``allow_fused=False``."""


class LipidReductionEngine:
    """Runs unwrap -> membrane COM -> lipid COM -> leaflets for batches of frames and keeps
    the results for all analysed frames resident on the device."""

    def __init__(self, masses, layout: BeadLayout, n_frames: int, device: int = 0, norm: int = 2,
                 leaflet_update_interval: int = 1, batch_frames: Optional[int] = None,
                 scratch_bytes: int = 4 << 30, allow_fused: bool = True, comtraj_rules: bool = False, rewrap: bool = False,
                 allow_stream: bool = True):
        if not torch.cuda.is_available():
            raise RuntimeError("pybilt_b200: a CUDA device is required (no CPU fallback)")
        self.ctx = _native.Context(device)
        self.dev = torch.device("cuda", device)
        self.masses_h = np.ascontiguousarray(masses, dtype=np.float64)
        self.M = int(self.masses_h.shape[0])
        self.layout = layout
        self.N = layout.n_beads
        self.F = int(n_frames)
        self.norm = int(norm)
        self.update_interval = int(leaflet_update_interval)
        self.total_mass = float(self.masses_h.sum())            # com_frame.py:175
        plan = build_plan(self.M)
        self.plan = plan
        d = self.dev
        self.d_mass = torch.from_numpy(self.masses_h).to(d)
        self.d_seg = torch.from_numpy(layout.seg).to(d)
        self.d_gather = None if layout.gather is None else torch.from_numpy(layout.gather).to(d)
        self.d_bead_mass = torch.from_numpy(np.ascontiguousarray(layout.bead_mass)).to(d)
        self.d_leaf_start = torch.from_numpy(plan.leaf_start).to(d)
        self.d_leaf_len = torch.from_numpy(plan.leaf_len).to(d)
        self.d_node_left = torch.from_numpy(plan.node_left).to(d) if plan.n_nodes else torch.zeros(1, dtype=torch.int32, device=d)
        self.d_node_right = torch.from_numpy(plan.node_right).to(d) if plan.n_nodes else torch.zeros(1, dtype=torch.int32, device=d)
        self.d_level_off = torch.from_numpy(plan.level_off).to(d)
        # streamed reduction (frames that do not fit in shared memory): tiles of lipids
        self.comtraj_rules = bool(comtraj_rules)
        self.rewrap = bool(rewrap)            # com_frame:rewrap (com_frame.py:233-270) after every reduction
        self._orient = None                   # leaflets:assign_method 'orientation' (set_orientation)
        self.stream_plan = None if (self.comtraj_rules or not allow_stream) else stream_bead_limit(layout)
        # per-frame scratch: image counts (6 B/atom) + tree nodes (generic kernels), or tree nodes +
        # compact normal components + tile partials (streamed reduction)
        nodes = plan.n_leaves + plan.n_nodes
        if self.stream_plan is not None:
            per_frame = nodes * 24 + self.N * 8 + self.M // 4 + (3 * self.N // 128 + 1) * 16 + 96
        else:
            per_frame = self.M * 6 + nodes * 24
        if batch_frames is None:
            batch_frames = max(1, min(self.F, scratch_bytes // max(per_frame, 1)))
        self.B = int(batch_frames)
        self._d_img = None
        self._d_nodes = None
        self._stream_scratch = None
        self.d_u_state = torch.empty((self.M, 3), dtype=torch.float32, device=d)
        self.d_status = torch.zeros(1, dtype=torch.int32, device=d)
        # resident representations of every analysed frame
        self.com = torch.empty((self.F, self.N, 3), dtype=torch.float64, device=d)
        self.com_unwrap = torch.empty((self.F, self.N, 3), dtype=torch.float64, device=d)
        self.mem_com = torch.empty((self.F, 3), dtype=torch.float64, device=d)
        self.labels = torch.empty((self.F, self.N), dtype=torch.uint8, device=d)
        self.box = torch.empty((self.F, 3), dtype=torch.float32, device=d)
        self.n_done = 0
        self.row_offset = 0
        self.stream_started = False
        self.launches = 0
        # COMTraj arithmetic (pybilt/com_trajectory/COMTraj.py:1067-1085): scalar float32 unwrap
        # rule and membrane COM accumulated atom by atom; only the generic kernels implement it
        self.allow_fused = bool(allow_fused) and not self.comtraj_rules
        self.fused_supported = True       # cleared when libpbk answers PBK_EUNSUPPORTED
        self.used_fused = False
        self.used_stream = False
        self._fused_scratch = None
        self._rdf_tables = {}

    # scratch of the generic kernels, allocated when they are first used
    @property
    def d_img(self) -> torch.Tensor:
        if self._d_img is None:
            self._d_img = torch.empty((self.B, self.M, 3), dtype=torch.int16, device=self.dev)
        return self._d_img

    @property
    def d_nodes(self) -> torch.Tensor:
        if self._d_nodes is None:
            p = self.plan
            self._d_nodes = torch.empty((self.B, p.n_leaves + p.n_nodes, 3), dtype=torch.float64, device=self.dev)
        return self._d_nodes

    def _stream_scratch_for(self, nb: int) -> torch.Tensor:
        import ctypes
        need = ctypes.c_int64(0)
        p = self.plan
        self.ctx.lib.pbk_reduce_stream_scratch_bytes(nb, self.M, self.N, p.n_leaves, p.n_nodes, ctypes.byref(need))
        if self._stream_scratch is None or self._stream_scratch.numel() < need.value:
            self._stream_scratch = torch.empty(int(need.value), dtype=torch.uint8, device=self.dev)
        return self._stream_scratch

    def _stream_call(self, xb, bb, nb, phase, init_b, net_out, net_frames, g0, prev):
        """pbk_reduce_stream over one batch; False when libpbk has no streamed path for the shape."""
        p = self.plan
        scratch = self._stream_scratch_for(nb)
        ok = self.ctx.try_call(
            "pbk_reduce_stream", xb.data_ptr(), bb.data_ptr(), self.d_mass.data_ptr(), self.d_seg.data_ptr(),
            _ptr(self.d_gather), self.d_bead_mass.data_ptr(), int(self.stream_plan), self.d_leaf_start.data_ptr(),
            self.d_leaf_len.data_ptr(), p.n_leaves, self.d_node_left.data_ptr(), self.d_node_right.data_ptr(),
            p.n_nodes, self.d_level_off.data_ptr(), p.n_levels, self.total_mass, self.d_u_state.data_ptr(),
            0 if self.stream_started else 1, nb, self.M, self.N, self.norm, int(phase), _ptr(init_b), _ptr(net_out),
            int(net_frames), g0 + self.row_offset, self.update_interval, prev, scratch.data_ptr(),
            int(scratch.numel()), self.com[g0:].data_ptr(), self.com_unwrap[g0:].data_ptr(),
            self.mem_com[g0:].data_ptr(), self.labels[g0:].data_ptr(), self.d_status.data_ptr(), self.stream_ptr())
        if ok:
            # k_box_table, k_sweep_p, k_sweep_s, k_sweep_a2 (flag-gated) [+ k_sweep_l, k_mem_tree, k_sweep_bt,
            # k_sweep_b (flag-gated), k_zpart, k_zmean, k_labels, k_leaflets_c]
            self.launches += 12 if (phase & 2) else 4
            self.used_stream = True
        return ok

    def _fused_scratch_for(self, nb: int) -> torch.Tensor:
        import ctypes
        need = ctypes.c_int64(0)
        self.ctx.lib.pbk_reduce_fused_scratch_bytes(nb, self.M, self.ctx.sm_count(), ctypes.byref(need),
                                                    None, None)
        if self._fused_scratch is None or self._fused_scratch.numel() < need.value:
            self._fused_scratch = torch.empty(int(need.value), dtype=torch.uint8, device=self.dev)
        return self._fused_scratch

    # ---------------------------------------------------------------------------------
    def reset(self):
        self.n_done = 0
        self.row_offset = 0             # analysed-frame index of engine row 0 (windowed runs)
        self.stream_started = False     # the next push holds the first frame of the stream
        self.d_status.zero_()

    def rewind(self, keep: int):
        """Start the next window of a streamed run: the last ``keep`` reduced rows move to the
        front of the resident buffers (context for multi-origin MSD blocks and for leaflet
        update intervals), the rest of the buffers is free again.  The unwrap state carries on."""
        n = self.n_done
        keep = int(min(keep, n))
        if keep and n > keep:
            for t in (self.com, self.com_unwrap, self.mem_com, self.labels, self.box):
                t[:keep].copy_(t[n - keep:n].clone())
        self.row_offset += n - keep
        self.n_done = keep

    def unwrap_only(self, x: torch.Tensor, box: torch.Tensor):
        """Advance the unwrap state over ``x`` without reducing the frames (first pass of a
        streamed frame shard: only the image counts at its end are wanted)."""
        assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()
        st = self.stream_ptr()
        for b0 in range(0, int(x.shape[0]), self.B):
            xb, bb = x[b0:b0 + self.B], box[b0:b0 + self.B]
            if self.stream_plan is not None and self._stream_call(xb, bb, int(xb.shape[0]), 1, None, None, 0, 0, None):
                self.stream_started = True
                continue
            self.ctx.call("pbk_unwrap_images", xb.data_ptr(), bb.data_ptr(), self.d_u_state.data_ptr(),
                          0 if self.stream_started else 1, None, 1 if self.comtraj_rules else 0,
                          int(xb.shape[0]), self.M, self.d_img.data_ptr(), self.d_status.data_ptr(), st)
            self.stream_started = True
            self.launches += 1

    def planes_to_aos(self, raw: torch.Tensor, sel: Optional[torch.Tensor], out: torch.Tensor,
                      offsets, byteswap: bool = False):
        """Device-side interleave + atom selection of DCD frames: ``raw`` float32 [f, frame_floats]
        (the frames as they lie in the file) -> ``out`` float32 [f, M, 3]; ``offsets`` = float index
        of the X, Y, Z records inside a frame; ``sel`` int32 [M] or None for the first M atoms."""
        assert raw.is_cuda and raw.dtype == torch.float32 and raw.is_contiguous()
        assert out.is_cuda and out.dtype == torch.float32 and out.is_contiguous()
        f, stride = raw.shape
        st = self.stream_ptr()
        for f0 in range(0, f, 65535):
            f1 = min(f, f0 + 65535)
            self.ctx.call("pbk_planes_to_aos", raw[f0:f1].data_ptr(), _ptr(sel), f1 - f0, int(stride),
                          int(offsets[0]), int(offsets[1]), int(offsets[2]), 1 if byteswap else 0,
                          int(out.shape[1]), out[f0:f1].data_ptr(), st)
            self.launches += 1

    def stream_ptr(self):
        return torch.cuda.current_stream(self.dev).cuda_stream

    def push(self, x: torch.Tensor, box: torch.Tensor, phase: int = 3,
             init_images: Optional[torch.Tensor] = None, net_out: Optional[torch.Tensor] = None,
             net_frames: int = 0):
        """Process the next ``x.shape[0]`` analysed frames.  ``x`` float32 [f, M, 3] and ``box``
        float32 [f, 3] are device tensors (any f; split into scratch-sized batches here).

        Frame shards (pybilt_b200/distributed.py) call this twice for their first push:
        ``phase=1`` only counts periodic images and writes the net change over the frames into
        ``net_out`` (int32 [M, 3]); after the ranks exchanged those, ``phase=2`` finishes the
        same frames with ``init_images`` (int32 [M, 3]), the image counts of frame 0.
        ``net_frames`` (both phases): the net change is taken after that many frames of the push
        (the frames behind it are the shard's forward halo); 0 = after all of them."""
        assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()
        assert box.is_cuda and box.dtype == torch.float32 and box.is_contiguous()
        f_total = int(x.shape[0])
        if self.n_done + f_total > self.F:
            raise RuntimeError("more frames pushed than the engine was sized for")
        if phase != 3 and (f_total > self.B or self.n_done != 0):
            raise RuntimeError("two-phase (sharded) pushes must be the first push and fit one engine batch")
        if init_images is not None and self.stream_started:
            raise RuntimeError("init_images belongs to the first frame of the stream")
        st = self.stream_ptr()
        c = self.ctx
        first_frame = not self.stream_started
        if self._orient is not None and phase != 3:
            raise RuntimeError("orientation leaflets are not combined with frame shards")
        try:
            n_before = self.n_done
            self._push_batches(x, box, phase, init_images, net_out, net_frames, f_total, st, c)
            if self._orient is not None and self.n_done > n_before:
                self._orientation_labels(x, box, n_before, first_frame, st)
        finally:
            if self.rewrap and self.n_done > n_before:
                n = self.n_done - n_before
                c.call("pbk_rewrap", self.com[n_before:].data_ptr(), self.com_unwrap[n_before:].data_ptr(),
                       self.mem_com[n_before:].data_ptr(), self.box[n_before:].data_ptr(), n, self.N,
                       self.d_status.data_ptr(), st)
                self.launches += 1

    def set_orientation(self, sel_atom, sel_mass, seg):
        """Label the leaflets by lipid orientation (bilayer_analyzer.py:849-876) instead of the normal
        position: ``seg[2 i .. 2 i + 2]`` delimits lipid i's tail and head atoms in ``sel_atom`` (indices
        into the pushed atoms) / ``sel_mass``."""
        d = self.dev
        self._orient = (torch.as_tensor(np.asarray(sel_atom, dtype=np.int32), device=d),
                        torch.as_tensor(np.asarray(sel_mass, dtype=np.float64), device=d),
                        torch.as_tensor(np.asarray(seg, dtype=np.int32), device=d),
                        torch.zeros(max(1, len(sel_atom)), dtype=torch.float32, device=d))

    def _orientation_labels(self, x, box, n_before, first_frame, st):
        sel_atom, sel_mass, seg, u_state = self._orient
        n = self.n_done - n_before
        self.ctx.call("pbk_orient_labels", x.data_ptr(), box.data_ptr(), self.mem_com[n_before:].data_ptr(),
                      sel_atom.data_ptr(), sel_mass.data_ptr(), seg.data_ptr(), u_state.data_ptr(),
                      1 if first_frame else 0, n, self.M, self.N, self.norm, n_before + self.row_offset,
                      self.update_interval, self.labels[n_before:].data_ptr(), self.d_status.data_ptr(), st)
        self.launches += 1
        if self.update_interval > 1:
            # frames between rebuilds carry the labels of the last rebuild frame (bilayer_analyzer.py:969-981)
            rows = torch.arange(n_before, self.n_done, device=self.dev)
            g = rows + self.row_offset
            src = g - g % self.update_interval - self.row_offset
            keep = src != rows
            if bool((src[keep] < 0).any()):
                raise RuntimeError("orientation leaflets: the last rebuild frame is not resident")
            self.labels[rows[keep]] = self.labels[src[keep]]

    def _push_batches(self, x, box, phase, init_images, net_out, net_frames, f_total, st, c):
        for b0 in range(0, f_total, self.B):
            b1 = min(f_total, b0 + self.B)
            nb = b1 - b0
            xb, bb = x[b0:b1], box[b0:b1]
            g0 = self.n_done
            init_b = init_images if b0 == 0 else None      # image counts of the stream's first frame
            if phase & 2:
                self.box[g0:g0 + nb].copy_(bb)
            p = self.plan
            prev = self.labels[g0 - 1].data_ptr() if g0 > 0 else None
            if self.allow_fused:
                scratch = self._fused_scratch_for(nb)
                do_labels = 1 if self.update_interval == 1 else 0
                ok = c.try_call("pbk_reduce_fused", xb.data_ptr(), bb.data_ptr(), self.d_mass.data_ptr(),
                                self.d_seg.data_ptr(), _ptr(self.d_gather), self.d_bead_mass.data_ptr(),
                                self.d_leaf_start.data_ptr(), self.d_leaf_len.data_ptr(), p.n_leaves,
                                self.d_node_left.data_ptr(), self.d_node_right.data_ptr(), p.n_nodes,
                                self.d_level_off.data_ptr(), p.n_levels, self.total_mass,
                                self.d_u_state.data_ptr(), 0 if self.stream_started else 1, nb, self.M, self.N,
                                self.norm, do_labels, int(phase), _ptr(init_b), _ptr(net_out), int(net_frames),
                                scratch.data_ptr(), int(scratch.numel()),
                                self.com[g0:].data_ptr(), self.com_unwrap[g0:].data_ptr(),
                                self.mem_com[g0:].data_ptr(), self.labels[g0:].data_ptr(),
                                self.d_status.data_ptr(), st)
                if ok:
                    # k_chunk_box, k_img3 | k_scan3, k_evsort, k_frames4 (two CTAs per SM; one fewer without k_evsort)
                    self.launches += 5 if phase == 3 else (2 if phase == 1 else 3)
                    self.used_fused = True
                    if not (phase & 2):
                        continue
                    if not do_labels:
                        c.call("pbk_leaflets", self.com_unwrap[g0:].data_ptr(), nb, self.N, self.norm,
                               g0 + self.row_offset, self.update_interval, prev, self.labels[g0:].data_ptr(),
                               self.d_status.data_ptr(), st)
                        self.launches += 1
                    self.n_done += nb
                    self.stream_started = True
                    continue
                self.allow_fused = False          # frame does not fit in shared memory
                self.fused_supported = False
            if self.stream_plan is not None:
                if self._stream_call(xb, bb, nb, phase, init_b, net_out if b0 == 0 else None, net_frames, g0, prev):
                    if phase & 2:
                        self.n_done += nb
                        self.stream_started = True
                    continue
                self.stream_plan = None           # shape outside the streamed path: generic kernels from here on
            if phase == 1:
                # generic path: a plain unwrap of the shard; its last row is the net change
                c.call("pbk_unwrap_images", xb.data_ptr(), bb.data_ptr(), self.d_u_state.data_ptr(), 1, None,
                       1 if self.comtraj_rules else 0, nb, self.M, self.d_img.data_ptr(),
                       self.d_status.data_ptr(), st)
                self.launches += 1
                if net_out is not None:
                    last = (net_frames if 0 < net_frames < nb else nb) - 1
                    net_out.copy_(self.d_img[last].to(torch.int32))
                continue
            c.call("pbk_unwrap_images", xb.data_ptr(), bb.data_ptr(), self.d_u_state.data_ptr(),
                   0 if self.stream_started else 1, _ptr(init_b), 1 if self.comtraj_rules else 0, nb, self.M,
                   self.d_img.data_ptr(), self.d_status.data_ptr(), st)
            if self.comtraj_rules:
                c.call("pbk_membrane_com_seq", xb.data_ptr(), bb.data_ptr(), self.d_img.data_ptr(),
                       self.d_mass.data_ptr(), nb, self.M, self.total_mass, self.mem_com[g0:].data_ptr(), st)
            else:
                c.call("pbk_membrane_com", xb.data_ptr(), bb.data_ptr(), self.d_img.data_ptr(),
                       self.d_mass.data_ptr(), nb, self.M, self.d_leaf_start.data_ptr(),
                       self.d_leaf_len.data_ptr(), p.n_leaves, self.d_node_left.data_ptr(),
                       self.d_node_right.data_ptr(), p.n_nodes, self.d_level_off.data_ptr(), p.n_levels,
                       self.total_mass, self.d_nodes.data_ptr(), self.mem_com[g0:].data_ptr(), st)
            c.call("pbk_lipid_com", xb.data_ptr(), bb.data_ptr(), self.d_img.data_ptr(),
                   self.d_mass.data_ptr(), self.mem_com[g0:].data_ptr(), self.d_seg.data_ptr(),
                   _ptr(self.d_gather), self.d_bead_mass.data_ptr(), nb, self.M, self.N,
                   self.com[g0:].data_ptr(), self.com_unwrap[g0:].data_ptr(), st)
            c.call("pbk_leaflets", self.com_unwrap[g0:].data_ptr(), nb, self.N, self.norm, g0 + self.row_offset,
                   self.update_interval, prev, self.labels[g0:].data_ptr(), self.d_status.data_ptr(), st)
            self.launches += 5
            self.n_done += nb
            self.stream_started = True

    def image_counts_last(self, x_last: torch.Tensor, box_last: torch.Tensor) -> torch.Tensor:
        """int32 [M, 3] image counts of the last pushed frame, recovered from the carried
        unwrapped state: u = f32(x + s*L)  =>  s = rint((u - x) / L)."""
        return torch.round((self.d_u_state.double() - x_last.double()) / box_last.double()[None, :]).to(torch.int32)

    def check_status(self, agree: bool = False, extra_bits: int = 0) -> int:
        """Synchronises; raises like the reference does for data-dependent failures.  ``agree``
        (frame shards): the status word is OR-ed over the ranks first, so that every rank raises --
        or carries on -- together and no rank is left alone in a collective; ``extra_bits`` are
        caller-defined bits (>= 32) that take part in the agreement and are returned."""
        s = int(self.d_status.item()) | int(extra_bits)
        if agree:
            from . import distributed as pd
            if pd.is_distributed():
                s = pd.agree_bits(s, self.dev)
        if s & _native.ST_INTERNAL:
            raise RuntimeError("libpbk: a bulk copy did not complete (internal error)")
        if s & (_native.ST_UNWRAP_AMBIGUOUS | _native.ST_FUSED_RANGE):
            raise FusedPathRejected("fused reduction rejected (status %d)" % s)
        if s & _native.ST_LEAFLET_TIE:
            raise KeyError("")      # bilayer_analyzer.py:838-845: pos stays '' -> leaflets['']
        if s & _native.ST_IMAGE_OVERFLOW:
            raise RuntimeError("unwrap: image count outside int16 or non-positive box length")
        return s

    # ---------------------------------------------------------------------------------
    # analyses over the resident representations
    # ---------------------------------------------------------------------------------
    def label_changes(self, frames: Optional[slice] = None) -> torch.Tensor:
        """int32 [frames]: lipids whose leaflet label differs from the previous resident row (0 for
        row 0, which has no predecessor)."""
        sl = frames or slice(0, self.n_done)
        F = sl.stop - sl.start
        out = torch.empty(F, dtype=torch.int32, device=self.dev)
        prev = self.labels[sl.start - 1].data_ptr() if sl.start > 0 else None
        self.ctx.call("pbk_label_changes", self.labels[sl.start:].data_ptr(), prev, F, self.N, out.data_ptr(),
                      self.stream_ptr())
        self.launches += 1
        return out

    def leaflet_distance(self, frames: Optional[slice] = None) -> torch.Tensor:
        """|mean(com_upper) - mean(com_lower)| along the bilayer normal, float64 [frames]."""
        sl = frames or slice(0, self.n_done)
        F = sl.stop - sl.start
        out = torch.empty(F, dtype=torch.float64, device=self.dev)
        self.ctx.call("pbk_leaflet_distance", self.com[sl.start:].data_ptr(), self.labels[sl.start:].data_ptr(),
                      F, self.N, self.norm, out.data_ptr(), self.stream_ptr())
        self.launches += 1
        return out

    def ct_nearest(self, idx_a, idx_b, lateral=(0, 1), exclude_self: bool = False, d1: bool = False,
                   frames: Optional[slice] = None, second: bool = False):
        """Nearest-partner search on the wrapped COMs (pbk_ct_nearest) over the resident rows `frames`:
        (position in idx_b int32 [F, na], distance float64 [F, na][, second smallest distance]).
        d1: distance_euclidean_pbc instead of COMTraj's minimum image."""
        sl = frames or slice(0, self.n_done)
        d, F = self.dev, sl.stop - sl.start
        a = torch.as_tensor(np.asarray(idx_a, dtype=np.int32), device=d)
        b = torch.as_tensor(np.asarray(idx_b, dtype=np.int32), device=d)
        pos = torch.empty((F, a.numel()), dtype=torch.int32, device=d)
        dist = torch.empty((F, a.numel()), dtype=torch.float64, device=d)
        dist2 = torch.empty((F, a.numel()), dtype=torch.float64, device=d) if second else None
        self.ctx.call("pbk_ct_nearest", self.com[sl.start:].data_ptr(), self.box[sl.start:].data_ptr(), a.data_ptr(),
                      a.numel(), b.data_ptr(), b.numel(), int(lateral[0]), int(lateral[1]), int(bool(exclude_self)),
                      int(bool(d1)), F, self.N, pos.data_ptr(), dist.data_ptr(), _ptr(dist2), self.stream_ptr())
        self.launches += 1
        return (pos, dist, dist2) if second else (pos, dist)

    def ct_adjacency(self, idx, dist: float, lateral=(0, 1), d1: bool = False, frames: Optional[slice] = None) -> torch.Tensor:
        """uint32-packed adjacency rows [F, n, ceil(n/32)] of the lipids `idx` under a lateral cut-off
        (pbk_ct_adjacency): COMTraj's cluster criterion, or (d1) distance_cutoff_clustering's matrix."""
        sl = frames or slice(0, self.n_done)
        d, F = self.dev, sl.stop - sl.start
        a = torch.as_tensor(np.asarray(idx, dtype=np.int32), device=d)
        n = a.numel()
        bits = torch.zeros((F, n, (n + 31) // 32), dtype=torch.int32, device=d)
        self.ctx.call("pbk_ct_adjacency", self.com[sl.start:].data_ptr(), self.box[sl.start:].data_ptr(), a.data_ptr(), n,
                      int(lateral[0]), int(lateral[1]), float(dist), int(bool(d1)), F, self.N, bits.data_ptr(),
                      self.stream_ptr())
        self.launches += 1
        return bits

    def msd_pairs(self, idx: np.ndarray, origins: Sequence[int], ends: Sequence[int],
                  lateral=(0, 1), lengths: bool = False) -> torch.Tensor:
        """mean squared displacement per (origin, end) pair; ``lengths``: mean displacement length."""
        d = self.dev
        d_idx, d_o, d_e = (self._i32(v) for v in (idx, origins, ends))
        out = torch.empty(len(d_o), dtype=torch.float64, device=d)
        self.ctx.call("pbk_ald_pairs" if lengths else "pbk_msd_pairs", self.com_unwrap.data_ptr(), self.n_done, self.N,
                      d_idx.data_ptr(), int(d_idx.numel()), int(lateral[0]), int(lateral[1]),
                      d_o.data_ptr(), d_e.data_ptr(), int(d_o.numel()), out.data_ptr(), self.stream_ptr())
        self.launches += 1
        return out

    def _i32(self, v) -> torch.Tensor:
        if torch.is_tensor(v):
            assert v.is_cuda and v.dtype == torch.int32
            return v
        return torch.as_tensor(np.ascontiguousarray(v, dtype=np.int32), device=self.dev)

    def rdf(self, type_code: torch.Tensor, type_a: int, type_b: int, leaflet_mask: int,
            edges, lateral=(0, 1), frames: Optional[slice] = None, edges_dev=None, thr_dev=None,
            dense=False):
        """``edges``: the float64 np.histogram edges (host); ``edges_dev`` / ``thr_dev``
        optional device copies of the edges and of ``rdf_thresholds(edges)`` so that no
        host->device copy happens inside the call.  ``dense`` forces the O(N^2) sqrt kernel."""
        d = self.dev
        sl = frames or slice(0, self.n_done)
        F = sl.stop - sl.start
        nb = len(edges) - 1
        if edges_dev is None or (thr_dev is None and not dense):
            # bin edges and their d^2 thresholds are pure functions of the edges: kept per engine
            key = np.ascontiguousarray(edges, dtype=np.float64).tobytes()
            cached = self._rdf_tables.get(key)
            if cached is None:
                cached = (torch.as_tensor(np.ascontiguousarray(edges, dtype=np.float64), device=d),
                          torch.as_tensor(rdf_thresholds(edges), device=d))
                self._rdf_tables[key] = cached
            edges_dev = edges_dev if edges_dev is not None else cached[0]
            thr_dev = thr_dev if thr_dev is not None else cached[1]
        d_edges = edges_dev
        if dense:
            thr_dev = None
        count = torch.zeros(nb, dtype=torch.int64, device=d)
        fc = torch.empty((F, 2, 3), dtype=torch.int32, device=d)
        self.ctx.call("pbk_rdf", self.com[sl.start:].data_ptr(), self.labels[sl.start:].data_ptr(),
                      type_code.data_ptr(), self.box[sl.start:].data_ptr(), F, self.N,
                      int(lateral[0]), int(lateral[1]), int(type_a), int(type_b), int(leaflet_mask),
                      nb, float(edges[0]), float(edges[-1]), d_edges.data_ptr(), _ptr(thr_dev),
                      count.data_ptr(), fc.data_ptr(), self.stream_ptr())
        self.launches += 2 if (self.N <= 512 and not dense) else 1      # (k_rdf5_lut + k_rdf5 for small systems)
        return count, fc

    def rdf_multi(self, type_code: torch.Tensor, n_types: int, edges, lateral=(0, 1),
                  frames: Optional[slice] = None):
        """One pair sweep for every ordered type pair and leaflet: (count int64 [2, T, T, n_bins],
        type_counts int32 [F, 2, T]), or None when libpbk has no table-per-pair path for this shape
        (the caller then asks per analysis)."""
        d = self.dev
        sl = frames or slice(0, self.n_done)
        F = sl.stop - sl.start
        nb = len(edges) - 1
        key = np.ascontiguousarray(edges, dtype=np.float64).tobytes()
        cached = self._rdf_tables.get(key)
        if cached is None:
            cached = (torch.as_tensor(np.ascontiguousarray(edges, dtype=np.float64), device=d),
                      torch.as_tensor(rdf_thresholds(edges), device=d))
            self._rdf_tables[key] = cached
        count = torch.zeros((2, n_types, n_types, nb), dtype=torch.int64, device=d)
        tc = torch.empty((F, 2, n_types), dtype=torch.int32, device=d)
        ok = self.ctx.try_call("pbk_rdf_multi", self.com[sl.start:].data_ptr(), self.labels[sl.start:].data_ptr(),
                               type_code.data_ptr(), self.box[sl.start:].data_ptr(), F, self.N, int(lateral[0]),
                               int(lateral[1]), int(n_types), nb, float(edges[0]), float(edges[-1]),
                               cached[0].data_ptr(), cached[1].data_ptr(), count.data_ptr(), tc.data_ptr(),
                               self.stream_ptr())
        if not ok:
            return None
        self.launches += 5
        return count, tc

    def nnf(self, type_code: torch.Tensor, type_a: int, type_b: int, leaflet_mask: int, k: int,
            lateral=(0, 1), frames: Optional[slice] = None):
        d = self.dev
        sl = frames or slice(0, self.n_done)
        F = sl.stop - sl.start
        s0 = sl.start
        n_b = torch.empty((F, self.N), dtype=torch.int32, device=d)
        frac = torch.empty(F, dtype=torch.float64, device=d)
        self.ctx.call("pbk_nnf", self.com[s0:].data_ptr(), self.labels[s0:].data_ptr(), type_code.data_ptr(),
                      self.box[s0:].data_ptr(), F, self.N, int(lateral[0]), int(lateral[1]), int(type_a),
                      int(type_b), int(leaflet_mask), int(k), n_b.data_ptr(), frac.data_ptr(),
                      self.stream_ptr())
        self.launches += 2
        return n_b, frac

    def nnf_multi(self, type_code: torch.Tensor, n_types: int, k: int, lateral=(0, 1),
                  frames: Optional[slice] = None):
        """k nearest leaflet members of every lipid, found once: (type_hits int32 [F, N, T],
        type_counts int32 [F, 2, T]) device tensors for nnf_select, or None (no shared path)."""
        d = self.dev
        sl = frames or slice(0, self.n_done)
        F = sl.stop - sl.start
        s0 = sl.start
        hits = torch.empty((F, self.N, n_types), dtype=torch.int32, device=d)
        tc = torch.empty((F, 2, n_types), dtype=torch.int32, device=d)
        ok = self.ctx.try_call("pbk_nnf_multi", self.com[s0:].data_ptr(), self.labels[s0:].data_ptr(),
                               type_code.data_ptr(), self.box[s0:].data_ptr(), F, self.N, int(lateral[0]),
                               int(lateral[1]), int(n_types), int(k), hits.data_ptr(), tc.data_ptr(),
                               self.stream_ptr())
        if not ok:
            return None
        self.launches += 4
        return hits, tc

    def nnf_select(self, hits: torch.Tensor, tc: torch.Tensor, type_code: torch.Tensor, type_a: int,
                   type_b: int, leaflet_mask: int, k: int, frames: Optional[slice] = None):
        """(n_b, frac) of one nnf analysis from the shared tallies of nnf_multi."""
        sl = frames or slice(0, self.n_done)
        F = sl.stop - sl.start
        n_b = torch.empty((F, self.N), dtype=torch.int32, device=self.dev)
        frac = torch.empty(F, dtype=torch.float64, device=self.dev)
        self.ctx.call("pbk_nnf_select", hits.data_ptr(), tc.data_ptr(), self.labels[sl.start:].data_ptr(),
                      type_code.data_ptr(), F, self.N, int(hits.shape[2]), int(type_a), int(type_b),
                      int(leaflet_mask), int(k), n_b.data_ptr(), frac.data_ptr(), self.stream_ptr())
        self.launches += 2
        return n_b, frac

    def nnf_select_many(self, hits: torch.Tensor, tc: torch.Tensor, type_code: torch.Tensor, specs, k: int,
                        frames: Optional[slice] = None):
        """(n_b [A, F, N], frac [A, F]) of the analyses ``specs`` = [(type_a, type_b, leaflet_mask)]
        from the shared tallies of nnf_multi, in two launches for all of them."""
        sl = frames or slice(0, self.n_done)
        F = sl.stop - sl.start
        A = len(specs)
        d_specs = torch.as_tensor(np.asarray(specs, dtype=np.int32).reshape(A, 3), device=self.dev)
        n_b = torch.empty((A, F, self.N), dtype=torch.int32, device=self.dev)
        frac = torch.empty((A, F), dtype=torch.float64, device=self.dev)
        self.ctx.call("pbk_nnf_select_many", hits.data_ptr(), tc.data_ptr(), self.labels[sl.start:].data_ptr(),
                      type_code.data_ptr(), F, self.N, int(hits.shape[2]), A, d_specs.data_ptr(), int(k),
                      n_b.data_ptr(), frac.data_ptr(), self.stream_ptr())
        self.launches += 2
        return n_b, frac

    def halperin_nelson(self, member: torch.Tensor, lateral=(0, 1), frames: Optional[slice] = None):
        """Mean hexatic invariant of the lipids ``member`` (uint8 [N], 1 = included) per frame,
        float64 [frames]: six nearest members from the 24-neighbour search, then the invariant."""
        sl = frames or slice(0, self.n_done)
        F = sl.stop - sl.start
        s0 = sl.start
        cls = member.to(torch.uint8).unsqueeze(0).expand(F, self.N).contiguous()     # class byte per frame
        knn = torch.empty((F, self.N, 24), dtype=torch.int32, device=self.dev)
        out = torch.empty(F, dtype=torch.float64, device=self.dev)
        st = self.stream_ptr()
        self.ctx.call("pbk_knn24", self.com[s0:].data_ptr(), cls.data_ptr(), self.box[s0:].data_ptr(), F, self.N,
                      int(lateral[0]), int(lateral[1]), knn.data_ptr(), st)
        self.ctx.call("pbk_halperin_nelson", self.com[s0:].data_ptr(), member.data_ptr(), self.box[s0:].data_ptr(),
                      knn.data_ptr(), F, self.N, int(lateral[0]), int(lateral[1]), int(member.sum().item()),
                      out.data_ptr(), st)
        self.launches += 2
        return out

    def lipid_grid(self, nx: int, ny: int, mode: str = "opt", lateral=(0, 1),
                   frames: Optional[slice] = None):
        d = self.dev
        sl = frames or slice(0, self.n_done)
        F = sl.stop - sl.start
        gi = torch.empty((F, 2, nx, ny), dtype=torch.int32, device=d)
        gz = torch.empty((F, 2, nx, ny), dtype=torch.float64, device=d)
        m = 1 if mode == "opt" else 0
        knn = torch.empty((F, self.N, 24), dtype=torch.int32, device=d) if m == 1 else None
        norm = [i for i in (0, 1, 2) if i not in tuple(lateral)][0]
        s = sl.start
        self.ctx.call("pbk_lipid_grid", self.com[s:].data_ptr(), self.com_unwrap[s:].data_ptr(),
                      self.labels[s:].data_ptr(), self.box[s:].data_ptr(), F, self.N, int(lateral[0]),
                      int(lateral[1]), norm, int(nx), int(ny), m, _ptr(knn), gi.data_ptr(),
                      gz.data_ptr(), self.stream_ptr())
        self.launches += 2 if m == 1 else 1
        return gi, gz, knn

    def grid_reduce(self, gi: torch.Tensor, gz: torch.Tensor, type_code: torch.Tensor, n_types: int,
                    lateral=(0, 1), frames: Optional[slice] = None):
        d = self.dev
        sl = frames or slice(0, self.n_done)
        F = sl.stop - sl.start
        nx, ny = int(gi.shape[2]), int(gi.shape[3])
        cells = torch.empty((F, self.N), dtype=torch.int32, device=d)
        thick = torch.empty((F, 2), dtype=torch.float64, device=d)
        apl = torch.empty((F, 1 + 2 * n_types), dtype=torch.float64, device=d)
        s = sl.start
        self.ctx.call("pbk_grid_reduce", gi.data_ptr(), gz.data_ptr(), self.labels[s:].data_ptr(),
                      type_code.data_ptr(), self.box[s:].data_ptr(), F, self.N, int(lateral[0]),
                      int(lateral[1]), nx, ny, int(n_types), cells.data_ptr(), thick.data_ptr(),
                      apl.data_ptr(), self.stream_ptr())
        self.launches += 2
        return cells, thick, apl
