#!/usr/bin/env python
"""Stage timings of the hot path on the OTHER named shapes of BASELINE.json (C3, C4, C5) at a
bounded number of frames (bench.py's line is C2).  Frames are resident in HBM; CUDA events.

    python tools/bench_configs.py [--configs C3,C4,C5] [--frames 32,8,16] [--reps 3]

Prints one JSON line per config: lipid-frames/s and algorithmic GB/s per stage.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timed(fn, reps):
    import torch
    fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(reps):
        out = fn()
    ev[1].record()
    torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / reps, out


def type_codes(types):
    names = []
    for t in types:
        if t not in names:
            names.append(t)
    return names, np.array([names.index(t) for t in types], dtype=np.int32)


def run(cfg: str, F: int, reps: int, dense_check: bool, tile: int = 1):
    print(json.dumps(measure(cfg, F, reps, dense_check, tile)), flush=True)


def measure(cfg: str, F: int, reps: int, dense_check: bool = False, tile: int = 1, peak_gbs: float = None):
    """Stage timings of one named shape (frames resident in HBM, CUDA events); see the module docstring."""
    import torch
    from pybilt_b200 import engine as eng, synthetic as syn
    spec, f_full = syn.config_spec(cfg)
    top, traj = syn.generate(spec, F)
    layout = eng.build_bead_layout(top.masses, top.names, top.resindex, top.resnames, top.resids)
    N, M = layout.n_beads, top.n_atoms
    A = M / N
    x = torch.from_numpy(traj.positions).cuda()
    box = torch.from_numpy(np.ascontiguousarray(traj.dimensions[:, :3])).cuda()
    if tile > 1:
        # the generated frames repeated along time: same per-frame work, enough frames to fill the
        # GPU for the kernels that run one CTA (or thread) per frame
        x = x.repeat(tile, 1, 1).contiguous()
        box = box.repeat(tile, 1).contiguous()
        F = F * tile
    e = eng.LipidReductionEngine(top.masses, layout, F, batch_frames=F)
    names, codes_h = type_codes(layout.types)
    codes = torch.from_numpy(codes_h).cuda()
    out = {"config": cfg, "lipids": N, "atoms": M, "frames": F, "frames_full": f_full, "stages": {}}

    def rec(name, ms, bytes_per_lf, extra=None):
        lf = N * F
        d = {"ms": round(ms, 4), "lipid_frames_per_s": lf / (ms * 1e-3),
             "GBps_algorithmic": bytes_per_lf * lf / (ms * 1e-3) / 1e9,
             "bytes_per_lipid_frame": bytes_per_lf,
             "full_config_s": ms * 1e-3 * f_full / F}
        if peak_gbs:
            d["frac"] = d["GBps_algorithmic"] / peak_gbs
        if extra:
            d.update(extra)
        out["stages"][name] = d

    def reduce():
        e.reset()
        e.push(x, box)
    ms, _ = timed(reduce, reps)
    e.check_status()
    rec("reduce", ms, 12 * A + 49, {"fused": bool(e.used_fused), "streamed": bool(e.used_stream)})

    lab0 = e.labels[0].cpu().numpy()
    idx = np.concatenate([np.nonzero(lab0 == 1)[0], np.nonzero(lab0 == 0)[0]]).astype(np.int32)
    idx_d = torch.from_numpy(idx).cuda()
    o = torch.zeros(F, dtype=torch.int32, device="cuda")
    en = torch.arange(F, dtype=torch.int32, device="cuda")
    ms, _ = timed(lambda: e.msd_pairs(idx_d, o, en), reps)
    rec("msd", ms, 16)

    if cfg in ("C3", "C5"):
        nb, ro = (80, 40.0) if cfg == "C3" else (25, 25.0)
        edges = np.histogram([-1], bins=nb, range=[0.0, ro])[1]
        if cfg == "C3":
            ta = tb = names.index("POPC")
        else:
            ta = tb = -1
        ms, (cnt, fc) = timed(lambda: e.rdf(codes, ta, tb, 3, edges), reps)
        rec("rdf(%d bins, %g A, both leaflets)" % (nb, ro), ms, 18, {"pairs_counted": int(cnt.sum())})
        if dense_check:
            cnt2, _ = e.rdf(codes, ta, tb, 3, edges, dense=True)
            out["rdf_equals_dense"] = bool(torch.equal(cnt, cnt2))
    if cfg == "C3":
        ta, tb = names.index("POPC"), names.index("DOPE")
        ms, (n_b, frac) = timed(lambda: e.nnf(codes, ta, tb, 3, 5), reps)
        rec("nnf(k=5, both leaflets)", ms, 18)
    if cfg == "C4":
        ms, (gi, gz, knn) = timed(lambda: e.lipid_grid(128, 128, "opt"), reps)
        rec("lipid_grid(opt,128x128)", ms, 25 + 2 * 128 * 128 * 12 / N)
        ms, _ = timed(lambda: e.grid_reduce(gi, gz, codes, len(names)), reps)
        rec("grid_reduce", ms, 2 * 128 * 128 * 12 / N)
    tot = sum(s["ms"] for s in out["stages"].values())
    out["total_ms"] = tot
    out["lipid_frames_per_s"] = N * F / (tot * 1e-3)
    del e
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="C3,C4,C5")
    ap.add_argument("--frames", default="256,32,48")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--dense-check", action="store_true")
    ap.add_argument("--tile", default="1,1,1", help="repeat the generated frames this many times per config")
    a = ap.parse_args()
    tiles = a.tile.split(",")
    for i, (cfg, f) in enumerate(zip(a.configs.split(","), a.frames.split(","))):
        run(cfg, int(f), a.reps, a.dense_check, int(tiles[i]) if i < len(tiles) else 1)


if __name__ == "__main__":
    main()
