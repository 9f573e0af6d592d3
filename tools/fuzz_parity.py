#!/usr/bin/env python
"""Randomised differential test: small random bilayers (sizes, compositions, box breathing, step
sizes, areas per lipid), every kernel family (fused / generic reduction; small-system, shared-memory
and cell-list pair kernels; dense and cell-list neighbour searches; grids) against the oracle port.

    python tools/fuzz_parity.py [--cases 60] [--seed 1]

TEST TOOL: imports oracle/ as the checker (like tests/).  Prints one line per failing case and a
summary; exit code 1 if anything differed.
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def type_codes(types):
    names = []
    for t in types:
        if t not in names:
            names.append(t)
    return names, np.array([names.index(t) for t in types], dtype=np.int32)


def one_case(rng, k):
    import torch
    from oracle import port
    from pybilt_b200 import engine as eng, synthetic as syn
    n = int(rng.choice([8, 16, 30, 64, 100, 160, 256, 400]))
    mix = [(syn.POPC_AA.with_fraction(0.5), syn.DOPE_AA.with_fraction(0.3), syn.TLCL2_AA.with_fraction(0.2)),
           (syn.POPC.with_fraction(0.5), syn.DOPE.with_fraction(0.3), syn.CHOL.with_fraction(0.2)),
           (syn.POPC,), (syn.POPC.with_fraction(0.9), syn.CHOL.with_fraction(0.1))][int(rng.integers(0, 4))]
    spec = syn.BilayerSpec(n, mix, seed=int(rng.integers(1, 1 << 30)), area_per_lipid=float(rng.choice([40.0, 64.0, 120.0, 400.0])),
                           step=float(rng.choice([0.3, 0.5, 3.0, 20.0])), npt=float(rng.choice([0.0, 1e-3, 2e-2])),
                           lattice_jitter=float(rng.choice([0.1, 0.35, 0.49])),
                           box_z=float(rng.choice([60.0, 80.0])), leaflet_z=float(rng.choice([8.0, 20.0])))
    F = int(rng.integers(2, 7))
    top, traj = syn.generate(spec, F)
    interval = int(rng.choice([1, 1, 2]))
    rep = {"leaflets": {"update_interval": interval}}
    _o, reps = port.run(top, traj, [], rep, return_reps=True)
    layout = eng.build_bead_layout(top.masses, top.names, top.resindex, top.resnames, top.resids)
    x = torch.from_numpy(traj.positions).cuda()
    box = torch.from_numpy(np.ascontiguousarray(traj.dimensions[:, :3])).cuda()
    names, codes = type_codes(layout.types)
    d_codes = torch.from_numpy(codes).cuda()
    bad = []
    for fused in (True, False):
        e = eng.LipidReductionEngine(top.masses, layout, F, leaflet_update_interval=interval, allow_fused=fused,
                                     batch_frames=int(rng.integers(1, F + 1)))
        e.push(x, box)
        try:
            e.check_status()
        except eng.FusedPathRejected:
            continue
        for nm, a, b in (("com", e.com, reps["com"]), ("com_unwrap", e.com_unwrap, reps["com_unwrap"]),
                         ("labels", e.labels, reps["labels"]), ("mem_com", e.mem_com, reps["mem_com"])):
            if not np.array_equal(a.cpu().numpy(), b):
                bad.append("reduce[fused=%s] %s" % (fused, nm))
    lab = reps["labels"]
    T = len(names)
    for env in ({}, {"PBK_FORCE_CELLS": "1"}, {"PBK_RDF_NOLIST": "1"}, {"PBK_RDF_NOLIST": "1", "PBK_RDF_V2": "1"}):
        for kk in ("PBK_FORCE_CELLS", "PBK_RDF_NOLIST", "PBK_RDF_V2"):
            os.environ.pop(kk, None)
        os.environ.update(env)
        tag = "+".join(env) or "default"
        for _ in range(2):
            ra, rb = int(rng.integers(-1, T)), int(rng.integers(-1, T))
            mask = int(rng.integers(1, 4))
            nb, lo, hi = [(25, 0.0, 25.0), (80, 0.0, 40.0), (30, 2.0, 32.0), (7, 1.0, 9.0)][int(rng.integers(0, 4))]
            leaflet = {3: "both", 1: "upper", 2: "lower"}[mask]
            na = "all" if ra < 0 else names[ra]
            nbn = "all" if rb < 0 else names[rb]
            _g, want = port.rdf(reps["com"], lab, reps["types"], reps["box"], leaflet, na, nbn, nb, lo, hi)
            edges = np.histogram([-1], bins=nb, range=[lo, hi])[1]
            cnt, _fc = e.rdf(d_codes, ra, rb, mask, edges)
            if not np.array_equal(cnt.cpu().numpy(), want):
                bad.append("rdf[%s] %s-%s mask %d bins %d" % (tag, na, nbn, mask, nb))
        if "PBK_RDF_NOLIST" in env:
            continue
        ia, ib = int(rng.integers(0, T)), int(rng.integers(0, T))
        kn = int(rng.choice([1, 5, 6, 24]))
        mask = int(rng.integers(1, 4))
        leaflet = {3: "both", 1: "upper", 2: "lower"}[mask]
        nl = sum((lab[0] == v).sum() for v, m in ((1, 1), (0, 2)) if mask & m)
        if kn < nl:
            res, counts = port.nnf(reps["com"], lab, reps["types"], reps["box"], reps["times"], leaflet, names[ia],
                                   names[ib], kn, return_counts=True)
            n_b, frac = e.nnf(d_codes, ia, ib, mask, kn)
            n_b = n_b.cpu().numpy()
            for f in range(F):
                got = np.concatenate([n_b[f][(lab[f] == 1) & (n_b[f] >= 0)], n_b[f][(lab[f] == 0) & (n_b[f] >= 0)]])
                if not np.array_equal(got, counts[f]):
                    bad.append("nnf[%s] k=%d frame %d" % (tag, kn, f))
                    break
            if not np.allclose(frac.cpu().numpy(), res[:, 1], rtol=1e-14, atol=0):
                bad.append("nnf[%s] frac" % tag)
        if min((lab[f] == 1).sum() for f in range(F)) > 0 and min((lab[f] == 0).sum() for f in range(F)) > 0:
            nx = int(rng.choice([4, 9, 16]))
            gi, gz, _knn = e.lipid_grid(nx, nx, "opt")
            gi = gi.cpu().numpy()
            for f in range(F):
                g = port.lipid_grid_frame(reps["com"][f], reps["com_unwrap"][f], lab[f], reps["box"][f], nx, nx, "opt")
                if not (np.array_equal(gi[f, 0], g["upper"][0]) and np.array_equal(gi[f, 1], g["lower"][0])):
                    bad.append("grid[%s] nx=%d frame %d" % (tag, nx, f))
                    break
    for kk in ("PBK_FORCE_CELLS", "PBK_RDF_NOLIST", "PBK_RDF_V2"):
        os.environ.pop(kk, None)
    # shared sweeps: every table / tally against the per-analysis kernels (both pinned to the port above)
    nb, lo, hi = [(25, 0.0, 25.0), (80, 0.0, 40.0)][int(rng.integers(0, 2))]
    edges = np.histogram([-1], bins=nb, range=[lo, hi])[1]
    r = e.rdf_multi(d_codes, T, edges)
    if r is not None:
        table = r[0].cpu().numpy()
        for leaf in range(2):
            for a_ in range(T):
                for b_ in range(T):
                    cnt, _fc = e.rdf(d_codes, a_, b_, 1 << leaf, edges)
                    if not np.array_equal(table[leaf, a_, b_], cnt.cpu().numpy()):
                        bad.append("rdf_multi table %d %d %d" % (leaf, a_, b_))
    kn = int(rng.choice([1, 5, 6]))
    if kn < min((lab[0] == 1).sum(), (lab[0] == 0).sum()):
        m = e.nnf_multi(d_codes, T, kn)
        if m is not None:
            for a_ in range(T):
                for b_ in range(T):
                    n1, f1 = e.nnf_select(m[0], m[1], d_codes, a_, b_, 3, kn)
                    n2, f2 = e.nnf(d_codes, a_, b_, 3, kn)
                    if not (torch.equal(n1, n2) and torch.equal(f1, f2)):
                        bad.append("nnf_multi %d %d k=%d" % (a_, b_, kn))
    # thickness from leaflet COMs and the hexatic invariant against the port
    if interval == 1 and min((lab[f] == 1).sum() for f in range(F)) > 0 and min((lab[f] == 0).sum() for f in range(F)) > 0:
        want = port.leaflet_distance(reps["com"], lab, reps["times"])[:, 1]
        if not np.array_equal(e.leaflet_distance().cpu().numpy(), want):
            bad.append("leaflet_distance")
        idx = list(np.nonzero(lab[0] == 1)[0])
        if 2 <= len(idx) <= 80:
            want = port.halperin_nelson(reps["com"], reps["box"], reps["times"], idx)[:, 1]
            member = torch.from_numpy((lab[0] == 1).astype(np.uint8)).cuda()
            got = e.halperin_nelson(member).cpu().numpy()
            if not np.allclose(got, want, rtol=1e-9, atol=1e-300):
                bad.append("halperin_nelson")
    desc = "case %d: N=%d F=%d types=%d apl=%g step=%g npt=%g jitter=%g interval=%d" % (
        k, n, F, T, spec.area_per_lipid, spec.step, spec.npt, spec.lattice_jitter, interval)
    return desc, bad


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=60)
    ap.add_argument("--seed", type=int, default=1)
    a = ap.parse_args()
    rng = np.random.default_rng(a.seed)
    n_bad = 0
    for k in range(a.cases):
        try:
            desc, bad = one_case(rng, k)
        except Exception as exc:                      # a crash is a finding too
            desc, bad = "case %d" % k, ["EXCEPTION %r" % (exc,)]
        if bad:
            n_bad += 1
            print("FAIL", desc, "::", "; ".join(bad[:6]), flush=True)
    print("fuzz: %d cases, %d failing" % (a.cases, n_bad))
    sys.exit(1 if n_bad else 0)


if __name__ == "__main__":
    main()
