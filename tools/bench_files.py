#!/usr/bin/env python
"""End-to-end from trajectory FILES: config C2 (512 lipids x 12 beads, 5,000 frames) plus 20 % extra
atoms standing in for solvent, written as psf + dcd to a scratch directory, then analysed through
BilayerAnalyzer(structure=psf, trajectory=dcd, selection='not resname TIP3') -- native readers,
device-side interleave and atom selection.  The file is in the page cache (just written).

    python tools/bench_files.py [--frames 5000]
"""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=5000)
    a = ap.parse_args()
    import torch
    from pybilt_b200 import BilayerAnalyzer, formats as fmt, synthetic as syn
    spec, _ = syn.config_spec("C2")
    top, traj = syn.generate(spec, a.frames)
    M = top.n_atoms
    n_w = M // 15                                   # 3-atom waters: +20 % atoms
    rng = np.random.default_rng(1)
    water = rng.uniform(0, 100, size=(1, 3 * n_w, 3)).astype(np.float32)
    pos = np.concatenate([traj.positions, np.broadcast_to(water, (a.frames, 3 * n_w, 3))], axis=1)
    tmp = tempfile.mkdtemp(prefix="pbk_files_")
    psf, dcd = os.path.join(tmp, "c2.psf"), os.path.join(tmp, "c2.dcd")
    fmt.write_psf(psf, ["MEMB"] * M + ["SOLV"] * (3 * n_w),
                  np.concatenate([np.asarray(top.resids)[top.resindex], np.repeat(np.arange(1, n_w + 1), 3)]),
                  [top.resnames[r] for r in top.resindex] + ["TIP3"] * (3 * n_w),
                  list(top.names) + ["OH2", "H1", "H2"] * n_w,
                  np.concatenate([top.masses, np.tile([15.9994, 1.008, 1.008], n_w)]))
    fmt.write_dcd(dcd, pos, traj.dimensions, dt_ps=spec.dt)
    del pos

    def make(structure, trajectory, sel):
        ba = BilayerAnalyzer(structure=structure, trajectory=trajectory, selection=sel)
        ba.add_analysis("msd_multi mm n_tau 50 n_sigma 50")
        ba.add_analysis("com_lateral_rdf rdf")
        return ba

    res = {}
    x_host = torch.from_numpy(traj.positions).pin_memory()
    for label, ba in (("files", make(psf, dcd, "not resname TIP3")),
                      ("pinned_arrays", make(top, (x_host, traj.dimensions, traj.times), "all"))):
        ts = []
        for _ in range(4):
            ba.reset()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ba.run_analysis()
            out = [ba.get_analysis_data(k) for k in ("msd_1", "mm", "rdf")]
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        res[label] = (min(ts), out)
    same = (np.array_equal(res["files"][1][0][:, 1], res["pinned_arrays"][1][0][:, 1]) and
            np.array_equal(res["files"][1][2][0], res["pinned_arrays"][1][2][0]))
    print(json.dumps({"workload": "C2 + 20%% solvent atoms from psf/dcd files, %d frames" % a.frames,
                      "file_bytes": os.path.getsize(dcd), "from_files_ms": res["files"][0] * 1e3,
                      "from_pinned_arrays_ms": res["pinned_arrays"][0] * 1e3,
                      "lipid_frames_per_s_from_files": 512 * a.frames / res["files"][0],
                      "outputs_identical": bool(same)}))
    for f in (psf, dcd):
        os.remove(f)
    os.rmdir(tmp)


if __name__ == "__main__":
    main()
