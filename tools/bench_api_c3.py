#!/usr/bin/env python
"""End-to-end timing of the C3 analysis pattern through the public API (BilayerAnalyzer):
8,192-lipid POPC/DOPE/CHOL bilayer, for every ordered resname pair x {upper, lower} one
com_lateral_rdf (80 bins, 40 A) and for every ordered pair one nnf (k = 5) -- what the reference's
prefab drivers register (prefab_analysis_protocols.py:880-890, 1508-1521): 18 RDF + 9 NNF analyses.

    python tools/bench_api_c3.py [--frames 256] [--lipids 8192] [--window W]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(frames, lipids, window, shared):
    import torch
    from pybilt_b200 import BilayerAnalyzer, synthetic as syn
    for k in ("PBK_NO_RDF_MULTI", "PBK_NO_NNF_MULTI"):
        os.environ.pop(k, None)
        if not shared:
            os.environ[k] = "1"
    spec, _ = syn.config_spec("C3")
    if lipids != 8192:
        spec = syn.BilayerSpec(lipids, spec.types, seed=spec.seed)
    import numpy as np
    from pybilt_b200.bilayer_analyzer import FrameWindow
    top, traj = syn.generate(spec, frames)
    x_host = torch.from_numpy(traj.positions).pin_memory()          # pinned, like bench.py's e2e leg
    ba = BilayerAnalyzer(structure=top, trajectory=(FrameWindow(x_host, 0, frames), traj.dimensions,
                                                    np.arange(frames) * spec.dt), selection="all")
    ba.remove_analysis("msd_1")
    names = ["POPC", "DOPE", "CHOL"]
    for a in names:
        for b in names:
            for leaf in ("upper", "lower"):
                ba.add_analysis("com_lateral_rdf rdf_%s_%s_%s resname_1 %s resname_2 %s leaflet %s "
                                "range_outer 40.0 n_bins 80" % (a, b, leaf, a, b, leaf))
            ba.add_analysis("nnf nnf_%s_%s resname_1 %s resname_2 %s" % (a, b, a, b))
    if window:
        ba.window_frames = window
    out = None
    times = []
    for _ in range(3):
        ba.reset()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ba.run_analysis()
        out = {k: ba.get_analysis_data(k) for k in ba.get_analysis_ids()}
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    return min(times), out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=256)
    ap.add_argument("--lipids", type=int, default=8192)
    ap.add_argument("--window", type=int, default=0)
    a = ap.parse_args()
    import numpy as np
    t1, o1 = run(a.frames, a.lipids, a.window, True)
    t0, o0 = run(a.frames, a.lipids, a.window, False)
    same = all(np.array_equal(np.asarray(o1[k][0] if isinstance(o1[k], tuple) else o1[k]),
                              np.asarray(o0[k][0] if isinstance(o0[k], tuple) else o0[k]), equal_nan=True)
               for k in o1)
    print(json.dumps({"workload": "C3 pattern: %d lipids, %d frames, 18 com_lateral_rdf + 9 nnf" % (a.lipids, a.frames),
                      "window": a.window or None,
                      "shared_sweeps_s": t1, "per_analysis_s": t0, "speedup": t0 / t1,
                      "lipid_frames_per_s_shared": a.lipids * a.frames / t1,
                      "outputs_identical": bool(same)}))


if __name__ == "__main__":
    main()
