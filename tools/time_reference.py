#!/usr/bin/env python
"""Wall-clock of the LITERAL reference (unmodified /root/reference, executed through oracle/ref_harness)
on the CPU-runnable cases: the 600-lipid stand-in for sample_bilayer (C1') and a prefix of config C2.
Build container only (the reference tree does not exist on the GPU box).

    python tools/time_reference.py > profiles/r02_reference_cpu.json
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_harness as rh          # noqa: E402
from pybilt_b200 import synthetic as syn      # noqa: E402

out = {"cpu": os.popen("grep -m1 'model name' /proc/cpuinfo").read().split(":")[-1].strip(), "cases": []}
for cfg, frames, analyses in (
        ("C1", 10, ["msd msd_1", "com_lateral_rdf rdf resname_1 POPC resname_2 POPC leaflet upper n_bins 80 range_outer 40.0",
                    "nnf nnf_a resname_1 DOPE resname_2 POPC leaflet upper n_neighbors 6"]),
        ("C2", 40, ["msd msd_1", "msd_multi mm n_tau 10 n_sigma 10", "com_lateral_rdf rdf"])):
    spec, _ = syn.config_spec(cfg)
    top, traj = syn.generate(spec, frames)
    t0 = time.perf_counter()
    rh.run_reference(top, traj, analyses=analyses)
    dt = time.perf_counter() - t0
    out["cases"].append({"config": cfg, "lipids": spec.n_lipids, "atoms": top.n_atoms, "frames": frames, "analyses": analyses,
                         "seconds": dt, "lipid_frames_per_s": spec.n_lipids * frames / dt, "cores": 1,
                         "note": "pybilt.bilayer_analyzer.BilayerAnalyzer.run_analysis(), single process, fake in-memory MDAnalysis"})
print(json.dumps(out, indent=1))
