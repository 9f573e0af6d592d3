#!/usr/bin/env python
"""Phase cycle counters of k_sweep_a (PBK_SWEEP_PROF=1): python tools/prof_sweep.py C4 32 8"""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from pybilt_b200 import engine as eng, synthetic as syn, _native
cfg, F, tile = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
spec, _ = syn.config_spec(cfg)
top, traj = syn.generate(spec, F)
layout = eng.build_bead_layout(top.masses, top.names, top.resindex, top.resnames, top.resids)
x = torch.from_numpy(traj.positions).cuda().repeat(tile, 1, 1).contiguous()
box = torch.from_numpy(np.ascontiguousarray(traj.dimensions[:, :3])).cuda().repeat(tile, 1).contiguous()
F *= tile
e = eng.LipidReductionEngine(top.masses, layout, F, batch_frames=F)
lib = ctypes.CDLL(os.path.join(ROOT, "pybilt_b200", "_pbk.so"))
out = (ctypes.c_ulonglong * 16)()
for it in range(2):
    e.reset(); e.push(x, box); torch.cuda.synchronize()
    lib.pbk_debug_sweep_prof(out)
names = ["task start", "state init", "reduce", "wait frame", "column", "fence", "barrier", "prefetch"]
L = e.plan.n_leaves
for base, who in ((0, "warp0 lane0"), (8, "warp3 lane0")):
    tot = sum(out[base + i] for i in range(8))
    print(who, "total cycles per leaf-frame: %.0f" % (tot / (L * F)))
    for i, n in enumerate(names):
        print("   %-12s %8.1f" % (n, out[base + i] / (L * F)))
