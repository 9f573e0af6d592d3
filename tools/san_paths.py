#!/usr/bin/env python
"""Small run of every kernel family for compute-sanitizer (tools/sanitize.sh): the streamed reduction
(more than one 32-frame block, so the block hand-over and the chunk snapshots are exercised), the generic
and the fused reduction, cell-list and small-system pair kernels, neighbour searches, grids."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from pybilt_b200 import engine as eng, synthetic as syn

spec = syn.BilayerSpec(300, (syn.POPC.with_fraction(0.5), syn.DOPE.with_fraction(0.3), syn.CHOL.with_fraction(0.2)),
                       seed=31, npt=1e-3, step=3.0)
top, traj = syn.generate(spec, 70)
layout = eng.build_bead_layout(top.masses, top.names, top.resindex, top.resnames, top.resids)
x = torch.from_numpy(traj.positions).cuda()
box = torch.from_numpy(np.ascontiguousarray(traj.dimensions[:, :3])).cuda()
names = sorted(set(layout.types))
codes = torch.tensor([names.index(t) for t in layout.types], dtype=torch.int32, device="cuda")
ref = None
for kw in (dict(allow_fused=False), dict(allow_fused=False, allow_stream=False), dict()):
    e = eng.LipidReductionEngine(top.masses, layout, 70, batch_frames=50, **kw)
    e.push(x, box)
    e.check_status()
    cur = (e.com.clone(), e.com_unwrap.clone(), e.labels.clone())
    if ref is not None:
        assert all(torch.equal(a, b) for a, b in zip(ref, cur))
    ref = cur
edges = np.histogram([-1], bins=25, range=[0.0, 25.0])[1]
for force in ("0", "1"):
    os.environ["PBK_FORCE_CELLS"] = force
    cnt, fc = e.rdf(codes, 0, 1, 3, edges)
    nb, frac = e.nnf(codes, 0, 1, 3, 5)
    gi, gz, knn = e.lipid_grid(16, 16, "opt")
    cells, th, apl = e.grid_reduce(gi, gz, codes, len(names))
idx = torch.arange(layout.n_beads, dtype=torch.int32, device="cuda")
o = torch.zeros(70, dtype=torch.int32, device="cuda")
en = torch.arange(70, dtype=torch.int32, device="cuda")
m = e.msd_pairs(idx, o, en)
torch.cuda.synchronize()
print("san_paths ok", int(cnt.sum()), float(m[-1]))
