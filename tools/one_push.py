"""One warm reduction of config C2 (for ncu captures): python tools/one_push.py [frames] [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pybilt_b200 import synthetic as syn, engine as eng
F = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
spec, _ = syn.config_spec("C2")
top, traj = syn.generate(spec, F)
layout = eng.build_bead_layout(top.masses, top.names, top.resindex, top.resnames, top.resids)
e = eng.LipidReductionEngine(top.masses, layout, F)
x = torch.from_numpy(traj.positions).cuda(); box = torch.from_numpy(np.ascontiguousarray(traj.dimensions[:, :3])).cuda()
for _ in range(reps):
    e.reset(); e.push(x, box)
torch.cuda.synchronize()
e.check_status()
print("ok", float(e.com_unwrap[-1].sum()))
