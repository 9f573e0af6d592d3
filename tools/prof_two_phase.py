#!/usr/bin/env python
"""Cost of the two-call (frame shard) form of the fused reduction on one GPU: push(phase=3) against
push(phase=1) + push(phase=2) with zero image offsets, same frames (config C2).  CUDA events."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from pybilt_b200 import engine as eng, synthetic as syn
F = int(sys.argv[1]) if len(sys.argv) > 1 else 5050
spec, _ = syn.config_spec("C2")
top, traj = syn.generate(spec, F)
layout = eng.build_bead_layout(top.masses, top.names, top.resindex, top.resnames, top.resids)
e = eng.LipidReductionEngine(top.masses, layout, F + 1, batch_frames=F)
x = torch.from_numpy(traj.positions).cuda()
box = torch.from_numpy(np.ascontiguousarray(traj.dimensions[:, :3])).cuda()
net = torch.zeros((top.n_atoms, 3), dtype=torch.int32, device="cuda")
off = torch.zeros((top.n_atoms, 3), dtype=torch.int32, device="cuda")
n_a = F - 49


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def one():
    e.reset(); e.push(x, box)


def two():
    e.reset()
    e.push(x, box, phase=1, net_out=net, net_frames=n_a)
    e.push(x, box, phase=2, init_images=off, net_frames=n_a)


def two_plain():
    e.reset()
    e.push(x, box, phase=1, net_out=net, net_frames=0)
    e.push(x, box, phase=2, init_images=off, net_frames=0)


print("frames", F, "one call %.4f ms   two calls (split tiling) %.4f ms   two calls (plain tiling) %.4f ms"
      % (timed(one), timed(two), timed(two_plain)))
