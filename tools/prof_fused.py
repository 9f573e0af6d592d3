"""Per-phase cycle counts of k_fused_frames (debug aid): PBK_FUSED_PROF=1 python tools/prof_fused.py"""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["PBK_FUSED_PROF"] = "1"
import numpy as np, torch
from pybilt_b200 import synthetic as syn, engine as eng
spec, _ = syn.config_spec("C2")
F = 5000
top, traj = syn.generate(spec, F)
layout = eng.build_bead_layout(top.masses, top.names, top.resindex, top.resnames, top.resids)
e = eng.LipidReductionEngine(top.masses, layout, F)
need = ctypes.c_int64(0); nch = ctypes.c_int32(0); ch = ctypes.c_int32(0)
e.ctx.lib.pbk_reduce_fused_scratch_bytes(F, e.M, e.ctx.sm_count(), ctypes.byref(need), ctypes.byref(nch), ctypes.byref(ch))
extra = 64 * nch.value + 4096
e._fused_scratch = torch.zeros(need.value + extra, dtype=torch.uint8, device="cuda")
x = torch.from_numpy(traj.positions).cuda(); box = torch.from_numpy(np.ascontiguousarray(traj.dimensions[:, :3])).cuda()
for _ in range(3):
    e.reset(); e.push(x, box)
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(5):
    e.reset(); e.push(x, box)
ev1.record(); torch.cuda.synchronize()
print("SKIP=%s reduce ms/step: %.3f" % (os.environ.get("PBK_FUSED_SKIP", "0"), ev0.elapsed_time(ev1) / 5))
off = ((need.value + 63) // 64) * 64
prof = e._fused_scratch[off:off + 64 * nch.value].view(torch.int64).reshape(nch.value, 8).cpu().numpy()
names = ["wait_frame", "leaf phase", "tree(+labels t-1)", "lipid phase", "final labels", "-", "-", "-"]
if os.environ.get("PBK_FUSED_V1"):
    names = ["wait_frame", "phaseU", "phaseL(sync)", "tree", "phaseC", "phaseZ", "tail", "phaseL(own)"]
tot = prof.sum(axis=1)
print("chunks", nch.value, "frames/chunk", ch.value, "cycles per frame (mean over chunks):")
for i, n in enumerate(names):
    print(f"  {n:18s} {prof[:, i].mean() / ch.value:10.0f}  ({100 * prof[:, i].sum() / tot.sum():5.1f}%)")
print("  total      %10.0f" % (tot.mean() / ch.value))
