"""Where the end-to-end time goes: python tools/prof_api2.py"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pybilt_b200 import BilayerAnalyzer, synthetic as syn, engine as eng
spec, _ = syn.config_spec("C2")
top, traj = syn.generate(spec, 5000)
x_host = torch.from_numpy(traj.positions).pin_memory()
def make(analyses):
    ba = BilayerAnalyzer(structure=top, trajectory=(x_host, traj.dimensions, traj.times), selection="all")
    ba.remove_analysis("msd_1")
    for s in analyses: ba.add_analysis(s)
    return ba
def bench(ba, keys, label, n=4):
    ts = []
    for i in range(n):
        ba.reset(); torch.cuda.synchronize(); t0 = time.perf_counter(); ba.run_analysis()
        r = [ba.get_analysis_data(k) for k in keys]; torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    print("%-44s %s" % (label, " ".join("%.2f" % t for t in ts)))
full = ["msd msd_1", "msd_multi mm n_tau 50 n_sigma 50", "com_lateral_rdf rdf"]
for bf in (None, 2500, 1250, 650, 325):
    ba = make(full); ba.batch_frames = bf
    bench(ba, ("msd_1", "mm", "rdf"), "full, batch_frames=%s" % bf)
ba = make(["msd msd_1"]); ba.batch_frames = 500
bench(ba, ("msd_1",), "msd only, batch 500")
orig = eng.LipidReductionEngine.push
def nopush(self, x, box, phase=3, init_images=None, net_out=None):
    self.n_done += int(x.shape[0])
eng.LipidReductionEngine.push = nopush
ba = make(["msd msd_1"]); ba.batch_frames = 500
bench(ba, ("msd_1",), "copies only (push = no-op), batch 500")
