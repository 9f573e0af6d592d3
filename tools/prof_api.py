import sys, time, cProfile, pstats, io
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from pybilt_b200 import BilayerAnalyzer, synthetic as syn
spec,_ = syn.config_spec("C2")
top, traj = syn.generate(spec, 5000)
x_host = torch.from_numpy(traj.positions).pin_memory()
ba = BilayerAnalyzer(structure=top, trajectory=(x_host, traj.dimensions, traj.times), selection="all")
ba.remove_analysis("msd_1")
for s in ["msd msd_1", "msd_multi mm n_tau 50 n_sigma 50", "com_lateral_rdf rdf"]: ba.add_analysis(s)
for bf in (500, 2500, None):
    ba.batch_frames = bf
    for i in range(3):
        ba.reset(); torch.cuda.synchronize(); t0=time.perf_counter(); ba.run_analysis(); r=[ba.get_analysis_data(k) for k in ("msd_1","mm","rdf")]; torch.cuda.synchronize(); print(bf, "step", i, (time.perf_counter()-t0)*1e3, "ms")
ba.batch_frames = None
ba.reset()
pr = cProfile.Profile(); pr.enable(); ba.run_analysis(); r=[ba.get_analysis_data(k) for k in ("msd_1","mm","rdf")]; pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(25); print(s.getvalue()[:4500])
