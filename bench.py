#!/usr/bin/env python
"""bench.py -- lipid-frames/s of the COM + MSD + lateral-RDF hot path on B200 (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one pass of the hot path over one batch of synthetic input: every frame of the
workload goes through unwrap -> membrane COM -> per-lipid COM -> leaflets, then single-origin
MSD, multi-origin MSD (n_tau = n_sigma = 50) and the COM lateral RDF (reference defaults:
25 bins, 0-25 A, both leaflets).  Workload at every N: BASELINE.json configs[1], the
512-lipid POPC CG bilayer (12 beads/lipid), 5,000 frames PER GPU (frame shards, weak
scaling; rank r holds frames r*5000 .. r*5000+4999 of one trajectory).

value   lipid-frames/s with the frames already resident in HBM (CUDA events, max over ranks)
e2e     the same through the public API (pybilt_b200.BilayerAnalyzer.run_analysis +
        get_analysis_data) from PINNED HOST buffers, host<->device copies inside the timing
roofline / cpu_baseline: see DESIGN.md "Measurement".

--impl reference times the CPU restatement of the reference (oracle/port.py; the reference
itself is pure Python + MDAnalysis, which cannot run on the GPU box) on all host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_LIPIDS, BEADS, FRAMES = 512, 12, 5000
ANALYSES = ["msd msd_1", "msd_multi mm n_tau 50 n_sigma 50", "com_lateral_rdf rdf"]
B_COM = 12 * BEADS + 48 + 1          # read f32 xyz of A atoms, write 2 f64 COMs + label
B_MSD = 16
B_RDF = 16 + 2
B_PIPE = B_COM + B_MSD + B_RDF       # = 12*A + 83
METRIC = "lipid-frames/s (COM+MSD+lateral RDF)"
WORKLOAD = ("C2: 512-lipid POPC CG bilayer (12 beads/lipid), %d frames per GPU; COM+unwrap+leaflets, msd, "
            "msd_multi(50,50), com_lateral_rdf(25 bins, 0-25 A, both leaflets)")
# the other named shapes of BASELINE.json, measured at a bounded number of frames (frames resident in
# HBM; generated frames x tile along time) and reported in the same line under "configs"
OTHER_CONFIGS = (("C3", 64, 4), ("C4", 32, 8), ("C5", 48, 4))


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def ncu_traffic(stage: str):
    """dram read+write bytes per step of a stage, from the committed ncu capture
    (profiles/traffic.json: per-launch bytes of each kernel); None if not captured."""
    kernels = {"reduce(unwrap+memCOM+lipidCOM+leaflets)": ["k_chunk_box", "k_img3", "k_scan3", "k_evsort", "k_frames4<12, 1>"],
               "rdf": ["k_rdf5<1>"], "msd": ["k_msd_pairs"], "msd_multi": ["k_msd_pairs"]}
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            per = json.load(fh)["bytes_per_launch"]
        return float(sum(per[k] for k in kernels[stage]))
    except Exception:
        return None


class ClockSampler(threading.Thread):
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        self.join(timeout=2)
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); smax.append(float(r[1]))
                for nm, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        busy = [v for v in sm if v > 0.5 * max(sm)] or sm
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(smax),
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------
def cpu_port_run(args):
    """One shard of the CPU baseline: the oracle port over `frames` frames (own process)."""
    rank, first, frames = args
    from oracle import port
    from pybilt_b200 import synthetic as syn
    spec, _ = syn.config_spec("C2")
    top, traj = syn.generate(spec, frames, first_frame=first)
    t0 = time.perf_counter()
    port.run(top, traj, ANALYSES)
    return time.perf_counter() - t0


def cpu_baseline(frames: int, procs: int):
    """lipid-frames/s of oracle/port.py over `frames` frames split over `procs` processes
    (each shard an independent frame range, like COMTraj.calc_msd_parallel's mp.Pool)."""
    import multiprocessing as mp
    from oracle import build as obuild
    obuild.build()
    per = max(50, frames // procs)
    jobs = [(i, i * per, per) for i in range(procs)]
    t0 = time.perf_counter()
    if procs == 1:
        cpu_port_run(jobs[0])
    else:
        with mp.get_context("spawn").Pool(procs) as pool:
            pool.map(cpu_port_run, jobs)
    wall = time.perf_counter() - t0
    return N_LIPIDS * per * procs / wall, per * procs, wall


def reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    procs = os.cpu_count() or 1
    # spawn start-up (numpy import + generating the shard) is excluded by timing inside the
    # workers; the line reports whole-job throughput over the slowest worker
    import multiprocessing as mp
    from oracle import build as obuild
    obuild.build()
    # exactly FRAMES frames (the GPU arm's workload), split as evenly as the core count allows
    procs = max(1, min(procs, FRAMES // 50))
    cuts = [FRAMES * i // procs for i in range(procs + 1)]
    jobs = [(i, cuts[i], cuts[i + 1] - cuts[i]) for i in range(procs)]
    n_frames = cuts[-1]
    per = max(j[2] for j in jobs)
    times = []
    with mp.get_context("spawn").Pool(procs) as pool:
        for _ in range(a.warmup):
            pool.map(cpu_port_run, jobs)
        for _ in range(a.steps):
            times.append(max(pool.map(cpu_port_run, jobs)))
    ms = 1e3 * sum(times) / len(times)
    value = N_LIPIDS * n_frames / (ms * 1e-3)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "lipid-frames/s",
            "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": WORKLOAD % n_frames, "frames_per_gpu": n_frames, "lipids": N_LIPIDS,
                       "atoms": N_LIPIDS * BEADS, "sharding": "frames"},
            "cpu_baseline": {"value": value, "unit": "lipid-frames/s", "cores": procs,
                             "kind": "port",
                             "sample": "all %d frames in %d frame-range shards of at most %d (oracle/port.py, "
                                       "one process per host core)" % (n_frames, procs, per)},
            "e2e": {"value": value, "unit": "lipid-frames/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=FRAMES, help="frames per GPU (default: config C2)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the C3/C4/C5 stage timings and the C5 sharded leg")
    a = ap.parse_args()
    if a.impl == "reference":
        reference_arm(a)
        return

    import torch
    import torch.distributed as dist
    from pybilt_b200 import BilayerAnalyzer, engine as eng
    from pybilt_b200.analysis_protocols import msd_multi_schedule

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    F = a.frames
    from pybilt_b200 import distributed as pd
    from pybilt_b200 import synthetic as syn
    from pybilt_b200.bilayer_analyzer import FrameWindow
    spec, _ = syn.config_spec("C2")
    F_total = F * world
    # frame shard of one F*world-frame trajectory: [1-frame back halo | F owned | 49-frame
    # forward halo for msd_multi(50,50)]; at N = 1 this is just the F frames
    sh = pd.make_shard(F_total, world, rank, forward_halo=49) if world > 1 else None
    first_local = sh.first_local if sh else 0
    n_local = sh.n_local if sh else F
    top, traj = syn.generate(spec, n_local, first_frame=first_local)
    M = top.n_atoms
    layout = eng.build_bead_layout(top.masses, top.names, top.resindex, top.resnames, top.resids)
    e = eng.LipidReductionEngine(top.masses, layout, n_local + 1, device=local, batch_frames=n_local)

    x_host = torch.from_numpy(traj.positions).pin_memory()
    box_host = torch.from_numpy(np.ascontiguousarray(traj.dimensions[:, :3])).pin_memory()
    x_dev = x_host.to(dev)
    box_dev = box_host.to(dev)
    codes = torch.zeros(layout.n_beads, dtype=torch.int32, device=dev)
    edges = np.histogram([-1], bins=25, range=[0.0, 25.0])[1]
    edges_dev = torch.from_numpy(edges).to(dev)
    thr_dev = torch.from_numpy(eng.rdf_thresholds(edges)).to(dev)
    fin, _nb = msd_multi_schedule(F_total, 50, 50)
    row = sh.local_of if sh else (lambda g: g)
    mine = [b for b in sorted(fin) if sh is None or sh.lo <= fin[b][0] < sh.hi]
    mm_o = torch.tensor([row(fin[b][0]) for b in mine], dtype=torch.int32, device=dev)
    mm_e = torch.tensor([row(fin[b][1]) for b in mine], dtype=torch.int32, device=dev)
    owned = sh.owned_local() if sh else slice(0, F)
    ref_row = n_local if sh else 0
    origins = torch.full((F,), ref_row, dtype=torch.int32, device=dev)
    ends = torch.arange(owned.start, owned.stop, dtype=torch.int32, device=dev)
    net = torch.zeros((M, 3), dtype=torch.int32, device=dev)
    n_a = (sh.back + sh.n_owned) if sh else F

    # collectives that nothing on the frame data path waits for (the broadcast of frame 0's COMs is
    # only read by the MSD kernels, the histogram all-reduce by nobody inside the step) run on a side
    # stream next to the RDF sweep
    comm = torch.cuda.Stream(dev) if world > 1 else None
    pc_side = pd.peer_comm(dev, 1 << 16, lane="side") if world > 1 else None
    if world > 1:
        pd.peer_comm(dev, M * 12)          # the image-count prefix (pd.image_offsets) uses the main lane
    ev_reduced, ev_ref = torch.cuda.Event(), torch.cuda.Event()

    def reduce_all():
        e.reset()
        if sh is None:
            e.push(x_dev, box_dev)
            return
        # [back halo | owned | forward halo] in one push; image counts exchanged after the owned frames
        e.push(x_dev, box_dev, phase=1, net_out=net, net_frames=n_a)  # count periodic images
        off = pd.image_offsets(net)                                   # NCCL all-gather + prefix
        e.push(x_dev, box_dev, phase=2, init_images=off, net_frames=n_a)   # finish from the true counts
        main = torch.cuda.current_stream(dev)
        ev_reduced.record(main)
        with torch.cuda.stream(comm):
            comm.wait_event(ev_reduced)
            ref = e.com_unwrap[0].clone()
            pd.broadcast_from_first(ref)                              # frame 0 of the whole run
            e.com_unwrap[n_local].copy_(ref)
            ev_ref.record(comm)

    # index list of 'msd ... leaflet both resname all': upper members then lower (frame 0 of the run)
    reduce_all()
    lab0 = e.labels[0].clone()
    if sh is not None:
        pd.broadcast_from_first(lab0)
    lab0 = lab0.cpu().numpy()
    idx = torch.from_numpy(np.concatenate([np.nonzero(lab0 == 1)[0], np.nonzero(lab0 == 0)[0]])
                           .astype(np.int32)).to(dev)

    stage_names = (["reduce(unwrap+memCOM+lipidCOM+leaflets)", "msd", "msd_multi", "rdf"] if world == 1 else
                   ["reduce(unwrap+memCOM+lipidCOM+leaflets)", "rdf", "msd", "msd_multi"])

    def step(events=None):
        def mark(i):
            if events is not None:
                events[i].record()
        main = torch.cuda.current_stream(dev)
        mark(0)
        reduce_all()
        mark(1)
        if world > 1:
            # the RDF sweep does not read the broadcast reference frame: it goes first and overlaps
            # the broadcast; its partial histograms are all-reduced on the side stream
            cnt, fc = e.rdf(codes, 0, 0, 3, edges, edges_dev=edges_dev, thr_dev=thr_dev, frames=owned)
            ev_rdf = torch.cuda.Event()
            ev_rdf.record(main)
            with torch.cuda.stream(comm):
                comm.wait_event(ev_rdf)
                # partial histograms of the frame shards: one libpbk kernel over NVLink peer memory
                # (pbk_allreduce_u64), NCCL where the ranks have no peer access
                if pc_side is not None:
                    pc_side.allreduce_(cnt)
                else:
                    dist.all_reduce(cnt)
                cnt.record_stream(comm)
            mark(2)
            main.wait_event(ev_ref)
            m1 = e.msd_pairs(idx, origins, ends)
            mark(3)
            m2 = e.msd_pairs(idx, mm_o, mm_e) if len(mm_o) else None
            mark(4)
            return m1, m2, cnt
        m1 = e.msd_pairs(idx, origins, ends)
        mark(2)
        m2 = e.msd_pairs(idx, mm_o, mm_e) if len(mm_o) else None
        mark(3)
        cnt, fc = e.rdf(codes, 0, 0, 3, edges, edges_dev=edges_dev, thr_dev=thr_dev, frames=owned)
        mark(4)
        return m1, m2, cnt

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # at N > 1 the first collectives of a process group are slow (channel set-up): never fewer than
    # five untimed steps there, whatever --warmup says
    for _ in range(a.warmup if world == 1 else max(a.warmup, 5)):
        step()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.15)
    l0 = e.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(a.steps):
        out = step()
    ev1.record()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    launches = e.launches - l0
    e.check_status()
    # per-stage device time (same steps, instrumented) -> dominant kernel for the roofline
    stage_ms = np.zeros(4)
    reps = max(3, min(a.steps, 10))
    for _ in range(reps):
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        step(evs)
        torch.cuda.synchronize()
        stage_ms += np.array([evs[i].elapsed_time(evs[i + 1]) for i in range(4)])
    stage_ms /= reps

    # ---- end to end through the public API, from pinned host memory ---------------------
    dims_all = syn.box_at(spec, np.arange(F_total))
    times_all = np.arange(F_total) * spec.dt
    ba = BilayerAnalyzer(structure=top, trajectory=(FrameWindow(x_host, first_local, F_total), dims_all,
                                                    times_all), selection="all", device=local)
    ba.remove_analysis("msd_1")
    for s in ANALYSES:
        ba.add_analysis(s)

    def api_step():
        ba.reset()
        ba.run_analysis()
        return [ba.get_analysis_data(k) for k in ("msd_1", "mm", "rdf")]

    for _ in range(max(1, a.warmup)):
        res = api_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        res = api_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop()
    d2h = layout.n_beads + F * 8 + len(mm_o) * 8 + 25 * 8 + F * 24 + 4   # labels[0], msd, mm, rdf count, frame counts, status
    h2d = int(x_host.numel() * 4 + box_host.numel() * 4)

    t = torch.tensor([ms_total, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms = float(t[0]), float(t[1])
    ms_step = ms_total / a.steps
    total_lf = N_LIPIDS * F * world
    value = total_lf / (ms_step * 1e-3)
    e2e_val = total_lf / (e2e_ms / a.steps * 1e-3)

    # ---- the C5 shape as a streamed, frame-sharded run through the public API (every rank) ----------
    c5_leg = None
    if not a.no_configs:
        del ba, e, x_dev
        torch.cuda.empty_cache()
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import bench_c5_sharded
        c5_leg = bench_c5_sharded.measure(frames_per_gpu=48, window=96, tile=4, reps=2)

    if rank == 0:
        peak, which = peaks()
        dom = int(np.argmax(stage_ms))
        per_lf = {"reduce(unwrap+memCOM+lipidCOM+leaflets)": B_COM, "msd": B_MSD,
                  "msd_multi": B_MSD * max(len(mm_o), 1) / F, "rdf": B_RDF}
        dom_bytes = per_lf[stage_names[dom]] * N_LIPIDS * F
        achieved = dom_bytes / (stage_ms[dom] * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": "lipid-frames/s", "n_gpus": world,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD % F,
                       "frames_per_gpu": F, "lipids": N_LIPIDS, "atoms": M,
                       "l2": "inputs larger than L2 (%.0f MB of positions per step)" % (h2d / 1e6),
                       "sharding": "frames"},
            "e2e": {"value": e2e_val, "unit": "lipid-frames/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms / a.steps},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": stage_names[dom], "achieved": achieved,
                         "peak": peak, "peak_source": which, "unit": "GB/s",
                         "frac": achieved / peak,
                         "traffic": ncu_traffic(stage_names[dom]) if (world == 1 and F == FRAMES) else None,
                         "algorithmic_bytes": dom_bytes,
                         "pipeline_achieved": B_PIPE * N_LIPIDS * F / (ms_step * 1e-3) / 1e9,
                         "pipeline_frac": B_PIPE * N_LIPIDS * F / (ms_step * 1e-3) / 1e9 / peak},
            "stages_ms": {n: float(v) for n, v in zip(stage_names, stage_ms)},
        }
        if c5_leg is not None:
            line["c5_sharded"] = c5_leg
        if not a.no_configs and world == 1:
            # stage times and roofline fractions of the other named shapes (one GPU)
            import bench_configs
            line["configs"] = {}
            for cfg, frames, tile in OTHER_CONFIGS:
                r = bench_configs.measure(cfg, frames, 3, False, tile, peak_gbs=peak)
                line["configs"][cfg] = {"lipids": r["lipids"], "atoms": r["atoms"], "frames": r["frames"],
                                        "stages": {k: {"ms": v["ms"], "GBps_algorithmic": v["GBps_algorithmic"],
                                                       "frac": v.get("frac"), "bytes_per_lipid_frame": v["bytes_per_lipid_frame"]}
                                                   for k, v in r["stages"].items()},
                                        "lipid_frames_per_s": r["lipid_frames_per_s"]}
        if not a.no_cpu_baseline and world == 1:
            v, nfr, wall = cpu_baseline(FRAMES, 1)
            line["cpu_baseline"] = {"value": v, "unit": "lipid-frames/s", "cores": 1, "kind": "port",
                                    "sample": "all %d frames of the workload, oracle/port.py in one "
                                              "process (%.1f s)" % (nfr, wall)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
