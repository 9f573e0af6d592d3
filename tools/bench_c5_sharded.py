#!/usr/bin/env python
"""Config C5 through the public API as a STREAMED, FRAME-SHARDED run: 262,144-lipid bilayer (one
bead per lipid), `msd` + `com_lateral_rdf resname_1 all resname_2 all`, frames split over the ranks
(torchrun), every rank streaming its shard from pinned host memory in windows.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/bench_c5_sharded.py \
        [--frames-per-gpu 48] [--window 16] [--lipids 262144]

Prints one JSON line (rank 0): whole-job lipid-frames/s end to end (host->device copies included).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames-per-gpu", type=int, default=48)
    ap.add_argument("--window", type=int, default=16)
    ap.add_argument("--lipids", type=int, default=262144)
    ap.add_argument("--tile", type=int, default=1, help="repeat the generated frames along time (generation is the slow part)")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    from pybilt_b200 import BilayerAnalyzer, synthetic as syn, distributed as pd
    from pybilt_b200.bilayer_analyzer import FrameWindow
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    spec, _ = syn.config_spec("C5")
    if a.lipids != 262144:
        spec = syn.BilayerSpec(a.lipids, spec.types, seed=spec.seed)
    F = a.frames_per_gpu * a.tile * world
    lo, hi = pd.shard_bounds(F, world, rank)
    first = max(0, lo - 1)                       # the stream starts one context frame early
    if a.tile > 1:
        # the same generated frames repeated along time (the seams are ordinary small jumps)
        top, traj = syn.generate(spec, a.frames_per_gpu)
        reps = -(-(hi - first) // a.frames_per_gpu) + 1
        off = first % a.frames_per_gpu
        x_host = torch.from_numpy(traj.positions).repeat(reps, 1, 1)[off:off + hi - first].contiguous().pin_memory()
        dims = np.tile(traj.dimensions, (F // a.frames_per_gpu + 1, 1))[:F]
    else:
        top, traj = syn.generate(spec, hi - first, first_frame=first)
        x_host = torch.from_numpy(traj.positions).pin_memory()
        dims = syn.box_at(spec, np.arange(F))
    ba = BilayerAnalyzer(structure=top, trajectory=(FrameWindow(x_host, first, F), dims, np.arange(F) * spec.dt),
                         selection="all", device=local)
    ba.add_analysis("com_lateral_rdf rdf resname_1 all resname_2 all")
    ba.window_frames = a.window
    times = []
    for _ in range(3):
        ba.reset()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ba.run_analysis()
        msd = ba.get_analysis_data("msd_1")
        rdf, bins = ba.get_analysis_data("rdf")
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        times.append(time.perf_counter() - t0)
    t = torch.tensor([min(times)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"workload": "C5 shape: %d lipids, %d frames per GPU, msd + com_lateral_rdf(all, all), "
                                      "streamed in windows of %d frames" % (a.lipids, a.frames_per_gpu * a.tile, a.window),
                          "n_gpus": world, "seconds": float(t[0]),
                          "lipid_frames_per_s_e2e": a.lipids * F / float(t[0]),
                          "msd_last": float(msd[-1, 1]), "rdf_peak": float(np.nanmax(rdf)),
                          "h2d_bytes_per_gpu": int(x_host.numel() * 4)}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
