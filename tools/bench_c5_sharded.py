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


def measure(frames_per_gpu=48, window=16, lipids=262144, tile=1, reps=3):
    """The run described above inside an initialised process group (or a single process); returns the
    JSON-able summary on rank 0 and None elsewhere."""
    import torch
    import torch.distributed as dist
    from pybilt_b200 import BilayerAnalyzer, synthetic as syn, distributed as pd
    from pybilt_b200.bilayer_analyzer import FrameWindow
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    local = torch.cuda.current_device()
    spec, _ = syn.config_spec("C5")
    if lipids != 262144:
        spec = syn.BilayerSpec(lipids, spec.lipid_types, seed=spec.seed)
    F = frames_per_gpu * tile * world
    lo, hi = pd.shard_bounds(F, world, rank)
    first = max(0, lo - 1)                       # the stream starts one context frame early
    if tile > 1:
        # the same generated frames repeated along time (the seams are ordinary small jumps)
        top, traj = syn.generate(spec, frames_per_gpu)
        nrep = -(-(hi - first) // frames_per_gpu) + 1
        off = first % frames_per_gpu
        x_host = torch.from_numpy(traj.positions).repeat(nrep, 1, 1)[off:off + hi - first].contiguous().pin_memory()
        dims = np.tile(traj.dimensions, (F // frames_per_gpu + 1, 1))[:F]
    else:
        top, traj = syn.generate(spec, hi - first, first_frame=first)
        x_host = torch.from_numpy(traj.positions).pin_memory()
        dims = syn.box_at(spec, np.arange(F))
    ba = BilayerAnalyzer(structure=top, trajectory=(FrameWindow(x_host, first, F), dims, np.arange(F) * spec.dt),
                         selection="all", device=local)
    ba.add_analysis("com_lateral_rdf rdf resname_1 all resname_2 all")
    ba.window_frames = window
    times = []
    for _ in range(reps):
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
    del ba
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    return {"workload": "C5 shape: %d lipids, %d frames per GPU, msd + com_lateral_rdf(all, all), frame shards "
                        "streamed from pinned host memory in windows of %d frames" % (lipids, frames_per_gpu * tile, window),
            "n_gpus": world, "seconds": float(t[0]), "lipid_frames_per_s_e2e": lipids * F / float(t[0]),
            "msd_last": float(msd[-1, 1]), "rdf_peak": float(np.nanmax(rdf)), "h2d_bytes_per_gpu": int(x_host.numel() * 4)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames-per-gpu", type=int, default=48)
    ap.add_argument("--window", type=int, default=16)
    ap.add_argument("--lipids", type=int, default=262144)
    ap.add_argument("--tile", type=int, default=1, help="repeat the generated frames along time (generation is the slow part)")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    out = measure(a.frames_per_gpu, a.window, a.lipids, a.tile)
    if out is not None:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
