"""Worker of the multi-rank tests: `python tests/dist_worker.py <mode> <case>` under torchrun
(or spawned with RANK / WORLD_SIZE / MASTER_* set).

mode cpu : the shard algebra (one-frame back halo, net image counts, exclusive prefix over
           ranks) with the oracle port standing in for the device, gloo backend, no GPU.
mode cpuwin : the same for streamed shards (context frames before the first owned frame, image
           counts taken where the next rank's stream starts).
mode gpuwin : like gpu, but as a streamed run (window_frames = $PBK_TEST_WINDOW).
mode gpu : pybilt_b200.BilayerAnalyzer sharded over the ranks (NCCL when every rank has its
           own GPU, otherwise gloo with the ranks sharing GPU 0) against the golden fixture.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist


def main():
    mode, case = sys.argv[1], sys.argv[2]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    from tests import cases
    from pybilt_b200 import distributed as pd
    top, traj = cases.make_inputs(case)
    c = cases.CASES[case]
    if mode == "cpu":
        from oracle import port
        dist.init_process_group("gloo")
        X, box = traj.positions, traj.dimensions[:, :3]
        F = X.shape[0]
        serial = port.unwrap(X, box)
        sh = pd.make_shard(F, world, rank, forward_halo=2)
        loc = slice(sh.first_local, sh.first_local + sh.n_local)
        U_loc = port.unwrap(X[loc], box[loc])                       # relative to the halo frame
        s_loc = np.rint((U_loc.astype(np.float64) - X[loc]) / box[loc][:, None, :]).astype(np.int32)
        last = sh.back + sh.n_owned - 1
        off = pd.image_offsets(torch.from_numpy(s_loc[last].copy())).numpy()
        u_true = (X[loc].astype(np.float64) + (s_loc + off[None]) * box[loc][:, None, :].astype(np.float64)).astype(np.float32)
        ok = np.array_equal(u_true, serial[loc])
        # slot combination: every rank contributes its owned frames of a per-frame vector
        vec = torch.zeros(F, dtype=torch.float64)
        vec[sh.lo:sh.hi] = torch.arange(sh.lo, sh.hi, dtype=torch.float64) + 0.5
        full = pd.combine_slots(vec).numpy()
        ok = ok and np.array_equal(full, np.arange(F) + 0.5)
        flag = torch.tensor([1 if ok else 0])
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        dist.destroy_process_group()
        sys.exit(0 if int(flag) == 1 else 3)
    if mode == "cpuwin":
        # streamed shards (BilayerAnalyzer._run_windowed): rank r's stream starts C frames before its
        # first owned frame; the image counts there = prefix over ranks of the counts each rank finds,
        # unwrapping alone, at the frame where the NEXT rank's stream starts
        from oracle import port
        dist.init_process_group("gloo")
        X, box = traj.positions, traj.dimensions[:, :3]
        F = X.shape[0]
        C = 2
        serial = port.unwrap(X, box)
        starts = [max(0, pd.shard_bounds(F, world, q)[0] - C) for q in range(world)]
        lo, hi = pd.shard_bounds(F, world, rank)
        a = starts[rank]
        net = np.zeros(X.shape[1:], dtype=np.int32)
        if rank < world - 1 and starts[rank + 1] > a:
            seg = slice(a, starts[rank + 1] + 1)
            U = port.unwrap(X[seg], box[seg])
            net = np.rint((U[-1].astype(np.float64) - X[seg][-1]) / box[seg][-1][None, :]).astype(np.int32)
        off = pd.image_offsets(torch.from_numpy(net.copy())).numpy()
        loc = slice(a, hi)
        U_loc = port.unwrap(X[loc], box[loc])
        s_loc = np.rint((U_loc.astype(np.float64) - X[loc]) / box[loc][:, None, :]).astype(np.int32)
        u_true = (X[loc].astype(np.float64) + (s_loc + off[None]) * box[loc][:, None, :].astype(np.float64)).astype(np.float32)
        ok = np.array_equal(u_true, serial[loc])
        flag = torch.tensor([1 if ok else 0])
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        dist.destroy_process_group()
        sys.exit(0 if int(flag) == 1 else 3)
    if mode == "peer":
        # libpbk's NVLink peer-memory collectives (pbk_comm_*) against NCCL, one GPU per rank
        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
        dist.init_process_group("nccl", device_id=dev)
        pc = pd.peer_comm(dev, 1 << 21)
        ok = pc is not None
        if ok:
            g = torch.Generator(device="cpu").manual_seed(100 + rank)
            for it in range(40):
                n = int(torch.randint(1, 200000, (1,), generator=g)) if it else 18432
                n = int(pd.combine_slots(torch.tensor([n if rank == 0 else 0], device=dev))[0])
                net = torch.randint(-5, 6, (n,), generator=g, dtype=torch.int32).to(dev)
                got = pc.prefix_i32(net)
                alln = [torch.empty_like(net) for _ in range(world)]
                dist.all_gather(alln, net)
                want = sum(alln[:rank]) if rank else torch.zeros_like(net)
                ok = ok and bool(torch.equal(got, want))
                h = torch.randint(0, 1 << 40, (n,), generator=g, dtype=torch.int64).to(dev)
                d = torch.rand((n,), generator=g, dtype=torch.float64).to(dev)
                h2, d2 = h.clone(), d.clone()
                pc.allreduce_(h2)
                pc.allreduce_(d2)
                dist.all_reduce(h)
                alld = [torch.empty_like(d) for _ in range(world)]
                dist.all_gather(alld, d)
                dsum = alld[0].clone()
                for q in range(1, world):
                    dsum += alld[q]
                ok = ok and bool(torch.equal(h2, h)) and bool(torch.equal(d2, dsum))
            pc.check()
        flag = torch.tensor([1 if ok else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        dist.destroy_process_group()
        sys.exit(0 if int(flag) == 1 else 3)
    # ---- gpu ----
    ngpu = torch.cuda.device_count()
    own_gpu = ngpu >= world
    dev = rank if own_gpu else 0
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl" if own_gpu else "gloo")
    from pybilt_b200 import BilayerAnalyzer
    gold = dict(np.load(os.path.join(ROOT, "tests", "golden", case + ".npz"), allow_pickle=False))
    ba = BilayerAnalyzer(structure=top, trajectory=traj, selection="all", device=dev)
    ba.remove_analysis("msd_1")
    for a in c["analyses"]:
        ba.add_analysis(a)
    for rep, d in (c.get("rep_settings") or {}).items():
        ba.rep_settings[rep].update(d)
    if c.get("frame_range"):
        ba.set_frame_range(*c["frame_range"])
    if mode == "gpuwin":
        ba.window_frames = int(os.environ.get("PBK_TEST_WINDOW", "4"))
    ba.run_analysis()
    flat = cases.flatten_outputs({k: ba.get_analysis_data(k) for k in ba.get_analysis_ids()})
    ok = True
    for k in [k for k in gold if k.startswith("out/")]:
        a, b = np.asarray(flat[k], dtype=float), np.asarray(gold[k], dtype=float)
        good = (np.isnan(a) & np.isnan(b)) | (np.abs(a - b) <= 1e-9 * np.maximum(np.abs(a), np.abs(b))
                                               + (cases.output_atol(k, b) if a.shape == b.shape else 0.0))
        if a.shape != b.shape or not good.all():
            print("rank", rank, "MISMATCH", k, flush=True)
            ok = False
    if mode == "gpuwin":
        flag = torch.tensor([1 if ok else 0])
        if own_gpu:
            flag = flag.cuda()
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        dist.destroy_process_group()
        sys.exit(0 if int(flag) == 1 else 3)
    sh = ba._shard
    e = ba._engine
    rows = sh.owned_local()
    ok = ok and np.array_equal(e.labels[rows].cpu().numpy(), gold["rep/labels"][sh.lo:sh.hi])
    ok = ok and np.array_equal(e.com[rows].cpu().numpy(), gold["rep/com"][sh.lo:sh.hi])
    ok = ok and np.array_equal(e.com_unwrap[rows].cpu().numpy(), gold["rep/com_unwrap"][sh.lo:sh.hi])
    ok = ok and np.array_equal(e.mem_com[rows].cpu().numpy(), gold["rep/mem_com"][sh.lo:sh.hi])
    flag = torch.tensor([1 if ok else 0])
    if own_gpu:
        flag = flag.cuda()
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if int(flag) == 1 else 3)


if __name__ == "__main__":
    main()
