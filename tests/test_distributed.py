"""Multi-rank paths: world_size-2/3 gloo on the CPU for the host logic, and the
sharded BilayerAnalyzer on the GPU (NCCL with one GPU per rank, else the ranks share GPU 0
and talk over gloo)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from pybilt_b200 import distributed as pd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def launch(world, mode, case, timeout=600):
    port = free_port()
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE=str(world),
                   MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "dist_worker.py"),
                                       mode, case], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=timeout)[0] for p in procs]
    return [p.returncode for p in procs], outs


def test_shard_bounds_cover_all_frames():
    for n in (1, 2, 7, 10, 5000, 5001):
        for world in (1, 2, 3, 8):
            cover = []
            for r in range(world):
                lo, hi = pd.shard_bounds(n, world, r)
                cover.extend(range(lo, hi))
            assert cover == list(range(n))
    sh = pd.make_shard(100, 4, 2, forward_halo=9)
    assert (sh.lo, sh.hi, sh.back, sh.fwd, sh.first_local, sh.n_local) == (50, 75, 1, 9, 49, 35)
    assert pd.net_frames_of(sh, 9) == sh.back + sh.n_owned
    # leaflet update interval 4: local row 0 is a rebuild frame, and the image counts are handed on at
    # local row 0 of the next rank
    for r in range(1, 4):
        al = pd.make_shard(100, 4, r, forward_halo=9, align=4)
        assert al.first_local % 4 == 0 and 1 <= al.back <= 4
        prv = pd.make_shard(100, 4, r - 1, forward_halo=9, align=4)
        assert prv.first_local + pd.net_frames_of(prv, 9, 4) - 1 == al.first_local
    assert sh.local_of(50) == 1 and sh.owned_local() == slice(1, 26)
    last = pd.make_shard(100, 4, 3, forward_halo=9)
    assert last.fwd == 0 and last.hi == 100
    first = pd.make_shard(100, 4, 0, forward_halo=9)
    assert first.back == 0 and first.first_local == 0


def test_msd_multi_blocks_have_one_owner():
    from pybilt_b200.analysis_protocols import msd_multi_schedule
    finished, n_blocks = msd_multi_schedule(97, 10, 4)
    owners = []
    for r in range(3):
        sh = pd.make_shard(97, 3, r, forward_halo=9)
        mine = pd.msd_multi_owner_blocks(finished, sh)
        for b in mine:      # the block's end frame lies inside the shard's rows
            assert sh.first_local <= finished[b][1] < sh.first_local + sh.n_local
        owners.extend(mine)
    assert sorted(owners) == sorted(finished)


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("case", ["big_steps", "tiny_all"])
def test_shard_unwrap_algebra_gloo_cpu(world, case):
    codes, outs = launch(world, "cpu", case)
    assert codes == [0] * world, outs


@pytest.mark.gpu
@pytest.mark.parametrize("world,case", [(2, "tiny_all"), (2, "c2_prefix"), (3, "c3_small"),
                                        (2, "name_dict"), (2, "grid_dense"), (3, "flip_flop"),
                                        (2, "leaflet_dist"),
                                        # leaflets rebuilt every 3rd frame while lipids hop: shards must copy
                                        # the labels of the right rebuild frame
                                        (2, "flip_interval"), (3, "flip_interval"),
                                        # jumps of many box lengths: the fast path is refused on SOME ranks
                                        # only; all ranks have to fall back together
                                        (2, "big_steps"), (3, "big_steps")])
def test_sharded_analyzer_matches_reference(world, case):
    codes, outs = launch(world, "gpu", case)
    assert codes == [0] * world, outs


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("case", ["big_steps", "tiny_all"])
def test_streamed_shard_unwrap_algebra_gloo_cpu(world, case):
    codes, outs = launch(world, "cpuwin", case)
    assert codes == [0] * world, outs


@pytest.mark.gpu
@pytest.mark.parametrize("world,case,window", [(2, "tiny_all", 4), (2, "c2_prefix", 7), (3, "c3_small", 2),
                                               (2, "name_dict", 3), (2, "grid_dense", 1000),
                                               (2, "flip_flop", 3), (2, "leaflet_dist", 5),
                                               (2, "flip_interval", 4), (3, "flip_interval", 2),
                                               (2, "big_steps", 5), (3, "big_steps", 3)])
def test_streamed_sharded_analyzer_matches_reference(world, case, window, monkeypatch):
    monkeypatch.setenv("PBK_TEST_WINDOW", str(window))
    codes, outs = launch(world, "gpuwin", case)
    assert codes == [0] * world, outs


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4, 8])
def test_peer_memory_collectives_match_nccl(world):
    """pbk_comm_prefix_i32 / pbk_allreduce_u64 / pbk_allreduce_f64 over CUDA IPC peer windows against
    NCCL on random payloads, 40 rounds back to back (needs one GPU per rank)."""
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    codes, outs = launch(world, "peer", "tiny_all")
    assert codes == [0] * world, outs
