"""CUDA path (libpbk through pybilt_b200.engine) against the golden fixtures of the literal
reference and against the oracle port, kernel by kernel."""
import os

import numpy as np
import pytest
import torch

from oracle import port
from tests import cases

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
COM_RTOL = 1e-9          # north_star: COMs and MSD within 1e-9 relative in fp64


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))


def rel_close(a, b, rtol, atol=0.0):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape
    err = np.abs(a - b)
    lim = rtol * np.maximum(np.abs(a), np.abs(b)) + atol
    assert (err <= lim).all(), f"max rel err {np.max(err / np.maximum(np.abs(b), 1e-300))}"


def run_engine(name, batch_frames=None, fused=True, stream=True):
    from pybilt_b200 import engine as eng
    c = cases.CASES[name]
    top, traj = cases.make_inputs(name)
    cf = (c.get("rep_settings") or {}).get("com_frame", {})
    lf = (c.get("rep_settings") or {}).get("leaflets", {})
    fr = port.analysed_frames(traj.positions.shape[0], c.get("frame_range", (0, -1, 1)))
    layout = eng.build_bead_layout(top.masses, top.names, top.resindex, top.resnames, top.resids,
                                   cf.get("name_dict"), cf.get("multi_bead", False))
    e = eng.LipidReductionEngine(top.masses, layout, len(fr), norm=2,
                                 leaflet_update_interval=lf.get("update_interval", 1),
                                 batch_frames=batch_frames, allow_fused=fused, allow_stream=stream,
                                 rewrap=cf.get("rewrap", False))
    if lf.get("assign_method") == "orientation":
        from pybilt_b200 import BilayerAnalyzer
        e.set_orientation(*BilayerAnalyzer._orientation_selection(top, layout, lf["orientation_atoms"]))
    x = torch.from_numpy(np.ascontiguousarray(traj.positions[fr])).cuda()
    box = torch.from_numpy(np.ascontiguousarray(traj.dimensions[fr][:, :3])).cuda()
    e.push(x, box)
    e.check_status()
    return e, top, traj, fr, layout


@pytest.mark.parametrize("name", list(cases.CASES))
@pytest.mark.parametrize("batch", [None, 5])
@pytest.mark.parametrize("path", ["fused", "stream", "stream_seq", "stream_seqb", "stream_v1", "generic"])
def test_representations_match_reference(name, batch, path, monkeypatch):
    from pybilt_b200.engine import FusedPathRejected
    gold = load(name)
    fused = path == "fused"
    # the streamed reduction's variants (pbk_sweep.cu): time-parallel front end (default), the
    # warp-per-leaf sweep it falls back to, sequential chunks in sweep B, the first generation
    if path == "stream_seq":
        monkeypatch.setenv("PBK_SWEEP_SEQ_A", "1")
    if path == "stream_seqb":
        monkeypatch.setenv("PBK_SWEEP_SEQ_B", "1")
    if path == "stream_v1":
        monkeypatch.setenv("PBK_SWEEP_V1", "1")
    if path.startswith("stream"):
        path = "stream"
    try:
        e, top, traj, fr, layout = run_engine(name, batch, fused, stream=path != "generic")
    except FusedPathRejected:
        # only the adversarial case (jumps of many box lengths) may be refused by the fast path
        assert fused and name == "big_steps"
        return
    if fused and name not in ("big_steps",):
        assert e.used_fused
    if path == "stream":
        assert e.used_stream and not e.used_fused
    if path == "generic":
        assert not e.used_stream and not e.used_fused
    com = e.com.cpu().numpy()
    comu = e.com_unwrap.cpu().numpy()
    rel_close(com, gold["rep/com"], COM_RTOL)
    rel_close(comu, gold["rep/com_unwrap"], COM_RTOL, atol=1e-12)
    # the membrane COM is reproduced bit for bit (NumPy pairwise order)
    assert np.array_equal(e.mem_com.cpu().numpy(), gold["rep/mem_com"])
    assert np.array_equal(e.labels.cpu().numpy(), gold["rep/labels"])          # bit-exact labels
    assert np.array_equal(e.d_u_state.cpu().numpy(), gold["rep/u_last"])       # bit-exact unwrap
    rel_close(layout.bead_mass, gold["rep/mass"], 1e-15)
    # stronger than required: the COMs come out bit-identical too
    assert np.array_equal(com, gold["rep/com"])
    assert np.array_equal(comu, gold["rep/com_unwrap"])


def type_codes(types):
    names = []
    for t in types:
        if t not in names:
            names.append(t)
    return names, np.array([names.index(t) for t in types], dtype=np.int32)


def port_reps(name):
    c = cases.CASES[name]
    top, traj = cases.make_inputs(name)
    out, reps = port.run(top, traj, [], c.get("rep_settings"), c.get("settings"),
                         c.get("frame_range", (0, -1, 1)), return_reps=True)
    return reps


@pytest.mark.parametrize("name", ["tiny_all", "c1_sample", "c2_prefix", "name_dict"])
def test_msd_pairs_match_port(name):
    e, top, traj, fr, layout = run_engine(name)
    reps = port_reps(name)
    F = len(fr)
    views0 = port.leaflet_views(reps["labels"][0], reps["types"])
    for leaflet, group in [("both", "all"), ("upper", "all"), ("lower", layout.types[0])]:
        idx = port._indices_for(views0, leaflet, group)
        want = port.msd(reps["com_unwrap"], reps["times"], idx)[:, 1]
        got = e.msd_pairs(idx, np.zeros(F, dtype=np.int32), np.arange(F, dtype=np.int32)).cpu().numpy()
        rel_close(got, want, COM_RTOL)
        assert got[0] == 0.0
    pairs, _ = port.msd_multi_blocks(F, 3, 2)
    idx = port._indices_for(views0, "both", "all")
    r = reps["com_unwrap"][:, idx][:, :, [0, 1]]
    want = np.array([port._frame_msd(r[b:b + 1], r[a:a + 1])[0] for a, b in pairs])
    got = e.msd_pairs(idx, [a for a, _ in pairs], [b for _, b in pairs]).cpu().numpy()
    rel_close(got, want, COM_RTOL)


RDF_SETTINGS = [("all", "all", 3, 25, 0.0, 25.0), ("A0", "A0", 1, 80, 0.0, 40.0),
                ("A0", "A1", 3, 30, 2.0, 32.0), ("all", "A1", 2, 25, 0.0, 25.0),
                ("A1", "all", 3, 7, 1.0, 9.0)]


@pytest.mark.parametrize("name", ["tiny_all", "c1_sample", "c3_small", "name_dict"])
def test_rdf_counts_match_port(name):
    e, top, traj, fr, layout = run_engine(name)
    reps = port_reps(name)
    names, codes = type_codes(layout.types)
    d_codes = torch.from_numpy(codes).cuda()
    for (ra, rb, mask, nb, lo, hi) in RDF_SETTINGS:
        def resolve(r):
            if r == "all":
                return "all", -1
            i = int(r[1:]) % len(names)
            return names[i], i
        (na, ca), (nbn, cb) = resolve(ra), resolve(rb)
        leaflet = {3: "both", 1: "upper", 2: "lower"}[mask]
        (g, bins), want = port.rdf(reps["com"], reps["labels"], reps["types"], reps["box"], leaflet,
                                   na, nbn, nb, lo, hi)
        edges = np.histogram([-1], bins=nb, range=[lo, hi])[1]
        count, fc = e.rdf(d_codes, ca, cb, mask, edges)
        got = count.cpu().numpy()
        assert np.array_equal(got, want), (name, ra, rb, mask, np.abs(got - want).max())
        count_d, fc_d = e.rdf(d_codes, ca, cb, mask, edges, dense=True)     # dense fallback kernel
        assert np.array_equal(count_d.cpu().numpy(), want)
        assert np.array_equal(fc.cpu().numpy(), fc_d.cpu().numpy())
        assert want.sum() > 0 or name == "tiny_all"


@pytest.mark.parametrize("name,chunk,skin", [("c2_prefix", 16, 4.0), ("c2_prefix", 16, 0.7), ("c2_prefix", 7, 2.0),
                                             ("c2_prefix", 64, 8.0), ("c3_small", 4, 4.0), ("c3_small", 3, 0.5),
                                             ("tiny_all", 12, 4.0)])
def test_rdf_pair_list_reuse_matches_dense(name, chunk, skin, monkeypatch):
    """k_rdf4 keeps a Verlet pair list over `chunk` frames (rebuilt when a lipid moved skin/2):
    counts must equal the dense sqrt kernel (itself pinned to the port) for any chunk / skin."""
    monkeypatch.setenv("PBK_RDF_CHUNK", str(chunk))
    monkeypatch.setenv("PBK_RDF_SKIN", str(skin))
    e, top, traj, fr, layout = run_engine(name)
    names, codes = type_codes(layout.types)
    d_codes = torch.from_numpy(codes).cuda()
    for (ra, rb, mask, nb, lo, hi) in RDF_SETTINGS:
        ca = -1 if ra == "all" else int(ra[1:]) % len(names)
        cb = -1 if rb == "all" else int(rb[1:]) % len(names)
        edges = np.histogram([-1], bins=nb, range=[lo, hi])[1]
        count, fc = e.rdf(d_codes, ca, cb, mask, edges)
        count_d, fc_d = e.rdf(d_codes, ca, cb, mask, edges, dense=True)
        assert np.array_equal(count.cpu().numpy(), count_d.cpu().numpy()), (name, chunk, skin, ra, rb)
        assert np.array_equal(fc.cpu().numpy(), fc_d.cpu().numpy())


@pytest.mark.parametrize("seq", [False, True])
@pytest.mark.parametrize("name", ["tiny_all", "c1_sample", "c3_small", "name_dict"])
def test_nnf_counts_match_port(name, seq, monkeypatch):
    # seq: the sequential Welford replays (bit-exact restatement); default: the parallel means
    if seq:
        monkeypatch.setenv("PBK_SEQ_WELFORD", "1")
    e, top, traj, fr, layout = run_engine(name)
    reps = port_reps(name)
    names, codes = type_codes(layout.types)
    d_codes = torch.from_numpy(codes).cuda()
    for (ia, ib, mask, k) in [(0, 1, 3, 5), (1, 0, 1, 6), (0, 0, 2, 3), (2, 1, 3, 24)]:
        ia, ib = ia % len(names), ib % len(names)
        leaflet = {3: "both", 1: "upper", 2: "lower"}[mask]
        res, counts = port.nnf(reps["com"], reps["labels"], reps["types"], reps["box"], reps["times"],
                               leaflet, names[ia], names[ib], k, return_counts=True)
        n_b, frac = e.nnf(d_codes, ia, ib, mask, k)
        n_b = n_b.cpu().numpy()
        lab = reps["labels"]
        for f in range(len(fr)):
            got = np.concatenate([n_b[f][(lab[f] == 1) & (n_b[f] >= 0)],
                                  n_b[f][(lab[f] == 0) & (n_b[f] >= 0)]])
            assert np.array_equal(got, counts[f]), (name, f)              # bit-exact neighbour counts
        # parallel frame mean: the sequential Welford recurrence up to its rounding (PBK_SEQ_WELFORD=1: 1e-15)
        rel_close(frac.cpu().numpy(), res[:, 1], 1e-15 if seq else 1e-12)


@pytest.mark.parametrize("name,nx", [("tiny_all", 8), ("grid_dense", 16), ("grid_sparse", 8),
                                     ("c3_small", 32)])
@pytest.mark.parametrize("mode", ["opt", "exact", "opt-seq"])
def test_lipid_grid_matches_port(name, nx, mode, monkeypatch):
    seq = mode.endswith("-seq")
    if seq:
        monkeypatch.setenv("PBK_SEQ_WELFORD", "1")
        mode = "opt"
    e, top, traj, fr, layout = run_engine(name)
    reps = port_reps(name)
    names, codes = type_codes(layout.types)
    d_codes = torch.from_numpy(codes).cuda()
    gi, gz, knn = e.lipid_grid(nx, nx, mode)
    cells, thick, apl = e.grid_reduce(gi, gz, d_codes, len(names))
    gi, gz = gi.cpu().numpy(), gz.cpu().numpy()
    cells, thick, apl = cells.cpu().numpy(), thick.cpu().numpy(), apl.cpu().numpy()
    gold = load(name)
    for f in range(len(fr)):
        g = port.lipid_grid_frame(reps["com"][f], reps["com_unwrap"][f], reps["labels"][f],
                                  reps["box"][f], nx, nx, mode)
        for li, leaf in enumerate(("upper", "lower")):
            assert np.array_equal(gi[f, li], g[leaf][0]), (name, f, leaf)   # bit-exact cell indices
            assert np.array_equal(gz[f, li], g[leaf][1])
            key = f"rep/grid_{leaf}" if mode == "opt" else f"rep/grid_exact_{leaf}"
            if key in gold and gold[key].shape[1] == nx:
                assert np.array_equal(gi[f, li], gold[key][f])              # literal reference
        t = port.grid_thickness(g)
        rel_close(thick[f], np.array(t), 1e-12)
        avg, by_res, area_of = port.grid_apl(g, reps["labels"][f], reps["types"])
        rel_close(apl[f, 0], avg, 1e-15 if seq else 1e-12)      # parallel Welford / sequential replay
        for ti, nm in enumerate(names):
            rel_close(apl[f, 1 + 2 * ti], by_res[nm][0], 1e-15 if seq else 1e-12)
            rel_close(apl[f, 2 + 2 * ti], by_res[nm][1], 1e-12)
        cnt_want = np.zeros(len(layout.types), dtype=np.int64)
        for li, leaf in enumerate(("upper", "lower")):
            cnt_want += np.bincount(g[leaf][0].reshape(-1), minlength=len(cnt_want))
        assert np.array_equal(cells[f], cnt_want)


def test_rdf_out_of_box_coordinates_follow_reference_formula():
    """Lipid COMs outside [0, L] (unwrapped input fed as wrapped): the reference's
    `box - |a'| - |b'|` branch gives odd values; the kernel must reproduce them, not the
    geometric minimum image (the cell list / prefilter is bypassed for such frames)."""
    e, top, traj, fr, layout = run_engine("c3_small")
    reps = port_reps("c3_small")
    names, codes = type_codes(layout.types)
    d_codes = torch.from_numpy(codes).cuda()
    com = reps["com"].copy()
    L = float(reps["box"][0, 0])
    com[1, ::7, 0] += L            # frame 1: every 7th lipid one image to the right
    com[2, ::5, 1] -= 0.75 * L     # frame 2: some lipids below the box
    e.com.copy_(torch.from_numpy(com).cuda())
    for (na, ca, nbn, cb, mask, nb, lo, hi) in [("all", -1, "all", -1, 3, 25, 0.0, 25.0),
                                                 (names[0], 0, names[1], 1, 3, 80, 0.0, 40.0)]:
        leaflet = {3: "both", 1: "upper", 2: "lower"}[mask]
        (g, bins), want = port.rdf(com, reps["labels"], reps["types"], reps["box"], leaflet,
                                   na, nbn, nb, lo, hi)
        edges = np.histogram([-1], bins=nb, range=[lo, hi])[1]
        count, fc = e.rdf(d_codes, ca, cb, mask, edges)
        assert np.array_equal(count.cpu().numpy(), want)


def test_division_by_reused_mass_is_correctly_rounded():
    """pbk_div_r (reciprocal once per lipid + two FMA corrections) against the device's own correctly
    rounded division on 2^27 pseudo-random operand pairs: not one differs."""
    from pybilt_b200 import _native
    ctx = _native.Context(0)
    bad = torch.zeros(1, dtype=torch.int64, device="cuda")
    for seed in (1, 2):
        ctx.call("pbk_debug_div_check", seed, 1 << 26, bad.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert int(bad.item()) == 0
