"""The other named shapes of BASELINE.json (C3 8,192 lipids / C4 65,536 / C5 262,144) at a handful
of frames: the oracle cannot run them in seconds, so the large-system kernels are held to
(a) the dense kernels (pinned to the oracle port on the small cases) on the SAME frames,
(b) batch-split invariance of the generic reduction, (c) the port on a lipid subset where the
algorithm is local (leaflets, per-lipid COMs), (d) streamed == resident through the API."""
import numpy as np
import pytest
import torch

from oracle import port

pytestmark = pytest.mark.gpu


def _engine(cfg, frames, **kw):
    from pybilt_b200 import engine as eng, synthetic as syn
    spec, _ = syn.config_spec(cfg)
    top, traj = syn.generate(spec, frames)
    layout = eng.build_bead_layout(top.masses, top.names, top.resindex, top.resnames, top.resids)
    e = eng.LipidReductionEngine(top.masses, layout, frames, **kw)
    x = torch.from_numpy(traj.positions).cuda()
    box = torch.from_numpy(np.ascontiguousarray(traj.dimensions[:, :3])).cuda()
    e.push(x, box)
    e.check_status()
    names = []
    for t in layout.types:
        if t not in names:
            names.append(t)
    codes = torch.from_numpy(np.array([names.index(t) for t in layout.types], dtype=np.int32)).cuda()
    return e, top, traj, layout, names, codes, x, box


def test_c3_shape_reduction_rdf_nnf(monkeypatch):
    from pybilt_b200 import engine as eng
    F = 6
    e, top, traj, layout, names, codes, x, box = _engine("C3", F)
    assert not e.used_fused                      # 1.1 MB per frame: the generic kernels
    # (b) ragged batches chain through the carried unwrap state
    e2 = eng.LipidReductionEngine(top.masses, layout, F, batch_frames=4)
    e2.push(x[:1], box[:1]); e2.push(x[1:], box[1:]); e2.check_status()
    for a, b in ((e.com, e2.com), (e.com_unwrap, e2.com_unwrap), (e.labels, e2.labels), (e.mem_com, e2.mem_com)):
        assert torch.equal(a, b)
    # (c) the port on the same input (3 frames; membrane COM and unwrap need all atoms)
    sub = type(traj)(traj.positions[:3], traj.dimensions[:3], traj.times[:3], traj.dt)
    _o, reps = port.run(top, sub, [], {}, return_reps=True)
    assert np.array_equal(e.com[:3].cpu().numpy(), reps["com"])
    assert np.array_equal(e.com_unwrap[:3].cpu().numpy(), reps["com_unwrap"])
    assert np.array_equal(e.labels[:3].cpu().numpy(), reps["labels"])
    assert np.array_equal(e.mem_com[:3].cpu().numpy(), reps["mem_com"])
    # (a) cell-list kernels against the dense kernels, prefab settings (80 bins, 40 A; k = 5)
    edges = np.histogram([-1], bins=80, range=[0.0, 40.0])[1]
    T = len(names)
    table, tc = e.rdf_multi(codes, T, edges)
    for a in range(T):
        for b in range(T):
            c_cl, fc_cl = e.rdf(codes, a, b, 3, edges)
            c_d, fc_d = e.rdf(codes, a, b, 3, edges, dense=True)
            assert torch.equal(c_cl, c_d) and torch.equal(fc_cl, fc_d)
            assert torch.equal(table[:, a, b].sum(dim=0), c_d)
    hits, tcn = e.nnf_multi(codes, T, 5)
    for a in range(T):
        for b in range(T):
            n_cl, f_cl = e.nnf(codes, a, b, 3, 5)
            n_s, f_s = e.nnf_select(hits, tcn, codes, a, b, 3, 5)
            monkeypatch.setenv("PBK_NO_CELLS", "1")
            n_d, f_d = e.nnf(codes, a, b, 3, 5)
            monkeypatch.delenv("PBK_NO_CELLS")
            assert torch.equal(n_cl, n_d) and torch.equal(f_cl, f_d)
            assert torch.equal(n_s, n_d) and torch.equal(f_s, f_d)


def test_c4_shape_lipid_grids(monkeypatch):
    F = 2
    e, top, traj, layout, names, codes, x, box = _engine("C4", F)
    gi, gz, knn = e.lipid_grid(128, 128, "opt")
    monkeypatch.setenv("PBK_NO_CELLS", "1")
    gi_d, gz_d, knn_d = e.lipid_grid(128, 128, "opt")
    monkeypatch.delenv("PBK_NO_CELLS")
    assert torch.equal(knn, knn_d)               # 24 nearest neighbours of 65,536 lipids, order included
    assert torch.equal(gi, gi_d) and torch.equal(gz, gz_d)
    cells, thick, apl = e.grid_reduce(gi, gz, codes, len(names))
    assert int(cells.sum()) == F * 2 * 128 * 128
    # every cell's lipid belongs to the cell's leaflet; APL * lipids = box area (both leaflets)
    lab = e.labels
    for f in range(F):
        assert bool((lab[f][gi[f, 0].reshape(-1).long()] == 1).all())
        assert bool((lab[f][gi[f, 1].reshape(-1).long()] == 0).all())
    area = (e.box[:, 0].double() * e.box[:, 1].double()).cpu().numpy()
    assert np.allclose(apl[:, 0].cpu().numpy() * layout.n_beads, 2 * area, rtol=1e-6)
    # the exact grid (lipid_grid.py) agrees with the reference's heuristic one (lipid_grid_opt.py)
    # wherever the heuristic found the true nearest lipid: about 80 % of the cells at 2 lipids per cell
    ge, _z, _k = e.lipid_grid(128, 128, "exact", frames=slice(0, 1))
    assert float((ge[0] == gi[0]).double().mean()) > 0.7


def test_c5_shape_msd_rdf_and_streamed_api():
    F = 3
    e, top, traj, layout, names, codes, x, box = _engine("C5", F)
    edges = np.histogram([-1], bins=25, range=[0.0, 25.0])[1]
    c_cl, fc_cl = e.rdf(codes, -1, -1, 3, edges)
    c_d, fc_d = e.rdf(codes, -1, -1, 3, edges, frames=slice(0, 1), dense=True)     # 131 ms per frame
    c_1, _ = e.rdf(codes, -1, -1, 3, edges, frames=slice(0, 1))
    assert torch.equal(c_1, c_d)
    a, _ = e.rdf(codes, -1, -1, 3, edges, frames=slice(1, F))
    assert torch.equal(c_1 + a, c_cl)            # frame additivity
    assert int(c_cl.sum()) % 2 == 0
    lab = e.labels.cpu().numpy()
    assert (lab.sum(axis=1) == layout.n_beads // 2).all()
    # streamed run through the API == resident run
    from pybilt_b200 import BilayerAnalyzer
    outs = []
    for window in (None, 2):
        ba = BilayerAnalyzer(structure=top, trajectory=traj, selection="all")
        ba.add_analysis("com_lateral_rdf rdf resname_1 all resname_2 all")
        ba.window_frames = window
        ba.run_analysis()
        outs.append((ba.get_analysis_data("msd_1"), ba.get_analysis_data("rdf")[0]))
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
    idx = np.concatenate([np.nonzero(lab[0] == 1)[0], np.nonzero(lab[0] == 0)[0]]).astype(np.int32)
    m = e.msd_pairs(idx, np.zeros(F, dtype=np.int32), np.arange(F, dtype=np.int32)).cpu().numpy()
    assert np.array_equal(m, outs[0][0][:, 1])


# ---------------------------------------------------------------------------------------------
# The named shapes against the ORACLE (oracle/port.py + portc.c) on real frames of each config --
# not against another GPU kernel.  The oracle's large-shape loops (cell-list RDF, matrix-free
# 24-neighbour table) are pinned to its dense loops in tests/test_oracle_golden.py.
# ---------------------------------------------------------------------------------------------
def _port_reps(top, traj, n):
    sub = type(traj)(traj.positions[:n], traj.dimensions[:n], traj.times[:n], traj.dt)
    _o, reps = port.run(top, sub, [], {}, return_reps=True)
    return reps, np.ascontiguousarray(sub.dimensions[:, :3]), sub.times


def test_c3_shape_rdf_nnf_against_oracle():
    """C3 (8,192 lipids, POPC/DOPE/CHOL): 2 frames, prefab settings -- com_lateral_rdf 80 bins / 40 A
    and nnf k = 5 (prefab_analysis_protocols.py:1517-1521, 886-890): counts bit-exact."""
    F = 2
    e, top, traj, layout, names, codes, x, box = _engine("C3", F)
    reps, boxh, times = _port_reps(top, traj, F)
    types = list(layout.types)
    assert np.array_equal(e.com.cpu().numpy(), reps["com"])
    assert np.array_equal(e.labels.cpu().numpy(), reps["labels"])
    edges = np.histogram([-1], bins=80, range=[0.0, 40.0])[1]
    table, tc = e.rdf_multi(codes, len(names), edges)
    for (ra, rb, leaflet, mask) in (("POPC", "POPC", "both", 3), ("POPC", "DOPE", "upper", 1),
                                    ("CHOL", "all", "lower", 2), ("all", "all", "both", 3)):
        (_g, _b), want = port.rdf(reps["com"], reps["labels"], types, boxh, leaflet, ra, rb, 80, 0.0, 40.0)
        ca = -1 if ra == "all" else names.index(ra)
        cb = -1 if rb == "all" else names.index(rb)
        got, _fc = e.rdf(codes, ca, cb, mask, edges)
        assert want.sum() > 0 and np.array_equal(got.cpu().numpy(), want), (ra, rb, leaflet)
        if ca >= 0 and cb >= 0:          # the shared-sweep tables hold the same counts
            rows = [0, 1] if mask == 3 else [0 if mask == 1 else 1]
            assert np.array_equal(table[rows, ca, cb].sum(dim=0).cpu().numpy(), want)
    lab = reps["labels"]
    for (ra, rb, leaflet, mask) in (("POPC", "DOPE", "both", 3), ("CHOL", "POPC", "lower", 2)):
        res, counts = port.nnf(reps["com"], lab, types, boxh, times, leaflet, ra, rb, 5, return_counts=True)
        n_b, frac = e.nnf(codes, names.index(ra), names.index(rb), mask, 5)
        n_b = n_b.cpu().numpy()
        for f in range(F):
            got = np.concatenate([n_b[f][(lab[f] == 1) & (n_b[f] >= 0)], n_b[f][(lab[f] == 0) & (n_b[f] >= 0)]])
            assert np.array_equal(got, counts[f]), (ra, rb, f)
        assert np.allclose(frac.cpu().numpy(), res[:, 1], rtol=1e-14, atol=0)


def test_c4_shape_grids_against_oracle():
    """C4 (65,536 lipids): one frame, both leaflets' 128 x 128 lipid_grid_opt grids -- 24-neighbour
    tables, cell indices and cell z bit-exact; APL / thickness rows to 1e-12."""
    e, top, traj, layout, names, codes, x, box = _engine("C4", 2)
    reps, boxh, _t = _port_reps(top, traj, 2)
    f = 1
    assert np.array_equal(e.com_unwrap.cpu().numpy(), reps["com_unwrap"])
    gi, gz, knn = e.lipid_grid(128, 128, "opt", frames=slice(f, f + 1))
    g = port.lipid_grid_frame(reps["com"][f], reps["com_unwrap"][f], reps["labels"][f], boxh[f], 128, 128,
                              "opt", lean_knn=True)
    knn = knn[0].cpu().numpy()
    for li, (leaf, lab) in enumerate((("upper", 1), ("lower", 0))):
        members = np.nonzero(reps["labels"][f] == lab)[0]
        assert np.array_equal(knn[members].astype(np.int64), g["knn_" + leaf]), leaf
        assert np.array_equal(gi[0, li].cpu().numpy(), g[leaf][0]), leaf
        assert np.array_equal(gz[0, li].cpu().numpy(), g[leaf][1]), leaf
    cells, thick, apl = e.grid_reduce(gi, gz, codes, len(names), frames=slice(f, f + 1))
    t = port.grid_thickness(g)
    assert np.allclose(thick[0].cpu().numpy(), np.array(t), rtol=1e-12, atol=0)
    avg, by_res, _area = port.grid_apl(g, reps["labels"][f], list(layout.types))
    apl = apl[0].cpu().numpy()
    assert abs(apl[0] - avg) <= 1e-12 * abs(avg)
    for ti, nm in enumerate(names):
        assert abs(apl[1 + 2 * ti] - by_res[nm][0]) <= 1e-12 * abs(by_res[nm][0])
        assert abs(apl[2 + 2 * ti] - by_res[nm][1]) <= 1e-9 * abs(by_res[nm][1])


def test_c5_shape_rdf_msd_against_oracle():
    """C5 (262,144 lipids): com_lateral_rdf all-all (25 bins, 25 A, both leaflets) of one frame
    bit-exact against the oracle's cell-list loop; the msd rows of 3 frames to 1e-9."""
    F = 3
    e, top, traj, layout, names, codes, x, box = _engine("C5", F)
    reps, boxh, times = _port_reps(top, traj, F)
    types = list(layout.types)
    assert np.array_equal(e.com.cpu().numpy(), reps["com"])
    assert np.array_equal(e.com_unwrap.cpu().numpy(), reps["com_unwrap"])
    assert np.array_equal(e.labels.cpu().numpy(), reps["labels"])
    edges = np.histogram([-1], bins=25, range=[0.0, 25.0])[1]
    (_g, _b), want = port.rdf(reps["com"][1:2], reps["labels"][1:2], types, boxh[1:2], "both", "all", "all",
                              25, 0.0, 25.0, cells=True)
    got, _fc = e.rdf(codes, -1, -1, 3, edges, frames=slice(1, 2))
    assert want.sum() > 7_000_000 and np.array_equal(got.cpu().numpy(), want)
    lab0 = reps["labels"][0]
    idx = np.concatenate([np.nonzero(lab0 == 1)[0], np.nonzero(lab0 == 0)[0]]).astype(np.int32)
    want_msd = port.msd(reps["com_unwrap"], times, idx)
    got_msd = e.msd_pairs(idx, np.zeros(F, dtype=np.int32), np.arange(F, dtype=np.int32)).cpu().numpy()
    assert np.allclose(got_msd, want_msd[:, 1], rtol=1e-9, atol=0)
