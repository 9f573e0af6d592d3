"""Cell-list kernels for large leaflets (pbk_cells.cu: k_rdf_cl, k_knn_cl) against the golden
fixtures / oracle port (forced onto the small parity cases) and against the dense kernels
(themselves pinned to the port) on mid-size systems where the port would take minutes."""
import numpy as np
import pytest
import torch

from pybilt_b200 import synthetic as syn
from tests import cases
from tests import test_gpu_engine as tge

pytestmark = pytest.mark.gpu


@pytest.fixture
def force_cells(monkeypatch):
    monkeypatch.setenv("PBK_FORCE_CELLS", "1")


@pytest.mark.parametrize("name", ["tiny_all", "c1_sample", "c3_small", "name_dict"])
def test_cells_rdf_matches_port(name, force_cells):
    tge.test_rdf_counts_match_port(name)


@pytest.mark.parametrize("name", ["tiny_all", "c1_sample", "c3_small", "name_dict"])
def test_cells_nnf_matches_port(name, force_cells, monkeypatch):
    tge.test_nnf_counts_match_port(name, False, monkeypatch)


@pytest.mark.parametrize("name,nx", [("tiny_all", 8), ("grid_dense", 16), ("grid_sparse", 8), ("c3_small", 32)])
def test_cells_grid_knn_matches_port(name, nx, force_cells, monkeypatch):
    tge.test_lipid_grid_matches_port(name, nx, "opt", monkeypatch)


def test_cells_rdf_out_of_box_frames_go_dense(force_cells):
    tge.test_rdf_out_of_box_coordinates_follow_reference_formula()


def _mid_engine(n_lipids, n_frames, seed, npt=1e-3, jitter=None):
    from pybilt_b200 import engine as eng
    kw = dict(seed=seed, npt=npt)
    if jitter is not None:
        kw["lattice_jitter"] = jitter
    spec = syn.BilayerSpec(n_lipids, cases.CG3, **kw)
    top, traj = syn.generate(spec, n_frames)
    layout = eng.build_bead_layout(top.masses, top.names, top.resindex, top.resnames, top.resids)
    e = eng.LipidReductionEngine(top.masses, layout, n_frames)
    e.push(torch.from_numpy(traj.positions).cuda(),
           torch.from_numpy(np.ascontiguousarray(traj.dimensions[:, :3])).cuda())
    e.check_status()
    names, codes = tge.type_codes(layout.types)
    return e, names, torch.from_numpy(codes).cuda()


@pytest.mark.parametrize("n_lipids,seed,jitter", [(2048, 21, None), (6000, 22, 0.49), (10000, 23, None)])
def test_cells_match_dense_kernels_mid_size(n_lipids, seed, jitter, monkeypatch):
    e, names, codes = _mid_engine(n_lipids, 3, seed, jitter=jitter)
    settings = [(-1, -1, 3, 25, 0.0, 25.0), (0, 0, 1, 80, 0.0, 40.0), (0, 1, 3, 30, 2.0, 32.0),
                (-1, 2, 2, 25, 0.0, 25.0), (1, -1, 3, 200, 0.0, 60.0)]
    knn_k = [(0, 1, 3, 5), (1, 0, 1, 6), (2, 2, 2, 30)]

    def run_all():
        out = []
        for (ca, cb, mask, nb, lo, hi) in settings:
            edges = np.histogram([-1], bins=nb, range=[lo, hi])[1]
            cnt, fc = e.rdf(codes, ca, cb, mask, edges)
            out += [cnt.cpu().numpy(), fc.cpu().numpy()]
        for (ia, ib, mask, k) in knn_k:
            n_b, frac = e.nnf(codes, ia, ib, mask, k)
            out += [n_b.cpu().numpy(), frac.cpu().numpy()]
        gi, gz, knn = e.lipid_grid(64, 64, "opt")
        out += [knn.cpu().numpy(), gi.cpu().numpy(), gz.cpu().numpy()]
        return out

    monkeypatch.setenv("PBK_FORCE_CELLS", "1")
    got = run_all()
    monkeypatch.delenv("PBK_FORCE_CELLS")
    monkeypatch.setenv("PBK_NO_CELLS", "1")
    want = run_all()
    for i, (g, w) in enumerate(zip(got, want)):
        assert np.array_equal(g, w), (n_lipids, i)
    assert got[0].sum() > 0


def test_cells_out_of_box_frames_nnf_and_grid(monkeypatch):
    e, names, codes = _mid_engine(2048, 3, 31)
    L = float(e.box[0, 0])
    com = e.com.clone()
    com[1, ::7, 0] += L
    com[2, ::5, 1] -= 0.75 * L
    e.com.copy_(com)

    def run_all():
        n_b, frac = e.nnf(codes, 0, 1, 3, 5)
        gi, gz, knn = e.lipid_grid(32, 32, "opt")
        edges = np.histogram([-1], bins=25, range=[0.0, 25.0])[1]
        cnt, fc = e.rdf(codes, -1, -1, 3, edges)
        return [n_b.cpu().numpy(), frac.cpu().numpy(), knn.cpu().numpy(), gi.cpu().numpy(),
                cnt.cpu().numpy(), fc.cpu().numpy()]

    monkeypatch.setenv("PBK_FORCE_CELLS", "1")
    got = run_all()
    monkeypatch.delenv("PBK_FORCE_CELLS")
    monkeypatch.setenv("PBK_NO_CELLS", "1")
    want = run_all()
    for i, (g, w) in enumerate(zip(got, want)):
        assert np.array_equal(g, w), i


@pytest.mark.parametrize("n_lipids,oob", [(512, False), (2048, False), (6000, True)])
def test_rdf_multi_tables_equal_per_pair_calls(n_lipids, oob):
    """pbk_rdf_multi: one sweep, one table per (leaflet, type_i, type_j); every table must equal the
    per-analysis kernel's histogram, and type_counts its (N_a, N_b, N_leaflet) rows."""
    e, names, codes = _mid_engine(n_lipids, 3, 41)
    if oob:
        L = float(e.box[0, 0])
        com = e.com.clone()
        com[1, ::7, 0] += L
        e.com.copy_(com)
    T = len(names)
    for nb, lo, hi in [(25, 0.0, 25.0), (80, 0.0, 40.0), (30, 2.0, 32.0)]:
        edges = np.histogram([-1], bins=nb, range=[lo, hi])[1]
        table, tc = e.rdf_multi(codes, T, edges)
        table, tc = table.cpu().numpy(), tc.cpu().numpy()
        for leaf in range(2):
            for a in range(T):
                for b in range(T):
                    cnt, fc = e.rdf(codes, a, b, 1 << leaf, edges)
                    assert np.array_equal(table[leaf, a, b], cnt.cpu().numpy()), (nb, leaf, a, b)
                    fc = fc.cpu().numpy()
                    on = (tc[:, leaf, a] > 0) & (tc[:, leaf, b] > 0)
                    assert np.array_equal(fc[:, leaf, 0], np.where(on, tc[:, leaf, a], 0))
                    assert np.array_equal(fc[:, leaf, 1], np.where(on, tc[:, leaf, b], 0))
                    assert np.array_equal(fc[:, leaf, 2], tc[:, leaf].sum(axis=1))
        cnt_all, _ = e.rdf(codes, -1, -1, 3, edges)
        assert np.array_equal(table.sum(axis=(0, 1, 2)), cnt_all.cpu().numpy())


@pytest.mark.parametrize("n_lipids,oob", [(256, False), (2048, False), (3000, True)])
def test_nnf_multi_tallies_equal_per_analysis_calls(n_lipids, oob):
    """pbk_nnf_multi + pbk_nnf_select (one neighbour search, types tallied) against pbk_nnf."""
    e, names, codes = _mid_engine(n_lipids, 3, 43)
    if oob:
        L = float(e.box[0, 0])
        com = e.com.clone()
        com[1, ::7, 0] += L
        e.com.copy_(com)
    T = len(names)
    for k in (5, 6, 30):
        hits, tc = e.nnf_multi(codes, T, k)
        for (ta, tb, mask) in [(0, 1, 3), (1, 0, 1), (2, 2, 2), (-1, 1, 3), (0, -1, 3), (1, 5, 3)]:
            n_b, frac = e.nnf_select(hits, tc, codes, ta if ta < T else -2, tb if tb < T else -2, mask, k)
            n_b2, frac2 = e.nnf(codes, ta if ta < T else -2, tb if tb < T else -2, mask, k)
            assert np.array_equal(n_b.cpu().numpy(), n_b2.cpu().numpy()), (k, ta, tb, mask)
            assert np.array_equal(frac.cpu().numpy(), frac2.cpu().numpy())
        specs = [(0, 1, 3), (1, 0, 1), (2, 2, 2), (-1, 1, 3), (0, -1, 3), (-2, 1, 3)]
        nb_all, fr_all = e.nnf_select_many(hits, tc, codes, specs, k)          # all analyses in two launches
        for a, (ta, tb, mask) in enumerate(specs):
            n_b2, frac2 = e.nnf(codes, ta, tb, mask, k)
            assert torch.equal(nb_all[a], n_b2) and torch.equal(fr_all[a], frac2), (k, ta, tb, mask)


def test_halperin_nelson_cells_equal_dense(monkeypatch):
    """The hexatic invariant from the cell-list neighbour search equals the one from the dense
    search (same six neighbours in the same order, hence the same bits)."""
    e, names, codes = _mid_engine(3000, 3, 47)
    member = (e.labels[0] == 1).to(torch.uint8)
    monkeypatch.setenv("PBK_FORCE_CELLS", "1")
    a = e.halperin_nelson(member).cpu().numpy()
    monkeypatch.delenv("PBK_FORCE_CELLS")
    monkeypatch.setenv("PBK_NO_CELLS", "1")
    b = e.halperin_nelson(member).cpu().numpy()
    assert np.array_equal(a, b) and (a > 0).all() and (a < 1).all()
