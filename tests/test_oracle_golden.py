"""The oracle port (oracle/port.py) against outputs of the literal reference executed on the
same synthetic inputs (tests/golden/*.npz, made by oracle/make_golden.py)."""
import os

import numpy as np
import pytest

from oracle import port
from tests import cases

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
FLOAT_RTOL = 1e-12      # observed: bit-identical in the build container


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))


def close(a, b, rtol=FLOAT_RTOL):
    if np.asarray(b).dtype.kind in "USO":            # names: equal, not close
        assert np.array_equal(np.asarray(a).astype(str), np.asarray(b).astype(str))
        return
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape
    both_nan = np.isnan(a) & np.isnan(b)
    both_inf = np.isinf(a) & np.isinf(b) & (np.sign(a) == np.sign(b))
    ok = both_nan | both_inf | (np.abs(a - b) <= rtol * np.maximum(np.abs(a), np.abs(b)))
    assert ok.all(), f"max abs diff {np.nanmax(np.abs(a - b))}"


@pytest.mark.parametrize("name", list(cases.CASES))
def test_port_matches_reference(name):
    c = cases.CASES[name]
    gold = load(name)
    top, traj = cases.make_inputs(name)
    assert cases.input_digest(traj) == str(gold["meta/input_sha256"]), "synthetic generator drifted"
    out, reps = port.run(top, traj, c["analyses"], c.get("rep_settings"), c.get("settings"),
                         c.get("frame_range", (0, -1, 1)), return_reps=True)
    # representations
    close(reps["com"], gold["rep/com"])
    close(reps["com_unwrap"], gold["rep/com_unwrap"])
    close(reps["mem_com"], gold["rep/mem_com"])
    assert np.array_equal(reps["labels"], gold["rep/labels"])            # bit-exact labels
    assert np.array_equal(reps["u_last"], gold["rep/u_last"])            # bit-exact f32 unwrap
    assert list(reps["types"]) == list(gold["rep/types"])
    close(reps["times"], gold["rep/time"])
    if "rep/grid_upper" in gold:
        for f, g in enumerate(reps["grids"]):
            for leaf in ("upper", "lower"):
                assert np.array_equal(g[leaf][0], gold[f"rep/grid_{leaf}"][f])   # bit-exact cells
                close(g[leaf][1], gold[f"rep/gridz_{leaf}"][f])
    # analysis outputs
    flat = cases.flatten_outputs(out)
    keys = [k for k in gold if k.startswith("out/")]
    assert keys
    for k in keys:
        close(flat[k], gold[k])


@pytest.mark.parametrize("name", ["grid_dense", "grid_sparse", "tiny_all"])
def test_port_exact_grid_matches_lipid_grid_py(name):
    c = cases.CASES[name]
    gold = load(name)
    top, traj = cases.make_inputs(name)
    out, reps = port.run(top, traj, c["analyses"], c.get("rep_settings"), return_reps=True,
                         grid_mode="exact")
    for f, g in enumerate(reps["grids"]):
        for leaf in ("upper", "lower"):
            assert np.array_equal(g[leaf][0], gold[f"rep/grid_exact_{leaf}"][f])


def test_opt_and_exact_grids_differ_when_sparse():
    """at <1 cell per lipid the heuristic grid is not the exact argmin."""
    gold = load("grid_sparse")
    diff = sum((gold[f"rep/grid_{l}"] != gold[f"rep/grid_exact_{l}"]).sum() for l in ("upper", "lower"))
    total = sum(gold[f"rep/grid_{l}"].size for l in ("upper", "lower"))
    assert 0 <= diff <= total


def test_pairwise_sum_is_numpy_order():
    rng = np.random.default_rng(0)
    for n in [0, 1, 5, 7, 8, 9, 15, 16, 17, 100, 127, 128, 129, 130, 255, 256, 257, 1000, 6144,
              7080, 12345, 91750, 100003]:
        a = rng.standard_normal(n) * 10.0 ** rng.integers(-3, 6, n)
        assert port.pairwise_sum(a) == a.sum(), n


def test_msd_multi_block_schedule():
    pairs, n_orig = port.msd_multi_blocks(20, 3, 3)
    assert pairs == [(0, 2), (3, 5), (6, 8), (9, 11), (12, 14), (15, 17)]
    pairs, _ = port.msd_multi_blocks(12, 4, 2)
    assert pairs == [(0, 3), (2, 5), (4, 7), (6, 9), (8, 11)]


def test_histogram_rule_is_numpy():
    """pc_rdf_accumulate's bin rule == np.histogram on values sitting on / next to edges."""
    import ctypes
    lo, hi, nb = 0.0, 40.0, 80
    edges = np.histogram([-1], bins=nb, range=[lo, hi])[1]
    vals = np.concatenate([edges, np.nextafter(edges, -1), np.nextafter(edges, 100),
                           np.random.default_rng(1).uniform(-1, 41, 5000)])
    want = np.histogram(vals, bins=nb, range=[lo, hi])[0]
    # two points per value on a line so that each ordered pair (0, j) has r = vals[j]
    pos = np.zeros((len(vals) + 1, 2))
    pos[1:, 0] = np.abs(vals)
    ok = vals >= 0
    ia = np.zeros(1, dtype=np.int64)
    ib = np.arange(1, len(vals) + 1, dtype=np.int64)[ok]
    count = np.zeros(nb, dtype=np.int64)
    port.lib().pc_rdf_accumulate(port._p(pos), port._p(ia), ctypes.c_int64(1), port._p(ib),
                                 ctypes.c_int64(len(ib)), ctypes.c_float(1000.0), ctypes.c_float(1000.0),
                                 ctypes.c_int(nb), ctypes.c_double(lo), ctypes.c_double(hi),
                                 port._p(edges), port._p(count))
    centre = np.float32(1000.0) / 2.0
    # distances pass through (x - c) - (0 - c): recompute what the routine saw
    seen = np.abs((pos[ib, 0] - float(centre)) - (0.0 - float(centre)))
    want = np.histogram(seen, bins=nb, range=[lo, hi])[0]
    assert np.array_equal(count, want)


@pytest.mark.parametrize("name", ["big_steps", "name_dict", "c2_prefix"])
def test_port_comtraj_matches_reference(name):
    """oracle.port.comtraj against the literal pybilt.com_trajectory.COMTraj (scalar float32
    unwrap, sequential membrane COM, leaflets from frame 0)."""
    gold = load("comtraj_" + name)
    top, traj = cases.make_inputs(name)
    assert cases.input_digest(traj) == str(gold["meta/input_sha256"])
    ct = port.comtraj(top, traj)
    close(ct["com"], gold["com"])
    close(ct["com_unwrap"], gold["com_unwrap"])
    assert np.array_equal(ct["labels0"], gold["labels0"])
    close(port.comtraj_msd(ct), gold["msd_both"])
    close(port.comtraj_msd(ct, "upper"), gold["msd_upper"])
    close(port.comtraj_msd(ct, "lower", str(gold["first_group"])), gold["msd_lower_first"])
    assert np.array_equal(port.comtraj_group_indices(ct["labels0"], ct["types"], "upper", "all"),
                          gold["upper_all"])


def test_port_leaflet_distance_raises_with_update_interval():
    top, traj = cases.make_inputs("leaflet_dist")
    with pytest.raises(IndexError):
        port.run(top, traj, ["thickness_leaflet_distance tld"], {"leaflets": {"update_interval": 3}})


def test_large_shape_oracle_variants_equal_the_dense_ones():
    """portc.c's cell-list RDF loop and matrix-free 24-neighbour table (used for the named shapes
    C3-C5 in tests/test_gpu_named_shapes.py) are the dense loops reorganised: identical outputs."""
    from pybilt_b200 import synthetic as syn
    spec = syn.BilayerSpec(2048, cases.CG3, seed=17)
    top, traj = syn.generate(spec, 2)
    _o, reps = port.run(top, traj, [], {}, return_reps=True)
    types = list(top.resnames)
    box = np.ascontiguousarray(traj.dimensions[:, :3])
    for kw in (dict(resname_1="all", resname_2="all", n_bins=25, range_outer=25.0),
               dict(resname_1="POPC", resname_2="DOPE", leaflet="lower", n_bins=80, range_outer=40.0),
               dict(resname_1="CHOL", resname_2="all", n_bins=30, range_inner=2.0, range_outer=31.0)):
        (_g, _b), dense = port.rdf(reps["com"], reps["labels"], types, box, **kw)
        (_g, _b), cl = port.rdf(reps["com"], reps["labels"], types, box, cells=True, **kw)
        assert dense.sum() > 0 and np.array_equal(dense, cl)
    a = port.lipid_grid_frame(reps["com"][1], reps["com_unwrap"][1], reps["labels"][1], box[1], 32, 32)
    b = port.lipid_grid_frame(reps["com"][1], reps["com_unwrap"][1], reps["labels"][1], box[1], 32, 32,
                              lean_knn=True)
    for k in ("knn_upper", "knn_lower"):
        assert np.array_equal(a[k], b[k])
    for k in ("upper", "lower"):
        assert np.array_equal(a[k][0], b[k][0]) and np.array_equal(a[k][1], b[k][1])
    # a COM outside the box: the cell-list form refuses instead of miscounting
    com = reps["com"].copy()
    com[0, 0, 0] = -1.0
    with pytest.raises(ValueError):
        port.rdf(com, reps["labels"], types, box, resname_1="all", resname_2="all", cells=True)
