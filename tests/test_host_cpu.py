"""CPU-only checks: the C ABI loads and exports what include/pbk.h declares, the host mirror of
the plugin interface behaves like the reference's, and the product refuses to run without a
CUDA device (no CPU fallback)."""
import os
import re
import warnings

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "pbk.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pbk_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from pybilt_b200 import _native
    lib = _native.load()                      # dlopen only, no CUDA call
    syms = declared_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), s
    assert sorted(_native.EXPORTS) == syms    # the ctypes stub binds exactly the header
    assert lib.pbk_abi_version() == 8


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from pybilt_b200 import BilayerAnalyzer
    from tests import cases
    top, traj = cases.make_inputs("tiny_all")
    ba = BilayerAnalyzer(structure=top, trajectory=traj, selection="all")
    with pytest.raises(RuntimeError):
        ba.run_analysis()


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "pybilt_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), fn


def test_registry_and_argument_parsing():
    from pybilt_b200 import analysis_protocols as ap
    assert set(ap.valid_analysis) == {"msd", "msd_multi", "com_lateral_rdf", "nnf", "apl_grid",
                                      "thickness_grid", "thickness_leaflet_distance", "area_fluctuation",
                                      "halperin_nelson", "apl_box", "acm", "ald", "ac", "vcm", "flip_flop",
                                      "disp_vec", "msd_dist", "dc_cluster"}
    assert ap.analysis_obj_name_dict["apl_grid"] == "lipid_grid"
    a = ap.Analyses(["msd msd_1", "com_lateral_rdf r n_bins 80 range_outer 40.0 leaflet upper"])
    assert a.analysis_ids == ["msd_1", "r"] and not a.use_objects["lipid_grid"]
    r = a["r"]
    assert r.settings["n_bins"] == 80 and r.settings["range_outer"] == 40.0
    assert r.settings["resname_1"] == "first" and r.save_file_name == "r.pickle"
    a.add_analysis(["nnf", "n1", {"n_neighbors": 6}])
    a.add_analysis({"analysis_key": "thickness_grid", "analysis_id": "bt", "analysis_settings": {}})
    assert a.use_objects["lipid_grid"] and len(a) == 4
    with pytest.raises(RuntimeError):
        a.add_analysis("msd msd_1")                       # duplicate id
    with pytest.raises(RuntimeError):
        a.add_analysis("not_an_analysis x")
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        a.add_analysis("msd m2 bogus_key 3")
        assert any("invalid argument key" in str(x.message) for x in w)
    a.remove_analysis("m2")
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        a.remove_analysis("m2")
        assert w
    a.remove_all()
    assert len(a) == 0


def test_input_script_and_frame_range(tmp_path):
    from pybilt_b200 import BilayerAnalyzer
    from tests import cases
    top, traj = cases.make_inputs("tiny_all")
    ba = BilayerAnalyzer(input_dict={"structure": top, "trajectory": traj, "selection": "all",
                                     "analysis": ["msd m", "nnf n n_neighbors 4"],
                                     "frames": ["first", 2, "last", -3, "interval", 2],
                                     "lipid_grid": ["n_xbins", 12, "n_ybins", 12]})
    assert ba.settings["frame_range"] == [2, -3, 2]
    assert ba.rep_settings["lipid_grid"]["n_xbins"] == 12
    assert list(ba.analysed_frame_indices()) == [2, 4, 6, 8]
    assert ba.get_analysis_ids() == ["m", "n"]
    ba.set_frame_range(0.25, -1, 1)
    assert ba.settings["frame_range"][0] == 3
    with pytest.raises(RuntimeError):
        BilayerAnalyzer(input_dict={"structure": top})
    with pytest.raises(RuntimeError):
        BilayerAnalyzer()
    script = tmp_path / "setup.in"
    script.write_text("# comment\nstructure a.psf\ntrajectory b.dcd\nselection 'resname POPC'\n"
                      "analysis msd msd_1\nanalysis nnf n1 n_neighbors 3\nframes first 1 last 5\n")
    cmds = BilayerAnalyzer.parse_input_script(BilayerAnalyzer.__new__(BilayerAnalyzer), str(script))
    assert cmds["selection"] == "resname POPC" and cmds["analysis"] == ["msd msd_1", "nnf n1 n_neighbors 3"]
    bad = tmp_path / "bad.in"
    bad.write_text("structure a\ntrajectory b\nselection 'all'\nbogus x y\n")
    with pytest.raises(RuntimeError):
        BilayerAnalyzer.parse_input_script(BilayerAnalyzer.__new__(BilayerAnalyzer), str(bad))


def test_selection_of_whole_residues():
    from pybilt_b200.bilayer_analyzer import TrajectoryData
    from tests import cases
    top, traj = cases.make_inputs("tiny_all")
    md = TrajectoryData(top, traj, "resname POPC DOPE")
    keep = [r for r in top.resnames if r in ("POPC", "DOPE")]
    assert md.resnames == keep and md.n_residues == len(keep)
    x, d, t = md.frames(np.array([0, 1, 2]))
    assert x.shape == (3, md.natoms, 3) and md.natoms < top.n_atoms
    assert np.array_equal(x[1], traj.positions[1][md.indices])


def test_msd_multi_schedule_matches_the_oracle():
    from oracle import port
    from pybilt_b200.analysis_protocols import msd_multi_schedule
    for (F, nt, ns) in [(20, 3, 3), (12, 4, 2), (97, 10, 4), (5000, 50, 50), (30, 50, 50), (9, 1, 1)]:
        fin, nb = msd_multi_schedule(F, nt, ns)
        pairs, n_orig = port.msd_multi_blocks(F, nt, ns)
        assert [fin[b] for b in sorted(fin)] == pairs and nb == n_orig


def test_running_stats_float32_promotion():
    """pybilt/common/running_stats.py: the first push keeps float32 (RDF area, APL)."""
    from oracle.port import Welford
    from pybilt_b200.running_stats import RunningStats
    rng = np.random.default_rng(3)
    vals = (rng.uniform(60, 70, 50).astype(np.float32))
    a, b = RunningStats(), Welford()
    for v in vals:
        a.push(v); b.push(v)
        assert a.mean() == b.mean() and a.deviation() == b.deviation()
    assert RunningStats().mean() == 0.0


def test_bead_layout_matches_the_oracle():
    from oracle import port
    from pybilt_b200.engine import build_bead_layout
    from tests import cases
    top, _ = cases.make_inputs("name_dict")
    nd = cases.CASES["name_dict"]["rep_settings"]["com_frame"]["name_dict"]
    for multi in (False, True):
        lay = build_bead_layout(top.masses, top.names, top.resindex, top.resnames, top.resids, nd, multi)
        g, seg, types, resids = port.build_beads(top, nd, multi)
        assert np.array_equal(lay.gather, g) and np.array_equal(lay.seg, seg) and lay.types == types
        assert np.array_equal(lay.bead_mass, port.bead_masses(top.masses, g, seg))


def test_diffusion_estimators():
    """tests/test_diffusion_coefficent_estimation.py of the reference: MSD = 4 D t + noise."""
    from pybilt_b200 import diffusion as dc
    t = np.linspace(0.0, 100.0, 400)
    rng = np.random.default_rng(0)
    msd = 4.0 * 2.0 * t + rng.normal(0, 0.05, t.size)
    assert abs(dc.diffusion_coefficient_Einstein(t, msd) - 2.0) < 0.01
    d, err = dc.diffusion_coefficient_linear_fit(t, msd)
    assert abs(d - 2.0) < 0.01 and err < 0.01
    p = dc.diffusion_coefficient_anomalous_fit(t[1:], msd[1:] - msd[1] + 4.0 * 2.0 * 0.0)
    assert abs(p[1] - 1.0) < 0.05
    t0 = t.copy()
    dc.diffusion_coefficient_anomalous_fit(t0[1:], msd[1:])
    assert np.array_equal(t0, t)              # unlike the reference, the input is left alone


def test_plot_protocol_registry_matches_reference_strings():
    """add_plot / remove_plot / ids / errors of the plot registry (plot_protocols.py:15-118) on an
    analyzer that never touches the device."""
    from pybilt_b200 import BilayerAnalyzer
    from tests import cases
    top, traj = cases.make_inputs("tiny_all")
    ba = BilayerAnalyzer(structure=top, trajectory=traj, selection="all")
    ba.add_analysis("msd msd_popc resname POPC")
    ba.add_analysis("apl_grid apl")
    ba.add_analysis("thickness_grid bt")
    ba.add_plot("msd msd_a msd_1 All")
    ba.add_plot("msd msd_l msd_1 All msd_popc POPC")
    ba.add_plot("apl apl_p apl None")
    ba.add_plot("bilayer_thickness bt_p bt None")
    assert ba.get_plot_ids() == ["msd_a", "msd_l", "apl_p", "bt_p"]
    assert ba._plot_protocol.command_protocol["msd_l"].include == ["msd_1", "msd_popc"]
    assert ba._plot_protocol.command_protocol["msd_l"].save_file_name == "msd_l.pdf"
    with pytest.raises(RuntimeError):
        ba.add_plot("msd msd_a msd_1 All")              # id used twice
    with pytest.raises(RuntimeError):
        ba.add_plot("msd other apl Nope")               # not an msd analysis
    with pytest.raises(RuntimeError):
        ba.add_plot("rainbow p1 msd_1 All")             # unknown plot key
    ba.remove_plot("msd_a")
    assert ba.get_plot_ids() == ["msd_l", "apl_p", "bt_p"]
    ba.print_plot_protocol()


def test_diffusion_estimators_match_reference_fixture():
    """pybilt/diffusion/diffusion_coefficients.py:17-167 executed by oracle/make_golden.py
    (tests/golden/estimators.npz): Einstein, linear fit (slope and its error) and the anomalous fit,
    2-D and 3-D, with and without time_range, on the c2_prefix msd rows and on a noisy line."""
    import os
    import warnings
    from pybilt_b200 import diffusion as dc
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "estimators.npz"))
    for key in ("c2", "line"):
        t, y = g[key + "/t"], g[key + "/msd"]
        for tag in ("full", "range"):
            tr = g[f"{key}/{tag}/time_range"]
            tr = None if np.isnan(tr).any() else [float(tr[0]), float(tr[1])]
            for dim in (2, 3):
                assert np.isclose(dc.diffusion_coefficient_Einstein(t, y, dim=dim, time_range=tr),
                                  g[f"{key}/{tag}/einstein/{dim}"], rtol=1e-14, atol=0)
                assert np.allclose(dc.diffusion_coefficient_linear_fit(t, y, dim=dim, time_range=tr),
                                   g[f"{key}/{tag}/linear/{dim}"], rtol=1e-12, atol=0)
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    p = dc.diffusion_coefficient_anomalous_fit(t[1:], y[1:], dim=dim, time_range=tr)
                assert np.allclose(p, g[f"{key}/{tag}/anomalous/{dim}"], rtol=1e-9, atol=0)


def test_host_blas_ddot_fuses_like_the_fixtures_assume():
    """The reference's np.dot(d, d) of two float64 goes through the host BLAS' ddot; the golden
    fixtures (and the kernels' d2 = fma(dy, dy, dx*dx)) encode the FMA accumulation of OpenBLAS' x86
    kernels (DESIGN.md "Arithmetic").  A host whose BLAS does not fuse would make the oracle differ
    from the fixtures by 1 ulp at bin-edge ties: detect that here instead of in a parity test."""
    from fractions import Fraction
    rng = np.random.default_rng(5)
    v = rng.normal(0, 30, (4000, 2))
    fused = plain = 0
    for a, b in v:
        got = float(np.dot(np.array([a, b]), np.array([a, b])))
        want_fma = float(Fraction(b) * Fraction(b) + Fraction(a * a))     # one rounding: fma(b, b, a*a)
        fused += got == want_fma
        plain += got == a * a + b * b
    assert fused == len(v), "np.dot(float64[2]) is not fma(b, b, a*a) on this host (%d/%d; a*a+b*b: %d)" % (
        fused, len(v), plain)


def test_binned_average_and_block_averager_host_logic():
    """Host pieces of the round-2 analyses against the oracle's restatement (no GPU): msd_dist's
    binned_average (first closed bin that holds a position wins, empty bins dropped) and the prefab
    drivers' BlockAverager."""
    from oracle import port
    from pybilt_b200.analysis_protocols import binned_average, _interval_pairs
    from pybilt_b200.running_stats import BlockAverager, RunningStats
    rng = np.random.default_rng(5)
    data, pos = rng.normal(size=400), rng.uniform(-2.0, 27.0, size=400)
    pos[:5] = [0.0, 25.0, 5.0, 10.0, 12.5]                    # bin edges: the lower bin takes them
    got = binned_average(data, pos, n_bins=10, position_range=[0.0, 25.0])
    want = port.binned_average(data, pos, 10, [0.0, 25.0])
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
    # interval bookkeeping: frames 3, 5, 7, ... with interval 4 pair up every other analysed frame
    class B:                                                  # the two attributes _interval_pairs reads
        n_frames = 8
        frame_numbers = np.arange(3, 19, 2)
    assert _interval_pairs(B, 3, 4) == port.interval_pairs(B.frame_numbers, 4) == [(0, 2), (2, 4), (4, 6)]
    assert _interval_pairs(B, 3, 3) == []
    # block averages: 10 values per block, the last short block (3 < min 5 points) is left out
    vals = rng.normal(size=53)
    ba = BlockAverager(points_per_block=10, min_points_in_block=5)
    ba.push_container(vals)
    means = np.array([vals[i:i + 10].mean() for i in range(0, 50, 10)])
    avg, err = ba.get()
    assert ba.number_of_blocks() == 6 and len(ba.averages_of_blocks()) == 5
    assert abs(avg - means.mean()) < 1e-12 and abs(err - means.std() / np.sqrt(5)) < 1e-12
    rs = RunningStats()
    for v in vals[:10]:
        rs.push(v)
    assert abs(ba.averages_of_blocks()[0] - rs.mean()) == 0.0
