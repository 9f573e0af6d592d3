"""The drop-in API (pybilt_b200.BilayerAnalyzer + the plugin registry) against the golden
outputs of the literal reference: same analysis strings, same get_analysis_data shapes."""
import os

import numpy as np
import pytest

from tests import cases

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))


def close(a, b, rtol, atol=0.0):
    if np.asarray(b).dtype.kind in "USO":            # names (disp_vec resnames): equal, not close
        assert np.array_equal(np.asarray(a).astype(str), np.asarray(b).astype(str))
        return
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    both_nan = np.isnan(a) & np.isnan(b)
    ok = both_nan | (np.abs(a - b) <= rtol * np.maximum(np.abs(a), np.abs(b)) + atol)
    assert ok.all(), float(np.nanmax(np.abs(a - b)))


def make_analyzer(name):
    from pybilt_b200 import BilayerAnalyzer
    c = cases.CASES[name]
    top, traj = cases.make_inputs(name)
    ba = BilayerAnalyzer(structure=top, trajectory=traj, selection="all")
    ba.remove_analysis("msd_1")
    for a in c["analyses"]:
        ba.add_analysis(a)
    for rep, d in (c.get("rep_settings") or {}).items():
        for k, v in d.items():
            ba.rep_settings[rep][k] = v
    if c.get("frame_range"):
        ba.set_frame_range(*c["frame_range"])
    return ba, c


@pytest.mark.parametrize("name", list(cases.CASES))
def test_api_outputs_match_reference(name):
    gold = load(name)
    ba, c = make_analyzer(name)
    ba.run_analysis()
    out = {a_id: ba.get_analysis_data(a_id) for a_id in ba.get_analysis_ids()}
    flat = cases.flatten_outputs(out)
    keys = [k for k in gold if k.startswith("out/")]
    assert sorted(keys) == sorted(flat)
    for k in keys:
        # msd / msd_multi / nnf / thickness / apl rows: 1e-9 relative (north_star);
        # RDF: counts are exact, so the normalised g(r) agrees to rounding
        close(flat[k], gold[k], 1e-9, atol=cases.output_atol(k, gold[k]))


def test_iterator_views_match_reference():
    name = "tiny_all"
    gold = load(name)
    ba, c = make_analyzer(name)
    n = 0
    for _ in ba:
        frame = ba.reps["com_frame"]
        f = n
        assert len(frame) == gold["rep/com"].shape[1]
        got = np.array([l.com for l in frame.lipidcom])
        assert np.array_equal(got, gold["rep/com"][f])
        got_u = np.array([l.com_unwrap for l in frame.lipidcom])
        assert np.array_equal(got_u, gold["rep/com_unwrap"][f])
        assert np.array_equal(frame.mem_com, gold["rep/mem_com"][f])
        up = ba.reps["leaflets"]["upper"].get_member_indices()
        assert np.array_equal(np.nonzero(gold["rep/labels"][f] == 1)[0], up)
        grid = ba.reps["lipid_grid"]
        for leaf in ("upper", "lower"):
            assert np.array_equal(grid.leaf_grid[leaf].lipid_grid, gold[f"rep/grid_{leaf}"][f])
        th = grid.average_thickness()
        close(np.array(th), gold["rep/thickness"][f], 1e-12)
        avg, per_res, area = grid.area_per_lipid()
        close(np.array([area[i] for i in sorted(area)]), gold["rep/apl_lipid"][f], 1e-12)
        assert frame.time == gold["rep/time"][f]
        n += 1
    assert n == gold["rep/com"].shape[0]
    # results are the same as a batched run
    close(ba.get_analysis_data("msd_1"), gold["out/msd_1"], 1e-9)


def test_plain_per_frame_plugin_still_works():
    """A plugin written against the reference recipe (analysis_protocols.py:9-37)."""
    from pybilt_b200 import analysis_protocols as ap

    class ZSpread(ap.AnalysisProtocol):
        analysis_key = "zspread"

        def __init__(self, args):
            self.analysis_id = "none"
            self.settings = {"leaflet": "upper"}
            self._valid_settings = list(self.settings)
            self._parse_args(args)
            self.save_file_name = self.analysis_id + ".pickle"
            self.analysis_output = []

        def run_analysis(self, ba_settings, ba_reps, ba_mda_data):
            idx = ba_reps["leaflets"][self.settings["leaflet"]].get_member_indices()
            z = [ba_reps["com_frame"].lipidcom[i].com_unwrap[ba_settings["norm"]] for i in idx]
            self.analysis_output.append([ba_reps["com_frame"].time, float(np.std(z))])

    ap.valid_analysis.append("zspread")
    ap.analysis_obj_name_dict["zspread"] = "com_frame"
    ap.command_protocols["zspread"] = ZSpread
    try:
        gold = load("tiny_all")
        ba, c = make_analyzer("tiny_all")
        ba.add_analysis("zspread zs leaflet lower")
        ba.run_analysis()
        got = ba.get_analysis_data("zs")
        assert got.shape == (gold["rep/com"].shape[0], 2)
        lab = gold["rep/labels"]
        want = [np.std(gold["rep/com_unwrap"][f][lab[f] == 0][:, 2]) for f in range(len(lab))]
        close(got[:, 1], np.array(want), 1e-12)
    finally:
        ap.valid_analysis.remove("zspread")
        del ap.analysis_obj_name_dict["zspread"], ap.command_protocols["zspread"]


def test_reset_and_rerun_and_frame_range():
    ba, c = make_analyzer("c2_prefix")
    ba.run_analysis()
    first = ba.get_analysis_data("msd_1").copy()
    ba.run_analysis()                      # exhausted iterator: nothing happens
    assert np.array_equal(ba.get_analysis_data("msd_1"), first)
    ba.reset()
    ba.set_frame_range(10, 40, 3)
    ba.run_analysis()
    again = ba.get_analysis_data("msd_1")
    assert again.shape == (11, 2)
    assert again[0, 1] == 0.0 and again[0, 0] == 100.0


@pytest.mark.parametrize("name", ["big_steps", "name_dict", "c2_prefix"])
def test_comtraj_matches_reference(name):
    """pybilt_b200.com_trajectory.COMTraj against the literal COMTraj (com_trajectory objects
    of the north star): COMs bit-identical, leaflets exact, calc_msd within 1e-9."""
    from pybilt_b200.com_trajectory import COMTraj
    gold = load("comtraj_" + name)
    top, traj = cases.make_inputs(name)
    ct = COMTraj(traj, top)
    assert ct.nframes == gold["com"].shape[0] and ct.nlipids == gold["com"].shape[1]
    e = ct._engine
    assert np.array_equal(e.com.cpu().numpy(), gold["com"])
    assert np.array_equal(e.com_unwrap.cpu().numpy(), gold["com_unwrap"])
    lab = np.array([1 if l == "upper" else 0 for l in ct.com_leaflet], dtype=np.uint8)
    assert np.array_equal(lab, gold["labels0"])
    assert ct.leaflets["upper"].get_group_indices("all") == list(gold["upper_all"])
    close(ct.calc_msd(), gold["msd_both"], 1e-9)
    close(ct.calc_msd(leaflet="upper"), gold["msd_upper"], 1e-9)
    close(ct.calc_msd(leaflet="lower", group=str(gold["first_group"])), gold["msd_lower_first"], 1e-9)
    close(ct.calc_msd_parallel(nprocs=4), gold["msd_both"], 1e-9)
    fr = ct.frame[3]
    assert np.array_equal(np.array([l.com_unwrap for l in fr.lipidcom]), gold["com_unwrap"][3])
    assert fr.lipidcom[0].type == top.resnames[0] and len(ct.frame) == ct.nframes


@pytest.mark.parametrize("name", ["big_steps", "name_dict", "c2_prefix"])
def test_comtraj_analyses_beyond_msd_match_reference(name):
    """thickness from nearest cross-leaflet partners, closest-neighbour-circle and
    box areas, clusters, step vectors, tessellation inputs, leaflet COM motion removal -- against the
    literal COMTraj (oracle/ref_harness.run_comtraj_reference).  Partner choices and cluster lists are
    exact; floating outputs to 1e-12 (they come out bit-identical)."""
    from pybilt_b200.com_trajectory import COMTraj
    gold = load("comtraj_" + name)
    top, traj = cases.make_inputs(name)
    ct = COMTraj(traj, top)
    first = str(gold["first_group"])
    zavgs, zmaps = ct.calc_membrane_thickness()
    assert np.array_equal(zmaps, gold["thick_maps"])          # partner choice + host arithmetic: identical
    close(zavgs, gold["thick_avgs"], 1e-12)
    close(ct.calc_membrane_thickness_parallel(nprocs=3)[0], gold["thick_avgs"], 1e-12)
    close(ct.calc_area_per_lipid_closest_neighbor_circle(), gold["apl_circle_both"], 1e-12)
    close(ct.calc_area_per_lipid_closest_neighbor_circle(leaflet="upper", group=first), gold["apl_circle_upper_first"], 1e-12)
    close(ct.calc_area_per_lipid_box(), gold["apl_box_both"], 1e-12)
    close(ct.calc_area_per_lipid_box(leaflet="lower"), gold["apl_box_lower"], 1e-12)
    close(ct.check_clustering(leaflet="upper", dist=9.0), gold["cluster_upper_9"], 1e-12)
    rows = np.array([(f, c, m) for f, fr in enumerate(ct.clusters) for c, cl in enumerate(fr) for m in cl],
                    dtype=np.int64).reshape(-1, 3)
    assert np.array_equal(rows, gold["cluster_upper_9_lists"])  # same clusters, same member order
    exported = ct.export_clusters_for_plotting()
    assert len(exported) == ct.nframes and len(exported[0][0]) == int((rows[:, 0] == 0).sum())
    close(ct.check_clustering(group=first, dist=12.0), gold["cluster_both_first_12"], 1e-12)
    assert np.array_equal(np.array(ct.step_vector(fstep=2)), gold["step_vec_2"])
    assert np.array_equal(np.array(ct.step_vector(leaflet="upper", fstep=3, wrapped=True)), gold["step_vec_2_wrapped_upper"])
    assert np.array_equal(ct.step_vector_frames(fstep=2), gold["step_vec_frames_2"])
    vor = ct.voronoi_tesselate(leaflet="upper")
    assert np.array_equal(vor[0].points, gold["voronoi_upper_points0"])
    assert np.array_equal(np.array([len(v.regions) for v in vor]), gold["voronoi_upper_nregions"])
    tri = ct.delaunay_tesselate(leaflet="upper")
    assert len(tri) == ct.nframes and np.array_equal(tri[0][0], gold["voronoi_upper_points0"][:, 0])
    ct.remove_leaflet_com_motion()
    assert np.array_equal(ct._engine.com_unwrap.cpu().numpy(), gold["com_unwrap_no_leaflet_motion"])
    close(ct.calc_msd(), gold["msd_both_no_leaflet_motion"], 1e-9)


@pytest.mark.parametrize("name", list(cases.CASES))
@pytest.mark.parametrize("window", [3, 5, 1000])
def test_windowed_run_matches_resident_run(name, window):
    """Streamed run (bounded engine, windows of `window` frames with context rows) against the
    resident run and the literal reference: same outputs, whatever the window size."""
    gold = load(name)
    ba, c = make_analyzer(name)
    ba.run_analysis()
    want = cases.flatten_outputs({a_id: ba.get_analysis_data(a_id) for a_id in ba.get_analysis_ids()})
    bw, c = make_analyzer(name)
    bw.window_frames = window
    if name == "orientation":
        with pytest.raises(RuntimeError, match="resident"):     # orientation leaflets: resident runs only
            bw.run_analysis()
        return
    if name == "n3_rest":
        # disp_vec / msd_dist / dc_cluster look at pairs of frames far apart and at per-frame host
        # views: they say so instead of running on a window
        with pytest.raises(RuntimeError, match="resident"):
            bw.run_analysis()
        return
    bw.run_analysis()
    got = cases.flatten_outputs({a_id: bw.get_analysis_data(a_id) for a_id in bw.get_analysis_ids()})
    assert sorted(got) == sorted(want)
    for k in want:
        assert np.array_equal(np.asarray(got[k]), np.asarray(want[k]), equal_nan=True), k
        if k in gold:
            close(got[k], gold[k], 1e-9, atol=cases.output_atol(k, gold[k]))


def test_window_is_chosen_automatically_when_resident_state_exceeds_budget():
    ba, c = make_analyzer("c2_prefix")
    assert ba._window_choice() is None
    ba.resident_budget_bytes = 1 << 10
    assert ba._window_choice() >= 256
    ba.run_analysis()
    gold = load("c2_prefix")
    flat = cases.flatten_outputs({a_id: ba.get_analysis_data(a_id) for a_id in ba.get_analysis_ids()})
    for k in [k for k in gold if k.startswith("out/")]:
        close(flat[k], gold[k], 1e-9, atol=cases.output_atol(k, gold[k]))


@pytest.mark.parametrize("name,window", [("tiny_all", None), ("c3_small", None), ("name_dict", None), ("tiny_all", 5)])
def test_run_from_psf_dcd_files_matches_in_memory_run(name, window, tmp_path):
    """Native psf + dcd ingestion (pybilt_b200/formats.py + the device-side plane interleave and
    atom selection) against the same system handed over as arrays: identical representations
    and analysis outputs (the time column only to float32(delta) accuracy, as DCD stores it)."""
    from pybilt_b200 import BilayerAnalyzer
    from tests.test_formats_cpu import export_system
    psf, dcds, top, traj, lipid, pos = export_system(tmp_path, name=name, split=3)
    c = cases.CASES[name]
    bas = []
    for structure, trajectory, sel in ((top, traj, "all"),
                                       (psf, dcds, "not resname CLA and not resname TIP3 and not resname POT")):
        ba = BilayerAnalyzer(structure=structure, trajectory=trajectory, selection=sel)
        ba.remove_analysis("msd_1")
        for a in c["analyses"]:
            ba.add_analysis(a)
        for rep, d in (c.get("rep_settings") or {}).items():
            for k, v in d.items():
                ba.rep_settings[rep][k] = v
        if c.get("frame_range"):
            ba.set_frame_range(*c["frame_range"])
        ba.window_frames = window
        ba.run_analysis()
        bas.append(ba)
    mem, fil = bas
    if window is None:
        for attr in ("com", "com_unwrap", "mem_com", "labels"):
            assert np.array_equal(getattr(mem._engine, attr).cpu().numpy(), getattr(fil._engine, attr).cpu().numpy())
    want = cases.flatten_outputs({k: mem.get_analysis_data(k) for k in mem.get_analysis_ids()})
    got = cases.flatten_outputs({k: fil.get_analysis_data(k) for k in fil.get_analysis_ids()})
    assert sorted(want) == sorted(got)
    for k in want:
        a, b = np.asarray(got[k], dtype=float), np.asarray(want[k], dtype=float)
        if a.ndim == 2 and a.shape[1] in (2, 4) and not np.array_equal(a, b):
            assert np.array_equal(a[:, 1:], b[:, 1:], equal_nan=True), k      # data columns exact
            close(a[:, 0], b[:, 0], 1e-6)                                       # time column
        else:
            close(a, b, 1e-6, atol=1e-300)        # msd_multi's D divides by tau = t[n_tau-1] - t[0]


def test_prefab_style_driver_sequence(tmp_path, monkeypatch, capsys):
    """The calls the reference's prefab drivers make on an analyzer (msd_diffusion,
    prefab_analysis_protocols.py:80-140; area_per_lipid :363-375): analyses per lipid type and
    leaflet, plots over them, protocol print-outs, run, dump to a path prefix, save the plots,
    diffusion estimators on the msd rows."""
    from pybilt_b200 import BilayerAnalyzer, diffusion as dc
    monkeypatch.chdir(tmp_path)
    top, traj = cases.make_inputs("c3_small")
    ba = BilayerAnalyzer(structure=top, trajectory=traj, selection="all")
    ba.set_frame_range(0, -1, 1)
    resnames = ["POPC", "DOPE"]
    for lt in resnames:
        ba.add_analysis("msd msd_" + lt + " resname " + lt)
        ba.add_analysis("msd msd_" + lt + "_upper resname " + lt + " leaflet upper")
        ba.add_analysis("msd msd_" + lt + "_lower resname " + lt + " leaflet lower")
    ba.rep_settings['lipid_grid']['n_xbins'] = ba.rep_settings['lipid_grid']['n_ybins'] = 16
    ba.add_analysis("apl_grid apl_grid")
    ba.add_analysis("thickness_grid bt")
    ba.add_plot('msd msd_a msd_1 All')
    ba.add_plot("msd msd_l " + " ".join("msd_%s %s" % (lt, lt) for lt in resnames))
    for lt in resnames:
        ba.add_plot("msd %s_cul msd_%s Composite msd_%s_upper Upper msd_%s_lower Lower" % (lt, lt, lt, lt))
    ba.add_plot("apl apl_p apl_grid None")
    ba.add_plot("bilayer_thickness bt_p bt None")
    ba.print_analysis_protocol()
    ba.print_plot_protocol()
    ba.run_analysis()
    ba.dump_data(path=str(tmp_path) + "/")
    ba.save_all_plots()
    assert (tmp_path / "msd_1.pickle").exists() and (tmp_path / "apl_grid.pickle").exists()
    msd = ba.get_analysis_data('msd_1')
    ser = ba.generate_plot("msd_l")
    assert [s[0] for s in ser] == resnames
    assert np.array_equal(ser[0][2], ba.get_analysis_data("msd_POPC")[:, 1])
    assert np.allclose(ser[0][1] * 1000.0, msd[:, 0])
    apl = ba.generate_plot("apl_p")
    assert [s[0] for s in apl][0] == "composite" and len(apl) == 1 + 3
    assert np.array_equal(apl[0][2], ba.get_analysis_data("apl_grid")["composite"][:, 2])
    bt = ba.generate_plot("bt_p")
    assert np.array_equal(bt[0][2], ba.get_analysis_data("bt")[:, 2])
    times, vals = msd[:, 0], msd[:, 1]
    d_e = dc.diffusion_coefficient_Einstein(times, vals)
    d_l = dc.diffusion_coefficient_linear_fit(times, vals)
    assert np.isfinite(d_e) and np.isfinite(d_l[0])
    out = capsys.readouterr().out
    assert "with plots:" in out and "msd_l" in out


def test_leaflet_distance_with_update_interval_raises_like_the_reference():
    """Between leaflet rebuilds the reference's COM frame carries no leaflet names; its mean over an
    empty selection is a nan scalar and indexing it raises IndexError (analysis_protocols.py:1158)."""
    from pybilt_b200 import BilayerAnalyzer
    top, traj = cases.make_inputs("leaflet_dist")
    ba = BilayerAnalyzer(structure=top, trajectory=traj, selection="all")
    ba.add_analysis("thickness_leaflet_distance tld")
    ba.rep_settings["leaflets"]["update_interval"] = 3
    with pytest.raises(IndexError):
        ba.run_analysis()


def test_unwrap_coordinates_up_to_continues_the_unwrap():
    """bilayer_analyzer.py:1034-1079: frames 0..6 are only unwrapped, the run over frames 7.. starts from
    that state (golden from the executed reference, adversarial unwrap case); reset() forgets it."""
    from pybilt_b200 import BilayerAnalyzer
    gold = load("unwrap_up_to")
    top, traj = cases.make_inputs("big_steps")
    ba = BilayerAnalyzer(structure=top, trajectory=traj, selection="all")
    ba.add_analysis("msd msd_up leaflet upper")
    ba.unwrap_coordinates_up_to(6)
    ba.set_frame_range(7, -1, 1)
    ba.run_analysis()
    close(ba.get_analysis_data("msd_1"), gold["out/msd_1"], 1e-9)
    close(ba.get_analysis_data("msd_up"), gold["out/msd_up"], 1e-9)
    assert np.array_equal(ba._engine.d_u_state.cpu().numpy(), gold["rep/u_last"])
    # without the pre-pass the first analysed frame is its own reference: different rows
    ba.reset()
    ba.run_analysis()
    assert not np.array_equal(ba._engine.d_u_state.cpu().numpy(), gold["rep/u_last"])


_REPORT = ("Composite for", "Values from", "  Basic", "    Diffusion coefficient", "  Linear fit", "  Anomalous",
           "Final running", "  via the", "Will do block", "Block Averaged", "Nearest neighbor fraction")


def _report_lines(lines):
    import re
    per_type = re.compile(r"^    \w+: -?\d+\.\d{4} \+- ")          # '    POPC: 67.6784 +- 2.1069'
    return [l for l in lines if l.startswith(_REPORT) or per_type.match(l)]


@pytest.mark.parametrize("driver", ["msd_diffusion", "area_per_lipid", "nearest_neighbor_fraction", "com_lateral_rdf"])
def test_prefab_drivers_match_the_reference_drivers(driver, tmp_path, capsys):
    """the prefab drivers on the device analyzer against the reference's own drivers
    executed unmodified (oracle/make_golden.py make_prefab): the summary they print (numbers to 1e-6, they
    are formatted or come out of scipy fits) and the pickles dump_data writes (1e-9)."""
    import pickle
    import re
    from pybilt_b200 import prefab_analysis_protocols as pap
    from oracle.make_golden import PREFAB_RUNS
    gold = load("prefab")
    case, kw = PREFAB_RUNS[driver]
    top, traj = cases.make_inputs(case)
    dump = str(tmp_path) + "/"
    cwd = os.getcwd()
    os.chdir(dump)
    try:
        getattr(pap, driver)(top, traj, "all", dump_path=dump, **kw)
    finally:
        os.chdir(cwd)
    got_lines = _report_lines(capsys.readouterr().out.splitlines())
    want_lines = _report_lines([str(l) for l in gold[driver + "/stdout"]])
    assert len(got_lines) == len(want_lines) and (len(want_lines) > 0 or driver == "com_lateral_rdf")
    num = re.compile(r"[-+]?(?:\d+\.\d*|\.\d+|\d+)(?:[eE][-+]?\d+)?|nan|inf")
    for g, w in zip(got_lines, want_lines):
        assert num.sub("#", g) == num.sub("#", w), (g, w)
        gv, wv = [float(x) for x in num.findall(g)], [float(x) for x in num.findall(w)]
        close(gv, wv, 1e-6, atol=1e-4)
    dumped = {}
    for name in os.listdir(dump):
        if name.endswith(".pickle"):
            with open(os.path.join(dump, name), "rb") as fh:
                dumped[name[:-len(".pickle")]] = pickle.load(fh)
    flat = cases.flatten_outputs(dumped)
    keys = [k[len(driver) + 1:] for k in gold if k.startswith(driver + "/out/")]
    assert keys and sorted(keys) == sorted(flat)
    for k in keys:
        close(flat[k], gold[driver + "/" + k], 1e-9, atol=cases.output_atol(k, gold[driver + "/" + k]))
