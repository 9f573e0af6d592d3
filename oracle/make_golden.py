"""Generate tests/golden/*.npz by executing the UNMODIFIED reference (/root/reference).

TEST INFRASTRUCTURE.  Run in the build container only (the reference tree does not exist
on the GPU box):

    python -m oracle.make_golden            # all cases of tests/cases.py
    python -m oracle.make_golden tiny_all   # one case

Each fixture stores the sha256 of the synthetic inputs it was made from, the
representations a plugin sees every analysed frame (wrapped / unwrapped lipid COMs,
membrane COM, leaflet labels, and for grid cases the lipid grids of both
lipid_grid_opt.py -- what BilayerAnalyzer builds -- and the exact lipid_grid.py), and
``get_analysis_data`` of every analysis.
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_harness as rh          # noqa: E402
from tests import cases                       # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def make(name: str) -> str:
    c = cases.CASES[name]
    top, traj = cases.make_inputs(name)
    wants_grid = any(a.split()[0] in ("apl_grid", "thickness_grid") for a in c["analyses"])
    out, frames, ba = rh.run_reference(top, traj, analyses=c["analyses"],
                                       rep_settings=c.get("rep_settings"),
                                       frame_range=c.get("frame_range"),
                                       settings=c.get("settings"), capture_grid=wants_grid)
    flat = cases.flatten_outputs(out)
    N = len(frames[0]["com"])
    labels = np.zeros((len(frames), N), dtype=np.uint8)
    for f, fr in enumerate(frames):
        labels[f][fr["upper"]] = 1
        assert len(fr["upper"]) + len(fr["lower"]) == N
    flat["rep/com"] = np.array([fr["com"] for fr in frames])
    flat["rep/com_unwrap"] = np.array([fr["com_unwrap"] for fr in frames])
    flat["rep/mem_com"] = np.array([fr["mem_com"] for fr in frames])
    flat["rep/mass"] = frames[0]["mass"]
    flat["rep/labels"] = labels
    flat["rep/time"] = np.array([fr["time"] for fr in frames])
    flat["rep/u_last"] = frames[-1]["u"]
    flat["rep/types"] = np.array(frames[0]["types"])
    if wants_grid:
        import pybilt.lipid_grid.lipid_grid as lg_exact
        for leaf in ("upper", "lower"):
            flat[f"rep/grid_{leaf}"] = np.array([fr["grid_" + leaf] for fr in frames])
            flat[f"rep/gridz_{leaf}"] = np.array([fr["gridz_" + leaf] for fr in frames])
        flat["rep/thickness"] = np.array([fr["thickness"] for fr in frames])
        flat["rep/apl_lipid"] = np.array([fr["apl_lipid"] for fr in frames])
        # the exact (brute force) grids of lipid_grid.py on the same frames
        ba2 = rh.make_analyzer(top, traj, [], rep_settings=c.get("rep_settings"),
                               frame_range=c.get("frame_range"))
        nx = c["rep_settings"]["lipid_grid"]["n_xbins"]
        ny = c["rep_settings"]["lipid_grid"]["n_ybins"]
        ex = {"upper": [], "lower": []}
        with rh.quiet():
            for _ in ba2:
                g = lg_exact.LipidGrids(ba2.reps["com_frame"], ba2.reps["leaflets"], [0, 1],
                                        nxbins=nx, nybins=ny)
                for leaf in ex:
                    ex[leaf].append(np.array(g.leaf_grid[leaf].lipid_grid))
        for leaf in ex:
            flat[f"rep/grid_exact_{leaf}"] = np.array(ex[leaf])
    flat["meta/input_sha256"] = np.array(cases.input_digest(traj))
    flat["meta/numpy"] = np.array(np.__version__)
    os.makedirs(GOLDEN, exist_ok=True)
    path = os.path.join(GOLDEN, name + ".npz")
    np.savez_compressed(path, **flat)
    return path


COMTRAJ_CASES = ["big_steps", "name_dict", "c2_prefix"]


def make_comtraj(name: str) -> str:
    """pybilt.com_trajectory.COMTraj.COMTraj (literal) on the inputs of case `name`."""
    top, traj = cases.make_inputs(name)
    ref = rh.run_comtraj_reference(top, traj)
    flat = {"com": ref["com"], "com_unwrap": ref["com_unwrap"],
            "labels0": np.array([1 if l == "upper" else 0 for l in ref["com_leaflet"]], dtype=np.uint8),
            "msd_both": ref["msd_both"], "msd_upper": ref["msd_upper"],
            "msd_lower_first": ref["msd_lower_first"], "first_group": np.array(ref["first_group"]),
            "upper_all": np.array(ref["upper_all"]),
            "meta/input_sha256": np.array(cases.input_digest(traj))}
    for key in ("thick_avgs", "thick_maps", "apl_circle_both", "apl_circle_upper_first", "apl_box_both",
                "apl_box_lower", "cluster_upper_9", "cluster_both_first_12", "step_vec_2",
                "step_vec_2_wrapped_upper", "step_vec_frames_2", "voronoi_upper_points0", "voronoi_upper_nregions",
                "com_unwrap_no_leaflet_motion", "msd_both_no_leaflet_motion"):
        flat[key] = np.asarray(ref[key])
    # cluster lists: frame, cluster number and member, one row per clustered lipid
    flat["cluster_upper_9_lists"] = np.array([(f, c, m) for f, fr in enumerate(ref["cluster_upper_9_lists"])
                                              for c, cl in enumerate(fr) for m in cl], dtype=np.int64).reshape(-1, 3)
    path = os.path.join(GOLDEN, "comtraj_" + name + ".npz")
    np.savez_compressed(path, **flat)
    return path


def make_unwrap_up_to() -> str:
    """BilayerAnalyzer.unwrap_coordinates_up_to (bilayer_analyzer.py:1034-1079) on the adversarial
    unwrap case: frames 0..6 are only unwrapped, the analysis runs over frames 7.. and continues from
    that state."""
    top, traj = cases.make_inputs("big_steps")
    ba = rh.make_analyzer(top, traj, ["msd msd_1", "msd msd_up leaflet upper"], "all", None, None, None)
    with rh.quiet():
        ba.unwrap_coordinates_up_to(6)
        ba.set_frame_range(7, -1, 1)
        ba.run_analysis()
    flat = {"out/msd_1": ba.get_analysis_data("msd_1"), "out/msd_up": ba.get_analysis_data("msd_up"),
            "rep/u_last": np.array(ba._oldcoords, dtype=np.float32),
            "meta/input_sha256": np.array(cases.input_digest(traj))}
    path = os.path.join(GOLDEN, "unwrap_up_to.npz")
    np.savez_compressed(path, **flat)
    return path


PREFAB_RUNS = {
    # driver -> (case whose inputs are used, keyword arguments)
    "msd_diffusion": ("name_dict", dict(resnames=["POPC", "DOPE"])),
    "area_per_lipid": ("tiny_all", dict(n_xbins=8, n_ybins=8)),
    "nearest_neighbor_fraction": ("name_dict", dict(lipid_resnames=["POPC", "DOPE"], frame_interval=2)),
    "com_lateral_rdf": ("c3_small", dict(resnames=["POPC", "CHOL"])),
}


def make_prefab() -> str:
    """The reference's own prefab drivers (pybilt/bilayer_analyzer/prefab_analysis_protocols.py:34,313,834,
    1463) executed unmodified on in-memory inputs: what they print and what dump_data pickles."""
    import contextlib
    import glob
    import io
    import pickle
    import tempfile
    rh.install()
    with rh.quiet():
        import pybilt.bilayer_analyzer.prefab_analysis_protocols as pap
    flat = {}
    for driver, (case, kw) in PREFAB_RUNS.items():
        top, traj = cases.make_inputs(case)
        tmp = tempfile.mkdtemp(prefix="prefab_") + "/"
        buf = io.StringIO()
        cwd = os.getcwd()
        os.chdir(tmp)                      # figures / plot files land in the scratch directory
        try:
            with contextlib.redirect_stdout(buf):
                getattr(pap, driver)(top, (traj.positions, traj.dimensions, traj.times), "all", dump_path=tmp, **kw)
        finally:
            os.chdir(cwd)
        flat[f"{driver}/stdout"] = np.array(buf.getvalue().splitlines())
        dumped = {}
        for path in sorted(glob.glob(tmp + "*.pickle")):
            with open(path, "rb") as fh:
                dumped[os.path.basename(path)[:-len(".pickle")]] = pickle.load(fh)
        for k, v in cases.flatten_outputs(dumped).items():
            flat[f"{driver}/{k}"] = v
    path = os.path.join(GOLDEN, "prefab.npz")
    np.savez_compressed(path, **flat)
    return path


def make_estimators() -> str:
    """The reference's diffusion estimators (pybilt/diffusion/diffusion_coefficients.py:17-167)
    on the msd rows of fixture c2_prefix and on a noisy straight line, with and without time_range."""
    with rh.quiet():
        import pybilt.diffusion.diffusion_coefficients as dc
    g = np.load(os.path.join(GOLDEN, "c2_prefix.npz"))
    rows = g["out/msd_1"]
    rng = np.random.default_rng(0)
    t_lin = np.linspace(0.0, 100.0, 400)
    sets = {"c2": (rows[:, 0].copy(), rows[:, 1].copy()),
            "line": (t_lin, 4.0 * 2.0 * t_lin + rng.normal(0, 0.05, t_lin.size))}
    flat = {}
    for key, (t, y) in sets.items():
        flat[key + "/t"], flat[key + "/msd"] = t.copy(), y.copy()
        for tag, tr in (("full", None), ("range", [float(t[len(t) // 4]), float(t[3 * len(t) // 4])])):
            flat[f"{key}/{tag}/time_range"] = np.array(tr if tr else [np.nan, np.nan])
            for dim in (2, 3):
                flat[f"{key}/{tag}/einstein/{dim}"] = np.array(dc.diffusion_coefficient_Einstein(t.copy(), y, dim=dim, time_range=tr))
                flat[f"{key}/{tag}/linear/{dim}"] = np.array(dc.diffusion_coefficient_linear_fit(t.copy(), y, dim=dim, time_range=tr))
                flat[f"{key}/{tag}/anomalous/{dim}"] = np.array(dc.diffusion_coefficient_anomalous_fit(t[1:].copy(), y[1:], dim=dim, time_range=tr))
    path = os.path.join(GOLDEN, "estimators.npz")
    np.savez_compressed(path, **flat)
    return path


if __name__ == "__main__":
    names = sys.argv[1:] or (list(cases.CASES) + ["comtraj:" + n for n in COMTRAJ_CASES] + ["estimators", "unwrap_up_to", "prefab"])
    for n in names:
        if n == "estimators":
            p = make_estimators()
        elif n == "unwrap_up_to":
            p = make_unwrap_up_to()
        elif n == "prefab":
            p = make_prefab()
        else:
            p = make_comtraj(n.split(":", 1)[1]) if n.startswith("comtraj:") else make(n)
        print(n, "->", p, os.path.getsize(p), "bytes")
