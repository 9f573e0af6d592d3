"""Run the UNMODIFIED PyBILT reference (``/root/reference``) on in-memory arrays.

TEST INFRASTRUCTURE ONLY.  Works only where ``/root/reference`` exists (the build
container); nothing that runs on the GPU box may import this module.  Recipe from
SURVEY.md Appendix C:

1. restore ``np.int`` / ``np.float`` (used at lipid_grid_opt.py:233, bilayer_analyzer.py:726);
2. permissive stubs for matplotlib / seaborn / MDAnalysis.analysis (import-time only);
3. ``oracle.fake_mda`` registered as ``sys.modules['MDAnalysis']``.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import types
import warnings

import numpy as np

REFERENCE_ROOT = os.environ.get("PYBILT_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "pybilt"))


class _Absorb:
    """Attribute/call absorbing dummy used by the stub modules."""

    def __call__(self, *a, **k):
        return _Absorb()

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Absorb()

    def __iter__(self):
        return iter(())

    def __getitem__(self, item):
        return _Absorb()

    def __setitem__(self, key, value):
        pass


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Absorb()


_installed = False


def install():
    """Idempotently make ``import pybilt`` work in this process."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError("reference tree not found at " + REFERENCE_ROOT)
    if not hasattr(np, "int"):
        np.int = int          # type: ignore[attr-defined]
    if not hasattr(np, "float"):
        np.float = float      # type: ignore[attr-defined]
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.cm", "matplotlib.colors",
                 "seaborn", "MDAnalysis.analysis", "MDAnalysis.analysis.align",
                 "MDAnalysis.lib", "MDAnalysis.lib.mdamath"):
        if name not in sys.modules:
            m = _StubModule(name)
            m.__path__ = []   # type: ignore[attr-defined]
            sys.modules[name] = m
    from . import fake_mda
    mda = types.ModuleType("MDAnalysis")
    mda.__path__ = []         # type: ignore[attr-defined]
    mda.Universe = fake_mda.Universe
    mda.analysis = sys.modules["MDAnalysis.analysis"]
    mda.lib = sys.modules["MDAnalysis.lib"]
    sys.modules["MDAnalysis"] = mda
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _installed = True


@contextlib.contextmanager
def quiet():
    """The reference prints every frame (bilayer_analyzer.py:1029, analysis_protocols.py:4967)."""
    with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        yield


def make_analyzer(top, traj, analyses=(), selection="all", rep_settings=None,
                  frame_range=None, settings=None):
    """A real ``pybilt.bilayer_analyzer.BilayerAnalyzer`` over (top, traj) arrays."""
    install()
    with quiet():
        from pybilt.bilayer_analyzer.bilayer_analyzer import BilayerAnalyzer
        ba = BilayerAnalyzer(structure=top,
                             trajectory=(traj.positions, traj.dimensions, traj.times),
                             selection=selection)
        ba.remove_analysis("msd_1")
        for a in analyses:
            ba.add_analysis(a)
        if rep_settings:
            for rep, d in rep_settings.items():
                for k, v in d.items():
                    ba.rep_settings[rep][k] = v
        if settings:
            ba.settings.update(settings)
        if frame_range is not None:
            ba.set_frame_range(*frame_range)
    return ba


def run_reference(top, traj, analyses=(), selection="all", rep_settings=None,
                  frame_range=None, settings=None, capture=True, capture_grid=False):
    """Run the literal reference loop; return final analysis outputs and, per analysed
    frame, the representations a plugin sees (COMs, unwrapped COMs, membrane COM,
    leaflet labels, unwrapped atom coordinates, optionally the lipid grids)."""
    ba = make_analyzer(top, traj, analyses, selection, rep_settings, frame_range, settings)
    frames = []
    with quiet():
        for _ in ba:
            if not capture:
                continue
            cf = ba.reps["com_frame"]
            rec = {
                "com": np.array([l.com for l in cf.lipidcom]),
                "com_unwrap": np.array([l.com_unwrap for l in cf.lipidcom]),
                "mass": np.array([l.mass for l in cf.lipidcom]),
                "mem_com": np.array(cf.mem_com),
                "box": np.array(cf.box),
                "time": cf.time,
                "number": cf.number,
                "u": np.array(ba._oldcoords),
                "v": np.array(ba.reps["current_mda_frame"]._pos),
                "upper": np.array(ba.reps["leaflets"]["upper"].get_member_indices(), dtype=np.int64),
                "lower": np.array(ba.reps["leaflets"]["lower"].get_member_indices(), dtype=np.int64),
                "upper_groups": list(ba.reps["leaflets"]["upper"].get_group_names()),
                "lower_groups": list(ba.reps["leaflets"]["lower"].get_group_names()),
                "leaflet_attr": list(cf.leaflets()),
                "types": list(cf.resnames()),
            }
            if capture_grid and ba.reps.get("lipid_grid") is not None:
                lg = ba.reps["lipid_grid"]
                for leaf in ("upper", "lower"):
                    g = lg.leaf_grid[leaf]
                    rec["grid_" + leaf] = np.array(g.lipid_grid)
                    rec["gridz_" + leaf] = np.array(g.lipid_grid_z)
                    rec["xc_" + leaf] = np.array(g.x_centers)
                    rec["yc_" + leaf] = np.array(g.y_centers)
                    rec["incr_" + leaf] = np.array([g.x_incr, g.y_incr])
                rec["thickness"] = np.array(lg.average_thickness())
                apl = lg.area_per_lipid()
                rec["apl_avg"] = apl[0]
                rec["apl_res"] = {k: tuple(v) for k, v in apl[1].items()}
                rec["apl_lipid"] = np.array([apl[2][i] for i in sorted(apl[2])])
            frames.append(rec)
        out = {a_id: ba.get_analysis_data(a_id) for a_id in ba.get_analysis_ids()}
    return out, frames, ba


def run_comtraj_reference(top, traj, plane="xy", fstart=0, fend=-1, fskip=1):
    """The literal ``pybilt.com_trajectory.COMTraj.COMTraj`` over (top, traj) arrays."""
    install()
    import tempfile
    from . import fake_mda
    with quiet():
        from pybilt.com_trajectory import COMTraj as ct
        u = fake_mda.Universe(top, (traj.positions, traj.dimensions, traj.times))
        sel = u.select_atoms("all")
        tmp = tempfile.mkdtemp(prefix="comtraj_") + "/"
        obj = ct.COMTraj(u.trajectory, sel, plane=plane, fstart=fstart, fend=fend, fskip=fskip,
                         frame_path=tmp)
        n = obj.nframes
        com = np.array([[l.com for l in obj.frame[f].lipidcom] for f in range(n)])
        comu = np.array([[l.com_unwrap for l in obj.frame[f].lipidcom] for f in range(n)])
        out = {"com": com, "com_unwrap": comu, "com_leaflet": list(obj.com_leaflet),
               "msd_both": obj.calc_msd(), "msd_upper": obj.calc_msd(leaflet="upper"),
               "upper_all": obj.leaflets["upper"].get_group_indices("all"),
               "n_groups": len(obj.leaflets["upper"].groups)}
        first = obj.leaflets["upper"].get_group_names()[0]
        out["msd_lower_first"] = obj.calc_msd(leaflet="lower", group=first)
        out["first_group"] = first
        # analyses beyond the MSD
        zavgs, zmaps = obj.calc_membrane_thickness()
        out["thick_avgs"], out["thick_maps"] = zavgs, zmaps
        out["apl_circle_both"] = obj.calc_area_per_lipid_closest_neighbor_circle()
        out["apl_circle_upper_first"] = obj.calc_area_per_lipid_closest_neighbor_circle(leaflet="upper", group=first)
        out["apl_box_both"] = obj.calc_area_per_lipid_box()
        out["apl_box_lower"] = obj.calc_area_per_lipid_box(leaflet="lower")
        cl = obj.check_clustering(leaflet="upper", dist=9.0)
        out["cluster_upper_9"] = cl
        out["cluster_upper_9_lists"] = [[list(map(int, c)) for c in fr] for fr in obj.clusters]
        out["cluster_both_first_12"] = obj.check_clustering(group=first, dist=12.0)
        sv = obj.step_vector(fstep=2)
        out["step_vec_2"] = np.array(sv)
        out["step_vec_2_wrapped_upper"] = np.array(obj.step_vector(leaflet="upper", fstep=3, wrapped=True))
        np_int = getattr(np, "int", None)
        out["step_vec_frames_2"] = np.array(obj.step_vector_frames(fstep=2))
        vor = obj.voronoi_tesselate(leaflet="upper")
        out["voronoi_upper_points0"] = np.array(vor[0].points)
        out["voronoi_upper_nregions"] = np.array([len(v.regions) for v in vor])
        # last: this one rewrites the stored frames
        obj.remove_leaflet_com_motion()
        out["com_unwrap_no_leaflet_motion"] = np.array([[l.com_unwrap for l in obj.frame[f].lipidcom] for f in range(n)])
        out["msd_both_no_leaflet_motion"] = obj.calc_msd()
    return out
