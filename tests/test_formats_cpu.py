"""PSF / DCD readers and the selection subset (pybilt_b200/formats.py): round trips through files
written here, the MDAnalysis conventions written down in the module docstring."""
import numpy as np
import pytest

from pybilt_b200 import formats as fmt
from pybilt_b200 import synthetic as syn
from tests import cases


def export_system(tmp_path, name="tiny_all", n_water=7, endian="<", split=None):
    """A synthetic bilayer + some water / ions as psf + dcd files.  Returns (psf, [dcd...], top, traj,
    lipid atom mask)."""
    top, traj = cases.make_inputs(name)
    M = top.n_atoms
    F = traj.positions.shape[0]
    seg = np.array(["MEMB"] * M + ["SOLV"] * (3 * n_water) + ["IONS"] * 2, dtype=object)
    resids = np.concatenate([np.asarray(top.resids)[top.resindex], np.repeat(np.arange(1, n_water + 1), 3), [1, 2]])
    resnames = np.array([top.resnames[r] for r in top.resindex] + ["TIP3"] * (3 * n_water) + ["POT", "CLA"],
                        dtype=object)
    names = np.concatenate([np.asarray(top.names, dtype=object), np.array(["OH2", "H1", "H2"] * n_water, dtype=object),
                            np.array(["POT", "CLA"], dtype=object)])
    masses = np.concatenate([top.masses, np.tile([15.9994, 1.008, 1.008], n_water), [39.0983, 35.45]])
    rng = np.random.default_rng(5)
    extra = rng.uniform(0, 50, size=(F, 3 * n_water + 2, 3)).astype(np.float32)
    pos = np.concatenate([traj.positions, extra], axis=1)
    psf = str(tmp_path / "sys.psf")
    fmt.write_psf(psf, seg, resids, resnames, names, masses)
    dt = float(traj.times[1] - traj.times[0]) if F > 1 else 1.0
    dcds = []
    cuts = [0, F] if not split else [0, split, F]
    for k in range(len(cuts) - 1):
        path = str(tmp_path / ("traj%d.dcd" % k))
        fmt.write_dcd(path, pos[cuts[k]:cuts[k + 1]], traj.dimensions[cuts[k]:cuts[k + 1]], dt_ps=dt, endian=endian)
        dcds.append(path)
    lipid = np.arange(pos.shape[1]) < M
    return psf, dcds, top, traj, lipid, pos


@pytest.mark.parametrize("endian", ["<", ">"])
def test_dcd_round_trip(tmp_path, endian):
    psf, dcds, top, traj, lipid, pos = export_system(tmp_path, endian=endian)
    d = fmt.DCDFile(dcds[0])
    F = pos.shape[0]
    assert (d.n_frames, d.n_atoms, d.has_unitcell) == (F, pos.shape[1], True)
    fr = np.arange(F)
    assert np.array_equal(d.positions(fr), pos)                         # float32 values bit for bit
    assert np.array_equal(d.positions(fr[::3], np.nonzero(lipid)[0]), pos[::3][:, lipid])
    pl = d.planes(2, 5)
    assert pl.shape == (3, 3, pos.shape[1]) and np.array_equal(pl.transpose(0, 2, 1), pos[2:5])
    dims = d.dimensions(fr)
    assert dims.dtype == np.float32
    assert np.array_equal(dims[:, :3], traj.dimensions[:, :3])          # A, B, C as float32(double(float32))
    assert np.allclose(dims[:, 3:], 90.0)                               # cosines 0 -> 90 degrees
    dt = traj.times[1] - traj.times[0]
    assert np.allclose(d.times(fr), fr * dt, rtol=1e-6)                 # delta is stored as float32


def test_dcd_chain_reads_several_files_as_one(tmp_path):
    psf, dcds, top, traj, lipid, pos = export_system(tmp_path, split=5)
    ch = fmt.DCDChain(dcds)
    F = pos.shape[0]
    assert ch.n_frames == F
    assert np.array_equal(ch.positions(np.arange(F)), pos)
    assert np.array_equal(ch.planes(3, 9).transpose(0, 2, 1), pos[3:9])        # spans the file boundary
    assert np.array_equal(ch.all_dimensions()[:, :3], traj.dimensions[:, :3])


def test_psf_and_selections(tmp_path):
    psf, dcds, top, traj, lipid, pos = export_system(tmp_path)
    t = fmt.PSFTopology(psf)
    assert t.n_atoms == pos.shape[1]
    assert np.array_equal(t.masses[lipid], top.masses)
    assert list(t.names[lipid]) == list(top.names)
    # residues = runs of (resid, resname, segid); lipids first, then 7 waters, 2 ions
    assert t.resindex[lipid].max() + 1 == top.n_residues and t.resindex[-1] == top.n_residues + 7 + 1
    sel = fmt.select_atoms
    assert np.array_equal(sel(t, "not resname CLA and not resname TIP3 and not resname POT"), lipid)
    assert np.array_equal(sel(t, "segid MEMB"), lipid)
    assert np.array_equal(sel(t, "all"), np.ones(t.n_atoms, dtype=bool))
    names = sorted(set(top.resnames))
    assert np.array_equal(sel(t, "resname " + " ".join(names)), lipid)
    assert np.array_equal(sel(t, "resname TIP* or (segid IONS and name POT)"),
                          (t.resnames == "TIP3") | (t.names == "POT"))
    assert sel(t, "resid 1:3 and segid SOLV").sum() == 9
    assert sel(t, "bynum 1-4").sum() == 4 and sel(t, "index 0").sum() == 1
    with pytest.raises(RuntimeError):
        sel(t, "around 5 resname POPC")
    with pytest.raises(RuntimeError):
        sel(t, "resname")


def test_trajectory_data_from_files_matches_arrays(tmp_path):
    from pybilt_b200.bilayer_analyzer import TrajectoryData
    psf, dcds, top, traj, lipid, pos = export_system(tmp_path, split=4)
    md = TrajectoryData(psf, dcds, "not resname CLA and not resname TIP3 and not resname POT")
    ref = TrajectoryData(top, traj, "all")
    assert md.natoms == ref.natoms and md.resnames == ref.resnames
    assert np.array_equal(md.masses, ref.masses) and np.array_equal(md.resindex, ref.resindex)
    assert np.array_equal(md.resids, ref.resids) and list(md.names) == list(ref.names)
    fr = np.array([0, 1, 2, 5, 6, 9])
    x, d, t = md.frames(fr)
    xr, dr, tr = ref.frames(fr)
    assert np.array_equal(x, xr) and np.array_equal(d[:, :3], dr[:, :3]) and np.allclose(t, tr, rtol=1e-6)
    fb, ff, offs, foreign = md.dcd.raw_layout
    out = np.zeros((8, fb), dtype=np.uint8)
    raw = md.read_raw_frames(fr, out).view(np.float32)                 # frames as they lie in the files
    assert raw.shape == (len(fr), ff) and not foreign
    got = np.stack([raw[:, o:o + pos.shape[1]] for o in offs], axis=2)
    assert np.array_equal(got, pos[fr])
    with pytest.raises(RuntimeError):
        TrajectoryData(psf, dcds, "name OH2")          # part of a residue
