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


def run_engine(name, batch_frames=None):
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
                                 batch_frames=batch_frames)
    x = torch.from_numpy(np.ascontiguousarray(traj.positions[fr])).cuda()
    box = torch.from_numpy(np.ascontiguousarray(traj.dimensions[fr][:, :3])).cuda()
    e.push(x, box)
    e.check_status()
    return e, top, traj, fr, layout


@pytest.mark.parametrize("name", list(cases.CASES))
@pytest.mark.parametrize("batch", [None, 5])
def test_representations_match_reference(name, batch):
    gold = load(name)
    e, top, traj, fr, layout = run_engine(name, batch)
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
