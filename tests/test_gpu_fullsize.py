"""Full-size checks on BASELINE.json configs[1] (512 lipids x 12 beads, 5,000 frames): the oracle
cannot run this in seconds, so the CUDA path is held to size-independent properties --
batch-split invariance of the reduction (bit for bit), agreement of three independently written
RDF kernels, exact agreement with the oracle port on a prefix, and algebraic identities of the
outputs over the whole run.  Everything goes through the C ABI (pybilt_b200.engine -> libpbk)."""
import numpy as np
import pytest
import torch

from oracle import port

pytestmark = pytest.mark.gpu
F_FULL = 5000


@pytest.fixture(scope="module")
def c2_full():
    from pybilt_b200 import engine as eng, synthetic as syn
    spec, _ = syn.config_spec("C2")
    top, traj = syn.generate(spec, F_FULL)
    layout = eng.build_bead_layout(top.masses, top.names, top.resindex, top.resnames, top.resids)
    x = torch.from_numpy(traj.positions).cuda()
    box = torch.from_numpy(np.ascontiguousarray(traj.dimensions[:, :3])).cuda()
    e = eng.LipidReductionEngine(top.masses, layout, F_FULL, batch_frames=F_FULL)
    e.push(x, box)
    e.check_status()
    assert e.used_fused
    return dict(top=top, traj=traj, layout=layout, x=x, box=box, e=e)


def test_reduction_is_invariant_under_batch_splits(c2_full):
    """One 5,000-frame batch (148 chunks of 34 frames) against 7 ragged batches (chunks of 5 frames,
    carried unwrap state between batches) and against the generic multi-kernel path."""
    from pybilt_b200 import engine as eng
    d = c2_full
    ref = d["e"]
    for kwargs in (dict(batch_frames=733), dict(batch_frames=1250, allow_fused=False)):
        e = eng.LipidReductionEngine(d["top"].masses, d["layout"], F_FULL, **kwargs)
        e.push(d["x"], d["box"])
        e.check_status()
        assert torch.equal(e.labels, ref.labels)
        assert torch.equal(e.mem_com, ref.mem_com)
        assert torch.equal(e.com, ref.com)
        assert torch.equal(e.com_unwrap, ref.com_unwrap)
        assert torch.equal(e.d_u_state, ref.d_u_state)


def test_reduction_matches_oracle_on_prefix_and_run_identities(c2_full):
    """The port on frames 0..39 (bit-identical; a different chunking than the 60-frame fixture), then
    identities over all 5,000 frames."""
    d = c2_full
    e, top, traj = d["e"], d["top"], d["traj"]
    sub = type(traj)(traj.positions[:40], traj.dimensions[:40], traj.times[:40], traj.dt)
    _out, reps = port.run(top, sub, ["msd msd_1"], {}, return_reps=True)
    assert np.array_equal(e.com[:40].cpu().numpy(), reps["com"])
    assert np.array_equal(e.com_unwrap[:40].cpu().numpy(), reps["com_unwrap"])
    assert np.array_equal(e.labels[:40].cpu().numpy(), reps["labels"])
    # algebra on the whole run: wrapped COMs inside the box hull, leaflets balanced and stable,
    # membrane COM removed (mass-weighted mean of the unwrapped COMs vanishes to rounding)
    com = e.com.cpu().numpy()
    L = traj.dimensions[:, None, :3].astype(np.float64)
    assert (com >= 0).all() and (com <= L).all()
    lab = e.labels.cpu().numpy()
    assert (lab.sum(axis=1) == 256).all() and (lab == lab[0]).all()
    bm = d["layout"].bead_mass
    mean_u = (e.com_unwrap.cpu().numpy() * bm[None, :, None]).sum(axis=1) / bm.sum()
    assert np.abs(mean_u).max() < 1e-4          # float32 rounding of v = f32(u - mem_com), per atom


def test_three_rdf_kernels_agree_at_full_size(c2_full, monkeypatch):
    """Pair list over frame runs (k_rdf5), per-frame cell sweep with float32 look-up (k_rdf3) and the
    block-per-leaflet prefilter + exact queue kernel (k_rdf2): identical counts on 10,000 leaflets;
    the dense sqrt kernel on the first 200 frames; sum rule against a direct count."""
    d = c2_full
    e = d["e"]
    codes = torch.zeros(d["layout"].n_beads, dtype=torch.int32, device="cuda")
    for nb, hi in ((25, 25.0), (80, 40.0)):
        edges = np.histogram([-1], bins=nb, range=[0.0, hi])[1]
        c5, fc5 = e.rdf(codes, 0, 0, 3, edges)
        monkeypatch.setenv("PBK_RDF_NOLIST", "1")
        c3, fc3 = e.rdf(codes, 0, 0, 3, edges)
        monkeypatch.setenv("PBK_RDF_V2", "1")
        c2, fc2 = e.rdf(codes, 0, 0, 3, edges)
        monkeypatch.delenv("PBK_RDF_NOLIST")
        monkeypatch.delenv("PBK_RDF_V2")
        assert torch.equal(c5, c3) and torch.equal(c5, c2)
        assert torch.equal(fc5, fc3) and torch.equal(fc5, fc2)
        assert int(c5.sum()) > 0 and int(c5.sum()) % 2 == 0        # same-type pairs are counted in both orders
        cd, _ = e.rdf(codes, 0, 0, 3, edges, frames=slice(0, 200), dense=True)
        c5s, _ = e.rdf(codes, 0, 0, 3, edges, frames=slice(0, 200))
        assert torch.equal(cd, c5s)
    # frame-range additivity: counts of [0, 5000) = [0, 1234) + [1234, 5000)
    edges = np.histogram([-1], bins=25, range=[0.0, 25.0])[1]
    a, _ = e.rdf(codes, 0, 0, 3, edges, frames=slice(0, 1234))
    b, _ = e.rdf(codes, 0, 0, 3, edges, frames=slice(1234, F_FULL))
    c, _ = e.rdf(codes, 0, 0, 3, edges)
    assert torch.equal(a + b, c)


def test_msd_identities_at_full_size(c2_full):
    d = c2_full
    e = d["e"]
    lab0 = e.labels[0].cpu().numpy()
    idx = np.concatenate([np.nonzero(lab0 == 1)[0], np.nonzero(lab0 == 0)[0]]).astype(np.int32)
    ends = np.arange(F_FULL, dtype=np.int32)
    m = e.msd_pairs(idx, np.zeros(F_FULL, dtype=np.int32), ends).cpu().numpy()
    assert m[0] == 0.0 and (m[1:] > 0).all()
    # symmetric in (origin, end); equal to the NumPy evaluation of the resident COMs to 1e-12
    m_rev = e.msd_pairs(idx, ends[::-1].copy(), np.full(F_FULL, F_FULL - 1, dtype=np.int32)).cpu().numpy()
    want = ((e.com_unwrap[:, idx, :2] - e.com_unwrap[0, idx, :2]) ** 2).sum(dim=2).mean(dim=1).cpu().numpy()
    assert np.allclose(m, want, rtol=1e-12, atol=0)
    assert m_rev[0] == 0.0 and m_rev[-1] == m[-1]
