"""oracle/port.py -- NumPy/C restatement of PyBILT's lipid-reduction hot path.

TEST INFRASTRUCTURE, NOT PRODUCT CODE (see oracle/__init__.py).  This is the CPU
checker the CUDA path is compared with, and the "port" CPU baseline ``bench.py``
reports.  It restates, array-at-a-time, what the reference does object-at-a-time;
every function cites the reference file:line (relative to /root/reference) it follows.
It is pinned against the literal reference executed here (tests/test_oracle_golden.py,
fixtures from oracle/make_golden.py).

Conventions: F analysed frames, M selected atoms, N lipids (or beads).  Leaflet label
1 = 'upper', 0 = 'lower'.
"""
from __future__ import annotations

import ctypes
import warnings
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import build as _build

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(_build.build())
        _lib.pc_pairwise_sum.restype = ctypes.c_double
        _lib.pc_leaflet_labels.restype = ctypes.c_int64
        _lib.pc_rdf_accumulate_cells.restype = ctypes.c_int64
    return _lib


def _p(a):
    return ctypes.c_void_p(a.ctypes.data)


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


# --------------------------------------------------------------------------------------
# small scalar helpers
# --------------------------------------------------------------------------------------
class Welford:
    """pybilt/common/running_stats.py:17-85 (RunningStats), NumPy-scalar arithmetic so
    that float32 inputs promote exactly as they do there (first push keeps the dtype)."""

    def __init__(self):
        self.n = 0
        self._m = self._s = np.float64(0.0)

    def push(self, val):
        self.n += 1
        if self.n == 1:
            self._m = np.array([val])[0]
            self._s = np.float64(0.0)
        else:
            n = np.float64(self.n)
            m_new = self._m + (val - self._m) / n
            self._s = self._s + (val - self._m) * (val - m_new)
            self._m = m_new

    def mean(self):
        return self._m if self.n >= 1 else 0.0

    def variance(self):
        if self.n > 1:
            return self._s / (np.float64(self.n) - np.float64(1.0))
        return 0.0

    def deviation(self):
        if self.n > 1:
            return np.sqrt(self._s / (np.float64(self.n) - np.float64(1.0)))
        return np.sqrt(0.0)


def pairwise_sum(a: np.ndarray) -> float:
    """NumPy's float64 pairwise summation order, restated; the order the
    membrane-COM sums at com_frame.py:175-178 are evaluated in."""
    a = np.ascontiguousarray(a, dtype=np.float64)
    return float(lib().pc_pairwise_sum(_p(a), ctypes.c_int64(a.shape[0])))


# --------------------------------------------------------------------------------------
# A1: PBC unwrap            pybilt/mda_tools/mda_unwrap_vectorized.py:12-93
# --------------------------------------------------------------------------------------
def unwrap_frame(x: np.ndarray, u_prev: np.ndarray, abc: np.ndarray) -> np.ndarray:
    """One frame: per axis d=f32(x-u_prev); integer shift from the strict >/< loops in
    float64; u=f32(f64(x)+s*L).  Vectorised over atoms (the loops run while any atom
    still violates, which per atom equals the scalar rule of SURVEY.md B.1)."""
    u = x.copy()
    for k in range(3):
        L = abc[k]                          # np.float32
        h = L / 2.0                         # float32 (exact halving)
        d = x[:, k] - u_prev[:, k]          # float32
        s = np.zeros(x.shape[0])            # float64
        gt = d > h
        while gt.any():
            s[gt] -= 1.0
            gt = (d + s * L) > h            # float64 comparison
        lt = d < -h
        while lt.any():
            s[lt] += 1.0
            lt = (d + s * L) < -h
        u[:, k] = (x[:, k].astype(np.float64) + s * L).astype(np.float32)
    return u


def unwrap(X: np.ndarray, box: np.ndarray, u_state: Optional[np.ndarray] = None) -> np.ndarray:
    """All analysed frames; frame 0 is its own reference unless ``u_state`` (the unwrapped
    coordinates of the frame before) is given (bilayer_analyzer.py:936-954)."""
    U = np.empty_like(X)
    prev = u_state
    for f in range(X.shape[0]):
        U[f] = X[f] if prev is None else unwrap_frame(X[f], prev, box[f])
        prev = U[f]
    return U


# --------------------------------------------------------------------------------------
# A2-A5: lipid COMs and membrane-COM removal     pybilt/bilayer_analyzer/com_frame.py
# --------------------------------------------------------------------------------------
def build_beads(top, name_dict=None, multi_bead=False):
    """Atom gather lists of every COM bead.

    default           : one bead per residue over all its atoms (com_frame.py:163-168)
    name_dict         : atoms named name_dict[resname][0], then [1], ... (com_frame.py:63-77)
    + multi_bead      : one bead per listed name (com_frame.py:188-231)
    Returns (gather[T], seg[Nb+1], types[Nb], resids[Nb]).
    """
    rseg = top.segments()
    gather, seg, types, resids = [], [0], [], []
    names = np.asarray(top.names, dtype=object)
    for r in range(top.n_residues):
        lo, hi = int(rseg[r]), int(rseg[r + 1])
        rn = top.resnames[r]
        if name_dict is None:
            groups = [np.arange(lo, hi)]
        else:
            per_name = [lo + np.nonzero(names[lo:hi] == nm)[0] for nm in name_dict[rn]]
            groups = per_name if multi_bead else [np.concatenate(per_name)]
        for g in groups:
            gather.append(g)
            seg.append(seg[-1] + len(g))
            types.append(rn)
            resids.append(int(top.resids[r]))
    return (np.concatenate(gather).astype(np.int64), np.asarray(seg, dtype=np.int64),
            types, np.asarray(resids, dtype=np.int64))


def bead_masses(masses, gather, seg):
    """LipidCOM.mass = atom_group.total_mass() = masses.sum() (com_frame.py:95)."""
    return np.array([masses[gather[seg[i]:seg[i + 1]]].sum() for i in range(len(seg) - 1)])


def lipid_coms(P: np.ndarray, masses: np.ndarray, gather: np.ndarray, seg: np.ndarray) -> np.ndarray:
    """MDAnalysis center_of_mass per bead: (pos.astype(f64)*m[:,None]).sum(axis=0)/m.sum()
    with rows accumulated in atom order (com_frame.py:82-91, oracle/fake_mda.py)."""
    F = P.shape[0]
    nb = len(seg) - 1
    counts = np.diff(seg)
    out = np.empty((F, nb, 3))
    wsum_all = bead_masses(masses, gather, seg)
    for c in np.unique(counts):
        lip = np.nonzero(counts == c)[0]
        idx = gather[seg[lip][:, None] + np.arange(c)[None, :]]           # [n, c]
        w = masses[idx]
        acc = np.zeros((F, len(lip), 3))
        for a in range(int(c)):
            acc += P[:, idx[:, a], :].astype(np.float64) * w[None, :, a, None]
        out[:, lip, :] = acc / wsum_all[lip][None, :, None]
    return out


def membrane_com(U: np.ndarray, masses: np.ndarray) -> np.ndarray:
    """com_frame.py:175-179: per axis (f64(u[:,k])*m).sum()/m.sum(), NumPy pairwise sums."""
    total = masses.sum()
    F = U.shape[0]
    mem = np.empty((F, 3))
    for f in range(F):
        prod = np.ascontiguousarray((U[f].astype(np.float64) * masses[:, None]).T)   # [3, M]
        for k in range(3):
            mem[f, k] = prod[k].sum() / total
    return mem


def com_frames(X, box, masses, gather, seg, u_state=None, chunk=64, keep_v=None):
    """COMFrame for every analysed frame: wrapped COM from raw x, unwrap, membrane COM,
    v=f32(f64(u)-mem), unwrapped COM from v (com_frame.py:152-186).
    Returns dict(com, com_unwrap, mem_com, u_last)."""
    F = X.shape[0]
    nb = len(seg) - 1
    com = np.empty((F, nb, 3))
    comu = np.empty((F, nb, 3))
    mem = np.empty((F, 3))
    prev = u_state
    vs = []
    for c0 in range(0, F, chunk):
        c1 = min(F, c0 + chunk)
        U = unwrap(X[c0:c1], box[c0:c1], prev)
        prev = U[-1].copy()
        m = membrane_com(U, masses)
        V = (U.astype(np.float64) - m[:, None, :]).astype(np.float32)       # com_frame.py:180
        com[c0:c1] = lipid_coms(X[c0:c1], masses, gather, seg)
        comu[c0:c1] = lipid_coms(V, masses, gather, seg)
        mem[c0:c1] = m
        if keep_v is not None:              # the timestep as COMFrame leaves it, for the listed atoms
            vs.append(V[:, keep_v, :].copy())
    out = {"com": com, "com_unwrap": comu, "mem_com": mem, "u_last": prev}
    if keep_v is not None:
        out["v"] = np.concatenate(vs, axis=0)
    return out


# --------------------------------------------------------------------------------------
# B1/B2: leaflets            pybilt/bilayer_analyzer/bilayer_analyzer.py:826-847, 969-981
# --------------------------------------------------------------------------------------
def leaflet_labels(com_unwrap: np.ndarray, norm: int = 2, update_interval: int = 1) -> np.ndarray:
    F, N, _ = com_unwrap.shape
    lab = np.empty((F, N), dtype=np.uint8)
    for f in range(F):
        if f == 0 or (f % update_interval) == 0:
            z = np.ascontiguousarray(com_unwrap[f, :, norm])
            rc = lib().pc_leaflet_labels(_p(z), ctypes.c_int64(N), _p(lab[f]))
            if rc != 0:
                raise KeyError("")          # z == zavg (bilayer_analyzer.py:838-845)
        else:
            lab[f] = lab[f - 1]
    return lab


def orientation_labels(v_sel, sel_mass, seg, norm=2, update_interval=1):
    """BilayerAnalyzer._build_leaflets_orientation (bilayer_analyzer.py:849-876) on the rebuild frames:
    v_sel [F, n_sel, 3] the COMFrame-modified coordinates of the selected atoms, seg[2 i .. 2 i + 2] the
    tail / head atoms of lipid i."""
    F, N = v_sel.shape[0], (len(seg) - 1) // 2
    lab = np.empty((F, N), dtype=np.uint8)
    nv = np.zeros(3)
    nv[norm] = 1.0
    for f in range(F):
        if f != 0 and (f % update_interval) != 0:
            lab[f] = lab[f - 1]
            continue
        for i in range(N):
            com = []
            for a0, a1 in ((seg[2 * i], seg[2 * i + 1]), (seg[2 * i + 1], seg[2 * i + 2])):
                c = v_sel[f, a0:a1].astype(np.float64)
                w = sel_mass[a0:a1]
                com.append((c * w[:, None]).sum(axis=0) / w.sum())
            vec = com[1] - com[0]
            cost = np.dot(vec, nv) / np.sqrt(np.dot(vec, vec))
            if cost > 0.0:
                lab[f, i] = 1
            elif cost < 0.0:
                lab[f, i] = 0
            else:
                raise KeyError("")
    return lab


def orientation_selection(top, types, resids, orientation_atoms):
    names = np.asarray(top.names, dtype=object)
    resindex = np.asarray(top.resindex)
    sel, seg = [], [0]
    for resname, resid in zip(types, resids):
        res = [r for r in range(len(top.resnames)) if top.resnames[r] == resname and int(top.resids[r]) == int(resid)]
        for which in (0, 1):
            for r in res:
                sel += [int(a) for a in np.nonzero((resindex == r) & (names == orientation_atoms[resname][which]))[0]]
            seg.append(len(sel))
    sel = np.asarray(sel, dtype=np.int64)
    return sel, np.asarray(top.masses, dtype=np.float64)[sel], seg


class LeafletView:
    """Ordered member / group lists of one leaflet (pybilt/bilayer_analyzer/leaflet.py)."""

    def __init__(self, members: np.ndarray, types: Sequence[str]):
        self.members = _i64(members)
        self.group_names: List[str] = []
        tl = [types[i] for i in self.members]
        for t in tl:
            if t not in self.group_names:
                self.group_names.append(t)
        tarr = np.asarray(tl, dtype=object)
        self.groups = {g: self.members[tarr == g] for g in self.group_names}

    def has_group(self, name):
        return name == "all" or name in self.groups              # leaflet.py:201-210

    def get_group_indices(self, name):
        if name == "all" or name not in self.groups:             # leaflet.py:93-118 (unknown -> all)
            return self.members
        return self.groups[name]


def leaflet_views(label_row: np.ndarray, types: Sequence[str]) -> Dict[str, LeafletView]:
    return {"upper": LeafletView(np.nonzero(label_row == 1)[0], types),
            "lower": LeafletView(np.nonzero(label_row == 0)[0], types)}


def _indices_for(views, leaflet, group):
    """MSDProtocol index list (analysis_protocols.py:586-605): both = upper then lower."""
    if leaflet in ("upper", "lower"):
        return np.asarray(views[leaflet].get_group_indices(group))
    return np.concatenate([views["upper"].get_group_indices(group),
                           views["lower"].get_group_indices(group)])


# --------------------------------------------------------------------------------------
# C1/C2: MSD                 pybilt/bilayer_analyzer/analysis_protocols.py:582-667, 781-907
# --------------------------------------------------------------------------------------
def _frame_msd(cur: np.ndarray, ref: np.ndarray) -> np.ndarray:
    """Welford mean over lipids of (dx^2 + dy^2), for a stack of frames [f, n, 2]."""
    dr = cur - ref
    drs = dr * dr
    mag = np.ascontiguousarray(drs[..., 0] + drs[..., 1])       # drs_curr.sum() of 2 terms
    out = np.zeros(mag.shape[0])
    if mag.shape[1] > 0:
        lib().pc_welford_mean_rows(_p(mag), ctypes.c_int64(mag.shape[0]),
                                   ctypes.c_int64(mag.shape[1]), _p(out))
    return out


def msd(com_unwrap, times, idx, lateral=(0, 1)) -> np.ndarray:
    r = com_unwrap[:, idx][:, :, list(lateral)]
    out = np.empty((r.shape[0], 2))
    out[:, 0] = times
    out[:, 1] = _frame_msd(r, r[0:1])
    return out


def msd_multi_blocks(n_frames: int, n_tau: int, n_sigma: int) -> List[Tuple[int, int]]:
    """(origin frame, evaluated frame) of every finished block, by replaying the counter
    logic of MSDMultiProtocol.run_analysis (analysis_protocols.py:833-888)."""
    pairs: Dict[int, Tuple[int, int]] = {}
    origins = [0]
    counters = [0]
    tau_counter = 0
    sigma_counter = 0
    for f in range(n_frames):
        if sigma_counter == n_sigma:
            origins.append(f)
            counters.append(0)
        if sigma_counter <= n_sigma:
            for b in range(tau_counter, len(origins)):
                if counters[b] < n_tau:
                    counters[b] += 1
                if counters[b] == n_tau:
                    pairs[b] = (origins[b], f)
                    tau_counter += 1
            if sigma_counter == n_sigma:
                sigma_counter = 0
        sigma_counter += 1
    return [pairs[b] for b in sorted(pairs)], len(origins)


def msd_multi(com_unwrap, times, idx, n_tau=50, n_sigma=50, lateral=(0, 1)) -> np.ndarray:
    F = com_unwrap.shape[0]
    pairs, _ = msd_multi_blocks(F, n_tau, n_sigma)
    r = com_unwrap[:, idx][:, :, list(lateral)]
    vals = np.array([_frame_msd(r[e:e + 1], r[o:o + 1])[0] for o, e in pairs])
    vals = vals[vals > 0.0]
    tau = times[n_tau - 1] - times[0] if F >= n_tau else 0.0
    with np.errstate(all="ignore"), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        avg = vals.mean()
        err = vals.std() / np.sqrt(len(vals))
        return np.array([avg, err, avg / (4.0 * tau), err / (4.0 * tau)])


# --------------------------------------------------------------------------------------
# D2: COM lateral RDF        analysis_protocols.py:4900-5062
# --------------------------------------------------------------------------------------
def rdf(com, labels, types, box, leaflet="both", resname_1="first", resname_2="first",
        n_bins=25, range_inner=0.0, range_outer=25.0, lateral=(0, 1), cells=False):
    """``cells``: use the cell-list form of the pair loop (portc.c: pc_rdf_accumulate_cells, same
    arithmetic, for the 10^5-lipid leaflets of configs C4/C5); raises if a COM lies outside the box."""
    F = com.shape[0]
    lx, ly = lateral
    do_leaflet = ["upper", "lower"] if leaflet == "both" else [leaflet]
    edges = np.histogram([-1], bins=n_bins, range=[range_inner, range_outer])[1]
    count = np.zeros(n_bins, dtype=np.int64)
    area = Welford()
    N_a = N_b = N_tot = 0
    for f in range(F):
        views = leaflet_views(labels[f], types)
        if f == 0:
            ltypes = []
            for ln in do_leaflet:
                for g in views[ln].group_names:
                    if g not in ltypes:
                        ltypes.append(g)
            if resname_1 == "first":
                resname_1 = ltypes[0]
            if resname_2 == "first":
                resname_2 = ltypes[0]
        pos = np.ascontiguousarray(com[f][:, [lx, ly]])
        bx, by = box[f, lx], box[f, ly]
        for ln in do_leaflet:
            v = views[ln]
            N_tot += len(v.members)
            if v.has_group(resname_1) and v.has_group(resname_2):
                ia = _i64(v.get_group_indices(resname_1))
                ib = _i64(v.get_group_indices(resname_2))
                N_a += len(ia)
                N_b += len(ib)
                fn = lib().pc_rdf_accumulate_cells if cells else lib().pc_rdf_accumulate
                rc = fn(_p(pos), _p(ia), ctypes.c_int64(len(ia)), _p(ib),
                        ctypes.c_int64(len(ib)), ctypes.c_float(bx),
                        ctypes.c_float(by), ctypes.c_int(n_bins),
                        ctypes.c_double(range_inner), ctypes.c_double(range_outer),
                        _p(edges), _p(count))
                if cells and rc != 0:
                    raise ValueError("cell-list RDF oracle: a COM outside the box or a box below 3 cells")
        area.push(box[f][[lx, ly]].prod())                        # float32 product (:5004)
    return rdf_finalize(count, edges, N_a, N_b, N_tot, float(len(do_leaflet)), F, area.mean(),
                        resname_1, resname_2), count


def rdf_finalize(count, edges, N_a, N_b, N_tot, n_leaf, n_frames, box_area, resname_1, resname_2):
    """COMLateralRDFProtocol.get_data (analysis_protocols.py:5026-5062)."""
    shell = np.power(edges[1:], 2) - np.power(edges[:-1], 2)
    shell *= np.pi
    na = N_a / (n_leaf * n_frames)
    nb = N_b / (n_leaf * n_frames)
    n = na * nb
    if resname_1 == resname_2:
        n -= na
    elif resname_1 == "all" and resname_2 != "all":
        n -= nb
    elif resname_2 == "all" and resname_1 != "all":
        n -= na
    with np.errstate(all="ignore"):
        pair_density = n / box_area
        normf = 1.0 / (pair_density * shell)
        g = (count.astype(np.float64) * normf) / n_frames
    return g, 0.5 * (edges[:-1] + edges[1:])


# --------------------------------------------------------------------------------------
# D3: nearest-neighbour fraction      analysis_protocols.py:2174-2257
# --------------------------------------------------------------------------------------
def nnf(com, labels, types, box, times, leaflet="both", resname_1="first", resname_2="first",
        n_neighbors=5, lateral=(0, 1), return_counts=False):
    F = com.shape[0]
    lx, ly = lateral
    do_leaflet = ["upper", "lower"] if leaflet == "both" else [leaflet]
    running = Welford()
    out = np.zeros((F, 4))
    all_counts = []
    tarr = np.asarray(types, dtype=object)
    for f in range(F):
        views = leaflet_views(labels[f], types)
        # the "first frame" block runs every frame (analysis_protocols.py:2181: never cleared)
        ltypes, nlip = [], 0
        for ln in do_leaflet:
            nlip += len(views[ln].members)
            for g in views[ln].group_names:
                if g not in ltypes:
                    ltypes.append(g)
        if resname_1 == "first":
            resname_1 = ltypes[0]
        if resname_2 == "first":
            resname_2 = ltypes[0]
        if n_neighbors > nlip:
            n_neighbors = nlip - 1
        is_b = np.ascontiguousarray((tarr == resname_2).astype(np.uint8))
        pos = np.ascontiguousarray(com[f][:, [lx, ly]])
        bx, by = box[f, lx], box[f, ly]
        avg = Welford()
        frame_counts = []
        for ln in do_leaflet:
            v = views[ln]
            if v.has_group(resname_1) and v.has_group(resname_2):
                ia = _i64(v.get_group_indices(resname_1))
                iall = _i64(v.members)
                nb = np.zeros(len(ia), dtype=np.int32)
                lib().pc_nnf_counts(_p(pos), _p(ia), ctypes.c_int64(len(ia)), _p(iall),
                                    ctypes.c_int64(len(iall)), _p(is_b), ctypes.c_int(n_neighbors),
                                    ctypes.c_float(bx), ctypes.c_float(by), _p(nb))
                frame_counts.append(nb)
                for c in nb:
                    avg.push(float(c) / n_neighbors)
        fc = avg.mean()
        running.push(fc)
        out[f] = [times[f], fc, running.mean(), running.deviation()]
        all_counts.append(np.concatenate(frame_counts) if frame_counts else np.zeros(0, np.int32))
    return (out, all_counts) if return_counts else out


# --------------------------------------------------------------------------------------
# N3: Halperin-Nelson hexatic invariant      analysis_protocols.py:4337-4425
# --------------------------------------------------------------------------------------
def halperin_nelson(com, box, times, idx, lateral=(0, 1)) -> np.ndarray:
    """rows [time, mean_l |(1/6) sum_{j in 6nn(l)} exp(6i a_lj)|^2] over the lipids ``idx`` (the
    first frame's leaflet members).  Every pair (lower, higher position) is evaluated once with
    distance_euclidean_pbc; a stable sort of all pairs by distance hands each lipid its six nearest
    in (distance, partner index) order; the "angle" is the x component of the unit separation
    vector, used as an angle as the reference does."""
    import cmath
    lx, ly = lateral
    n = len(idx)
    out = []
    for f in range(com.shape[0]):
        bl = np.array([box[f, lx], box[f, ly]], dtype=np.float32)
        c = bl / 2.0
        pos = com[f][:, [lx, ly]]
        pairs = []
        for a in range(n - 1):
            for b in range(a + 1, n):
                va, vb = pos[idx[a]] - c, pos[idx[b]] - c
                dv = va - vb
                for k in range(2):
                    if abs(dv[k]) > bl[k] / 2.0:
                        dv[k] = bl[k] - abs(va[k]) - abs(vb[k])
                pairs.append((idx[a], idx[b], np.sqrt(np.dot(dv, dv)), dv))
        pairs.sort(key=lambda p: p[2])
        knn = {i: [] for i in idx}
        for i, j, r, dv in pairs:
            if len(knn[i]) < 6:
                knn[i].append((r, -dv))
            if len(knn[j]) < 6:
                knn[j].append((r, dv))
        hn = []
        for i in idx:
            acc = 0.0
            for r, vec in knn[i]:
                acc += cmath.exp(6j * (np.dot(vec, np.array([1.0, 0.0])) / r))
            hn.append(abs(acc / 6.0) ** 2)
        out.append([times[f], np.array(hn).mean()])
    return np.array(out)


# --------------------------------------------------------------------------------------
# N3: flip-flop events      analysis_protocols.py:3784-3833
# --------------------------------------------------------------------------------------
def flip_flop(labels, types, resids, times, frame_numbers) -> dict:
    """{resname: {'count', 'events': [[frame, time, resid, from, to], ...]}}.  Only the first leaflet
    ('upper') is examined; the reference leaflets are replaced whenever anything changed, i.e. they
    are always the previous frame's.  Events of a frame come in the iteration order of the Python
    sets ``curr.difference(ref)`` then ``ref.difference(curr)`` -- reproduced by making the same sets."""
    views0 = leaflet_views(labels[0], types)
    names = []
    for ln in ("upper", "lower"):
        for g in views0[ln].group_names:
            if g not in names:
                names.append(g)
    out = {r: {'count': 0, 'events': []} for r in names}
    ref = set(np.nonzero(labels[0] == 1)[0].tolist())
    for f in range(1, labels.shape[0]):
        curr = set(np.nonzero(labels[f] == 1)[0].tolist())
        fwd, bwd = curr.difference(ref), ref.difference(curr)
        for index in fwd:
            out[types[index]]['count'] += 1
            out[types[index]]['events'].append([frame_numbers[f], times[f], resids[index], 'lower', 'upper'])
        for index in bwd:
            out[types[index]]['count'] += 1
            out[types[index]]['events'].append([frame_numbers[f], times[f], resids[index], 'upper', 'lower'])
        if fwd or bwd:
            ref = curr
    return out


# --------------------------------------------------------------------------------------
# N3: average lateral displacement      analysis_protocols.py:3262-3339
# --------------------------------------------------------------------------------------
def ald(com_unwrap, times, idx, lateral=(0, 1)) -> np.ndarray:
    """rows [time, running mean over frames of L_t], L_t = sum_i |r_i(t) - r_i(0)| / n over the lipids
    ``idx`` (unwrapped lateral COMs; |.| = sqrt(np.dot(dr, dr)), np.sum of the list)."""
    run_ = Welford()
    ref = com_unwrap[0][idx][:, list(lateral)]
    out = np.zeros((com_unwrap.shape[0], 2))
    for f in range(com_unwrap.shape[0]):
        dr = com_unwrap[f][idx][:, list(lateral)] - ref
        m_dr = [np.sqrt(np.dot(v, v)) for v in dr]
        L = np.abs(m_dr).sum() / len(m_dr)
        run_.push(L)
        out[f] = [times[f], run_.mean()]
    return out


# --------------------------------------------------------------------------------------
# box-only analyses: apl_box (analysis_protocols.py:991-1006), acm (:3181-3207)
# --------------------------------------------------------------------------------------
def apl_box(dims, times, n_residues, lateral=(0, 1)) -> np.ndarray:
    """rows [time, 2 A_xy / N_lipids, running mean, running deviation]; float32 box arithmetic."""
    run_ = Welford()
    out = np.zeros((dims.shape[0], 4))
    for f in range(dims.shape[0]):
        plane = dims[f][0:3][list(lateral)]
        apl = plane[0] * plane[1] * 2.0 / n_residues
        run_.push(apl)
        out[f] = [times[f], apl, run_.mean(), run_.deviation()]
    return out


def acm(dims, times, norm=2, temperature=298.15) -> np.ndarray:
    """Area compressibility modulus rows [time, Ka, Ka error, area] from the running mean and
    variance of the lateral box area (box product / normal edge, float32)."""
    import scipy.constants as scicon
    run_ = Welford()
    out = []
    first = True
    with np.errstate(all="ignore"):
        for f in range(dims.shape[0]):
            d = dims[f][0:3]
            area = d.prod() / d[norm]
            run_.push(area)
            ka, ka_err = 1.0, 0.0
            if first:
                first = False
                ka = 0.0
            else:
                ka = (run_.mean() * scicon.k * temperature) / run_.variance()
                sd, mean, n = run_.deviation(), run_.mean(), run_.n
                var_of_var = (np.sqrt(2.0) * sd ** 2) / np.sqrt(n - 1)
                sd_of_var = np.sqrt(var_of_var)
                ka_err = np.sqrt((sd / mean) ** 2 + (sd_of_var / sd ** 2) ** 2) * ka
            ka *= 10.0 ** 23
            ka_err *= 10.0 ** 23
            out.append([times[f], ka, ka_err, area])
    return np.array(out)


def area_compressibility(dims, times, norm=2, temperature=298.15) -> np.ndarray:
    """ac (analysis_protocols.py:3431-3445): rows [time, var(A) / (<A> k T) * 1e-23]."""
    import scipy.constants as scicon
    run_ = Welford()
    out = []
    with np.errstate(all="ignore"):
        for f in range(dims.shape[0]):
            d = dims[f][0:3]
            area = d.prod() / d[norm]
            run_.push(area)
            x_t = run_.variance() / (run_.mean() * scicon.k * temperature)
            x_t *= 10.0 ** (-23)
            out.append([times[f], x_t])
    return np.array(out)


def volume_compressibility_modulus(dims, times, temperature=298.15) -> np.ndarray:
    """vcm (analysis_protocols.py:3071-3091): rows [time, <V> k T / dev(V)^2], 0 on the first frame."""
    import scipy.constants as scicon
    run_ = Welford()
    out = []
    with np.errstate(all="ignore"):
        for f in range(dims.shape[0]):
            run_.push(dims[f][0:3].prod())
            kv = 0.0 if f == 0 else (run_.mean() * scicon.k * temperature) / run_.deviation() ** 2
            out.append([times[f], kv])
    return np.array(out)


# --------------------------------------------------------------------------------------
# N3: lateral box-area fluctuation      analysis_protocols.py:4486-4498
# --------------------------------------------------------------------------------------
def area_fluctuation(box, times, lateral=(0, 1)) -> np.ndarray:
    """rows [time, area, sqrt(<A^2> - <A>^2)] with the float32 box product of the COM frame and
    Welford running means of A and A*A (float32 on the first push, running_stats.py:34-52)."""
    run_a, run_sq = Welford(), Welford()
    out = []
    for f in range(box.shape[0]):
        area = box[f][list(lateral)].prod()              # float32 product
        run_a.push(area)
        run_sq.push(area * area)
        if f == 0:
            out.append([times[f], area, 0.0])
        else:
            out.append([times[f], area, np.sqrt(run_sq.mean() - run_a.mean() ** 2)])
    return np.array(out)


# --------------------------------------------------------------------------------------
# N3: thickness from the leaflets' centres of mass      analysis_protocols.py:1154-1170
# --------------------------------------------------------------------------------------
def leaflet_distance(com, labels, times, norm=2) -> np.ndarray:
    """rows [time, |mean(com_upper) - mean(com_lower)|[norm], running mean, running deviation].
    ``coordinates(leaflet=...)`` (com_frame.py:369-385) collects the WRAPPED COMs of the leaflet's
    lipids in lipid order; ``ndarray.mean(axis=0)`` of that [n, 3] array adds the rows one after
    the other and divides by n."""
    F = com.shape[0]
    run_ = Welford()
    out = np.zeros((F, 4))
    for f in range(F):
        up = com[f][labels[f] == 1]
        lo = com[f][labels[f] == 0]
        cur = np.abs(up.mean(axis=0) - lo.mean(axis=0))[norm]
        run_.push(cur)
        out[f] = [times[f], cur, run_.mean(), run_.deviation()]
    return out


# --------------------------------------------------------------------------------------
# G1-G4: lipid grids         pybilt/lipid_grid/lipid_grid_opt.py, lipid_grid.py
# --------------------------------------------------------------------------------------
def grid_geometry(boxl: np.float32, nbins: int):
    """Edges/centres exactly as lipid_grid_opt.py:204-215 under NumPy >= 2 (float32
    linspace): returns (centres f64[n], incr f32)."""
    edges = np.linspace(0.0, boxl, nbins + 1, endpoint=True)
    incr = edges[1] - edges[0]
    h = incr / 2.0
    centres = np.zeros(nbins)
    for j in range(nbins):
        centres[j] = edges[j] + h
    return centres, incr


def grid_bins(box_xy, nxbins, nybins, area_per_bin=None):
    """Fixed-resolution bin counts (lipid_grid_opt.py:192-198)."""
    boxx, boxy = box_xy
    if area_per_bin is not None:
        nbins = np.sqrt((boxx * boxy) / area_per_bin)
        x2y_sr = np.sqrt(boxx / boxy)
        nxbins = int(nbins * x2y_sr)
        nybins = int(nbins / x2y_sr)
    return nxbins, nybins


def lipid_grid_frame(com_f, comu_f, label_f, box_f, nxbins=10, nybins=10, mode="opt",
                     plane=(0, 1), area_per_bin=None, lean_knn=False):
    """``lean_knn``: the 24-neighbour table without the n x n matrix (portc.c: pc_knn_lean).
    LipidGrids for one frame: per leaflet (grid_idx[nx,ny] i64, grid_z[nx,ny] f64),
    plus (x_incr, y_incr).  mode 'opt' = lipid_grid_opt.py:237-334 (what BilayerAnalyzer
    builds), 'exact' = lipid_grid.py:217-248."""
    ix, iy = plane
    iz = [i for i in (0, 1, 2) if i not in plane][0]
    box = box_f[list(plane)]
    boxx, boxy = box[ix], box[iy]                   # quirk: indexes the 2-vector with plane ids
    nx, ny = grid_bins((boxx, boxy), nxbins, nybins, area_per_bin)
    xc, xincr = grid_geometry(boxx, nx)
    yc_full, yincr = grid_geometry(boxy, ny)
    yc = np.zeros(ny)                               # y_centers loop bound uses x edges (:225)
    m = min(nx, ny)
    yc[:m] = yc_full[:m]
    pos = np.ascontiguousarray(com_f[:, [ix, iy]])
    z = np.ascontiguousarray(comu_f[:, iz])
    out = {}
    for name, lab in (("upper", 1), ("lower", 0)):
        idx = _i64(np.nonzero(label_f == lab)[0])
        n = len(idx)
        gidx = np.zeros((nx, ny), dtype=np.int64)
        gz = np.zeros((nx, ny))
        if mode == "exact":
            lib().pc_grid_exact(_p(pos), _p(z), _p(idx), ctypes.c_int64(n), _p(xc), ctypes.c_int(nx),
                                _p(yc), ctypes.c_int(ny), ctypes.c_float(boxx), ctypes.c_float(boxy),
                                _p(gidx), _p(gz))
        else:
            nn_k = 24
            knn = np.full((n, nn_k), -1, dtype=np.int64)
            (lib().pc_knn_lean if lean_knn else lib().pc_knn)(
                _p(pos), _p(idx), ctypes.c_int64(n), ctypes.c_int(nn_k),
                ctypes.c_float(boxx), ctypes.c_float(boxy), _p(knn))
            pos_of = np.full(com_f.shape[0], -1, dtype=np.int64)
            pos_of[idx] = np.arange(n)
            lib().pc_grid_opt(_p(pos), _p(z), _p(idx), ctypes.c_int64(n), _p(knn), ctypes.c_int(nn_k),
                              _p(pos_of), _p(xc), ctypes.c_int(nx), _p(yc), ctypes.c_int(ny),
                              ctypes.c_float(boxx), ctypes.c_float(boxy), _p(gidx), _p(gz))
            out["knn_" + name] = knn
        out[name] = (gidx, gz)
    out["incr"] = (xincr, yincr)
    return out


def grid_thickness(grid) -> Tuple[float, float]:
    """LipidGrids.average_thickness (lipid_grid_opt.py:496-526)."""
    t = grid["upper"][1] - grid["lower"][1]
    return t.mean(), t.std()


def grid_apl(grid, label_f, types):
    """LipidGrids.area_per_lipid (lipid_grid_opt.py:553-597): composite Welford mean and
    per-resname (mean, sample dev); lipids in upper-then-lower member order."""
    resnames = []
    views = leaflet_views(label_f, types)
    for ln in ("upper", "lower"):
        for g in views[ln].group_names:
            if g not in resnames:
                resnames.append(g)
    per = {r: Welford() for r in resnames}
    run = Welford()
    area_of = {}
    xincr, yincr = grid["incr"]
    apb = xincr * yincr
    for ln in ("upper", "lower"):
        gidx = grid[ln][0]
        cnt = np.bincount(gidx.reshape(-1), minlength=len(label_f))
        for i in views[ln].members:
            a = apb * int(cnt[i])
            area_of[int(i)] = a
            per[types[i]].push(a)
            run.push(a)
    return run.mean(), {r: (per[r].mean(), per[r].deviation()) for r in resnames}, area_of


# --------------------------------------------------------------------------------------
# T1: COMTraj               pybilt/com_trajectory/COMTraj.py:916-1117, 1142-1240
# --------------------------------------------------------------------------------------
def unwrap_frame_scalar(x, u_prev, abc):
    """pybilt/mda_tools/mda_unwrap.py:12-85: like unwrap_frame, but `dz + shift*C` is formed
    from NumPy float32 scalars, i.e. the loop conditions are evaluated in float32."""
    u = x.copy()
    for k in range(3):
        L = abc[k]
        h = L / 2.0
        d = x[:, k] - u_prev[:, k]
        s = np.zeros(x.shape[0])
        for i in np.nonzero((d > h) | (d < -h))[0]:
            di, shift = d[i], 0
            if di > h:
                shift -= 1
                while (di + shift * L) > h:
                    shift -= 1
            else:
                shift += 1
                while (di + shift * L) < -h:
                    shift += 1
            s[i] = shift
        u[:, k] = (x[:, k].astype(np.float64) + s * L).astype(np.float32)
    return u


def comtraj(top, traj, plane=(0, 1), norm=2, frames=None):
    """COMTraj.__init__: wrapped COMs, scalar unwrap, membrane COM = center_of_mass() of the
    selection (rows accumulated in atom order), unwrapped COMs, leaflets from frame 0."""
    fr = np.arange(traj.positions.shape[0]) if frames is None else np.asarray(frames)
    X = traj.positions[fr]
    box = np.array(traj.dimensions[fr][:, :3])
    for f in range(1, len(box)):
        if (box[f] == 0).any():
            box[f] = box[f - 1]
    gather, seg, types, _resids = build_beads(top)
    m = top.masses
    total = m.sum()
    F = X.shape[0]
    comu = np.empty((F, len(types), 3))
    prev = None
    for f in range(F):
        u = X[f] if prev is None else unwrap_frame_scalar(X[f], prev, box[f])
        prev = u
        mem = (u.astype(np.float64) * m[:, None]).sum(axis=0) / total
        v = (u.astype(np.float64) - mem).astype(np.float32)
        comu[f] = lipid_coms(v[None], m, gather, seg)[0]
    com = lipid_coms(X, m, gather, seg)
    labels0 = leaflet_labels(comu[:1], norm)[0]
    return {"com": com, "com_unwrap": comu, "labels0": labels0, "types": types,
            "times": np.asarray(traj.times)[fr]}


def comtraj_group_indices(labels0, types, leaflet, group):
    """COMTraj Leaflet.get_group_indices (:525-554): 'all' concatenates group by group."""
    out = []
    for lab in ([1, 0] if leaflet == "both" else [1 if leaflet == "upper" else 0]):
        members = np.nonzero(labels0 == lab)[0]
        names = []
        for i in members:
            if types[i] not in names:
                names.append(types[i])
        pick = [group] if (group != "all" and group in names) else names
        for g in pick:
            out += [int(i) for i in members if types[i] == g]
    return np.asarray(out, dtype=np.int64)


def comtraj_msd(ct, leaflet="both", group="all", plane=(0, 1)):
    idx = comtraj_group_indices(ct["labels0"], ct["types"], leaflet, group)
    res = msd(ct["com_unwrap"], ct["times"], idx, plane)
    res[0] = 0.0                                    # row 0 is never written (COMTraj.py:1203-1236)
    return res


# --------------------------------------------------------------------------------------
# N3: com_frame:rewrap, disp_vec, msd_dist, dc_cluster
#   pybilt/bilayer_analyzer/com_frame.py:233-270; analysis_protocols.py:1655-1897, 5523-5830, 2826-2983;
#   pybilt/common/distance_cutoff_clustering.py:27-71,127-205; pybilt/common/running_stats.py:364-431
# (plain loops: these run on the small parity cases only)
# --------------------------------------------------------------------------------------
def rewrap(com, com_unwrap, mem_com, box):
    out = com.copy()
    for f in range(com.shape[0]):
        for r in range(com.shape[1]):
            pos_c = com_unwrap[f, r] + mem_com[f]
            diff = pos_c - com[f, r]
            if np.sqrt(np.dot(diff, diff)) > 1.0:
                for i in range(3):
                    p_i, b = pos_c[i], box[f, i]
                    while p_i < 0.0:
                        p_i += b
                    while p_i > b:
                        p_i -= b
                    out[f, r, i] = p_i
    return out


def pbc_dist(v_a, v_b, bl):
    """distance_euclidean_pbc(v_a, v_b, bl, center='box_half')"""
    c = np.array(bl) / 2.0
    v_a, v_b = v_a - c, v_b - c
    d_v = v_a - v_b
    for i in range(len(d_v)):
        if np.absolute(d_v[i]) > bl[i] / 2.0:
            d_v[i] = bl[i] - np.absolute(v_a[i]) - np.absolute(v_b[i])
    return np.sqrt(np.dot(d_v, d_v))


def interval_pairs(md_frames, interval):
    pairs, last_g, last_md = [], 0, int(md_frames[0])
    for g in range(1, len(md_frames)):
        if int(md_frames[g]) - last_md == interval:
            pairs.append((last_g, g))
            last_g, last_md = g, int(md_frames[g])
    return pairs


def disp_vec(com, com_unwrap, labels, types, box, md_frames, kv, lateral=(0, 1)):
    lat = list(lateral)
    truth = lambda k: kv.get(k, "False") in ("True", "true")
    scale, to_max, wrapped = truth("scale"), truth("scale_to_max"), truth("wrapped")
    if scale and to_max:
        scale = False
    leaflet, resname = kv.get("leaflet", "both"), kv.get("resname", "all")
    out, boxes = [], []
    for gp, gc in interval_pairs(md_frames, int(kv.get("interval", 5))):
        idx = list(_indices_for(leaflet_views(labels[gc], types), leaflet, resname))
        bl = box[gp][lat]
        boxes.append(bl)
        ci, cj, cw = com_unwrap[gc][idx][:, lat], com_unwrap[gp][idx][:, lat], com[gp][idx][:, lat]
        if scale:
            ci, cj, cw = ci / bl, cj / bl, cw / bl
        vec = np.zeros((len(idx), 4))
        vec[:, 0:2] = cw if wrapped else cj
        vec[:, 2:4] = ci - cj
        out.append([vec, [types[i] for i in idx]])
    if to_max and boxes:
        b = np.array(boxes)
        mx, my = b[:, 0].max(), b[:, 1].max()
        scaled = []
        for vec, names in out:
            vec = vec.copy()
            vec[:, 0] /= mx; vec[:, 2] /= mx; vec[:, 1] /= my; vec[:, 3] /= my
            scaled.append([vec, names])
        return scaled
    return out


def binned_average(data, positions, n_bins, position_range):
    lower, upper = position_range
    edges = np.linspace(lower, upper, num=n_bins + 1, endpoint=True)
    bins = np.linspace(lower, upper, num=n_bins, endpoint=False)
    counts, sums = np.zeros(len(bins)).astype(np.int64), np.zeros(len(bins))
    for c_val, pos in zip(data, positions):
        for j in range(1, len(bins) + 1):
            if (pos >= edges[j - 1]) and (pos <= edges[j]):
                counts[j - 1] += 1
                sums[j - 1] += c_val
                break
    keep = [i for i in range(len(counts)) if counts[i] > 0]
    return bins[keep], sums[keep] / counts[keep]


def msd_dist(com, com_unwrap, labels, types, box, md_frames, kv, lateral=(0, 1)):
    lat = list(lateral)
    leaflet = kv.get("leaflet", "both")
    views0 = leaflet_views(labels[0], types)
    do_leaflet = ["upper", "lower"] if leaflet == "both" else [leaflet]
    ltypes = []
    for ln in do_leaflet:
        for g in views0[ln].group_names:
            if g not in ltypes:
                ltypes.append(g)
    r1n = ltypes[0] if kv.get("resname_1", "first") == "first" else kv["resname_1"]
    r2n = ltypes[0] if kv.get("resname_2", "first") == "first" else kv["resname_2"]
    r1, r2 = [], []
    for ln in (["upper", "lower"] if leaflet == "both" else [leaflet]):
        r1 += list(views0[ln].get_group_indices(r1n))
        r2 += list(views0[ln].get_group_indices(r2n))
    dist_avg = kv.get("dist_avg", "False") in ("True", "true")

    def ordered(f):
        bl = np.array([box[f][lat[0]], box[f][lat[1]]])
        table = {}
        for i in r1:
            d = sorted(pbc_dist(com[f][i][lat], com[f][j][lat], bl) for j in r2)
            table[i] = d
        return table
    sds, distances = [], []
    pick = 1 if r1n == r2n else 0
    for gp, gc in interval_pairs(md_frames, int(kv.get("interval", 5))):
        for i in r1:
            disp = com_unwrap[gc][i][lat] - com_unwrap[gp][i][lat]
            sds.append(np.dot(disp, disp))
        prev, curr = ordered(gp), ordered(gc)
        for key in prev:
            d = prev[key][pick]
            if dist_avg:
                d = (d + curr[key][pick]) / 2.0
            distances.append(d)
    return binned_average(np.array(sds), np.array(distances), int(kv.get("n_bins", 25)),
                          [float(kv.get("range_inner", 0.0)), float(kv.get("range_outer", 25.0))])


def dc_cluster(com, labels, types, box, times, kv, lateral=(0, 1)):
    lat = list(lateral)
    leaflet, cutoff = kv.get("leaflet", "upper"), float(kv.get("cutoff", 12.0))
    do_leaflet = ["upper", "lower"] if leaflet == "both" else [leaflet]
    resname = kv.get("resname", "first")
    out = {k: [] for k in ("clusters", "nclusters", "max_size", "min_size", "avg_size")}
    rs = {k: Welford() for k in ("nclusters", "max_size", "min_size", "avg_size")}
    for f in range(com.shape[0]):
        views = leaflet_views(labels[f], types)
        if resname == "first":
            names = []
            for ln in do_leaflet:
                for g in views[ln].group_names:
                    if g not in names:
                        names.append(g)
            resname = names[0]
        idx = []
        for ln in do_leaflet:
            idx += list(views[ln].get_group_indices(resname))
        pos = [com[f][i][lat] for i in idx]
        bl = box[f][lat]
        n = len(pos)
        adj = np.zeros((n, n), dtype=bool)
        for i in range(n - 1):
            for j in range(i + 1, n):
                adj[i, j] = adj[j, i] = pbc_dist(pos[i], pos[j], bl) <= cutoff
        clusters, master = [], list(range(n))
        while master:
            nb, taken, i = [master[0]], {master[0]}, 0
            while i < len(nb):
                for b in master:
                    if b not in taken and adj[nb[i], b]:
                        nb.append(b)
                        taken.add(b)
                i += 1
            master = [b for b in master if b not in taken]
            if len(nb) > 1:
                clusters.append(nb)
        ncl, mn, mx, avg = len(clusters), 0, 0, 0.0
        for cl in clusters:
            m = len(cl)
            if m > mx:
                mx = m
            elif m > 0 and mx == 0:
                mn = m
            elif mx > 0 and m < mx:
                mn = m
            avg += m
        if ncl > 0:
            avg /= ncl
        out["clusters"].append(clusters)
        for key, val in (("nclusters", ncl), ("max_size", mx), ("min_size", mn), ("avg_size", avg)):
            rs[key].push(val)
            out[key].append([times[f], val, rs[key].mean(), rs[key].deviation()])
    return {k: (v if k == "clusters" else np.array(v)) for k, v in out.items()}


# --------------------------------------------------------------------------------------
# driver: the BilayerAnalyzer frame loop over arrays (bilayer_analyzer.py:901-1031)
# --------------------------------------------------------------------------------------
def parse_analysis(s: str):
    w = s.split()
    key, a_id = w[0], w[1]
    kv = {w[i]: w[i + 1] for i in range(2, len(w) - 1, 2)}
    return key, a_id, kv


def analysed_frames(n_total: int, frame_range) -> np.ndarray:
    first, last, interval = frame_range
    if last < 0:
        last += n_total
    return np.arange(first, last + 1, interval)


def run(top, traj, analyses=(), rep_settings=None, settings=None, frame_range=(0, -1, 1),
        return_reps=False, grid_mode="opt"):
    """Array-at-a-time equivalent of BilayerAnalyzer.run_analysis + get_analysis_data for the
    in-scope analyses.  Returns {analysis_id: data} with the reference's output shapes."""
    rep_settings = rep_settings or {}
    settings = settings or {}
    cf = rep_settings.get("com_frame", {})
    lf = rep_settings.get("leaflets", {})
    lgs = rep_settings.get("lipid_grid", {})
    norm = settings.get("norm", 2)
    lateral = tuple(settings.get("lateral", [0, 1]))
    fr = analysed_frames(traj.positions.shape[0], frame_range)
    X = traj.positions[fr]
    dims = traj.dimensions[fr]
    box = dims[:, :3]
    times = np.asarray(traj.times)[fr]
    gather, seg, types, resids = build_beads(top, cf.get("name_dict"), cf.get("multi_bead", False))
    orient = lf.get("assign_method", "avg_norm") == "orientation"
    if orient:
        o_sel, o_mass, o_seg = orientation_selection(top, types, resids, lf["orientation_atoms"])
    reps = com_frames(X, box, top.masses, gather, seg, keep_v=o_sel if orient else None)
    if cf.get("rewrap"):
        reps["com"] = rewrap(reps["com"], reps["com_unwrap"], reps["mem_com"], box)
    com, comu = reps["com"], reps["com_unwrap"]
    labels = (orientation_labels(reps["v"], o_mass, o_seg, norm, lf.get("update_interval", 1)) if orient
              else leaflet_labels(comu, norm, lf.get("update_interval", 1)))
    out = {}
    views0 = leaflet_views(labels[0], types)
    grids = None
    for a in analyses:
        key, a_id, kv = parse_analysis(a)
        if key == "msd":
            idx = _indices_for(views0, kv.get("leaflet", "both"), kv.get("resname", "all"))
            out[a_id] = msd(comu, times, idx, lateral)
        elif key == "msd_multi":
            idx = _indices_for(views0, kv.get("leaflet", "both"), kv.get("resname", "all"))
            out[a_id] = msd_multi(comu, times, idx, int(kv.get("n_tau", 50)),
                                  int(kv.get("n_sigma", 50)), lateral)
        elif key == "com_lateral_rdf":
            (g, bins), cnt = rdf(com, labels, types, box, kv.get("leaflet", "both"),
                                 kv.get("resname_1", "first"), kv.get("resname_2", "first"),
                                 int(kv.get("n_bins", 25)), float(kv.get("range_inner", 0.0)),
                                 float(kv.get("range_outer", 25.0)), lateral)
            out[a_id] = (g, bins)
            out[a_id + ":count"] = cnt
        elif key == "nnf":
            res, counts = nnf(com, labels, types, box, times, kv.get("leaflet", "both"),
                              kv.get("resname_1", "first"), kv.get("resname_2", "first"),
                              int(kv.get("n_neighbors", 5)), lateral, return_counts=True)
            out[a_id] = res
            out[a_id + ":counts"] = counts
        elif key == "halperin_nelson":
            leaf = kv.get("leaflet", "upper")
            v = views0[leaf if leaf in ("upper", "lower") else "upper"]
            out[a_id] = halperin_nelson(com, box, times, list(v.members), lateral)
        elif key == "flip_flop":
            out[a_id] = flip_flop(labels, types, resids, times, fr)
        elif key == "disp_vec":
            out[a_id] = disp_vec(com, comu, labels, types, box, fr, kv, lateral)
        elif key == "msd_dist":
            out[a_id] = msd_dist(com, comu, labels, types, box, fr, kv, lateral)
        elif key == "dc_cluster":
            out[a_id] = dc_cluster(com, labels, types, box, times, kv, lateral)
        elif key == "ald":
            idx = _indices_for(views0, kv.get("leaflet", "both"), kv.get("resname", "all"))
            out[a_id] = ald(comu, times, idx, lateral)
        elif key == "ac":
            out[a_id] = area_compressibility(dims, times, norm, float(kv.get("temperature", 298.15)))
        elif key == "vcm":
            out[a_id] = volume_compressibility_modulus(dims, times, float(kv.get("temperature", 298.15)))
        elif key == "apl_box":
            out[a_id] = apl_box(dims, times, top.n_residues, lateral)
        elif key == "acm":
            out[a_id] = acm(dims, times, norm, float(kv.get("temperature", 298.15)))
        elif key == "area_fluctuation":
            out[a_id] = area_fluctuation(box, times, lateral)
        elif key == "thickness_leaflet_distance":
            if int(lf.get("update_interval", 1)) != 1 and len(fr) > 1:
                # frames between leaflet rebuilds carry no leaflet names: the reference's mean of an
                # empty selection is a nan scalar and indexing it raises (analysis_protocols.py:1158)
                raise IndexError("invalid index to scalar variable.")
            out[a_id] = leaflet_distance(com, labels, times, norm)
        elif key in ("apl_grid", "thickness_grid"):
            if grids is None:
                grids = [lipid_grid_frame(com[f], comu[f], labels[f], box[f],
                                          lgs.get("n_xbins", 10), lgs.get("n_ybins", 10),
                                          grid_mode, lateral, lgs.get("area_per_bin"))
                         for f in range(len(fr))]
            if key == "thickness_grid":
                run_ = Welford()
                rows = np.zeros((len(fr), 4))
                for f, g in enumerate(grids):
                    t = grid_thickness(g)[0]
                    run_.push(t)
                    rows[f] = [times[f], t, run_.mean(), run_.deviation()]
                out[a_id] = rows
            else:
                comp, per = Welford(), {}
                res = {"composite": []}
                for f, g in enumerate(grids):
                    avg, by_res, _ = grid_apl(g, labels[f], types)
                    comp.push(avg)
                    if f == 0:
                        for r in by_res:
                            per[r] = Welford()
                            res[r] = []
                    res["composite"].append([times[f], avg, comp.mean(), comp.deviation()])
                    for r in by_res:
                        per[r].push(by_res[r][0])
                        res[r].append([times[f], by_res[r][0], per[r].mean(), per[r].deviation()])
                out[a_id] = {k: np.array(v) for k, v in res.items()}
        else:
            raise RuntimeError("analysis key outside the hot path: " + key)
    if return_reps:
        reps.update(labels=labels, types=types, resids=resids, box=box, times=times, grids=grids,
                    gather=gather, seg=seg)
        return out, reps
    return out
