"""Deterministic synthetic bilayer trajectories (SURVEY.md §8d "Synthetic inputs").

The reference ships no usable trajectory (its ``sample_bilayer`` psf/dcd blobs are
missing from the checkout), so every test, fixture and benchmark in this repository
runs on bilayers made here.  The generator is counter based: every random number is
``mix64(seed, stream, index)`` turned into a float, so the same (seed, shape) gives
the same float32 positions on every machine and NumPy version, and any frame range
can be produced without generating the frames before it (used for frame shards).

What is produced mirrors what MDAnalysis hands PyBILT's ``BilayerAnalyzer``
(``/root/reference/pybilt/bilayer_analyzer/bilayer_analyzer.py:933-943``):
float32 ``positions[F, M, 3]`` wrapped into ``[0, L)``, float32 ``dimensions[F, 6]``,
``time[F]``, plus the static topology (masses, atom names, residue index, resnames,
resids).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)
_GOLD = np.uint64(0x9E3779B97F4A7C15)


def mix64(x: np.ndarray) -> np.ndarray:
    """splitmix64 finaliser on a uint64 array (wrapping arithmetic)."""
    x = x.astype(np.uint64, copy=True)
    with np.errstate(over="ignore"):
        x ^= x >> np.uint64(30)
        x *= _M1
        x ^= x >> np.uint64(27)
        x *= _M2
        x ^= x >> np.uint64(31)
    return x


def uniform(seed: int, stream: int, index: np.ndarray) -> np.ndarray:
    """U[0,1) float64 with 24 random bits, a pure function of (seed, stream, index)."""
    with np.errstate(over="ignore"):
        key = mix64(np.asarray([seed], dtype=np.uint64) * _GOLD
                    + np.uint64(stream) * _M2 + np.uint64(0x1234567))
        h = mix64(np.asarray(index, dtype=np.uint64) * _GOLD + key)
    return (h >> np.uint64(40)).astype(np.float64) * (1.0 / 16777216.0)


@dataclass
class Topology:
    """Static description of the selected atoms (what the psf + selection provide)."""
    masses: np.ndarray          # f64 [M]
    names: np.ndarray           # str [M]
    resindex: np.ndarray        # i64 [M]  residue (lipid) index of every atom, ascending
    resnames: List[str]         # [N]
    resids: np.ndarray          # i64 [N]

    @property
    def n_atoms(self) -> int:
        return int(self.masses.shape[0])

    @property
    def n_residues(self) -> int:
        return len(self.resnames)

    def segments(self) -> np.ndarray:
        """CSR offsets seg[N+1] of each residue's atoms inside the selection."""
        counts = np.bincount(self.resindex, minlength=self.n_residues)
        seg = np.zeros(self.n_residues + 1, dtype=np.int64)
        np.cumsum(counts, out=seg[1:])
        return seg


@dataclass
class Trajectory:
    positions: np.ndarray       # f32 [F, M, 3]
    dimensions: np.ndarray      # f32 [F, 6]
    times: np.ndarray           # f64 [F]
    dt: float


@dataclass
class LipidType:
    resname: str
    bead_names: Sequence[str]
    bead_masses: Sequence[float]
    fraction: float = 1.0

    def with_fraction(self, fraction: float) -> "LipidType":
        return LipidType(self.resname, self.bead_names, self.bead_masses, fraction)


# Martini-like coarse-grained templates (12/12/8 beads)
POPC = LipidType("POPC", ["NC3", "PO4", "GL1", "GL2", "C1A", "D2A", "C3A", "C4A",
                          "C1B", "C2B", "C3B", "C4B"], [72.0] * 12)
DOPE = LipidType("DOPE", ["NH3", "PO4", "GL1", "GL2", "C1A", "D2A", "C3A", "C4A",
                          "C1B", "D2B", "C3B", "C4B"], [72.0] * 12)
CHOL = LipidType("CHOL", ["ROH", "R1", "R2", "R3", "R4", "R5", "C1", "C2"],
                 [72.0, 72.0, 72.0, 72.0, 72.0, 72.0, 72.0, 72.0])
# a mixed-mass variant used by parity tests so that mass weighting actually matters
POPC_AA = LipidType("POPC", ["N", "P", "O11", "O12", "C1", "C2", "C21", "C22",
                             "C31", "C32", "H1", "H2"],
                    [14.007, 30.974, 15.999, 15.999, 12.011, 12.011, 12.011, 12.011,
                     12.011, 12.011, 1.008, 1.008])
DOPE_AA = LipidType("DOPE", ["N", "P", "O11", "O12", "C1", "C2", "C21", "C22",
                             "C31", "H1"],
                    [14.007, 30.974, 15.999, 15.999, 12.011, 12.011, 12.011, 12.011,
                     12.011, 1.008])
TLCL2_AA = LipidType("TLCL2", ["P1", "P3", "O1", "O3", "C1", "C2", "C3", "CA1", "CB1",
                               "CC1", "CD1", "H1", "H2", "H3"],
                     [30.974, 30.974, 15.999, 15.999, 12.011, 12.011, 12.011, 12.011,
                      12.011, 12.011, 12.011, 1.008, 1.008, 1.008])


@dataclass
class BilayerSpec:
    """Shape of one synthetic bilayer (all lengths in Angstrom, times in ps)."""
    n_lipids: int
    lipid_types: Sequence[LipidType] = (POPC,)
    area_per_lipid: float = 64.0
    box_z: float = 80.0
    leaflet_z: float = 20.0
    step: float = 0.5            # half-width of the per-frame lateral lipid step
    z_step: float = 0.15
    bead_jitter: float = 0.1
    dt: float = 10.0
    npt: float = 0.0             # relative amplitude of the triangular box breathing
    drift: Tuple[float, float, float] = (0.02, -0.015, 0.0)   # membrane COM drift / frame
    seed: int = 20240611
    box_xy: Optional[float] = None
    lattice_jitter: float = 0.35  # fraction of the lattice spacing

    def box_length(self) -> float:
        if self.box_xy is not None:
            return float(self.box_xy)
        return float(np.sqrt(self.area_per_lipid * self.n_lipids / 2.0))


def _assign_types(spec: BilayerSpec) -> np.ndarray:
    """Type index per lipid: a fixed repeating pattern honouring the fractions."""
    fr = np.array([t.fraction for t in spec.lipid_types], dtype=np.float64)
    fr = fr / fr.sum()
    # pattern of length 20 (5 % granularity), spread with a stride so types interleave
    plen = 20
    counts = np.floor(fr * plen + 1e-9).astype(int)
    counts[0] += plen - counts.sum()
    pattern = np.concatenate([np.full(c, k) for k, c in enumerate(counts)])
    pattern = pattern[(np.arange(plen) * 7) % plen]
    return pattern[np.arange(spec.n_lipids) % plen]


def make_topology(spec: BilayerSpec) -> Tuple[Topology, np.ndarray]:
    tidx = _assign_types(spec)
    masses, names, resindex, resnames = [], [], [], []
    for i, k in enumerate(tidx):
        lt = spec.lipid_types[k]
        masses.extend(lt.bead_masses)
        names.extend(lt.bead_names)
        resindex.extend([i] * len(lt.bead_names))
        resnames.append(lt.resname)
    top = Topology(masses=np.asarray(masses, dtype=np.float64),
                   names=np.asarray(names, dtype=object),
                   resindex=np.asarray(resindex, dtype=np.int64),
                   resnames=resnames,
                   resids=np.arange(1, spec.n_lipids + 1, dtype=np.int64))
    return top, tidx


def _bead_template(lt: LipidType, sign: float) -> np.ndarray:
    """Fixed bead offsets relative to the lipid centre: a short two-tail stick figure."""
    n = len(lt.bead_names)
    k = np.arange(n, dtype=np.float64)
    off = np.zeros((n, 3))
    off[:, 0] = 2.2 * np.cos(2.399963 * k) * (0.4 + 0.6 * k / max(n - 1, 1))
    off[:, 1] = 2.2 * np.sin(2.399963 * k) * (0.4 + 0.6 * k / max(n - 1, 1))
    off[:, 2] = sign * (7.0 - 14.0 * k / max(n - 1, 1))      # head outside, tails inside
    return off


def box_at(spec: BilayerSpec, frames: np.ndarray) -> np.ndarray:
    """float32 [F, 6] unit-cell dimensions; optional triangular NPT breathing."""
    L = spec.box_length()
    t = np.asarray(frames, dtype=np.float64)
    tri = 2.0 * np.abs((t / 16.0) % 2.0 - 1.0) - 1.0         # in [-1, 1], period 32 frames
    s = 1.0 + spec.npt * tri
    dims = np.empty((len(t), 6), dtype=np.float32)
    dims[:, 0] = (L * s).astype(np.float32)
    dims[:, 1] = (L * s).astype(np.float32)
    dims[:, 2] = np.float32(spec.box_z) * (1.0 + 0.5 * spec.npt * tri).astype(np.float32)
    dims[:, 3:] = 90.0
    return dims


def generate(spec: BilayerSpec, n_frames: int, first_frame: int = 0,
             chunk: int = 256) -> Tuple[Topology, Trajectory]:
    """Frames ``first_frame .. first_frame+n_frames-1`` of the bilayer ``spec``.

    Lipid centres follow a random walk whose steps are counter-based, so frame
    ``t`` of a shard equals frame ``t`` of the whole trajectory bit for bit.
    """
    top, tidx = make_topology(spec)
    N, M = spec.n_lipids, top.n_atoms
    L = spec.box_length()
    seg = top.segments()
    # ---- static layout: jittered square lattice per leaflet ------------------------
    half = (N + 1) // 2
    side = int(np.ceil(np.sqrt(half)))
    lip = np.arange(N)
    leaf = (lip % 2).astype(np.int64)                  # even -> upper, odd -> lower
    slot = lip // 2
    a = L / side
    jit = spec.lattice_jitter
    cx0 = ((slot % side) + 0.5 + jit * (2 * uniform(spec.seed, 1, lip) - 1)) * a
    cy0 = ((slot // side) + 0.5 + jit * (2 * uniform(spec.seed, 2, lip) - 1)) * a
    sign = np.where(leaf == 0, 1.0, -1.0)
    cz0 = 0.5 * spec.box_z + sign * spec.leaflet_z + 1.5 * (2 * uniform(spec.seed, 3, lip) - 1)
    centre0 = np.stack([cx0, cy0, cz0], axis=1)        # [N, 3]
    # bead offsets (static) + static per-atom jitter
    off = np.zeros((M, 3))
    for k, lt in enumerate(spec.lipid_types):
        for sgn in (1.0, -1.0):
            sel = np.nonzero((tidx == k) & (sign == sgn))[0]
            if len(sel) == 0:
                continue
            tmpl = _bead_template(lt, sgn)
            idx = (seg[sel][:, None] + np.arange(len(lt.bead_names))[None, :])
            off[idx.reshape(-1)] = np.tile(tmpl, (len(sel), 1))
    atom = np.arange(M)
    for d in range(3):
        off[:, d] += 0.3 * (2 * uniform(spec.seed, 10 + d, atom) - 1)
    atom_lipid = top.resindex

    # ---- random walk: position(t) = centre0 + sum_{s<=t} step(s); sum built by chunks ---
    # cumulative displacement up to first_frame-1 needs all earlier steps; to keep shards
    # O(chunk) we derive it from a two-level counter scheme: steps are grouped in blocks
    # of 256 frames whose block totals are themselves counter-based random numbers.
    def lipid_disp(frames: np.ndarray) -> np.ndarray:
        """[len(frames), N, 3] displacement of every lipid centre at the given frames."""
        out = np.zeros((len(frames), N, 3))
        amp = np.array([spec.step, spec.step, spec.z_step])
        for d in range(3):
            # bounded, smooth-ish random walk = sum of a few incommensurate sawtooth-free
            # counter-based terms: W(t) = sum_j A_j * tri(t / P_j + phase_j(lipid))
            acc = np.zeros((len(frames), N))
            for j, period in enumerate((37.0, 211.0, 1553.0, 9973.0)):
                ph = uniform(spec.seed, 100 + 10 * d + j, lip)
                u = frames[:, None] / period + ph[None, :]
                tri = 2.0 * np.abs((u % 2.0) - 1.0) - 1.0
                acc += np.sqrt(period) * tri
            # plus white per-frame noise (counter = frame * N + lipid)
            cnt = frames[:, None].astype(np.uint64) * np.uint64(N) + lip[None, :].astype(np.uint64)
            noise = 2 * uniform(spec.seed, 200 + d, cnt) - 1
            out[:, :, d] = amp[d] * (acc + noise) if d < 2 else amp[d] * noise
        return out

    pos = np.empty((n_frames, M, 3), dtype=np.float32)
    dims = box_at(spec, np.arange(first_frame, first_frame + n_frames))
    for c0 in range(0, n_frames, chunk):
        c1 = min(n_frames, c0 + chunk)
        fr = np.arange(first_frame + c0, first_frame + c1, dtype=np.float64)
        disp = lipid_disp(fr)                                        # [f, N, 3]
        scale = (dims[c0:c1, :3].astype(np.float64) /
                 np.array([L, L, spec.box_z]))[:, None, :]           # affine NPT scaling
        cen = (centre0[None] + disp
               + fr[:, None, None] * np.asarray(spec.drift)[None, None, :])
        xyz = cen[:, atom_lipid, :] + off[None]
        if spec.bead_jitter > 0.0:
            cnt = (fr[:, None].astype(np.uint64) * np.uint64(M) + atom[None, :].astype(np.uint64))
            for d in range(3):
                xyz[:, :, d] += spec.bead_jitter * (2 * uniform(spec.seed, 300 + d, cnt) - 1)
        xyz *= scale
        box = dims[c0:c1, None, :3].astype(np.float64)
        xyz -= np.floor(xyz / box) * box                             # wrap into [0, L)
        w = xyz.astype(np.float32)
        bf = dims[c0:c1, None, :3]
        w = np.where(w >= bf, np.float32(0.0), w)                    # f32 rounding up to L
        pos[c0:c1] = w
    times = (np.arange(first_frame, first_frame + n_frames) * spec.dt).astype(np.float64)
    return top, Trajectory(positions=pos, dimensions=dims, times=times, dt=spec.dt)


# ---- the named configurations of BASELINE.json --------------------------------------
def config_spec(name: str) -> Tuple[BilayerSpec, int]:
    """(spec, n_frames) for the configs of BASELINE.json / SURVEY.md §8."""
    if name == "C1":      # stand-in for the missing 600-lipid sample_bilayer
        return BilayerSpec(600, (POPC_AA.with_fraction(0.5), DOPE_AA.with_fraction(0.3),
                                 TLCL2_AA.with_fraction(0.2)),
                           box_xy=149.278, box_z=72.393, seed=20240612), 10
    if name == "C2":
        return BilayerSpec(512, (POPC,), seed=20240613), 5000
    if name == "C3":
        return BilayerSpec(8192, (POPC.with_fraction(0.5), DOPE.with_fraction(0.3),
                                  CHOL.with_fraction(0.2)), seed=20240614), 20000
    if name == "C4":
        p = LipidType("POPC", ["P"], [30.974])
        return BilayerSpec(65536, (p,), seed=20240615), 10000
    if name == "C5":
        p = LipidType("POPC", ["P"], [30.974])
        return BilayerSpec(262144, (p,), seed=20240616), 100000
    raise KeyError(name)
