"""Native readers for the files PyBILT users hand to ``BilayerAnalyzer`` -- CHARMM / NAMD PSF
topologies and DCD trajectories (the reference's bundled ``sample_bilayer`` is a psf + dcd pair,
BASELINE configs[0]) -- and the subset of the MDAnalysis selection language its scripts use.

The reference gets all of this from MDAnalysis (``pybilt/bilayer_analyzer/mda_data.py:14-38``:
``mda.Universe(psf, dcd)``, ``select_atoms``, ``ts.positions / dimensions / time``).  MDAnalysis
is not part of this image, so what it would return is restated here from the file formats:

* PSF ``!NATOM`` records: segid, resid, resname, atom name, type, charge, mass; residues are the
  runs of consecutive atoms with one (resid, resname, segid) (MDAnalysis ``PSFParser``).
* DCD: Fortran unformatted records -- 84-byte ``CORD`` header (frames, istart, nsavc, delta,
  unit-cell flag, CHARMM version), title block, atom count, then per frame an optional unit-cell
  record of six doubles ``A, gamma, B, beta, alpha, C`` followed by the X, Y and Z records of
  float32.  Angles stored as cosines (CHARMM >= 25 / NAMD) are turned into degrees the way
  MDAnalysis' DCDReader does (``90 - asin(c) * 90 / (pi/2)``); ``dimensions`` are float32
  ``[A, B, C, alpha, beta, gamma]``; ``time = (frame + istart / nsavc) * delta * nsavc`` with delta
  in AKMA units (4.888821e-2 ps).  Parity of this piece against MDAnalysis itself is UNPINNED
  (nothing to execute here); positions are the file's float32 values bit for bit either way.

X
``[atom, xyz]`` layout ON THE DEVICE (``pbk_planes_to_aos``), together with the atom selection,
so the host never touches individual atoms.
"""
from __future__ import annotations

import fnmatch
import os
import struct
import sys
from concurrent.futures import ThreadPoolExecutor
from typing import List, Optional, Sequence, Tuple

import numpy as np

AKMA_PS = 4.888821e-2          # one AKMA time unit in ps (MDAnalysis.units)

_POOLS = {}


def _pool(threads: int) -> ThreadPoolExecutor:
    if threads not in _POOLS:
        _POOLS[threads] = ThreadPoolExecutor(max_workers=threads)
    return _POOLS[threads]


# ---------------------------------------------------------------------------------------
# PSF
# ---------------------------------------------------------------------------------------
class PSFTopology(object):
    """Per-atom arrays of a PSF file: segids, resids, resnames, names, types, charges, masses."""

    def __init__(self, path: str):
        self.path = path
        with open(path) as fh:
            lines = fh.read().splitlines()
        if not lines or not lines[0].startswith("PSF"):
            raise RuntimeError("%s is not a PSF file" % path)
        start = None
        for k, ln in enumerate(lines):
            if "!NATOM" in ln:
                start = k
                break
        if start is None:
            raise RuntimeError("%s has no !NATOM section" % path)
        n = int(lines[start].split()[0])
        segids, resids, resnames, names, types, charges, masses = [], [], [], [], [], [], []
        for ln in lines[start + 1:start + 1 + n]:
            f = ln.split()
            if len(f) < 8:
                raise RuntimeError("unparsable PSF atom record: %r" % ln)
            segids.append(f[1]); resids.append(int(f[2])); resnames.append(f[3]); names.append(f[4])
            types.append(f[5]); charges.append(float(f[6])); masses.append(float(f[7]))
        if len(masses) != n:
            raise RuntimeError("%s: expected %d atoms, found %d" % (path, n, len(masses)))
        self.n_atoms = n
        self.segids = np.asarray(segids, dtype=object)
        self.resids = np.asarray(resids, dtype=np.int64)
        self.resnames = np.asarray(resnames, dtype=object)
        self.names = np.asarray(names, dtype=object)
        self.types = np.asarray(types, dtype=object)
        self.charges = np.asarray(charges, dtype=np.float64)
        self.masses = np.asarray(masses, dtype=np.float64)
        # residues: runs of consecutive atoms with one (resid, resname, segid)
        change = np.ones(n, dtype=bool)
        if n > 1:
            change[1:] = ((self.resids[1:] != self.resids[:-1]) | (self.resnames[1:] != self.resnames[:-1]) |
                          (self.segids[1:] != self.segids[:-1]))
        self.resindex = np.cumsum(change) - 1


def write_psf(path: str, segids, resids, resnames, names, masses, types=None, charges=None):
    """Minimal standard-format PSF (what the tests and `synthetic` exports need)."""
    n = len(masses)
    types = types if types is not None else names
    charges = charges if charges is not None else np.zeros(n)
    with open(path, "w") as fh:
        fh.write("PSF\n\n       1 !NTITLE\n REMARKS written by pybilt_b200.formats.write_psf\n\n")
        fh.write("%8d !NATOM\n" % n)
        for i in range(n):
            fh.write("%8d %-4s %-4d %-4s %-4s %-4s %10.6f %13.4f %11d\n" % (
                i + 1, segids[i], int(resids[i]), resnames[i], names[i], types[i], charges[i],
                masses[i], 0))
        fh.write("\n%8d !NBOND: bonds\n\n" % 0)


# ---------------------------------------------------------------------------------------
# selection language (subset)
# ---------------------------------------------------------------------------------------
_KEYWORDS = ("all", "resname", "name", "segid", "type", "resid", "resnum", "index", "bynum")
_OPS = ("and", "or", "not", "(", ")")


def select_atoms(top: PSFTopology, selection: str) -> np.ndarray:
    """Boolean mask of the atoms matched by an MDAnalysis-style selection string.

    Supported: ``all``; ``resname / name / segid / type V1 V2 ...`` (shell wildcards);
    ``resid / resnum A B:C D-E``; ``index`` (0-based) / ``bynum`` (1-based) ranges; ``not``,
    ``and``, ``or``, parentheses.  Anything else raises ``RuntimeError``."""
    toks = selection.replace("(", " ( ").replace(")", " ) ").split()
    pos = [0]
    n = top.n_atoms

    def peek():
        return toks[pos[0]] if pos[0] < len(toks) else None

    def take():
        t = peek()
        pos[0] += 1
        return t

    def values():
        out = []
        while peek() is not None and peek() not in _OPS and peek() not in _KEYWORDS:
            out.append(take())
        if not out:
            raise RuntimeError("selection keyword without a value in %r" % selection)
        return out

    def match_str(arr, vals):
        m = np.zeros(n, dtype=bool)
        uniq, inv = np.unique(arr.astype(str), return_inverse=True)
        for v in vals:
            hit = np.array([fnmatch.fnmatchcase(u, v) for u in uniq], dtype=bool)
            m |= hit[inv]
        return m

    def match_range(arr, vals):
        m = np.zeros(n, dtype=bool)
        for v in vals:
            sep = ":" if ":" in v else ("-" if "-" in v[1:] else None)
            if sep:
                k = v.index(sep, 1) if sep == "-" else v.index(sep)
                lo, hi = int(v[:k]), int(v[k + 1:])
                m |= (arr >= lo) & (arr <= hi)
            else:
                m |= arr == int(v)
        return m

    def factor():
        t = take()
        if t is None:
            raise RuntimeError("unexpected end of selection %r" % selection)
        if t == "not":
            return ~factor()
        if t == "(":
            m = expr()
            if take() != ")":
                raise RuntimeError("unbalanced parentheses in %r" % selection)
            return m
        if t == "all":
            return np.ones(n, dtype=bool)
        if t == "resname":
            return match_str(top.resnames, values())
        if t == "name":
            return match_str(top.names, values())
        if t == "segid":
            return match_str(top.segids, values())
        if t == "type":
            return match_str(top.types, values())
        if t in ("resid", "resnum"):
            return match_range(top.resids, values())
        if t == "index":
            return match_range(np.arange(n), values())
        if t == "bynum":
            return match_range(np.arange(1, n + 1), values())
        raise RuntimeError("selection token %r is outside the supported subset "
                           "(all, resname, name, segid, type, resid, index, bynum, not/and/or)" % t)

    def term():
        m = factor()
        while peek() == "and":
            take()
            m = m & factor()
        return m

    def expr():
        m = term()
        while peek() == "or":
            take()
            m = m | term()
        return m

    mask = expr()
    if peek() is not None:
        raise RuntimeError("trailing tokens in selection %r" % selection)
    return mask


# ---------------------------------------------------------------------------------------
# DCD
# ---------------------------------------------------------------------------------------
class DCDFile(object):
    """Memory-mapped CHARMM / NAMD DCD trajectory."""

    def __init__(self, path: str):
        self.path = path
        size = os.path.getsize(path)
        with open(path, "rb") as fh:
            head = fh.read(92)
            if len(head) < 92:
                raise RuntimeError("%s is too short for a DCD header" % path)
            if struct.unpack("<i", head[:4])[0] == 84:
                self.endian = "<"
            elif struct.unpack(">i", head[:4])[0] == 84:
                self.endian = ">"
            else:
                raise RuntimeError("%s: not a DCD file (first record is not 84 bytes)" % path)
            e = self.endian
            if head[4:8] != b"CORD":
                raise RuntimeError("%s: DCD magic 'CORD' missing" % path)
            icntrl = struct.unpack(e + "20i", head[8:88])
            self.n_frames_header = icntrl[0]
            self.istart, self.nsavc = icntrl[1], max(1, icntrl[2])
            self.n_fixed = icntrl[8]
            self.charmm_version = icntrl[19]
            if self.charmm_version != 0:
                self.delta = struct.unpack(e + "f", head[8 + 36:8 + 40])[0]
                self.has_unitcell = icntrl[10] != 0
                self.has_4dims = icntrl[11] != 0
            else:                               # X-PLOR: delta is a double, no unit cell
                self.delta = struct.unpack(e + "d", head[8 + 36:8 + 44])[0]
                self.has_unitcell = False
                self.has_4dims = False
            if self.n_fixed:
                raise RuntimeError("%s: DCD files with fixed atoms are not supported" % path)
            # title block
            blk = struct.unpack(e + "i", fh.read(4))[0]
            fh.seek(blk, 1)
            if struct.unpack(e + "i", fh.read(4))[0] != blk:
                raise RuntimeError("%s: corrupt title block" % path)
            four, natoms, four2 = struct.unpack(e + "3i", fh.read(12))
            if four != 4 or four2 != 4:
                raise RuntimeError("%s: corrupt atom-count block" % path)
            self.n_atoms = natoms
            self.header_bytes = fh.tell()
        fields = []
        if self.has_unitcell:
            fields += [("c0", e + "i4"), ("cell", e + "f8", (6,)), ("c1", e + "i4")]
        for ax in "xyz":
            fields += [(ax + "0", e + "i4"), (ax, e + "f4", (natoms,)), (ax + "1", e + "i4")]
        if self.has_4dims:
            fields += [("w0", e + "i4"), ("w", e + "f4", (natoms,)), ("w1", e + "i4")]
        self.frame_dtype = np.dtype(fields)
        self.frame_bytes = self.frame_dtype.itemsize
        self.n_frames = (size - self.header_bytes) // self.frame_bytes
        if self.n_frames <= 0:
            raise RuntimeError("%s holds no complete frame" % path)
        self._mm = np.memmap(path, dtype=self.frame_dtype, mode="r", offset=self.header_bytes,
                             shape=(self.n_frames,))
        self.dt = float(self.delta) * self.nsavc * AKMA_PS      # ps between stored frames
        # raw layout for the device-side interleave: float32 index of the X / Y / Z records in a frame
        self.frame_floats = self.frame_bytes // 4
        self.plane_offsets = tuple(self.frame_dtype.fields[ax][1] // 4 for ax in "xyz")
        self.foreign_endian = (self.endian == ">") != (sys.byteorder == "big")
        self._fd = None

    def read_raw(self, f0: int, f1: int, out: np.ndarray, threads: Optional[int] = None):
        """Frames f0..f1-1 exactly as they lie in the file into ``out`` (uint8 [>= f1-f0, frame_bytes],
        e.g. pinned memory): positional reads straight into the buffer, several threads (os.preadv
        releases the GIL), no per-atom work on the host."""
        if self._fd is None:
            self._fd = os.open(self.path, os.O_RDONLY)
        if threads is None:
            threads = int(os.environ.get("PBK_READ_THREADS", min(16, os.cpu_count() or 1)))
        nbytes = (f1 - f0) * self.frame_bytes
        flat = out.reshape(-1)[:nbytes]
        base = self.header_bytes + f0 * self.frame_bytes
        piece = max(1 << 20, -(-nbytes // max(1, threads)))
        jobs = [(o, min(piece, nbytes - o)) for o in range(0, nbytes, piece)]

        def rd(job):
            o, n = job
            got = 0
            while got < n:
                r = os.preadv(self._fd, [memoryview(flat[o + got:o + n])], base + o + got)
                if r <= 0:
                    raise RuntimeError("short read from %s" % self.path)
                got += r

        if len(jobs) == 1:
            rd(jobs[0])
        else:
            list(_pool(threads).map(rd, jobs))
        return out[:f1 - f0]

    # -- what MDAnalysis' DCDReader hands over ---------------------------------------------
    def dimensions(self, frames) -> np.ndarray:
        """float32 [f, 6] = A, B, C, alpha, beta, gamma (zeros without a unit cell)."""
        frames = np.asarray(frames)
        out = np.zeros((len(frames), 6), dtype=np.float32)
        if not self.has_unitcell:
            return out
        cell = np.asarray(self._mm["cell"][frames], dtype=np.float64)      # A gamma B beta alpha C
        uc = cell[:, [0, 2, 5, 4, 3, 1]].copy()
        ang = uc[:, 3:]
        cosines = (np.abs(ang) <= 1.0).all(axis=1)
        conv = 90.0 - np.arcsin(np.clip(ang, -1.0, 1.0)) * 90.0 / (np.pi / 2)
        uc[:, 3:] = np.where(cosines[:, None], conv, ang)
        out[:] = uc
        return out

    def times(self, frames) -> np.ndarray:
        frames = np.asarray(frames, dtype=np.float64)
        return (frames + self.istart / float(self.nsavc)) * self.dt

    def planes(self, f0: int, f1: int, out: Optional[np.ndarray] = None) -> np.ndarray:
        """float32 [f1-f0, 3, n_atoms]: the X, Y and Z records of a contiguous run of frames,
        native byte order, copied plane by plane (no per-atom work on the host)."""
        nf = f1 - f0
        if out is None:
            out = np.empty((nf, 3, self.n_atoms), dtype=np.float32)
        for k, ax in enumerate("xyz"):
            out[:nf, k, :] = self._mm[ax][f0:f1]          # astype to native order happens in the copy
        return out[:nf]

    def positions(self, frames, atoms: Optional[np.ndarray] = None) -> np.ndarray:
        """float32 [f, M, 3] in MDAnalysis' layout (host path; the pipeline uses `planes`)."""
        frames = np.asarray(frames)
        sel = slice(None) if atoms is None else atoms
        return np.stack([np.asarray(self._mm[ax][frames])[:, sel] for ax in "xyz"], axis=2).astype(np.float32)


def write_dcd(path: str, positions: np.ndarray, dimensions: Optional[np.ndarray] = None,
              dt_ps: float = 1.0, istart: int = 0, nsavc: int = 1, endian: str = "<"):
    """CHARMM-style DCD (version 24, unit cell stored as A, cos(gamma), B, cos(beta), cos(alpha), C)."""
    positions = np.asarray(positions, dtype=np.float32)
    F, M, _ = positions.shape
    e = endian
    icntrl = [0] * 20
    icntrl[0], icntrl[1], icntrl[2], icntrl[3] = F, istart, nsavc, F * nsavc
    icntrl[10] = 1 if dimensions is not None else 0
    icntrl[19] = 24
    head = bytearray(struct.pack(e + "20i", *icntrl))
    head[36:40] = struct.pack(e + "f", dt_ps / nsavc / AKMA_PS)
    title = b"REMARKS written by pybilt_b200.formats.write_dcd".ljust(80)
    with open(path, "wb") as fh:
        fh.write(struct.pack(e + "i", 84) + b"CORD" + bytes(head) + struct.pack(e + "i", 84))
        fh.write(struct.pack(e + "2i", 84, 1) + title + struct.pack(e + "i", 84))
        fh.write(struct.pack(e + "3i", 4, M, 4))
        rec = struct.pack(e + "i", 4 * M)
        for f in range(F):
            if dimensions is not None:
                A, B, C, al, be, ga = [float(v) for v in dimensions[f]]
                cell = [A, np.cos(np.radians(ga)), B, np.cos(np.radians(be)), np.cos(np.radians(al)), C]
                cell = [0.0 if abs(c) < 1e-15 else c for c in cell]
                fh.write(struct.pack(e + "i", 48) + struct.pack(e + "6d", *cell) + struct.pack(e + "i", 48))
            for k in range(3):
                fh.write(rec + positions[f, :, k].astype(e + "f4").tobytes() + rec)


class DCDChain(object):
    """One or several DCD files read as one trajectory (MDAnalysis ChainReader: frames in file
    order; the time of frame g is taken as ``(g + istart/nsavc) * dt`` of the first file)."""

    def __init__(self, paths: Sequence[str]):
        self.files = [DCDFile(p) for p in paths]
        n = {f.n_atoms for f in self.files}
        if len(n) != 1:
            raise RuntimeError("the trajectory files hold different numbers of atoms")
        self.n_atoms = self.files[0].n_atoms
        self.offsets = np.concatenate([[0], np.cumsum([f.n_frames for f in self.files])])
        self.n_frames = int(self.offsets[-1])
        self.dt = self.files[0].dt

    def _locate(self, g: int) -> Tuple[int, int]:
        k = int(np.searchsorted(self.offsets, g, side="right") - 1)
        return k, g - int(self.offsets[k])

    def dimensions(self, frames) -> np.ndarray:
        frames = np.asarray(frames)
        out = np.empty((len(frames), 6), dtype=np.float32)
        for i, g in enumerate(frames):
            k, l = self._locate(int(g))
            out[i] = self.files[k].dimensions([l])[0]
        return out

    def all_dimensions(self) -> np.ndarray:
        return np.concatenate([f.dimensions(np.arange(f.n_frames)) for f in self.files], axis=0)

    def times(self, frames) -> np.ndarray:
        return self.files[0].times(frames)

    def planes(self, f0: int, f1: int, out: Optional[np.ndarray] = None) -> np.ndarray:
        nf = f1 - f0
        if out is None:
            out = np.empty((nf, 3, self.n_atoms), dtype=np.float32)
        g = f0
        while g < f1:
            k, l = self._locate(g)
            take = min(f1 - g, self.files[k].n_frames - l)
            self.files[k].planes(l, l + take, out[g - f0:g - f0 + take])
            g += take
        return out[:nf]

    @property
    def raw_layout(self):
        """(frame_bytes, frame_floats, plane_offsets, foreign_endian) when every file shares one
        frame layout (then batches can be shipped to the device as raw file bytes), else None."""
        f0 = self.files[0]
        key = (f0.frame_bytes, f0.frame_floats, f0.plane_offsets, f0.foreign_endian)
        for f in self.files[1:]:
            if (f.frame_bytes, f.frame_floats, f.plane_offsets, f.foreign_endian) != key:
                return None
        return key

    def read_raw(self, f0: int, f1: int, out: np.ndarray):
        g = f0
        while g < f1:
            k, l = self._locate(g)
            take = min(f1 - g, self.files[k].n_frames - l)
            self.files[k].read_raw(l, l + take, out[g - f0:g - f0 + take])
            g += take
        return out[:f1 - f0]

    def positions(self, frames, atoms=None) -> np.ndarray:
        frames = np.asarray(frames)
        out = []
        for g in frames:
            k, l = self._locate(int(g))
            out.append(self.files[k].positions([l], atoms)[0])
        return np.stack(out, axis=0) if out else np.empty((0, 0, 3), dtype=np.float32)
