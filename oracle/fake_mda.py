"""In-memory stand-in for the parts of MDAnalysis that PyBILT's hot path touches.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  MDAnalysis is a third-party
dependency of the reference (README.md:36,45 pin "MDAnalysis > 1.0 / tested 1.1.1";
there is no lockfile) and is not installed in this image, so the literal reference
can only be executed against a look-alike.  This module *defines* the MDAnalysis 1.x
semantics the oracle assumes:

* ``Timestep.positions`` is the float32 array ``_pos`` itself (writes to ``_pos`` are
  visible through ``positions``), ``dimensions`` is float32 ``[a,b,c,alpha,beta,gamma]``;
* every frame yielded by the trajectory is freshly loaded (``COMFrame`` overwrites
  ``_pos``: /root/reference/pybilt/bilayer_analyzer/com_frame.py:172-180);
* ``AtomGroup.center_of_mass()`` is MDAnalysis ``GroupBase.center``:
  ``(positions.astype(f64) * masses[:, None]).sum(axis=0) / masses.sum()``;
* ``Residue.atoms`` are the residue's atoms in topology order, ``select_atoms('name X')``
  keeps topology order, ``a + b`` concatenates preserving order;
* ``AtomGroup.residues`` are in ascending residue index.

Call sites this serves: com_frame.py:57-95,153-186; bilayer_analyzer.py:921-943;
mda_data.py:16-27.
"""
from __future__ import annotations

import numpy as np


class Timestep:
    def __init__(self, pos, dims, time, frame):
        self._pos = pos
        self.dimensions = dims
        self.time = time
        self.frame = frame

    @property
    def positions(self):
        return self._pos

    @positions.setter
    def positions(self, value):
        self._pos[:] = value


class _Trajectory:
    def __init__(self, universe, X, D, times):
        self._u = universe
        self._X = X
        self._D = D
        self._t = times
        self.ts = self._load(0)

    def _load(self, f):
        self.ts = Timestep(np.array(self._X[f], dtype=np.float32, copy=True),
                           np.array(self._D[f], dtype=np.float32, copy=True),
                           float(self._t[f]), int(f))
        return self.ts

    def __len__(self):
        return int(self._X.shape[0])

    @property
    def time(self):
        return self.ts.time

    def __iter__(self):
        for f in range(len(self)):
            yield self._load(f)

    def __getitem__(self, item):
        if isinstance(item, slice):
            rng = range(*item.indices(len(self)))

            def gen():
                for f in rng:
                    yield self._load(f)
            return gen()
        f = int(item)
        if f < 0:
            f += len(self)
        return self._load(f)


class Residue:
    def __init__(self, universe, resindex):
        self._u = universe
        self.resindex = int(resindex)
        self.resname = universe._resnames[resindex]
        self.resid = int(universe._resids[resindex])

    @property
    def atoms(self):
        lo, hi = self._u._seg[self.resindex], self._u._seg[self.resindex + 1]
        return AtomGroup(self._u, np.arange(lo, hi, dtype=np.int64))

    def center_of_mass(self):
        return self.atoms.center_of_mass()

    def total_mass(self):
        return self.atoms.total_mass()


class AtomGroup:
    def __init__(self, universe, indices):
        self.universe = universe
        self._ix = np.asarray(indices, dtype=np.int64)

    @property
    def indices(self):
        return self._ix

    @property
    def atoms(self):
        return self

    @property
    def masses(self):
        return self.universe._masses[self._ix]

    @property
    def names(self):
        return self.universe._names[self._ix]

    @property
    def positions(self):
        return self.universe.trajectory.ts._pos[self._ix]

    @property
    def residues(self):
        ridx = np.unique(self.universe._resindex[self._ix])
        return [Residue(self.universe, r) for r in ridx]

    @property
    def n_residues(self):
        return len(np.unique(self.universe._resindex[self._ix]))

    def __len__(self):
        return int(self._ix.shape[0])

    def __add__(self, other):
        return AtomGroup(self.universe, np.concatenate([self._ix, other._ix]))

    def select_atoms(self, sel):
        if " and " in sel:
            # "resname X and resid R and name A" (bilayer_analyzer.py:857-861): the intersection
            out = self
            for part in sel.split(" and "):
                out = out.select_atoms(part.strip())
            return out
        words = sel.split()
        if words[0] == "resid":
            rid = np.asarray(self.universe._resids)[self.universe._resindex[self._ix]]
            keep = np.isin(rid, [int(w) for w in words[1:]])
            return AtomGroup(self.universe, self._ix[keep])
        if words[0] == "all":
            return AtomGroup(self.universe, self._ix.copy())
        if words[0] == "name":
            keep = np.isin(self.universe._names[self._ix], words[1:])
            return AtomGroup(self.universe, self._ix[keep])
        if words[0] == "resname":
            rn = np.asarray(self.universe._resnames, dtype=object)[self.universe._resindex[self._ix]]
            keep = np.isin(rn, words[1:])
            return AtomGroup(self.universe, self._ix[keep])
        raise NotImplementedError("fake MDAnalysis selection: " + sel)

    def center_of_mass(self):
        coords = self.positions.astype(np.float64, copy=False)
        weights = self.masses.astype(np.float64, copy=False)
        return (coords * weights[:, None]).sum(axis=0) / weights.sum()

    def total_mass(self):
        return self.masses.sum()


class Universe:
    """``Universe(top, traj)`` with ``top`` a pybilt_b200.synthetic.Topology-like object
    (masses, names, resindex, resnames, resids) and ``traj = (X, D, times)``."""

    def __init__(self, top, traj):
        self._masses = np.asarray(top.masses, dtype=np.float64)
        self._names = np.asarray(top.names, dtype=object)
        self._resindex = np.asarray(top.resindex, dtype=np.int64)
        self._resnames = list(top.resnames)
        self._resids = np.asarray(top.resids)
        counts = np.bincount(self._resindex, minlength=len(self._resnames))
        self._seg = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        X, D, times = traj
        self.trajectory = _Trajectory(self, X, D, times)
        self.atoms = AtomGroup(self, np.arange(self._masses.shape[0], dtype=np.int64))

    def select_atoms(self, sel):
        return self.atoms.select_atoms(sel)

    def load_new(self, traj):
        X, D, times = traj
        self.trajectory = _Trajectory(self, X, D, times)
