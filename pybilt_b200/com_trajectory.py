"""``COMTraj``: the stand-alone centre-of-mass trajectory of PyBILT on the device.

Mirror of the core of pybilt/com_trajectory/COMTraj.py: ``LipidCOM`` (:46), ``Frame`` (:105),
``Leaflet`` / ``LipidGroup`` (:442, :597) and ``COMTraj`` (:907) with ``calc_msd`` (:1142) and
``calc_msd_parallel`` (:2139).  The reference builds every frame in Python and keeps them in
a ``shelve`` database; here the per-lipid COMs of all frames are computed by libpbk and
``COMTraj.frame[f]`` is a host view.

COMTraj's arithmetic differs from BilayerAnalyzer's in three documented ways (SURVEY.md §8a
row T1), all reproduced: the scalar unwrap of ``pybilt/mda_tools/mda_unwrap.py:12-85``
evaluates ``dz + shift*C`` in float32; the membrane COM is ``mem_sel.center_of_mass()``
(:1084, rows accumulated in atom order instead of NumPy's pairwise order); leaflets are
assigned once, from frame 0 (:1099-1115).

Not provided: thickness, clustering, Voronoi/Delaunay, step vectors.
"""
from __future__ import annotations

import numpy as np
import torch

from .bilayer_analyzer import TrajectoryData
from .engine import LipidReductionEngine, build_bead_layout


class LipidCOM(object):
    def __init__(self):
        self.type = "UNK"
        self.com = np.zeros(3)
        self.com_unwrap = np.zeros(3)
        self.mass = 1.0


class Frame(object):
    def __init__(self, nlipids):
        self.lipidcom = [LipidCOM() for _ in range(nlipids)]
        self.box = np.zeros(3)
        self.time = np.zeros(1)
        self.number = np.zeros(1, dtype=int)
        self.mdnumber = np.zeros(1, dtype=int)

    def set_box(self, box_lengths):
        self.box = box_lengths

    def set_time(self, time):
        self.time = time

    def __len__(self):
        return len(self.lipidcom)

    def com(self, wrapped=True):
        out, total = np.zeros(3), 0.0
        for lipid in self.lipidcom:
            out += (lipid.com if wrapped else lipid.com_unwrap) * lipid.mass
            total += lipid.mass
        return out / total


class LipidGroup(object):
    def __init__(self, name):
        self.lg_members = []
        self.lg_name = name

    def add_member(self, new_mem):
        self.lg_members.append(new_mem)

    def name(self):
        return self.lg_name


class Leaflet(object):
    """COMTraj.py:442-595 -- note get_group_indices('all') concatenates group by group."""

    def __init__(self, name):
        self.name = name
        self.members = []
        self.groups = []
        self.group_dict = {}

    def __repr__(self):
        return '%s leaflet of a Center of mass trajectory with %s members and %s lipid groups' % (
            self.name, len(self.members), len(self.groups))

    __str__ = __repr__

    def __len__(self):
        return len(self.members)

    def add_member(self, index, resname):
        self.members.append([index, resname])
        if resname not in self.group_dict:
            self.group_dict[resname] = len(self.groups)
            self.groups.append(LipidGroup(resname))
        self.groups[self.group_dict[resname]].add_member(index)

    def get_group_indices(self, group_name):
        if group_name != "all" and group_name in self.group_dict:
            return list(self.groups[self.group_dict[group_name]].lg_members)
        if group_name != "all":
            print("!! Warning - request for unknown Lipid Group \'", group_name, "\' from the ",
                  self.name, " leaflet")
            print("!! using the default \"all\"")
        indices = []
        for element in self.group_dict:
            indices += self.groups[self.group_dict[element]].lg_members
        return list(indices)

    def get_member_indices(self):
        return [m[0] for m in self.members]

    def has_group(self, group_name):
        return group_name in self.group_dict

    def num_groups(self):
        return len(self.groups)

    def get_group_names(self):
        return [g.lg_name for g in self.groups]


class _FrameViews(object):
    """``COMTraj.frame``: indexable like the reference's FrameShelve (:241-379)."""

    def __init__(self, owner):
        self._o = owner

    def __len__(self):
        return self._o.nframes

    def __getitem__(self, f):
        o = self._o
        if f < 0:
            f += o.nframes
        h = o._host()
        fr = Frame(o.nlipids)
        for i, lc in enumerate(fr.lipidcom):
            lc.type = o._layout.types[i]
            lc.com = h['com'][f, i].copy()
            lc.com_unwrap = h['com_unwrap'][f, i].copy()
            lc.mass = o._layout.bead_mass[i]
        fr.set_box(o._dims[f, :3])
        fr.set_time(o._times[f])
        fr.number = f
        fr.mdnumber = int(o._frames[f])
        return fr


class COMTraj(object):
    def __init__(self, mda_traj, mem_sel, plane="xy", fstart=0, fend=-1, fskip=1, frame_path='Default',
                 frame_save=False, nprocs=1, device=0):
        """``mda_traj`` / ``mem_sel``: an MDAnalysis trajectory and AtomGroup, or -- in memory --
        a trajectory (positions f32 [F,M,3], dimensions f32 [F,6], times) and a topology
        (masses, names, resindex, resnames, resids)."""
        ii, jj, kk = 0, 1, 2
        if plane in ("yz", "zy"):
            ii, jj, kk = 1, 2, 0
        if plane in ("xz", "zx"):
            ii, jj, kk = 0, 2, 1
        self.plane = [ii, jj]
        self.norm = kk
        if hasattr(mem_sel, 'universe'):          # MDAnalysis objects
            raise RuntimeError("MDAnalysis input: build a BilayerAnalyzer-style TrajectoryData first "
                               "(MDAnalysis is not installed in this image)")
        md = TrajectoryData(mem_sel, mda_traj, "all")
        n = md.nframes
        while fstart < 0:
            fstart += n
        while fend < 0:
            fend += n
        self._frames = np.arange(fstart, fend + 1, fskip)
        self.nframes = len(self._frames)
        self._layout = build_bead_layout(md.masses, md.names, md.resindex, md.resnames, md.resids)
        self.nlipids = self._layout.n_beads
        eng = LipidReductionEngine(md.masses, self._layout, self.nframes, device=device, norm=self.norm,
                                   leaflet_update_interval=1, comtraj_rules=True)
        x, d, t = md.frames(self._frames)
        # zero box of a later frame -> previous frame's box (COMTraj.py:1031-1033)
        d = np.array(d, dtype=np.float32)
        for f in range(1, len(d)):
            if (d[f, :3] == 0).any():
                d[f, :3] = d[f - 1, :3]
        xh = x if torch.is_tensor(x) else torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
        eng.push(xh.to(eng.dev), torch.from_numpy(np.ascontiguousarray(d[:, :3])).to(eng.dev))
        eng.check_status()
        self._engine, self._dims, self._times = eng, d, np.asarray(t, dtype=np.float64)
        self._host_cache = None
        # leaflets from frame 0 only (COMTraj.py:1099-1115)
        lab0 = eng.labels[0].cpu().numpy()
        self.leaflets = {'upper': Leaflet('upper'), 'lower': Leaflet('lower')}
        self.com_leaflet = []
        for i in range(self.nlipids):
            pos = 'upper' if lab0[i] == 1 else 'lower'
            self.com_leaflet.append(pos)
            self.leaflets[pos].add_member(i, self._layout.types[i])
        self.clusters = []
        self.frame = _FrameViews(self)

    def _host(self):
        if self._host_cache is None:
            e = self._engine
            self._host_cache = {'com': e.com.cpu().numpy(), 'com_unwrap': e.com_unwrap.cpu().numpy()}
        return self._host_cache

    def __repr__(self):
        return 'Center of mass trajectory with %s frames and %s lipids/components' % (self.nframes,
                                                                                   self.nlipids)

    __str__ = __repr__

    def number_of_unique_groups(self):
        names = []
        for leaflet in self.leaflets.values():
            for group in leaflet.groups:
                if group.name() not in names:
                    names.append(group.name())
        return len(names)

    def _indices(self, leaflet, group):
        if leaflet in ("upper", "lower"):
            return self.leaflets[leaflet].get_group_indices(group)
        if leaflet != "both":
            print("!! Warning - request for unknown leaflet name \'", leaflet)
            print("!! the options are \"upper\", \"lower\", or \"both\"--using the default \"both\"")
        out = []
        for name in self.leaflets:
            out += self.leaflets[name].get_group_indices(group)
        return out

    def calc_msd(self, leaflet="both", group="all"):
        """[nframes, 2] (time, lateral MSD vs frame 0); row 0 is (0, 0)  (COMTraj.py:1142-1240)."""
        idx = np.asarray(self._indices(leaflet, group), dtype=np.int32)
        F = self.nframes
        vals = self._engine.msd_pairs(idx, np.zeros(F, dtype=np.int32), np.arange(F, dtype=np.int32),
                                      self.plane).cpu().numpy()
        out = np.zeros((F, 2))
        out[1:, 0] = self._times[1:]
        out[1:, 1] = vals[1:]
        return out

    def calc_msd_parallel(self, leaflet="both", group="all", nprocs=2):
        """The reference shards frame ranges over ``mp.Pool`` (COMTraj.py:2139-2232); the device
        already processes all frames concurrently."""
        return self.calc_msd(leaflet=leaflet, group=group)
