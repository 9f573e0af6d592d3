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

The analyses beyond the MSD run their pair searches on the device
(``pbk_ct_nearest`` / ``pbk_ct_adjacency``, csrc/pbk_comtraj.cu) over all frames at once and finish on
the host with the reference's own NumPy / Welford arithmetic: ``calc_membrane_thickness`` (:1244),
``check_clustering`` (:1411), ``calc_area_per_lipid_closest_neighbor_circle`` (:1655),
``calc_area_per_lipid_box`` (:1768), ``voronoi_tesselate`` / ``delaunay_tesselate`` (:1843, :1888, SciPy on
the host like the reference), ``step_vector`` / ``step_vector_frames`` (:1929, :1990) and
``remove_leaflet_com_motion`` (:2081).  Colour tables (``step_vector_colors``, the colours of
``export_clusters_for_plotting``) need matplotlib and are not rebuilt.
"""
from __future__ import annotations

import numpy as np
import torch

from .bilayer_analyzer import TrajectoryData
from .engine import LipidReductionEngine, build_bead_layout
from .running_stats import RunningStats


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
        torch.cuda.set_device(device)             # libpbk launches on the current device
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


    # ------------------------------------------------------------------ analyses beyond the MSD (N4)
    def _do_leaflets(self, leaflet):
        if leaflet == "both":
            return ['upper', 'lower']
        if leaflet in ("upper", "lower"):
            return [leaflet]
        print("!! Warning - request for unknown leaflet name \'", leaflet)
        print("!! the options are \"upper\", \"lower\", or \"both\"--using the default \"both\"")
        return []

    def calc_membrane_thickness(self):
        """(zavgs [F,5], zmaps [F,N,6]): thickness from every lipid's laterally nearest partner in the
        other leaflet (COMTraj.py:1244-1408); upper members first, then lower, in member order."""
        F, N = self.nframes, self.nlipids
        h = self._host()
        com = h['com']
        xi, yi, zi = self.plane[0], self.plane[1], self.norm
        up = np.asarray(self.leaflets['upper'].get_member_indices(), dtype=np.int64)
        lo = np.asarray(self.leaflets['lower'].get_member_indices(), dtype=np.int64)
        zdists = np.zeros((F, N, 1))
        zmaps = np.zeros((F, N, 6))
        n0 = 0
        for mine, other, mine_is_upper in ((up, lo, True), (lo, up, False)):
            if len(mine) == 0:
                continue
            pos = self._engine.ct_nearest(mine, other, self.plane)[0].cpu().numpy() if len(other) else \
                np.full((F, len(mine)), -1, dtype=np.int32)
            found = pos >= 0
            partner = other[np.where(found, pos, 0)] if len(other) else np.zeros_like(pos, dtype=np.int64)
            fidx = np.arange(F)[:, None]
            cm, cp = com[:, mine, :], com[fidx, partner, :]
            comcup, comclo = (cm, cp) if mine_is_upper else (cp, cm)
            comavg = (comcup + comclo) / 2.0
            distz = np.absolute((cm - cp)[:, :, zi])
            sl = slice(n0, n0 + len(mine))
            zdists[:, sl, 0] = np.where(found, distz, 0.0)
            zmaps[:, sl, 0] = self._times[:, None]
            zmaps[:, sl, 1] = np.where(found, comavg[:, :, xi], 0.0)
            zmaps[:, sl, 2] = np.where(found, comavg[:, :, yi], 0.0)
            zmaps[:, sl, 3] = np.where(found, comclo[:, :, zi], 0.0)
            zmaps[:, sl, 4] = np.where(found, comcup[:, :, zi], 0.0)
            zmaps[:, sl, 5] = zdists[:, sl, 0]
            n0 += len(mine)
        zavgs = np.zeros((F, 5))
        stat = RunningStats()
        for f in range(F):
            curr = zdists[f, :]
            zavgcurr = curr.mean()
            stat.push(zavgcurr)
            zavgs[f] = (self._times[f], zavgcurr, curr.std(), stat.mean(), stat.deviation())
        return (zavgs, zmaps)

    def check_clustering(self, leaflet="both", group="all", dist=15.0):
        """[F,5] running statistics of the clusters grown with the lateral cut-off `dist`
        (COMTraj.py:1411-1576); the cluster lists of every frame are kept in ``self.clusters``."""
        indices = self._indices(leaflet, group)
        n, F = len(indices), self.nframes
        bits = self._engine.ct_adjacency(indices, dist, self.plane).cpu().numpy().view(np.uint32) if n else None
        self.clusters = []
        outdata = np.zeros((F, 5))
        ncstat, asstat = RunningStats(), RunningStats()
        for f in range(F):
            adj = np.unpackbits(bits[f].view(np.uint8), axis=1, bitorder='little')[:, :n].astype(bool) if n else None
            clusters = []
            master = list(range(n))                 # positions in `indices`, not yet in a cluster
            while master:
                neighborlist = [master[0]]
                taken = {master[0]}
                i = 0
                while i < len(neighborlist):
                    row = adj[neighborlist[i]]
                    for j in master:
                        if j not in taken and row[j]:
                            neighborlist.append(j)
                            taken.add(j)
                    i += 1
                master = [j for j in master if j not in taken]
                if len(neighborlist) > 1:
                    clusters.append([indices[j] for j in neighborlist])
            clsizestat = RunningStats()
            for cluster in clusters:
                clsizestat.push(len(cluster))
            ncstat.push(len(clusters))
            asstat.push(clsizestat.mean())
            outdata[f] = (self._times[f], ncstat.mean(), ncstat.deviation(), asstat.mean(), asstat.deviation())
            self.clusters.append(list(clusters))
        return outdata

    def export_clusters_for_plotting(self):
        """Per frame [x, y, cluster number] of the clustered lipids (COMTraj.py:1578-1652; the reference
        maps the cluster number through matplotlib's rainbow colour table)."""
        if len(self.clusters) == 0:
            print("Warning!! - call to \'ExportClustersForPlotting\' of a MemSys object with no cluster lists")
            print("      the \'CheckClustering\' function needs to be called first!")
            return
        com = self._host()['com']
        xi, yi = self.plane
        output = []
        for f, frame_clusters in enumerate(self.clusters):
            xcoord, ycoord, number = [], [], []
            for c, cluster in enumerate(frame_clusters):
                for index in cluster:
                    xcoord.append(com[f, index, xi])
                    ycoord.append(com[f, index, yi])
                    number.append(c)
            output.append([xcoord, ycoord, number])
        return output

    def calc_area_per_lipid_closest_neighbor_circle(self, leaflet="both", group="all"):
        """[F,4] (time, configurational mean area, running mean, running deviation): area of the
        non-overlapping circles around a lipid and its nearest leaflet neighbour (COMTraj.py:1655-1766)."""
        F = self.nframes
        sub_fact = (2.0 * np.pi / 3.0 - np.sqrt(3.0) / 2.0)
        rmins = []
        for name in self._do_leaflets(leaflet):
            idx = self.leaflets[name].get_group_indices(group)
            allm = self.leaflets[name].get_member_indices()
            if len(idx):
                rmins.append(self._engine.ct_nearest(idx, allm, self.plane, exclude_self=True)[1].cpu().numpy())
        areas = np.zeros((F, 4))
        stat = RunningStats()
        if rmins:
            r = np.concatenate(rmins, axis=1)
            a = np.pi * r * r - (r * r) * sub_fact
            mean = a[:, 0].copy()                     # RunningStats.push over the lipids, all frames side by side
            for i in range(1, a.shape[1]):
                mean = mean + (a[:, i] - mean) / np.float64(i + 1)
        else:
            mean = np.zeros(F)
        for f in range(F):
            stat.push(mean[f])
            areas[f] = (self._times[f], mean[f], stat.mean(), stat.deviation())
        return areas

    def calc_area_per_lipid_box(self, leaflet="both"):
        """[F,4]: lateral box area over the number of lipids per leaflet (COMTraj.py:1768-1840)."""
        do = self._do_leaflets(leaflet)
        nlip = 0
        if leaflet == "both":
            nlip = [float(len(self.leaflets[name])) for name in do]
        elif do:
            nlip = len(self.leaflets[leaflet])
        xi, yi = self.plane
        areas = np.zeros((self.nframes, 4))
        stat = RunningStats()
        for f in range(self.nframes):
            boxc = self._dims[f, :3]
            lat_area = (boxc[xi] / 2.0) * (boxc[yi] / 2.0) * 4.0
            area_per_lip = lat_area / nlip
            if leaflet == 'both':
                area_per_lip = (lat_area / 2.0) * ((nlip[0] + nlip[1]) / (nlip[0] * nlip[1]))
            stat.push(area_per_lip)
            areas[f] = (self._times[f], area_per_lip, stat.mean(), stat.deviation())
        return areas

    def _plane_points(self, leaflet, group):
        idx = self._indices(leaflet, group)
        return self._host()['com_unwrap'][:, idx][:, :, self.plane]

    def voronoi_tesselate(self, leaflet="both", group="all"):
        """One scipy.spatial.Voronoi per frame over the unwrapped lateral COMs (COMTraj.py:1843-1885)."""
        from scipy.spatial import Voronoi
        return [Voronoi(np.array(p)) for p in self._plane_points(leaflet, group)]

    def delaunay_tesselate(self, leaflet="both", group="all"):
        """[x, y, scipy.spatial.Delaunay] per frame (COMTraj.py:1888-1925)."""
        from scipy.spatial import Delaunay
        out = []
        for p in self._plane_points(leaflet, group):
            p = np.array(p)
            out.append([p[:, 0], p[:, 1], Delaunay(p)])
        return out

    def step_vector(self, leaflet="both", group="all", fstart=0, fend=-1, fstep=1000, wrapped=False):
        """List of [n,4] arrays (start x, start y, dx, dy) of the unwrapped lateral displacement between
        frames f - fstep and f (COMTraj.py:1929-1987)."""
        if fstart < 0:
            fstart += self.nframes
        if fend < 0:
            fend += self.nframes
        if fstep == 'single':
            fstep = fend - fstart
        idx = self._indices(leaflet, group)
        h = self._host()
        out = []
        for f in range(fstart + fstep, fend + 1, fstep):
            cur = h['com_unwrap'][f][idx][:, self.plane]
            prev = h['com_unwrap'][f - fstep][idx][:, self.plane]
            start = h['com'][f - fstep][idx][:, self.plane] if wrapped else prev
            vec = np.zeros((len(idx), 4))
            vec[:, 0:2] = start
            vec[:, 2:4] = cur - prev
            out.append(vec)
        return out

    def step_vector_frames(self, fstart=0, fend=-1, fstep=1000):
        if fstart < 0:
            fstart += self.nframes
        if fend < 0:
            fend += self.nframes
        if fstep == 'single':
            fstep = fend - fstart
        return np.array([[f - fstep, f] for f in range(fstart + fstep, fend + 1, fstep)], dtype=int)

    def remove_leaflet_com_motion(self, leaflet="both"):
        """Subtracts every leaflet's own (mass-weighted, unwrapped) COM from its members in every frame,
        keeping component 2 (COMTraj.py:2081-2131: ``lcom[2] = 0.0``); later analyses see the result."""
        do = self._do_leaflets(leaflet) or ['upper', 'lower']
        h = self._host()
        cu = h['com_unwrap']
        mass = np.asarray(self._layout.bead_mass, dtype=np.float64)
        for name in do:
            idx = self.leaflets[name].get_member_indices()
            if not idx:
                continue
            lcom = np.zeros((self.nframes, 3))
            masst = 0.0
            for i in idx:                              # the reference's accumulation order, all frames side by side
                lcom += cu[:, i, :] * mass[i]
                masst += mass[i]
            lcom /= masst
            lcom[:, 2] = 0.0
            cu[:, idx, :] -= lcom[:, None, :]
        self._engine.com_unwrap[:self.nframes].copy_(torch.from_numpy(cu).to(self._engine.dev))
        return


COMTraj.calc_membrane_thickness_parallel = lambda self, nprocs=2: self.calc_membrane_thickness()
