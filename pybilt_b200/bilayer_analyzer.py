"""``BilayerAnalyzer``: PyBILT's analysis driver with the frame loop moved onto the B200.

Drop-in mirror of pybilt/bilayer_analyzer/bilayer_analyzer.py:112-1079: same constructor
(``structure`` / ``trajectory`` / ``selection``, ``input_file`` or ``input_dict``), same
``settings`` / ``rep_settings`` / ``reps`` dictionaries, ``add_analysis`` /
``remove_analysis`` / ``get_analysis_data`` / ``dump_data`` / ``set_frame_range`` /
``reset`` / ``run_analysis`` and the iterator protocol.

What differs: ``__next__`` in the reference (:901-1031) unwraps, builds a COMFrame, assigns
leaflets, optionally builds lipid grids and calls every plugin -- all in Python, one frame
at a time.  Here all analysed frames are pushed through ``LipidReductionEngine`` (libpbk) in
batches, the representations stay in HBM, the batched protocols run their kernels over all
frames at once, and ``reps[...]`` are host views materialised per frame only for iterator
users and for plain per-frame plugins.

Trajectory input: an MDAnalysis Universe is built when ``structure`` / ``trajectory`` are
file names (MDAnalysis must be importable); in-memory input is accepted as
``structure=<topology>`` (attributes masses, names, resindex, resnames, resids) and
``trajectory=<trajectory>`` (attributes positions f32 [F,M,3], dimensions f32 [F,6], times)
or the tuple ``(positions, dimensions, times)``.
"""
from __future__ import annotations

import ast
import os
import pickle

import numpy as np
import torch

from . import analysis_protocols as ap
from . import com_frame as cf
from . import leaflet as lf
from . import lipid_grid as lg
from . import distributed as pd
from . import formats as fmt
from . import plot_protocols as pp
from .engine import FusedPathRejected, LipidReductionEngine, build_bead_layout

default_analysis_commands = ['msd msd_1']


class _HostFrame(object):
    """What plugins see as ``reps['current_mda_frame']``: dimensions, time, frame, positions."""

    def __init__(self, positions, dimensions, time, frame):
        self._pos = positions
        self.dimensions = dimensions
        self.time = time
        self.frame = frame

    @property
    def positions(self):
        return self._pos


class FrameWindow(object):
    """Positions of a contiguous window of a longer trajectory: ``window[g0:g1]`` with GLOBAL
    frame indices returns the rows held locally.  Lets a frame shard keep only its own frames
    (+ halos) in host memory while box / time arrays describe the whole run."""

    def __init__(self, positions_local, first_global, n_global):
        self.local = positions_local
        self.first = int(first_global)
        self.n_global = int(n_global)

    def __getitem__(self, item):
        if not isinstance(item, slice):
            raise TypeError("FrameWindow supports contiguous slices of global frame indices")
        lo, hi, step = item.indices(self.n_global)
        if step != 1 or lo < self.first or hi > self.first + self.local.shape[0]:
            raise IndexError("frames %d:%d are outside this rank's window" % (lo, hi))
        return self.local[lo - self.first:hi - self.first]


class TrajectoryData(object):
    """Equivalent of pybilt/bilayer_analyzer/mda_data.py:9-38 (``MDAData``): topology of the
    selected atoms plus frame access."""

    def __init__(self, structure, trajectory, sel_string):
        self.psf_file = structure if isinstance(structure, str) else None
        self.trajectory_file = trajectory if isinstance(trajectory, (str, list)) else None
        self.bilayer_sel_string = sel_string
        self.mda_universe = None
        self.dcd = None                 # pybilt_b200.formats.DCDChain when read natively
        if isinstance(structure, str):
            traj_list = [trajectory] if isinstance(trajectory, str) else list(trajectory)
            native = (structure.lower().endswith(".psf") and
                      all(isinstance(t, str) and t.lower().endswith(".dcd") for t in traj_list))
            if native and os.environ.get("PBK_USE_MDANALYSIS") is None:
                self._from_psf_dcd(structure, traj_list, sel_string)
            else:
                self._from_mdanalysis(structure, trajectory, sel_string)
        else:
            self._from_arrays(structure, trajectory, sel_string)
        self.natoms = int(self.masses.shape[0])
        self.n_residues = len(self.resnames)

    # -- in-memory -------------------------------------------------------------------
    def _from_arrays(self, top, traj, sel):
        if isinstance(traj, (tuple, list)):
            positions, dimensions, times = traj
        else:
            positions, dimensions, times = traj.positions, traj.dimensions, traj.times
        resnames_all = list(top.resnames)
        resindex = np.asarray(top.resindex, dtype=np.int64)
        words = (sel or "all").split()
        if words[0] == "all":
            keep_res = np.ones(len(resnames_all), dtype=bool)
        elif words[0] == "resname":
            keep_res = np.isin(np.asarray(resnames_all, dtype=object), words[1:])
        else:
            raise RuntimeError("in-memory input supports the selections 'all' and "
                               "'resname A B ...' (got %r)" % sel)
        atom_keep = keep_res[resindex]
        self.indices = np.nonzero(atom_keep)[0]
        self.masses = np.asarray(top.masses, dtype=np.float64)[self.indices]
        self.names = np.asarray(top.names, dtype=object)[self.indices]
        new_index = np.cumsum(keep_res) - 1
        self.resindex = new_index[resindex[self.indices]]
        self.resnames = [r for r, k in zip(resnames_all, keep_res) if k]
        self.resids = np.asarray(top.resids)[keep_res]
        self._positions = positions
        self._dimensions = np.asarray(dimensions, dtype=np.float32)
        self._times = np.asarray(times, dtype=np.float64)
        self.nframes = int(self._dimensions.shape[0])
        self._all_atoms = bool(atom_keep.all())

    # -- psf + dcd, read natively (pybilt_b200/formats.py) ------------------------------------
    def _from_psf_dcd(self, structure, traj_list, sel):
        top = fmt.PSFTopology(structure)
        self.dcd = fmt.DCDChain(traj_list)
        if self.dcd.n_atoms != top.n_atoms:
            raise RuntimeError("structure (%d atoms) and trajectory (%d atoms) do not match"
                               % (top.n_atoms, self.dcd.n_atoms))
        mask = fmt.select_atoms(top, sel or "all")
        if not mask.any():
            raise RuntimeError("the selection %r matches no atom" % sel)
        # like group.residues.atoms == group: a residue is selected as a whole or not at all
        res_any = np.zeros(int(top.resindex[-1]) + 1, dtype=bool)
        res_any[top.resindex[mask]] = True
        if (res_any[top.resindex] != mask).any():
            raise RuntimeError("the selection must cover whole residues")
        self.indices = np.nonzero(mask)[0]
        self.masses = top.masses[self.indices]
        self.names = top.names[self.indices]
        ridx = top.resindex[self.indices]
        first = np.concatenate([[True], ridx[1:] != ridx[:-1]])
        self.resindex = np.cumsum(first) - 1
        self.resnames = list(top.resnames[self.indices][first])
        self.resids = top.resids[self.indices][first]
        self.n_atoms_total = top.n_atoms
        self._all_atoms = bool(mask.all())
        self._dimensions = self.dcd.all_dimensions()
        self._times = self.dcd.times(np.arange(self.dcd.n_frames))
        self.nframes = self.dcd.n_frames

    def read_raw_frames(self, frame_idx, out):
        """The frames exactly as they lie in the file(s) into ``out`` (uint8 [>= f, frame_bytes],
        pinned); runs of consecutive frames are read in one go."""
        frame_idx = np.asarray(frame_idx)
        k = 0
        while k < len(frame_idx):
            j = k
            while j + 1 < len(frame_idx) and frame_idx[j + 1] == frame_idx[j] + 1:
                j += 1
            self.dcd.read_raw(int(frame_idx[k]), int(frame_idx[j]) + 1, out[k:j + 1])
            k = j + 1
        return out[:len(frame_idx)]

    def frames(self, frame_idx):
        """float32 positions [f, M, 3] (selected atoms), dimensions [f, 6], times [f]."""
        if self.mda_universe is not None:
            return self._read_mda(frame_idx)
        if self.dcd is not None:
            frame_idx = np.asarray(frame_idx)
            x = self.dcd.positions(frame_idx, None if self._all_atoms else self.indices)
            return x, self._dimensions[frame_idx], self._times[frame_idx]
        x = self._positions
        contiguous = len(frame_idx) > 0 and (np.diff(frame_idx) == 1).all()
        if isinstance(x, FrameWindow):
            if not contiguous:
                raise RuntimeError("a FrameWindow trajectory needs a contiguous frame range")
            sel = x[int(frame_idx[0]):int(frame_idx[-1]) + 1]
            if not self._all_atoms:
                sel = sel[:, torch.as_tensor(self.indices)] if torch.is_tensor(sel) else sel[:, self.indices]
        elif torch.is_tensor(x):
            sel = x[int(frame_idx[0]):int(frame_idx[-1]) + 1] if contiguous else x[torch.as_tensor(frame_idx)]
            if not self._all_atoms:
                sel = sel[:, torch.as_tensor(self.indices)]
        else:
            sel = x[int(frame_idx[0]):int(frame_idx[-1]) + 1] if contiguous else x[frame_idx]
            if not self._all_atoms:
                sel = sel[:, self.indices]
        return sel, self._dimensions[frame_idx], self._times[frame_idx]

    # -- MDAnalysis ------------------------------------------------------------------
    def _from_mdanalysis(self, structure, trajectory, sel):
        try:
            import MDAnalysis as mda
        except ImportError as exc:       # pragma: no cover - MDAnalysis absent in this image
            raise RuntimeError("reading structure/trajectory files needs MDAnalysis, which is not "
                               "installed; pass in-memory arrays instead") from exc
        self.mda_universe = mda.Universe(structure, trajectory)
        group = self.mda_universe.select_atoms(sel)
        whole = group.residues.atoms
        if len(whole) != len(group):
            raise RuntimeError("the selection must cover whole residues")
        self.indices = group.indices
        self.masses = np.asarray(group.masses, dtype=np.float64)
        self.names = np.asarray(group.names, dtype=object)
        res = group.residues
        ridx = group.resindices
        remap = {r: i for i, r in enumerate(res.resindices)}
        self.resindex = np.array([remap[r] for r in ridx], dtype=np.int64)
        self.resnames = list(res.resnames)
        self.resids = np.asarray(res.resids)
        self.nframes = len(self.mda_universe.trajectory)

    def _read_mda(self, frame_idx):      # pragma: no cover - MDAnalysis absent in this image
        traj = self.mda_universe.trajectory
        x = np.empty((len(frame_idx), len(self.indices), 3), dtype=np.float32)
        d = np.empty((len(frame_idx), 6), dtype=np.float32)
        t = np.empty(len(frame_idx))
        for k, f in enumerate(frame_idx):
            ts = traj[int(f)]
            x[k] = ts.positions[self.indices]
            d[k] = ts.dimensions
            t[k] = ts.time
        return x, d, t


class _FrameFeeder(object):
    """Host -> device frame batches, double buffered on a copy stream.  In-memory trajectories are
    copied as they are ([f, M, 3]); natively read DCD files go to the device as the raw bytes of
    their frames (threaded positional reads into pinned staging) and are interleaved / atom-selected
    / byte-swapped there (`pbk_planes_to_aos`)."""

    def __init__(self, md, eng):
        dev = eng.dev
        self.md, self.eng, self.B = md, eng, eng.B
        self.stream = torch.cuda.Stream(dev)
        self.main = torch.cuda.current_stream(dev)
        self.bufs = [torch.empty((self.B, md.natoms, 3), dtype=torch.float32, device=dev) for _ in range(2)]
        self.boxes = [torch.empty((self.B, 3), dtype=torch.float32, device=dev) for _ in range(2)]
        self.ready = [torch.cuda.Event() for _ in range(2)]
        self.consumed = [torch.cuda.Event() for _ in range(2)]
        self.copied = [None, None]            # event: the pinned staging buffer of a slot may be rewritten
        self.n = [0, 0]
        self.k_issue = 0
        self.k_take = 0
        lay = md.dcd.raw_layout if md.dcd is not None else None
        self.planar = lay is not None           # natively read DCD: raw file frames go to the device
        if self.planar:
            fb, self.frame_floats, self.plane_offsets, self.byteswap = lay
            self.stage = [torch.empty((self.B, fb), dtype=torch.uint8).pin_memory() for _ in range(2)]
            self.pdev = [torch.empty((self.B, fb), dtype=torch.uint8, device=dev) for _ in range(2)]
            self.sel = None if md._all_atoms else torch.from_numpy(md.indices.astype(np.int32)).to(dev)

    def issue(self, frames):
        """Start the copy of the next batch (at most B frames); returns (dimensions, times)."""
        md = self.md
        slot = self.k_issue % 2
        nf = len(frames)
        if self.planar:
            if self.copied[slot] is not None:
                self.copied[slot].synchronize()       # the staging buffer is free again
            md.read_raw_frames(frames, self.stage[slot].numpy())
            d, t = md._dimensions[frames], md._times[frames]
            src, dst = self.stage[slot][:nf], self.pdev[slot][:nf]
        else:
            x, d, t = md.frames(frames)
            src = x if torch.is_tensor(x) else torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
            dst = self.bufs[slot][:nf]
        bh = torch.from_numpy(np.ascontiguousarray(d[:, :3]))
        with torch.cuda.stream(self.stream):
            if self.k_issue >= 2:
                self.stream.wait_event(self.consumed[slot])
            dst.copy_(src, non_blocking=True)
            self.boxes[slot][:nf].copy_(bh, non_blocking=True)
            self.ready[slot].record(self.stream)
            if self.planar:
                ev = torch.cuda.Event()
                ev.record(self.stream)
                self.copied[slot] = ev
        self.n[slot] = nf
        self.k_issue += 1
        return d, t

    def take(self):
        """Device views (x [f, M, 3], box [f, 3]) of the oldest issued batch, valid until release()."""
        slot = self.k_take % 2
        nf = self.n[slot]
        self.main.wait_event(self.ready[slot])
        if self.planar:
            self.eng.planes_to_aos(self.pdev[slot][:nf].view(torch.float32), self.sel, self.bufs[slot][:nf],
                                   self.plane_offsets, self.byteswap)
        return self.bufs[slot][:nf], self.boxes[slot][:nf]

    def release(self):
        self.consumed[self.k_take % 2].record(self.main)
        self.k_take += 1


class GridBatch(object):
    """Device lipid grids of all analysed frames + their host reductions."""

    def __init__(self, grid_idx, grid_z, cells, thickness_host, apl_host, runs):
        self.grid_idx, self.grid_z, self.cells = grid_idx, grid_z, cells
        self.thickness_host, self.apl_host = thickness_host, apl_host
        self.runs = runs          # [(frame slice, nx, ny, index into the lists above)]

    def frame(self, f):
        for k, (sl, _nx, _ny) in enumerate(self.runs):
            if sl.start <= f < sl.stop:
                j = f - sl.start
                return (self.grid_idx[k][j].cpu().numpy(), self.grid_z[k][j].cpu().numpy(),
                        self.cells[k][j].cpu().numpy())
        raise IndexError(f)


class ReductionBatch(object):
    """What a batched protocol's ``run_batch`` receives: the engine holding the resident
    representations, host metadata, and the leaflets of the first analysed frame."""

    def __init__(self, analyzer, engine, times, box_host, first_leaflets, shard=None, window=None):
        self._analyzer = analyzer
        self.engine = engine
        self.shard = shard                  # pybilt_b200.distributed.Shard or None
        self.lateral = list(analyzer.settings['lateral'])
        self.frame_numbers = analyzer.analysed_frame_indices()      # MD frame index of every analysed frame
        # n_frames / times / box_host describe ALL analysed frames; the engine rows of a
        # shard are [back halo | owned | forward halo | reference row]
        self.n_frames = len(times)
        if window is None:
            # resident run: every analysed frame (of this rank's shard) is in the engine
            self.windowed = False
            self.rows = shard.owned_local() if shard is not None else slice(0, engine.n_done)
            self.ref_row = shard.n_local if shard is not None else 0     # frame 0 of the whole run
            self.gframes = slice(shard.lo, shard.hi) if shard is not None else slice(0, len(times))
            self.row0_global = shard.first_local if shard is not None else 0
            self.first_window = self.last_window = True
            self.sharded = shard is not None
        else:
            # streamed run: the engine holds [context rows | the window's frames]; see
            # BilayerAnalyzer._run_windowed
            self.windowed = True
            self.rows = window['rows']
            self.ref_row = window['ref_row']
            self.gframes = window['gframes']
            self.row0_global = window['row0_global']
            self.first_window, self.last_window = window['first'], window['last']
            self.sharded = window['sharded']
        self.times = times
        self.box_host = box_host
        self.types = engine.layout.types
        self.resids = engine.layout.resids
        cache = getattr(engine, '_type_code_cache', None)
        if cache is None:
            tarr = np.asarray(self.types, dtype=object).astype(str)
            uniq, first, inv = np.unique(tarr, return_index=True, return_inverse=True)
            order = np.argsort(first, kind="stable")             # first-seen order of the resnames
            rank = np.empty(len(uniq), dtype=np.int32)
            rank[order] = np.arange(len(uniq), dtype=np.int32)
            names = [self.types[int(first[k])] for k in order]
            codes = rank[inv].astype(np.int32)
            cache = (names, codes, torch.from_numpy(codes).to(engine.dev))
            engine._type_code_cache = cache
        self.type_names, self.type_code_host, self.type_code_dev = cache
        self.first_leaflets = first_leaflets
        self._grids = None
        self._rdf_multi = {}
        self._nnf_multi = {}

    def row_of_global(self, g):
        """engine row of analysed frame g"""
        return int(g) - self.row0_global

    def allreduce(self, values):
        """Sum of a small host array over the ranks (identity when not sharded)."""
        values = np.asarray(values)
        if not self.sharded:
            return values
        dt = torch.int64 if values.dtype.kind in "iu" else torch.float64
        t = torch.as_tensor(values, dtype=dt, device=self.engine.dev)
        return pd.combine_slots(t).cpu().numpy()

    def type_code_of(self, name):
        if name == 'all':
            return -1
        return self.type_names.index(name) if name in self.type_names else -2

    def rdf_shared(self, edges, lateral):
        """(count [2,T,T,n_bins], type_counts [F,2,T]) host arrays of ONE pair sweep over this
        batch's frames shared by every com_lateral_rdf analysis with these bins, or None when only
        one analysis wants them / the library has no shared path for the shape."""
        key = (np.ascontiguousarray(edges, dtype=np.float64).tobytes(), tuple(lateral))
        if key not in self._rdf_multi:
            res = None
            if self._analyzer._rdf_sharers(key) >= 2 and os.environ.get("PBK_NO_RDF_MULTI") is None:
                r = self.engine.rdf_multi(self.type_code_dev, len(self.type_names), edges, lateral,
                                          frames=self.rows)
                if r is not None:
                    res = (r[0].cpu().numpy(), r[1].cpu().numpy().astype(np.int64))
            self._rdf_multi[key] = res
        return self._rdf_multi[key]

    def nnf_shared(self, k, lateral, spec):
        """(n_b [F, N], frac [F]) device tensors of the nnf analysis ``spec`` = (type_a, type_b,
        leaflet_mask) from ONE neighbour search over this batch's frames shared by every nnf
        analysis with this n_neighbors, or None (a single analysis, or no shared path)."""
        key = (int(k), tuple(lateral))
        if key not in self._nnf_multi:
            res = None
            protos = [self._analyzer._analysis_protocol.command_protocol[a]
                      for a in self._analyzer._analysis_protocol.analysis_ids]
            mine = [p_ for p_ in protos if p_.analysis_key == 'nnf' and int(p_.settings['n_neighbors']) == int(k)]
            if len(mine) >= 2 and os.environ.get("PBK_NO_NNF_MULTI") is None:
                m = self.engine.nnf_multi(self.type_code_dev, len(self.type_names), k, lateral, frames=self.rows)
                if m is not None:
                    specs = []
                    for p_ in mine:
                        st = p_.settings
                        sp = (self.type_code_of(st['resname_1']), self.type_code_of(st['resname_2']),
                              {"both": 3, "upper": 1, "lower": 2}.get(st['leaflet'], 0))
                        if sp not in specs:
                            specs.append(sp)
                    # 'first' resnames are resolved by each protocol before it asks: specs that still hold
                    # an unresolved name (-2) simply never match a request
                    n_b, frac = self.engine.nnf_select_many(m[0], m[1], self.type_code_dev, specs, k,
                                                            frames=self.rows)
                    res = {sp: (n_b[i], frac[i]) for i, sp in enumerate(specs)}
                    res['_tallies'] = m
            self._nnf_multi[key] = res
        table = self._nnf_multi[key]
        if table is None:
            return None
        if spec not in table:           # e.g. a resname resolved from 'first' after the table was made
            m = table['_tallies']
            n_b, frac = self.engine.nnf_select(m[0], m[1], self.type_code_dev, spec[0], spec[1], spec[2], k,
                                               frames=self.rows)
            table[spec] = (n_b, frac)
        return table[spec]

    def lipid_grids(self) -> GridBatch:
        if self._grids is None:
            self._grids = self._analyzer._build_lipid_grids(self)
        return self._grids


class BilayerAnalyzer(object):
    _valid_commands = ["structure", "trajectory", "analysis", "selection", "frames", "plot",
                       "lipid_grid"]
    _required_commands = ['structure', 'trajectory', 'selection']

    def __init__(self, structure=None, trajectory=None, selection=None, input_file=None,
                 input_dict=None, device=0, distributed=True):
        self.next = self.__next__
        self.distributed = distributed      # shard frames across ranks when torch.distributed is up
        self.settings = {'frame_range': [0, -1, 1], 'print_interval': 5, 'norm': 2,
                         'lateral': [0, 1], 'lateral_dimension': "xy", 'normal_dimension': "z",
                         'frame_index': 0}
        self.reps = {'current_mda_frame': None, 'com_frame': None, 'leaflets': None,
                     'lipid_grid': None, 'vector_frame': None, 'first_com_frame': None}
        self.rep_settings = {
            'com_frame': {'dump': False, 'dump_path': "./", 'name_dict': None, 'multi_bead': False,
                          'rewrap': False, 'make_whole': False},
            'leaflets': {'dump': False, 'dump_path': "./", 'update_interval': 1,
                         'assign_method': 'avg_norm', 'orientation_atoms': None},
            'lipid_grid': {'dump': False, 'dump_path': "./", 'n_xbins': 10, 'n_ybins': 10,
                           'area_per_bin': None},
            'vector_frame': {'dump': False, 'dump_path': "./", 'ref_atoms': None}}
        self.device = device
        self.batch_frames = None    # frames per host->device batch (None: about 96 MB of positions)
        # Streamed runs: None = keep every analysed frame's representations resident when they fit
        # `resident_budget_bytes`, else stream windows of frames through a bounded engine; an int
        # forces windows of that many frames (see _run_windowed)
        self.window_frames = None
        self._incremental = True    # batched analyses consume each host->device batch as it arrives
        self._incremental_done = False
        self.resident_budget_bytes = 48 << 30
        self.positions_budget_bytes = 64 << 30   # streamed shards keep their positions on the device up to this
        self.grid_mode = 'opt'      # 'opt' = lipid_grid_opt.py (reference default), 'exact' = lipid_grid.py
        self._input_script_name = os.path.abspath(input_file) if input_file is not None else None
        self._frame_loop_count = 0
        if input_file is not None:
            self._commands = self.parse_input_script(self._input_script_name)
            self._apply_frames_and_grid(self._commands, cast=lambda v: v)
        elif structure is not None and trajectory is not None and selection is not None:
            self._commands = {'structure': structure, 'trajectory': trajectory,
                              'selection': selection}
        elif input_dict is not None and isinstance(input_dict, dict):
            for key in self._required_commands:
                if key not in input_dict:
                    raise RuntimeError("key '{}' needs to be included in the input "
                                       "dictionary.".format(key))
            self._commands = {k: v for k, v in input_dict.items() if k in self._valid_commands}
            self._apply_frames_and_grid(self._commands, cast=int)
        else:
            raise RuntimeError('Must provide input_file or all three options for'
                               '"structure", "trajectory", and "selection"')
        self._analysis_protocol = ap.Analyses(self._commands.get('analysis',
                                                                 default_analysis_commands))
        self._plot_commands = self._commands.get('plot', [])
        self._plot_protocol = pp.PlotProtocol(self._plot_commands, self._analysis_protocol)
        self._mda_data = TrajectoryData(self._commands['structure'], self._commands['trajectory'],
                                        self._commands['selection'])
        self._current_frame = self.settings['frame_range'][0]
        self._last_frame = self.settings['frame_range'][1]
        if self._last_frame < 0:
            self._last_frame += self._mda_data.nframes
        self._engine = None
        self._batch = None
        self._host = None
        self._unwrap_state = None           # unwrap_coordinates_up_to: unwrapped coordinates to continue from
        self._cursor = 0

    def __repr__(self):
        return 'BilayerAnalyzer'

    # -- input script ---------------------------------------------------------------------
    def _apply_frames_and_grid(self, commands, cast):
        if 'frames' in commands:
            f_args = commands['frames']
            if isinstance(f_args, list) and len(f_args) == 1 and isinstance(f_args[0], str):
                f_args = f_args[0].split()
            for i in range(0, len(f_args) - 1, 2):
                slot = {'first': 0, 'last': 1, 'interval': 2}.get(f_args[i])
                if slot is not None:
                    self.settings['frame_range'][slot] = int(f_args[i + 1])
        if 'lipid_grid' in commands:
            lg_args = commands['lipid_grid']
            if isinstance(lg_args, list) and len(lg_args) == 1 and isinstance(lg_args[0], str):
                lg_args = lg_args[0].split()
            for i in range(0, len(lg_args) - 1, 2):
                if lg_args[i] in ('n_xbins', 'n_ybins'):
                    self.rep_settings['lipid_grid'][lg_args[i]] = int(lg_args[i + 1])

    def parse_input_script(self, input_script_name):
        """Line-oriented setup script (bilayer_analyzer.py:479-531)."""
        commands = {}
        with open(input_script_name) as ifile:
            for line in ifile:
                line = line.rstrip("\n")
                words = line.split()
                if len(words) <= 1 or words[0][0] == '#':
                    continue
                key = words[0]
                if key in ('structure', 'trajectory', 'selection'):
                    value = line.replace(key, '', 1).strip()
                    if key == 'selection':
                        value = ast.literal_eval(value)
                    commands[key] = value
                elif key in self._valid_commands:
                    commands.setdefault(key, []).append(" ".join(words[1:]))
                else:
                    raise RuntimeError("invalid input command '{}'".format(key))
        for required in self._required_commands:
            if required not in commands:
                raise RuntimeError("the {} needs to be specified in the input "
                                   "script".format(required))
        return commands

    # -- analysis protocol ----------------------------------------------------------------
    def print_analysis_protocol(self):
        self._analysis_protocol.print_protocol()

    # -- plots (bilayer_analyzer.py:592-653) ------------------------------------------------------
    def print_plot_protocol(self):
        self._plot_protocol.print_protocol()

    def add_plot(self, plot_string):
        self._plot_protocol.add_plot(plot_string, self._analysis_protocol)

    def remove_plot(self, plot_id):
        self._plot_protocol.remove_plot(plot_id)

    def print_plot_ids(self):
        print(self._plot_protocol.plot_ids)

    def get_plot_ids(self):
        return self._plot_protocol.plot_ids

    def show_plot(self, plot_id):
        return self._plot_protocol.command_protocol[plot_id].show_plot(self._analysis_protocol)

    def generate_plot(self, plot_id):
        return self._plot_protocol.command_protocol[plot_id].generate_plot(self._analysis_protocol)

    def save_all_plots(self):
        self._plot_protocol.save_plots(self._analysis_protocol)

    def add_analysis(self, analysis_in):
        self._analysis_protocol.add_analysis(analysis_in)

    def remove_analysis(self, analysis_id):
        self._analysis_protocol.remove_analysis(analysis_id)

    def remove_all_analyses(self):
        self._analysis_protocol.remove_all()

    def get_analysis_ids(self):
        return self._analysis_protocol.analysis_ids

    def get_analysis_data(self, analysis_id):
        return self._analysis_protocol.command_protocol[analysis_id].get_data()

    def dump_data(self, path=None):
        self._analysis_protocol.dump_data(path=path)

    @staticmethod
    def available_analysis():
        return list(ap.command_protocols.keys())

    @staticmethod
    def print_available_analysis():
        print(list(ap.command_protocols.keys()))

    # -- settings -------------------------------------------------------------------------
    def set_frame_range(self, first=0, last=-1, interval=1):
        """bilayer_analyzer.py:716-747 (fractions in (0,1) scale with the trajectory length)."""
        n = self._mda_data.nframes
        if isinstance(first, (float, np.floating)):
            first = int(first * n) if 0.0 < first < 1.0 else int(first)
        if isinstance(last, (float, np.floating)):
            last = int(last * n) if 0.0 < last < 1.0 else int(last)
        if first != self.settings['frame_range'][0]:
            self.settings['frame_range'][0] = first
            self._current_frame = first
        if last != self.settings['frame_range'][1]:
            self.settings['frame_range'][1] = last
            self._last_frame = last
            if self._last_frame < 0:
                self._last_frame += n
        if interval != self.settings['frame_range'][2]:
            self.settings['frame_range'][2] = interval
        self._invalidate()

    def adjust_rep_setting(self, rep_key, setting_key, setting_value):
        self.rep_settings[rep_key][setting_key] = setting_value
        self._invalidate()

    def print_rep_settings(self):
        for rep in self.rep_settings.keys():
            print("represetation: {}".format(rep))
            for setting in self.rep_settings[rep].keys():
                print("  setting:value {}:{}".format(setting, self.rep_settings[rep][setting]))

    def set_dump_com_frame(self, on=True, path="./"):
        self.rep_settings['com_frame']['dump'] = on
        self.rep_settings['com_frame']['dump_path'] = path

    def set_dump_leaflet(self, on=True, path="./"):
        self.rep_settings['leaflets']['dump'] = on
        self.rep_settings['leaflets']['dump_path'] = path

    def set_dump_lipid_grid(self, on=True, path="./"):
        self.rep_settings['lipid_grid']['dump'] = on
        self.rep_settings['lipid_grid']['dump_path'] = path

    def turn_off_representation(self, rep_key):
        self._analysis_protocol.use_objects[rep_key] = False

    def turn_on_representation(self, rep_key):
        self._analysis_protocol.use_objects[rep_key] = True

    def reset(self):
        """Clear analysis outputs and rewind the frame loop (bilayer_analyzer.py:761-774)."""
        self.settings['frame_index'] = 0
        self._current_frame = self.settings['frame_range'][0]
        self._frame_loop_count = 0
        self._unwrap_state = None
        self._analysis_protocol.reset()
        self._invalidate()

    def unwrap_coordinates_up_to(self, frame_index):
        with torch.cuda.device(self.device):       # libpbk launches on the current device
            return self._unwrap_coordinates_up_to(frame_index)

    def _unwrap_coordinates_up_to(self, frame_index):
        """Unwrap the trajectory from frame 0 up to ``frame_index`` (stepping by the frame interval)
        without analysing anything, so that a run whose frame range starts later continues from
        that unwrapped state instead of treating its first frame as the reference
        (bilayer_analyzer.py:1034-1079).  ``reset()`` forgets the state, as it does there."""
        md = self._mda_data
        last = md.nframes if frame_index == -1 else int(frame_index)
        frames = np.arange(0, min(last + 1, md.nframes), self.settings['frame_range'][2])
        if len(frames) == 0:
            return
        cset = self.rep_settings['com_frame']
        layout = build_bead_layout(md.masses, md.names, md.resindex, md.resnames, md.resids,
                                   cset['name_dict'], cset['multi_bead'])
        nb = int(max(1, min(len(frames), (96 << 20) // max(1, md.natoms * 12))))
        eng = LipidReductionEngine(md.masses, layout, 1, device=self.device, norm=self.settings['norm'],
                                   batch_frames=nb, allow_fused=False)
        if self._unwrap_state is not None:
            eng.d_u_state.copy_(self._unwrap_state)
            eng.stream_started = True
        self._push_frames(eng, md, frames, unwrap_only=True)
        eng.check_status()
        self._unwrap_state = eng.d_u_state.clone()
        self._feeder_cache = None

    @staticmethod
    def _orientation_selection(md, layout, orientation_atoms):
        """Atoms behind ``"resname R and resid I and name A"`` (bilayer_analyzer.py:857-861) for the tail
        (``orientation_atoms[R][0]``) and head (``[1]``) of every lipid, as indices into the selected atoms."""
        names = np.asarray(md.names, dtype=object)
        res_of_atom = np.asarray(md.resindex)
        resnames = np.asarray(md.resnames, dtype=object)
        resids = np.asarray(md.resids)
        by_res = {}
        for r in range(len(resnames)):
            by_res.setdefault((resnames[r], int(resids[r])), []).append(r)
        order = np.argsort(res_of_atom, kind="stable")
        starts = np.searchsorted(res_of_atom[order], np.arange(len(resnames) + 1))
        sel_atom, seg = [], [0]
        for i, (resname, resid) in enumerate(zip(layout.types, layout.resids)):
            pair = orientation_atoms[resname]                 # KeyError for a lipid type without an entry, as there
            for which in (0, 1):
                for r in by_res.get((resname, int(resid)), []):
                    atoms = order[starts[r]:starts[r + 1]]
                    sel_atom += [int(a) for a in np.sort(atoms[names[atoms] == pair[which]])]
                seg.append(len(sel_atom))
        sel_atom = np.asarray(sel_atom, dtype=np.int32)
        return sel_atom, np.asarray(md.masses, dtype=np.float64)[sel_atom], np.asarray(seg, dtype=np.int32)

    def _invalidate(self):
        if getattr(self, '_engine', None) is not None:
            self._engine_cache = (self._engine_key, self._engine)
        self._engine = None
        self._batch = None
        self._host = None
        self._cursor = 0

    # -- the device pipeline ----------------------------------------------------------------
    def analysed_frame_indices(self):
        first, _last, interval = self.settings['frame_range']
        return np.arange(first, self._last_frame + 1, interval)

    def _reduce(self):
        """Push every analysed frame through the engine (unwrap -> COMs -> leaflets)."""
        if self._batch is not None:
            return self._batch
        self._incremental_done = False
        try:
            return self._reduce_with(allow_fused=True)
        except FusedPathRejected:
            # a jump the chunked fast path cannot decide locally: sequential kernels instead
            self._incremental_done = False
            return self._reduce_with(allow_fused=False)

    def _reduce_with(self, allow_fused):
        lset = self.rep_settings['leaflets']
        if lset['assign_method'] not in ('avg_norm', 'orientation'):
            raise RuntimeError("leaflets:assign_method '%s' is outside the accelerated path "
                               "('avg_norm', 'orientation')" % lset['assign_method'])
        if self.rep_settings['com_frame']['make_whole']:
            raise RuntimeError("com_frame:make_whole is outside the accelerated path")
        md = self._mda_data
        cset = self.rep_settings['com_frame']
        layout = None
        cached = getattr(self, '_engine_cache', None)
        if cached is None or cached[0][:3] != (id(md), repr(cset['name_dict']), bool(cset['multi_bead'])):
            layout = build_bead_layout(md.masses, md.names, md.resindex, md.resnames, md.resids,
                                       cset['name_dict'], cset['multi_bead'])
        fr_all = self.analysed_frame_indices()
        shard = None
        if self.distributed and pd.is_distributed():
            import torch.distributed as dist
            fwd = 0
            for a_id in self._analysis_protocol.analysis_ids:
                proto = self._analysis_protocol.command_protocol[a_id]
                if proto.analysis_key == 'msd_multi':
                    fwd = max(fwd, int(proto.settings['n_tau']) - 1)
            shard = pd.make_shard(len(fr_all), dist.get_world_size(), dist.get_rank(), fwd,
                                  align=int(lset['update_interval']))
            self._shard_fwd = fwd
            if shard.n_owned == 0:
                raise RuntimeError("more ranks than analysed frames")
            fr = fr_all[shard.first_local:shard.first_local + shard.n_local]
        else:
            fr = fr_all
        self._frames = fr
        self._frames_all = fr_all
        self._shard = shard
        n_rows = len(fr) + (1 if shard is not None else 0)
        # the engine (device buffers, summation plan, libpbk context) is reused across reset()
        # as long as nothing that shapes it changed
        batch_frames = self.batch_frames if shard is None else len(fr)
        if batch_frames is None:
            # ~96 MB of positions per host->device batch: the copy of batch k+1 overlaps the
            # reduction of batch k, and only the last batch's kernels are exposed
            batch_frames = int(max(64, min(len(fr), (96 << 20) // max(1, md.natoms * 12))))
        key = (id(md), repr(cset['name_dict']), bool(cset['multi_bead']), n_rows, self.device,
               self.settings['norm'], int(lset['update_interval']), batch_frames, bool(cset['rewrap']))
        cached = getattr(self, '_engine_cache', None)
        if cached is not None and cached[0] == key:
            eng = cached[1]
            eng.reset()
            eng.allow_fused = allow_fused and eng.fused_supported
            layout = eng.layout
        else:
            if layout is None:
                layout = cached[1].layout
            eng = LipidReductionEngine(md.masses, layout, n_rows, device=self.device,
                                       norm=self.settings['norm'],
                                       leaflet_update_interval=int(lset['update_interval']),
                                       batch_frames=batch_frames, allow_fused=allow_fused, rewrap=bool(cset['rewrap']))
        self._engine_key = key
        self._engine_cache = None
        eng._orient = None
        if lset['assign_method'] == 'orientation':
            if shard is not None:
                raise RuntimeError("leaflets:assign_method 'orientation' is not combined with frame shards")
            eng.set_orientation(*self._orientation_selection(md, layout, lset['orientation_atoms']))
        if self._unwrap_state is not None:
            if shard is not None:
                raise RuntimeError("unwrap_coordinates_up_to is not combined with frame shards")
            eng.d_u_state.copy_(self._unwrap_state)
            eng.stream_started = True
        if shard is not None:
            return self._reduce_sharded(eng, md, fr, fr_all, shard, layout)
        # host -> device in engine-batch sized pieces, double-buffered on a copy stream
        times = np.empty(len(fr))
        dims_all = np.empty((len(fr), 6), dtype=np.float32)
        B = eng.B
        feeder = self._feeder_for(md, eng)
        # batched analyses consume every host->device batch as soon as it is reduced (their kernels
        # and bookkeeping hide under the copy of the next batches); per-frame plugins, iteration and
        # representation dumps read the resident representations afterwards
        protos = [self._analysis_protocol.command_protocol[a] for a in self._analysis_protocol.analysis_ids]
        inc = [p_ for p_ in protos if getattr(p_, 'batched', False)] if self._incremental else []
        starts = list(range(0, len(fr), B))
        if len(inc) * len(starts) > 32:
            # many analyses x many batches: per-launch latencies and host bookkeeping of the
            # per-batch calls would cost more than the overlap hides -- analyse once at the end
            inc = []
        first_leaflets = None

        def issue_copy(k):
            b0 = starts[k]
            b1 = min(len(fr), b0 + B)
            d, t = feeder.issue(fr[b0:b1])
            times[b0:b1] = t
            dims_all[b0:b1] = d

        if starts:
            issue_copy(0)
        for k, b0 in enumerate(starts):
            b1 = min(len(fr), b0 + B)
            xb, bb = feeder.take()
            eng.push(xb, bb)
            feeder.release()
            if k + 1 < len(starts):
                issue_copy(k + 1)            # in flight while this batch is reduced and analysed
            if inc:
                if first_leaflets is None:
                    first_leaflets = lf.leaflets_from_labels(eng.labels[0].cpu().numpy(), layout.types,
                                                             layout.resids)
                win = {'rows': slice(b0, b1), 'ref_row': 0, 'gframes': slice(b0, b1), 'row0_global': 0,
                       'first': b0 == 0, 'last': b1 == len(fr), 'sharded': False}
                wb = ReductionBatch(self, eng, times, dims_all[:, :3], first_leaflets, None, win)
                for p_ in inc:
                    p_.run_batch(self.settings, wb, md)
        eng.check_status()
        if first_leaflets is None:
            first_leaflets = lf.leaflets_from_labels(eng.labels[0].cpu().numpy(), layout.types, layout.resids)
        self._engine = eng
        self._times = times
        self._dims = dims_all
        self._batch = ReductionBatch(self, eng, times, dims_all[:, :3].copy(), first_leaflets)
        self._incremental_done = bool(inc)
        return self._batch

    def _reduce_sharded(self, eng, md, fr, fr_all, shard, layout):
        """Frame-shard variant of the reduction (pybilt_b200/distributed.py)."""
        dev = eng.dev
        # metadata of ALL analysed frames (tiny) is known to every rank
        dims_all = np.asarray(md._dimensions[fr_all], dtype=np.float32) if hasattr(md, '_dimensions') else None
        times_all = np.asarray(md._times[fr_all], dtype=np.float64) if hasattr(md, '_times') else None
        # frames after which the image counts are handed on: up to local row 0 of the next rank
        n_a = pd.net_frames_of(shard, getattr(self, '_shard_fwd', 0), int(self.rep_settings['leaflets']['update_interval']))
        eng.row_offset = shard.first_local      # leaflet rebuild frames are counted in analysed frames of the RUN
        x, d, t = md.frames(fr)
        if dims_all is None:                 # MDAnalysis input: only the shard was read
            raise RuntimeError("sharded runs need in-memory box/time arrays for all frames")
        xh = x if torch.is_tensor(x) else torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
        xd = xh.to(dev, non_blocking=True)
        bd = torch.from_numpy(np.ascontiguousarray(d[:, :3])).to(dev)
        net = torch.zeros((md.natoms, 3), dtype=torch.int32, device=dev)
        # one push of [back halo | owned | forward halo]; the net image counts are taken after the
        # owned frames (net_frames), where the next rank's back halo sits
        eng.push(xd, bd, phase=1, net_out=net, net_frames=n_a)
        off = pd.image_offsets(net)
        seen = torch.zeros_like(net)
        eng.push(xd, bd, phase=2, init_images=off, net_frames=n_a, net_out=seen)
        mismatch = 0
        if not eng.used_fused:
            # streamed / generic kernels: the image counts of the last owned frame must be the
            # rank-local ones plus the offsets (the fused path checks its own decisions:
            # PBK_ST_UNWRAP_AMBIGUOUS)
            if not eng.used_stream:
                seen = eng.d_img[n_a - 1].to(torch.int32)
            mismatch = 0 if bool(torch.equal(seen, off + net)) else 32
        # the status words are OR-ed over the ranks before anything else is exchanged: a rejection (or
        # an error) on one rank is raised on every rank, never leaving the others in a collective
        if eng.check_status(agree=True, extra_bits=mismatch) & 32:
            raise RuntimeError("frame-sharded unwrap: a jump within rounding distance of half a box "
                               "makes the shard decomposition ambiguous; run on one rank")
        # frame 0 of the run: leaflets (index lists) and unwrapped COMs (MSD reference)
        lab0 = eng.labels[0].clone()
        pd.broadcast_from_first(lab0)
        ref = eng.com_unwrap[0].clone()
        pd.broadcast_from_first(ref)
        eng.com_unwrap[shard.n_local].copy_(ref)
        first_leaflets = lf.leaflets_from_labels(lab0.cpu().numpy(), layout.types, layout.resids)
        self._engine = eng
        self._times = times_all[shard.first_local:shard.first_local + shard.n_local]
        self._dims = dims_all[shard.first_local:shard.first_local + shard.n_local]
        self._batch = ReductionBatch(self, eng, times_all, dims_all[:, :3].copy(), first_leaflets, shard)
        return self._batch

    # -- streamed (windowed) runs ----------------------------------------------------------------
    def _check_rep_settings(self):
        lset = self.rep_settings['leaflets']
        if lset['assign_method'] != 'avg_norm':
            raise RuntimeError("leaflets:assign_method '%s' needs a resident run on one device "
                               "(streamed windows build their leaflets with 'avg_norm')" % lset['assign_method'])
        if self.rep_settings['com_frame']['make_whole']:
            raise RuntimeError("com_frame:make_whole is outside the accelerated path")

    def _default_batch_frames(self, n):
        md = self._mda_data
        if self.batch_frames is not None:
            return int(self.batch_frames)
        return int(max(64, min(n, (96 << 20) // max(1, md.natoms * 12))))

    def _window_choice(self):
        """Frames per window of a streamed run, or None for a resident run."""
        protos = [self._analysis_protocol.command_protocol[a] for a in self._analysis_protocol.analysis_ids]
        streamable = (all(getattr(p, 'batched', False) for p in protos) and
                      not any(self.rep_settings[r]['dump'] for r in ('com_frame', 'leaflets', 'lipid_grid')))
        if self.window_frames is not None:
            if not streamable:
                raise RuntimeError("window_frames needs batched analyses only and no representation dumps")
            return int(self.window_frames)
        if not streamable:
            return None
        md = self._mda_data
        n_res = len(md.resnames)
        F = len(self.analysed_frame_indices())
        per_frame = n_res * 56 + 64                  # com + com_unwrap + label + mem_com + box
        if self._uses_lipid_grids():
            g = self.rep_settings['lipid_grid']
            per_frame += 2 * int(g['n_xbins']) * int(g['n_ybins']) * 12 + n_res * 100
        if F * per_frame <= self.resident_budget_bytes:
            return None
        return int(max(256, (8 << 30) // per_frame))

    def _rdf_sharers(self, key):
        """number of registered com_lateral_rdf analyses with the bins / lateral axes of `key`"""
        n = 0
        for a in self._analysis_protocol.analysis_ids:
            p_ = self._analysis_protocol.command_protocol[a]
            if p_.analysis_key != 'com_lateral_rdf':
                continue
            st = p_.settings
            _c, e = np.histogram([-1], bins=st['n_bins'], range=[st['range_inner'], st['range_outer']])
            if (e.tobytes(), tuple(self.settings['lateral'])) == key:
                n += 1
        return n

    def _uses_lipid_grids(self):
        return any(self._analysis_protocol.command_protocol[a].analysis_key in ('apl_grid', 'thickness_grid')
                   for a in self._analysis_protocol.analysis_ids)

    def _feeder_for(self, md, eng):
        """The host->device pipeline of (trajectory, engine): device and pinned staging buffers are
        kept as long as the engine is."""
        fc = getattr(self, '_feeder_cache', None)
        if fc is None or fc[0] is not md or fc[1] is not eng or fc[2].B != eng.B:
            fc = (md, eng, _FrameFeeder(md, eng))
            self._feeder_cache = fc
        return fc[2]

    def _push_frames(self, eng, md, frames, times_out=None, dims_out=None, init_images=None, unwrap_only=False,
                     keep=None, copy_only=False):
        """Host -> device in engine-batch sized pieces (double-buffered on a copy stream) and through
        the engine.  Returns the device views (x, box) of the last frame pushed.  ``keep`` =
        (x [n, M, 3], box [n, 3], first row): the batches are also retained there on the device."""
        feeder = self._feeder_for(md, eng)
        B = eng.B
        starts = list(range(0, len(frames), B))

        def issue(k):
            b0 = starts[k]
            b1 = min(len(frames), b0 + B)
            d, t = feeder.issue(frames[b0:b1])
            if times_out is not None:
                times_out[b0:b1] = t
                dims_out[b0:b1] = d

        last = None
        if starts:
            issue(0)
        for k, b0 in enumerate(starts):
            xb, bb = feeder.take()
            if keep is not None:
                r0 = keep[2] + b0
                keep[0][r0:r0 + xb.shape[0]].copy_(xb)
                keep[1][r0:r0 + xb.shape[0]].copy_(bb)
            if copy_only:
                pass
            elif unwrap_only:
                eng.unwrap_only(xb, bb)
            else:
                eng.push(xb, bb, init_images=init_images if b0 == 0 else None)
            last = (xb[-1].clone(), bb[-1].clone())
            feeder.release()
            if k + 1 < len(starts):
                issue(k + 1)
        return last

    def _run_windowed(self, W, allow_fused=True):
        """Streamed run: the analysed frames go through a bounded engine in windows of ``W`` frames;
        every batched analysis consumes a window (``run_batch``) before the next one replaces it and
        combines / finalises at the end (``finish``).  Resident per window: [C context rows | W
        frames | reference row], C = max(1, largest msd_multi n_tau - 1) -- the context rows are the
        tail of the previous window (origins of multi-origin MSD blocks, previous leaflet labels).

        Frame shards (torch.distributed): rank r streams frames [lo_r - C, hi_r); the first C are
        its back halo (context only).  The image counts at the first frame of its stream come from
        a first pass in which every rank advances the unwrap state alone (no reduction) up to the
        frame where the next rank's stream starts; the counts there are all-gathered and prefixed
        (pybilt_b200/distributed.py).  No collective runs inside the window loop: ranks may have
        different numbers of windows."""
        self._check_rep_settings()
        md = self._mda_data
        cset, lset = self.rep_settings['com_frame'], self.rep_settings['leaflets']
        lkey = (id(md), repr(cset['name_dict']), bool(cset['multi_bead']))
        wc = getattr(self, '_win_cache', None) or {}
        if wc.get('lkey') != lkey:
            wc = {'lkey': lkey, 'layout': build_bead_layout(md.masses, md.names, md.resindex, md.resnames,
                                                            md.resids, cset['name_dict'], cset['multi_bead'])}
        self._win_cache = wc
        layout = wc['layout']
        fr_all = self.analysed_frame_indices()
        F = len(fr_all)
        protos = [self._analysis_protocol.command_protocol[a] for a in self._analysis_protocol.analysis_ids]
        C = 1
        for p_ in protos:
            if p_.analysis_key == 'msd_multi':
                C = max(C, int(p_.settings['n_tau']) - 1)
        world, rank = 1, 0
        if self.distributed and pd.is_distributed():
            import torch.distributed as dist
            world, rank = dist.get_world_size(), dist.get_rank()
        lo, hi = pd.shard_bounds(F, world, rank)
        if hi <= lo:
            raise RuntimeError("more ranks than analysed frames")
        # a stream starts C frames before the first owned one, moved back to a leaflet rebuild frame
        ival = max(1, int(lset['update_interval']))
        starts = [max(0, pd.shard_bounds(F, world, q)[0] - C) for q in range(world)]
        starts = [s_ - s_ % ival for s_ in starts]
        a = starts[rank]
        W = int(max(W, C + 1))
        if hasattr(md, '_dimensions') and hasattr(md, '_times'):
            dims_all = np.asarray(md._dimensions[fr_all], dtype=np.float32)
            times_all = np.asarray(md._times[fr_all], dtype=np.float64)
        elif world > 1:
            raise RuntimeError("sharded runs need in-memory box/time arrays for all frames")
        else:
            dims_all = np.zeros((F, 6), dtype=np.float32)
            times_all = np.zeros(F)
        ekey = (C + W + 1, self.device, self.settings['norm'], int(lset['update_interval']),
                self._default_batch_frames(W), bool(self.rep_settings['com_frame']['rewrap']))
        if wc.get('ekey') == ekey:                     # buffers, plan and context are reused across reset()
            eng = wc['engine']
            eng.reset()
            eng.allow_fused = allow_fused and eng.fused_supported
        else:
            eng = LipidReductionEngine(md.masses, layout, C + W + 1, device=self.device, norm=self.settings['norm'],
                                       leaflet_update_interval=int(lset['update_interval']),
                                       batch_frames=self._default_batch_frames(W), allow_fused=allow_fused,
                                       rewrap=bool(self.rep_settings['com_frame']['rewrap']))
            wc['ekey'], wc['engine'] = ekey, eng
        if self._unwrap_state is not None:
            if world > 1:
                raise RuntimeError("unwrap_coordinates_up_to is not combined with frame shards")
            eng.d_u_state.copy_(self._unwrap_state)
            eng.stream_started = True
        ref_row = C + W
        init = None
        retained = None
        if world > 1:
            # pass A: image counts where the next rank's stream starts
            net = torch.zeros((md.natoms, 3), dtype=torch.int32, device=eng.dev)
            if rank < world - 1 and starts[rank + 1] > a:
                stop = starts[rank + 1]
                # when the shard's positions fit on the device they are retained during this pass and
                # the window loop reads them from there (one host->device transfer instead of two)
                if (hi - a) * md.natoms * 12 <= self.positions_budget_bytes:
                    retained = (torch.empty((hi - a, md.natoms, 3), dtype=torch.float32, device=eng.dev),
                                torch.empty((hi - a, 3), dtype=torch.float32, device=eng.dev))
                last = self._push_frames(eng, md, fr_all[a:stop + 1], unwrap_only=True,
                                         keep=None if retained is None else (retained[0], retained[1], 0))
                net = eng.image_counts_last(last[0], last[1])
                if retained is not None and stop + 1 < hi:
                    self._push_frames(eng, md, fr_all[stop + 1:hi], copy_only=True,
                                      keep=(retained[0], retained[1], stop + 1 - a))
            elif rank == world - 1 and (hi - a) * md.natoms * 12 <= self.positions_budget_bytes:
                # the last rank has no counts to report: it uses the wait for the others' first pass to
                # bring its positions over
                retained = (torch.empty((hi - a, md.natoms, 3), dtype=torch.float32, device=eng.dev),
                            torch.empty((hi - a, 3), dtype=torch.float32, device=eng.dev))
                self._push_frames(eng, md, fr_all[a:hi], copy_only=True, keep=(retained[0], retained[1], 0))
            init = pd.image_offsets(net)
            eng.reset()
        eng.row_offset = a                      # analysed-frame index (of the run) of engine row 0
        rejected = False
        failure = None
        first_leaflets = None
        g = a
        first = True
        try:
            while g < hi:
                g1 = min(hi, g + W)
                if not first:
                    eng.rewind(C)
                ctx = eng.n_done
                if retained is not None:
                    xs, bs = retained[0][g - a:g1 - a], retained[1][g - a:g1 - a]
                    for b0 in range(0, g1 - g, eng.B):
                        eng.push(xs[b0:b0 + eng.B], bs[b0:b0 + eng.B],
                                 init_images=init if (first and b0 == 0) else None)
                elif hasattr(md, '_dimensions'):
                    self._push_frames(eng, md, fr_all[g:g1], init_images=init if first else None)
                else:
                    self._push_frames(eng, md, fr_all[g:g1], times_all[g:g1], dims_all[g:g1],
                                      init_images=init if first else None)
                # first window: every rank is here, the status is agreed on before the broadcasts below;
                # later windows differ in number between ranks, a failure there is exchanged after the loop
                eng.check_status(agree=first and world > 1)
                if first:
                    lab0 = eng.labels[0].clone()
                    ref = eng.com_unwrap[0].clone()
                    if world > 1:                       # frame 0 of the run lives on rank 0
                        pd.broadcast_from_first(lab0)
                        pd.broadcast_from_first(ref)
                    eng.com_unwrap[ref_row].copy_(ref)
                    first_leaflets = lf.leaflets_from_labels(lab0.cpu().numpy(), layout.types, layout.resids)
                own0 = max(g, lo)
                if g1 > own0:
                    win = {'rows': slice(ctx + own0 - g, ctx + g1 - g), 'ref_row': ref_row,
                           'gframes': slice(own0, g1), 'row0_global': g - ctx, 'first': own0 == lo,
                           'last': g1 == hi, 'sharded': world > 1}
                    batch = ReductionBatch(self, eng, times_all, dims_all[:, :3].copy(), first_leaflets,
                                           None, win)
                    for p_ in protos:
                        p_.run_batch(self.settings, batch, md)
                g = g1
                first = False
        except FusedPathRejected:
            if world == 1 or first:              # (first window: raised on every rank together)
                raise
            rejected = True
        except (KeyError, RuntimeError) as exc:
            if world == 1 or first:
                raise
            failure = exc
        if world > 1:
            bits = pd.agree_bits((1 if rejected else 0) | (2 if isinstance(failure, KeyError) else 0) |
                                 (4 if isinstance(failure, RuntimeError) else 0), eng.dev)
            if failure is not None:
                raise failure
            if bits & 2:
                raise KeyError("")
            if bits & 4:
                raise RuntimeError("a rank's reduction failed")
            if bits & 1:
                raise FusedPathRejected("a rank's fused reduction was rejected")
        for p_ in protos:
            p_.finish(batch)
        self._frames_all = fr_all
        self._frames = fr_all[a:hi]
        self._engine = None
        self._batch = None
        return batch


    def _build_lipid_grids(self, batch):
        gs = self.rep_settings['lipid_grid']
        lat = self.settings['lateral']
        # grids are built for the owned frames; index 0 = first owned frame (engine row r0)
        r0 = batch.rows.start
        F = batch.rows.stop - batch.rows.start
        g0 = batch.gframes.start
        if gs['area_per_bin'] is not None:
            bins = [lg.fixed_resolution_bins(batch.box_host[g0 + f][lat[0]], batch.box_host[g0 + f][lat[1]],
                                             gs['area_per_bin']) for f in range(F)]
        else:
            bins = [(int(gs['n_xbins']), int(gs['n_ybins']))] * F
        runs, start = [], 0
        for f in range(1, F + 1):
            if f == F or bins[f] != bins[start]:
                runs.append((slice(start, f), bins[start][0], bins[start][1]))
                start = f
        eng = batch.engine
        gi_l, gz_l, cells_l = [], [], []
        thick = np.empty((F, 2))
        apl = np.empty((F, 1 + 2 * len(batch.type_names)))
        for sl, nx, ny in runs:
            esl = slice(r0 + sl.start, r0 + sl.stop)            # engine rows of this run
            gi, gz, _knn = eng.lipid_grid(nx, ny, self.grid_mode, lat, frames=esl)
            cells, th, ap_ = eng.grid_reduce(gi, gz, batch.type_code_dev, len(batch.type_names), lat,
                                             frames=esl)
            gi_l.append(gi); gz_l.append(gz); cells_l.append(cells)
            thick[sl] = th.cpu().numpy()
            apl[sl] = ap_.cpu().numpy()
        return GridBatch(gi_l, gz_l, cells_l, thick, apl, runs)

    # -- iterator: per-frame host views -------------------------------------------------------
    def __iter__(self):
        return self

    def _start(self):
        batch = self._reduce()
        for a_id in self._analysis_protocol.analysis_ids:
            proto = self._analysis_protocol.command_protocol[a_id]
            if getattr(proto, 'batched', False):
                if not self._incremental_done:
                    proto.run_batch(self.settings, batch, self._mda_data)
                proto.finish(batch)
        self._legacy = [self._analysis_protocol.command_protocol[a] for a in
                        self._analysis_protocol.analysis_ids
                        if not getattr(self._analysis_protocol.command_protocol[a], 'batched', False)]
        self._started = True

    def _host_arrays(self):
        if self._host is None:
            e = self._engine
            self._host = {'com': e.com.cpu().numpy(), 'com_unwrap': e.com_unwrap.cpu().numpy(),
                          'mem_com': e.mem_com.cpu().numpy(), 'labels': e.labels.cpu().numpy()}
        return self._host

    def frame_views(self, f, need_grid=None):
        """Host views of analysed frame ``f``: the objects the reference rebuilds per frame."""
        h = self._host_arrays()
        lay = self._engine.layout
        g = int(f)
        upd = int(self.rep_settings['leaflets']['update_interval'])
        src = g - (g % upd)
        names = None
        if g == src:          # _update_com_frame_leaflet_positions only on rebuild frames (A.10)
            names = np.where(h['labels'][g] == 1, 'upper', 'lower')
        frame = cf.COMFrame(h['com'][g], h['com_unwrap'][g], lay.bead_mass, lay.types, lay.resids,
                            self._dims[g, :3], self._times[g], int(self._frames[g]), h['mem_com'][g],
                            names)
        leaflets = lf.leaflets_from_labels(h['labels'][g], lay.types, lay.resids)
        grid = None
        if need_grid if need_grid is not None else self._analysis_protocol.use_objects['lipid_grid']:
            gb = self._batch.lipid_grids()
            o = g - self._batch.rows.start          # grids are indexed by owned frame
            gi, gz, cells = gb.frame(o)
            grid = lg.LipidGrids(frame, leaflets, self.settings['lateral'], gi, gz,
                                 gb.thickness_host[o], gb.apl_host[o], cells, self._batch.type_names)
        return frame, leaflets, grid

    def __next__(self):
        with torch.cuda.device(self.device):
            return self._next_frame()

    def _next_frame(self):
        if self._batch is None or not getattr(self, '_started', False):
            if self._current_frame > self._last_frame:
                raise StopIteration()
            self._started = False
            self._start()
            self._cursor = self._batch.rows.start
        if self._cursor >= self._batch.rows.stop:
            self._started = False
            raise StopIteration()
        g = self._cursor
        self.settings['frame_index'] = int(self._frames[g])
        need_views = bool(self._legacy) or self._iterating_for_user
        if need_views:
            frame, leaflets, grid = self.frame_views(g)
            x, d, t = self._mda_data.frames(self._frames[g:g + 1])
            self.reps['current_mda_frame'] = _HostFrame(np.asarray(x[0]), d[0], t[0], int(self._frames[g]))
            self.reps['com_frame'] = frame
            if g == self._batch.rows.start:
                self.reps['first_com_frame'] = frame
            self.reps['leaflets'] = leaflets
            self.reps['lipid_grid'] = grid
            self._dump_reps(int(self._frames[g]))
            for proto in self._legacy:
                proto.run_analysis(self.settings, self.reps, self._mda_data)
        self._cursor += 1
        self._frame_loop_count += 1
        self._current_frame = int(self._frames[g]) + self.settings['frame_range'][2]
        return

    _iterating_for_user = True

    def run_analysis(self):
        with torch.cuda.device(self.device):
            return self._run_analysis()

    def _run_analysis(self):
        """Run every registered analysis over the analysed frames (bilayer_analyzer.py:777-790).
        Batched run: no per-frame host views unless a plain per-frame plugin needs them."""
        if self._current_frame > self._last_frame and not getattr(self, '_started', False):
            return
        self._iterating_for_user = False
        try:
            if not getattr(self, '_started', False):
                W = self._window_choice()
                if W is not None:
                    try:
                        self._run_windowed(W, allow_fused=True)
                    except FusedPathRejected:
                        self._analysis_protocol.reset()
                        self._run_windowed(W, allow_fused=False)
                    n_own = len(self._frames)
                    self._frame_loop_count += n_own
                    self.settings['frame_index'] = int(self._frames_all[-1])
                    self._current_frame = int(self._frames_all[-1]) + self.settings['frame_range'][2]
                    self._started = False
                    return
                self._started = False
                self._start()
                self._cursor = self._batch.rows.start
            if self._legacy:
                for _frame in self:
                    pass
            else:                       # nothing needs per-frame host views: skip the frame loop
                n = self._batch.rows.stop
                self._frame_loop_count += n - self._cursor
                self._cursor = n
                self.settings['frame_index'] = int(self._frames_all[-1])
                self._current_frame = int(self._frames_all[-1]) + self.settings['frame_range'][2]
                self._started = False
        finally:
            self._iterating_for_user = True

    def _dump_reps(self, frame_number):
        """Per-frame pickles of the representations (bilayer_analyzer.py:982-1003)."""
        for rep, prefix in (('com_frame', 'com_frame_'), ('leaflets', 'leaflets_'),
                            ('lipid_grid', 'lipid_grid_')):
            if self.rep_settings[rep]['dump'] and self.reps[rep] is not None:
                name = self.rep_settings[rep]['dump_path'] + prefix + str(frame_number) + ".pickle"
                with open(name, 'wb') as ofile:
                    pickle.dump(self.reps[rep], ofile)
