"""Analysis plugin registry and the in-scope analysis protocols, frame-batched on the device.

Drop-in mirror of the plugin surface of pybilt/bilayer_analyzer/analysis_protocols.py:
module registries ``command_protocols`` / ``valid_analysis`` / ``analysis_obj_name_dict`` /
``use_objects`` (:70-76), the ``Analyses`` container (:119-329) and the ``AnalysisProtocol``
base class (:332-501) with the same constructor argument forms ("id key value ..." strings
or dicts), settings handling, ``get_data()`` shapes, ``save_data`` pickles and error
behaviour (``RuntimeError`` for bad input, ``warnings.warn`` for unknown setting keys).

What changes is *when* the work happens.  The reference calls ``run_analysis`` once per
frame and computes in Python; the six protocols registered here ('msd', 'msd_multi',
'com_lateral_rdf', 'nnf', 'apl_grid', 'thickness_grid') implement ``run_batch`` instead,
which receives the device-resident representations of all analysed frames and launches
libpbk kernels over them.  A user plugin written against the reference recipe (a subclass
with a per-frame ``run_analysis(ba_settings, ba_reps, ba_mda_data)``) still works: the
analyzer feeds it host views frame by frame.
"""
from __future__ import annotations

import pickle
import functools
import warnings

import numpy as np
import torch

from .running_stats import RunningStats, running_rows

command_protocols = {}
valid_analysis = []
analysis_obj_name_dict = {}
use_objects = {"mda_frame": True, "com_frame": True, "lipid_grid": False, "vector_frame": False}


def word_list_to_string(word_list, delimeter=" "):
    return str(delimeter.join(word_list))


class Analyses(object):
    """The set of analyses a BilayerAnalyzer runs (analysis_protocols.py:119-329)."""

    def __init__(self, analysis_commands):
        self.use_objects = use_objects.copy()
        self.in_commands = analysis_commands
        self.arguments = []
        self.analysis_keys = []
        self.analysis_ids = []
        self.command_protocol = dict()
        self.n_commands = 0
        for command in analysis_commands:
            self.add_analysis(command)
        if self.use_objects['lipid_grid']:
            self.use_objects['com_frame'] = True

    def __getitem__(self, item):
        return self.command_protocol[item]

    def __len__(self):
        return len(self.analysis_ids)

    def add_analysis(self, inputs):
        if isinstance(inputs, str):
            words = inputs.split()
            if len(words) < 2:
                raise RuntimeError("wrong number of arguments for analysis {}".format(words))
            self._register(words[0], words[1], word_list_to_string(words[1:]))
        elif isinstance(inputs, (list, tuple)):
            if len(inputs) != 3:
                raise RuntimeError("wrong number of arguments in the input list/tuple. Should be "
                                   "three: [a_key, a_id, a_set_dict]")
            args = inputs[2]
            args['analysis_id'] = inputs[1]
            self._register(inputs[0], inputs[1], args)
        elif isinstance(inputs, dict):
            for key in ['analysis_key', 'analysis_id', "analysis_settings"]:
                if key not in inputs:
                    raise RuntimeError("required key: {} , not in the input dictionary.".format(key))
            args = inputs['analysis_settings']
            args['analysis_id'] = inputs['analysis_id']
            self._register(inputs['analysis_key'], inputs['analysis_id'], args)

    def _register(self, key, a_id, args):
        if key not in valid_analysis:
            raise RuntimeError("invalid analysis_key '{}' : {}".format(key, args))
        if a_id in self.analysis_ids:
            raise RuntimeError("analysis_id '{}' has already been used!".format(a_id))
        self.use_objects[analysis_obj_name_dict[key]] = True
        self.arguments.append(args)
        self.analysis_ids.append(a_id)
        self.analysis_keys.append(key)
        self.command_protocol[a_id] = command_protocols[key](args)
        self.n_commands += 1

    def remove_analysis(self, analysis_id):
        if analysis_id in self.analysis_ids:
            del self.command_protocol[analysis_id]
            index = self.analysis_ids.index(analysis_id)
            del self.arguments[index]
            del self.analysis_keys[index]
            del self.analysis_ids[index]
            self.n_commands -= 1
        else:
            warnings.warn("no analysis with id '{}'".format(analysis_id))

    def remove_all(self):
        for a_id in list(self.analysis_ids):
            self.remove_analysis(a_id)

    def print_protocol(self):
        print('build objects:')
        for key in self.use_objects.keys():
            if self.use_objects[key]:
                print(key)
        print("with analysis:")
        for analysis_id in self.analysis_ids:
            self.command_protocol[analysis_id].print_protocol()

    def dump_data(self, path=None):
        print('dumping analysis data to pickle files...')
        for analysis_id in self.analysis_ids:
            print("analysis id: {} ---> {} ".format(
                analysis_id, self.command_protocol[analysis_id].save_file_name))
            self.command_protocol[analysis_id].save_data(path=path)

    def reset(self):
        for analysis_id in self.analysis_ids:
            self.command_protocol[analysis_id].reset()


class AnalysisProtocol(object):
    """Base class of every analysis plugin (analysis_protocols.py:332-501)."""
    _pickleable = True
    _short_description = 'None'
    analysis_key = 'none'
    _related = None
    batched = False         # True: implements run_batch(ba_settings, batch, ba_mda_data)

    def __init__(self, args):
        self._short_description = "parent analysis protocol"
        self._return_length = 1
        self.analysis_key = 'none'
        self.analysis_id = 'none'
        self.settings = dict()
        self.settings['none'] = None
        self._valid_settings = list(self.settings.keys())
        self.save_file_name = self.analysis_id + ".pickle"
        self._parse_args(args)
        self.analysis_output = []

    def _setup(self, defaults, args):
        """Common constructor tail of the concrete protocols."""
        self.analysis_id = 'none'
        self.settings = dict(defaults)
        self._valid_settings = list(self.settings.keys())
        self._parse_args(args)
        self.save_file_name = self.analysis_id + ".pickle"
        self.analysis_output = []

    def _parse_args(self, args):
        if isinstance(args, str):
            self._parse_string(args)
        elif isinstance(args, dict):
            self._parse_dict(args)
        else:
            raise RuntimeError("Invalid input structure for arguments/settings to analysis type "
                               + self.analysis_key)

    def _parse_string(self, args):
        self._parse_dict(self._cast_settings(self._parse_str_to_dict(args)))

    def _parse_str_to_dict(self, args):
        words = args.split()
        arg_dict = {'analysis_id': words[0]}
        for i in range(1, len(words) - 1, 2):
            if words[i] in self._valid_settings:
                arg_dict[words[i]] = words[i + 1]
            else:
                warnings.warn("ignoring invalid argument key " + words[i] + " for analysis "
                              + self.analysis_key)
        return arg_dict

    _casts = {}

    def _cast_settings(self, arg_dict):
        for key, cast in self._casts.items():
            if key in arg_dict:
                arg_dict[key] = cast(arg_dict[key])
        return arg_dict

    def _parse_dict(self, args):
        if 'analysis_id' not in args:
            raise RuntimeError("required key 'anlaysis_id' not assigned in input dict for analysis "
                               "type: '" + self.analysis_key + "'")
        for key, value in args.items():
            if key in self._valid_settings:
                self.settings[key] = value
            elif key == 'analysis_id':
                self.analysis_id = value
            else:
                warnings.warn("ignoring invalid argument key " + key + " for analysis "
                              + self.analysis_key)

    def short_description(self):
        return self._short_description

    def pickleable(self):
        return self._pickleable

    def print_protocol(self):
        print("Analysis: " + self._short_description)
        print("  with analysis_id: {} ".format(self.analysis_id))
        print("   and settings: ")
        for key in self.settings.keys():
            print("    {}: {} ".format(key, self.settings[key]))

    def run_analysis(self, ba_settings, ba_reps, ba_mda_data):
        """Per-frame entry point of the reference recipe.  Batched protocols do their work in
        run_batch and ignore per-frame calls; plain plugins override this."""
        if not self.batched:
            self.analysis_output.append(np.zeros(self._return_length))

    def save_data(self, path=None):
        save_file = self.save_file_name if path is None else path + self.save_file_name
        with open(save_file, 'wb') as outfile:
            pickle.dump(self.get_data(), outfile)

    def get_data(self):
        return np.array(self.analysis_output)

    def reset(self):
        self.analysis_output = []


def _protocol(key, rep):
    def deco(cls):
        valid_analysis.append(key)
        analysis_obj_name_dict[key] = rep
        command_protocols[key] = cls
        cls.analysis_key = key
        return cls
    return deco


def _first_frame_indices(leaflets, leaflet, group):
    """Index list frozen at the first analysed frame (analysis_protocols.py:586-605):
    'both' (and any unknown name, after a warning) = upper members then lower members."""
    if leaflet in ("upper", "lower"):
        return list(leaflets[leaflet].get_group_indices(group))
    if leaflet != "both":
        print("!! Warning - request for unknown leaflet name \'", leaflet)
        print("!! the options are \"upper\", \"lower\", or \"both\"--using the default \"both\"")
    out = []
    for name in leaflets:
        out += leaflets[name].get_group_indices(group)
    return out


def _leaflet_mask(leaflet):
    return {"both": 3, "upper": 1, "lower": 2}.get(leaflet, 0)


def _lipid_types_in(leaflets, do_leaflet):
    names, nlipids = [], 0
    for leaf in do_leaflet:
        nlipids += len(leaflets[leaf].get_member_indices())
        for g in leaflets[leaf].get_group_names():
            if g not in names:
                names.append(g)
    return names, nlipids


# ---------------------------------------------------------------------------------------
@_protocol('msd', 'com_frame')
class MSDProtocol(AnalysisProtocol):
    """Single time origin lateral MSD, rows [time, msd] (analysis_protocols.py:509-667)."""
    _short_description = "Single time origin mean squared displacement."
    _related = list(['msd_multi'])
    batched = True

    def __init__(self, args):
        self._indices = []
        self._setup({'leaflet': 'both', 'resname': 'all'}, args)

    def run_batch(self, ba_settings, batch, ba_mda_data):
        """One window of analysed frames (all of them in a resident run)."""
        if batch.first_window:
            self._indices = _first_frame_indices(batch.first_leaflets, self.settings['leaflet'],
                                                 self.settings['resname'])
            self._vals = np.zeros(batch.n_frames)
            self._pending = []               # (global frame slice, device values): read back in finish()
            self._idx_dev = batch.engine._i32(np.asarray(self._indices, dtype=np.int32))
        rows = batch.rows                    # engine rows of the frames this rank owns
        dev = batch.engine.dev               # built on the device: no small host->device copy has to
        ends = torch.arange(rows.start, rows.stop, dtype=torch.int32, device=dev)   # queue behind a bulk one
        origins = torch.full((rows.stop - rows.start,), batch.ref_row, dtype=torch.int32, device=dev)
        vals = batch.engine.msd_pairs(self._idx_dev, origins, ends, ba_settings['lateral'])
        self._pending.append((batch.gframes, vals))

    def finish(self, batch):
        for g, v in self._pending:
            self._vals[g] = v.cpu().numpy()
        self._pending = []
        vals = batch.allreduce(self._vals)       # every frame has exactly one owner: x + 0 = x
        self.analysis_output = np.stack([np.asarray(batch.times, dtype=np.float64), vals], axis=1)


@_protocol('flip_flop', 'com_frame')
class FlipFlopProtocol(AnalysisProtocol):
    """Lipid flip-flop events: {resname: {'count': n, 'events': [[frame, time, resid, from, to]]}}
    (analysis_protocols.py:3731-3868).  The device flags the frames in which any label changed
    (`pbk_label_changes`); for those the events are listed on the host with the same Python set
    operations the reference uses, so that their order within a frame is the reference's."""
    _short_description = "Count lipid flip flops."
    _related = None
    batched = True

    def __init__(self, args):
        self._setup({'none': None}, args)
        self.analysis_output = {}
        self._resnames = []

    def run_batch(self, ba_settings, batch, ba_mda_data):
        if batch.first_window:
            self._resnames = []
            for leaflet in batch.first_leaflets.keys():
                for resname in batch.first_leaflets[leaflet].get_group_names():
                    if resname not in self._resnames:
                        self._resnames.append(resname)
            self._events = []                     # (resname, event) in frame order, this rank's frames
        rows, g0 = batch.rows, batch.gframes.start
        changed = batch.engine.label_changes(frames=rows).cpu().numpy()
        labels = batch.engine.labels
        for k in np.nonzero(changed)[0]:
            r, g = rows.start + int(k), g0 + int(k)
            if g == 0:
                continue                          # the first analysed frame only sets the reference
            pair = labels[r - 1:r + 1].cpu().numpy()
            ref = set(np.nonzero(pair[0] == 1)[0].tolist())
            curr = set(np.nonzero(pair[1] == 1)[0].tolist())
            frame, time = int(batch.frame_numbers[g]), batch.times[g]
            for index in curr.difference(ref):                       # lower -> upper
                self._events.append((batch.types[index], [frame, time, batch.resids[index], 'lower', 'upper']))
            for index in ref.difference(curr):                       # upper -> lower
                self._events.append((batch.types[index], [frame, time, batch.resids[index], 'upper', 'lower']))

    def finish(self, batch):
        events = self._events
        if batch.sharded:
            import torch.distributed as dist
            parts = [None] * dist.get_world_size()
            dist.all_gather_object(parts, events)                    # rank order = frame order
            events = [e for part in parts for e in part]
        self.analysis_output = {r: {'count': 0, 'events': []} for r in self._resnames}
        for resname, ev in events:
            self.analysis_output[resname]['count'] += 1
            self.analysis_output[resname]['events'].append(ev)

    def get_data(self):
        return self.analysis_output

    def reset(self):
        self.analysis_output = {}
        self._resnames = []


@_protocol('ald', 'com_frame')
class ALDProtocol(AnalysisProtocol):
    """Average lateral displacement, rows [time, running mean over frames of the mean displacement
    length from the first frame] (analysis_protocols.py:3218-3345)."""
    _short_description = "Average lateral displacement."
    batched = True

    def __init__(self, args):
        self._indices = []
        self._setup({'leaflet': 'both', 'resname': 'all'}, args)
        self.L_stat = RunningStats()

    def run_batch(self, ba_settings, batch, ba_mda_data):
        if batch.first_window:
            self._indices = _first_frame_indices(batch.first_leaflets, self.settings['leaflet'],
                                                 self.settings['resname'])
            self._vals = np.zeros(batch.n_frames)
            self._pending = []
            self._idx_dev = batch.engine._i32(np.asarray(self._indices, dtype=np.int32))
        rows = batch.rows
        dev = batch.engine.dev
        ends = torch.arange(rows.start, rows.stop, dtype=torch.int32, device=dev)
        origins = torch.full((rows.stop - rows.start,), batch.ref_row, dtype=torch.int32, device=dev)
        vals = batch.engine.msd_pairs(self._idx_dev, origins, ends, ba_settings['lateral'], lengths=True)
        self._pending.append((batch.gframes, vals))

    def finish(self, batch):
        for g, v in self._pending:
            self._vals[g] = v.cpu().numpy()
        self._pending = []
        vals = batch.allreduce(self._vals)
        for t, L in zip(np.asarray(batch.times, dtype=np.float64), vals):
            self.L_stat.push(L)
            self.analysis_output.append(np.array([t, self.L_stat.mean()]))

    def reset(self):
        self.L_stat.reset()
        self.analysis_output = []
        self._indices = []


def msd_multi_schedule(n_frames, n_tau, n_sigma):
    """Replay of the block counters of MSDMultiProtocol.run_analysis
    (analysis_protocols.py:833-888): returns ({block: (origin, end frame)}, n_blocks).
    A pure function of its arguments, memoised (an O(n_frames) Python loop)."""
    finished, n_blocks = _msd_multi_schedule(int(n_frames), int(n_tau), int(n_sigma))
    return dict(finished), n_blocks


@functools.lru_cache(maxsize=64)
def _msd_multi_schedule(n_frames, n_tau, n_sigma):
    finished = {}
    origins, counters = [0], [0]
    tau_counter = sigma_counter = 0
    for f in range(n_frames):
        opening = sigma_counter == n_sigma
        if opening:
            origins.append(f)
            counters.append(0)
        if sigma_counter < n_sigma or opening:
            for b in range(tau_counter, len(origins)):
                if counters[b] < n_tau:
                    counters[b] += 1
                if counters[b] == n_tau:
                    finished[b] = (origins[b], f)
                    tau_counter += 1
        if opening:
            sigma_counter = 0
        sigma_counter += 1
    return finished, len(origins)


@_protocol('msd_multi', 'com_frame')
class MSDMultiProtocol(AnalysisProtocol):
    """Multiple time origin MSD -> [msd, stderr, D, D stderr] (analysis_protocols.py:678-935)."""
    _short_description = "Multiple time origin mean squared displacement."
    _related = list(['msd'])
    _casts = {'n_tau': int, 'n_sigma': int}
    batched = True

    def __init__(self, args):
        self._indices = []
        self._tau = 0.0
        self._setup({'leaflet': 'both', 'resname': 'all', 'n_tau': 50, 'n_sigma': 50}, args)

    def run_batch(self, ba_settings, batch, ba_mda_data):
        n_tau, n_sigma = self.settings['n_tau'], self.settings['n_sigma']
        F = batch.n_frames
        if batch.first_window:
            self._indices = _first_frame_indices(batch.first_leaflets, self.settings['leaflet'],
                                                 self.settings['resname'])
            self._finished, n_blocks = msd_multi_schedule(F, n_tau, n_sigma)
            self._out = np.zeros(n_blocks)
            self._pending = []
            self._idx_dev = batch.engine._i32(np.asarray(self._indices, dtype=np.int32))
        if batch.first_window:
            # all (origin, end) frame pairs go to the device once; blocks finish in frame order, so
            # the blocks of a window are a contiguous run of them
            self._blocks = np.array(sorted(self._finished), dtype=np.int64)
            self._b_origin = np.array([self._finished[b][0] for b in self._blocks], dtype=np.int64)
            self._b_end = np.array([self._finished[b][1] for b in self._blocks], dtype=np.int64)
            self._b_origin_dev = batch.engine._i32(self._b_origin.astype(np.int32))
            self._b_end_dev = batch.engine._i32(self._b_end.astype(np.int32))
        shard = batch.shard
        if batch.windowed:
            # streamed run: a window evaluates the blocks that END in it; the origin frame lies at
            # most n_tau - 1 frames back, i.e. in the window or in the context rows kept before it
            g = batch.gframes
            sel = np.nonzero((self._b_end >= g.start) & (self._b_end < g.stop))[0]
        elif shard is not None:
            # resident shard: a rank evaluates the blocks whose origin frame it owns (the end frame
            # lies in its forward halo)
            sel = np.nonzero((self._b_origin >= shard.lo) & (self._b_origin < shard.hi))[0]
        else:
            sel = np.arange(len(self._blocks))
        if len(sel):
            i0, i1 = int(sel[0]), int(sel[-1]) + 1
            assert i1 - i0 == len(sel)
            vals = batch.engine.msd_pairs(self._idx_dev, self._b_origin_dev[i0:i1] - batch.row0_global,
                                          self._b_end_dev[i0:i1] - batch.row0_global, ba_settings['lateral'])
            self._pending.append((self._blocks[i0:i1], vals))

    def finish(self, batch):
        for blocks, v in self._pending:
            self._out[blocks] = v.cpu().numpy()
        self._pending = []
        n_tau = self.settings['n_tau']
        self.analysis_output = list(batch.allreduce(self._out))
        self._tau = batch.times[n_tau - 1] - batch.times[0] if batch.n_frames >= n_tau else 0.0

    def _process_output(self, analysis_output):
        vals = np.array([v for v in analysis_output if v > 0.0])
        with np.errstate(all="ignore"), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            avg = vals.mean()
            err = vals.std() / np.sqrt(len(vals))
            return np.array([avg, err, avg / (4.0 * self._tau), err / (4.0 * self._tau)])

    def get_data(self):
        return self._process_output(self.analysis_output)

    def reset(self):
        self.analysis_output = []
        self._indices = []


# ---------------------------------------------------------------------------------------
@_protocol('com_lateral_rdf', 'com_frame')
class COMLateralRDFProtocol(AnalysisProtocol):
    """2-d lateral RDF of lipid COMs -> (rdf[n_bins], bin centres)
    (analysis_protocols.py:4792-5062)."""
    _short_description = "Lipid-lipid RDF in the bilayer lateral plane."
    _related = list(['nnf'])
    _casts = {'n_bins': int, 'range_inner': float, 'range_outer': float}
    batched = True

    def __init__(self, args):
        self._setup({'leaflet': 'both', 'resname_1': 'first', 'resname_2': 'first', 'n_bins': 25,
                     'range_inner': 0.0, 'range_outer': 25.0}, args)
        self.reset()

    def reset(self):
        self.analysis_output = []
        self._count = None
        self._edges = None
        self._bins = None
        self._n_frames = 0
        self._area_run = RunningStats()
        self._N_a = self._N_b = self._N = 0
        self._n_leaf = 0

    def run_batch(self, ba_settings, batch, ba_mda_data):
        s = self.settings
        if batch.first_window:
            do_leaflet = ['upper', 'lower'] if s['leaflet'] == 'both' else [s['leaflet']]
            rdf_range = [s['range_inner'], s['range_outer']]
            # same call the reference uses to lay out the bins (:4909-4916)
            _c, edges = np.histogram([-1], bins=s['n_bins'], range=rdf_range)
            self._edges = edges
            self._bins = 0.5 * (edges[:-1] + edges[1:])
            lipid_types, _n = _lipid_types_in(batch.first_leaflets, do_leaflet)
            if s['resname_1'] == 'first':
                s['resname_1'] = lipid_types[0]
            if s['resname_2'] == 'first':
                s['resname_2'] = lipid_types[0]
            self.lipid_types, self.n_ltypes = lipid_types, len(lipid_types)
            self._n_leaf = float(len(do_leaflet))
            self._count_acc = np.zeros(s['n_bins'], dtype=np.int64)
            self._sums_acc = np.zeros(3, dtype=np.int64)
            self._pending = []               # device (count, frame counts) per window, read back in finish()
        shared = batch.rdf_shared(self._edges, ba_settings['lateral'])
        if shared is not None:
            # one sweep binned every (leaflet, type_i, type_j) table; this analysis is a sum of tables
            table, tc = shared                       # [2, T, T, n_bins], [F, 2, T]
            T = table.shape[1]
            ca, cb = batch.type_code_of(s['resname_1']), batch.type_code_of(s['resname_2'])
            A = list(range(T)) if ca == -1 else ([ca] if 0 <= ca < T else [])
            B = list(range(T)) if cb == -1 else ([cb] if 0 <= cb < T else [])
            mask = _leaflet_mask(s['leaflet'])
            for l in range(2):
                if not (mask >> l) & 1:
                    continue
                if A and B:
                    self._count_acc += table[l][np.ix_(A, B)].sum(axis=(0, 1))
                n_a = tc[:, l, A].sum(axis=1) if A else np.zeros(len(tc), dtype=np.int64)
                n_b = tc[:, l, B].sum(axis=1) if B else np.zeros(len(tc), dtype=np.int64)
                on = (n_a > 0) & (n_b > 0)           # the leaflet holds both groups (:4965)
                self._sums_acc += np.array([n_a[on].sum(), n_b[on].sum(), tc[:, l, :].sum()])
            return
        count, fc = batch.engine.rdf(batch.type_code_dev, batch.type_code_of(s['resname_1']),
                                     batch.type_code_of(s['resname_2']), _leaflet_mask(s['leaflet']),
                                     self._edges, ba_settings['lateral'], frames=batch.rows)
        self._pending.append((count, fc))

    def finish(self, batch):
        for count, fc in self._pending:
            fc = fc.cpu().numpy().astype(np.int64)
            self._count_acc += count.cpu().numpy()
            self._sums_acc += np.array([fc[:, :, 0].sum(), fc[:, :, 1].sum(), fc[:, :, 2].sum()])
        self._pending = []
        # frame shards: partial histograms and (N_a, N_b, N) sums are all-reduced
        count = batch.allreduce(self._count_acc).astype(np.float64)
        self._count = count if self._count is None else self._count + count
        sums = batch.allreduce(self._sums_acc)
        self._N_a += int(sums[0])
        self._N_b += int(sums[1])
        self._N += int(sums[2])
        lat = list(batch.lateral)
        # box[lateral].prod() per frame is a float32 product (:5004) pushed into a Welford mean
        areas = batch.box_host[:, lat[0]] * batch.box_host[:, lat[1]]
        if self._area_run.n == 0 and len(areas) and (areas == areas[0]).all():
            # constant box: every Welford update adds exactly zero
            self._area_run.push(areas[0])
            self._area_run.n = len(areas)
            if len(areas) > 1:
                self._area_run._mean = self._area_run._mean + np.float64(0.0)   # float64 after push 2
        else:
            for a in areas:
                self._area_run.push(a)
        self._n_frames += batch.n_frames

    def get_data(self):
        shell = np.power(self._edges[1:], 2) - np.power(self._edges[:-1], 2)
        shell *= np.pi
        n_a = self._N_a / (self._n_leaf * self._n_frames)
        n_b = self._N_b / (self._n_leaf * self._n_frames)
        n_pairs = n_a * n_b
        r1, r2 = self.settings['resname_1'], self.settings['resname_2']
        if r1 == r2:
            n_pairs -= n_a
        elif r1 == 'all' and r2 != 'all':
            n_pairs -= n_b
        elif r2 == 'all' and r1 != 'all':
            n_pairs -= n_a
        with np.errstate(all="ignore"):
            pair_density = n_pairs / self._area_run.mean()
            rdf = (self._count * (1.0 / (pair_density * shell))) / self._n_frames
        return rdf, self._bins


# ---------------------------------------------------------------------------------------
@_protocol('nnf', 'com_frame')
class NNFProtocol(AnalysisProtocol):
    """Nearest-neighbour fraction, rows [time, f, running mean, running deviation]
    (analysis_protocols.py:2075-2257)."""
    _short_description = "Lateral order nearest neighbor fraction."
    _casts = {'n_neighbors': int}
    batched = True

    def __init__(self, args):
        self._setup({'leaflet': 'both', 'n_neighbors': 5, 'resname_1': 'first',
                     'resname_2': 'first'}, args)
        self.running = RunningStats()
        self.neighbor_counts = None

    def reset(self):
        self.running.reset()
        self.analysis_output = []

    def run_batch(self, ba_settings, batch, ba_mda_data):
        s = self.settings
        if batch.first_window:
            do_leaflet = ['upper', 'lower'] if s['leaflet'] == 'both' else [s['leaflet']]
            lipid_types, nlipids = _lipid_types_in(batch.first_leaflets, do_leaflet)
            if s['resname_1'] == 'first':
                s['resname_1'] = lipid_types[0]
            if s['resname_2'] == 'first':
                s['resname_2'] = lipid_types[0]
            if s['n_neighbors'] > nlipids:                       # :2197-2198
                s['n_neighbors'] = nlipids - 1
            self.lipid_types, self.n_ltypes = lipid_types, len(lipid_types)
            self._frac = np.zeros(batch.n_frames)
        shared = batch.nnf_shared(s['n_neighbors'], ba_settings['lateral'],
                                  (batch.type_code_of(s['resname_1']), batch.type_code_of(s['resname_2']),
                                   _leaflet_mask(s['leaflet'])))
        if shared is not None:      # one neighbour search (and one selection pass) for every nnf analysis with this k
            n_b, frac = shared
        else:
            n_b, frac = batch.engine.nnf(batch.type_code_dev, batch.type_code_of(s['resname_1']),
                                         batch.type_code_of(s['resname_2']), _leaflet_mask(s['leaflet']),
                                         s['n_neighbors'], ba_settings['lateral'], frames=batch.rows)
        self.neighbor_counts = n_b            # device int32 [frames of the last window, N]; -1 = not a reference lipid
        self._frac[batch.gframes] = frac.cpu().numpy()

    def finish(self, batch):
        rows, self.running = running_rows(batch.times, batch.allreduce(self._frac))
        self.analysis_output = [list(r) for r in rows]


# ---------------------------------------------------------------------------------------
@_protocol('thickness_leaflet_distance', 'com_frame')
class LeaftoLeafDistProtocol(AnalysisProtocol):
    """Bilayer thickness as the distance between the leaflets' centres of mass along the normal,
    rows [time, thickness, running mean, running deviation] (analysis_protocols.py:1105-1176)."""
    _short_description = "Distance between the leaflets' centers of mass along the normal."
    _related = list(['lipid_length', 'thickness_grid', 'thickness_peak_distance',
                     'thickness_mass_weighted_std'])
    batched = True

    def __init__(self, args):
        self._setup({'none': None}, args)
        self.running = RunningStats()

    def run_batch(self, ba_settings, batch, ba_mda_data):
        if batch.first_window:
            self._vals = np.zeros(batch.n_frames)
            self._pending = []
        if batch.engine.update_interval != 1 and batch.n_frames > 1:
            # frames between leaflet rebuilds carry no leaflet names in the reference: its mean over
            # an empty selection is a nan scalar and `[norm_ind]` on it raises (:1158)
            raise IndexError("invalid index to scalar variable.")
        self._pending.append((batch.gframes, batch.engine.leaflet_distance(frames=batch.rows)))

    def finish(self, batch):
        for g, v in self._pending:
            self._vals[g] = v.cpu().numpy()
        self._pending = []
        rows, self.running = running_rows(batch.times, batch.allreduce(self._vals))
        self.analysis_output = [r for r in rows]

    def reset(self):
        self.running.reset()
        self.analysis_output = []


# ---------------------------------------------------------------------------------------
def _truth(v):
    return v in ['True', 'true']


def _resident_only(batch, key):
    if batch.windowed or batch.sharded:
        raise RuntimeError("analysis '%s' needs every analysed frame resident on one device "
                           "(no streamed windows / frame shards)" % key)


def _frame_leaflets(batch, g):
    """Host views of the leaflets of analysed frame g (rebuilt from the device labels)."""
    from . import leaflet as lf
    h = batch._analyzer._host_arrays()
    return lf.leaflets_from_labels(h['labels'][batch.row_of_global(g)], batch.types, batch.resids)


def _interval_pairs(batch, first_md_frame, interval):
    """(previous, current) analysed-frame pairs the reference's interval bookkeeping selects: the last
    used frame starts at frame_range[0] and moves to the current frame whenever the MD frame distance
    equals `interval` (analysis_protocols.py:1772-1781, :5709-5713)."""
    pairs, last_g, last_md = [], 0, int(first_md_frame)
    for g in range(1, batch.n_frames):
        md = int(batch.frame_numbers[g])
        if md - last_md == interval:
            pairs.append((last_g, g))
            last_g, last_md = g, md
    return pairs


@_protocol('disp_vec', 'com_frame')
class DispVecProtocol(AnalysisProtocol):
    """Lateral displacement vectors between frames `interval` MD frames apart: per pair of frames
    [[x, y, dx, dy] per lipid, resnames] with the lipids of the CURRENT frame's leaflets
    (analysis_protocols.py:1655-1897).  Gathers of the resident COMs, formed on the host."""
    _short_description = "Displacement vectors."
    _related = list(['disp_vec_corr', 'disp_vec_nncorr', 'disp_vec_corr_avg', 'spatial_velocity_corr'])
    _casts = {'interval': int, 'wrapped': _truth, 'scale': _truth, 'scale_to_max': _truth}
    batched = True

    def __init__(self, args):
        self._setup({'leaflet': 'both', 'resname': 'all', 'wrapped': False, 'interval': 5, 'scale': False,
                     'scale_to_max': False}, args)
        if self.settings['scale'] and self.settings['scale_to_max']:
            self.settings['scale'] = False
        self.boxes = []
        self.max_dim = None

    def run_batch(self, ba_settings, batch, ba_mda_data):
        pass                                  # everything happens in finish(), on the resident frames

    def finish(self, batch):
        _resident_only(batch, self.analysis_key)
        h = batch._analyzer._host_arrays()
        lat = list(batch.lateral)
        if self.settings['leaflet'] not in ('both', 'upper', 'lower'):
            warnings.warn("bad setting for \'leaflet\' in " + self.analysis_id + ". Using default \'both\'")
            self.settings['leaflet'] = 'both'
        for gp, gc in _interval_pairs(batch, batch._analyzer.settings['frame_range'][0], self.settings['interval']):
            leaflets = _frame_leaflets(batch, gc)
            if self.settings['leaflet'] == 'both':
                indices = []
                for name in leaflets:
                    indices += leaflets[name].get_group_indices(self.settings['resname'])
            else:
                indices = leaflets[self.settings['leaflet']].get_group_indices(self.settings['resname'])
            box_lateral = batch.box_host[gp][lat]
            self.boxes.append(box_lateral)
            rp, rc = batch.row_of_global(gp), batch.row_of_global(gc)
            com_i = h['com_unwrap'][rc][indices][:, lat]
            com_j = h['com_unwrap'][rp][indices][:, lat]
            com_j_w = h['com'][rp][indices][:, lat]
            if self.settings['scale']:
                com_i, com_j, com_j_w = com_i / box_lateral, com_j / box_lateral, com_j_w / box_lateral
            vec_ends = np.zeros((len(indices), 4))
            vec_ends[:, 0:2] = com_j_w if self.settings['wrapped'] else com_j
            vec_ends[:, 2:4] = com_i - com_j
            self.analysis_output.append([vec_ends, [batch.types[i] for i in indices]])

    def _process_output(self):
        if self.settings['scale_to_max']:
            boxes = np.array(self.boxes)
            max_x, max_y = boxes[:, 0].max(), boxes[:, 1].max()
            self.max_dim = [max_x, max_y]
            scaled_out = []
            for vec_ends, resnames in self.analysis_output:
                vec_ends = vec_ends.copy()
                vec_ends[:, 0] /= max_x
                vec_ends[:, 2] /= max_x
                vec_ends[:, 1] /= max_y
                vec_ends[:, 3] /= max_y
                scaled_out.append([vec_ends, resnames])
            return scaled_out
        return self.analysis_output

    def get_data(self):
        return self._process_output()

    def reset(self):
        self.analysis_output = []
        self.boxes = []
        self.max_dim = None


def binned_average(data, positions, n_bins=25, position_range=None, min_count=0):
    """(bin starts, averages) over the bins that received data; a position belongs to the FIRST bin whose
    closed interval [edge_j, edge_j+1] holds it (pybilt/common/running_stats.py:364-431)."""
    lower, upper = position_range if position_range is not None else (positions.min(), positions.max())
    edges = np.linspace(lower, upper, num=n_bins + 1, endpoint=True)
    bins = np.linspace(lower, upper, num=n_bins, endpoint=False)
    counts = np.zeros(len(bins)).astype(np.int64)
    sums = np.zeros(len(bins))
    for c_val, pos in zip(data, positions):
        hit = np.nonzero((pos >= edges[:-1]) & (pos <= edges[1:]))[0]
        if len(hit):
            counts[hit[0]] += 1
            sums[hit[0]] += c_val
    keep = np.nonzero(counts > 0)[0]
    return bins[keep], sums[keep] / counts[keep]


@_protocol('msd_dist', 'com_frame')
class MSDDistProtocol(AnalysisProtocol):
    """Squared lateral displacement over `interval` MD frames binned by the distance to the nearest
    lipid of type resname_2 (second nearest when resname_1 == resname_2) in the earlier frame
    (analysis_protocols.py:5523-5830).  The nearest distances of all frames come from one device search
    (`pbk_ct_nearest`, distance_euclidean_pbc on the wrapped COMs)."""
    _short_description = "Mean squared displacement as a function of distance."
    _related = list(['disp_vec', 'msd'])
    _casts = {'interval': int, 'n_bins': int, 'range_inner': float, 'range_outer': float, 'dist_avg': _truth}
    batched = True

    def __init__(self, args):
        self._setup({'leaflet': 'both', 'interval': 5, 'resname_1': 'first', 'resname_2': 'first', 'n_bins': 25,
                     'range_inner': 0.0, 'range_outer': 25.0, 'dist_avg': False}, args)
        self.reset()

    def reset(self):
        self._sds, self._distances = [], []
        self._r1_indices, self._r2_indices = [], []
        self.analysis_output = []
        self._processed_output = None

    def run_batch(self, ba_settings, batch, ba_mda_data):
        pass

    def finish(self, batch):
        _resident_only(batch, self.analysis_key)
        st, leaflets = self.settings, batch.first_leaflets
        do_leaflet = ['upper', 'lower'] if st['leaflet'] == 'both' else [st['leaflet']]
        lipid_types, _n = _lipid_types_in(leaflets, do_leaflet)
        if st['resname_1'] == 'first':
            st['resname_1'] = lipid_types[0]
        if st['resname_2'] == 'first':
            st['resname_2'] = lipid_types[0]
        r1, r2 = [], []
        for name in (leaflets if st['leaflet'] == 'both' else [st['leaflet']]):
            r1 += leaflets[name].get_group_indices(st['resname_1'])
            r2 += leaflets[name].get_group_indices(st['resname_2'])
        self._r1_indices, self._r2_indices = r1, r2
        lat = list(batch.lateral)
        _pos, d_first, d_second = batch.engine.ct_nearest(r1, r2, lat, d1=True, frames=batch.rows, second=True)
        dist = (d_second if st['resname_1'] == st['resname_2'] else d_first).cpu().numpy()
        h = batch._analyzer._host_arrays()
        # the reference keys its neighbour table by lipid index: a lipid listed twice keeps one entry
        keys = list(dict.fromkeys(r1))
        col = {i: k for k, i in enumerate(r1)}
        for gp, gc in _interval_pairs(batch, batch._analyzer.settings['frame_range'][0], st['interval']):
            rp, rc = batch.row_of_global(gp), batch.row_of_global(gc)
            disp = h['com_unwrap'][rc][r1][:, lat] - h['com_unwrap'][rp][r1][:, lat]
            # np.dot of two float64: fused accumulation (DESIGN.md "Arithmetic"); x*x + y*y agrees with it
            # to one rounding, far below the 1e-9 of the binned averages
            self._sds += list(disp[:, 0] * disp[:, 0] + disp[:, 1] * disp[:, 1])
            for i in keys:
                d = dist[gp - batch.gframes.start][col[i]]
                if st['dist_avg']:
                    d = (d + dist[gc - batch.gframes.start][col[i]]) / 2.0
                self._distances.append(d)

    def _process_output(self):
        if self._processed_output is None:
            self._processed_output = binned_average(np.array(self._sds), np.array(self._distances),
                                                    n_bins=self.settings['n_bins'],
                                                    position_range=[self.settings['range_inner'], self.settings['range_outer']],
                                                    min_count=0)

    def get_data(self):
        self._process_output()
        return self._processed_output


@_protocol('dc_cluster', 'com_frame')
class DCClusterProtocol(AnalysisProtocol):
    """Distance cut-off clusters of one lipid type per frame: {'clusters': per frame lists of positions in
    the frame's index list, 'nclusters' / 'max_size' / 'min_size' / 'avg_size': rows [time, value, running
    mean, running deviation]} (analysis_protocols.py:2826-2983, distance_cutoff_clustering.py:127-205).
    The cut-off matrix comes from the device (`pbk_ct_adjacency`); clusters are grown on the host in the
    reference's order."""
    _short_description = "Hiearchical clustering of lipids based on distance."
    _related = None
    _casts = {'cutoff': float}
    batched = True

    def __init__(self, args):
        self._setup({'resname': 'first', 'leaflet': 'upper', 'cutoff': 12.0}, args)
        self.reset()

    def reset(self):
        self.converted = False
        self.analysis_output = {k: [] for k in ('clusters', 'nclusters', 'max_size', 'min_size', 'avg_size')}
        self.running_stats = {k: RunningStats() for k in ('nclusters', 'max_size', 'min_size', 'avg_size')}

    def run_batch(self, ba_settings, batch, ba_mda_data):
        pass

    def finish(self, batch):
        _resident_only(batch, self.analysis_key)
        st = self.settings
        do_leaflet = ['upper', 'lower'] if st['leaflet'] == 'both' else [st['leaflet']]
        if st['resname'] == 'first':
            st['resname'] = _lipid_types_in(batch.first_leaflets, do_leaflet)[0][0]
        # index list per frame (it follows the frame's leaflets); frames that share a list share a launch
        lists = []
        for g in range(batch.n_frames):
            leaflets = _frame_leaflets(batch, g)
            idx = []
            for name in do_leaflet:
                idx += leaflets[name].get_group_indices(st['resname'])
            lists.append(idx)
        g = 0
        while g < batch.n_frames:
            g1 = g + 1
            while g1 < batch.n_frames and lists[g1] == lists[g]:
                g1 += 1
            idx, n = lists[g], len(lists[g])
            rows = slice(batch.row_of_global(g), batch.row_of_global(g1))
            bits = batch.engine.ct_adjacency(idx, st['cutoff'], batch.lateral, d1=True, frames=rows).cpu().numpy() if n else None
            for k in range(g1 - g):
                adj = np.unpackbits(bits[k].view(np.uint8), axis=1, bitorder='little')[:, :n].astype(bool) if n else None
                self._frame(batch.times[g + k], adj, n)
            g = g1

    def _frame(self, time, adj, n):
        clusters, master = [], list(range(n))
        while master:
            neighbors, taken = [master[0]], {master[0]}
            i = 0
            while i < len(neighbors):
                row = adj[neighbors[i]]
                for b in master:
                    if b not in taken and row[b]:
                        neighbors.append(b)
                        taken.add(b)
                i += 1
            master = [b for b in master if b not in taken]
            if len(neighbors) > 1:
                clusters.append(list(neighbors))
        nclusters, min_size, max_size, avg_size = len(clusters), 0, 0, 0.0
        for cluster in clusters:                  # (the reference's running min / max, as written)
            m = len(cluster)
            if m > max_size:
                max_size = m
            elif m > 0 and max_size == 0:
                min_size = m
            elif max_size > 0 and m < max_size:
                min_size = m
            avg_size += m
        if nclusters > 0:
            avg_size /= nclusters
        out, rs = self.analysis_output, self.running_stats
        out['clusters'].append(clusters)
        for key, val in (('nclusters', nclusters), ('max_size', max_size), ('min_size', min_size), ('avg_size', avg_size)):
            rs[key].push(val)
            out[key].append([time, val, rs[key].mean(), rs[key].deviation()])

    def get_data(self):
        if not self.converted:
            for key in self.analysis_output.keys():
                if key != 'clusters':
                    self.analysis_output[key] = np.array(self.analysis_output[key])
            self.converted = True
        return self.analysis_output


# ---------------------------------------------------------------------------------------
@_protocol('halperin_nelson', 'com_frame')
class HalperinNelsonProtocol(AnalysisProtocol):
    """Halperin and Nelson's hexagonal-packing invariant averaged over a leaflet, rows
    [time, mean phi] (analysis_protocols.py:4286-4440).  The lipid list is the first frame's
    leaflet ('upper' unless 'lower' is asked for)."""
    _short_description = "Halperin and Nelson's rotational invariant."
    _related = None
    batched = True

    def __init__(self, args):
        self._indices = []
        self._setup({'leaflet': 'upper'}, args)

    def run_batch(self, ba_settings, batch, ba_mda_data):
        if batch.first_window:
            leaf = self.settings['leaflet'] if self.settings['leaflet'] in ('upper', 'lower') else 'upper'
            self._indices = batch.first_leaflets[leaf].get_member_indices()
            mask = np.zeros(len(batch.types), dtype=np.uint8)
            mask[np.asarray(self._indices, dtype=np.int64)] = 1
            self._member_dev = torch.from_numpy(mask).to(batch.engine.dev)
            self._vals = np.zeros(batch.n_frames)
            self._pending = []
        self._pending.append((batch.gframes,
                              batch.engine.halperin_nelson(self._member_dev, ba_settings['lateral'],
                                                           frames=batch.rows)))

    def finish(self, batch):
        for g, v in self._pending:
            self._vals[g] = v.cpu().numpy()
        self._pending = []
        vals = batch.allreduce(self._vals)
        self.analysis_output = [[t, v] for t, v in zip(np.asarray(batch.times, dtype=np.float64), vals)]

    def reset(self):
        self.analysis_output = []
        self._indices = []


# ---------------------------------------------------------------------------------------
@_protocol('apl_box', 'mda_frame')
class APLBoxProtocol(AnalysisProtocol):
    """Area per lipid from the lateral box area, rows [time, 2 A_xy / N_lipids, running mean,
    running deviation] (analysis_protocols.py:945-1012).  Box arithmetic only (float32, like the
    trajectory's dimensions), done on the host in finish()."""
    _short_description = "Area per lipid using box dimensions."
    _related = list(['apl_grid'])
    batched = True

    def __init__(self, args):
        self._setup({'none': None}, args)
        self.running = RunningStats()

    def run_batch(self, ba_settings, batch, ba_mda_data):
        self._n_residues = ba_mda_data.n_residues

    def finish(self, batch):
        lat = list(batch.lateral)
        for f in range(batch.n_frames):
            plane = batch.box_host[f][lat]
            apl = plane[0] * plane[1] * 2.0 / self._n_residues
            self.running.push(apl)
            self.analysis_output.append(np.array([batch.times[f], apl, self.running.mean(),
                                                  self.running.deviation()]))

    def reset(self):
        self.running.reset()
        self.analysis_output = []


@_protocol('acm', 'mda_frame')
class AreaCompressibilityModulusProtocol(AnalysisProtocol):
    """Area compressibility modulus from the fluctuation of the lateral box area, rows
    [time, Ka, Ka error, area] (analysis_protocols.py:3100-3210).  Host-side box arithmetic."""
    _short_description = "Area compressibility modulus."
    _related = list(['area_fluctuation', 'ac'])
    _casts = {'temperature': float}
    batched = True

    def __init__(self, args):
        self._setup({'temperature': 298.15}, args)
        self.n_frames = 0
        self.area_run = RunningStats()
        self.first_comp = True

    def run_batch(self, ba_settings, batch, ba_mda_data):
        self._norm = ba_settings['norm']

    def finish(self, batch):
        import scipy.constants as scicon
        with np.errstate(all="ignore"):
            for f in range(batch.n_frames):
                d = batch.box_host[f]
                area = d.prod() / d[self._norm]
                self.area_run.push(area)
                ka, ka_err = 1.0, 0.0
                if self.first_comp:
                    self.first_comp = False
                    self.n_frames += 1
                    ka = 0.0
                else:
                    ka = (self.area_run.mean() * scicon.k * self.settings['temperature']) / self.area_run.variance()
                    sd, mean, n = self.area_run.deviation(), self.area_run.mean(), self.area_run.n
                    var_of_var = (np.sqrt(2.0) * sd ** 2) / np.sqrt(n - 1)
                    sd_of_var = np.sqrt(var_of_var)
                    ka_err = np.sqrt((sd / mean) ** 2 + (sd_of_var / sd ** 2) ** 2) * ka
                ka *= 10.0 ** 23
                ka_err *= 10.0 ** 23
                self.analysis_output.append([batch.times[f], ka, ka_err, area])
                self.n_frames += 1

    def reset(self):
        self.n_frames = 0
        self.area_run.reset()
        self.analysis_output = []
        self.first_comp = True


@_protocol('ac', 'mda_frame')
class AreaCompressibilityProtocol(AnalysisProtocol):
    """Isothermal area compressibility, rows [time, var(A) / (<A> k T) * 1e-23]
    (analysis_protocols.py:3349-3450).  Host-side box arithmetic."""
    _short_description = "Isothermal area compressibility."
    _related = list(['acm', 'vcm'])
    _casts = {'temperature': float}
    batched = True

    def __init__(self, args):
        self._setup({'temperature': 298.15}, args)
        self.n_frames = 0
        self.area_run = RunningStats()
        self.area_fluctuation = RunningStats()

    def run_batch(self, ba_settings, batch, ba_mda_data):
        self._norm = ba_settings['norm']

    def finish(self, batch):
        import scipy.constants as scicon
        with np.errstate(all="ignore"):
            for f in range(batch.n_frames):
                d = batch.box_host[f]
                area = d.prod() / d[self._norm]
                self.area_run.push(area)
                area_mean = self.area_run.mean()
                self.area_fluctuation.push((area - area_mean) ** 2)
                x_t = self.area_run.variance() / (area_mean * scicon.k * self.settings['temperature'])
                x_t *= 10.0 ** (-23)
                self.analysis_output.append([batch.times[f], x_t])
                self.n_frames += 1

    def reset(self):
        self.n_frames = 0
        self.area_run.reset()
        self.area_fluctuation.reset()
        self.analysis_output = []


@_protocol('vcm', 'mda_frame')
class VolumeCompressibilityModulusProtocol(AnalysisProtocol):
    """Volume compressibility modulus, rows [time, <V> k T / dev(V)^2] (0 on the first frame)
    (analysis_protocols.py:2997-3092).  Host-side box arithmetic."""
    _short_description = "Volume compressibility modulus."
    _related = list(['acm', 'ac'])
    _casts = {'temperature': float}
    batched = True

    def __init__(self, args):
        self._setup({'temperature': 298.15}, args)
        self.n_frames = 0
        self.volume_run = RunningStats()
        self.first_comp = True

    def run_batch(self, ba_settings, batch, ba_mda_data):
        pass

    def finish(self, batch):
        import scipy.constants as scicon
        with np.errstate(all="ignore"):
            for f in range(batch.n_frames):
                self.volume_run.push(batch.box_host[f].prod())
                if self.first_comp:
                    self.first_comp = False
                    self.n_frames += 1
                    kv = 0.0
                else:
                    kv = (self.volume_run.mean() * scicon.k * self.settings['temperature']) / \
                        self.volume_run.deviation() ** 2
                self.analysis_output.append([batch.times[f], kv])
                self.n_frames += 1

    def reset(self):
        self.n_frames = 0
        self.volume_run.reset()
        self.analysis_output = []
        self.first_comp = True


# ---------------------------------------------------------------------------------------
@_protocol('area_fluctuation', 'com_frame')
class AreaFluctuationProtocol(AnalysisProtocol):
    """Fluctuation of the lateral box area, rows [time, area, sqrt(<A^2> - <A>^2)]
    (analysis_protocols.py:4452-4510).  Box arithmetic only, O(frames): done on the host in
    finish(), in the reference's float32 / Welford order."""
    _short_description = "Bilayer lateral box area fluctuation."
    _related = list(['acm', 'ac'])
    batched = True

    def __init__(self, args):
        self._setup({'none': None}, args)
        self._first_frame = True
        self._area_run = RunningStats()
        self._area_sq_run = RunningStats()

    def run_batch(self, ba_settings, batch, ba_mda_data):
        pass                                  # nothing to do per window: the boxes are host metadata

    def finish(self, batch):
        lat = list(batch.lateral)
        with np.errstate(invalid="ignore"):
            for f in range(batch.n_frames):
                area = batch.box_host[f][lat].prod()          # float32 product (com_frame.box)
                self._area_run.push(area)
                self._area_sq_run.push(area * area)
                if self._first_frame:
                    self.analysis_output.append([batch.times[f], area, 0.0])
                    self._first_frame = False
                else:
                    fluct = np.sqrt(self._area_sq_run.mean() - self._area_run.mean() ** 2)
                    self.analysis_output.append([batch.times[f], area, fluct])

    def reset(self):
        self._first_frame = True
        self.analysis_output = []
        self._area_run.reset()
        self._area_sq_run.reset()


# ---------------------------------------------------------------------------------------
@_protocol('thickness_grid', 'lipid_grid')
class BTGridProtocol(AnalysisProtocol):
    """Bilayer thickness from the lipid grids, rows [time, thickness, running mean, running
    deviation] (analysis_protocols.py:1026-1095)."""
    _short_description = "Bilayer thickness using lipid_grid."
    _related = list(['lipid_length', 'thickness_leaflet_distance', 'thickness_peak_distance',
                     'thickness_mass_weighted_std'])
    batched = True

    def __init__(self, args):
        self._setup({'none': None}, args)
        self.running = RunningStats()

    def run_batch(self, ba_settings, batch, ba_mda_data):
        if batch.first_window:
            self._th = np.zeros(batch.n_frames)
        self._th[batch.gframes] = batch.lipid_grids().thickness_host[:, 0]

    def finish(self, batch):
        rows, self.running = running_rows(batch.times, batch.allreduce(self._th))
        self.analysis_output = [r for r in rows]

    def reset(self):
        self.running.reset()
        self.analysis_output = []


@_protocol('apl_grid', 'lipid_grid')
class APLGridProtocol(AnalysisProtocol):
    """Area per lipid from the lipid grids -> {'composite': [F,4], resname: [F,4]}
    (analysis_protocols.py:1549-1644)."""
    _short_description = "Area per lipid using 2D lipid grids."
    _related = list(['apl_box'])
    batched = True

    def __init__(self, args):
        self._setup({'none': None}, args)
        self.analysis_output = {}
        self.running = RunningStats()
        self.running_res = {}
        self.first_comp = True

    def run_batch(self, ba_settings, batch, ba_mda_data):
        grids = batch.lipid_grids()
        if batch.first_window:
            self._apl = np.zeros((batch.n_frames, grids.apl_host.shape[1]))
        self._apl[batch.gframes] = grids.apl_host

    def finish(self, batch):
        apl = batch.allreduce(self._apl)
        rows, self.running = running_rows(batch.times, apl[:, 0])
        out = {'composite': rows}
        # resnames in the order LipidGrids.area_per_lipid lists them at the first frame
        names, _n = _lipid_types_in(batch.first_leaflets, ['upper', 'lower'])
        for name in names:
            t = batch.type_names.index(name)
            out[name], self.running_res[name] = running_rows(batch.times, apl[:, 1 + 2 * t])
        self.analysis_output = out
        self.first_comp = False

    def get_data(self):
        for key in self.analysis_output.keys():
            self.analysis_output[key] = np.array(self.analysis_output[key])
        return self.analysis_output

    def reset(self):
        self.running.reset()
        self.analysis_output = {}
        self.first_comp = True
        self.running_res = {}
