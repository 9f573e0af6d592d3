"""``Leaflet`` / ``LipidGroup``: ordered index sets of one bilayer leaflet.

Mirror of pybilt/bilayer_analyzer/leaflet.py:9-268 with the same public methods; instances
are built from a row of device-computed leaflet labels instead of lipid by lipid.
"""
from __future__ import annotations

import numpy as np


_TYPE_TABLES = {}


def _type_table(types):
    """(object array of the resnames, int code per lipid, code -> resname) of a resname list; kept
    per list object, since the same topology is asked for every run (10^5 strings are slow to
    re-encode)."""
    key = id(types)
    hit = _TYPE_TABLES.get(key)
    if hit is not None and hit[0] is types:
        return hit[1]
    tarr = np.asarray(types, dtype=object)
    uniq, inv = np.unique(tarr.astype(str), return_inverse=True)
    names = [str(u) for u in uniq]
    # keep the caller's own string objects as names where possible
    first = {}
    for i, code in enumerate(inv.tolist()):
        if code not in first:
            first[code] = tarr[i]
            if len(first) == len(names):
                break
    names = [first[k] for k in range(len(names))]
    table = (tarr, inv.astype(np.int32), names)
    if len(_TYPE_TABLES) > 8:
        _TYPE_TABLES.clear()
    _TYPE_TABLES[key] = (types, table)
    return table


class LipidGroup(object):
    def __init__(self, name):
        self.lg_members = []
        self.lg_name = name

    def add_member(self, new_mem):
        self.lg_members.append(new_mem)

    def name(self):
        return self.lg_name


class Leaflet(object):
    def __init__(self, name):
        self.name = name
        self._members = []          # [com_index, resname, resid]
        self._lazy = None           # (indices, resnames, resids) arrays not yet turned into `members`
        self.groups = []
        self.group_dict = {}

    @property
    def members(self):
        """[com_index, resname, resid] per member (leaflet.py:58-70); built on first use when the
        leaflet came from label arrays (a Python list of 10^5 small lists is slow to make)."""
        if self._lazy is not None:
            idx, t, r = self._lazy
            self._lazy = None
            self._members = [list(m) for m in zip(idx.tolist(), t.tolist(), [int(v) for v in r.tolist()])]
        return self._members

    @members.setter
    def members(self, value):
        self._lazy = None
        self._members = value

    @classmethod
    def from_arrays(cls, name, indices, types, resids):
        """Members in ascending lipid index; groups in first-seen order (leaflet.py:58-91)."""
        leaf = cls(name)
        idx = np.asarray(indices, dtype=np.int64)
        if len(idx) == 0:
            return leaf
        tarr, codes, names = _type_table(types)
        r = np.asarray(resids)[idx]
        leaf._lazy = (idx, tarr[idx], r)
        # groups in first-seen order, members in ascending lipid index (what add_member builds)
        c = codes[idx]
        present, first = np.unique(c, return_index=True)
        for k in np.argsort(first, kind="stable"):
            nm = names[int(present[k])]
            g = LipidGroup(nm)
            g.lg_members = idx[c == present[k]].tolist()
            leaf.group_dict[nm] = len(leaf.groups)
            leaf.groups.append(g)
        return leaf

    def __str__(self):
        return '%s leaflet of a Membrane System with %s members and %s lipid groups' % (
            self.name, len(self.members), len(self.groups))

    __repr__ = __str__

    def __len__(self):
        return len(self._lazy[0]) if self._lazy is not None else len(self._members)

    def add_member(self, com_index, resname, resid):
        self.members.append([com_index, resname, resid])
        if resname not in self.group_dict:
            self.group_dict[resname] = len(self.groups)
            self.groups.append(LipidGroup(resname))
        self.groups[self.group_dict[resname]].add_member(com_index)

    def get_member_indices(self):
        if self._lazy is not None:
            return self._lazy[0].tolist()
        return [m[0] for m in self.members]

    def get_member_resids(self):
        if self._lazy is not None:
            return [int(v) for v in self._lazy[2].tolist()]
        return [m[2] for m in self.members]

    def get_member_resnames(self):
        if self._lazy is not None:
            return self._lazy[1].tolist()
        return [m[1] for m in self.members]

    def get_group_indices(self, group_name):
        if group_name == "all":
            return self.get_member_indices()
        if group_name in self.group_dict:
            return list(self.groups[self.group_dict[group_name]].lg_members)
        # leaflet.py:112-116: unknown group -> warning + every member
        print("!! Warning - request for unknown Lipid Group \'", group_name, "\' from the ",
              self.name, " leaflet")
        print("!! using the default \"all\"")
        return self.get_member_indices()

    def get_group_indices_per_resid(self, group_name):
        if group_name != "all" and group_name not in self.group_dict:
            return self.get_group_indices(group_name)
        out = {}
        for i in self.get_group_indices(group_name):
            out.setdefault(self.get_member_resid_from_index(i), []).append(i)
        return out

    def get_member_resname_from_resid(self, resid):
        resids = self.get_member_resids()
        if resid in resids:
            return self.get_member_resnames()[resids.index(resid)]
        raise ValueError('resid is not in this leaflet')

    def get_member_resid_from_index(self, index):
        indices = self.get_member_indices()
        if index in indices:
            return self.get_member_resids()[indices.index(index)]
        raise ValueError('resid is not in this leaflet')

    def has_group(self, group_name):
        return group_name == 'all' or group_name in self.group_dict

    def num_groups(self):
        return len(self.groups)

    def get_group_names(self):
        return [g.lg_name for g in self.groups]


def leaflets_from_labels(label_row, types, resids):
    """{'upper': Leaflet, 'lower': Leaflet} from one uint8 label row (1 = upper)."""
    label_row = np.asarray(label_row)
    return {'upper': Leaflet.from_arrays('upper', np.nonzero(label_row == 1)[0], types, resids),
            'lower': Leaflet.from_arrays('lower', np.nonzero(label_row == 0)[0], types, resids)}
