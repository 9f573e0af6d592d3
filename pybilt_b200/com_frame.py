"""``LipidCOM`` / ``COMFrame``: the centre-of-mass representation handed to plugins.

Mirror of pybilt/bilayer_analyzer/com_frame.py:13-453.  In the reference a COMFrame is
*built* from an MDAnalysis frame (N ``center_of_mass()`` calls, twice); here the numbers
come from the device (libpbk: unwrap -> membrane COM -> lipid COM) and this class is a host
view of one analysed frame with the same attributes and accessors.
"""
from __future__ import annotations

import numpy as np


class LipidCOM(object):
    def __init__(self):
        self.type = "UNK"
        self.com = np.zeros(3)
        self.com_unwrap = np.zeros(3)
        self.mass = 1.0
        self.leaflet = "UNK"
        self.resid = 0


class COMFrame(object):
    def __init__(self, com, com_unwrap, masses, types, resids, box, time, number, mem_com,
                 leaflet_names=None):
        self.lipidcom = []
        for i in range(len(types)):
            lc = LipidCOM()
            lc.type = types[i]
            lc.com = np.array(com[i])
            lc.com_unwrap = np.array(com_unwrap[i])
            lc.mass = masses[i]
            lc.resid = int(resids[i])
            if leaflet_names is not None:
                lc.leaflet = leaflet_names[i]
            self.lipidcom.append(lc)
        self.box = np.asarray(box, dtype=np.float32)       # com_frame.py:131
        self.time = time
        self.mdnumber = number
        self.number = number
        self.mem_com = np.array(mem_com)

    def __repr__(self):
        return 'COMFrame for frame %s with %s lipids' % (self.number, len(self.lipidcom))

    def __len__(self):
        return len(self.lipidcom)

    def set_box(self, box_lengths):
        self.box = box_lengths

    def set_time(self, time):
        self.time = time

    def com(self, wrapped=True):
        out = np.zeros(3)
        total = 0.0
        for lipid in self.lipidcom:
            out += (lipid.com if wrapped else lipid.com_unwrap) * lipid.mass
            total += lipid.mass
        out /= total
        return out

    def cog(self, wrapped=True):
        out = np.zeros(3)
        total = 0.0
        for lipid in self.lipidcom:
            out += lipid.com if wrapped else lipid.com_unwrap
            total += 1.0
        out /= total
        return out

    def coordinates(self, wrapped=True, leaflet=None):
        return np.array([(l.com if wrapped else l.com_unwrap) for l in self.lipidcom
                         if leaflet is None or l.leaflet == leaflet])

    def masses(self):
        return np.array([l.mass for l in self.lipidcom])

    def resids(self):
        return np.array([l.resid for l in self.lipidcom])

    def resnames(self):
        return [l.type for l in self.lipidcom]

    def leaflets(self):
        return [l.leaflet for l in self.lipidcom]

    def unique_resnames(self):
        out = []
        for l in self.lipidcom:
            if l.type not in out:
                out.append(l.type)
        return out

    def write_xyz(self, xyz_name, wrapped=True, name_by_leaflet=False):
        with open(xyz_name, "w") as fh:
            fh.write("%d\n" % len(self.lipidcom))
            fh.write("COMFrame " + str(self.number) + " MD Frame " + str(self.mdnumber) + "\n")
            for l in self.lipidcom:
                xy = l.com if wrapped else l.com_unwrap
                name = l.leaflet if name_by_leaflet else l.type
                fh.write("%s %s %s %s\n" % (name, xy[0], xy[1], l.com_unwrap[2]))
