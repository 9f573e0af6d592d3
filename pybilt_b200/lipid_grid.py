"""``LipidGrids`` / ``LipidGrid2d``: host views of the device-built GridMAT-style lipid grids.

Mirror of the objects pybilt/lipid_grid/lipid_grid_opt.py:155-334,470-597 builds per frame
(what ``BilayerAnalyzer`` stores in ``reps['lipid_grid']``).  The cell -> lipid assignment,
the cell counts per lipid and the thickness / area-per-lipid reductions are computed by
libpbk (``pbk_lipid_grid`` / ``pbk_grid_reduce``); this module only exposes them with the
reference's attribute and method names.
"""
from __future__ import annotations

import numpy as np


def grid_geometry(box_len, nbins):
    """Edges, increment and centres as lipid_grid_opt.py:204-227 computes them under
    NumPy >= 2 (float32 linspace)."""
    edges = np.linspace(0.0, box_len, nbins + 1, endpoint=True)
    incr = edges[1] - edges[0]
    centres = np.zeros(nbins)
    for j in range(nbins):
        centres[j] = edges[j] + incr / 2.0
    return edges, incr, centres


def fixed_resolution_bins(boxx, boxy, area_per_bin):
    """lipid_grid_opt.py:192-198."""
    nbins = np.sqrt((boxx * boxy) / area_per_bin)
    x2y_sr = np.sqrt(boxx / boxy)
    return int(nbins * x2y_sr), int(nbins / x2y_sr)


class LipidGrid2d(object):
    def __init__(self, grid_idx, grid_z, boxx, boxy):
        nx, ny = grid_idx.shape
        self.x_nbins, self.y_nbins = nx, ny
        self.x_min, self.x_max = 0.0, boxx
        self.y_min, self.y_max = 0.0, boxy
        self.x_edges, self.x_incr, self.x_centers = grid_geometry(boxx, nx)
        self.y_edges, self.y_incr, yc = grid_geometry(boxy, ny)
        self.y_centers = np.zeros(ny)
        m = min(nx, ny)                              # lipid_grid_opt.py:225 (x edge count)
        self.y_centers[:m] = yc[:m]
        self.x_length = self.x_max - self.x_min
        self.y_length = self.y_max - self.y_min
        self.lipid_grid = np.asarray(grid_idx, dtype=np.int64)
        self.lipid_grid_z = np.asarray(grid_z, dtype=np.float64)

    def get_index_at(self, ix, iy):
        return self.lipid_grid[ix, iy]

    def get_z_at(self, ix, iy):
        return self.lipid_grid_z[ix, iy]


class LipidGrids(object):
    def __init__(self, com_frame, leaflets, plane, grid_idx, grid_z, thickness, apl_row, cells,
                 type_names):
        self.frame = com_frame
        self.leaflets = leaflets
        self.plane = plane
        self.norm = [i for i in [0, 1, 2] if i not in plane][0]
        self.nbins_x, self.nbins_y = grid_idx.shape[1], grid_idx.shape[2]
        self.myframe = com_frame.mdnumber if com_frame is not None else 0
        box = com_frame.box[list(plane)]
        self.leaf_grid = {'upper': LipidGrid2d(grid_idx[0], grid_z[0], box[0], box[1]),
                          'lower': LipidGrid2d(grid_idx[1], grid_z[1], box[0], box[1])}
        self._thickness = thickness
        self._apl_row = apl_row
        self._cells = cells
        self._type_names = type_names

    def thickness_grid(self):
        return self.leaf_grid['upper'].lipid_grid_z - self.leaf_grid['lower'].lipid_grid_z

    def average_thickness(self, return_grid=False):
        avg_out = (self._thickness[0], self._thickness[1])
        if return_grid:
            return avg_out, self.thickness_grid()
        return avg_out

    def area_per_lipid(self):
        """(system average, {resname: (mean, deviation)}, {lipid index: area})."""
        resnames = []
        for leaf in ('upper', 'lower'):
            for g in self.leaflets[leaf].get_group_names():
                if g not in resnames:
                    resnames.append(g)
        per_res = {}
        for name in resnames:
            t = self._type_names.index(name)
            per_res[name] = (self._apl_row[1 + 2 * t], self._apl_row[2 + 2 * t])
        area = {}
        for leaf in ('upper', 'lower'):
            g = self.leaf_grid[leaf]
            apb = g.x_incr * g.y_incr
            for i in self.leaflets[leaf].get_member_indices():
                area[i] = apb * int(self._cells[i])
        return self._apl_row[0], per_res, area

    def map_to_grid(self, com_values_dict, leaflet='both'):
        do_leaflet = ['upper', 'lower'] if leaflet not in ('upper', 'lower') else [leaflet]
        out = {}
        for leaf in do_leaflet:
            idx = self.leaf_grid[leaf].lipid_grid
            out[leaf] = np.vectorize(lambda i: com_values_dict[i], otypes=[float])(idx)
        return out
