"""Prefab drivers for the accelerated analyses.

Mirror of the top-level user calls in pybilt/bilayer_analyzer/prefab_analysis_protocols.py that are
built on the analyses of this package: ``msd_diffusion`` (:34), ``area_per_lipid`` (:313),
``nearest_neighbor_fraction`` (:834) and ``com_lateral_rdf`` (:1463).  Same arguments, same analysis /
plot strings handed to the analyzer, same summary lines on standard output and the same pickles from
``dump_data``; the frames are reduced on the device by ``pybilt_b200.BilayerAnalyzer``.  Figures are
written through the analyzer's plot protocols (series files when matplotlib is absent); the drivers'
extra matplotlib figures (``plot_generation_functions``) are not rebuilt.
"""
from __future__ import annotations

import numpy as np

from .bilayer_analyzer import BilayerAnalyzer
from . import diffusion as dc
from .running_stats import BlockAverager


def _report_estimators(label, times, msd_vals):
    """The three estimators over the whole series, the first 10 ns and (if present) 10-100 ns."""
    t0 = times[0]
    t10, t100 = t0 + 10000.0, t0 + 100000.0
    ranges = [("whole time range", None), ("0-10 ns", [t0, t10])]
    if max(times) > t10:
        ranges.append(("10-100 ns", [t10, t100]))
    for name, tr in ranges:
        kw = {} if tr is None else {'time_range': tr}
        D_e = dc.diffusion_coefficient_Einstein(times, msd_vals, **kw)
        D_l = dc.diffusion_coefficient_linear_fit(times, msd_vals, **kw)
        D_a = dc.diffusion_coefficient_anomalous_fit(times, msd_vals, **kw)
        print(label + "; " + name)
        print("Values from estimators:")
        print("  Basic Einstein relation:")
        print(("    Diffusion coefficient: {}".format(D_e)))
        print("  Linear fit:")
        print(("    Diffusion coefficient: {} Std Error: {}".format(D_l[0], D_l[1])))
        print("  Anomalous diffusion fit:")
        print(("    Diffusion coefficient: {} Alpha value: {}".format(D_a[0], D_a[1])))


def msd_diffusion(structure_file, trajectory_file, selection_string, resnames=None, frame_start=0,
                  frame_end=-1, frame_interval=1, dump_path=None, **analyzer_kwargs):
    """MSD curves (all lipids; per listed resname: both / upper / lower leaflet) and the diffusion
    coefficients the three estimators give for them."""
    analyzer = BilayerAnalyzer(structure=structure_file, trajectory=trajectory_file,
                               selection=selection_string, **analyzer_kwargs)
    analyzer.set_frame_range(frame_start, frame_end, frame_interval)
    resnames = resnames or []
    for lipid_type in resnames:
        a_name = "msd_" + lipid_type
        analyzer.add_analysis("msd " + a_name + " resname " + lipid_type)
        analyzer.add_analysis("msd " + a_name + "_upper resname " + lipid_type + " leaflet upper")
        analyzer.add_analysis("msd " + a_name + "_lower resname " + lipid_type + " leaflet lower")
    analyzer.add_plot('msd msd_a msd_1 All')
    l_msd_string = "msd msd_l"
    for lipid_type in resnames:
        l_msd_string += " msd_" + lipid_type + " " + lipid_type
    analyzer.add_plot(l_msd_string)
    for lipid_type in resnames:
        analyzer.add_plot("msd " + lipid_type + "_cul msd_" + lipid_type + " Composite msd_" + lipid_type +
                          "_upper Upper msd_" + lipid_type + "_lower Lower")
    analyzer.print_analysis_protocol()
    analyzer.print_plot_protocol()
    analyzer.run_analysis()
    analyzer.dump_data(path=dump_path)
    analyzer.save_all_plots()
    msd_dat = analyzer.get_analysis_data('msd_1')
    _report_estimators("Composite for all lipids and both leaflets", msd_dat[:, 0], msd_dat[:, 1])
    for lipid_type in resnames:
        msd_dat = analyzer.get_analysis_data("msd_" + lipid_type)
        _report_estimators("Composite for " + lipid_type + " and both leaflets", msd_dat[:, 0], msd_dat[:, 1])
        for leaflet in ['upper', 'lower']:
            msd_dat = analyzer.get_analysis_data("msd_" + lipid_type + "_" + leaflet)
            _report_estimators("Composite for " + lipid_type + " and " + leaflet + " leaflet",
                               msd_dat[:, 0], msd_dat[:, 1])
    return


def area_per_lipid(structure_file, trajectory_file, selection_string, frame_start=0, frame_end=-1,
                   frame_interval=1, dump_path=None, name_dict=None, n_xbins=100, n_ybins=100,
                   **analyzer_kwargs):
    """Area per lipid from the box and from the lipid grids, with running and block averages."""
    analyzer = BilayerAnalyzer(structure=structure_file, trajectory=trajectory_file,
                               selection=selection_string, **analyzer_kwargs)
    analyzer.set_frame_range(frame_start, frame_end, frame_interval)
    analyzer.remove_analysis('msd_1')
    analyzer.rep_settings['com_frame']['name_dict'] = name_dict
    analyzer.add_analysis("apl_box apl_box")
    analyzer.add_analysis("apl_grid apl_grid")
    analyzer.print_analysis_protocol()
    analyzer.add_plot("apl apl_box apl_box None")
    analyzer.add_plot("apl apl_p apl_box Box apl_grid None")
    analyzer.add_plot("apl apl_grid apl_grid None")
    analyzer.rep_settings['lipid_grid']['n_xbins'] = n_xbins
    analyzer.rep_settings['lipid_grid']['n_ybins'] = n_ybins
    analyzer.print_plot_protocol()
    analyzer.run_analysis()
    analyzer.dump_data(path=dump_path)
    analyzer.save_all_plots()
    apl_box = analyzer.get_analysis_data('apl_box')
    print("Final running average Area per lipid estimates (squared Angstrom): ")
    print(("  via the box dimensions: {:0.4f} +- {:0.4f}".format(apl_box[-1][2], apl_box[-1][3])))
    print("  via the gridding procedure: ")
    apl_grid = analyzer.get_analysis_data('apl_grid')
    for item in apl_grid.keys():
        print(("    {}: {:0.4f} +- {:0.4f}".format(item, apl_grid[item][-1][2], apl_grid[item][-1][3])))
    ppb = len(apl_box) / 10
    if ppb > 1000:
        ppb = 1000
    n_b = 9
    while ppb < 10:
        ppb = len(apl_box) / n_b
        n_b -= 1
        if n_b == 0:
            ppb = len(apl_box)
            break
    print(("Will do block average with {} points per block".format(ppb)))
    block_averager = BlockAverager(points_per_block=ppb)
    block_averager.push_container(apl_box[:, 1])
    block_average, std_err = block_averager.get()
    print(("Block Averaged ({} points with {} blocks) Area per lipid estimates (squared Angstrom): ".format(
        len(apl_box), block_averager.number_of_blocks())))
    print(("  via the box dimensions: {:0.4f} +- {:0.4f}".format(block_average, std_err)))
    print("  via the gridding procedure: ")
    for item in apl_grid.keys():
        block_averager = BlockAverager(points_per_block=ppb)
        block_averager.push_container(apl_grid[item][:, 1])
        block_average, std_err = block_averager.get()
        print(("    {}: {:0.4f} +- {:0.4f}".format(item, block_average, std_err)))
    return


def nearest_neighbor_fraction(structure_file, trajectory_file, selection_string, lipid_resnames,
                              frame_start=0, frame_end=-1, frame_interval=1, dump_path=None,
                              name_dict=None, **analyzer_kwargs):
    """Nearest-neighbour fractions of every ordered pair of the listed lipid types."""
    analyzer = BilayerAnalyzer(structure=structure_file, trajectory=trajectory_file,
                               selection=selection_string, **analyzer_kwargs)
    analyzer.set_frame_range(frame_start, frame_end, frame_interval)
    analyzer.remove_analysis('msd_1')
    analyzer.rep_settings['com_frame']['name_dict'] = name_dict
    pairs = [[a, b] for a in lipid_resnames for b in lipid_resnames]
    for l1, l2 in pairs:
        analyzer.add_analysis("nnf nnf_" + l1 + "_" + l2 + " resname_1 " + l1 + " resname_2 " + l2)
    analyzer.print_analysis_protocol()
    print(" ")
    analyzer.run_analysis()
    analyzer.dump_data(path=dump_path)
    for l1, l2 in pairs:
        t_nnf = analyzer.get_analysis_data("nnf_" + l1 + "_" + l2)
        print(("Nearest neighbor fraction for lipid pair {} and {} : {:0.4f} +- {:0.4f}".format(
            l1, l2, t_nnf[-1][2], t_nnf[-1][3])))
    return


def com_lateral_rdf(structure_file, trajectory_file, bilayer_selection_string, resnames=None,
                    frame_start=0, frame_end=-1, frame_interval=1, dump_path="./", name_dict=None,
                    **analyzer_kwargs):
    """Lateral RDFs of every pair of the listed lipid types, per leaflet (80 bins up to 40 A).  Returns
    {leaflet: [(name, bins, rdf)]} (the reference draws these and returns nothing)."""
    analyzer = BilayerAnalyzer(structure=structure_file, trajectory=trajectory_file,
                               selection=bilayer_selection_string, **analyzer_kwargs)
    analyzer.set_frame_range(frame_start, frame_end, frame_interval)
    analyzer.remove_analysis('msd_1')
    analyzer.rep_settings['com_frame']['name_dict'] = name_dict
    res_pairs = []
    for a in (resnames or []):
        for b in resnames:
            if [a, b] not in res_pairs:
                res_pairs.append([a, b])
    for a, b in res_pairs:
        for leaflet in ("upper", "lower"):
            analyzer.add_analysis("com_lateral_rdf com_lateral_rdf_{}-{}_{} resname_1 {} resname_2 {} leaflet {} "
                                  "range_outer 40.0 n_bins 80".format(a, b, leaflet, a, b, leaflet))
    analyzer.print_analysis_protocol()
    analyzer.run_analysis()
    analyzer.dump_data(path=dump_path)
    curves = {'upper': [], 'lower': [], 'both': []}
    for a, b in res_pairs:
        rdf_u, bins = analyzer.get_analysis_data("com_lateral_rdf_{}-{}_upper".format(a, b))
        rdf_l, bins = analyzer.get_analysis_data("com_lateral_rdf_{}-{}_lower".format(a, b))
        name = "{}-{}".format(a, b)
        curves['upper'].append((name, bins, rdf_u))
        curves['lower'].append((name, bins, rdf_l))
        curves['both'].append((name, bins, (rdf_u + rdf_l) / 2.0))
    return curves
