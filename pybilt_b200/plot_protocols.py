"""Plot protocols of the in-scope analyses: ``msd``, ``apl`` and ``bilayer_thickness`` plots
(reference registry ``pybilt/bilayer_analyzer/plot_protocols.py:9-12,166-330,
444-520``).  Same plot strings (``"<plot key> <plot id> <analysis id> <legend> ..."``), ids, error
behaviour and file names (``<plot id>.pdf``), so the reference's prefab drivers can register and
save their plots on the drop-in analyzer.

What a plot shows is collected as plain series (``PlotFunctionProtocol.series``: label, x in ns,
y, optional error band); drawing them needs matplotlib, which this image does not have -- without
it ``generate_plot`` stores the series next to where the figure would go (``<plot id>.npz``) and
says so, instead of failing an analysis run at its last step.
"""
from __future__ import annotations

import numpy as np

command_protocols = {}
valid_plots = []
plot_analysis_dict = {}


def _matplotlib():
    try:
        import matplotlib
        matplotlib.use("Agg", force=False)
        import matplotlib.pyplot as plt
        return plt
    except Exception:
        return None


class PlotFunctionProtocol(object):
    """One plot: which analyses it draws, under which legend names."""
    plot_key = 'none'
    title = "plot"
    y_label = ""

    def __init__(self, args, analysis_keys, analysis_ids):
        self.plot_id = args[0]
        wanted = plot_analysis_dict[self.plot_key]
        wanted = wanted if isinstance(wanted, (list, tuple)) else [wanted]
        self.valid_args = [a_id for key, a_id in zip(analysis_keys, analysis_ids) if key in wanted]
        self.save_file_name = self.plot_id + ".pdf"
        self.interval = 1
        self.include = []
        self.names = {}
        if len(args) > 1:
            self.parse_args(args)
        self.my_plot = None

    def parse_args(self, args):
        if (len(args) - 1) % 2:
            raise RuntimeError("plot " + self.plot_id + ": every analysis id needs a legend name")
        for i in range(1, len(args), 2):
            a_id, legend = args[i], args[i + 1]
            if a_id not in self.valid_args:
                raise RuntimeError("ignoring invalid argument key " + a_id + " for plot " + self.plot_key)
            self.include.append(a_id)
            self.names[a_id] = legend

    def print_protocol(self):
        print("Plot " + self.plot_id + " for " + self.title + ":" + str(self.include))

    # -- data ---------------------------------------------------------------------------------
    def _rows(self, protocol, legend):
        """[(label, rows)] of one analysis; rows = [time, value, ...]."""
        return [(legend, np.asarray(protocol.get_data()))]

    def series(self, analysis_protocol):
        """[(label, t_ns, y, err or None)] -- times converted ps -> ns like the reference's plots."""
        out = []
        for a_id in self.include:
            proto = analysis_protocol.command_protocol[a_id]
            for label, rows in self._rows(proto, self.names[a_id]):
                rows = np.asarray(rows)[::self.interval]
                err = rows[:, 3] if rows.shape[1] >= 4 and self.plot_key != 'msd' else None
                y = rows[:, 2] if rows.shape[1] >= 4 and self.plot_key != 'msd' else rows[:, 1]
                out.append((label, rows[:, 0] / 1000.0, y, err))
        return out

    # -- rendering ----------------------------------------------------------------------------
    def _draw(self, analysis_protocol, save, show):
        ser = self.series(analysis_protocol)
        plt = _matplotlib()
        if plt is None:
            target = self.save_file_name[:-4] + ".npz"
            np.savez(target, **{"%d_%s_%s" % (k, lab, nm): arr for k, (lab, t, y, e) in enumerate(ser)
                                for nm, arr in (("t_ns", t), ("y", y)) })
            print("matplotlib is not installed: series of plot '%s' written to %s" % (self.plot_id, target))
            return ser
        fig, ax = plt.subplots()
        for lab, t, y, e in ser:
            ax.plot(t, y, label=lab, linewidth=2.0)
            if e is not None:
                ax.fill_between(t, y - e, y + e, alpha=0.2)
        ax.set_xlabel("Time (ns)")
        ax.set_ylabel(self.y_label)
        if len(ser) > 1:
            ax.legend(loc=2)
        if save:
            fig.savefig(self.save_file_name)
        if show:
            plt.show()
        plt.close(fig)
        return ser

    def generate_plot(self, analysis_protocol):
        return self._draw(analysis_protocol, save=True, show=False)

    def show_plot(self, analysis_protocol):
        return self._draw(analysis_protocol, save=False, show=True)


def _register(key, analyses):
    def deco(cls):
        valid_plots.append(key)
        plot_analysis_dict[key] = analyses
        command_protocols[key] = cls
        cls.plot_key = key
        return cls
    return deco


@_register('msd', 'msd')
class MSDPlotProtocol(PlotFunctionProtocol):
    title = "MSD analysis"
    y_label = "Distance in the lateral plane ($\\AA^2$)"


@_register('apl', ['apl_box', 'apl_grid'])
class APLPlotProtocol(PlotFunctionProtocol):
    """apl_grid analyses contribute one series per entry of their output dict (composite + one
    per lipid type, plot_protocols.py:282-296); plotted: running mean with its deviation."""
    title = "area per lipid (APL) analysis"
    y_label = "Area per lipid ($\\AA^2$)"

    def _rows(self, protocol, legend):
        data = protocol.get_data()
        if isinstance(data, dict):
            return [(k, np.asarray(v)) for k, v in data.items()]
        return [(legend, np.asarray(data))]


@_register('bilayer_thickness', ['thickness_grid', 'thickness_leaflet_distance'])
class BTPlotProtocol(PlotFunctionProtocol):
    title = "bilayer thickness (BT) analysis"
    y_label = "Bilayer thickness ($\\AA$)"


class PlotProtocol(object):
    """The plots of a BilayerAnalyzer, by id (plot_protocols.py:15-118)."""

    def __init__(self, plot_commands, analysis_protocol):
        self.in_commands = plot_commands
        self.arguments, self.plot_keys, self.plot_ids = [], [], []
        self.command_protocol = {}
        self.n_commands = 0
        for command in (plot_commands or []):
            self.add_plot(command, analysis_protocol)

    def add_plot(self, plot_string, analysis_protocol):
        command = plot_string.split() if isinstance(plot_string, str) else list(plot_string)
        if len(command) < 2:
            raise RuntimeError("wrong number of arguments for plot " + str(command))
        plot_key, plot_id, args = command[0], command[1], command[1:]
        if plot_key not in valid_plots:
            raise RuntimeError("invalid plot type '{}' : {}".format(plot_key, command))
        if plot_id in self.plot_ids:
            raise RuntimeError("plot id '{}' has already been used".format(plot_id))
        proto = command_protocols[plot_key](args, analysis_protocol.analysis_keys, analysis_protocol.analysis_ids)
        self.arguments.append(args)
        self.plot_ids.append(plot_id)
        self.plot_keys.append(plot_key)
        self.command_protocol[plot_id] = proto
        self.n_commands += 1

    def remove_plot(self, plot_id):
        if plot_id not in self.plot_ids:
            raise RuntimeWarning("no plot with id '{}'".format(plot_id))
        k = self.plot_ids.index(plot_id)
        del self.command_protocol[plot_id], self.arguments[k], self.plot_keys[k], self.plot_ids[k]
        self.n_commands -= 1

    def print_protocol(self):
        print("with plots:")
        if not self.plot_ids:
            print('None')
        for plot_id in self.plot_ids:
            self.command_protocol[plot_id].print_protocol()

    def save_plots(self, analysis_protocol):
        print('saving plots...')
        for plot_id in self.plot_ids:
            print("plot id: " + plot_id + " ---> " + self.command_protocol[plot_id].save_file_name)
            self.command_protocol[plot_id].generate_plot(analysis_protocol)
