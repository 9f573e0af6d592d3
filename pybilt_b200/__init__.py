"""pybilt_b200 -- B200-native lipid-reduction hot path behind PyBILT's BilayerAnalyzer API.

``from pybilt_b200 import BilayerAnalyzer`` is the drop-in for
``from pybilt.bilayer_analyzer import BilayerAnalyzer`` for the analyses 'msd', 'msd_multi',
'com_lateral_rdf', 'nnf', 'apl_grid' and 'thickness_grid'.  The numerical work is done by
libpbk (include/pbk.h, pybilt_b200/csrc) on an sm_100a device; there is no CPU fallback.
"""
from .bilayer_analyzer import BilayerAnalyzer  # noqa: F401
from . import analysis_protocols  # noqa: F401

__all__ = ["BilayerAnalyzer", "analysis_protocols"]
