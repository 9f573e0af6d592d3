"""ctypes binding of libpbk (include/pbk.h).  This is the stub INTEGRATION.md describes.

There is deliberately no fallback: if ``_pbk.so`` is missing or a CUDA device is not
available every call raises ``RuntimeError`` -- nothing under pybilt_b200 computes the hot
path on the CPU.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_double, c_int, c_int32, c_int64, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(HERE, "_pbk.so")

ST_IMAGE_OVERFLOW = 1
ST_LEAFLET_TIE = 2
ST_UNWRAP_AMBIGUOUS = 4
ST_FUSED_RANGE = 8
ST_INTERNAL = 16
E_UNSUPPORTED = -3

_SIGNATURES = {
    "pbk_abi_version": (c_int, []),
    "pbk_ctx_create": (c_int, [c_int, ctypes.POINTER(c_void_p)]),
    "pbk_ctx_destroy": (None, [c_void_p]),
    "pbk_last_error": (ctypes.c_char_p, [c_void_p]),
    "pbk_planes_to_aos": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_int64,
                                  c_int, c_int64, c_void_p, c_void_p]),
    "pbk_unwrap_images": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int64,
                                  c_int64, c_void_p, c_void_p, c_void_p]),
    "pbk_membrane_com_seq": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64,
                                     c_double, c_void_p, c_void_p]),
    "pbk_membrane_com": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64,
                                 c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_int32, c_void_p,
                                 c_int32, c_double, c_void_p, c_void_p, c_void_p]),
    "pbk_lipid_com": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                              c_void_p, c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p,
                              c_void_p]),
    "pbk_sm_count": (c_int, [c_void_p]),
    "pbk_reduce_fused_scratch_bytes": (c_int, [c_int64, c_int64, c_int, ctypes.POINTER(c_int64),
                                               ctypes.POINTER(c_int32), ctypes.POINTER(c_int32)]),
    "pbk_reduce_fused": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                 c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_int32, c_void_p,
                                 c_int32, c_double, c_void_p, c_int, c_int64, c_int64, c_int64, c_int,
                                 c_int, c_int, c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p,
                                 c_void_p, c_void_p, c_void_p, c_void_p]),
    "pbk_reduce_stream_scratch_bytes": (c_int, [c_int64, c_int64, c_int64, c_int32, c_int32,
                                                ctypes.POINTER(c_int64)]),
    "pbk_reduce_stream": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_int32, c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_int32, c_void_p,
                                  c_int32, c_double, c_void_p, c_int, c_int64, c_int64, c_int64, c_int, c_int,
                                  c_void_p, c_void_p, c_int64, c_int64, c_int, c_void_p, c_void_p, c_int64,
                                  c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "pbk_leaflets": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int, c_int64, c_int, c_void_p,
                             c_void_p, c_void_p, c_void_p]),
    "pbk_leaflet_distance": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_void_p, c_void_p]),
    "pbk_label_changes": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p]),
    "pbk_debug_div_check": (c_int, [c_void_p, ctypes.c_uint64, c_int64, c_void_p, c_void_p]),
    "pbk_orient_labels": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                  c_int64, c_int64, c_int64, c_int, c_int64, c_int, c_void_p, c_void_p, c_void_p]),
    "pbk_rewrap": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p]),
    "pbk_ct_nearest": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_void_p, c_int32, c_int, c_int, c_int,
                               c_int, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "pbk_ct_adjacency": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int, c_int, c_double, c_int, c_int64,
                                 c_int64, c_void_p, c_void_p]),
    "pbk_msd_pairs": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_int32, c_int, c_int,
                              c_void_p, c_void_p, c_int32, c_void_p, c_void_p]),
    "pbk_ald_pairs": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_int32, c_int, c_int,
                              c_void_p, c_void_p, c_int32, c_void_p, c_void_p]),
    "pbk_rdf": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int,
                        c_int, c_int, c_int, c_int, c_int32, c_double, c_double, c_void_p, c_void_p,
                        c_void_p, c_void_p, c_void_p]),
    "pbk_rdf_multi": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int,
                              c_int, c_int, c_int32, c_double, c_double, c_void_p, c_void_p, c_void_p,
                              c_void_p, c_void_p]),
    "pbk_nnf": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int,
                        c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "pbk_nnf_multi": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int,
                              c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "pbk_nnf_select": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int,
                               c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "pbk_knn24": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_int, c_void_p,
                          c_void_p]),
    "pbk_halperin_nelson": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int,
                                    c_int, c_int64, c_void_p, c_void_p]),
    "pbk_nnf_select_many": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int,
                                    c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "pbk_lipid_grid": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64,
                               c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                               c_void_p]),
    "pbk_grid_reduce": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64,
                                c_int64, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                c_void_p, c_void_p]),
    "pbk_comm_create": (c_int, [c_void_p, c_int, c_int, c_int64, ctypes.POINTER(c_void_p)]),
    "pbk_comm_handle": (c_int, [c_void_p, c_void_p]),
    "pbk_comm_connect": (c_int, [c_void_p, c_void_p]),
    "pbk_comm_destroy": (None, [c_void_p]),
    "pbk_comm_prefix_i32": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    "pbk_allreduce_u64": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    "pbk_allreduce_f64": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
}

EXPORTS = tuple(_SIGNATURES)

_lib = None


def load() -> ctypes.CDLL:
    """dlopen libpbk and bind every symbol of include/pbk.h (no CUDA call is made)."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise RuntimeError(
                "pybilt_b200: %s is missing -- build it with `python -m pybilt_b200.build` "
                "(there is no CPU fallback)" % SO_PATH)
        lib = ctypes.CDLL(SO_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        if lib.pbk_abi_version() != 8:
            raise RuntimeError("pybilt_b200: libpbk ABI version mismatch")
        _lib = lib
    return _lib


class Context:
    """Owns a ``pbk_ctx*`` for one CUDA device; methods mirror the C entry points and raise
    ``RuntimeError`` (the reference's exception type for bad input) on failure."""

    def __init__(self, device: int = 0):
        self.lib = load()
        self.device = int(device)
        h = c_void_p()
        rc = self.lib.pbk_ctx_create(self.device, ctypes.byref(h))
        if rc != 0:
            raise RuntimeError("pybilt_b200: no usable CUDA device %d (pbk_ctx_create -> %d); "
                               "the hot path has no CPU fallback" % (device, rc))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self.lib.pbk_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def try_call(self, name: str, *args) -> bool:
        """Like call(), but returns False when the entry point answers PBK_EUNSUPPORTED."""
        rc = getattr(self.lib, name)(self._h, *args)
        if rc == E_UNSUPPORTED:
            return False
        if rc != 0:
            msg = self.lib.pbk_last_error(self._h)
            raise RuntimeError("libpbk %s failed (%d): %s" % (name, rc, (msg or b"").decode()))
        return True

    def sm_count(self) -> int:
        return int(self.lib.pbk_sm_count(self._h))

    def call(self, name: str, *args):
        rc = getattr(self.lib, name)(self._h, *args)
        if rc != 0:
            msg = self.lib.pbk_last_error(self._h)
            raise RuntimeError("libpbk %s failed (%d): %s" % (name, rc, (msg or b"").decode()))
