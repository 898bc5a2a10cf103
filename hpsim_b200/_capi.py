"""ctypes binding of the C ABI declared in include/hpsim_b200.h.

`load()` opens the CUDA product library (hpsim_b200/csrc/libhpsim_b200.so) and
fails loudly when it is missing -- there is no CPU fallback.  `bind(cdll)` only
attaches prototypes, so the test-suite can bind the very same prototypes onto the
CPU oracle (oracle/_build/libhpsim_oracle.so), which implements the same ABI.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .lattice import ELEM_DTYPE, NPARAM, NAME_LEN

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HPSIM_B200_LIB") or os.path.join(HERE, "csrc", "libhpsim_b200.so")  # (the override is for build experiments)

OK, ERR_CUDA, ERR_ARG, ERR_STATE, ERR_LATTICE, ERR_NOTFOUND, ERR_COMM = range(7)

X, XP, Y, YP, PHI, W = range(6)

STAT_NAMES = [
    "x_avg", "x_avg_good", "xp_avg", "y_avg", "y_avg_good", "yp_avg", "phi_avg", "phi_avg_r",
    "phi_avg_good", "w_avg", "w_avg_good", "x_sig", "x_sig_good", "xp_sig", "y_sig", "y_sig_good",
    "yp_sig", "phi_sig", "phi_sig_r", "phi_sig_good", "w_sig", "w_sig_good", "x_emit", "y_emit",
    "z_emit", "r_max", "abs_phi_max", "num_loss", "num_lloss", "num_good",
]
STAT = {n: i for i, n in enumerate(STAT_NAMES)}

UPD_NAMES = [
    "loss", "lloss", "good_count", "avg_x", "avg_x_good", "avg_xp", "avg_y", "avg_y_good", "avg_yp",
    "avg_phi", "avg_phi_good", "avg_phi_r", "avg_phi_r_good", "avg_w", "avg_w_good", "sig_x",
    "sig_x_good", "sig_xp", "sig_y", "sig_y_good", "sig_yp", "sig_phi", "sig_phi_r", "sig_phi_r_good",
    "sig_w", "sig_w_good", "emittance", "stat_for_spch", "avg_xy", "max_r_phi", "rel_phi", "rel_phi_good",
]
UPD = {n: i for i, n in enumerate(UPD_NAMES)}

PROF_CLASSES = ["segment", "lloss", "moments", "deposit", "field"]
MON_FIELDS = ["xavg", "xsig", "xpavg", "xpsig", "xemit", "yavg", "ysig", "ypavg", "ypsig", "yemit", "phiavg", "phisig", "wavg", "wsig",
              "zemit", "loss_ratio", "loss_local"]  # members of the reference's PlotData, hps_monitor_field order
SC_INTERVAL, SC_ADJ_BUNCH, SC_ADJ_BUNCH_CUTOFF_W, SC_MESH_CUTOFF_W, SC_REMESH_THRESHOLD, SC_FRACTION, SC_NR, SC_NZ = range(8)


class Elem(C.Structure):
    _fields_ = [("type", C.c_uint32), ("flags", C.c_uint32), ("name", C.c_char * NAME_LEN),
                ("p", C.c_double * NPARAM)]


assert C.sizeof(Elem) == ELEM_DTYPE.itemsize


class RunStats(C.Structure):
    _fields_ = [("particles", C.c_uint64), ("elements", C.c_uint64), ("kicks", C.c_uint64),
                ("remesh", C.c_uint64), ("sweeps", C.c_uint64), ("extra_sweeps", C.c_uint64),
                ("stat_redo", C.c_uint64), ("launches", C.c_uint64), ("device_ms", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_dp = C.POINTER(C.c_double)
_up = C.POINTER(C.c_uint32)
_vp = C.c_void_p

# name -> (restype, argtypes); this table is also what tests check the exports against
PROTOTYPES = {
    "hps_abi_version": (C.c_int, []),
    "hps_ctx_create": (_vp, [C.c_int]),
    "hps_ctx_destroy": (None, [_vp]),
    "hps_last_error": (C.c_char_p, [_vp]),
    "hps_ctx_set_compat": (C.c_int, [_vp, C.c_int]),
    "hps_ctx_set_math": (C.c_int, [_vp, C.c_int]),
    "hps_ctx_get_math": (C.c_int, [_vp]),
    "hps_comm_unique_id": (C.c_int, [_vp]),
    "hps_ctx_comm_init": (C.c_int, [_vp, C.c_int, C.c_int, _vp]),
    "hps_ctx_comm_mode": (C.c_int, [_vp]),
    "hps_ctx_comm_stats": (C.c_int, [_vp, C.POINTER(C.c_uint64)]),
    "hps_ctx_synchronize": (C.c_int, [_vp]),
    "hps_host_alloc": (_vp, [C.c_size_t]),
    "hps_host_free": (None, [_vp]),
    "hps_ctx_launch_count": (C.c_uint64, [_vp]),
    "hps_ctx_set_profile": (C.c_int, [_vp, C.c_int]),
    "hps_ctx_get_profile": (C.c_int, [_vp, _vp, _vp]),
    "hps_ctx_event_record": (C.c_int, [_vp, C.c_int]),
    "hps_ctx_event_elapsed": (C.c_int, [_vp, C.c_int, C.c_int, _dp]),
    "hps_ctx_flush_l2": (C.c_int, [_vp]),
    "hps_beam_create": (_vp, [_vp, C.c_uint64, C.c_double, C.c_double, C.c_double]),
    "hps_beam_destroy": (None, [_vp]),
    "hps_beam_local_range": (C.c_int, [_vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "hps_beam_local_stride": (C.c_uint64, [_vp]),
    "hps_beam_upload": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "hps_beam_download": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "hps_beam_download_phi_relative": (C.c_int, [_vp, C.c_int, _vp]),
    "hps_beam_set_dc": (C.c_int, [_vp] + [C.c_double] * 9 + [C.c_uint32]),
    "hps_beam_set_waterbag": (C.c_int, [_vp] + [C.c_double] * 12 + [C.c_uint32]),
    "hps_beam_save": (C.c_int, [_vp, C.c_int]),
    "hps_beam_restore": (C.c_int, [_vp, C.c_int]),
    "hps_beam_set_ref_w": (C.c_int, [_vp, C.c_double]),
    "hps_beam_set_ref_phi": (C.c_int, [_vp, C.c_double]),
    "hps_beam_set_frequency": (C.c_int, [_vp, C.c_double]),
    "hps_beam_get_ref_w": (C.c_double, [_vp]),
    "hps_beam_get_ref_phi": (C.c_double, [_vp]),
    "hps_beam_get_frequency": (C.c_double, [_vp]),
    "hps_beam_get_mass": (C.c_double, [_vp]),
    "hps_beam_get_charge": (C.c_double, [_vp]),
    "hps_beam_get_current": (C.c_double, [_vp]),
    "hps_beam_get_size": (C.c_uint64, [_vp]),
    "hps_beam_shift": (C.c_int, [_vp, C.c_int, C.c_double]),
    "hps_beam_cut": (C.c_int, [_vp, C.c_int, C.c_double, C.c_double]),
    "hps_beam_update": (C.c_int, [_vp, C.c_int]),
    "hps_beam_get": (C.c_int, [_vp, C.c_int, _dp]),
    "hps_beam_moment_stats": (C.c_int, [_vp, C.POINTER(C.c_uint64), _dp, _dp]),
    "hps_monitor_create": (_vp, [_vp, C.c_uint32]),
    "hps_monitor_destroy": (None, [_vp]),
    "hps_ctx_set_monitor": (C.c_int, [_vp, _vp]),
    "hps_monitor_download": (C.c_int, [_vp, _vp, C.POINTER(C.c_uint32)]),
    "hps_lattice_create": (_vp, [_vp, _vp, C.c_uint32]),
    "hps_lattice_destroy": (None, [_vp]),
    "hps_lattice_size": (C.c_uint32, [_vp]),
    "hps_lattice_index": (C.c_int64, [_vp, C.c_char_p]),
    "hps_lattice_get_elem": (C.c_int, [_vp, C.c_uint32, C.POINTER(Elem)]),
    "hps_lattice_update_elem": (C.c_int, [_vp, C.c_uint32, C.POINTER(Elem)]),
    "hps_scheff_create": (_vp, [_vp, C.c_uint32, C.c_uint32, C.c_int]),
    "hps_scheff_destroy": (None, [_vp]),
    "hps_scheff_set": (C.c_int, [_vp, C.c_int, C.c_double]),
    "hps_scheff_get": (C.c_int, [_vp, C.c_int, _dp]),
    "hps_scheff_set_mesh_size": (C.c_int, [_vp, C.c_uint32, C.c_uint32]),
    "hps_scheff_download_tables": (C.c_int, [_vp, _vp, _vp, _vp]),
    "hps_scheff_phase_cycles": (C.c_int, [_vp, C.POINTER(C.c_uint64)]),
    "hps_simulate": (C.c_int, [_vp, _vp, _vp, _vp, C.c_uint32, C.c_uint32, C.POINTER(RunStats)]),
    "hps_engine_reset": (C.c_int, [_vp, _vp]),
}


class HpsError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"hpsim_b200 error {code}: {msg}")
        self.code = code


def bind(cdll: C.CDLL) -> C.CDLL:
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(cdll, name)  # AttributeError when a symbol is missing: intended
        fn.restype = res
        fn.argtypes = args
    return cdll


_lib = None


def load() -> C.CDLL:
    """The CUDA product library.  No fallback: raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(hpsim_b200 has no CPU fallback)")
        _lib = bind(C.CDLL(LIB_PATH, mode=C.RTLD_LOCAL))
    return _lib


def ptr(a):
    """void* of a C-contiguous numpy array (or None)."""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def u32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.uint32)
