"""ctypes binding of libjacket_b200.so (include/jacket_b200.h).

The library is built in-tree by ``__graft_entry__.build()``.  There is no CPU
path: if the shared object is missing, or there is no sm_100 device, the
product fails loudly here.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# JK_LIB selects an experimental build of the same sources (kernel A/B runs); default = the in-tree library
LIB_PATH = os.environ.get("JK_LIB") or os.path.join(_HERE, "libjacket_b200.so")

# keep in sync with include/jacket_b200.h
SEC_NPROP = 8
TABLE_NCOL = 16
MEMBER_NCOL = 7
DETAIL_NCOL = 4
ORDER_NATURAL, ORDER_RCM = 0, 1
SOLVER_BANDED, SOLVER_DENSE = 0, 1
NTIMERS = 13
TIMER_NAMES = ("assemble", "factor", "wave_setup", "morison", "rhs", "solve_fwd", "solve_bwd",
               "post", "reduce", "scan_total", "h2d", "d2h", "solve_fwd2")
TABLE_COLUMNS = ("t", "phase_deg", "total_kN", "drag_kN", "inertia_kN", "Fx_kN", "Fy_kN", "Fz_kN",
                 "max_disp_mm", "max_disp_node", "max_util", "max_util_member", "max_vm_MPa",
                 "sum_Rx", "sum_Ry", "sum_Rz")
MEMBER_COLUMNS = ("Fx_max_kN", "Fy_max_kN", "Fz_max_kN", "My_max_kNm", "Mz_max_kNm",
                  "von_mises_max_MPa", "utilization")
DETAIL_COLUMNS = ("drag_kN", "inertia_kN", "total_kN", "submerged_length")

ERRORS = {-1: "JK_EINVAL", -2: "JK_ECUDA", -3: "JK_ENODEVICE", -4: "JK_ENOTSPD", -5: "JK_ESTATE"}


class JacketError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"{ERRORS.get(code, code)}: {message}")
        self.code = code


class NotPositiveDefinite(JacketError):
    """Cholesky met a non-positive pivot (the reference would silently fall to lstsq, GUI.py:486-487)."""


_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_lib = None


def _sig(lib):
    H = C.c_void_p
    lib.jk_version.restype = C.c_int
    lib.jk_last_error.restype = C.c_char_p
    lib.jk_last_error.argtypes = [H]
    lib.jk_set_option.argtypes = [H, C.c_char_p, C.c_int]
    lib.jk_get_option.argtypes = [H, C.c_char_p, C.POINTER(C.c_int)]
    lib.jk_option_count.restype = C.c_int
    lib.jk_option_name.argtypes = [C.c_int]
    lib.jk_option_name.restype = C.c_char_p
    lib.jk_kinematics_points.argtypes = [H, C.c_int, _dp, C.c_double, _dp]
    lib.jk_create.argtypes = [C.c_int, C.c_void_p, C.c_int, _dp, C.c_int, _ip, _ip, C.c_int, _dp, C.POINTER(H)]
    lib.jk_destroy.argtypes = [H]
    lib.jk_set_supports.argtypes = [H, C.c_int, _ip, C.c_int, C.c_int]
    lib.jk_assemble.argtypes = [H, C.c_double, C.c_double]
    lib.jk_factor.argtypes = [H]
    lib.jk_factor_begin.argtypes = [H]
    lib.jk_set_static_load.argtypes = [H, _dp]
    lib.jk_set_wave_airy.argtypes = [H] + [C.c_double] * 6
    lib.jk_set_wave_fourier.argtypes = [H] + [C.c_double] * 5 + [C.c_int, _dp, _dp]
    lib.jk_set_morison.argtypes = [H] + [C.c_double] * 5 + [C.c_int, _dp, _dp]
    lib.jk_morison_scan.argtypes = [H, C.c_int, _dp, _dp, C.POINTER(C.c_int64)]
    lib.jk_morison_single.argtypes = [H, C.c_double, _dp, _dp, _dp]
    lib.jk_phase_scan.argtypes = [H, C.c_int, _dp, C.c_double, _dp, C.POINTER(C.c_int64)]
    lib.jk_phase_scan_dev.argtypes = [H, C.c_int, C.c_void_p, C.c_double]
    lib.jk_step_dev.argtypes = [H, C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_double]
    lib.jk_step.argtypes = [H, C.c_double, C.c_double, C.c_int, _dp, C.c_double, _dp, C.POINTER(C.c_int64)]
    lib.jk_read_table.argtypes = [H, C.c_int, _dp, C.POINTER(C.c_int64)]
    lib.jk_solve.argtypes = [H, C.c_int, _dp, C.c_double]
    lib.jk_ensemble_scan.argtypes = [H, C.c_int, C.c_int, _dp, _dp, _dp, _dp, _dp, _dp, C.c_double, _dp, C.POINTER(C.c_int64)]
    lib.jk_ensemble_scan_sea_states.argtypes = [H, C.c_int, C.c_int, _dp, _dp, _dp, C.c_double, _dp, C.c_double, _dp, C.POINTER(C.c_int64), _dp]
    lib.jk_fetch_phase.argtypes = [H, C.c_int, _dp, _dp, _dp, _dp, _dp]
    lib.jk_fetch_member_column.argtypes = [H, C.c_int, C.c_int, C.c_int, _dp]
    lib.jk_get_dims.argtypes = [H, _ip]
    lib.jk_get_order.argtypes = [H, _ip]
    lib.jk_get_K.argtypes = [H, _dp]
    lib.jk_get_elements.argtypes = [H, _dp, _dp, _dp, _dp]
    lib.jk_get_timings.argtypes = [H, _dp]
    lib.jk_residual.argtypes = [H, _dp]
    lib.jk_solver_stats.argtypes = [H, _dp]
    lib.jk_sweep_program.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, _ip, _ip, C.c_int, _ip]
    lib.jk_sweep_program.restype = C.c_int
    lib.jk_launch_count.argtypes = [H]
    lib.jk_launch_count.restype = C.c_int64
    lib.jk_stream.argtypes = [H]
    lib.jk_stream.restype = C.c_void_p
    for name in ("jk_table_dev", "jk_critical_value_dev", "jk_critical_index_dev"):
        getattr(lib, name).argtypes = [H]
        getattr(lib, name).restype = C.c_void_p
    for name in ("jk_create", "jk_destroy", "jk_set_supports", "jk_assemble", "jk_factor", "jk_factor_begin", "jk_set_static_load",
                 "jk_set_wave_airy", "jk_set_wave_fourier", "jk_set_morison", "jk_morison_scan", "jk_morison_single",
                 "jk_phase_scan", "jk_phase_scan_dev", "jk_read_table", "jk_solve", "jk_ensemble_scan", "jk_ensemble_scan_sea_states", "jk_fetch_phase",
                 "jk_fetch_member_column", "jk_get_dims", "jk_get_order", "jk_get_K", "jk_get_elements",
                 "jk_get_timings", "jk_residual", "jk_solver_stats", "jk_set_option", "jk_get_option", "jk_kinematics_points", "jk_step", "jk_step_dev"):
        getattr(lib, name).restype = C.c_int


def lib():
    """Load (once) and return the shared library; raise if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                "jacket_b200 has no CPU fallback.")
        _lib = C.CDLL(LIB_PATH)
        _sig(_lib)
    return _lib


def dptr(a):
    return None if a is None else a.ctypes.data_as(_dp)


def iptr(a):
    return None if a is None else a.ctypes.data_as(_ip)


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def check(rc, handle=None):
    if rc == 0:
        return
    msg = lib().jk_last_error(handle)
    msg = msg.decode() if msg else ""
    raise (NotPositiveDefinite if rc == -4 else JacketError)(rc, msg)
