"""ctypes binding of libstereo_b200.so (include/stereo_b200.h).

The library is the product: there is no Python or CPU fallback.  If the shared
object is missing, importing the bound functions raises; if it loads but no CUDA
device is visible every compute call returns SB_ENODEV and we raise.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_double, c_int, c_int32, c_int64, c_uint32

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libstereo_b200.so")

SB_OK, SB_EINVAL, SB_ENODEV, SB_ECUDA, SB_ENOTGRID, SB_ENOMEM, SB_EUNSUP = 0, -1, -2, -3, -4, -5, -6
SB_F32, SB_F64 = 0, 1
SB_MAX_LABELS = 256

_dp = POINTER(c_double)
_up = POINTER(c_uint32)
_ip = POINTER(c_int32)


class SbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"stereo_b200 error {code}: {msg}")
        self.code = code


class TrwsOptions(Structure):
    _fields_ = [("maxiter", c_double), ("max_relgap", c_double), ("precision", c_int),
                ("fuse_rounding", c_int), ("col_blocks", c_int), ("latency_mode", c_int), ("reserved", c_int * 4)]


class TrwsTiming(Structure):
    _fields_ = [("setup_ms", c_double), ("solve_ms", c_double), ("sweep_ms_avg", c_double),
                ("download_ms", c_double), ("kernel_launches", c_int64), ("sweep_kernel_ms", c_double),
                ("sweep_kernel_launches", c_int64), ("reserved", c_int64 * 1)]


_lib = None


def lib():
    """Load libstereo_b200.so (built by stereo_b200/csrc/Makefile) or fail loudly."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: build it with `make -C stereo_b200/csrc` "
                "(or python -c 'import __graft_entry__ as g; g.build()'). There is no fallback path.")
        L = ctypes.CDLL(LIB_PATH)
        L.sb_version.restype = c_char_p
        L.sb_last_error.restype = c_char_p
        L.sb_device_count.restype = c_int
        L.sb_set_device.argtypes = [c_int]
        L.sb_kernel_launches.restype = c_int64
        L.sb_trws_default_options.argtypes = [POINTER(TrwsOptions)]
        L.sb_trws_default_options.restype = None
        L.sb_trws_solve.argtypes = [c_int, c_int, c_int64, c_int64, _dp, _up, _dp, _dp, _dp, c_double,
                                    POINTER(TrwsOptions), _dp, _dp, _dp, _dp, POINTER(TrwsTiming)]
        L.sb_trws_create.argtypes = [c_int, c_int, c_int64, c_int64, _dp, _up, _dp, _dp, _dp, c_double,
                                     POINTER(TrwsOptions), POINTER(ctypes.c_void_p)]
        L.sb_trws_create_banded.argtypes = [c_int, c_int, c_int64, c_int64, _dp, _up, _dp, _dp, _dp, c_double,
                                            POINTER(TrwsOptions), c_int, c_int, POINTER(ctypes.c_void_p)]
        L.sb_trws_ipc_export.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
        L.sb_trws_ipc_attach.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_char_p]
        L.sb_trws_pass.argtypes = [ctypes.c_void_p, c_int, c_int, _dp]
        L.sb_trws_reset.argtypes = [ctypes.c_void_p]
        L.sb_trws_minimize.argtypes = [ctypes.c_void_p, c_double, c_double, _dp, _dp, _dp, POINTER(TrwsTiming)]
        L.sb_trws_get_labels.argtypes = [ctypes.c_void_p, _dp]
        L.sb_trws_destroy.argtypes = [ctypes.c_void_p]
        L.sb_trws_destroy.restype = None
        L.sb_trws_grid_ordering.argtypes = [c_int, c_int, _ip]
        L.sb_trws_plan_stats.argtypes = [c_int, c_int, c_int, c_int, POINTER(c_int64)]
        L.sb_grid_from_connectivity.argtypes = [c_int64, c_int64, _up, POINTER(c_int), POINTER(c_int)]
        vp = ctypes.c_void_p
        L.sb_trws_grid_create.argtypes = [c_int, c_int, c_int, c_int, c_double, POINTER(TrwsOptions), c_int, c_int, POINTER(vp)]
        L.sb_trws_grid_set_labels.argtypes = [vp, c_int, c_int, _dp, _dp, c_double, c_double]
        L.sb_trws_grid_set_weights.argtypes = [vp, _dp]
        L.sb_trws_grid_synth.argtypes = [vp, ctypes.c_uint64]
        L.sb_trws_grid_synth_window.argtypes = [vp, ctypes.c_uint64, c_int, c_int, c_int, c_int]
        L.sb_trws_grid_finalize.argtypes = [vp]
        L.sb_trws_grid_get_label.argtypes = [vp, c_int, _dp]
        L.sb_trws_grid_get_weights.argtypes = [vp, _dp]
        L.sb_trws_grid_reset.argtypes = [vp]
        L.sb_trws_grid_minimize.argtypes = [vp, c_double, c_double, _dp, _dp, _dp, POINTER(TrwsTiming)]
        L.sb_trws_grid_get_labels.argtypes = [vp, _dp]
        L.sb_trws_grid_ipc_export.argtypes = [vp, ctypes.c_char_p]
        L.sb_trws_grid_ipc_attach.argtypes = [vp, ctypes.c_char_p, ctypes.c_char_p]
        L.sb_trws_grid_pass.argtypes = [vp, c_int, c_int, _dp]
        L.sb_trws_grid_launch_pass.argtypes = [vp, c_int, c_int]
        L.sb_trws_grid_wait.argtypes = [vp, _dp, c_int, POINTER(c_int)]
        L.sb_trws_grid_attach_local.argtypes = [vp, vp, vp, c_int]
        L.sb_trws_grid_info.argtypes = [vp, POINTER(c_int64)]
        L.sb_trws_grid_latency_mode.argtypes = [vp, POINTER(c_int)]
        L.sb_trws_grid_counters.argtypes = [vp, _dp]
        L.sb_trws_grid_destroy.argtypes = [vp]
        L.sb_trws_grid_destroy.restype = None
        L.sb_trws_grid_plan_stats.argtypes = [c_int, c_int, c_int, c_int, POINTER(c_int64)]
        L.sb_trws_update_message.argtypes = [c_int, c_int, _dp, _dp, _dp, _dp, c_double, c_double, c_double, c_int, _dp, _dp]
        ip, i64, dbl = c_int, c_int64, c_double
        L.sb_rd_solve.argtypes = [c_int64, c_int64, _dp, _dp, _dp, _dp, _dp, _dp, _up, c_int, _dp, _dp, _dp, _dp]
        L.sb_binary_fusion_grid.argtypes = [ip, ip, ip, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                            ctypes.c_void_p, dbl, dbl, dbl, ip, ip, ctypes.c_void_p, _dp, _dp, _dp, _dp]
        L.sb_binary_fuse_until_convergence_grid.argtypes = [ip, ip, ip, ip, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                                            ctypes.c_void_p, ctypes.c_void_p, dbl, dbl, dbl, ip, ip,
                                                            POINTER(ctypes.c_int32), i64, ip, _dp, POINTER(c_int), _dp]
        L.sb_plane_from_disparity.argtypes = [ip, ip, ctypes.c_void_p, dbl, dbl, dbl, ip, ip, _dp, ctypes.c_void_p, _dp]
        L.sb_segpln_wta.argtypes = [ip, ip, ip, ip, _dp, _dp, ip, _dp, ip, dbl, dbl, _dp, _dp]
        L.sb_smooth_weights.argtypes = [ip, ip, POINTER(ctypes.c_uint32), dbl, dbl, dbl, _dp]
        L.sb_ncc_volume.argtypes = [ip, ip, ip, _dp, _dp, ip, _dp, ip, _dp]
        L.sb_ncc_vol_create.argtypes = [ip, ip, ip, _dp, _dp, ip, _dp, ip, POINTER(vp)]
        L.sb_ncc_vol_get.argtypes = [vp, _dp]
        L.sb_ncc_vol_best_disp.argtypes = [vp, _dp]
        L.sb_ncc_vol_sample.argtypes = [vp, _dp, dbl, ip, _dp]
        L.sb_ncc_vol_info.argtypes = [vp, _dp]
        L.sb_ncc_vol_destroy.argtypes = [vp]
        L.sb_ncc_vol_destroy.restype = None
        L.sb_ncc_best_disp.argtypes = [ip, ip, ip, _dp, _dp, _dp]
        L.sb_ncc_sample.argtypes = [ip, ip, ip, _dp, _dp, _dp, dbl, ip, _dp]
        L.sb_plane_disparity.argtypes = [i64, _dp, _dp, dbl, dbl, _dp]
        L.sb_interp2_linear.argtypes = [_dp, ip, ip, ip, _dp, _dp, i64, dbl, _dp]
        L.sb_photo_unary.argtypes = [ip, ip, ip, _dp, _dp, _dp, _dp, dbl, dbl, dbl, _dp]
        L.sb_pairwise_tables.argtypes = [ip, ip, ip, _dp, _dp, _dp, dbl, dbl, dbl, _dp, _dp, _dp, _dp]
        L.sb_fusion_positions.argtypes = [ip, ip, ip, _dp, dbl, dbl, _dp, _dp]
        L.sb_energy.argtypes = [ip, ip, ip, _dp, _dp, _dp, dbl, dbl, dbl, _dp]
        _lib = L
    return _lib


def check(rc):
    if rc != SB_OK:
        raise SbError(rc, lib().sb_last_error().decode("utf-8", "replace"))
