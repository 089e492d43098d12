"""ctypes binding of libmcl_b200.so (the C ABI declared in include/mcl.h).

Loading is explicit and loud: if the shared library is missing this module raises -- there is no
Python/NumPy fallback for any filter arithmetic.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# MCL_LIB selects an experimental build of the SAME sources (tools/, kernel tuning); the default is the product library
LIB_PATH = os.environ.get("MCL_LIB") or os.path.join(HERE, "libmcl_b200.so")

MCL_ABI_VERSION = 3

MCL_OK = 0
MCL_ERR_ARG = -1
MCL_ERR_CUDA = -2
MCL_ERR_STATE = -3
MCL_ERR_NOMEM = -4
MCL_ERR_NODEVICE = -5

MCL_LIK_PAIR = 0
MCL_LIK_MOMENT = 1
MCL_LIK_RB = 2
MCL_LIK_RB_FAN = 3
MCL_LIK_PAIR_F32 = 4
MCL_LIK_RB_F32 = 5
MCL_MOTION_A = 0
MCL_MOTION_B = 1

MCL_FORCE_RESAMPLE = 1
MCL_NO_RESAMPLE = 2
MCL_PHILOX_NOISE = 4

MCL_IPC_HANDLE_BYTES = 64
MCL_MAX_SHARDS = 16
MCL_STAGE_A, MCL_STAGE_B, MCL_STAGE_C, MCL_STAGE_D, MCL_STAGE_ALL = 1, 2, 4, 8, 15


class MclError(RuntimeError):
    def __init__(self, status, where, detail=""):
        self.status = status
        super().__init__(f"{where}: status {status} ({detail})")


class mcl_config(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32),
        ("device", C.c_int32),
        ("n_particles", C.c_int64),
        ("n_trials", C.c_int32),
        ("max_landmarks", C.c_int32),
        ("dt", C.c_double),
        ("vel", C.c_double),
        ("yaw_rate", C.c_double),
        ("q_std", C.c_double * 3),
        ("sigma_obs", C.c_double * 2),
        ("ess_threshold", C.c_double),
        ("motion_b_alpha", C.c_double * 6),
        ("likelihood", C.c_int32),
        ("motion", C.c_int32),
        ("seed", C.c_uint64),
        ("n_global", C.c_int64),
        ("slot_base", C.c_int64),
        ("fan_range", C.c_double),
        ("fan_half_angle", C.c_double),
        ("fan_miss_chi2", C.c_double),
        ("underflow_rule", C.c_int32),
        ("reserved0", C.c_int32),
    ]


class mcl_stats(C.Structure):
    _fields_ = [
        ("log_norm", C.c_double),
        ("ess", C.c_double),
        ("max_weight", C.c_double),
        ("max_idx", C.c_int64),
        ("est", C.c_double * 3),
        ("mean", C.c_double * 3),
        ("cov", C.c_double * 9),
        ("sum_w", C.c_double),
        ("sum_w2", C.c_double),
        ("max_logw", C.c_double),
        ("raw", C.c_double * 12),
        ("ref", C.c_double * 3),
        ("resampled", C.c_int32),
        ("step", C.c_int32),
        ("derived", C.c_int32),
        ("underflow", C.c_int32),
        ("abs_shift", C.c_double),
    ]


class mcl_shard_plan(C.Structure):
    _fields_ = [
        ("k_offset", C.c_uint64),
        ("ofs", C.c_double),
        ("first_slot", C.c_int64),
        ("n_slots", C.c_int64),
        ("totals", C.c_uint64 * MCL_MAX_SHARDS),
        ("win_lo", C.c_uint32),
        ("win_hi", C.c_uint32),
        ("gate", C.c_int32),
        ("resampled", C.c_int32),
        ("error", C.c_int32),
        ("epoch", C.c_uint32),
    ]


_P = C.POINTER
_vp = C.c_void_p

# name -> (restype, argtypes); every symbol include/mcl.h declares
PROTOTYPES = {
    "mcl_abi_version": (C.c_int, []),
    "mcl_status_string": (C.c_char_p, [C.c_int]),
    "mcl_last_error": (C.c_char_p, [_vp]),
    "mcl_config_default": (C.c_int, [_P(mcl_config), C.c_double]),
    "mcl_workspace_bytes": (C.c_int, [_P(mcl_config), _P(C.c_size_t)]),
    "mcl_create": (C.c_int, [_P(mcl_config), _vp, C.c_size_t, _P(_vp)]),
    "mcl_destroy": (C.c_int, [_vp]),
    "mcl_stride": (C.c_int64, [_vp]),
    "mcl_set_landmarks": (C.c_int, [_vp, _vp, C.c_int32, _vp]),
    "mcl_init_particles": (C.c_int, [_vp, _vp, _vp]),
    "mcl_set_particles": (C.c_int, [_vp, _vp, _vp, _vp, _vp, C.c_int, _vp]),
    "mcl_get_particles": (C.c_int, [_vp, _vp, _vp, _vp, _vp, C.c_int, _vp]),
    "mcl_resample": (C.c_int, [_vp, _vp, C.c_uint32, _vp]),
    "mcl_predict_weight": (C.c_int, [_vp, _vp, _vp, C.c_uint32, _vp]),
    "mcl_step": (C.c_int, [_vp, _vp, _vp, _vp, C.c_uint32, _vp]),
    "mcl_step_async": (C.c_int, [_vp, _vp, _vp, _vp, C.c_uint32, _vp]),
    "mcl_step_dev": (C.c_int, [_vp, _vp, _vp, _vp, C.c_uint32, _vp]),
    "mcl_pose_moments": (C.c_int, [_vp, _vp]),
    "mcl_get_stats": (C.c_int, [_vp, _P(mcl_stats), _vp]),
    "mcl_stats_dev": (C.c_int, [_vp, _P(_vp)]),
    "mcl_get_stats_begin": (C.c_int, [_vp, _vp]),
    "mcl_get_stats_end": (C.c_int, [_vp, _P(mcl_stats)]),
    "mcl_sync_stats": (C.c_int, [_vp, _vp]),
    "mcl_get_weights": (C.c_int, [_vp, _vp, _vp]),
    "mcl_get_qweights": (C.c_int, [_vp, _vp, _vp]),
    "mcl_get_ancestors": (C.c_int, [_vp, _vp, _vp]),
    "mcl_profile": (C.c_int, [_vp, C.c_int]),
    "mcl_profile_read": (C.c_int, [_vp, _P(C.c_double)]),
    "mcl_profile_read_peer": (C.c_int, [_vp, _P(C.c_double)]),
    "mcl_launch_count": (C.c_uint64, [_vp]),
    "mcl_debug_timeline": (C.c_int, [_vp, _vp, C.c_int64, _P(C.c_int32)]),
    "mcl_shard_total": (C.c_int, [_vp, _vp, _vp]),
    "mcl_shard_expand": (C.c_int, [_vp, C.c_uint64, C.c_double, _vp, _vp, _vp, C.c_int64,
                                   _P(C.c_int64), _P(C.c_int64), _vp]),
    "mcl_shard_step": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, C.c_int64, C.c_int64, C.c_uint32, _vp]),
    "mcl_shard_set_global": (C.c_int, [_vp, _P(mcl_stats), _vp]),
    "mcl_shard_merge": (C.c_int, [_vp, _vp, C.c_int, _vp]),
    "mcl_restore": (C.c_int, [_vp, _P(mcl_stats), C.c_uint32, _vp]),
    "mcl_shard_peer_alloc": (C.c_int, [_vp, C.c_int, C.c_int, _vp]),
    "mcl_shard_peer_connect": (C.c_int, [_vp, _vp, _P(_vp)]),
    "mcl_shard_step_peer": (C.c_int, [_vp, _vp, _vp, _vp, C.c_uint32, C.c_uint32, _vp]),
    "mcl_shard_sync_peer": (C.c_int, [_vp, C.c_uint32, _vp]),
    "mcl_shard_get_plan": (C.c_int, [_vp, _P(mcl_shard_plan), _vp]),
}

_lib = None


def load():
    """Load libmcl_b200.so once; raise if it is absent or its ABI differs."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python slam-robot_simu_b200/build.py` "
            "(the engine is CUDA-only; there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.mcl_abi_version() != MCL_ABI_VERSION:
        raise ImportError(f"{LIB_PATH}: ABI {lib.mcl_abi_version()} != expected {MCL_ABI_VERSION}")
    _lib = lib
    return lib


def check(status, where, handle=None):
    if status == MCL_OK:
        return
    lib = load()
    detail = lib.mcl_last_error(handle)
    detail = detail.decode() if detail else ""
    kind = lib.mcl_status_string(status).decode()
    raise MclError(status, where, f"{kind}: {detail}" if detail else kind)
