"""ctypes binding of the C ABI in include/mppi_b200.h (libmppi_b200.so, built in-tree).

The product path: there is no Python/NumPy fallback -- if the shared library is missing, or no
CUDA device is present, construction fails loudly.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MPPI_B200_LIB") or os.path.join(_HERE, "lib", "libmppi_b200.so")   # env: experimental variants

MPPI_OK = 0
STATUS_NAMES = {0: "MPPI_OK", 1: "MPPI_ERR_INVALID", 2: "MPPI_ERR_CUDA", 3: "MPPI_ERR_NO_DEVICE",
                4: "MPPI_ERR_UNSUPPORTED", 5: "MPPI_ERR_STATE", 6: "MPPI_ERR_NONFINITE", 7: "MPPI_ERR_RETRY"}
MPPI_ERR_RETRY = 7

MODEL_DIFF_DRIVE, MODEL_UNICYCLE_EULER, MODEL_BICYCLE, MODEL_USER = 0, 1, 2, 3
WEIGHT_COST_TO_GO, WEIGHT_TOTAL_COST = 0, 1
PRECISION_F32, PRECISION_F64, PRECISION_MIXED = 0, 1, 2
ABI_VERSION = 1


class MppiParams(C.Structure):
    """mirror of `mppi_params` (include/mppi_b200.h)."""
    _fields_ = [
        ("struct_size", C.c_uint32), ("abi_version", C.c_uint32),
        ("K", C.c_int32), ("T", C.c_int32), ("model", C.c_int32), ("weighting", C.c_int32),
        ("precision", C.c_int32), ("device", C.c_int32),
        ("dt", C.c_double), ("q", C.c_double * 3), ("r", C.c_double * 4), ("p1", C.c_double * 3),
        ("sig", C.c_double * 4), ("noise_std", C.c_double * 2), ("lambda_", C.c_double),
        ("u_max", C.c_double * 2), ("wheel_radius", C.c_double), ("wheel_base", C.c_double),
        ("eps_floor", C.c_double), ("seed", C.c_uint64),
        ("k_offset", C.c_int64), ("k_total", C.c_int64), ("world_size", C.c_int32), ("rank", C.c_int32),
        ("stream", C.c_void_p), ("refine_margin", C.c_double),
    ]


class MppiUserModel(C.Structure):
    """mirror of `mppi_user_model`."""
    _fields_ = [("source", C.c_char_p), ("integrator", C.c_int32), ("wrap_theta", C.c_int32), ("has_cost", C.c_int32),
                ("kind", C.c_int32), ("speed_max", C.c_double), ("yaw_rate_max", C.c_double)]


class MppiTiming(C.Structure):
    """mirror of `mppi_timing`."""
    _fields_ = [
        ("step_ms", C.c_float), ("rollout_ms", C.c_float), ("reduce_ms", C.c_float), ("finalize_ms", C.c_float),
        ("launches", C.c_int32), ("steps", C.c_int32), ("refine_candidates", C.c_int32),
        ("refine_overflow", C.c_int32), ("refine_max_dev", C.c_double), ("refine_head_room", C.c_double),
    ]


class MppiError(RuntimeError):
    def __init__(self, status, where, text):
        RuntimeError.__init__(self, "%s failed: %s (%s)" % (where, STATUS_NAMES.get(status, status), text))
        self.status = status


_dp = C.POINTER(C.c_double)
_H = C.c_void_p

# name -> argtypes; every function returns mppi_status unless listed in _OTHER_RET
_SIGNATURES = {
    "mppi_default_params": [C.POINTER(MppiParams)],
    "mppi_create": [C.POINTER(MppiParams), C.POINTER(_H)],
    "mppi_create_user": [C.POINTER(MppiParams), C.POINTER(MppiUserModel), C.POINTER(_H)],
    "mppi_check_user_model": [C.POINTER(MppiUserModel)],
    "mppi_destroy": [_H],
    "mppi_reset": [_H],
    "mppi_set_goal": [_H, _dp],
    "mppi_set_sampling": [_H, _dp, C.c_double],
    "mppi_set_noise_std": [_H, _dp],
    "mppi_step": [_H, _dp, _dp, _dp],
    "mppi_get_nominal": [_H, _dp],
    "mppi_set_nominal": [_H, _dp],
    "mppi_get_last_update": [_H, _dp],
    "mppi_set_grid": [_H, C.POINTER(C.c_int8), C.c_int32, C.c_int32, C.c_double, C.c_double, C.c_double, C.c_double],
    "mppi_update_grid": [_H, C.POINTER(C.c_int8), C.c_int32, C.c_int32, C.c_int32, C.c_int32],
    "mppi_clear_grid": [_H],
    "mppi_set_noise": [_H, _dp],
    "mppi_use_philox": [_H, C.c_uint64],
    "mppi_get_noise": [_H, _dp],
    "mppi_set_capture": [_H, C.c_int32],
    "mppi_get_cost_to_go": [_H, _dp],
    "mppi_cost_to_go": [_H, _dp, _dp, _dp, _dp, _dp],
    "mppi_update_action": [_H, _dp, _dp, _dp, _dp],
    "mppi_perform_action": [_H, _dp, _dp, _dp],
    "mppi_model_step": [_H, _dp, _dp, C.c_int32, _dp],
    "mppi_step_local": [_H, _dp],
    "mppi_exchange_buffers": [_H, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)],
    "mppi_step_finish": [_H, _dp, _dp],
    "mppi_read_record": [_H, _dp],
    "mppi_write_gather": [_H, _dp],
    "mppi_p2p_export": [_H, C.c_void_p],
    "mppi_p2p_connect": [_H, C.c_void_p],
    "mppi_bench": [_H, _dp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(MppiTiming)],
    "mppi_last_stats": [_H, C.POINTER(MppiTiming)],
    "mppi_measure_fp32_peak": [C.c_int32, _dp, _dp],
    "mppi_io_bytes": [_H, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)],
    "mppi_launch_info": [_H, C.POINTER(C.c_int32)],
    "mppi_debug_flush_l2": [_H],
    "mppi_debug_reduce_timestamps": [_H, C.POINTER(C.c_uint64)],
    "mppi_debug_rollout_timestamps": [_H, C.POINTER(C.c_uint64), C.c_size_t],
    "mppi_debug_host_timing": [_H, _dp],
    "mppi_last_error": [],
    "mppi_version": [],
    "mppi_device_count": [],
}
_OTHER_RET = {"mppi_last_error": C.c_char_p, "mppi_version": C.c_char_p, "mppi_device_count": C.c_int32}

_lib = None


def exported_symbols():
    """Names the header declares (used by the CPU test that the .so exports all of them)."""
    return sorted(_SIGNATURES)


def load():
    """dlopen libmppi_b200.so and type its entry points. Raises if the extension is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("motion_planning_b200: %s is missing -- build it with "
                          "`python -m motion_planning_b200.build` (there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, args in _SIGNATURES.items():
        fn = getattr(lib, name)            # AttributeError if the symbol is not exported
        fn.argtypes = args
        fn.restype = _OTHER_RET.get(name, C.c_int)
    _lib = lib
    return lib


def check(status, where):
    if status != MPPI_OK:
        raise MppiError(status, where, load().mppi_last_error().decode("utf-8", "replace"))


def dptr(a):
    """float64 C-contiguous ndarray -> double*."""
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_dp)


def f64(a, shape=None):
    out = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
    if shape is not None:
        out = out.reshape(shape)
    return out
