"""ctypes binding of include/ipp_b200.h.  No CPU fallback: a missing library or device is an error."""
import ctypes as C
import os

from . import build as _build

MAX_AGENTS = 8
MAX_ALT = 8
MAX_LATTICE = 128
N_ACTIONS = 6
FLAG_QUADS = 640


class IppConfig(C.Structure):
    _fields_ = [
        ("gx", C.c_int32), ("gy", C.c_int32), ("map_stride", C.c_int32), ("gt_stride", C.c_int32),
        ("code_stride", C.c_int32), ("n_seg", C.c_int32),
        ("px", C.c_int32), ("py", C.c_int32), ("n_alt", C.c_int32),
        ("n_agents", C.c_int32), ("n_envs", C.c_int32), ("spacing", C.c_int32),
        ("min_altitude", C.c_int32), ("max_altitude", C.c_int32),
        ("x_dim_m", C.c_int32), ("y_dim_m", C.c_int32), ("budget", C.c_int32),
        ("seed", C.c_uint32), ("comm_d2_max", C.c_int32), ("fail_thresh24", C.c_uint32),
        ("prior", C.c_float), ("k_out", C.c_float), ("p_min", C.c_float), ("p_max", C.c_float),
        ("o_min", C.c_float), ("o_max", C.c_float),
        ("radius_x", C.c_int32 * MAX_ALT), ("radius_y", C.c_int32 * MAX_ALT),
        ("k_hi", C.c_float * MAX_ALT), ("k_lo", C.c_float * MAX_ALT),
        ("y_hi", C.c_float * MAX_ALT), ("y_lo", C.c_float * MAX_ALT),
        ("flip_thresh", C.c_uint32 * MAX_ALT),
        ("cell_x", C.c_int32 * MAX_LATTICE), ("cell_y", C.c_int32 * MAX_LATTICE),
        ("fix_range", C.c_int32), ("comm_d2_table", C.c_int32 * 4),
        ("n_meas", C.c_int32), ("meas_y", C.c_float * 16), ("meas_ly", C.c_float * 16),
        ("l_prior", C.c_double),
    ]


class IppState(C.Structure):
    _fields_ = [
        ("local_maps", C.c_void_p), ("global_map", C.c_void_p),
        ("ground_truth", C.c_void_p), ("episodes", C.c_void_p), ("meas_codes", C.c_void_p),
        ("map_flags", C.c_void_p),
    ]


class IppStepIO(C.Structure):
    _fields_ = [
        ("pos_in", C.c_void_p), ("pos_out", C.c_void_p), ("actions_in", C.c_void_p),
        ("probs_in", C.c_void_p), ("greedy", C.c_int32), ("actions_out", C.c_void_p),
        ("mask_out", C.c_void_p), ("comm_out", C.c_void_p), ("reward_rel", C.c_void_p),
        ("reward_abs", C.c_void_p), ("stuck_out", C.c_void_p),
    ]


EXPORTS = [
    "ipp_status_string", "ipp_last_error", "ipp_version", "ipp_create", "ipp_destroy", "ipp_scratch_bytes",
    "ipp_set_step_variant", "ipp_get_step_variant",
    "ipp_reset", "ipp_step", "ipp_run_steps", "ipp_step_host", "ipp_step_phases", "ipp_observe", "ipp_act", "ipp_features_actor", "ipp_features_critic",
    "ipp_export_beliefs", "ipp_ig_plan", "ipp_eval_metrics", "ipp_project_fov", "ipp_measure", "ipp_update_cells",
    "ipp_shannon_entropy", "ipp_fuse_map", "ipp_utility_reward",
]

_lib = None


class IppError(RuntimeError):
    pass


def library_path():
    return _build.LIB


def load():
    """Load (building first if the sources changed and nvcc is available)."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.build()
    if not os.path.exists(path):
        raise IppError("CUDA library %s is missing; run `python -m ipp_marl_b200.build`" % path)
    lib = C.CDLL(path)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    lib.ipp_status_string.restype = C.c_char_p
    lib.ipp_status_string.argtypes = [C.c_int]
    lib.ipp_last_error.restype = C.c_char_p
    lib.ipp_last_error.argtypes = [vp]
    lib.ipp_version.restype = C.c_int
    lib.ipp_create.argtypes = [C.POINTER(IppConfig), C.POINTER(vp)]
    lib.ipp_destroy.argtypes = [vp]
    lib.ipp_scratch_bytes.restype = i64
    lib.ipp_scratch_bytes.argtypes = [vp]
    lib.ipp_set_step_variant.argtypes = [vp, i32]
    lib.ipp_get_step_variant.argtypes = [vp]
    lib.ipp_reset.argtypes = [vp, C.POINTER(IppState), vp, vp]
    for name in ("ipp_step", "ipp_observe", "ipp_act"):
        getattr(lib, name).argtypes = [vp, C.POINTER(IppState), i32, C.POINTER(IppStepIO), vp]
    lib.ipp_run_steps.argtypes = [vp, C.POINTER(IppState), i32, vp, i32, i32, C.POINTER(IppStepIO), vp]
    lib.ipp_features_actor.argtypes = [vp, C.POINTER(IppState), i32, C.POINTER(IppStepIO), vp, vp]
    lib.ipp_features_critic.argtypes = [vp, C.POINTER(IppState), i32, vp, vp, vp, vp, vp]
    lib.ipp_step_host.argtypes = [vp, C.POINTER(IppState), i32, C.POINTER(IppStepIO), vp, vp, vp, vp, vp, vp]
    lib.ipp_export_beliefs.argtypes = [vp, C.POINTER(IppState), vp, vp, vp]
    lib.ipp_ig_plan.argtypes = [vp, C.POINTER(IppState), vp, i32, vp, vp, vp, vp, vp]
    lib.ipp_eval_metrics.argtypes = [vp, C.POINTER(IppState), vp, vp, vp]
    lib.ipp_step_phases.argtypes = [vp, C.POINTER(IppState), i32, C.POINTER(IppStepIO), i32, vp]
    lib.ipp_project_fov.argtypes = [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]
    lib.ipp_measure.argtypes = [vp, vp, C.POINTER(i32), i32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_float,
                                C.c_float, vp]
    lib.ipp_update_cells.argtypes = [vp, vp, i32, vp, i32, i32, i64, vp]
    lib.ipp_shannon_entropy.argtypes = [vp, vp, i32, i64, vp]
    lib.ipp_fuse_map.argtypes = [vp, vp, vp, i32, i64, vp]
    lib.ipp_utility_reward.argtypes = [vp, vp, i32, vp, i32, i64, vp]
    _lib = lib
    return lib


def check(lib, handle, status, what):
    if status != 0:
        msg = lib.ipp_status_string(status).decode()
        detail = lib.ipp_last_error(handle).decode() if handle else ""
        raise IppError("%s failed: %s %s" % (what, msg, detail))
