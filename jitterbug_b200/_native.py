"""ctypes binding of ``libjbsim.so`` (C-ABI declared in include/jbsim.h).

There is deliberately no fallback: if the CUDA library has not been built the loader raises, and if no CUDA device
is present ``jb_sim_create`` fails.  Nothing in this package imports the CPU oracle.
"""

from __future__ import annotations

import ctypes as C
import os

from .model import JbModelD, JbModelF

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("JB_LIB", os.path.join(_HERE, "_lib", "libjbsim.so"))   # JB_LIB: tuning variants


class JbSimConfig(C.Structure):
    _fields_ = [("num_envs", C.c_int), ("device", C.c_int), ("n_sub_steps", C.c_int), ("step_limit", C.c_int),
                ("random_pose", C.c_int), ("obs_stride", C.c_int), ("seed", C.c_uint64),
                ("env_id_offset", C.c_uint32), ("pad", C.c_uint32)]


class JbEncoderDesc(C.Structure):
    _fields_ = [("obs_dim", C.c_int), ("k_encode", C.c_int), ("n_layers", C.c_int), ("dims", C.c_int * 5),
                ("act", C.c_int * 4), ("vae", C.c_int), ("normalize", C.c_int), ("n_models", C.c_int),
                ("env_id_offset", C.c_uint32)]


AUG_NGEOM, AUG_NDELTA = 22, 31


class JbAugBase(C.Structure):
    _fields_ = [("model", JbModelD), ("density", C.c_double * AUG_NGEOM), ("follows_default", C.c_int * AUG_NGEOM)]


class JbAugSpec(C.Structure):
    _fields_ = [("modify_legs", C.c_int), ("modify_mass", C.c_int), ("modify_core1", C.c_int), ("modify_core2", C.c_int),
                ("modify_global_density", C.c_int), ("modify_gear", C.c_int),
                ("sd_legs", C.c_double * 3), ("sd_mass_pos", C.c_double * 3), ("sd_core1_density", C.c_double),
                ("sd_core2_density", C.c_double), ("sd_global_density", C.c_double), ("sd_gear", C.c_double)]


# every symbol include/jbsim.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "jb_sim_create": (C.c_int, [C.POINTER(_P), C.POINTER(JbSimConfig), C.POINTER(JbModelF), C.c_int, C.POINTER(C.c_int32)]),
    "jb_sim_create_indexed": (C.c_int, [C.POINTER(_P), C.POINTER(JbSimConfig), C.POINTER(JbModelF), C.c_int,
                                        C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "jb_sim_create_packed": (C.c_int, [C.POINTER(_P), C.POINTER(JbSimConfig), _P, C.c_int, C.POINTER(C.c_int32)]),
    "jb_sim_destroy": (None, [_P]),
    "jb_sim_reset": (C.c_int, [_P, _P, _P, _P]),
    "jb_sim_step": (C.c_int, [_P, _P, _P, _P, _P, _P]),
    "jb_sim_step_host": (C.c_int, [_P, _P, _P, _P, _P, _P]),
    "jb_sim_observe": (C.c_int, [_P, _P, _P, _P]),
    "jb_sim_heuristic_actions": (C.c_int, [_P, _P, _P, _P]),
    "jb_sim_get_state": (C.c_int, [_P, _P, _P, _P, _P]),
    "jb_sim_set_state": (C.c_int, [_P, _P, _P, _P, _P]),
    "jb_sim_get_target": (C.c_int, [_P, _P, _P]),
    "jb_sim_set_target": (C.c_int, [_P, _P, _P]),
    "jb_sim_get_step_count": (C.c_int, [_P, _P, _P]),
    "jb_sim_set_step_count": (C.c_int, [_P, _P, _P]),
    "jb_sim_get_contacts": (C.c_int, [_P, _P, _P, _P, C.c_int, _P]),
    "jb_sim_set_slot_order": (C.c_int, [_P, C.c_int, _P, _P]),
    "jb_sim_get_slot_order": (C.c_int, [_P, _P, _P]),
    "jb_sim_get_load": (C.c_int, [_P, _P, _P]),
    "jb_sim_self_clearance": (C.c_int, [_P, _P, C.POINTER(C.c_double), _P]),
    "jb_sim_get_counters": (C.c_int, [_P, C.POINTER(C.c_ulonglong)]),
    "jb_sim_get_phase_cycles": (C.c_int, [_P, C.POINTER(C.c_ulonglong)]),
    "jb_measure_fp32_peak": (C.c_int, [C.c_int, C.POINTER(C.c_double)]),
    "jb_packed_model_size": (C.c_int, []),
    "jb_augment_sample": (C.c_int, [C.POINTER(JbAugSpec), C.c_uint64, C.c_uint32, C.c_int, _P, _P]),
    "jb_augment_compile": (C.c_int, [C.POINTER(JbAugBase), _P, C.c_int, _P, _P, _P]),
    "jb_encoder_num_params": (C.c_int, [C.POINTER(JbEncoderDesc)]),
    "jb_encoder_out_dim": (C.c_int, [C.POINTER(JbEncoderDesc)]),
    "jb_encode_obs": (C.c_int, [C.POINTER(JbEncoderDesc), _P, _P, _P, C.c_int, C.c_int, _P, C.c_int, C.c_uint64,
                                C.c_uint64, _P]),
    "jb_sim_num_envs": (C.c_int, [_P]),
    "jb_sim_launch_count": (C.c_ulonglong, [_P]),
    "jb_sim_snapshot_size": (C.c_size_t, [_P]),
    "jb_sim_snapshot": (C.c_int, [_P, C.c_void_p, C.c_void_p]),
    "jb_sim_restore": (C.c_int, [_P, C.c_void_p, C.c_void_p]),
    "jb_sim_obs_stride": (C.c_int, [_P]),
    "jb_last_error": (C.c_char_p, []),
    "jb_version": (C.c_char_p, []),
}

_lib = None


def load() -> C.CDLL:
    """Load libjbsim.so; raises if the CUDA extension is missing (no CPU fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: the CUDA extension is not built. Run `python -m jitterbug_b200.build` "
                "(needs nvcc). jitterbug_b200 has no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise RuntimeError("libjbsim: " + load().jb_last_error().decode())
