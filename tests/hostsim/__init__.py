"""TEST-ONLY ctypes wrapper of tests/hostsim/hostsim.cpp (host build of the GPU sub-step algorithm)."""
import ctypes as C
import os
import subprocess

import numpy as np

from jitterbug_b200 import model as jbmodel

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libhostsim.so")
_SRC = [os.path.join(_HERE, "hostsim.cpp"),
        os.path.join(_HERE, "ref_thread_per_env.cuh"),
        os.path.join(_HERE, "..", "..", "jitterbug_b200", "csrc", "jb_model.h")]


def _diag_struct(real):
    class Diag(C.Structure):
        _fields_ = [("M", (real * 15) * 15), ("bias", real * 15), ("fsmooth", real * 15), ("asmooth", real * 15),
                    ("qacc", real * 15), ("fcon", real * 15), ("Rw", real * 9),
                    ("ncon", C.c_int), ("geom", C.c_int * 64), ("dist", real * 64), ("cpos", (real * 3) * 64),
                    ("newton_iters", C.c_int), ("overflow", C.c_int)]
    return Diag


DiagD, DiagF = _diag_struct(C.c_double), _diag_struct(C.c_float)
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(s) for s in _SRC):
            subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", "-std=c++14", "-ffp-contract=off",
                                   "-o", _SO, _SRC[0]])
        _lib = C.CDLL(_SO)
        assert _lib.hs_sizeof_model_f32() == C.sizeof(jbmodel.JbModelF)
        assert _lib.hs_sizeof_model_f64() == C.sizeof(jbmodel.JbModelD)
        assert _lib.hs_sizeof_diag_f64() == C.sizeof(DiagD), (_lib.hs_sizeof_diag_f64(), C.sizeof(DiagD))
        assert _lib.hs_sizeof_diag_f32() == C.sizeof(DiagF)
    return _lib


def substeps(model_struct, qpos, qvel, warm, ctrl, nsub=1, double=True, want_diag=False):
    """Advance one env by ``nsub`` sub-steps with the host build.  Returns (qpos, qvel, warm, iters, diag)."""
    L = lib()
    dt = np.float64 if double else np.float32
    real = C.c_double if double else C.c_float
    q = np.array(qpos, dtype=dt).copy(); v = np.array(qvel, dtype=dt).copy(); w = np.array(warm, dtype=dt).copy()
    diag = (DiagD if double else DiagF)() if want_diag else None
    fn = L.hs_substeps_f64 if double else L.hs_substeps_f32
    fn.restype = C.c_int
    it = fn(C.byref(model_struct), q.ctypes.data_as(C.POINTER(real)), v.ctypes.data_as(C.POINTER(real)),
            w.ctypes.data_as(C.POINTER(real)), real(ctrl), C.c_int(nsub), C.byref(diag) if want_diag else None)
    return q, v, w, it, diag


# ------------------------------------------------------------------------------------------------ quad emulation
_QSO = os.path.join(_HERE, "libquadsim.so")
_QSRC = [os.path.join(_HERE, "quadsim.cpp")] + [os.path.join(_HERE, "..", "..", "jitterbug_b200", "csrc", f)
                                                 for f in ("jb_quad.cuh", "jb_math.cuh", "jb_packed.h", "jb_model.h", "jb_task.cuh")]
_qlib = None


def qlib():
    global _qlib
    if _qlib is None:
        if not os.path.exists(_QSO) or os.path.getmtime(_QSO) < max(os.path.getmtime(s) for s in _QSRC):
            subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", "-std=c++14", "-ffp-contract=off",
                                   "-o", _QSO, _QSRC[0]])
        _qlib = C.CDLL(_QSO)
    return _qlib


def pack(model_struct, double=True):
    L = qlib()
    n = L.hq_packed_size()
    out = np.zeros(n, dtype=np.float64 if double else np.float32)
    (L.hq_pack_f64 if double else L.hq_pack_f32)(C.byref(model_struct), out.ctypes.data_as(C.c_void_p))
    return out


def quad_substeps(packed, qpos, qvel, warm, ctrl, nsub=1, forward_only=False, want_contacts=False):
    """Advance one env with the fiber-emulated quad kernel.  Returns (qpos, qvel, warm, iters, ncon, contacts, overflow)."""
    L = qlib()
    double = packed.dtype == np.float64
    dt = packed.dtype
    real = C.c_double if double else C.c_float
    q = np.array(qpos, dtype=dt).copy(); v = np.array(qvel, dtype=dt).copy(); w = np.array(warm, dtype=dt).copy()
    ncon, ovf = C.c_int(), C.c_int()
    geom = np.full(64, -1, dtype=np.int32); dist = np.zeros(64, dtype=dt)      # >= 4 * JBQ_CAP
    fn = L.hq_substeps_f64 if double else L.hq_substeps_f32
    fn.restype = C.c_int
    P = C.c_void_p
    it = fn(packed.ctypes.data_as(P), q.ctypes.data_as(P), v.ctypes.data_as(P), w.ctypes.data_as(P), real(ctrl),
            C.c_int(nsub), C.c_int(int(forward_only)), C.byref(ncon),
            geom.ctypes.data_as(P) if want_contacts else None, dist.ctypes.data_as(P) if want_contacts else None,
            C.byref(ovf))
    if it < 0:
        raise RuntimeError("replicated quad state diverged between lanes")
    cons = sorted((int(g), float(d)) for g, d in zip(geom, dist) if g >= 0) if want_contacts else None
    return q, v, w, it, ncon.value, cons, ovf.value


_DEFAULT_PACKED = None


def default_packed_f32():
    global _DEFAULT_PACKED
    if _DEFAULT_PACKED is None:
        from jitterbug_b200 import mjcf
        _DEFAULT_PACKED = pack(jbmodel.fill_struct(jbmodel.flatten_model(mjcf.compile_xml())), double=False)
    return _DEFAULT_PACKED


def reset_states(seed, env_ids, episode, task, random_pose, packed=None):
    """Host run of the device reset code: rows of (quat w,x,y,z, target x, y, qw, qz)."""
    L = qlib()
    P = default_packed_f32() if packed is None else packed
    ids = np.ascontiguousarray(env_ids, dtype=np.uint32)
    out = np.zeros((len(ids), 8), dtype=np.float32)
    L.hq_reset_states(P.ctypes.data_as(C.c_void_p), C.c_ulonglong(seed), ids.ctypes.data_as(C.c_void_p), C.c_int(len(ids)),
                      C.c_uint(episode), C.c_int(task), C.c_int(random_pose), out.ctypes.data_as(C.c_void_p))
    return out


def reset_from_uniforms(u, task, random_pose, packed):
    """Host run of the device reset mapping from given uniforms u[n, 6] (dtype follows ``packed``): rows of
    (quat w,x,y,z, target x, y, qw, qz)."""
    L = qlib()
    dt = packed.dtype
    uu = np.ascontiguousarray(u, dtype=dt)
    out = np.zeros((len(uu), 8), dtype=dt)
    fn = L.hq_reset_from_uniforms_f64 if dt == np.float64 else L.hq_reset_from_uniforms_f32
    fn(packed.ctypes.data_as(C.c_void_p), uu.ctypes.data_as(C.c_void_p), C.c_int(len(uu)), C.c_int(task), C.c_int(random_pose),
       out.ctypes.data_as(C.c_void_p))
    return out


def obs_reward(qpos, qvel, task, target, packed=None):
    L = qlib()
    P = default_packed_f32() if packed is None else packed
    q = np.ascontiguousarray(qpos, dtype=np.float32); v = np.ascontiguousarray(qvel, dtype=np.float32)
    t = np.ascontiguousarray(target, dtype=np.float32)
    obs = np.zeros(19, dtype=np.float32); rew = C.c_float()
    n = L.hq_obs_reward(P.ctypes.data_as(C.c_void_p), q.ctypes.data_as(C.c_void_p), v.ctypes.data_as(C.c_void_p), C.c_int(task),
                        t.ctypes.data_as(C.c_void_p), obs.ctypes.data_as(C.c_void_p), C.byref(rew))
    return obs[:n].copy(), rew.value


def heuristic(task, obs):
    L = qlib()
    L.hq_heuristic.restype = C.c_float
    o = np.zeros(19, dtype=np.float32); o[:len(obs)] = obs
    return float(L.hq_heuristic(C.c_int(task), o.ctypes.data_as(C.c_void_p)))


# ------------------------------------------------------------------------------------------------ model compiler
_ASO = os.path.join(_HERE, "libaugsim.so")
_ASRC = [os.path.join(_HERE, "augsim.cpp")] + [os.path.join(_HERE, "..", "..", "jitterbug_b200", "csrc", f)
                                                for f in ("jb_augment.cuh", "jb_packed.h", "jb_model.h")] + \
    [os.path.join(_HERE, "..", "..", "include", "jbsim.h")]
_alib = None


def augment(base, deltas):
    """Host build of the device model compiler: deltas [n, 31] -> (packed f32 [n, P], JbModelF array, JbModelD array)."""
    global _alib
    from jitterbug_b200 import _native
    if _alib is None:
        if not os.path.exists(_ASO) or os.path.getmtime(_ASO) < max(os.path.getmtime(s) for s in _ASRC):
            subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", "-std=c++14", "-ffp-contract=off",
                                   "-o", _ASO, _ASRC[0]])
        _alib = C.CDLL(_ASO)
        assert _alib.ha_sizeof_base() == C.sizeof(_native.JbAugBase)
    d = np.ascontiguousarray(deltas, dtype=np.float64)
    n = d.shape[0]
    packed = np.zeros((n, qlib().hq_packed_size()), dtype=np.float32)
    mf = (jbmodel.JbModelF * n)()
    md = (jbmodel.JbModelD * n)()
    _alib.ha_augment(C.byref(base), d.ctypes.data_as(C.c_void_p), C.c_int(n), packed.ctypes.data_as(C.c_void_p), mf, md)
    return packed, mf, md


def quad_debug_counts():
    """(line-search iterations, Newton trips, contacts moved between lanes by the spread step) since the library loaded."""
    out = (C.c_longlong * 3)()
    qlib().hq_debug_counts(out)
    return tuple(int(x) for x in out)
