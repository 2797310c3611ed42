"""Flatten a compiled Jitterbug model into the device layout of ``csrc/jb_model.h``.

The kernels work in the root body's frame at the home configuration (product of exponentials), so the
MuJoCo-style local-coordinate tree from :mod:`jitterbug_b200.mjcf` is re-expressed here once, on the host, in
float64; the C-ABI struct is then filled in float32.  The structure of the Jitterbug family (reference:
jitterbug_dmc/jitterbug.xml:37-110 and its augmented variants, jitterbug_dmc/augmented_jitterbug.py:95-267) is
fixed: a free root body with two boxes, a cylinder and an ellipsoid; four two-link legs (cylinder + sphere on
each link, sprung hinges); one motor-driven rotor (cylinder + ellipsoid); a floor plane.  Anything else raises.
"""

from __future__ import annotations

import ctypes as C
from typing import Dict

import numpy as np

from . import mjcf

NLEG, NSLOT = 4, 10


def _geom_struct(real):
    class Geom(C.Structure):
        _fields_ = [("center", real * 3), ("rot", real * 9), ("size", real * 3), ("rbound", real),
                    ("geom_id", C.c_int), ("pad", C.c_int * 3)]
    return Geom


def _model_struct(real, geom):
    class Model(C.Structure):
        _fields_ = [
            ("timestep", real), ("gravity", real),
            ("m0", real), ("c0", real * 3), ("I0", real * 6),
            ("p1", (real * 3) * NLEG), ("a1", (real * 3) * NLEG),
            ("p2", (real * 3) * NLEG), ("a2", (real * 3) * NLEG),
            ("mu", real * NLEG), ("cu", (real * 3) * NLEG), ("Iu", (real * 6) * NLEG),
            ("ml", real * NLEG), ("cl", (real * 3) * NLEG), ("Il", (real * 6) * NLEG),
            ("stiff", (real * 2) * NLEG), ("damp", (real * 2) * NLEG),
            ("pm", real * 3), ("am", real * 3), ("mm", real), ("cm", real * 3), ("Im", real * 6),
            ("damp_m", real),
            ("gear", real), ("gain", real), ("bias_v", real), ("ctrl_lo", real), ("ctrl_hi", real),
            ("invw", real * NSLOT), ("friction", real), ("solref_K", real), ("solref_B", real),
            ("solimp", real * 5), ("impratio", real),
            ("g_box", geom * 2), ("g_rcyl", geom), ("g_rell", geom),
            ("g_ucyl", geom * NLEG), ("g_knee", geom * NLEG), ("g_lcyl", geom * NLEG), ("g_foot", geom * NLEG),
            ("g_mcyl", geom), ("g_mell", geom),
            ("meaninertia", real), ("root_pos0", real * 3), ("target_pos0", real * 3), ("pad_", real),
        ]
    return Model


JbGeomF = _geom_struct(C.c_float)
JbModelF = _model_struct(C.c_float, JbGeomF)
JbGeomD = _geom_struct(C.c_double)
JbModelD = _model_struct(C.c_double, JbGeomD)


def _sym6(m):
    return np.array([m[0, 0], m[1, 1], m[2, 2], m[0, 1], m[0, 2], m[1, 2]])


def flatten_model(cm: mjcf.CompiledModel) -> Dict[str, object]:
    """Compiled model -> dict with the field names of ``JbModel`` (float64 values)."""
    fk = mjcf.forward_kinematics(cm, cm.qpos0)
    free = [j for j in range(cm.njnt) if cm.jnt_type[j] == mjcf.JNT_FREE]
    if len(free) != 1:
        raise ValueError("expected exactly one free joint")
    root = int(cm.jnt_bodyid[free[0]])
    R0, x0 = fk["xmat"][root], fk["xpos"][root]

    def loc(p):
        return R0.T @ (np.asarray(p) - x0)

    def locdir(v):
        return R0.T @ np.asarray(v)

    def body_inertial(b):
        iw = fk["ximat"][b] @ np.diag(cm.body_inertia[b]) @ fk["ximat"][b].T
        return float(cm.body_mass[b]), loc(fk["xipos"][b]), _sym6(R0.T @ iw @ R0)

    def geom(g, want_type):
        if cm.geom_type[g] != want_type:
            raise ValueError(f"geom {g} ({cm.geom_names[g]}): unexpected type {cm.geom_type[g]}")
        size = cm.geom_size[g].copy()
        if want_type == mjcf.GEOM_SPHERE:
            rb = size[0]
        elif want_type == mjcf.GEOM_CYLINDER:
            rb = np.hypot(size[0], size[1])
        elif want_type == mjcf.GEOM_BOX:
            rb = np.linalg.norm(size)
        else:
            rb = size.max()
        return dict(center=loc(fk["geom_xpos"][g]), rot=(R0.T @ fk["geom_xmat"][g]).reshape(9), size=size,
                    rbound=float(rb), geom_id=int(g))

    def geoms_of(b, types):
        gs = [g for g in range(cm.ngeom) if cm.geom_bodyid[g] == b]
        if [int(cm.geom_type[g]) for g in gs] != list(types):
            raise ValueError(f"body {cm.body_names[b]}: geoms {[int(cm.geom_type[g]) for g in gs]} != {list(types)}")
        return gs

    children = [b for b in range(1, cm.nbody) if cm.body_parentid[b] == root]
    legs, rotor = [], None
    for b in children:
        kids = [c for c in range(1, cm.nbody) if cm.body_parentid[c] == b]
        if len(kids) == 1:
            legs.append((b, kids[0]))
        elif len(kids) == 0:
            rotor = b
    if len(legs) != NLEG or rotor is None:
        raise ValueError("expected four two-link legs and one rotor under the root body")

    out: Dict[str, object] = {}
    out["timestep"] = float(cm.timestep)
    if cm.gravity[0] != 0 or cm.gravity[1] != 0:
        raise ValueError("gravity must point along -z")
    out["gravity"] = float(-cm.gravity[2])
    out["m0"], out["c0"], out["I0"] = body_inertial(root)
    gs = geoms_of(root, [mjcf.GEOM_BOX, mjcf.GEOM_BOX, mjcf.GEOM_CYLINDER, mjcf.GEOM_ELLIPSOID])
    out["g_box"] = [geom(gs[0], mjcf.GEOM_BOX), geom(gs[1], mjcf.GEOM_BOX)]
    out["g_rcyl"] = geom(gs[2], mjcf.GEOM_CYLINDER)
    out["g_rell"] = geom(gs[3], mjcf.GEOM_ELLIPSOID)

    for key in ("p1", "a1", "p2", "a2", "cu", "cl"):
        out[key] = np.zeros((NLEG, 3))
    out["Iu"], out["Il"] = np.zeros((NLEG, 6)), np.zeros((NLEG, 6))
    out["mu"], out["ml"] = np.zeros(NLEG), np.zeros(NLEG)
    out["stiff"], out["damp"] = np.zeros((NLEG, 2)), np.zeros((NLEG, 2))
    for k in ("g_ucyl", "g_knee", "g_lcyl", "g_foot"):
        out[k] = []
    invw = np.zeros(NSLOT)
    invw[0] = cm.body_invweight0[root, 0]
    for k, (bu, bl) in enumerate(legs):
        ju, jl = int(cm.body_jntadr[bu]), int(cm.body_jntadr[bl])
        if cm.body_jntnum[bu] != 1 or cm.body_jntnum[bl] != 1 or cm.jnt_type[ju] != mjcf.JNT_HINGE \
                or cm.jnt_type[jl] != mjcf.JNT_HINGE:
            raise ValueError("leg bodies must carry exactly one hinge each")
        if cm.jnt_dofadr[ju] != 6 + 2 * k or cm.jnt_dofadr[jl] != 7 + 2 * k:
            raise ValueError("unexpected dof ordering")
        out["p1"][k], out["a1"][k] = loc(fk["xanchor"][ju]), locdir(fk["xaxis"][ju])
        out["p2"][k], out["a2"][k] = loc(fk["xanchor"][jl]), locdir(fk["xaxis"][jl])
        out["mu"][k], out["cu"][k], out["Iu"][k] = body_inertial(bu)
        out["ml"][k], out["cl"][k], out["Il"][k] = body_inertial(bl)
        out["stiff"][k] = [cm.jnt_stiffness[ju], cm.jnt_stiffness[jl]]
        out["damp"][k] = [cm.dof_damping[cm.jnt_dofadr[ju]], cm.dof_damping[cm.jnt_dofadr[jl]]]
        if cm.qpos_spring[cm.jnt_qposadr[ju]] != 0 or cm.qpos_spring[cm.jnt_qposadr[jl]] != 0:
            raise ValueError("non-zero springref is not supported")
        gu = geoms_of(bu, [mjcf.GEOM_CYLINDER, mjcf.GEOM_SPHERE])
        gl = geoms_of(bl, [mjcf.GEOM_CYLINDER, mjcf.GEOM_SPHERE])
        out["g_ucyl"].append(geom(gu[0], mjcf.GEOM_CYLINDER))
        out["g_knee"].append(geom(gu[1], mjcf.GEOM_SPHERE))
        out["g_lcyl"].append(geom(gl[0], mjcf.GEOM_CYLINDER))
        out["g_foot"].append(geom(gl[1], mjcf.GEOM_SPHERE))
        invw[1 + 2 * k] = cm.body_invweight0[bu, 0]
        invw[2 + 2 * k] = cm.body_invweight0[bl, 0]

    jm = int(cm.body_jntadr[rotor])
    if cm.jnt_type[jm] != mjcf.JNT_HINGE or cm.jnt_dofadr[jm] != 14 or cm.jnt_stiffness[jm] != 0:
        raise ValueError("rotor must carry one unsprung hinge on dof 14")
    out["pm"], out["am"] = loc(fk["xanchor"][jm]), locdir(fk["xaxis"][jm])
    out["mm"], out["cm"], out["Im"] = body_inertial(rotor)
    out["damp_m"] = float(cm.dof_damping[14])
    gm = geoms_of(rotor, [mjcf.GEOM_CYLINDER, mjcf.GEOM_ELLIPSOID])
    out["g_mcyl"], out["g_mell"] = geom(gm[0], mjcf.GEOM_CYLINDER), geom(gm[1], mjcf.GEOM_ELLIPSOID)
    invw[9] = cm.body_invweight0[rotor, 0]
    out["invw"] = invw

    if cm.nu != 1 or cm.actuator_trnid[0] != jm:
        raise ValueError("expected one actuator on the rotor joint")
    if cm.actuator_biasprm[0, 0] != 0 or cm.actuator_biasprm[0, 1] != 0:
        raise ValueError("only velocity-affine actuator bias is supported")
    out["gear"] = float(cm.actuator_gear[0])
    out["gain"] = float(cm.actuator_gainprm[0, 0])
    out["bias_v"] = float(cm.actuator_biasprm[0, 2])
    if cm.actuator_ctrllimited[0]:
        out["ctrl_lo"], out["ctrl_hi"] = map(float, cm.actuator_ctrlrange[0])
    else:
        out["ctrl_lo"], out["ctrl_hi"] = -1e30, 1e30

    # contact parameters: every floor-vs-geom pair mixes identical defaults in this family
    coll = [g for g in range(cm.ngeom) if cm.geom_contype[g] and cm.geom_bodyid[g] != 0 and cm.geom_bodyid[g] != cm.nbody - 1]
    g0 = coll[0]
    for g in coll:
        for f in ("geom_friction", "geom_solref", "geom_solimp", "geom_margin", "geom_gap"):
            if not np.array_equal(getattr(cm, f)[g], getattr(cm, f)[g0]):
                raise ValueError("per-geom contact parameters are not supported")
    out["friction"] = float(max(cm.geom_friction[g0, 0], cm.geom_friction[0, 0], 1e-5))
    tc, dr = float(cm.geom_solref[g0, 0]), float(cm.geom_solref[g0, 1])
    dmax = float(np.clip(cm.geom_solimp[g0, 1], 1e-4, 0.9999))
    tc = max(tc, 2 * cm.timestep)          # refsafe
    out["solref_K"] = 1.0 / max(1e-15, dmax * dmax * tc * tc * dr * dr)
    out["solref_B"] = 2.0 / max(1e-15, dmax * tc)
    out["solimp"] = cm.geom_solimp[g0].copy()
    out["impratio"] = float(cm.impratio)
    out["meaninertia"] = float(cm.meaninertia)
    out["root_pos0"] = fk["xpos"][root].copy()
    if not np.allclose(fk["xquat"][root], [1, 0, 0, 0]):
        raise ValueError("root body must be axis-aligned at qpos0")
    tb = cm.body_names.index("target")
    tg = cm.geom_names.index("target")
    out["target_pos0"] = fk["geom_xpos"][tg].copy()      # geom_xpos["target"] at the home target pose
    if not np.allclose(fk["geom_xpos"][tg], fk["xpos"][tb]):
        raise ValueError("target geom must sit at the target body origin")
    return out


def fill_struct(flat: Dict[str, object], double: bool = False):
    """dict from :func:`flatten_model` -> ctypes struct (float32 for the C-ABI, float64 for test builds)."""
    Model, Geom = (JbModelD, JbGeomD) if double else (JbModelF, JbGeomF)
    m = Model()

    def set_geom(dst, g):
        np.ctypeslib.as_array(dst.center)[:] = g["center"]
        np.ctypeslib.as_array(dst.rot)[:] = g["rot"]
        np.ctypeslib.as_array(dst.size)[:] = g["size"]
        dst.rbound = g["rbound"]
        dst.geom_id = g["geom_id"]

    for name, ctype in Model._fields_:
        if name in ("pad_",):
            continue
        val = flat[name]
        field = getattr(m, name)
        if name.startswith("g_"):
            if isinstance(val, list):
                for i, g in enumerate(val):
                    set_geom(field[i], g)
            else:
                set_geom(field, val)
        elif isinstance(val, (float, int)):
            setattr(m, name, val)
        else:
            np.ctypeslib.as_array(field)[...] = np.asarray(val)
    return m
