"""MJCF-subset compiler for the Jitterbug model family.

Compiles ``jitterbug.xml`` (reference: jitterbug_dmc/jitterbug.xml:1-148) into a
MuJoCo-``mjModel``-like set of float64 numpy arrays, in the *local* (body
relative) coordinates MuJoCo uses after compilation.  The reference hands the
XML text to MuJoCo 2.0's compiler through ``Physics.from_xml_string``
(jitterbug_dmc/jitterbug.py:81); no MuJoCo exists in this image and modern
MuJoCo rejects ``<compiler coordinate="global"/>`` (jitterbug.xml:4), so this
module restates the part of the compiler the model needs:

* ``coordinate="global"``: every pos/quat/fromto/axis is a world-frame quantity
  at qpos0; bodies without ``pos`` take their frame from their inertial frame.
* geom volumes / masses / inertias for box, cylinder, ellipsoid, sphere with the
  ``<default><geom density=...>`` class default (jitterbug.xml:22).
* body inertial frames (COM + principal axes) from the geoms (inertiafromgeom).
* kinematic tree bookkeeping (``dof_parentid``, addresses), joint anchors/axes in
  body coordinates, springs/dampers, one ``general`` actuator, contact defaults.
* ``body_invweight0`` / ``dof_invweight0`` / ``meaninertia`` at qpos0 (MuJoCo's
  ``mj_setConst``), which scale the contact regularisation.

It also carries a small numpy forward-kinematics + Jacobian routine; the
compiler needs it for the invweight computation and the tests reuse it as an
independent check of the dynamics code.
"""

from __future__ import annotations

import dataclasses
import os
import xml.etree.ElementTree as ET
from typing import Dict, List, Optional

import numpy as np

# MuJoCo enum values (mjtGeom, mjtJoint)
GEOM_PLANE, GEOM_SPHERE, GEOM_ELLIPSOID, GEOM_CYLINDER, GEOM_BOX = 0, 2, 4, 5, 6
JNT_FREE, JNT_HINGE = 0, 3
_GEOM_TYPES = {
    "plane": GEOM_PLANE, "sphere": GEOM_SPHERE, "ellipsoid": GEOM_ELLIPSOID,
    "cylinder": GEOM_CYLINDER, "box": GEOM_BOX,
}

DEFAULT_XML = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets", "jitterbug.xml")


# --------------------------------------------------------------------------- quaternion helpers (w, x, y, z)
def quat_mul(a, b):
    aw, ax, ay, az = a
    bw, bx, by, bz = b
    return np.array([
        aw * bw - ax * bx - ay * by - az * bz,
        aw * bx + ax * bw + ay * bz - az * by,
        aw * by - ax * bz + ay * bw + az * bx,
        aw * bz + ax * by - ay * bx + az * bw,
    ])


def quat_conj(q):
    return np.array([q[0], -q[1], -q[2], -q[3]])


def quat2mat(q):
    w, x, y, z = q
    return np.array([
        [w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)],
        [2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)],
        [2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z],
    ])


def mat2quat(m):
    """Rotation matrix -> unit quaternion (w >= 0 branch-stable)."""
    t = np.trace(m)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        q = np.array([0.25 * s, (m[2, 1] - m[1, 2]) / s, (m[0, 2] - m[2, 0]) / s, (m[1, 0] - m[0, 1]) / s])
    else:
        i = int(np.argmax(np.diag(m)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(1.0 + m[i, i] - m[j, j] - m[k, k]) * 2
        q = np.zeros(4)
        q[0] = (m[k, j] - m[j, k]) / s
        q[1 + i] = 0.25 * s
        q[1 + j] = (m[j, i] + m[i, j]) / s
        q[1 + k] = (m[k, i] + m[i, k]) / s
    q /= np.linalg.norm(q)
    if q[0] < 0:
        q = -q
    return q


def axis_angle_quat(axis, angle):
    s = np.sin(0.5 * angle)
    return np.array([np.cos(0.5 * angle), axis[0] * s, axis[1] * s, axis[2] * s])


def z2quat(vec):
    """Quaternion rotating the z axis onto ``vec`` (MuJoCo's fromto convention)."""
    vec = np.asarray(vec, dtype=np.float64)
    vec = vec / np.linalg.norm(vec)
    axis = np.cross([0.0, 0.0, 1.0], vec)
    s = np.linalg.norm(axis)
    if s < 1e-10:
        axis = np.array([1.0, 0.0, 0.0])
    else:
        axis = axis / s
    ang = np.arctan2(s, vec[2])
    return axis_angle_quat(axis, ang)


# --------------------------------------------------------------------------- compiled model container
@dataclasses.dataclass
class CompiledModel:
    """mjModel-like arrays (float64 / int32), local coordinates."""
    # sizes
    nq: int
    nv: int
    nu: int
    nbody: int
    njnt: int
    ngeom: int
    # options
    timestep: float
    gravity: np.ndarray
    tolerance: float
    iterations: int
    impratio: float
    # bodies
    body_parentid: np.ndarray
    body_rootid: np.ndarray
    body_jntnum: np.ndarray
    body_jntadr: np.ndarray
    body_dofnum: np.ndarray
    body_dofadr: np.ndarray
    body_pos: np.ndarray
    body_quat: np.ndarray
    body_ipos: np.ndarray
    body_iquat: np.ndarray
    body_mass: np.ndarray
    body_inertia: np.ndarray
    body_invweight0: np.ndarray
    # joints
    jnt_type: np.ndarray
    jnt_qposadr: np.ndarray
    jnt_dofadr: np.ndarray
    jnt_bodyid: np.ndarray
    jnt_pos: np.ndarray
    jnt_axis: np.ndarray
    jnt_stiffness: np.ndarray
    qpos0: np.ndarray
    qpos_spring: np.ndarray
    # dofs
    dof_bodyid: np.ndarray
    dof_jntid: np.ndarray
    dof_parentid: np.ndarray
    dof_damping: np.ndarray
    dof_invweight0: np.ndarray
    # geoms
    geom_type: np.ndarray
    geom_bodyid: np.ndarray
    geom_contype: np.ndarray
    geom_density: np.ndarray          # compile-time only (MuJoCo drops it); kept for the on-device augmentation
    geom_density_default: np.ndarray  # 1 where the geom took <default><geom density>
    geom_conaffinity: np.ndarray
    geom_condim: np.ndarray
    geom_size: np.ndarray
    geom_pos: np.ndarray
    geom_quat: np.ndarray
    geom_friction: np.ndarray
    geom_solref: np.ndarray
    geom_solimp: np.ndarray
    geom_margin: np.ndarray
    geom_gap: np.ndarray
    geom_solmix: np.ndarray
    geom_mass: np.ndarray
    # actuator (general, joint transmission)
    actuator_trnid: np.ndarray
    actuator_gear: np.ndarray
    actuator_gainprm: np.ndarray
    actuator_biasprm: np.ndarray
    actuator_ctrlrange: np.ndarray
    actuator_ctrllimited: np.ndarray
    # sensor: framelinvel of this body (objtype=body => inertial frame)
    sensor_body: int
    # statistics
    meaninertia: float
    # names
    body_names: List[str]
    jnt_names: List[str]
    geom_names: List[str]

    def body_id(self, name: str) -> int:
        return self.body_names.index(name)

    def geom_id(self, name: str) -> int:
        return self.geom_names.index(name)

    def jnt_id(self, name: str) -> int:
        return self.jnt_names.index(name)


# --------------------------------------------------------------------------- geometry: mass / inertia of primitive geoms
def geom_mass_inertia(gtype: int, size: np.ndarray, density: float):
    """Mass and diagonal inertia (geom frame) as MuJoCo's compiler computes them."""
    if gtype == GEOM_SPHERE:
        r = size[0]
        m = density * 4.0 / 3.0 * np.pi * r ** 3
        i = np.full(3, 0.4 * m * r * r)
    elif gtype == GEOM_CYLINDER:
        r, hl = size[0], size[1]
        height = 2.0 * hl
        m = density * np.pi * r * r * height
        ixy = m * (3.0 * r * r + height * height) / 12.0
        i = np.array([ixy, ixy, 0.5 * m * r * r])
    elif gtype == GEOM_ELLIPSOID:
        a, b, c = size
        m = density * 4.0 / 3.0 * np.pi * a * b * c
        i = m / 5.0 * np.array([b * b + c * c, a * a + c * c, a * a + b * b])
    elif gtype == GEOM_BOX:
        a, b, c = size
        m = density * 8.0 * a * b * c
        i = m / 3.0 * np.array([b * b + c * c, a * a + c * c, a * a + b * b])
    elif gtype == GEOM_PLANE:
        m, i = 0.0, np.zeros(3)
    else:
        raise ValueError(f"unsupported geom type {gtype}")
    return m, i


def _floats(s: str) -> np.ndarray:
    return np.array([float(x) for x in s.split()], dtype=np.float64)


class _Body:
    def __init__(self, name, parent):
        self.name = name
        self.parent = parent          # index
        self.pos = None               # global, may be None
        self.quat = None
        self.geoms = []
        self.joints = []


def compile_xml(xml: Optional[str] = None, *, path: Optional[str] = None) -> CompiledModel:
    """Compile the MJCF subset used by jitterbug.xml.

    Args:
        xml: XML text (as ``get_model_and_assets()[0]`` returns it,
            jitterbug_dmc/jitterbug.py:61-69); ``<include>`` elements only carry
            visual assets and are ignored.
        path: alternatively a file path.  Default: the packaged jitterbug.xml.
    """
    if xml is None:
        with open(path or DEFAULT_XML, "rb") as f:
            xml = f.read()
    root = ET.fromstring(xml)

    comp = root.find("compiler")
    coordinate = comp.attrib.get("coordinate", "local") if comp is not None else "local"
    if coordinate not in ("global", "local"):
        raise NotImplementedError("unknown <compiler coordinate>")
    local = coordinate == "local"      # e.g. to_local_xml()'s output: frames are composed into world frames while walking

    opt = root.find("option")
    timestep = float(opt.attrib.get("timestep", 0.002)) if opt is not None else 0.002

    default_density = 1000.0
    for d in root.findall("default"):
        for g in d.findall("geom"):
            if "density" in g.attrib:
                default_density = float(g.attrib["density"])

    # ---- walk the body tree, collecting global-frame specs
    bodies: List[_Body] = [_Body("world", -1)]
    bodies[0].pos = np.zeros(3)
    bodies[0].quat = np.array([1.0, 0, 0, 0])
    geoms = []   # dicts
    joints = []

    def to_world(bidx, pos, quat=None, vec=None):
        """local-coordinate models: compose an element's body-relative pose / direction with its body's world frame"""
        if not local:
            return pos, quat, vec
        b = bodies[bidx]
        rb = quat2mat(b.quat)
        return (b.pos + rb @ pos if pos is not None else None, quat_mul(b.quat, quat) if quat is not None else None,
                rb @ vec if vec is not None else None)

    def add_geom(el, bidx):
        gtype = _GEOM_TYPES[el.attrib.get("type", "sphere")]
        size = np.zeros(3)
        s = _floats(el.attrib["size"])
        size[:len(s)] = s
        if "fromto" in el.attrib:
            ft = _floats(el.attrib["fromto"])
            p_from, _, _ = to_world(bidx, ft[0:3])
            p_to, _, _ = to_world(bidx, ft[3:6])
            vec = p_from - p_to              # MuJoCo: z axis points from "to" to "from"
            size[1] = 0.5 * np.linalg.norm(vec)
            pos = 0.5 * (p_from + p_to)
            quat = z2quat(vec)
        else:
            pos = _floats(el.attrib.get("pos", "0 0 0"))
            quat = _floats(el.attrib.get("quat", "1 0 0 0"))
            quat = quat / np.linalg.norm(quat)
            pos, quat, _ = to_world(bidx, pos, quat)
        geoms.append(dict(
            name=el.attrib.get("name", ""), type=gtype, body=bidx, size=size, pos=pos, quat=quat,
            density=float(el.attrib.get("density", default_density)), density_default=int("density" not in el.attrib),
            contype=int(el.attrib.get("contype", 1)), conaffinity=int(el.attrib.get("conaffinity", 1)),
        ))
        bodies[bidx].geoms.append(len(geoms) - 1)

    def add_joint(el, bidx):
        jtype = {"free": JNT_FREE, "hinge": JNT_HINGE}[el.attrib.get("type", "hinge")]
        pos, _, axis = to_world(bidx, _floats(el.attrib.get("pos", "0 0 0")), None, _floats(el.attrib.get("axis", "0 0 1")))
        joints.append(dict(
            name=el.attrib.get("name", ""), type=jtype, body=bidx,
            pos=pos if jtype != JNT_FREE else None,
            axis=axis,
            stiffness=float(el.attrib.get("stiffness", 0.0)),
            damping=float(el.attrib.get("damping", 0.0)),
        ))
        bodies[bidx].joints.append(len(joints) - 1)

    def walk(el, bidx):
        # MuJoCo orders geoms/joints per body in document order, bodies depth-first
        for child in el:
            if child.tag == "geom":
                add_geom(child, bidx)
            elif child.tag == "joint":
                add_joint(child, bidx)
        for child in el:
            if child.tag == "body":
                b = _Body(child.attrib.get("name", ""), bidx)
                if "pos" in child.attrib or local:
                    b.pos = _floats(child.attrib.get("pos", "0 0 0"))
                    b.quat = _floats(child.attrib.get("quat", "1 0 0 0"))
                    b.quat = b.quat / np.linalg.norm(b.quat)
                    if local:
                        par = bodies[bidx]
                        b.pos = par.pos + quat2mat(par.quat) @ b.pos
                        b.quat = quat_mul(par.quat, b.quat)
                bodies.append(b)
                walk(child, len(bodies) - 1)

    walk(root.find("worldbody"), 0)

    # MuJoCo numbers geoms/joints in body order (depth-first body ids); re-sort accordingly
    nbody = len(bodies)
    geom_order = [g for b in range(nbody) for g in bodies[b].geoms]
    jnt_order = [j for b in range(nbody) for j in bodies[b].joints]
    geoms = [geoms[g] for g in geom_order]
    joints = [joints[j] for j in jnt_order]
    ngeom, njnt = len(geoms), len(joints)
    for b in bodies:
        b.geoms, b.joints = [], []
    for gi, g in enumerate(geoms):
        bodies[g["body"]].geoms.append(gi)
    for ji, j in enumerate(joints):
        bodies[j["body"]].joints.append(ji)

    # ---- geom masses, body inertial frames (global), body frames
    geom_mass = np.zeros(ngeom)
    geom_inertia = np.zeros((ngeom, 3))
    for gi, g in enumerate(geoms):
        geom_mass[gi], geom_inertia[gi] = geom_mass_inertia(g["type"], g["size"], g["density"])

    xpos = np.zeros((nbody, 3))
    xquat = np.zeros((nbody, 4))
    xquat[:, 0] = 1.0
    body_mass = np.zeros(nbody)
    body_inertia = np.zeros((nbody, 3))
    xipos = np.zeros((nbody, 3))
    xiquat = np.zeros((nbody, 4))
    xiquat[:, 0] = 1.0
    for bi in range(1, nbody):
        b = bodies[bi]
        m = sum(geom_mass[g] for g in b.geoms)
        if m > 0:
            com = sum(geom_mass[g] * geoms[g]["pos"] for g in b.geoms) / m
            inert = np.zeros((3, 3))
            for g in b.geoms:
                rg = quat2mat(geoms[g]["quat"])
                d = geoms[g]["pos"] - com
                inert += rg @ np.diag(geom_inertia[g]) @ rg.T
                inert += geom_mass[g] * (np.dot(d, d) * np.eye(3) - np.outer(d, d))
            evals, evecs = np.linalg.eigh(inert)
            # descending order like mju_eig3, right-handed
            order = np.argsort(-evals)
            evals, evecs = evals[order], evecs[:, order]
            if np.linalg.det(evecs) < 0:
                evecs[:, 2] = -evecs[:, 2]
            iq = mat2quat(evecs)
        else:
            com = b.pos.copy() if b.pos is not None else xpos[b.parent].copy()
            evals, iq = np.zeros(3), np.array([1.0, 0, 0, 0])
        body_mass[bi], body_inertia[bi] = m, evals
        xipos[bi], xiquat[bi] = com, iq
        if b.pos is None:
            # global coordinates + undefined body frame => body frame := inertial frame
            xpos[bi], xquat[bi] = com, iq
        else:
            xpos[bi], xquat[bi] = b.pos, b.quat

    # ---- localise
    body_parentid = np.array([max(b.parent, 0) for b in bodies], dtype=np.int32)
    body_pos = np.zeros((nbody, 3))
    body_quat = np.zeros((nbody, 4)); body_quat[:, 0] = 1
    body_ipos = np.zeros((nbody, 3))
    body_iquat = np.zeros((nbody, 4)); body_iquat[:, 0] = 1
    for bi in range(1, nbody):
        p = body_parentid[bi]
        rp = quat2mat(xquat[p])
        body_pos[bi] = rp.T @ (xpos[bi] - xpos[p])
        body_quat[bi] = quat_mul(quat_conj(xquat[p]), xquat[bi])
        rb = quat2mat(xquat[bi])
        body_ipos[bi] = rb.T @ (xipos[bi] - xpos[bi])
        body_iquat[bi] = quat_mul(quat_conj(xquat[bi]), xiquat[bi])

    geom_pos = np.zeros((ngeom, 3))
    geom_quat = np.zeros((ngeom, 4))
    for gi, g in enumerate(geoms):
        bi = g["body"]
        rb = quat2mat(xquat[bi])
        geom_pos[gi] = rb.T @ (g["pos"] - xpos[bi])
        geom_quat[gi] = quat_mul(quat_conj(xquat[bi]), g["quat"])

    # ---- joints / dofs
    jnt_type = np.array([j["type"] for j in joints], dtype=np.int32)
    jnt_bodyid = np.array([j["body"] for j in joints], dtype=np.int32)
    jnt_pos = np.zeros((njnt, 3))
    jnt_axis = np.zeros((njnt, 3))
    jnt_qposadr = np.zeros(njnt, dtype=np.int32)
    jnt_dofadr = np.zeros(njnt, dtype=np.int32)
    jnt_stiffness = np.array([j["stiffness"] for j in joints])
    nq = nv = 0
    dof_bodyid, dof_jntid, dof_damping = [], [], []
    qpos0 = []
    for ji, j in enumerate(joints):
        bi = j["body"]
        rb = quat2mat(xquat[bi])
        jnt_qposadr[ji], jnt_dofadr[ji] = nq, nv
        if j["type"] == JNT_FREE:
            jnt_axis[ji] = [0, 0, 1]
            qpos0 += list(xpos[bi]) + list(xquat[bi])
            nq += 7
            nd = 6
        else:
            jnt_pos[ji] = rb.T @ (j["pos"] - xpos[bi])
            ax = rb.T @ j["axis"]
            jnt_axis[ji] = ax / np.linalg.norm(ax)
            qpos0.append(0.0)
            nq += 1
            nd = 1
        for _ in range(nd):
            dof_bodyid.append(bi)
            dof_jntid.append(ji)
            dof_damping.append(j["damping"])
        nv += nd
    qpos0 = np.array(qpos0)
    dof_bodyid = np.array(dof_bodyid, dtype=np.int32)
    dof_jntid = np.array(dof_jntid, dtype=np.int32)
    dof_damping = np.array(dof_damping)

    body_jntnum = np.array([len(b.joints) for b in bodies], dtype=np.int32)
    body_jntadr = np.array([b.joints[0] if b.joints else -1 for b in bodies], dtype=np.int32)
    body_dofnum = np.zeros(nbody, dtype=np.int32)
    body_dofadr = np.full(nbody, -1, dtype=np.int32)
    for d, bi in enumerate(dof_bodyid):
        if body_dofadr[bi] < 0:
            body_dofadr[bi] = d
        body_dofnum[bi] += 1
    # dof_parentid: previous dof in the same body, else last dof of the nearest ancestor with dofs
    dof_parentid = np.full(nv, -1, dtype=np.int32)
    for d in range(nv):
        bi = dof_bodyid[d]
        if d > body_dofadr[bi]:
            dof_parentid[d] = d - 1
        else:
            p = body_parentid[bi]
            while p > 0 and body_dofnum[p] == 0:
                p = body_parentid[p]
            if p > 0:
                dof_parentid[d] = body_dofadr[p] + body_dofnum[p] - 1
    body_rootid = np.zeros(nbody, dtype=np.int32)
    for bi in range(1, nbody):
        p = bi
        while body_parentid[p] != 0:
            p = body_parentid[p]
        body_rootid[bi] = p

    # ---- actuator
    act = root.find("actuator")
    gens = list(act) if act is not None else []
    nu = len(gens)
    jnames = [j["name"] for j in joints]
    actuator_trnid = np.zeros(nu, dtype=np.int32)
    actuator_gear = np.zeros(nu)
    actuator_gainprm = np.zeros((nu, 3))
    actuator_biasprm = np.zeros((nu, 3))
    actuator_ctrlrange = np.zeros((nu, 2))
    actuator_ctrllimited = np.zeros(nu, dtype=np.int32)
    for ai, g in enumerate(gens):
        if g.tag != "general" or g.attrib.get("dyntype", "none") != "none" \
                or g.attrib.get("gaintype", "fixed") != "fixed" or g.attrib.get("biastype", "none") not in ("none", "affine"):
            raise NotImplementedError("only stateless general actuators with fixed gain / affine bias are supported")
        actuator_trnid[ai] = jnames.index(g.attrib["joint"])
        actuator_gear[ai] = _floats(g.attrib.get("gear", "1"))[0]
        gp = _floats(g.attrib.get("gainprm", "1 0 0")); actuator_gainprm[ai, :len(gp)] = gp[:3]
        bp = _floats(g.attrib.get("biasprm", "0 0 0")); actuator_biasprm[ai, :len(bp)] = bp[:3]
        if g.attrib.get("biastype", "none") == "none":
            actuator_biasprm[ai] = 0
        actuator_ctrllimited[ai] = 1 if g.attrib.get("ctrllimited", "false") == "true" else 0
        actuator_ctrlrange[ai] = _floats(g.attrib.get("ctrlrange", "0 0"))

    # ---- sensor (framelinvel objtype=body)
    sensor_body = -1
    sens = root.find("sensor")
    bnames = [b.name for b in bodies]
    if sens is not None:
        for s in sens:
            if s.tag == "framelinvel" and s.attrib.get("objtype") == "body":
                sensor_body = bnames.index(s.attrib["objname"])

    model = CompiledModel(
        nq=nq, nv=nv, nu=nu, nbody=nbody, njnt=njnt, ngeom=ngeom,
        timestep=timestep, gravity=np.array([0.0, 0.0, -9.81]), tolerance=1e-8, iterations=100, impratio=1.0,
        body_parentid=body_parentid, body_rootid=body_rootid, body_jntnum=body_jntnum, body_jntadr=body_jntadr,
        body_dofnum=body_dofnum, body_dofadr=body_dofadr,
        body_pos=body_pos, body_quat=body_quat, body_ipos=body_ipos, body_iquat=body_iquat,
        body_mass=body_mass, body_inertia=body_inertia, body_invweight0=np.zeros((nbody, 2)),
        jnt_type=jnt_type, jnt_qposadr=jnt_qposadr, jnt_dofadr=jnt_dofadr, jnt_bodyid=jnt_bodyid,
        jnt_pos=jnt_pos, jnt_axis=jnt_axis, jnt_stiffness=jnt_stiffness, qpos0=qpos0, qpos_spring=qpos0.copy(),
        dof_bodyid=dof_bodyid, dof_jntid=dof_jntid, dof_parentid=dof_parentid, dof_damping=dof_damping,
        dof_invweight0=np.zeros(nv),
        geom_type=np.array([g["type"] for g in geoms], dtype=np.int32),
        geom_bodyid=np.array([g["body"] for g in geoms], dtype=np.int32),
        geom_contype=np.array([g["contype"] for g in geoms], dtype=np.int32),
        geom_density=np.array([g["density"] for g in geoms], dtype=np.float64),
        geom_density_default=np.array([g["density_default"] for g in geoms], dtype=np.int32),
        geom_conaffinity=np.array([g["conaffinity"] for g in geoms], dtype=np.int32),
        geom_condim=np.full(ngeom, 3, dtype=np.int32),
        geom_size=np.array([g["size"] for g in geoms]),
        geom_pos=geom_pos, geom_quat=geom_quat,
        geom_friction=np.tile(np.array([1.0, 0.005, 0.0001]), (ngeom, 1)),
        geom_solref=np.tile(np.array([0.02, 1.0]), (ngeom, 1)),
        geom_solimp=np.tile(np.array([0.9, 0.95, 0.001, 0.5, 2.0]), (ngeom, 1)),
        geom_margin=np.zeros(ngeom), geom_gap=np.zeros(ngeom), geom_solmix=np.ones(ngeom),
        geom_mass=geom_mass,
        actuator_trnid=actuator_trnid, actuator_gear=actuator_gear, actuator_gainprm=actuator_gainprm,
        actuator_biasprm=actuator_biasprm, actuator_ctrlrange=actuator_ctrlrange,
        actuator_ctrllimited=actuator_ctrllimited,
        sensor_body=sensor_body, meaninertia=1.0,
        body_names=bnames, jnt_names=jnames, geom_names=[g["name"] for g in geoms],
    )
    _set_const(model)
    return model


# --------------------------------------------------------------------------- export
def to_local_xml(m: CompiledModel) -> str:
    """The compiled model as a LOCAL-coordinate MJCF that current MuJoCo releases load.

    ``jitterbug.xml`` uses ``<compiler coordinate="global"/>`` (jitterbug_dmc/jitterbug.xml:4), which MuJoCo removed in
    2.3.4, so ``pip install mujoco`` cannot load the reference model.  This writes the same model with every frame
    expressed in its parent (what MuJoCo 2.0's compiler produced internally): body frames as compiled here (bodies
    without ``pos`` in the global file sit at their inertial frame), geoms by type / size / pos / quat instead of
    ``fromto``, unit hinge axes, MuJoCo 2.0's option defaults spelled out.  Masses and inertias are NOT written:
    MuJoCo recomputes them from the geoms, which is one of the things ``tests/test_mujoco_parity.py`` compares.
    The ``coordinate`` attribute itself is omitted (local is the only mode left).
    """
    def f(a):
        return " ".join(repr(float(x)) for x in np.atleast_1d(a))

    gtypes = {v: k for k, v in _GEOM_TYPES.items()}
    children = {b: [c for c in range(1, m.nbody) if m.body_parentid[c] == b] for b in range(m.nbody)}
    out = ['<mujoco model="jitterbug_local">',
           '  <compiler angle="radian" inertiafromgeom="true"/>',
           f'  <option timestep="{f(m.timestep)}" gravity="{f(m.gravity)}" integrator="Euler" cone="pyramidal" '
           f'solver="Newton" iterations="{int(m.iterations)}" tolerance="{f(m.tolerance)}" impratio="{f(m.impratio)}"/>',
           '  <worldbody>']

    def geom_xml(g, ind):
        t = int(m.geom_type[g])
        nsize = {GEOM_PLANE: 3, GEOM_SPHERE: 1, GEOM_CYLINDER: 2, GEOM_ELLIPSOID: 3, GEOM_BOX: 3}[t]
        name = f' name="{m.geom_names[g]}"' if m.geom_names[g] else ""
        dens = "" if t == GEOM_PLANE else f' density="{f(m.geom_density[g])}"'
        return (f'{ind}<geom{name} type="{gtypes[t]}" size="{f(m.geom_size[g][:nsize])}" pos="{f(m.geom_pos[g])}" '
                f'quat="{f(m.geom_quat[g])}"{dens} contype="{int(m.geom_contype[g])}" '
                f'conaffinity="{int(m.geom_conaffinity[g])}" condim="{int(m.geom_condim[g])}" '
                f'friction="{f(m.geom_friction[g])}" solref="{f(m.geom_solref[g])}" solimp="{f(m.geom_solimp[g])}" '
                f'margin="{f(m.geom_margin[g])}" gap="{f(m.geom_gap[g])}"/>')

    def body_xml(b, ind):
        for g in range(m.ngeom):
            if m.geom_bodyid[g] == b:
                out.append(geom_xml(g, ind))
        for j in range(m.njnt):
            if m.jnt_bodyid[j] == b:
                if m.jnt_type[j] == JNT_FREE:
                    out.append(f'{ind}<joint name="{m.jnt_names[j]}" type="free"/>')
                else:
                    d = int(m.jnt_dofadr[j])
                    out.append(f'{ind}<joint name="{m.jnt_names[j]}" type="hinge" pos="{f(m.jnt_pos[j])}" '
                               f'axis="{f(m.jnt_axis[j])}" stiffness="{f(m.jnt_stiffness[j])}" '
                               f'damping="{f(m.dof_damping[d])}" springref="0"/>')
        for c in children[b]:
            out.append(f'{ind}<body name="{m.body_names[c]}" pos="{f(m.body_pos[c])}" quat="{f(m.body_quat[c])}">')
            body_xml(c, ind + "  ")
            out.append(f'{ind}</body>')

    body_xml(0, "    ")
    out.append('  </worldbody>')
    if m.sensor_body >= 0:
        out += ['  <sensor>', f'    <framelinvel name="jitterbug_framelinvel" objtype="body" '
                              f'objname="{m.body_names[m.sensor_body]}"/>', '  </sensor>']
    out.append('  <actuator>')
    for a in range(m.nu):
        out.append(f'    <general joint="{m.jnt_names[int(m.actuator_trnid[a])]}" '
                   f'ctrllimited="{"true" if m.actuator_ctrllimited[a] else "false"}" ctrlrange="{f(m.actuator_ctrlrange[a])}" '
                   f'gear="{f(m.actuator_gear[a])}" dyntype="none" gaintype="fixed" gainprm="{f(m.actuator_gainprm[a])}" '
                   f'biastype="affine" biasprm="{f(m.actuator_biasprm[a])}"/>')
    out += ['  </actuator>', '</mujoco>', '']
    return "\n".join(out)


# --------------------------------------------------------------------------- numpy kinematics / Jacobians (compiler + tests)
def forward_kinematics(m: CompiledModel, qpos: np.ndarray) -> Dict[str, np.ndarray]:
    """World poses of bodies, inertial frames, joints and geoms (mj_kinematics)."""
    qpos = np.asarray(qpos, dtype=np.float64)
    xpos = np.zeros((m.nbody, 3)); xquat = np.zeros((m.nbody, 4)); xquat[0, 0] = 1
    xmat = np.zeros((m.nbody, 3, 3)); xmat[0] = np.eye(3)
    xanchor = np.zeros((m.njnt, 3)); xaxis = np.zeros((m.njnt, 3))
    for b in range(1, m.nbody):
        p = m.body_parentid[b]
        jn, ja = m.body_jntnum[b], m.body_jntadr[b]
        if jn == 1 and m.jnt_type[ja] == JNT_FREE:
            a = m.jnt_qposadr[ja]
            xp = qpos[a:a + 3].copy()
            xq = qpos[a + 3:a + 7] / np.linalg.norm(qpos[a + 3:a + 7])
            xanchor[ja], xaxis[ja] = xp, [0, 0, 1]
        else:
            xp = xpos[p] + xmat[p] @ m.body_pos[b]
            xq = quat_mul(xquat[p], m.body_quat[b])
            for j in range(ja, ja + jn):
                r = quat2mat(xq)
                xaxis[j] = r @ m.jnt_axis[j]
                xanchor[j] = r @ m.jnt_pos[j] + xp
                ang = qpos[m.jnt_qposadr[j]] - m.qpos0[m.jnt_qposadr[j]]
                xq = quat_mul(xq, axis_angle_quat(m.jnt_axis[j], ang))
                xp = xanchor[j] - quat2mat(xq) @ m.jnt_pos[j]
        xq = xq / np.linalg.norm(xq)
        xpos[b], xquat[b], xmat[b] = xp, xq, quat2mat(xq)
    xipos = np.array([xpos[b] + xmat[b] @ m.body_ipos[b] for b in range(m.nbody)])
    ximat = np.array([quat2mat(quat_mul(xquat[b], m.body_iquat[b])) for b in range(m.nbody)])
    gpos = np.array([xpos[m.geom_bodyid[g]] + xmat[m.geom_bodyid[g]] @ m.geom_pos[g] for g in range(m.ngeom)])
    gmat = np.array([quat2mat(quat_mul(xquat[m.geom_bodyid[g]], m.geom_quat[g])) for g in range(m.ngeom)])
    return dict(xpos=xpos, xquat=xquat, xmat=xmat, xipos=xipos, ximat=ximat,
                xanchor=xanchor, xaxis=xaxis, geom_xpos=gpos, geom_xmat=gmat)


def jacobian(m: CompiledModel, fk: Dict[str, np.ndarray], point: np.ndarray, body: int):
    """Translational (3,nv) and rotational (3,nv) Jacobians of a world point attached to ``body`` (mj_jac)."""
    jacp = np.zeros((3, m.nv)); jacr = np.zeros((3, m.nv))
    b = body
    while b > 0:
        for j in range(m.body_jntadr[b], m.body_jntadr[b] + m.body_jntnum[b]):
            d = m.jnt_dofadr[j]
            if m.jnt_type[j] == JNT_FREE:
                jacp[:, d:d + 3] = np.eye(3)
                for k in range(3):
                    ax = fk["xmat"][b][:, k]
                    jacr[:, d + 3 + k] = ax
                    jacp[:, d + 3 + k] = np.cross(ax, point - fk["xpos"][b])
            else:
                ax = fk["xaxis"][j]
                jacr[:, d] = ax
                jacp[:, d] = np.cross(ax, point - fk["xanchor"][j])
        b = m.body_parentid[b]
    return jacp, jacr


def mass_matrix(m: CompiledModel, fk: Dict[str, np.ndarray]) -> np.ndarray:
    """Joint-space inertia from body Jacobians: M = sum_b Jp'mJp + Jr' I_b Jr (independent of CRB)."""
    M = np.zeros((m.nv, m.nv))
    for b in range(1, m.nbody):
        if m.body_mass[b] == 0 or _is_static(m, b):
            continue
        jp, jr = jacobian(m, fk, fk["xipos"][b], b)
        iw = fk["ximat"][b] @ np.diag(m.body_inertia[b]) @ fk["ximat"][b].T
        M += m.body_mass[b] * jp.T @ jp + jr.T @ iw @ jr
    return M


def _is_static(m: CompiledModel, b: int) -> bool:
    while b > 0:
        if m.body_dofnum[b] > 0:
            return False
        b = m.body_parentid[b]
    return True


def _set_const(m: CompiledModel) -> None:
    """mj_setConst: body_invweight0, dof_invweight0, meaninertia at qpos0."""
    fk = forward_kinematics(m, m.qpos0)
    M = mass_matrix(m, fk)
    Minv = np.linalg.inv(M)
    m.meaninertia = float(np.mean(np.diag(M)))
    for b in range(1, m.nbody):
        if _is_static(m, b):
            m.body_invweight0[b] = 0
            continue
        jp, jr = jacobian(m, fk, fk["xipos"][b], b)
        J = np.vstack([jp, jr])
        A = J @ Minv @ J.T
        m.body_invweight0[b, 0] = (A[0, 0] + A[1, 1] + A[2, 2]) / 3
        m.body_invweight0[b, 1] = (A[3, 3] + A[4, 4] + A[5, 5]) / 3
    dinv = np.diag(Minv)
    for j in range(m.njnt):
        d = m.jnt_dofadr[j]
        if m.jnt_type[j] == JNT_FREE:
            m.dof_invweight0[d:d + 3] = dinv[d:d + 3].mean()
            m.dof_invweight0[d + 3:d + 6] = dinv[d + 3:d + 6].mean()
        else:
            m.dof_invweight0[d] = dinv[d]
