"""TEST INFRASTRUCTURE ONLY -- ctypes binding of the fp64 CPU oracle (oracle/jb_oracle.c).

Pinning status (see jb_oracle.h): the task layer and the heuristic policies are pinned to outputs of the reference's
own code (tests/test_reference_golden.py); the step pipeline is PARITY UNPINNED -- no MuJoCo / dm_control in this
image and no golden vectors in the reference (its published training curves anchor return distributions under a random
policy, tests/test_published_curves.py, nothing finer).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module; the product package ``jitterbug_b200`` never does.

Besides the C step pipeline this module restates, in numpy, the reference's task layer so the CUDA task
kernels can be checked against an implementation that reads like the reference:
  * observations  -- jitterbug_dmc/jitterbug.py:673-763 (+ Physics helpers :180-317)
  * rewards       -- jitterbug_dmc/jitterbug.py:840-925 with dm_control.utils.rewards.tolerance
  * heuristics    -- jitterbug_dmc/heuristic_policies.py:6-136
"""

from __future__ import annotations

import collections
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libjb_oracle.so")

MAXB, MAXJ, MAXV, MAXQ, MAXG, MAXU, MAXCON = 16, 16, 16, 20, 32, 2, 64
MAXEFC = 4 * MAXCON
TASKS = ("move_from_origin", "face_direction", "move_in_direction", "move_to_position", "move_to_pose")


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (recipe: oracle/Makefile)."""
    src = os.path.join(_HERE, "jb_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < max(
            os.path.getmtime(src), os.path.getmtime(os.path.join(_HERE, "jb_oracle.h"))):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _LIB_PATH


def _arr(t, *dims):
    for d in reversed(dims):
        t = t * d
    return t


D, I = C.c_double, C.c_int


class JboModel(C.Structure):
    _fields_ = [
        ("nq", I), ("nv", I), ("nu", I), ("nbody", I), ("njnt", I), ("ngeom", I),
        ("iterations", I), ("sensor_body", I),
        ("timestep", D), ("gravity", _arr(D, 3)), ("tolerance", D), ("impratio", D), ("meaninertia", D),
        ("body_parentid", _arr(I, MAXB)), ("body_rootid", _arr(I, MAXB)), ("body_jntnum", _arr(I, MAXB)),
        ("body_jntadr", _arr(I, MAXB)), ("body_dofnum", _arr(I, MAXB)), ("body_dofadr", _arr(I, MAXB)),
        ("body_pos", _arr(D, MAXB, 3)), ("body_quat", _arr(D, MAXB, 4)), ("body_ipos", _arr(D, MAXB, 3)),
        ("body_iquat", _arr(D, MAXB, 4)), ("body_mass", _arr(D, MAXB)), ("body_inertia", _arr(D, MAXB, 3)),
        ("body_invweight0", _arr(D, MAXB, 2)),
        ("jnt_type", _arr(I, MAXJ)), ("jnt_qposadr", _arr(I, MAXJ)), ("jnt_dofadr", _arr(I, MAXJ)),
        ("jnt_bodyid", _arr(I, MAXJ)),
        ("jnt_pos", _arr(D, MAXJ, 3)), ("jnt_axis", _arr(D, MAXJ, 3)), ("jnt_stiffness", _arr(D, MAXJ)),
        ("qpos0", _arr(D, MAXQ)), ("qpos_spring", _arr(D, MAXQ)),
        ("dof_bodyid", _arr(I, MAXV)), ("dof_jntid", _arr(I, MAXV)), ("dof_parentid", _arr(I, MAXV)),
        ("dof_damping", _arr(D, MAXV)),
        ("geom_type", _arr(I, MAXG)), ("geom_bodyid", _arr(I, MAXG)), ("geom_contype", _arr(I, MAXG)),
        ("geom_conaffinity", _arr(I, MAXG)),
        ("geom_size", _arr(D, MAXG, 3)), ("geom_pos", _arr(D, MAXG, 3)), ("geom_quat", _arr(D, MAXG, 4)),
        ("geom_friction", _arr(D, MAXG, 3)), ("geom_solref", _arr(D, MAXG, 2)), ("geom_solimp", _arr(D, MAXG, 5)),
        ("geom_margin", _arr(D, MAXG)), ("geom_gap", _arr(D, MAXG)), ("geom_solmix", _arr(D, MAXG)),
        ("actuator_trnid", _arr(I, MAXU)), ("actuator_ctrllimited", _arr(I, MAXU)),
        ("actuator_gear", _arr(D, MAXU)), ("actuator_gainprm", _arr(D, MAXU, 3)),
        ("actuator_biasprm", _arr(D, MAXU, 3)), ("actuator_ctrlrange", _arr(D, MAXU, 2)),
        ("target_body", I), ("target_geom", I),
    ]


class JboData(C.Structure):
    _fields_ = [
        ("qpos", _arr(D, MAXQ)), ("qvel", _arr(D, MAXV)), ("ctrl", _arr(D, MAXU)),
        ("qacc_warmstart", _arr(D, MAXV)), ("time", D),
        ("xpos", _arr(D, MAXB, 3)), ("xquat", _arr(D, MAXB, 4)), ("xmat", _arr(D, MAXB, 9)),
        ("xipos", _arr(D, MAXB, 3)), ("ximat", _arr(D, MAXB, 9)),
        ("xanchor", _arr(D, MAXJ, 3)), ("xaxis", _arr(D, MAXJ, 3)),
        ("geom_xpos", _arr(D, MAXG, 3)), ("geom_xmat", _arr(D, MAXG, 9)),
        ("subtree_com", _arr(D, MAXB, 3)), ("cinert", _arr(D, MAXB, 10)), ("crb", _arr(D, MAXB, 10)),
        ("cdof", _arr(D, MAXV, 6)),
        ("qM", _arr(D, MAXV, MAXV)),
        ("ncon", I), ("con_geom1", _arr(I, MAXCON)), ("con_geom2", _arr(I, MAXCON)),
        ("con_exclude", _arr(I, MAXCON)),
        ("con_dist", _arr(D, MAXCON)), ("con_pos", _arr(D, MAXCON, 3)), ("con_frame", _arr(D, MAXCON, 9)),
        ("con_friction", _arr(D, MAXCON, 5)), ("con_solref", _arr(D, MAXCON, 2)),
        ("con_solimp", _arr(D, MAXCON, 5)), ("con_mu", _arr(D, MAXCON)),
        ("nefc", I), ("efc_id", _arr(I, MAXEFC)),
        ("efc_J", _arr(D, MAXEFC, MAXV)), ("efc_pos", _arr(D, MAXEFC)), ("efc_margin", _arr(D, MAXEFC)),
        ("efc_diagApprox", _arr(D, MAXEFC)),
        ("efc_R", _arr(D, MAXEFC)), ("efc_D", _arr(D, MAXEFC)), ("efc_KBIP", _arr(D, MAXEFC, 4)),
        ("efc_vel", _arr(D, MAXEFC)), ("efc_aref", _arr(D, MAXEFC)),
        ("efc_force", _arr(D, MAXEFC)),
        ("cvel", _arr(D, MAXB, 6)), ("cdof_dot", _arr(D, MAXV, 6)),
        ("actuator_length", _arr(D, MAXU)), ("actuator_velocity", _arr(D, MAXU)),
        ("actuator_force", _arr(D, MAXU)),
        ("qfrc_passive", _arr(D, MAXV)), ("qfrc_bias", _arr(D, MAXV)), ("qfrc_actuator", _arr(D, MAXV)),
        ("qfrc_smooth", _arr(D, MAXV)),
        ("qacc_smooth", _arr(D, MAXV)), ("qfrc_constraint", _arr(D, MAXV)), ("qacc", _arr(D, MAXV)),
        ("sensordata", _arr(D, 3)),
        ("solver_iter", I), ("warning_bad", I),
        ("total_solver_iter", C.c_long), ("total_substeps", C.c_long), ("total_contacts", C.c_long),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        assert _lib.jbo_sizeof_model() == C.sizeof(JboModel), "JboModel layout mismatch"
        assert _lib.jbo_sizeof_data() == C.sizeof(JboData), "JboData layout mismatch"
        for fn in ("jbo_reset_data", "jbo_forward", "jbo_step1", "jbo_step2", "jbo_physics_step"):
            getattr(_lib, fn).argtypes = [C.POINTER(JboModel), C.POINTER(JboData)]
            getattr(_lib, fn).restype = None
        _lib.jbo_control_step.argtypes = [C.POINTER(JboModel), C.POINTER(JboData), C.POINTER(D), I]
        _lib.jbo_control_step.restype = None
        _lib.jbo_task_obs_reward.argtypes = [C.POINTER(JboModel), C.POINTER(JboData), I, C.POINTER(D), C.POINTER(D)]
        _lib.jbo_task_obs_reward.restype = I
        _lib.jbo_rollout.argtypes = [C.POINTER(JboModel), C.POINTER(JboData), I, I, C.POINTER(D), I, I, I,
                                     C.POINTER(D), C.POINTER(D)]
        _lib.jbo_rollout.restype = None
        _lib.jbo_heuristic.argtypes = [I, C.POINTER(D)]
        _lib.jbo_heuristic.restype = D
        _lib.jbo_rollout_heuristic.argtypes = [C.POINTER(JboModel), C.POINTER(JboData), I, I, I, I, I, C.POINTER(D)]
        _lib.jbo_rollout_heuristic.restype = None
    return _lib


def heuristic_c(task: int, obs) -> float:
    """The C restatement of heuristic_policies.py used by the closed-loop rollouts (pinned by tests/test_reference_golden.py)."""
    o = (D * 32)(*[float(x) for x in np.ravel(obs)])
    return float(lib().jbo_heuristic(int(task), o))


def batch_rollout(cm, task: int, qpos0, targets, n_steps: int, actions=None, n_sub_steps: int = 50, threads: int = 0):
    """N independent episodes on the host threads.  ``qpos0[N, 16]`` initial configurations (velocities zero, like a
    reset), ``targets[N, 4]`` = x, y, quat_w, quat_z (the simulator's target layout); ``actions[n_steps, N]`` for an
    open-loop run, ``None`` for the closed loop under the task's heuristic policy (evaluate_policy.py:10-35).
    Returns rewards[n_steps, N]."""
    import os
    L = lib()
    n = len(qpos0)
    models, datas = (JboModel * n)(), (JboData * n)()
    proto = make_model(cm)
    for e in range(n):
        C.memmove(C.byref(models[e]), C.byref(proto), C.sizeof(JboModel))
        b = models[e].target_body
        models[e].body_pos[b][0], models[e].body_pos[b][1] = float(targets[e][0]), float(targets[e][1])
        models[e].body_quat[b][0], models[e].body_quat[b][3] = float(targets[e][2]), float(targets[e][3])
        L.jbo_reset_data(C.byref(models[e]), C.byref(datas[e]))
        for i in range(16):
            datas[e].qpos[i] = float(qpos0[e][i])
        L.jbo_forward(C.byref(models[e]), C.byref(datas[e]))
    rew = np.zeros((n_steps, n))
    dp = C.POINTER(D)
    threads = threads or (os.cpu_count() or 1)
    if actions is None:
        L.jbo_rollout_heuristic(models, datas, n, int(task), n_steps, n_sub_steps, threads, rew.ctypes.data_as(dp))
    else:
        a = np.ascontiguousarray(actions, dtype=np.float64)
        assert a.shape == (n_steps, n)
        L.jbo_rollout(models, datas, n, int(task), a.ctypes.data_as(dp), n_steps, n_sub_steps, threads,
                      rew.ctypes.data_as(dp), None)
    return rew


def _np_view(struct, name):
    """numpy view (no copy) of an array field of a ctypes struct."""
    f = getattr(struct, name)
    return np.ctypeslib.as_array(f)


def make_model(cm) -> JboModel:
    """Fill a JboModel from a ``jitterbug_b200.mjcf.CompiledModel``-shaped object (duck-typed)."""
    m = JboModel()
    for k in ("nq", "nv", "nu", "nbody", "njnt", "ngeom", "iterations", "sensor_body"):
        setattr(m, k, int(getattr(cm, k)))
    for k in ("timestep", "tolerance", "impratio", "meaninertia"):
        setattr(m, k, float(getattr(cm, k)))
    _np_view(m, "gravity")[:] = cm.gravity
    skip = {"nq", "nv", "nu", "nbody", "njnt", "ngeom", "iterations", "sensor_body", "timestep", "gravity",
            "tolerance", "impratio", "meaninertia", "target_body", "target_geom"}
    for name, _ in JboModel._fields_:
        if name in skip:
            continue
        src = np.asarray(getattr(cm, name))
        dst = _np_view(m, name)
        dst[...] = 0
        dst[tuple(slice(0, s) for s in src.shape)] = src
    m.target_body = cm.body_names.index("target") if "target" in cm.body_names else -1
    m.target_geom = cm.geom_names.index("target") if "target" in cm.geom_names else -1
    return m


class OracleSim:
    """One fp64 environment: the moral equivalent of ``dm_control.mujoco.Physics`` for this model."""

    def __init__(self, cm):
        self.cm = cm
        self.m = make_model(cm)
        self.d = JboData()
        self.L = lib()
        self.reset()

    # -- state
    def reset(self):
        self.L.jbo_reset_data(C.byref(self.m), C.byref(self.d))

    def view(self, name):
        return _np_view(self.d, name)

    @property
    def qpos(self):
        return _np_view(self.d, "qpos")[:self.cm.nq]

    @property
    def qvel(self):
        return _np_view(self.d, "qvel")[:self.cm.nv]

    @property
    def warmstart(self):
        return _np_view(self.d, "qacc_warmstart")[:self.cm.nv]

    def set_state(self, qpos, qvel, warmstart=None, ctrl=0.0):
        self.qpos[:] = qpos
        self.qvel[:] = qvel
        self.warmstart[:] = 0 if warmstart is None else warmstart
        self.d.ctrl[0] = float(ctrl)

    def set_target(self, x=None, y=None, yaw=None):
        """model.body_pos/body_quat["target"] writes of jitterbug.py:613-651."""
        b = self.m.target_body
        if x is not None:
            self.m.body_pos[b][0] = x
        if y is not None:
            self.m.body_pos[b][1] = y
        if yaw is not None:
            q = [np.cos(yaw / 2), 0.0, 0.0, np.sin(yaw / 2)]
            for k in range(4):
                self.m.body_quat[b][k] = q[k]

    # -- pipeline
    def forward(self):
        self.L.jbo_forward(C.byref(self.m), C.byref(self.d))

    def step1(self):
        self.L.jbo_step1(C.byref(self.m), C.byref(self.d))

    def step2(self):
        self.L.jbo_step2(C.byref(self.m), C.byref(self.d))

    def physics_step(self):
        self.L.jbo_physics_step(C.byref(self.m), C.byref(self.d))

    def control_step(self, ctrl, n_sub_steps=50):
        c = (D * 1)(float(ctrl))
        self.L.jbo_control_step(C.byref(self.m), C.byref(self.d), c, int(n_sub_steps))

    def contacts(self):
        n = self.d.ncon
        g2 = _np_view(self.d, "con_geom2")[:n].copy()
        dist = _np_view(self.d, "con_dist")[:n].copy()
        excl = _np_view(self.d, "con_exclude")[:n].copy()
        pos = _np_view(self.d, "con_pos")[:n].copy()
        return g2, dist, excl, pos

    def obs_reward_c(self, task: int):
        obs = (D * 32)()
        r = D()
        n = self.L.jbo_task_obs_reward(C.byref(self.m), C.byref(self.d), int(task), obs, C.byref(r))
        return np.array(obs[:n]), r.value


# ----------------------------------------------------------------------------------------------------------------
# numpy restatement of the reference's task layer
# ----------------------------------------------------------------------------------------------------------------
def _quat2mat(q):
    w, x, y, z = q
    return np.array([
        [w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)],
        [2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)],
        [2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z],
    ])


def _wrap(angle):
    while angle > np.pi:
        angle -= 2 * np.pi
    while angle <= -np.pi:
        angle += 2 * np.pi
    return angle


def tolerance(x, bounds=(0.0, 0.0), margin=0.0, sigmoid="gaussian", value_at_margin=0.1):
    """dm_control.utils.rewards.tolerance, scalar form (call sites jitterbug.py:846, :860, :874, :884)."""
    lower, upper = bounds
    in_bounds = lower <= x <= upper
    if margin == 0:
        return 1.0 if in_bounds else 0.0
    if in_bounds:
        return 1.0
    d = (lower - x if x < lower else x - upper) / margin
    if sigmoid == "gaussian":
        scale = np.sqrt(-2 * np.log(value_at_margin))
        return float(np.exp(-0.5 * (d * scale) ** 2))
    if sigmoid == "linear":
        sx = d * (1 - value_at_margin)
        return float(1 - sx) if abs(sx) < 1 else 0.0
    if sigmoid == "cosine":
        sx = d * np.arccos(2 * value_at_margin - 1) / np.pi
        return float((1 + np.cos(np.pi * sx)) / 2) if abs(sx) < 1 else 0.0
    raise ValueError(sigmoid)


_NORM_ALL = np.array([[-2.0, 2.0], [-2.0, 2.0], [0.0, 0.1], [-1, 1], [-1, 1], [-1, 1], [-1, 1],
                      [-1, 1], [-1, 1], [-1, 1], [-35, 35], [-35, 35], [-35, 35],
                      [-np.pi, np.pi], [-180.0, 180.0]])


def _norm(v, lo, hi):
    return (v - lo) / (hi - lo) * 2.0 - 1.0


class TaskOracle:
    """Reads a state the way ``Physics`` + ``Jitterbug`` do (jitterbug.py:177-317, :673-925)."""

    def __init__(self, sim: OracleSim, task: str):
        assert task in TASKS
        self.sim, self.task = sim, task

    # Physics helpers
    def jitterbug_direction_yaw(self):
        mat = _quat2mat(self.sim.qpos[3:7])
        return np.arctan2(mat[1, 0], mat[0, 0]) - np.pi / 2

    def motor_position(self):
        return _wrap(self.sim.qpos[15] + np.pi / 2)

    def target_position_xyz(self):
        return self.sim.view("geom_xpos")[self.sim.m.target_geom].copy()

    def target_position_quat(self):
        return self.sim.view("xquat")[self.sim.m.target_body].copy()

    def target_direction_yaw(self):
        mat = _quat2mat(self.target_position_quat())
        return np.arctan2(mat[1, 0], mat[0, 0])

    def target_position_in_jitterbug_frame(self):
        rel = self.target_position_xyz() - self.sim.qpos[0:3]
        return np.linalg.inv(_quat2mat(self.sim.qpos[3:7])) @ rel

    def jitterbug_velocity_in_target_frame(self):
        vel = self.sim.view("sensordata")[:3].copy()
        return np.linalg.inv(_quat2mat(self.target_position_quat())) @ vel

    def angle_jitterbug_to_target(self):
        return np.array([_wrap(self.target_direction_yaw() - self.jitterbug_direction_yaw())])

    # Task
    def get_observation(self):
        s = self.sim
        obs = collections.OrderedDict()
        obs["position"] = _norm(s.qpos[0:7].copy(), _NORM_ALL[0:7, 0], _NORM_ALL[0:7, 1])
        obs["velocity"] = _norm(s.qvel[0:6].copy(), _NORM_ALL[7:13, 0], _NORM_ALL[7:13, 1])
        obs["motor_position"] = np.array([_norm(self.motor_position(), -np.pi, np.pi)])
        obs["motor_velocity"] = np.array([_norm(s.qvel[14], -180.0, 180.0)])
        t = self.task
        if t in ("face_direction", "move_in_direction"):
            obs["angle_to_target"] = _norm(self.angle_jitterbug_to_target(), -np.pi, np.pi)
        if t == "move_in_direction":
            obs["speed_in_target_frame"] = _norm(self.jitterbug_velocity_in_target_frame(), -1.0, 1.0)
        if t in ("move_to_position", "move_to_pose"):
            obs["target_in_jitterbug_frame"] = _norm(
                self.target_position_in_jitterbug_frame(), np.array([-3.0, -3.0, -0.1]), np.array([3.0, 3.0, 0.1]))
        if t == "move_to_pose":
            obs["angle_to_target"] = _norm(self.angle_jitterbug_to_target(), -np.pi, np.pi)
        return obs

    def heading_reward(self):
        return tolerance(self.angle_jitterbug_to_target()[0], bounds=(0, 0), margin=np.pi / 2,
                         value_at_margin=0, sigmoid="cosine")

    def velocity_reward(self):
        return tolerance(self.jitterbug_velocity_in_target_frame()[0], bounds=(0.1, float("inf")), margin=0.1,
                         value_at_margin=0, sigmoid="linear")

    def position_reward(self):
        return tolerance(np.linalg.norm(self.target_position_in_jitterbug_frame()), bounds=(0, 0), margin=0.05)

    def upright_reward(self):
        return tolerance(self.sim.view("xmat")[1][8], bounds=(1, 1), margin=0.5)

    def get_reward(self):
        t = self.task
        if t == "move_from_origin":
            r = 1 - self.position_reward()
        elif t == "face_direction":
            r = self.heading_reward()
        elif t == "move_in_direction":
            r = self.velocity_reward()
        elif t == "move_to_position":
            r = self.position_reward()
        else:
            r = self.position_reward() * self.heading_reward()
        return r * self.upright_reward()


# ---- heuristic policies on an observation dict (heuristic_policies.py:6-136)
def h_face_direction(obs, angle_to_target=None):
    if angle_to_target is None:
        angle_to_target = obs["angle_to_target"]
    return 0.9 * max(min(3 * float(np.ravel(angle_to_target)[0]) / np.pi, 1), -1)


def h_move_forward(obs, kick_angle=np.deg2rad(45), speed=0.3, orientation="forward"):
    motor_angle = obs["motor_position"][0]
    motor_vel = obs["motor_velocity"][0]
    offset = {"left": np.pi / 2, "right": -np.pi / 2}.get(orientation, 0)
    if motor_angle < offset - kick_angle:
        return speed
    if motor_angle > offset + kick_angle:
        return -speed
    return speed if motor_vel > 0 else -speed


def h_optimal_orientation(angle):
    orientation = "None"
    d45 = 45 * np.pi / 180
    if -d45 <= angle <= d45:
        orientation = "forward"
    elif d45 < angle <= np.pi:
        orientation = "left"
        angle = abs(abs(angle) - 90 * np.pi / 180)
    elif -np.pi <= angle < -d45:
        orientation = "right"
        angle = -abs(abs(angle) - 90 * np.pi / 180)
    return angle, orientation


def heuristic(task, obs, angle_threshold=np.deg2rad(20)):
    if task == "move_from_origin":
        return h_move_forward(obs)
    if task == "face_direction":
        return h_face_direction(obs)
    if task == "move_in_direction":
        a, o = h_optimal_orientation(float(obs["angle_to_target"][0]))
        return h_face_direction(obs, a) if abs(a) > angle_threshold else h_move_forward(obs, orientation=o)
    dx, dy = obs["target_in_jitterbug_frame"][0:2]
    a = np.arctan2(dx, -dy)
    if task == "move_to_position":
        a, o = h_optimal_orientation(a)
        return h_face_direction(obs, a) if abs(a) > angle_threshold else h_move_forward(obs, orientation=o)
    if abs(a) > angle_threshold:
        return h_face_direction(obs, a)
    if np.linalg.norm([dx, dy]) > 0.01:
        return h_move_forward(obs)
    return h_face_direction(obs)


# ---------------------------------------------------------------------------------------------- observation encoders
def normalize_obs(obs, extrema, zero_one=False):
    """denoising_autoencoder.py:218-232 (normalize_obs) and its 0..1 sibling normalize_obs01."""
    obs, ex = np.asarray(obs, dtype=np.float64), np.asarray(extrema, dtype=np.float64)
    u = (obs - ex[:, 0]) / (ex[:, 1] - ex[:, 0])
    return u if zero_one else -1.0 + 2.0 * u


def dae_encode(x, w1, b1):
    """Hidden layer of the tanh auto-encoder: hid_layer1 = tanh(X w1 + b1) (denoising_autoencoder.py:68, :98, :195-205)."""
    return np.tanh(np.asarray(x, dtype=np.float64) @ w1 + b1)


def vae_encode(x, fc1, fc2, out_mean, out_std, eps):
    """VAE.Encoder.forward (VAE.py:32-46) then VAE.encode (:134-146): z = mean + std * eps, std head used as is.
    Each layer is (W[in, out], b[out])."""
    h = np.maximum(np.asarray(x, dtype=np.float64) @ fc1[0] + fc1[1], 0.0)
    h = np.maximum(h @ fc2[0] + fc2[1], 0.0)
    return (h @ out_mean[0] + out_mean[1]) + (h @ out_std[0] + out_std[1]) * eps


def encode_obs(obs, encode, k_encode=None, extrema=None, zero_one=False):
    """Jitterbug.encode_obs (jitterbug.py:927-993) for one flat observation: normalise (if extrema are given), encode the
    first k features with ``encode`` (a callable or a list of callables = use_several_autoencoders, averaged :983-991),
    append the rest unchanged (:957-981)."""
    v = np.asarray(obs, dtype=np.float64)
    if extrema is not None:
        v = normalize_obs(v, extrema, zero_one)
    k = len(v) if k_encode is None else k_encode
    encs = encode if isinstance(encode, (list, tuple)) else [encode]
    z = sum(e(v[:k]) for e in encs) / len(encs)
    return np.concatenate([z, v[k:]])
