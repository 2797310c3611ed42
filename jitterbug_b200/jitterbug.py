"""The Jitterbug domain: drop-in for ``jitterbug_dmc/jitterbug.py`` on top of the batched CUDA simulator.

Same public surface as the reference module: ``SUITE`` with the five tagged task factories (jitterbug.py:72-174),
``Physics`` with the derived-quantity helpers (jitterbug.py:177-317), ``Jitterbug`` task (jitterbug.py:320-431 --
canonical behaviour: raw, always-normalised observation dict; the autoencoder experiments hard-wired at HEAD,
jitterbug.py:468-599, are research scaffolding and not part of the documented API), and
``control.Environment``-style ``reset()/step()/action_spec()/observation_spec()`` semantics.

A single-env ``Environment`` is a batch of one on the GPU; the batched API is :class:`BatchedJitterbugEnv`.
"""

from __future__ import annotations

import collections
import os

import numpy as np

from . import mjcf
from .batched_env import (BatchedJitterbugEnv, DEFAULT_CONTROL_TIMESTEP, DEFAULT_TIME_LIMIT, OBS_DIM, STEP_FIRST,
                          STEP_LAST, TASKS, observation_layout)
from .dm import Array, BoundedArray, StepType, TaggedTasks, TimeStep

SUITE = TaggedTasks()
TARGET_SPEED = 0.1


def get_model_and_assets():
    """Returns a tuple containing the model XML string and a dict of assets (no visual assets are needed here)."""
    with open(mjcf.DEFAULT_XML, "rb") as f:
        return f.read(), {}


def tolerance(x, bounds=(0.0, 0.0), margin=0.0, sigmoid='gaussian', value_at_margin=0.1):
    """Scalar ``dm_control.utils.rewards.tolerance`` for the three sigmoids the Jitterbug rewards use (call sites
    jitterbug.py:846, :860, :874, :884): 1 inside ``bounds``, decaying to ``value_at_margin`` one ``margin`` outside."""
    lower, upper = bounds
    if lower <= x <= upper:
        return 1.0
    if margin == 0:
        return 0.0
    d = (lower - x if x < lower else x - upper) / margin
    if sigmoid == 'gaussian':
        return float(np.exp(-0.5 * (d * np.sqrt(-2 * np.log(value_at_margin))) ** 2))
    if sigmoid == 'linear':
        sx = d * (1 - value_at_margin)
        return float(1 - sx) if abs(sx) < 1 else 0.0
    if sigmoid == 'cosine':
        sx = d * np.arccos(2 * value_at_margin - 1) / np.pi
        return float((1 + np.cos(np.pi * sx)) / 2) if abs(sx) < 1 else 0.0
    raise ValueError("Unknown sigmoid type {!r}".format(sigmoid))


def _quat2mat(q):
    return mjcf.quat2mat(np.asarray(q, dtype=np.float64))


class Physics:
    """State views and derived quantities of one env, mirroring ``jitterbug.Physics`` (jitterbug.py:177-317)."""

    def __init__(self, batched: BatchedJitterbugEnv, index: int = 0, model=None):
        self._b, self._i = batched, index
        self.model = model if model is not None else batched.model
        self._cache = None

    @classmethod
    def from_state(cls, model, qpos, qvel, target=(0.0, 0.0, 1.0, 0.0)):
        """Detached host-side physics view of a given state (``target`` = x, y, quat_w, quat_z): the helpers and the
        task's ``get_observation`` / ``get_reward`` then work without a simulator (analysis, tests)."""
        ph = cls.__new__(cls)
        ph._b, ph._i, ph.model = None, 0, model
        ph._cache = (np.asarray(qpos, dtype=np.float64).copy(), np.asarray(qvel, dtype=np.float64).copy(),
                     np.asarray(target, dtype=np.float64).copy())
        return ph

    # ---- raw state (host copies, float64)
    def _state(self):
        if self._cache is None:
            qpos, qvel, _ = self._b.get_state()
            tgt = self._b.get_target()
            self._cache = (qpos[self._i].double().cpu().numpy(), qvel[self._i].double().cpu().numpy(),
                           tgt[self._i].double().cpu().numpy())
        return self._cache

    def _invalidate(self):
        if self._b is not None:
            self._cache = None

    @property
    def qpos(self):
        return self._state()[0]

    @property
    def qvel(self):
        return self._state()[1]

    def timestep(self):
        return float(self.model.timestep)

    # ---- helpers of the reference Physics
    def jitterbug_position(self):
        return self.qpos[0:7]

    def jitterbug_position_xyz(self):
        return self.jitterbug_position()[:3]

    def jitterbug_position_quat(self):
        return self.jitterbug_position()[3:]

    def jitterbug_direction_yaw(self):
        mat = _quat2mat(self.jitterbug_position_quat())
        return np.arctan2(mat[1, 0], mat[0, 0]) - np.pi / 2

    def jitterbug_velocity(self):
        return self.qvel[0:6]

    def jitterbug_velocity_xyz(self):
        return self.jitterbug_velocity()[:3]

    def jitterbug_velocity_rpy(self):
        return self.jitterbug_velocity()[3:]

    def motor_position(self):
        angle = self.qpos[15] + np.pi / 2
        angle = angle - 2 * np.pi * np.ceil((angle - np.pi) / (2 * np.pi))
        return np.array([angle])

    def motor_velocity(self):
        return self.qvel[14:15]

    def target_position_xyz(self):
        t = self._state()[2]
        return np.array([t[0], t[1], float(np.asarray(self.model.body_pos)[self.model.body_id("target"), 2])])

    def target_position_quat(self):
        t = self._state()[2]
        return np.array([t[2], 0.0, 0.0, t[3]])

    def target_position(self):
        return np.concatenate((self.target_position_xyz(), self.target_position_quat()), axis=0)

    def target_direction_yaw(self):
        mat = _quat2mat(self.target_position_quat())
        return np.arctan2(mat[1, 0], mat[0, 0])

    def target_position_in_jitterbug_frame(self):
        rel = self.target_position_xyz() - self.jitterbug_position_xyz()
        return _quat2mat(self.jitterbug_position_quat()).T @ rel

    def jitterbug_framelinvel(self):
        """The ``jitterbug_framelinvel`` sensor: world velocity of the root body's own COM."""
        rot = _quat2mat(self.jitterbug_position_quat())
        ipos = np.asarray(self.model.body_ipos)[self.model.body_id("jitterbug")]
        return self.qvel[0:3] + rot @ np.cross(self.qvel[3:6], ipos)

    def jitterbug_velocity_in_target_frame(self):
        return _quat2mat(self.target_position_quat()).T @ self.jitterbug_framelinvel()

    def angle_jitterbug_to_target(self):
        angle = self.target_direction_yaw() - self.jitterbug_direction_yaw()
        angle = angle - 2 * np.pi * np.ceil((angle - np.pi) / (2 * np.pi))
        return np.array([angle])

    def jitterbug_upright(self):
        """``physics.named.data.xmat['jitterbug', 'zz']`` (jitterbug.py:886): body z axis . world z."""
        return float(_quat2mat(self.jitterbug_position_quat())[2, 2])


class Jitterbug:
    """The task description (constructor kwargs and normalisation tables of ``jitterbug.Jitterbug``)."""

    _NORM_ALL = np.array([[-2.0, 2.0], [-2.0, 2.0], [0.0, 0.1], [-1.0, 1.0], [-1.0, 1.0], [-1.0, 1.0], [-1.0, 1.0],
                          [-1.0, 1.0], [-1.0, 1.0], [-1.0, 1.0], [-35.0, 35.0], [-35.0, 35.0], [-35.0, 35.0],
                          [-np.pi, np.pi], [-180.0, 180.0]])
    _NORM_TASKS = dict(
        move_from_origin=np.array([]),
        face_direction=np.array([[-np.pi, np.pi]]),
        move_in_direction=np.array([[-np.pi, np.pi], [-1.0, 1.0], [-1.0, 1.0], [-1.0, 1.0]]),
        move_to_position=np.array([[-3.0, 3.0], [-3.0, 3.0], [-0.1, 0.1]]),
        move_to_pose=np.array([[-3.0, 3.0], [-3.0, 3.0], [-0.1, 0.1], [-np.pi, np.pi]]),
    )

    def __init__(self, random=None, task="move_from_origin", random_pose=True, norm_obs=False):
        self.task_names = list(TASKS)
        assert task in self.task_names, "Invalid task {}, options are {}".format(task, self.task_names)
        self.task = task
        self.random_pose = random_pose
        self.norm_obs = norm_obs        # accepted and stored, never read: observations are always normalised (App. C3)
        if isinstance(random, np.random.RandomState):
            self.random = random
        else:
            self.random = np.random.RandomState(random)
        self.visualize_reward = False

    @staticmethod
    def _norm(v, min, max):  # noqa: A002 - reference signature
        return (v - min) / (max - min) * 2.0 - 1.0

    # ---- host-callable twins of the task methods (jitterbug.py:673-763, :840-925).  The batched simulator computes the
    # same quantities on the device (csrc/jb_task.cuh); these serve single-state analysis through ``env.task``.
    def get_observation(self, physics):
        """Returns an observation of the state and the target position (always normalised, App. C3)."""
        na, nt = Jitterbug._NORM_ALL, Jitterbug._NORM_TASKS[self.task]
        obs = collections.OrderedDict()
        obs['position'] = Jitterbug._norm(physics.jitterbug_position(), na[0:7, 0], na[0:7, 1])
        obs['velocity'] = Jitterbug._norm(physics.jitterbug_velocity(), na[7:13, 0], na[7:13, 1])
        obs['motor_position'] = Jitterbug._norm(physics.motor_position(), na[13, 0], na[13, 1])
        obs['motor_velocity'] = Jitterbug._norm(physics.motor_velocity(), na[14, 0], na[14, 1])
        if self.task in ("face_direction", "move_in_direction"):
            obs['angle_to_target'] = Jitterbug._norm(physics.angle_jitterbug_to_target(), nt[0, 0], nt[0, 1])
        if self.task == "move_in_direction":
            obs['speed_in_target_frame'] = Jitterbug._norm(physics.jitterbug_velocity_in_target_frame(), nt[1:, 0], nt[1:, 1])
        if self.task in ("move_to_position", "move_to_pose"):
            obs['target_in_jitterbug_frame'] = Jitterbug._norm(physics.target_position_in_jitterbug_frame(),
                                                               nt[0:3, 0], nt[0:3, 1])
        if self.task == "move_to_pose":
            obs['angle_to_target'] = Jitterbug._norm(physics.angle_jitterbug_to_target(), nt[3, 0], nt[3, 1])
        return obs

    def obsdict2vec(self, obs):
        """Observation dict -> (vector, column names), jitterbug.py:765-838."""
        columns = ["X", "Y", "Z", "QuatX", "QuatY", "QuatZ", "QuatW", "VelX", "VelY", "VelZ", "VelRoll", "VelPitch",
                   "VelYaw", "MotorYaw", "MotorVelYaw"]
        extra = {"face_direction": ["TargetYaw"], "move_in_direction": ["TargetYaw", "TargetVelX", "TargetVelY", "TargetVelZ"],
                 "move_to_position": ["TargetX", "TargetY", "TargetZ"],
                 "move_to_pose": ["TargetX", "TargetY", "TargetZ", "TargetYaw"]}.get(self.task, [])
        return np.concatenate([np.ravel(v) for v in obs.values()]), columns + extra

    def heading_reward(self, physics):
        return tolerance(physics.angle_jitterbug_to_target()[0], bounds=(0, 0), margin=np.pi / 2, value_at_margin=0,
                         sigmoid='cosine')

    def velocity_reward(self, physics):
        return tolerance(physics.jitterbug_velocity_in_target_frame()[0], bounds=(TARGET_SPEED, float('inf')),
                         margin=TARGET_SPEED, value_at_margin=0, sigmoid='linear')

    def position_reward(self, physics):
        return tolerance(np.linalg.norm(physics.target_position_in_jitterbug_frame()), bounds=(0, 0), margin=0.05)

    def upright_reward(self, physics):
        return tolerance(physics.jitterbug_upright(), bounds=(1, 1), margin=0.5)

    def get_reward(self, physics):
        if self.task == "move_from_origin":
            r = 1 - self.position_reward(physics)
        elif self.task == "face_direction":
            r = self.heading_reward(physics)
        elif self.task == "move_in_direction":
            r = self.velocity_reward(physics)
        elif self.task == "move_to_position":
            r = self.position_reward(physics)
        else:
            r = self.position_reward(physics) * self.heading_reward(physics)
        return r * self.upright_reward(physics)


class Environment:
    """``dm_control.rl.control.Environment`` semantics for one Jitterbug running on the GPU (a batch of one)."""

    def __init__(self, physics_model, task: Jitterbug, time_limit=DEFAULT_TIME_LIMIT,
                 control_timestep=DEFAULT_CONTROL_TIMESTEP, flat_observation=False, device="cuda:0", **_unused):
        seed = int(task.random.randint(0, 2 ** 31 - 1))
        self._batched = BatchedJitterbugEnv(task.task, num_envs=1, device=device, seed=seed,
                                            random_pose=task.random_pose, time_limit=time_limit,
                                            control_timestep=control_timestep, model=physics_model)
        self._task = task
        self._physics = Physics(self._batched, 0, physics_model)
        self._flat_observation = flat_observation
        self._n_sub_steps = self._batched.n_sub_steps
        self._step_limit = float("inf") if self._batched.step_limit == 0 else float(self._batched.step_limit)
        self._layout = observation_layout(task.task)
        self._reset_next_step = True
        import torch
        self._torch = torch
        self._action = torch.zeros(1, device=self._batched.device)

    # -- dm_env.Environment
    def _observation(self, obs_row):
        obs = collections.OrderedDict((k, obs_row[s].astype(np.float64)) for k, s in self._layout.items())
        if self._flat_observation:
            return collections.OrderedDict(observations=np.concatenate([v.ravel() for v in obs.values()]))
        return obs

    def reset(self):
        self._reset_next_step = False
        obs = self._batched.reset()
        self._physics._invalidate()
        return TimeStep(StepType.FIRST, None, None, self._observation(obs[0].cpu().numpy()))

    def step(self, action):
        if self._reset_next_step:
            return self.reset()
        self._action[0] = float(np.asarray(action, dtype=np.float64).reshape(-1)[0])
        obs, reward, step_type = self._batched.step(self._action)
        self._physics._invalidate()
        st = int(step_type[0])
        if st == STEP_FIRST:            # diverged physics: the kernel restarted the episode (PhysicsError upstream)
            return TimeStep(StepType.FIRST, None, None, self._observation(obs[0].cpu().numpy()))
        ts_obs = self._observation(obs[0].cpu().numpy())
        if st == STEP_LAST:
            self._reset_next_step = True
            return TimeStep(StepType.LAST, float(reward[0]), 1.0, ts_obs)
        return TimeStep(StepType.MID, float(reward[0]), 1.0, ts_obs)

    def action_spec(self):
        m = self._physics.model
        return BoundedArray(shape=(1,), dtype=np.float64, minimum=m.actuator_ctrlrange[:, 0],
                            maximum=m.actuator_ctrlrange[:, 1], name="action")

    def observation_spec(self):
        if self._flat_observation:
            return collections.OrderedDict(observations=Array((OBS_DIM[self._task.task],), np.float64, "observations"))
        return collections.OrderedDict((k, Array((s.stop - s.start,), np.float64, k)) for k, s in self._layout.items())

    def reward_spec(self):
        return Array(shape=(), dtype=np.float64, name="reward")

    def discount_spec(self):
        return BoundedArray(shape=(), dtype=np.float64, minimum=0.0, maximum=1.0, name="discount")

    def close(self):
        self._batched.close()

    @property
    def physics(self):
        return self._physics

    @property
    def task(self):
        return self._task

    def control_timestep(self):
        return self._batched.control_timestep()

    @property
    def batched(self):
        return self._batched


def _make(task_name, time_limit, control_timestep, random, environment_kwargs, model=None, **kwargs):
    physics_model = model if model is not None else mjcf.compile_xml(get_model_and_assets()[0])
    task = Jitterbug(random=random, task=task_name, **kwargs)
    environment_kwargs = environment_kwargs or {}
    return Environment(physics_model, task, time_limit=time_limit, control_timestep=control_timestep,
                       **environment_kwargs)


@SUITE.add("benchmarking", "easy")
def move_from_origin(time_limit=DEFAULT_TIME_LIMIT, control_timestep=DEFAULT_CONTROL_TIMESTEP, random=None,
                     environment_kwargs=None, **kwargs):
    """Move the Jitterbug away from the origin"""
    return _make("move_from_origin", time_limit, control_timestep, random, environment_kwargs, **kwargs)


@SUITE.add("benchmarking", "easy")
def face_direction(time_limit=DEFAULT_TIME_LIMIT, control_timestep=DEFAULT_CONTROL_TIMESTEP, random=None,
                   environment_kwargs=None, **kwargs):
    """Move the Jitterbug to face a certain yaw angle"""
    return _make("face_direction", time_limit, control_timestep, random, environment_kwargs, **kwargs)


@SUITE.add("benchmarking", "easy")
def move_in_direction(time_limit=DEFAULT_TIME_LIMIT, control_timestep=DEFAULT_CONTROL_TIMESTEP, random=None,
                      environment_kwargs=None, **kwargs):
    """Move the Jitterbug in a certain direction"""
    return _make("move_in_direction", time_limit, control_timestep, random, environment_kwargs, **kwargs)


@SUITE.add("benchmarking", "hard")
def move_to_position(time_limit=DEFAULT_TIME_LIMIT, control_timestep=DEFAULT_CONTROL_TIMESTEP, random=None,
                     environment_kwargs=None, **kwargs):
    """Move the Jitterbug to a certain XYZ position"""
    return _make("move_to_position", time_limit, control_timestep, random, environment_kwargs, **kwargs)


@SUITE.add("benchmarking", "hard")
def move_to_pose(time_limit=DEFAULT_TIME_LIMIT, control_timestep=DEFAULT_CONTROL_TIMESTEP, random=None,
                 environment_kwargs=None, **kwargs):
    """Move the Jitterbug to a certain XYZRPY pose"""
    return _make("move_to_pose", time_limit, control_timestep, random, environment_kwargs, **kwargs)
