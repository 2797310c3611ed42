"""Batched Jitterbug environment: N independent envs stepped in lock-step by one CUDA kernel launch per control step.

This is the batched counterpart of what ``suite.load(domain_name="jitterbug", task_name=...)`` builds in the
reference (jitterbug_dmc/jitterbug.py:72-174: ``control.Environment(Physics, Jitterbug, time_limit=10,
control_timestep=0.01)``).  ``step(actions[N])`` returns torch CUDA tensors; host code is Python/PyTorch plumbing over
the C-ABI in include/jbsim.h.
"""

from __future__ import annotations

import collections
import ctypes as C
from typing import Optional, Sequence, Union

import numpy as np
import torch

from . import _native, mjcf
from .model import JbModelF, fill_struct, flatten_model

TASKS = ("move_from_origin", "face_direction", "move_in_direction", "move_to_position", "move_to_pose")
TASK_ID = {t: i for i, t in enumerate(TASKS)}
OBS_DIM = {"move_from_origin": 15, "face_direction": 16, "move_in_direction": 19, "move_to_position": 18,
           "move_to_pose": 19}
STEP_FIRST, STEP_MID, STEP_LAST = 0, 1, 2

DEFAULT_TIME_LIMIT = 10          # jitterbug.py:56
DEFAULT_CONTROL_TIMESTEP = 0.01  # jitterbug.py:57


def observation_layout(task: str) -> "collections.OrderedDict[str, slice]":
    """Column slices of the flat observation, in the reference's OrderedDict order (jitterbug.py:673-763)."""
    lay = collections.OrderedDict()
    lay["position"] = slice(0, 7)
    lay["velocity"] = slice(7, 13)
    lay["motor_position"] = slice(13, 14)
    lay["motor_velocity"] = slice(14, 15)
    if task == "face_direction":
        lay["angle_to_target"] = slice(15, 16)
    elif task == "move_in_direction":
        lay["angle_to_target"] = slice(15, 16)
        lay["speed_in_target_frame"] = slice(16, 19)
    elif task == "move_to_position":
        lay["target_in_jitterbug_frame"] = slice(15, 18)
    elif task == "move_to_pose":
        lay["target_in_jitterbug_frame"] = slice(15, 18)
        lay["angle_to_target"] = slice(18, 19)
    return lay


class BatchedJitterbugEnv:
    """N Jitterbugs on one GPU.

    Args:
        task: a task name, or a sequence of N task names / ids for a mixed batch (observations are then padded with
            zeros to the widest task in the batch).
        num_envs: N.
        device: CUDA device (``"cuda:0"`` or an ordinal).  There is no CPU path.
        seed: Philox key for reset randomisation; streams are keyed by global env id (``env_id_offset + i``).
        random_pose / time_limit / control_timestep: the reference factory kwargs (jitterbug.py:73-79, :374-380).
        xml / model: MJCF text or a pre-compiled :class:`mjcf.CompiledModel`; default = packaged jitterbug.xml.
        models: optional sequence of compiled models (per-env augmented Jitterbugs): either N models, or K models
            together with ``model_index`` (N ints in [0, K)) assigning one of them to every env.
        packed_models: float32 CUDA tensor ``[N, jb_packed_model_size()]`` from
            ``augmented_jitterbug.compile_models_device`` (models compiled on the device); it is copied.
    """

    def __init__(self, task: Union[str, Sequence] = "move_from_origin", num_envs: int = 1, device="cuda:0", seed: int = 0,
                 random_pose: bool = True, time_limit: float = DEFAULT_TIME_LIMIT,
                 control_timestep: float = DEFAULT_CONTROL_TIMESTEP, xml: Optional[str] = None,
                 model: Optional[mjcf.CompiledModel] = None, models: Optional[Sequence] = None,
                 model_index: Optional[Sequence[int]] = None, env_id_offset: int = 0, packed_models=None):
        self._lib = _native.load()
        if not torch.cuda.is_available():
            raise RuntimeError("BatchedJitterbugEnv needs a CUDA device (no CPU fallback)")
        self.device = torch.device(device if not isinstance(device, int) else f"cuda:{device}")
        if self.device.type != "cuda":
            raise RuntimeError("BatchedJitterbugEnv needs a CUDA device (no CPU fallback)")
        self.num_envs = int(num_envs)
        if isinstance(task, str):
            names = [task] * self.num_envs
        else:
            names = [TASKS[int(t)] if not isinstance(t, str) else t for t in task]
            if len(names) != self.num_envs:
                raise ValueError("len(task) must equal num_envs")
        for t in names:
            assert t in TASKS, "Invalid task {}, options are {}".format(t, TASKS)   # jitterbug.py:425
        self.task_names = names
        self.task = names[0] if len(set(names)) == 1 else None
        self.task_ids = np.array([TASK_ID[t] for t in names], dtype=np.int32)
        self.obs_dim = max(OBS_DIM[t] for t in set(names))

        index_ptr = None
        if packed_models is not None:
            if models is not None or model is not None or xml is not None:
                raise ValueError("packed_models excludes xml / model / models")
            if (not packed_models.is_cuda or packed_models.dtype != torch.float32 or not packed_models.is_contiguous()
                    or tuple(packed_models.shape) != (self.num_envs, self._lib.jb_packed_model_size())
                    or packed_models.device != self.device):
                raise ValueError("packed_models must be a contiguous float32 [num_envs, jb_packed_model_size()] tensor "
                                 "on the env's device")
            self.model = None
            model_ptr, n_models = None, self.num_envs
            timestep = float(packed_models[0, 0])          # G_H: the time step leads the packed layout
        elif models is not None:
            if model_index is None and len(models) != self.num_envs:
                raise ValueError("len(models) must equal num_envs (or pass model_index)")
            self.model = models[0]
            structs = (JbModelF * len(models))()
            for i, cm in enumerate(models):
                structs[i] = cm if isinstance(cm, JbModelF) else fill_struct(flatten_model(cm))
            model_ptr, n_models = structs, len(models)
            timestep = float(structs[0].timestep)
            if model_index is not None:
                self.model_index = np.ascontiguousarray(model_index, dtype=np.int32)
                if self.model_index.shape != (self.num_envs,):
                    raise ValueError("model_index must have num_envs entries")
                index_ptr = self.model_index.ctypes.data_as(C.POINTER(C.c_int32))
        else:
            self.model = model if model is not None else mjcf.compile_xml(xml)
            self._struct = fill_struct(flatten_model(self.model))
            model_ptr, n_models = C.pointer(self._struct), 1
            timestep = float(self.model.timestep)
        self._models_keepalive = model_ptr

        self.n_sub_steps = int(round(control_timestep / timestep))
        self._control_timestep = self.n_sub_steps * timestep
        if time_limit == float("inf"):
            self.step_limit = 0
        else:
            self.step_limit = int(np.ceil(time_limit / (timestep * self.n_sub_steps) - 1e-9))
        cfg = _native.JbSimConfig(self.num_envs, self.device.index or 0, self.n_sub_steps, self.step_limit,
                                  int(bool(random_pose)), self.obs_dim, int(seed) & (2 ** 64 - 1), int(env_id_offset), 0)
        self._h = C.c_void_p()
        if packed_models is not None:
            _native.check(self._lib.jb_sim_create_packed(
                C.byref(self._h), C.byref(cfg), packed_models.data_ptr(), n_models,
                self.task_ids.ctypes.data_as(C.POINTER(C.c_int32))))
        else:
            _native.check(self._lib.jb_sim_create_indexed(
                C.byref(self._h), C.byref(cfg), C.cast(model_ptr, C.POINTER(JbModelF)), n_models, index_ptr,
                self.task_ids.ctypes.data_as(C.POINTER(C.c_int32))))
        with torch.cuda.device(self.device):
            self._obs = torch.zeros(self.num_envs, self.obs_dim, device=self.device, dtype=torch.float32)
            self._reward = torch.zeros(self.num_envs, device=self.device, dtype=torch.float32)
            self._step_type = torch.zeros(self.num_envs, device=self.device, dtype=torch.int32)

    # ------------------------------------------------------------------ lifecycle
    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.jb_sim_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def control_timestep(self):
        return self._control_timestep

    # ------------------------------------------------------------------ dm_control-like batched API
    def reset(self, mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Start a new episode for all envs (or those selected by a bool/uint8 mask). Returns obs[N, D]."""
        mp = None
        if mask is not None:
            mask = mask.to(device=self.device, dtype=torch.uint8).contiguous()
            mp = C.c_void_p(mask.data_ptr())
        _native.check(self._lib.jb_sim_reset(self._h, mp, C.c_void_p(self._obs.data_ptr()), self._stream()))
        return self._obs

    def step(self, actions: torch.Tensor):
        """One control step for every env.  actions: [N] or [N, 1] float tensor on the env's device.

        Returns ``(obs[N, D] f32, reward[N] f32, step_type[N] i32)``; the tensors are owned by the env and are
        overwritten by the next call (clone them to keep a history).
        """
        a = actions.reshape(-1)
        if a.shape[0] != self.num_envs:
            raise ValueError(f"expected {self.num_envs} actions, got {tuple(actions.shape)}")
        if a.device != self.device or a.dtype != torch.float32 or not a.is_contiguous():
            a = a.to(device=self.device, dtype=torch.float32).contiguous()
        _native.check(self._lib.jb_sim_step(self._h, C.c_void_p(a.data_ptr()), C.c_void_p(self._obs.data_ptr()),
                                            C.c_void_p(self._reward.data_ptr()),
                                            C.c_void_p(self._step_type.data_ptr()), self._stream()))
        return self._obs, self._reward, self._step_type

    def step_host(self, actions: np.ndarray):
        """Host-buffer entry point: numpy actions in, numpy (obs, reward, step_type) out (copies included).  The returned
        arrays are reused by the next call; pass actions in page-locked memory (e.g. ``torch.empty(N, pin_memory=True)
        .numpy()``) to skip the library's staging copy of the inputs as well."""
        a = np.ascontiguousarray(np.asarray(actions, dtype=np.float32).reshape(-1))
        if a.shape[0] != self.num_envs:
            raise ValueError(f"expected {self.num_envs} actions")
        if not hasattr(self, "_h_obs"):
            # page-locked result buffers: the library DMAs straight into them (pageable caller memory would go through
            # its staging copy); the numpy arrays returned are views of these tensors and are overwritten by the next call
            self._pin = [torch.empty((self.num_envs, self.obs_dim), dtype=torch.float32, pin_memory=True),
                         torch.empty(self.num_envs, dtype=torch.float32, pin_memory=True),
                         torch.empty(self.num_envs, dtype=torch.int32, pin_memory=True)]
            self._h_obs, self._h_rew, self._h_st = (t.numpy() for t in self._pin)
        _native.check(self._lib.jb_sim_step_host(self._h, a.ctypes.data_as(C.c_void_p),
                                                 self._h_obs.ctypes.data_as(C.c_void_p),
                                                 self._h_rew.ctypes.data_as(C.c_void_p),
                                                 self._h_st.ctypes.data_as(C.c_void_p), self._stream()))
        return self._h_obs, self._h_rew, self._h_st

    def observe(self):
        """``task.get_observation`` / ``task.get_reward`` at the current state (no physics advance): fresh
        ``(obs[N, D], reward[N])`` tensors."""
        obs = torch.empty(self.num_envs, self.obs_dim, device=self.device, dtype=torch.float32)
        rew = torch.empty(self.num_envs, device=self.device, dtype=torch.float32)
        _native.check(self._lib.jb_sim_observe(self._h, C.c_void_p(obs.data_ptr()), C.c_void_p(rew.data_ptr()),
                                               self._stream()))
        return obs, rew

    def heuristic_actions(self, obs: Optional[torch.Tensor] = None) -> torch.Tensor:
        """heuristic_policies.<task>(ts) for every env, on device (heuristic_policies.py:59-118)."""
        obs = self._obs if obs is None else obs.contiguous()
        out = torch.empty(self.num_envs, device=self.device, dtype=torch.float32)
        _native.check(self._lib.jb_sim_heuristic_actions(self._h, C.c_void_p(obs.data_ptr()),
                                                         C.c_void_p(out.data_ptr()), self._stream()))
        return out

    # ------------------------------------------------------------------ state access (parity / checkpointing)
    def get_state(self):
        """(qpos[N,16], qvel[N,15], qacc_warmstart[N,15]) in MuJoCo ordering."""
        qpos = torch.empty(self.num_envs, 16, device=self.device)
        qvel = torch.empty(self.num_envs, 15, device=self.device)
        warm = torch.empty(self.num_envs, 15, device=self.device)
        _native.check(self._lib.jb_sim_get_state(self._h, C.c_void_p(qpos.data_ptr()), C.c_void_p(qvel.data_ptr()),
                                                 C.c_void_p(warm.data_ptr()), self._stream()))
        return qpos, qvel, warm

    def set_state(self, qpos=None, qvel=None, warm=None, step_count=None, target=None):
        def prep(t, n):
            if t is None:
                return None, None
            t = torch.as_tensor(t, dtype=torch.float32, device=self.device).reshape(self.num_envs, n).contiguous()
            return t, C.c_void_p(t.data_ptr())
        q, qp = prep(qpos, 16)
        v, vp = prep(qvel, 15)
        w, wp = prep(warm, 15)
        _native.check(self._lib.jb_sim_set_state(self._h, qp, vp, wp, self._stream()))
        if target is not None:
            t, tp = prep(target, 4)
            _native.check(self._lib.jb_sim_set_target(self._h, tp, self._stream()))
        if step_count is not None:
            c = torch.as_tensor(step_count, dtype=torch.int32, device=self.device).reshape(self.num_envs).contiguous()
            _native.check(self._lib.jb_sim_set_step_count(self._h, C.c_void_p(c.data_ptr()), self._stream()))
        torch.cuda.current_stream(self.device).synchronize()

    def snapshot(self) -> torch.Tensor:
        """Exact checkpoint of the episode state of every env (uint8 device tensor): positions incl. the motor's
        revolution counter, velocities, warm start, targets, step counts, episode indices (the reset RNG key) and pending
        resets.  ``restore`` on an env built with the same arguments continues bit for bit."""
        blob = torch.empty(int(self._lib.jb_sim_snapshot_size(self._h)), dtype=torch.uint8, device=self.device)
        _native.check(self._lib.jb_sim_snapshot(self._h, C.c_void_p(blob.data_ptr()), self._stream()))
        return blob

    def restore(self, blob: torch.Tensor) -> None:
        if blob.dtype != torch.uint8 or blob.device != self.device or not blob.is_contiguous() \
                or blob.numel() != int(self._lib.jb_sim_snapshot_size(self._h)):
            raise ValueError("blob must be the contiguous uint8 tensor snapshot() returned for an env of this size and device")
        _native.check(self._lib.jb_sim_restore(self._h, C.c_void_p(blob.data_ptr()), self._stream()))

    def get_target(self) -> torch.Tensor:
        t = torch.empty(self.num_envs, 4, device=self.device)
        _native.check(self._lib.jb_sim_get_target(self._h, C.c_void_p(t.data_ptr()), self._stream()))
        return t

    def get_step_count(self) -> torch.Tensor:
        c = torch.empty(self.num_envs, dtype=torch.int32, device=self.device)
        _native.check(self._lib.jb_sim_get_step_count(self._h, C.c_void_p(c.data_ptr()), self._stream()))
        return c

    def contacts(self, max_con: int = 16):
        """Active contacts at the current state: (ncon[N], geom[N, max_con] (-1 padded), dist[N, max_con])."""
        ncon = torch.empty(self.num_envs, dtype=torch.int32, device=self.device)
        geom = torch.empty(self.num_envs, max_con, dtype=torch.int32, device=self.device)
        dist = torch.empty(self.num_envs, max_con, dtype=torch.float32, device=self.device)
        _native.check(self._lib.jb_sim_get_contacts(self._h, C.c_void_p(ncon.data_ptr()), C.c_void_p(geom.data_ptr()),
                                                    C.c_void_p(dist.data_ptr()), max_con, self._stream()))
        return ncon, geom, dist

    # ------------------------------------------------------------------ slot order of the step kernel (scheduling only)
    def set_slot_order(self, mode: Union[str, torch.Tensor] = "auto"):
        """``"auto"`` (default: load-sorted on the device after every step), ``"identity"``, or an int32 CUDA tensor
        holding a permutation of the env ids.  Results per env do not depend on it (see include/jbsim.h)."""
        if isinstance(mode, str):
            code = {"auto": 0, "identity": 1}[mode]
            _native.check(self._lib.jb_sim_set_slot_order(self._h, code, None, self._stream()))
        else:
            perm = mode.to(device=self.device, dtype=torch.int32).contiguous()
            if perm.shape != (self.num_envs,):
                raise ValueError("the slot order must have num_envs entries")
            _native.check(self._lib.jb_sim_set_slot_order(self._h, 2, C.c_void_p(perm.data_ptr()), self._stream()))

    def get_slot_order(self) -> torch.Tensor:
        perm = torch.empty(self.num_envs, dtype=torch.int32, device=self.device)
        _native.check(self._lib.jb_sim_get_slot_order(self._h, C.c_void_p(perm.data_ptr()), self._stream()))
        return perm

    def load(self) -> torch.Tensor:
        """Load record of the last step per env, int32 [N, 4]: contacts summed over the sub-steps in the fullest lane,
        sub-steps with >= 3 / >= 2 contacts in one lane, Newton trips."""
        ld = torch.empty(self.num_envs, 4, dtype=torch.int32, device=self.device)
        _native.check(self._lib.jb_sim_get_load(self._h, C.c_void_p(ld.data_ptr()), self._stream()))
        return ld

    def self_clearance(self, per_env: bool = False):
        """Self-collision monitor (see include/jbsim.h): dict with the env-states checked so far, how many had a
        conservative clearance bound below zero and the smallest bound seen (metres); with ``per_env=True`` the monitor
        is also evaluated now and the per-env bounds [N] are returned under ``"clearance"``."""
        out = (C.c_double * 3)()
        cl = torch.empty(self.num_envs, device=self.device, dtype=torch.float32) if per_env else None
        _native.check(self._lib.jb_sim_self_clearance(self._h, C.c_void_p(cl.data_ptr()) if per_env else None, out,
                                                      self._stream()))
        res = {"checked": int(out[0]), "below_zero": int(out[1]), "min_clearance_m": float(out[2])}
        if per_env:
            res["clearance"] = cl
        return res

    def launch_count(self) -> int:
        """Kernels of libjbsim launched so far for this env (step segments, slot-order sorts, task kernels)."""
        return int(self._lib.jb_sim_launch_count(self._h))

    def counters(self):
        """dict of device counters since creation (synchronises)."""
        out = (C.c_ulonglong * 5)()
        _native.check(self._lib.jb_sim_get_counters(self._h, out))
        keys = ("substeps", "newton_iterations", "contacts", "contact_overflows", "diverged_resets")
        return dict(zip(keys, [int(x) for x in out]))

    def phase_cycles(self):
        """Per-phase clock64 sums (profiling builds with -DJBQ_PHASE_TIMERS only; zeros otherwise)."""
        out = (C.c_ulonglong * 8)()
        _native.check(self._lib.jb_sim_get_phase_cycles(self._h, out))
        return [int(x) for x in out]

    def observation_layout(self, task: Optional[str] = None):
        return observation_layout(task or self.task or self.task_names[0])
