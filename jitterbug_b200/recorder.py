"""MuJoCo-layout state interchange and trajectory recording.

Dumps per-control-step ``qpos`` (16), ``qvel`` (15), ``ctrl`` (1), ``qacc_warmstart`` (15), ``time``, plus the task
observation and reward, in ``mjData`` field order, so a MuJoCo installation elsewhere can replay the same states for an
offline parity check; ``.mat`` output matches what the reference's notebooks read (observation dumps,
fig-heatmap.ipynb cell 2), ``.npz`` is the compact alternative.
"""

from __future__ import annotations

from typing import Callable, Optional

import numpy as np
import torch

from .batched_env import BatchedJitterbugEnv


def record(env: BatchedJitterbugEnv, policy: Optional[Callable[[torch.Tensor], torch.Tensor]], steps: int,
           env_ids=None):
    """Roll ``steps`` control steps and return a dict of arrays with leading shape [steps + 1, len(env_ids)].

    ``policy(obs) -> actions`` runs on device (``None`` = uniform random actions).  Entry 0 is the reset state.
    """
    ids = torch.arange(env.num_envs, device=env.device) if env_ids is None else \
        torch.as_tensor(env_ids, device=env.device, dtype=torch.long)
    obs = env.reset()
    out = {k: [] for k in ("qpos", "qvel", "qacc_warmstart", "ctrl", "time", "observation", "reward", "step_type")}

    def snap(ctrl, reward, step_type, t):
        qpos, qvel, warm = env.get_state()
        out["qpos"].append(qpos[ids].double().cpu().numpy())
        out["qvel"].append(qvel[ids].double().cpu().numpy())
        out["qacc_warmstart"].append(warm[ids].double().cpu().numpy())
        out["ctrl"].append(ctrl[ids].double().cpu().numpy()[:, None])
        out["time"].append(np.full(len(ids), t))
        out["observation"].append(env._obs[ids].double().cpu().numpy())
        out["reward"].append(reward[ids].double().cpu().numpy())
        out["step_type"].append(step_type[ids].cpu().numpy())

    zeros = torch.zeros(env.num_envs, device=env.device)
    snap(zeros, zeros, torch.zeros(env.num_envs, dtype=torch.int32, device=env.device), 0.0)
    for k in range(steps):
        act = policy(obs) if policy is not None else torch.rand(env.num_envs, device=env.device) * 2 - 1
        obs, rew, st = env.step(act)
        snap(act.reshape(-1).float(), rew, st, (k + 1) * env.control_timestep())
    res = {k: np.stack(v) for k, v in out.items()}
    res["task_names"] = np.array([env.task_names[int(i)] for i in ids.cpu().numpy()])
    res["control_timestep"] = np.array(env.control_timestep())
    return res


def save(path: str, traj: dict) -> None:
    """Write a recorded trajectory as ``.npz`` or MATLAB ``.mat`` (by extension)."""
    if path.endswith(".mat"):
        from scipy.io import savemat
        savemat(path, {k: v for k, v in traj.items()})
    else:
        np.savez_compressed(path, **traj)
