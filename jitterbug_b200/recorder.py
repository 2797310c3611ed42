"""MuJoCo-layout state interchange and trajectory recording.

Dumps per-control-step ``qpos`` (16), ``qvel`` (15), ``ctrl`` (1), ``qacc_warmstart`` (15), ``time``, plus the task
observation and reward, in ``mjData`` field order, so a MuJoCo installation elsewhere can replay the same states for an
offline parity check; ``.mat`` output matches what the reference's notebooks read (observation dumps,
fig-heatmap.ipynb cell 2), ``.npz`` is the compact alternative.

``qacc_warmstart`` caveat: with active contacts it is the constraint solver's ``qacc`` like MuJoCo's; in contact-free
sub-steps the kernel keeps the Euler (implicit-damping) acceleration where MuJoCo stores ``M^-1 f_smooth`` -- they
differ by O(h * damping / inertia) ~ 1e-3 relative.  The field is only the next solve's initial guess (the minimiser does
not depend on it), so replays from these dumps reproduce qpos / qvel, not this field, in airborne steps.

``python -m jitterbug_b200.recorder out.npz [--envs 256] [--steps 40] [--task move_from_origin]`` writes a replay file
(states, controls and the simulator's next states) for ``tests/scripts/replay_with_mujoco.py``.
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


def dump_replay(path: str, task: str = "move_from_origin", num_envs: int = 256, steps: int = 40, seed: int = 0,
                device="cuda:0") -> dict:
    """Record ``steps`` control steps of ``num_envs`` envs under uniform random actions, with the contact set at every
    recorded state: everything a MuJoCo installation needs to re-simulate each control step from the identical state
    (``tests/scripts/replay_with_mujoco.py``).  Entry k+1 is the simulator's result of stepping entry k with ctrl[k+1]."""
    env = BatchedJitterbugEnv(task, num_envs=num_envs, seed=seed, device=device, time_limit=float("inf"))
    gen = torch.Generator(device=env.device).manual_seed(seed)
    obs = env.reset()
    keys = ("qpos", "qvel", "qacc_warmstart", "ctrl", "target", "ncon", "contact_geom", "contact_dist", "observation", "reward")
    out = {k: [] for k in keys}
    act = torch.zeros(num_envs, device=env.device)
    rew = torch.zeros(num_envs, device=env.device)
    for k in range(steps + 1):
        qpos, qvel, warm = env.get_state()
        ncon, geom, dist = env.contacts(16)
        for name, t in (("qpos", qpos), ("qvel", qvel), ("qacc_warmstart", warm), ("ctrl", act), ("target", env.get_target()),
                        ("ncon", ncon), ("contact_geom", geom), ("contact_dist", dist), ("observation", obs), ("reward", rew)):
            out[name].append(t.double().cpu().numpy() if t.is_floating_point() else t.cpu().numpy())
        if k < steps:
            act = torch.rand(num_envs, device=env.device, generator=gen) * 2 - 1
            obs, rew, _ = env.step(act)
            obs, rew = obs.clone(), rew.clone()
    res = {k: np.stack(v) for k, v in out.items()}
    res["task"] = np.array(task)
    res["n_sub_steps"] = np.array(env.n_sub_steps)
    save(path, res)
    return res


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("path")
    ap.add_argument("--task", default="move_from_origin")
    ap.add_argument("--envs", type=int, default=256)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()
    r = dump_replay(a.path, a.task, a.envs, a.steps, a.seed)
    print(f"wrote {a.path}: {r['qpos'].shape[0] - 1} control steps x {r['qpos'].shape[1]} envs")
