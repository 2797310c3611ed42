"""On-device rollout loop: a vectorised-env facade + torch policy stepping the batched simulator (config 5 of
BASELINE.json; replaces the reference's ``SubprocVecEnv`` + ``Monitor`` plumbing, benchmarks/benchmark.py:138-185).

Everything stays on the GPU: observations, actions, rewards, episode-return accounting.  Envs shard across GPUs by env
index (one process per GPU); the only collective is the end-of-rollout statistics all-reduce.
"""

from __future__ import annotations

import time
from typing import Dict, Optional

import torch
import torch.nn as nn

from .batched_env import STEP_LAST, BatchedJitterbugEnv
from .distributed import gather_episode_stats


class VecJitterbugEnv:
    """VecEnv-style auto-resetting facade: ``step`` returns ``(obs, reward, done)``; when an env finishes its episode
    the observation returned for it is the first observation of its next episode (stable-baselines convention) and its
    return / length are recorded in the on-device monitor."""

    def __init__(self, env: BatchedJitterbugEnv, encoder=None):
        """``encoder``: optional :class:`jitterbug_b200.encoder.ObservationEncoder` applied to every observation
        returned (the reference's ``encode_obs`` hook, jitterbug.py:927-993); ``obs_dim`` is then its output width."""
        self.env = env
        self.num_envs = env.num_envs
        self.device = env.device
        self.encoder = encoder
        self.obs_dim = env.obs_dim if encoder is None else encoder.out_dim
        self._ret = torch.zeros(self.num_envs, device=self.device)
        self._len = torch.zeros(self.num_envs, dtype=torch.int32, device=self.device)
        self.episode_returns = []          # list of tensors of finished-episode returns
        self.episode_count = 0

    def reset(self) -> torch.Tensor:
        self._ret.zero_()
        self._len.zero_()
        return self._encode(self.env.reset())

    def _encode(self, obs):
        return obs if self.encoder is None else self.encoder(obs)

    def step(self, actions: torch.Tensor):
        obs, reward, step_type = self.env.step(actions)
        self._ret += reward
        self._len += 1
        done = step_type == STEP_LAST
        if bool(done.any()):
            self.episode_returns.append(self._ret[done].clone())
            self.episode_count += int(done.sum())
            self._ret[done] = 0
            self._len[done] = 0
            obs = self.env.reset(mask=done)        # masked reset kernel: new first observations in place
        return self._encode(obs), reward, done

    def finished_returns(self) -> torch.Tensor:
        if not self.episode_returns:
            return torch.zeros(0, device=self.device)
        return torch.cat(self.episode_returns)


class MLPPolicy(nn.Module):
    """The reference's policy shape: fully-connected 350-250 ReLU (benchmarks/benchmark.py:51, :64), tanh-squashed
    Gaussian action in [-1, 1]."""

    def __init__(self, obs_dim: int, hidden=(350, 250), log_std: float = -0.5):
        super().__init__()
        layers, d = [], obs_dim
        for h in hidden:
            layers += [nn.Linear(d, h), nn.ReLU()]
            d = h
        layers.append(nn.Linear(d, 1))
        self.net = nn.Sequential(*layers)
        self.log_std = nn.Parameter(torch.full((1,), log_std))

    @torch.no_grad()
    def act(self, obs: torch.Tensor, deterministic: bool = False) -> torch.Tensor:
        mean = self.net(obs).squeeze(-1)
        if not deterministic:
            mean = mean + torch.randn_like(mean) * self.log_std.exp()
        return torch.tanh(mean)


def rollout(vec: VecJitterbugEnv, policy: Optional[MLPPolicy], n_steps: int, group=None) -> Dict[str, float]:
    """Run ``n_steps`` control steps with ``policy`` (``None`` = uniform random actions).  Returns throughput
    (this rank) and the global episode statistics."""
    obs = vec.reset()
    dev = vec.device
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    total_reward = torch.zeros((), device=dev, dtype=torch.float64)
    for _ in range(n_steps):
        if policy is None:
            act = torch.rand(vec.num_envs, device=dev) * 2 - 1
        else:
            act = policy.act(obs)
        obs, rew, done = vec.step(act)
        total_reward += rew.double().sum()
    torch.cuda.synchronize(dev)
    dt = time.perf_counter() - t0
    stats = gather_episode_stats(vec.finished_returns(), group) if vec.episode_count else {"count": 0.0}
    return {"steps_per_s": vec.num_envs * n_steps / dt, "seconds": dt,
            "mean_step_reward": float(total_reward) / (vec.num_envs * n_steps), "episodes": stats}
