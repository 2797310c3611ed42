"""On-device RL data collection over the batched simulator (config 5 of BASELINE.json).

Replaces the reference's ``SubprocVecEnv`` + ``Monitor`` plumbing around ``JitterbugGymEnv`` (benchmarks/benchmark.py:
146-185: 8 worker processes, pickled numpy over pipes, a CSV monitor per env) and the rollout half of its PPO2 / SAC
agents: observations, actions, log-probabilities, values, rewards, episode accounting and advantage estimation all stay
on the GPU, and nothing in the collection loop synchronises with the host.

* :class:`VecJitterbugEnv` -- the stable-baselines ``VecEnv`` contract (``num_envs``, ``reset``, ``step_async`` /
  ``step_wait`` / ``step`` returning ``(obs, rewards, dones, infos)``, auto-reset with the first observation of the next
  episode, ``Monitor``-style episode return / length in ``infos``), on torch CUDA tensors.
* :class:`ActorCritic` -- the reference's 350-250 ReLU network (benchmark.py:51, :64) with a tanh-squashed Gaussian
  head: action, log-probability and value per env.
* :class:`RolloutBuffer` + :func:`collect` -- fixed-horizon PPO collection into preallocated device tensors, GAE(lambda)
  on the device; ``collect(..., cuda_graph=True)`` captures one policy + env step in a CUDA graph and replays it.

Envs shard across GPUs by env index (one process per GPU); the only collective is the end-of-rollout statistics
all-reduce (:func:`jitterbug_b200.distributed.gather_episode_stats`).
"""

from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn as nn

from .batched_env import STEP_FIRST, STEP_LAST, BatchedJitterbugEnv
from .distributed import gather_episode_stats


class EpisodeInfos:
    """``infos`` of a vectorised step, kept on the device: ``done`` mask plus ``Monitor``-style episode return (``r``) and
    length (``l``) of the envs whose episode just ended.  ``to_list()`` materialises the stable-baselines list of dicts
    (one host synchronisation) for consumers that want it."""

    __slots__ = ("done", "r", "l")

    def __init__(self, done, r, l):  # noqa: E741 - Monitor's key names
        self.done, self.r, self.l = done, r, l

    def to_list(self):
        done, r, l = self.done.cpu().numpy(), self.r.cpu().numpy(), self.l.cpu().numpy()  # noqa: E741
        return [({"episode": {"r": float(r[i]), "l": int(l[i])}} if done[i] else {}) for i in range(len(done))]


class VecJitterbugEnv:
    """``VecEnv`` over a :class:`BatchedJitterbugEnv`.

    ``step`` returns ``(obs, rewards, dones, infos)``.  When an env finishes its episode (time limit -> ``LAST``, or a
    diverged state the simulator restarted) ``dones`` is set, the observation returned for it is already the first
    observation of its next episode (stable-baselines convention) and ``infos`` carries the finished episode's return
    and length.  No host synchronisation: the masked reset kernel is enqueued after every step with the device-side
    ``done`` mask (CTAs without a finished env return immediately).
    """

    def __init__(self, env: BatchedJitterbugEnv, encoder=None):
        """``encoder``: optional :class:`jitterbug_b200.encoder.ObservationEncoder` applied to every observation
        returned (the reference's ``encode_obs`` hook, jitterbug.py:927-993); ``obs_dim`` is then its output width."""
        self.env = env
        self.num_envs = env.num_envs
        self.device = env.device
        self.encoder = encoder
        self.obs_dim = env.obs_dim if encoder is None else encoder.out_dim
        dev = self.device
        self._ret = torch.zeros(self.num_envs, device=dev)
        self._len = torch.zeros(self.num_envs, dtype=torch.int32, device=dev)
        # Monitor: running sums over finished episodes (device scalars; read them with episode_statistics())
        self._ep_count = torch.zeros((), dtype=torch.int64, device=dev)
        self._ep_sum = torch.zeros((), dtype=torch.float64, device=dev)
        self._ep_sumsq = torch.zeros((), dtype=torch.float64, device=dev)
        self._ep_min = torch.full((), float("inf"), dtype=torch.float64, device=dev)
        self._ep_max = torch.full((), float("-inf"), dtype=torch.float64, device=dev)
        self._last_ret = torch.zeros(self.num_envs, device=dev)
        self._last_len = torch.zeros(self.num_envs, dtype=torch.int32, device=dev)
        self._pending = None

    # ---- VecEnv
    def reset(self) -> torch.Tensor:
        self._ret.zero_()
        self._len.zero_()
        return self._encode(self.env.reset())

    def _encode(self, obs):
        return obs if self.encoder is None else self.encoder(obs)

    def step_async(self, actions: torch.Tensor) -> None:
        self._pending = actions

    def step_wait(self):
        actions, self._pending = self._pending, None
        obs, reward, step_type = self.env.step(actions)
        timeout = step_type == STEP_LAST
        done = timeout | (step_type == STEP_FIRST)        # FIRST out of step(): the simulator restarted a diverged env
        self._ret += reward
        self._len += 1
        # Monitor bookkeeping, all masked arithmetic (no data-dependent host control flow)
        fin = torch.where(done, self._ret, torch.zeros_like(self._ret)).double()
        self._ep_count += done.sum()
        self._ep_sum += fin.sum()
        self._ep_sumsq += (fin * fin).sum()
        inf = torch.full_like(self._ret, float("inf"))
        self._ep_min.copy_(torch.minimum(self._ep_min, torch.where(done, self._ret, inf).min().double()))
        self._ep_max.copy_(torch.maximum(self._ep_max, torch.where(done, self._ret, -inf).max().double()))
        # in-place updates throughout: a CUDA-graph replay of this method must keep writing the same buffers
        self._last_ret.copy_(torch.where(done, self._ret, self._last_ret))
        self._last_len.copy_(torch.where(done, self._len, self._last_len))
        infos = EpisodeInfos(done, self._last_ret, self._last_len)
        self._ret.masked_fill_(done, 0.0)
        self._len.masked_fill_(done, 0)
        # new episodes for the envs at their time limit: first observations written in place
        obs = self.env.reset(mask=timeout)
        return self._encode(obs), reward, done, infos

    def step(self, actions: torch.Tensor):
        self.step_async(actions)
        return self.step_wait()

    def close(self):
        self.env.close()

    # ---- Monitor
    def episode_statistics(self, group=None) -> Dict[str, float]:
        """Finished-episode statistics so far, all-reduced over ``group`` when torch.distributed is initialised."""
        return gather_episode_stats(None, group, moments=(self._ep_count, self._ep_sum, self._ep_sumsq, self._ep_min,
                                                          self._ep_max))

    @property
    def episode_count(self) -> int:
        return int(self._ep_count)


class ActorCritic(nn.Module):
    """The reference's policy shape -- fully-connected 350-250 ReLU (benchmarks/benchmark.py:51, :64) -- with a Gaussian
    action head squashed by tanh into the actuator's [-1, 1] range and a value head on a second trunk of the same shape."""

    def __init__(self, obs_dim: int, hidden=(350, 250), log_std: float = -0.5):
        super().__init__()

        def trunk():
            layers, d = [], obs_dim
            for h in hidden:
                layers += [nn.Linear(d, h), nn.ReLU()]
                d = h
            return nn.Sequential(*layers, nn.Linear(d, 1))
        self.pi = trunk()
        self.vf = trunk()
        self.log_std = nn.Parameter(torch.full((1,), log_std))

    def distribution(self, obs):
        return self.pi(obs).squeeze(-1), self.log_std.expand(obs.shape[0])

    @staticmethod
    def _log_prob(u, mean, log_std):
        """log-density of a = tanh(u), u ~ N(mean, exp(log_std)^2)."""
        z = (u - mean) * torch.exp(-log_std)
        gauss = -0.5 * z * z - log_std - 0.5 * math.log(2 * math.pi)
        return gauss - 2.0 * (math.log(2.0) - u - nn.functional.softplus(-2.0 * u))      # - log(1 - tanh(u)^2)

    @torch.no_grad()
    def act(self, obs: torch.Tensor, deterministic: bool = False):
        """(action[N], log_prob[N], value[N])."""
        mean, log_std = self.distribution(obs)
        u = mean if deterministic else mean + torch.randn_like(mean) * log_std.exp()
        return torch.tanh(u), self._log_prob(u, mean, log_std), self.vf(obs).squeeze(-1)

    def evaluate(self, obs, actions):
        """(log_prob, value, entropy estimate) of stored actions, with gradients (the learner's side)."""
        mean, log_std = self.distribution(obs)
        u = torch.atanh(actions.clamp(-1 + 1e-6, 1 - 1e-6))
        return self._log_prob(u, mean, log_std), self.vf(obs).squeeze(-1), log_std + 0.5 * math.log(2 * math.pi * math.e)


MLPPolicy = ActorCritic     # round-1 name


class RolloutBuffer:
    """Fixed-horizon on-device storage of a PPO collection phase: [T, N, ...] tensors, written in place."""

    def __init__(self, horizon: int, num_envs: int, obs_dim: int, device):
        T, N = horizon, num_envs
        f = dict(device=device, dtype=torch.float32)
        self.obs = torch.zeros(T, N, obs_dim, **f)
        self.actions = torch.zeros(T, N, **f)
        self.log_probs = torch.zeros(T, N, **f)
        self.values = torch.zeros(T, N, **f)
        self.rewards = torch.zeros(T, N, **f)
        self.dones = torch.zeros(T, N, device=device, dtype=torch.bool)
        self.advantages = torch.zeros(T, N, **f)
        self.returns = torch.zeros(T, N, **f)
        self.horizon, self.num_envs = T, N

    def compute_gae(self, last_value: torch.Tensor, gamma: float = 0.99, lam: float = 0.95) -> None:
        """GAE(lambda) backwards over the horizon, all envs at once.  ``dones[t]`` marks that the episode ended WITH
        transition t: nothing is bootstrapped across it (the time-limit truncation is treated as termination, as
        stable-baselines' PPO2 does with the ``dones`` of a ``VecEnv``)."""
        adv = torch.zeros_like(last_value)
        next_value = last_value
        for t in range(self.horizon - 1, -1, -1):
            nonterminal = (~self.dones[t]).float()
            delta = self.rewards[t] + gamma * next_value * nonterminal - self.values[t]
            adv = delta + gamma * lam * nonterminal * adv
            self.advantages[t] = adv
            next_value = self.values[t]
        self.returns = self.advantages + self.values


def collect(vec: VecJitterbugEnv, policy: Optional[ActorCritic], buf: RolloutBuffer, obs: torch.Tensor, gamma: float = 0.99,
            lam: float = 0.95, cuda_graph: bool = False, _graph_cache: Optional[dict] = None) -> torch.Tensor:
    """Fill ``buf`` with ``buf.horizon`` control steps starting from observation ``obs``; returns the observation to
    continue from.  ``policy=None`` draws uniform random actions (log-prob of the uniform density, value 0).

    ``cuda_graph=True``: the per-step work (policy forward, env step with its segment launches and slot-order sorts,
    the VecEnv bookkeeping and the masked reset) is captured once and replayed ``horizon`` times; pass a dict as
    ``_graph_cache`` to reuse the capture across calls."""
    T = buf.horizon
    dev = vec.device

    def one_step(o):
        if policy is None:
            a = torch.rand(vec.num_envs, device=dev) * 2 - 1
            lp, v = torch.full_like(a, -math.log(2.0)), torch.zeros_like(a)
        else:
            a, lp, v = policy.act(o)
        o2, r, d, _ = vec.step(a)
        return a, lp, v, r, d, o2

    if not cuda_graph:
        for t in range(T):
            buf.obs[t] = obs
            a, lp, v, r, d, obs = one_step(obs)
            buf.actions[t], buf.log_probs[t], buf.values[t], buf.rewards[t], buf.dones[t] = a, lp, v, r, d
        obs = obs.clone()                                          # the env reuses its observation tensor
    else:
        cache = _graph_cache if _graph_cache is not None else {}
        if "graph" not in cache:
            static_obs = obs.clone()
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):                       # warm-up outside capture (lazy initialisations)
                for _ in range(2):
                    one_step(static_obs)
            torch.cuda.current_stream(dev).wait_stream(side)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                outs = one_step(static_obs)
                static_obs.copy_(outs[5])
            cache.update(graph=g, obs=static_obs, outs=outs)
            cache["obs"].copy_(obs)
        g, static_obs, outs = cache["graph"], cache["obs"], cache["outs"]
        static_obs.copy_(obs)
        for t in range(T):
            buf.obs[t] = static_obs
            g.replay()
            buf.actions[t], buf.log_probs[t], buf.values[t], buf.rewards[t], buf.dones[t] = outs[:5]
        obs = static_obs.clone()
    if policy is None:
        last_value = torch.zeros(vec.num_envs, device=dev)
    else:
        with torch.no_grad():
            last_value = policy.vf(obs).squeeze(-1)
    buf.compute_gae(last_value, gamma, lam)
    return obs


def rollout(vec: VecJitterbugEnv, policy: Optional[ActorCritic], n_steps: int, group=None) -> Dict[str, float]:
    """Run ``n_steps`` control steps with ``policy`` (``None`` = uniform random actions), CUDA-event timed.  Returns
    throughput (this rank) and the global episode statistics."""
    obs = vec.reset()
    dev = vec.device
    total_reward = torch.zeros((), device=dev, dtype=torch.float64)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(dev)
    e0.record()
    for _ in range(n_steps):
        if policy is None:
            act = torch.rand(vec.num_envs, device=dev) * 2 - 1
        else:
            act = policy.act(obs)[0]
        obs, rew, done, _ = vec.step(act)
        total_reward += rew.double().sum()
    e1.record()
    torch.cuda.synchronize(dev)
    dt = e0.elapsed_time(e1) * 1e-3
    return {"steps_per_s": vec.num_envs * n_steps / dt, "seconds": dt,
            "mean_step_reward": float(total_reward) / (vec.num_envs * n_steps), "episodes": vec.episode_statistics(group)}
