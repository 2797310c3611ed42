#!/usr/bin/env python
"""A complete on-device PPO loop on the batched Jitterbug: what `benchmarks/benchmark.py --alg ppo2` does with
stable-baselines + SubprocVecEnv in the reference, here with the collection phase of `jitterbug_b200.rollout`
(VecEnv on CUDA tensors, CUDA-graph replay of policy + env step, GAE on device) and a plain clipped-surrogate update.

    python examples/ppo_train.py --task face_direction --iters 200
    python examples/ppo_train.py --task move_from_origin --iters 100

It is a CALLER of the hot path, not part of it (the reference's RL training scripts are out of scope, DESIGN.md section
1): it exists to show the simulator doing its job -- a learner consuming ~10^7 control steps per second -- and to put a
learning curve next to the paper's (fig-rl-perf: PPO2 on face_direction climbs from ~ 230 to ~ 900 return).
Prints one JSON line whenever a batch of episodes finished, and a summary line at the end.
"""
import argparse
import json
import os
import sys
import time

import torch
from torch import nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jitterbug_b200 import BatchedJitterbugEnv  # noqa: E402
from jitterbug_b200.rollout import ActorCritic, RolloutBuffer, VecJitterbugEnv, collect  # noqa: E402


class Scale(nn.Module):
    def __init__(self, s):
        super().__init__()
        self.s = s

    def forward(self, x):
        return x * self.s


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--task", default="face_direction")
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--horizon", type=int, default=50)
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--epochs", type=int, default=4)
    ap.add_argument("--minibatches", type=int, default=8)
    ap.add_argument("--lr", type=float, default=3e-4)
    ap.add_argument("--gamma", type=float, default=0.99)
    ap.add_argument("--lam", type=float, default=0.95)
    ap.add_argument("--clip", type=float, default=0.2)
    ap.add_argument("--ent-coef", type=float, default=0.0)
    ap.add_argument("--log-std", type=float, default=-0.5)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--no-graph", action="store_true")
    args = ap.parse_args()

    torch.manual_seed(args.seed)
    torch.backends.cuda.matmul.allow_tf32 = True
    env = BatchedJitterbugEnv(args.task, num_envs=args.envs, seed=args.seed)
    vec = VecJitterbugEnv(env)
    dev = env.device
    policy = ActorCritic(vec.obs_dim, log_std=args.log_std).to(dev)
    policy.vf = nn.Sequential(policy.vf, Scale(1.0 / (1.0 - args.gamma))).to(dev)     # the critic's output in units of 1/(1-gamma)
    opt = torch.optim.Adam(policy.parameters(), lr=args.lr, eps=1e-5, capturable=False)
    buf = RolloutBuffer(args.horizon, args.envs, vec.obs_dim, dev)
    cache = {}
    obs = vec.reset()
    n, mb = args.horizon * args.envs, args.horizon * args.envs // args.minibatches
    v_scale = 1.0 - args.gamma
    t0 = time.perf_counter()
    t_collect = t_update = 0.0
    count0, sum0 = 0, 0.0
    curve = []
    for it in range(args.iters):
        tc = time.perf_counter()
        obs = collect(vec, policy, buf, obs, gamma=args.gamma, lam=args.lam, cuda_graph=not args.no_graph, _graph_cache=cache)
        torch.cuda.synchronize()
        t_collect += time.perf_counter() - tc
        tu = time.perf_counter()
        b_obs, b_act = buf.obs.reshape(n, -1), buf.actions.reshape(n)
        b_lp, b_adv, b_ret = buf.log_probs.reshape(n), buf.advantages.reshape(n), buf.returns.reshape(n)
        for _ in range(args.epochs):
            perm = torch.randperm(n, device=dev)
            for k in range(args.minibatches):
                idx = perm[k * mb:(k + 1) * mb]
                lp, v, ent = policy.evaluate(b_obs[idx], b_act[idx])
                adv = b_adv[idx]
                adv = (adv - adv.mean()) / (adv.std() + 1e-8)
                ratio = torch.exp(lp - b_lp[idx])
                pg = -torch.min(ratio * adv, ratio.clamp(1 - args.clip, 1 + args.clip) * adv).mean()
                vf = 0.5 * ((v - b_ret[idx]) * v_scale).pow(2).mean()
                loss = pg + 0.5 * vf - args.ent_coef * ent.mean()
                opt.zero_grad(set_to_none=True)
                loss.backward()
                nn.utils.clip_grad_norm_(policy.parameters(), 0.5)
                opt.step()
        torch.cuda.synchronize()
        t_update += time.perf_counter() - tu
        st = vec.episode_statistics()                         # Monitor sums over all finished episodes so far
        count, total = int(st["count"]), st["mean"] * st["count"]
        if count > count0:
            rec = {"iter": it + 1, "env_steps": (it + 1) * n, "wall_s": round(time.perf_counter() - t0, 2),
                   "episodes": count - count0, "mean_return": round((total - sum0) / (count - count0), 1),
                   "policy_std": round(float(policy.log_std.detach().exp()), 3)}
            print(json.dumps(rec), flush=True)
            curve.append(rec)
            count0, sum0 = count, total
    wall = time.perf_counter() - t0
    print(json.dumps({"task": args.task, "envs": args.envs, "horizon": args.horizon, "iters": args.iters,
                      "env_steps": args.iters * n, "wall_s": round(wall, 2), "collect_s": round(t_collect, 2),
                      "update_s": round(t_update, 2), "collect_steps_per_s": round(args.iters * n / t_collect),
                      "first_mean_return": curve[0]["mean_return"] if curve else None,
                      "best_mean_return": max(c["mean_return"] for c in curve) if curve else None,
                      "final_mean_return": curve[-1]["mean_return"] if curve else None,
                      "diverged_resets": env.counters()["diverged_resets"]}))


if __name__ == "__main__":
    main()
