#!/usr/bin/env python
"""Config 5 of BASELINE.json: rollout loop with a torch MLP policy on device over env-sharded batched Jitterbugs.

    python examples/rollout_bench.py --envs 262144 --steps 200                       # 1 GPU
    torchrun --nproc-per-node 8 examples/rollout_bench.py --envs 2097152 --steps 200  # 8 GPUs, envs = global total
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jitterbug_b200.distributed import make_sharded_env  # noqa: E402
from jitterbug_b200.rollout import ActorCritic, RolloutBuffer, VecJitterbugEnv, collect, rollout  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=262144, help="global number of envs")
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--task", default="move_to_pose")
    ap.add_argument("--time-limit", type=float, default=10.0)
    ap.add_argument("--burn-in", type=int, default=150, help="untimed control steps before measuring (steady state)")
    ap.add_argument("--horizon", type=int, default=32, help="PPO collection horizon (control steps per buffer)")
    ap.add_argument("--collections", type=int, default=3)
    ap.add_argument("--no-graph", action="store_true")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    env = make_sharded_env(args.task, args.envs, seed=0, time_limit=args.time_limit)
    vec = VecJitterbugEnv(env)
    torch.backends.cuda.matmul.allow_tf32 = True          # the policy's small GEMMs on the tensor cores (TF32)
    policy = ActorCritic(env.obs_dim).to(env.device)
    # burn-in: from reset the Jitterbugs are airborne / spinning up, which is cheaper than the steady state
    obs = vec.reset()
    for _ in range(args.burn_in):
        obs, _, _, _ = vec.step(torch.rand(env.num_envs, device=env.device) * 2 - 1)
    r_env = rollout(vec, None, args.steps)
    r_pol = rollout(vec, policy, args.steps)
    # PPO collection: action + log-prob + value per step into the on-device buffer, GAE at the end, CUDA-graph replay
    buf = RolloutBuffer(args.horizon, env.num_envs, vec.obs_dim, env.device)
    cache = {}
    obs = collect(vec, policy, buf, vec.reset(), cuda_graph=not args.no_graph, _graph_cache=cache)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.collections):
        obs = collect(vec, policy, buf, obs, cuda_graph=not args.no_graph, _graph_cache=cache)
    e1.record()
    torch.cuda.synchronize()
    ppo = env.num_envs * args.horizon * args.collections / (e0.elapsed_time(e1) * 1e-3)
    agg = torch.tensor([r_env["steps_per_s"], r_pol["steps_per_s"], ppo], device=env.device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(agg)
    if int(os.environ.get("RANK", "0")) == 0:
        print(json.dumps({"task": args.task, "global_envs": args.envs, "n_gpus": world, "control_steps": args.steps,
                          "burn_in_steps": args.burn_in, "timing": "CUDA events, summed over ranks",
                          "env_only_steps_per_s": float(agg[0]), "env_plus_policy_steps_per_s": float(agg[1]),
                          "ppo_collection_steps_per_s": float(agg[2]), "ppo_horizon": args.horizon,
                          "cuda_graph": not args.no_graph, "episodes": r_pol["episodes"]}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
