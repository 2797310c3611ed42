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
from jitterbug_b200.rollout import MLPPolicy, VecJitterbugEnv, rollout  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=262144, help="global number of envs")
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--task", default="move_to_pose")
    ap.add_argument("--time-limit", type=float, default=10.0)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    env = make_sharded_env(args.task, args.envs, seed=0, time_limit=args.time_limit)
    vec = VecJitterbugEnv(env)
    policy = MLPPolicy(env.obs_dim).to(env.device)
    r_env = rollout(vec, None, args.steps)
    r_pol = rollout(vec, policy, args.steps)
    agg = torch.tensor([r_env["steps_per_s"], r_pol["steps_per_s"]], device=env.device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(agg)
    if int(os.environ.get("RANK", "0")) == 0:
        print(json.dumps({"task": args.task, "global_envs": args.envs, "n_gpus": world, "control_steps": args.steps,
                          "env_only_steps_per_s": float(agg[0]), "env_plus_policy_steps_per_s": float(agg[1]),
                          "episodes": r_pol["episodes"]}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
