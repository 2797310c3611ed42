"""Self-clearance monitor over one full episode of a large batch: python tests/scripts/gpu_clearance_run.py [N] [steps]
(uniform random actions and, second, the heuristic policy; writes gpurun_out/r02_self_clearance.json)."""
import json, os, sys
import torch
sys.path.insert(0, '.')
from jitterbug_b200.batched_env import BatchedJitterbugEnv
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
T = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
out = {"envs": N, "control_steps": T, "monitor_period": 8}
for policy in ("random", "heuristic"):
    env = BatchedJitterbugEnv("move_from_origin", num_envs=N, seed=17)
    obs = env.reset()
    g = torch.Generator(device='cuda').manual_seed(5)
    for k in range(T):
        a = torch.rand(N, device='cuda', generator=g) * 2 - 1 if policy == "random" else env.heuristic_actions(obs)
        obs, _, _ = env.step(a)
    res = env.self_clearance(per_env=True)
    cl = res.pop("clearance")
    res["final_state_min_clearance_m"] = float(cl.min())
    res["final_state_p001_clearance_m"] = float(torch.quantile(cl[:1 << 20].float(), 0.001))
    res["counters"] = env.counters()
    out[policy] = res
    print(policy, res, flush=True)
    del env
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/r02_self_clearance.json", "w"), indent=1)
