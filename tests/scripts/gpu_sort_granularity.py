"""Round-2 experiment: re-sorting the slot order every S sub-steps (S = sub-steps per launch).
python tests/scripts/gpu_sort_granularity.py [N]"""
import json, sys, os
import torch
sys.path.insert(0, '.')
from jitterbug_b200.batched_env import BatchedJitterbugEnv
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
dev = 'cuda'
g = torch.Generator(device=dev).manual_seed(1)
acts = [torch.rand(N, device=dev, generator=g) * 2 - 1 for _ in range(8)]
out = {}
SLIST = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [50, 25, 10, 5, 2]
for S in SLIST:
    rep = 50 // S
    env = BatchedJitterbugEnv("move_from_origin", num_envs=N, seed=0, control_timestep=0.0002 * S, time_limit=float("inf"))
    env.reset()
    for order in ("identity", "auto"):
        env.set_slot_order(order)
        for i in range(100 if order == "identity" else 10):      # burn-in: 100 control steps of 50 sub-steps
            for r in range(rep):
                env.step(acts[i % 8])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        K = 6
        e0.record()
        for i in range(K):
            for r in range(rep):
                env.step(acts[i % 8])
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / K
        ld = env.load().long()
        out[f"S{S}_{order}"] = {"ms_per_50_substeps": ms, "steps_per_s": N / ms * 1e3,
                                "frac_heavy3": float((ld[:, 1] > 0).float().mean())}
        print(S, order, out[f"S{S}_{order}"], flush=True)
    del env
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/sort_granularity.json", "w"), indent=1)
