"""Steady-state timing of the step: python tests/scripts/gpu_time.py N burn_in K [task]  (env vars select tuning variants)."""
import os, sys, json
import torch
sys.path.insert(0, '.')
from jitterbug_b200.batched_env import BatchedJitterbugEnv
N = int(sys.argv[1]); BURN = int(sys.argv[2]); K = int(sys.argv[3])
task = sys.argv[4] if len(sys.argv) > 4 else "move_from_origin"
if os.environ.get("JB_AUG"):              # one augmented model per env (the per-env-model kernel, config 4)
    from jitterbug_b200 import augmented_jitterbug as aug
    env = aug.make_batched_env(task, num_envs=N, seed=0)
else:
    env = BatchedJitterbugEnv(task, num_envs=N, seed=0)
if os.environ.get("JB_ORDER"):
    env.set_slot_order(os.environ["JB_ORDER"])
env.reset()
g = torch.Generator(device='cuda').manual_seed(0)
acts = [torch.rand(N, device='cuda', generator=g) * 2 - 1 for _ in range(8)]
for i in range(BURN):
    env.step(acts[i % 8])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(K):
    env.step(acts[i % 8])
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
c = env.counters()
tag = {k: v for k, v in os.environ.items() if k.startswith("JBSIM_") or k in ("JB_ORDER", "JB_LIB", "JB_AUG")}
print(json.dumps({"tag": tag, "N": N, "ms_per_step": round(ms, 3), "steps_per_s": round(N / ms * 1e3),
                  "newton_per_substep": round(c['newton_iterations'] / c['substeps'], 4),
                  "ncon_per_substep": round(c['contacts'] / c['substeps'], 4), "overflows": c['contact_overflows'],
                  "diverged": c['diverged_resets']}), flush=True)
