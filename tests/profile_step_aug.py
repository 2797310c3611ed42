"""Profiling driver for the per-env-model instantiation of the step kernel (config 4), see profiles/README.md."""
import sys
import torch
sys.path.insert(0, '.')
from jitterbug_b200 import augmented_jitterbug as aug

N = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
warm = int(sys.argv[2]) if len(sys.argv) > 2 else 30
env = aug.make_batched_env("move_from_origin", num_envs=N, seed=0)
env.reset()
g = torch.Generator(device='cuda').manual_seed(0)
acts = [torch.rand(N, device='cuda', generator=g) * 2 - 1 for _ in range(4)]
for i in range(warm + 3):
    env.step(acts[i % 4])
torch.cuda.synchronize()
print("done", env.counters())
