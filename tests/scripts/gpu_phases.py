"""In-kernel phase breakdown (profiling build -DJBQ_PHASE_TIMERS=1, see jb_quad.cuh JBQ_TICK): share of clock64 cycles
per phase of the sub-step in steady state.  JB_LIB=<variant> python tests/scripts/gpu_phases.py [N] [burn_in]"""
import sys
import torch
sys.path.insert(0, '.')
from jitterbug_b200.batched_env import BatchedJitterbugEnv
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
B = int(sys.argv[2]) if len(sys.argv) > 2 else 150
env = BatchedJitterbugEnv("move_from_origin", num_envs=N, seed=0)
env.reset()
g = torch.Generator(device='cuda').manual_seed(0)
acts = [torch.rand(N, device='cuda', generator=g) * 2 - 1 for _ in range(8)]
for i in range(B):
    env.step(acts[i % 8])
p0 = env.phase_cycles()
for i in range(10):
    env.step(acts[i % 8])
p1 = env.phase_cycles()
d = [b - a for a, b in zip(p0, p1)]
tot = sum(d)
names = ["-", "leg kin+col (+barrier 1)", "root/rotor", "extra geoms", "smooth force", "newton entry", "newton loop (+barrier 6)", "euler"]
print(", ".join(f"{n}={100 * c / tot:.1f}%" for n, c in zip(names, d) if c))
