import sys, torch
sys.path.insert(0, '.')
from jitterbug_b200.batched_env import BatchedJitterbugEnv
N = int(sys.argv[1])
env = BatchedJitterbugEnv("move_from_origin", num_envs=N, seed=0)
env.reset()
g = torch.Generator(device='cuda').manual_seed(0)
acts = [torch.rand(N, device='cuda', generator=g)*2-1 for _ in range(8)]
for i in range(40): env.step(acts[i%8])
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 10
ev0.record()
for i in range(K): env.step(acts[i%8])
ev1.record(); torch.cuda.synchronize()
ms = ev0.elapsed_time(ev1)/K
c = env.counters()
print(f"N={N} ms/step={ms:.3f} steps/s={N/ms*1e3:.3e} newton/substep={c['newton_iterations']/c['substeps']:.3f} ncon={c['contacts']/c['substeps']:.3f} ovf={c['contact_overflows']} div={c['diverged_resets']}")
pc = env.phase_cycles()
if sum(pc):
    tot = sum(pc)
    names = ["-", "leg kin+col", "root/rotor", "extra geoms", "factorM+solve", "warmstart sel", "newton loop", "euler"]
    print("phases: " + ", ".join(f"{n}={100*c/tot:.1f}%" for n, c in zip(names, pc) if c))
