"""Round-2 experiment: how persistent is an env's load from one control step to the next, and what does a load-sorted
slot order buy?  python tests/scripts/gpu_sort_experiment.py [N] [burn_in] [K]"""
import json
import sys
import torch
sys.path.insert(0, '.')
from jitterbug_b200.batched_env import BatchedJitterbugEnv

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
BURN = int(sys.argv[2]) if len(sys.argv) > 2 else 150
K = int(sys.argv[3]) if len(sys.argv) > 3 else 8
dev = 'cuda'
out = {"N": N, "burn_in": BURN, "K": K}

# ---- 0. results do not depend on the slot order (small batch, 40 steps, bitwise)
n0 = 8192
ea = BatchedJitterbugEnv("move_from_origin", num_envs=n0, seed=3); ea.set_slot_order("identity")
eb = BatchedJitterbugEnv("move_from_origin", num_envs=n0, seed=3); eb.set_slot_order("auto")
ec = BatchedJitterbugEnv("move_from_origin", num_envs=n0, seed=3)
ec.set_slot_order(torch.randperm(n0, device=dev).int())
for e in (ea, eb, ec): e.reset()
g = torch.Generator(device=dev).manual_seed(1)
same = True
for i in range(60):
    a = torch.rand(n0, device=dev, generator=g) * 2 - 1
    ra, rb, rc = ea.step(a), eb.step(a), ec.step(a)
    for x, y, z in zip(ra, rb, rc):
        same &= bool(torch.equal(x, y)) and bool(torch.equal(x, z))
    if i % 7 == 0:
        ec.set_slot_order(torch.randperm(n0, device=dev).int())
sa, sb, sc = ea.get_state(), eb.get_state(), ec.get_state()
for x, y, z in zip(sa, sb, sc):
    same &= bool(torch.equal(x, y)) and bool(torch.equal(x, z))
p = eb.get_slot_order().long()
out["order_invariant_bitwise"] = same
out["auto_perm_is_permutation"] = bool(torch.equal(torch.sort(p).values, torch.arange(n0, device=dev)))
print(json.dumps(out), flush=True)
del ea, eb, ec

env = BatchedJitterbugEnv("move_from_origin", num_envs=N, seed=0)
env.set_slot_order("identity")
env.reset()
acts = [torch.rand(N, device=dev, generator=g) * 2 - 1 for _ in range(8)]
for i in range(BURN):
    env.step(acts[i % 8])
torch.cuda.synchronize()


def timed_steps(k, after_step=None):
    tot = 0.0
    for i in range(k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        env.step(acts[i % 8])
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
        if after_step:
            after_step()
    return tot / k


# ---- 1. persistence of the load record
loads = []
for i in range(6):
    env.step(acts[i % 8])
    loads.append(env.load().clone())
L = torch.stack(loads).long()          # [6, N, 4]
heavy3 = L[:, :, 1] > 0
heavy2 = L[:, :, 2] > 0
extra = (L[:, :, 3] > 50)
def persist(m):
    return [float((m[t] & m[t + 1]).sum()) / max(float(m[t].sum()), 1) for t in range(m.shape[0] - 1)]
out["frac_heavy3"] = float(heavy3.float().mean()); out["persist_heavy3"] = persist(heavy3)
out["frac_heavy2"] = float(heavy2.float().mean()); out["persist_heavy2"] = persist(heavy2)
out["frac_extra_newton"] = float(extra.float().mean()); out["persist_extra_newton"] = persist(extra)
out["frac_airborne"] = float((L[:, :, 0] == 0).float().mean()); out["persist_airborne"] = persist(L[:, :, 0] == 0)
x = L[:-1, :, 0].float().flatten(); y = L[1:, :, 0].float().flatten()
out["corr_ldn_t_t1"] = float(torch.corrcoef(torch.stack([x, y]))[0, 1])
x = L[:-1, :, 3].float().flatten(); y = L[1:, :, 3].float().flatten()
out["corr_iters_t_t1"] = float(torch.corrcoef(torch.stack([x, y]))[0, 1])
out["hist_ld3"] = torch.bincount(L[-1, :, 1].clamp(max=50), minlength=51).tolist()
out["hist_ld2"] = torch.bincount(L[-1, :, 2].clamp(max=50), minlength=51).tolist()
out["hist_iters"] = torch.bincount((L[-1, :, 3] - 0).clamp(min=0, max=80), minlength=81).tolist()
print(json.dumps(out), flush=True)

# ---- 2. slot orders
res = {}
env.set_slot_order("identity")
res["identity"] = timed_steps(K)
env.set_slot_order("auto")
env.step(acts[0])
res["auto"] = timed_steps(K)


def explicit(keyfn):
    def f():
        ld = env.load().long()
        key = keyfn(ld)
        env.set_slot_order(torch.sort(key, stable=True).indices.int())
    return f


cands = {
    "random_perm": lambda ld: torch.randint(0, 1 << 30, (N,), device=dev),
    "heavy3_binary": lambda ld: (ld[:, 1] > 0).long(),
    "heavy2_binary": lambda ld: (ld[:, 2] > 0).long(),
    "ld_n_full": lambda ld: ld[:, 0],
    "ld3_ld2": lambda ld: ld[:, 1] * 64 + ld[:, 2],
    "iters": lambda ld: ld[:, 3],
    "ld3_ld2_iters": lambda ld: (ld[:, 1] * 64 + ld[:, 2]) * 256 + ld[:, 3].clamp(max=255),
    "ld3_iters_ld2": lambda ld: (ld[:, 1] * 256 + ld[:, 3].clamp(max=255)) * 64 + ld[:, 2],
    "iters_ld3_ld2": lambda ld: (ld[:, 3].clamp(max=255) * 64 + ld[:, 1]) * 64 + ld[:, 2],
    "ldn_iters": lambda ld: ld[:, 0] * 256 + ld[:, 3].clamp(max=255),
}
for name, fn in cands.items():
    f = explicit(fn)
    env.step(acts[0]); f()
    res[name] = timed_steps(K, after_step=f)
    print(name, res[name], flush=True)
env.set_slot_order("identity")
res["identity_again"] = timed_steps(K)
out["ms_per_step"] = res
out["steps_per_s"] = {k: N / v * 1e3 for k, v in res.items()}
print(json.dumps(out), flush=True)
import os
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/sort_experiment.json", "w"), indent=1)
