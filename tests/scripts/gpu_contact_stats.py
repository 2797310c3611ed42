"""Tuning helper (GPU): distribution and persistence of contact load per env in the bench workload."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from jitterbug_b200 import BatchedJitterbugEnv  # noqa: E402

E = 98304
env = BatchedJitterbugEnv("move_from_origin", num_envs=E, seed=0)
g = torch.Generator(device="cuda").manual_seed(1)
env.reset()
for i in range(150):
    env.step(torch.rand(E, device="cuda", generator=g) * 2 - 1)
lane_of = np.full(32, -1)
lane_of[1], lane_of[2], lane_of[3], lane_of[4] = 0, 1, 2, 2
for k in range(4):
    lane_of[5 + 4 * k:9 + 4 * k] = k
lane_of[21] = lane_of[22] = 3
T = 60
lmax = np.zeros((T, E), dtype=np.int8)
for t in range(T):
    env.step(torch.rand(E, device="cuda", generator=g) * 2 - 1)
    ncon, geom, dist = [x.cpu().numpy() for x in env.contacts(32)]
    cnt = np.zeros((E, 4), dtype=np.int64)
    for j in range(32):
        gg = geom[:, j]
        ok = gg >= 0
        np.add.at(cnt, (np.nonzero(ok)[0], lane_of[gg[ok]]), 1)
    lmax[t] = cnt.max(1)
print("max contacts in one lane, per env snapshot:", (np.bincount(lmax.ravel(), minlength=8)[:8] / lmax.size).round(4))
heavy = lmax >= 2
print("P(heavy)=", heavy.mean().round(4))
for lag in (1, 5, 20, 50):
    a, b = heavy[:-lag], heavy[lag:]
    print(f"lag {lag}: P(heavy at t+lag | heavy at t) = {(a & b).sum() / max(a.sum(), 1):.3f}   P(heavy | not heavy) = {(~a & b).sum() / max((~a).sum(), 1):.3f}")
frac = heavy.mean(0)
print("per-env fraction of heavy snapshots, quantiles 50/90/99/99.9:", np.quantile(frac, [0.5, 0.9, 0.99, 0.999]).round(3))
