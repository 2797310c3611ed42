"""Tuning helper (GPU): distribution of contacts per lane / warp / CTA in the bench workload (one snapshot per control step)."""
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
hist = np.zeros(12, dtype=np.int64)
wmax = np.zeros(12, dtype=np.int64)
cmax = np.zeros(12, dtype=np.int64)
emax_hist = np.zeros(33, dtype=np.int64)
geom_hist = np.zeros(32, dtype=np.int64)
for snap in range(20):
    env.step(torch.rand(E, device="cuda", generator=g) * 2 - 1)
    ncon, geom, dist = [t.cpu().numpy() for t in env.contacts(32)]
    cnt = np.zeros((E, 4), dtype=np.int64)
    for j in range(32):
        gg = geom[:, j]
        ok = gg >= 0
        np.add.at(cnt, (np.nonzero(ok)[0], lane_of[gg[ok]]), 1)
        np.add.at(geom_hist, gg[ok], 1)
    hist += np.bincount(cnt.ravel(), minlength=12)[:12]
    wmax += np.bincount(cnt.reshape(-1, 8 * 4).max(1), minlength=12)[:12]
    cmax += np.bincount(cnt.reshape(-1, 96 * 4).max(1), minlength=12)[:12]
    emax_hist += np.bincount(ncon, minlength=33)[:33]
print("per-lane contact count histogram      ", (hist / hist.sum()).round(4))
print("max over the 32 lanes of a warp       ", (wmax / wmax.sum()).round(4))
print("max over the 384 lanes of a CTA       ", (cmax / cmax.sum()).round(4))
print("contacts per env histogram            ", (emax_hist / emax_hist.sum()).round(4)[:20])
print("contacts by geom id                   ", (geom_hist / geom_hist.sum()).round(3))
