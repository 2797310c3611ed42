"""Offline parity replay for anyone who has MuJoCo (this project's containers do not):

    python -m jitterbug_b200.recorder replay.npz --envs 256 --steps 40        # on a GPU box with this repo
    python tests/scripts/replay_with_mujoco.py replay.npz                     # anywhere `import mujoco` works

For every recorded (env, control step) the script loads the identical qpos / qvel / qacc_warmstart / ctrl into MuJoCo
(model: jitterbug_b200/assets/jitterbug_local.xml), compares the active contact set, advances n_sub_steps and compares
the next qpos / qvel with what the simulator produced (north_star tolerances: 1e-5 relative after one sub-step; after a
50-sub-step control step the bulk of envs stays below 1e-4, contact make / break events amplify fp32 rounding)."""
import os
import sys

import numpy as np

import mujoco

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
m = mujoco.MjModel.from_xml_path(os.path.join(ROOT, "jitterbug_b200", "assets", "jitterbug_local.xml"))
d = mujoco.MjData(m)
r = np.load(sys.argv[1])
floor = mujoco.mj_name2id(m, mujoco.mjtObj.mjOBJ_GEOM, "floor")
tid = mujoco.mj_name2id(m, mujoco.mjtObj.mjOBJ_BODY, "target")
T, N = r["qpos"].shape[0] - 1, r["qpos"].shape[1]
n_sub = int(r["n_sub_steps"])
eq, ev, bad_sets = [], [], 0
for k in range(T):
    for e in range(N):
        mujoco.mj_resetData(m, d)
        d.qpos[:], d.qvel[:], d.qacc_warmstart[:] = r["qpos"][k, e], r["qvel"][k, e], r["qacc_warmstart"][k, e]
        d.ctrl[:] = r["ctrl"][k + 1, e]
        m.body_pos[tid, :2] = r["target"][k, e, :2]
        m.body_quat[tid] = (r["target"][k, e, 2], 0, 0, r["target"][k, e, 3])
        mujoco.mj_forward(m, d)
        got = sorted(int(c.geom2 if c.geom1 == floor else c.geom1) for c in d.contact[:d.ncon] if c.dist < 0)
        want = [int(g) for g in r["contact_geom"][k, e, :r["ncon"][k, e]]]
        bad_sets += got != want
        for _ in range(n_sub):
            mujoco.mj_step(m, d)
        eq.append(np.abs(d.qpos - r["qpos"][k + 1, e]).max())
        ev.append((np.abs(d.qvel - r["qvel"][k + 1, e]) / np.maximum(np.abs(d.qvel), 1)).max())
eq, ev = np.array(eq), np.array(ev)
print(f"{T * N} control steps replayed: contact-set mismatches {bad_sets}; |dqpos| median {np.median(eq):.2e} "
      f"p99 {np.quantile(eq, .99):.2e} max {eq.max():.2e}; scaled |dqvel| median {np.median(ev):.2e} p99 {np.quantile(ev, .99):.2e}")
