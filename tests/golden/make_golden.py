"""Generates tests/golden/*.npz from the fp64 oracle (oracle/jb_oracle.c).  Run from the repo root:

    python tests/golden/make_golden.py

The reference holds no golden vectors and MuJoCo / dm_control cannot be imported here, so
these fixtures pin the ORACLE's behaviour (a regression anchor for oracle, host emulation and CUDA kernel alike); they
are not MuJoCo outputs.  Contents:
  substeps.npz : 240 pre-step states (qpos, qvel, warm, ctrl) sampled along random-policy rollouts, the oracle's state
                 after ONE physics sub-step and after ONE control step (50 sub-steps), and the active contact geoms.
  tasks.npz    : for each of the 5 tasks, 40 states + target poses with the reference-order observation and reward.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from jitterbug_b200 import mjcf  # noqa: E402
from oracle import oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    cm = mjcf.compile_xml()
    rng = np.random.default_rng(20261018)
    pre_q, pre_v, pre_w, ctrls, sub_q, sub_v, sub_w, ctl_q, ctl_v, cons = [], [], [], [], [], [], [], [], [], []
    probe = O.OracleSim(cm)
    for env in range(8):
        sim = O.OracleSim(cm)
        a = rng.uniform(0, 2 * np.pi)
        ax = np.array([rng.uniform(-.025, .025), rng.uniform(-.025, .025), 1.0])
        ax /= np.linalg.norm(ax)
        sim.qpos[3:7] = [np.cos(a / 2), *(np.sin(a / 2) * ax)]
        sim.forward()
        for step in range(40):
            ctrl = rng.uniform(-1, 1)
            sim.d.ctrl[0] = ctrl
            for k in range(50):
                sim.physics_step()
                if step >= 10 and k == int(rng.integers(0, 50)) and len(pre_q) < 30 * (env + 1):
                    q, v, w = sim.qpos.copy(), sim.qvel.copy(), sim.warmstart.copy()
                    c2 = rng.uniform(-1, 1)
                    probe.set_state(q, v, w, ctrl=c2)
                    probe.step1()
                    g2, dist, excl, _ = probe.contacts()
                    if any(abs(d) < 1e-6 for d in dist):
                        continue
                    act = sorted(int(g) for g, x in zip(g2, excl) if not x)
                    probe.step2()
                    pre_q.append(q); pre_v.append(v); pre_w.append(w); ctrls.append(c2)
                    sub_q.append(probe.qpos.copy()); sub_v.append(probe.qvel.copy()); sub_w.append(probe.warmstart.copy())
                    probe.set_state(q, v, w)
                    probe.step1()
                    probe.control_step(c2, 50)
                    ctl_q.append(probe.qpos.copy()); ctl_v.append(probe.qvel.copy())
                    cons.append(act + [-1] * (16 - len(act)))
    np.savez_compressed(os.path.join(HERE, "substeps.npz"), qpos=np.array(pre_q), qvel=np.array(pre_v),
                        warm=np.array(pre_w), ctrl=np.array(ctrls), sub_qpos=np.array(sub_q), sub_qvel=np.array(sub_v),
                        sub_warm=np.array(sub_w), ctl_qpos=np.array(ctl_q), ctl_qvel=np.array(ctl_v),
                        contacts=np.array(cons, dtype=np.int32))
    print("substeps:", len(pre_q))

    out = {}
    sim = O.OracleSim(cm)
    for ti, task in enumerate(O.TASKS):
        qs, vs, tg, obs, rew = [], [], [], [], []
        for i in range(40):
            q = cm.qpos0.copy()
            q[0:2] = rng.uniform(-0.3, 0.3, 2)
            q[2] = rng.uniform(0.03, 0.04)
            tilt = rng.normal(size=4) * np.array([0, .2, .2, 1]) + np.array([1, 0, 0, 0])
            q[3:7] = tilt / np.linalg.norm(tilt)
            q[7:15] = rng.uniform(-.05, .05, 8)
            q[15] = rng.uniform(-20, 20)
            v = np.concatenate([rng.uniform(-.3, .3, 3), rng.uniform(-10, 10, 3), rng.uniform(-20, 20, 8),
                                [rng.uniform(-150, 150)]])
            tx, ty, yaw = rng.uniform(-.2, .2), rng.uniform(-.2, .2), rng.uniform(0, 2 * np.pi)
            sim.set_state(q, v)
            sim.set_target(tx, ty, yaw)
            sim.step1()
            tk = O.TaskOracle(sim, task)
            qs.append(q); vs.append(v); tg.append([tx, ty, np.cos(yaw / 2), np.sin(yaw / 2)])
            obs.append(np.concatenate([np.ravel(x) for x in tk.get_observation().values()]))
            rew.append(tk.get_reward())
        out[f"{task}_qpos"] = np.array(qs); out[f"{task}_qvel"] = np.array(vs); out[f"{task}_target"] = np.array(tg)
        out[f"{task}_obs"] = np.array(obs); out[f"{task}_reward"] = np.array(rew)
    np.savez_compressed(os.path.join(HERE, "tasks.npz"), **out)
    print("tasks: 5 x 40")


if __name__ == "__main__":
    main()
