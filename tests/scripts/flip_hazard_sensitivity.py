"""How sensitive the low tail of tests/test_published_curves.py is to contact parameters: move_from_origin returns of the
fp64 oracle under the PPO2-initial policy a_t = clip(N(0,1), -1, 1) with the sliding friction or the contact time
constant (solref[0]) scaled.  The low-return episodes are robots that tipped onto their back for good; the tip-over
rate changes by an order of magnitude for a factor 2 in either parameter, so the published tail (p10 ~ 200, i.e. ~ 10 %
of episodes on their back within the first 2 s) does not discriminate between simulators the way the medians do.
usage: python tests/scripts/flip_hazard_sensitivity.py [episodes]   (CPU, ~40 s per row at 384 episodes on 8 threads)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import hostsim  # noqa: E402
from jitterbug_b200 import mjcf  # noqa: E402
from oracle import oracle as O  # noqa: E402

M, T, ti = int(sys.argv[1]) if len(sys.argv) > 1 else 384, 1000, 0
rs = hostsim.reset_states(99, np.arange(M), 0, ti, 1)
a = np.clip(np.random.RandomState(3).randn(T, M), -1, 1)


def run(tag, edit):
    cm = mjcf.compile_xml()
    edit(cm)
    q0 = np.tile(np.asarray(cm.qpos0, dtype=np.float64), (M, 1))
    q0[:, 3:7] = rs[:, :4]
    rew = O.batch_rollout(cm, ti, q0, rs[:, 4:8].astype(np.float64), T, actions=a)
    ret = rew.sum(0)
    print(f"{tag:22s} median {np.median(ret):6.1f}  p10 {np.quantile(ret, .1):6.1f}  p90 {np.quantile(ret, .9):6.1f}  "
          f"P(return<200) {(ret < 200).mean():.3f}  on its back at the end {(rew[-50:].max(0) < 0.05).mean():.3f}", flush=True)


def friction(s):
    def edit(cm):
        cm.geom_friction = np.asarray(cm.geom_friction) * np.array([s, 1, 1])
    return edit


def timeconst(tc):
    def edit(cm):
        g = np.asarray(cm.geom_solref).copy()
        g[:, 0] = tc
        cm.geom_solref = g
    return edit


print(f"move_from_origin, {M} episodes x {T} control steps, published (MuJoCo, PPO2 at initialisation): median 825  p10 199  p90 952")
run("as compiled", lambda cm: None)
for s in (0.5, 0.75, 1.5, 2.0):
    run(f"friction x{s}", friction(s))
for tc in (0.01, 0.005):
    run(f"solref[0] = {tc}", timeconst(tc))
