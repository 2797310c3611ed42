"""The step pipeline against the only MuJoCo-EXECUTED numbers the reference repository holds:
the left edge of its paper's PPO2 training curves (tests/golden/published_training_start.json, digitised from the figure
stored in /root/reference/fig-rl-perf.ipynb by tests/golden/read_published_curves.py).  At initialisation
stable-baselines' PPO2 acts with a_t = clip(N(0, 1), -1, 1), independent per control step
(benchmarks/benchmark.py:290-325), and the plotted quantity is the per-episode return (1000 control steps, Monitor)
percentiles over all seeds' monitor files -- a protocol any simulator can replay.

What it can and cannot show.  Returns under a random policy depend on how fast the robot jitters away from the origin
(move_from_origin), on the size of its velocity fluctuations against TARGET_SPEED (move_in_direction) and on how often
it tips over for good (the upright factor), so wrong contact, friction, actuator or inertia numbers move them; but the
reading error is ~5 return units, every published percentile is an estimate from 40-80 episodes per point, the code
version behind the figure (September 2019) is not the checked-out one, and trajectories are chaotic.  So this is a
coarse, distribution-level pin -- it does not replace the MuJoCo-gated tests of tests/test_mujoco_parity.py.

Asserted (bands stated next to each number):
* at sigma = 1 (the policy exactly as initialised): the medians and 90th percentiles of move_from_origin and
  move_in_direction, and "essentially zero" for move_to_position / move_to_pose.  The 10th percentiles are reported
  only: the published low tail is heavier (move_from_origin p10 ~ 200 published vs ~ 650 at sigma = 1).  In both
  simulators the low-return episodes are robots that tipped onto their back and stayed there.
* that ONE policy std reproduces ALL SIX published numbers (median, p10, p90 of both tasks) -- it does, at sigma ~ 1.75:
  the first plotted point is already 1.6e5 environment steps into training with an entropy bonus that widens the
  Gaussian, and the tip-over rate rises steeply with sigma (tests/scripts/flip_hazard_sensitivity.py shows it is as
  steep in friction and contact stiffness).  DESIGN.md section 2 discusses what this does and does not establish.
"""
import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
TASKS = ("move_from_origin", "face_direction", "move_in_direction", "move_to_position", "move_to_pose")
T = 1000

# |ours - published| bands, in return units (published reading error 5; sampling error of a median of ~60 published
# episodes smoothed over 20: ~15)
BANDS = {
    "move_from_origin": dict(median=0.10, p90=0.02),      # relative
    "move_in_direction": dict(median=0.06, p90=0.08),
}


def _fixture():
    with open(os.path.join(HERE, "golden", "published_training_start.json")) as f:
        return json.load(f)


def _published():
    return _fixture()["curves"]


def _report(task, ret, pub, where):
    q = lambda p: float(np.quantile(ret, p))
    print(f"{where} {task}: return median {q(.5):.1f} p10 {q(.1):.1f} p90 {q(.9):.1f} P(<200) {(ret < 200).mean():.3f}"
          f"   published median {pub['median']:.1f} p10 {pub['p10']:.1f} p90 {pub['p90']:.1f}")
    return q(.5), q(.1), q(.9)


def _check(task, ret, where, slack=0.0):
    pub = _published()[f"{task}/ppo2"]
    assert pub["usable"]
    med, p10, p90 = _report(task, ret, pub, where)
    if task in BANDS:
        b = BANDS[task]
        assert abs(med - pub["median"]) <= (b["median"] + slack) * pub["median"], (task, med, pub["median"])
        assert abs(p90 - pub["p90"]) <= (b["p90"] + slack) * pub["p90"], (task, p90, pub["p90"])
        assert p10 < med < p90
    else:                                   # move_to_position / move_to_pose: a random policy earns next to nothing
        assert med < 15.0, (task, med)
        assert p90 < 200.0, (task, p90)


def test_fixture_matches_the_reference_figure():
    """The committed readings are what the digitiser extracts from the reference's notebook (skips without it)."""
    if not os.path.exists("/root/reference/fig-rl-perf.ipynb"):
        pytest.skip("reference checkout not present")
    pytest.importorskip("PIL")
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import read_published_curves as R
    assert R.digitise() == _published()
    assert R.digitise_over_training() == _fixture()["ppo2_over_training"]


@pytest.mark.parametrize("task", ["move_from_origin", "move_in_direction"])
def test_oracle_random_policy_returns_match_the_published_curves(compiled_model, task):
    """fp64 oracle, 96 episodes per task on the host threads (the band is widened by the sampling error of 96)."""
    import hostsim
    from oracle import oracle as O
    ti, M = TASKS.index(task), 96
    rs = hostsim.reset_states(1234 + ti, np.arange(M), 0, ti, 1)
    q0 = np.tile(np.asarray(compiled_model.qpos0, dtype=np.float64), (M, 1))
    q0[:, 3:7] = rs[:, :4]
    a = np.clip(np.random.RandomState(5 + ti).randn(T, M), -1, 1)
    rew = O.batch_rollout(compiled_model, ti, q0, rs[:, 4:8].astype(np.float64), T, actions=a)
    _check(task, rew.sum(0), "oracle", slack=0.04)


def _cuda_returns(task, sigma, M=8192, seed=77):
    """One 1000-step episode per env through the public batched API under a_t = clip(sigma N(0,1), -1, 1); rewards
    summed like stable-baselines' Monitor."""
    import torch
    from jitterbug_b200 import BatchedJitterbugEnv
    env = BatchedJitterbugEnv(task, num_envs=M, seed=seed)
    env.reset()
    g = torch.Generator(device=env.device).manual_seed(11)
    ret = torch.zeros(M, device=env.device)
    for k in range(T):
        a = (sigma * torch.randn(M, device=env.device, generator=g)).clamp_(-1, 1)
        _, r, st = env.step(a)
        ret += r
    assert int(st.min()) == 2 and int(st.max()) == 2                  # every env ended its episode on step 1000
    assert env.counters()["diverged_resets"] == 0
    return ret.cpu().numpy().astype(np.float64)


def _need_gpu():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


@pytest.mark.gpu
@pytest.mark.parametrize("task", [t for t in TASKS if t != "face_direction"])
def test_cuda_random_policy_returns_match_the_published_curves(task):
    """CUDA path, 8192 episodes per task, the policy exactly as PPO2 initialises it (sigma = 1)."""
    _need_gpu()
    _check(task, _cuda_returns(task, 1.0), "cuda")


SIGMAS = (1.0, 1.25, 1.5, 1.75, 2.0, 2.5, 3.0, 4.0, 6.0)
_FAMILY = {}


def _family(task):
    """sigma -> (median, p10, p90) of the per-episode return on the CUDA path (4096 episodes each), cached per module."""
    if task not in _FAMILY:
        _FAMILY[task] = {}
        for sigma in SIGMAS:
            ret = _cuda_returns(task, sigma, M=4096)
            _FAMILY[task][sigma] = tuple(float(np.quantile(ret, q)) for q in (.5, .1, .9))
    return _FAMILY[task]


@pytest.mark.gpu
def test_cuda_one_policy_std_reproduces_all_published_percentiles():
    """The first plotted point already averages the first ~20 episodes of every monitor file, 1.6e5 environment steps
    into training with an entropy bonus (ent_coef 0.01, benchmarks/benchmark.py:295) that widens the Gaussian; the
    policy std behind it is therefore >= 1 but unknown.  One parameter must then explain six published numbers: there is
    a sigma for which median, 10th and 90th percentile of BOTH move_from_origin and move_in_direction agree with the
    figure -- including the low tail (robots tipped onto their back), which sigma = 1 leaves too light."""
    _need_gpu()
    pub = _published()
    tasks = ("move_from_origin", "move_in_direction")
    tol = dict(median=0.06, p90=0.03, p10=50.0)            # relative, relative, absolute return units
    table, ok = {}, []
    for sigma in SIGMAS[:6]:
        good, row = True, []
        for task in tasks:
            (med, p10, p90), p = _family(task)[sigma], pub[f"{task}/ppo2"]
            row += [med, p10, p90]
            good &= abs(med - p["median"]) <= tol["median"] * p["median"] and abs(p90 - p["p90"]) <= tol["p90"] * p["p90"] \
                and abs(p10 - p["p10"]) <= tol["p10"]
        table[sigma] = row
        print(f"sigma {sigma:4.2f}: mfo median/p10/p90 {row[0]:6.1f} {row[1]:6.1f} {row[2]:6.1f}   "
              f"mid {row[3]:6.1f} {row[4]:6.1f} {row[5]:6.1f}   {'all six within tolerance' if good else ''}")
        if good:
            ok.append(sigma)
    print("published      : mfo median/p10/p90 " + " ".join(f"{pub[f'{t}/ppo2'][k]:6.1f}" for t in tasks for k in ("median", "p10", "p90")))
    assert ok, table
    assert all(1.25 <= s_ <= 2.5 for s_ in ok), ok


@pytest.mark.gpu
@pytest.mark.parametrize("task,p10_tol", [("move_from_origin", 100.0), ("move_in_direction", 60.0)])
def test_cuda_std_family_traces_the_published_ppo2_curves(task, p10_tol):
    """PPO2 never learns these two tasks in the paper; over its 2-3.5e6 training steps the plotted (median, p10, p90)
    drifts (move_from_origin: 826/206/954 -> 423/78/918) the way a zero-mean policy of growing std does.  Every
    published point whose median lies inside the range our sigma family spans must then have the family's p10 and p90 at
    that median (interpolated in sigma): a relation between the percentiles that does not need sigma(t).  It checks the
    structure of the return distribution (permanent tip-overs at a constant hazard on top of a narrow "upright" mode),
    not a single number.  Loose bands -- p10 within 100 / 60 return units at every point and within 60 % of that on
    average, p90 within 6 % --: the published percentiles come from 40-80 episodes per point, and PPO2's policy mean
    moves as well.  Measured: our p10 sits below the published one at equal median (by ~ 50 on move_from_origin, ~ 22 on
    move_in_direction), i.e. the published "upright" mode is ~ 4 % lower than ours, the same offset the sigma = 1.75
    fit of the first point shows."""
    _need_gpu()
    fam = _family(task)
    pts = np.array([fam[s_] for s_ in SIGMAS if s_ >= 1.5])             # medians fall monotonically from sigma = 1.5 on
    order = np.argsort(pts[:, 0])
    med, p10, p90 = pts[order].T
    assert np.all(np.diff(med) > 0)
    curve = np.array(_fixture()["ppo2_over_training"][task])
    inside = curve[(curve[:, 1] >= med[0]) & (curve[:, 1] <= med[-1])]
    assert len(inside) >= 0.7 * len(curve), (len(inside), len(curve), med[0], med[-1])
    d10 = np.interp(inside[:, 1], med, p10) - inside[:, 2]
    d90 = np.interp(inside[:, 1], med, p90) / inside[:, 3] - 1.0
    print(f"{task}: {len(inside)} of {len(curve)} published points inside the family's median range [{med[0]:.0f}, {med[-1]:.0f}]; "
          f"p10 ours - published: mean {d10.mean():+.1f}, worst {np.abs(d10).max():.1f}; p90 relative: mean {d90.mean():+.4f}, "
          f"worst {np.abs(d90).max():.4f}")
    assert np.abs(d10).max() <= p10_tol and abs(d10.mean()) <= 0.6 * p10_tol
    assert np.abs(d90).max() <= 0.06
