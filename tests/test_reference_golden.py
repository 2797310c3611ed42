"""Parity against outputs of the REFERENCE'S OWN CODE (tests/golden/reference_*.npz, produced by
tests/golden/make_reference_golden.py executing /root/reference's heuristic_policies.py, jitterbug.py and
augmented_jitterbug.py; see that script for what is stubbed).  Covered per SURVEY.md section 8:

  a6-a15 Physics helpers, a16-a19 normalisation / observation / reward, a20 initialize_episode (the mapping from the
  reference's uniforms to initial state and target), a3 factories, a22 heuristic policies, a25 augment_Jitterbug

each for the host module, the oracle (numpy + C), the device function compiled for the host, and -- under ``-m gpu`` --
the CUDA kernels through the C-ABI.
"""
import collections
import inspect
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import reference_inputs as RI  # noqa: E402

import hostsim  # noqa: E402
from jitterbug_b200 import augmented_jitterbug as jaug  # noqa: E402
from jitterbug_b200 import heuristic_policies as hp  # noqa: E402
from jitterbug_b200 import jitterbug as jb  # noqa: E402
from jitterbug_b200 import mjcf, model as jbm  # noqa: E402
from oracle import oracle as O  # noqa: E402

TASKS = RI.TASKS
TS = collections.namedtuple("TS", "observation")


def _load(name):
    return np.load(os.path.join(HERE, "golden", name), allow_pickle=False)


def _obs_dict(task, row):
    d = collections.OrderedDict()
    d["position"], d["velocity"] = row[0:7], row[7:13]
    d["motor_position"], d["motor_velocity"] = row[13:14], row[14:15]
    c = 15
    for key in RI.EXTRA_KEYS[task]:
        d[key] = row[c:c + RI.KEY_DIMS[key]]
        c += RI.KEY_DIMS[key]
    return d


def _target4(tgt):
    """(x, y, yaw) -> (x, y, quat_w, quat_z)."""
    return np.stack([tgt[:, 0], tgt[:, 1], np.cos(tgt[:, 2] / 2), np.sin(tgt[:, 2] / 2)], axis=1)


# ------------------------------------------------------------------------------------------------ a22 heuristic policies
def _knife_edge(task, row, eps=4e-6):
    """True if the (proven-equal) host policy changes its decision under an eps perturbation of the inputs it reads."""
    base = float(np.ravel(hp.POLICIES[task](TS(_obs_dict(task, row))))[0])
    for j in range(13, len(row)):
        for sgn in (-1, 1):
            r = row.copy(); r[j] += sgn * eps
            if abs(float(np.ravel(hp.POLICIES[task](TS(_obs_dict(task, r))))[0]) - base) > 1e-4:
                return True
    return False


def test_host_heuristic_policies_equal_the_reference():
    g = _load("reference_heuristics.npz")
    for task in TASKS:
        obs = RI.heuristic_observations(task).astype(np.float64)
        assert float(obs.sum()) == float(g[task + "_obs_checksum"])
        act = np.array([float(np.ravel(hp.POLICIES[task](TS(_obs_dict(task, o))))[0]) for o in obs])
        assert np.array_equal(act, g[task + "_action"]), task          # same float64 arithmetic: exact


def _check_fp32_actions(task, obs32, act32, ref):
    diff = np.abs(act32.astype(np.float64) - ref)
    bad = np.nonzero(diff > 2e-6)[0]
    for i in bad:
        assert _knife_edge(task, obs32[i].astype(np.float64)), (task, int(i), act32[i], ref[i])
    assert len(bad) <= 1e-3 * len(ref), (task, len(bad))
    return len(bad)


def test_device_heuristic_code_on_host_equals_the_reference():
    g = _load("reference_heuristics.npz")
    for ti, task in enumerate(TASKS):
        obs = RI.heuristic_observations(task)
        act = np.array([hostsim.heuristic(ti, o) for o in obs], dtype=np.float32)
        _check_fp32_actions(task, obs, act, g[task + "_action"])


def test_oracle_heuristic_restatement_equals_the_reference():
    g = _load("reference_heuristics.npz")
    for task in TASKS:
        obs = RI.heuristic_observations(task)[:2000].astype(np.float64)
        act = np.array([float(np.ravel(O.heuristic(task, _obs_dict(task, o)))[0]) for o in obs])
        assert np.array_equal(act, g[task + "_action"][:2000]), task
        # and the C version that drives the oracle's closed-loop rollouts (long-horizon distribution tests)
        act_c = np.array([O.heuristic_c(TASKS.index(task), o) for o in obs])
        np.testing.assert_allclose(act_c, g[task + "_action"][:2000], rtol=0, atol=1e-15)    # libm vs numpy atan2: 1 ulp


# ------------------------------------------------------------------------------------------------ a6-a19 task layer
def test_host_task_methods_equal_the_reference(compiled_model):
    g = _load("reference_tasks.npz")
    for task in TASKS:
        qpos, qvel, tgt = RI.task_states(task)
        assert float(qpos.sum() + qvel.sum() + tgt.sum()) == float(g[task + "_state_checksum"])
        t4 = _target4(tgt)
        tk = jb.Jitterbug(random=0, task=task)
        assert list(g[task + "_keys"]) == list(tk.get_observation(jb.Physics.from_state(compiled_model, qpos[0], qvel[0], t4[0])))
        for i in range(len(qpos)):
            ph = jb.Physics.from_state(compiled_model, qpos[i], qvel[i], t4[i])
            o = tk.get_observation(ph)
            assert [np.size(v) for v in o.values()] == list(g[task + "_dims"])
            np.testing.assert_allclose(np.concatenate([np.ravel(v) for v in o.values()]), g[task + "_obs"][i], rtol=0, atol=1e-12)
            assert tk.get_reward(ph) == pytest.approx(g[task + "_reward"][i], abs=1e-12)
            comps = (tk.heading_reward(ph), tk.velocity_reward(ph), tk.position_reward(ph), tk.upright_reward(ph))
            np.testing.assert_allclose(comps, g[task + "_reward_components"][i], rtol=0, atol=1e-12)
            helpers = np.concatenate([[ph.jitterbug_direction_yaw()], ph.motor_position(), [ph.target_direction_yaw()],
                                      ph.target_position_in_jitterbug_frame(), ph.jitterbug_velocity_in_target_frame(),
                                      ph.angle_jitterbug_to_target(), [ph.jitterbug_upright()]])
            np.testing.assert_allclose(helpers, g[task + "_helpers"][i], rtol=0, atol=1e-12)


def test_oracle_task_layer_equals_the_reference(compiled_model):
    """The checker itself (numpy TaskOracle and the C task layer) against reference-executed outputs."""
    g = _load("reference_tasks.npz")
    sim = O.OracleSim(compiled_model)
    for ti, task in enumerate(TASKS):
        qpos, qvel, tgt = RI.task_states(task)
        tk = O.TaskOracle(sim, task)
        for i in range(0, len(qpos), 2):
            sim.set_state(qpos[i], qvel[i])
            sim.set_target(tgt[i, 0], tgt[i, 1], tgt[i, 2])
            sim.step1()
            o = np.concatenate([np.ravel(v) for v in tk.get_observation().values()])
            np.testing.assert_allclose(o, g[task + "_obs"][i], rtol=0, atol=1e-12)
            assert tk.get_reward() == pytest.approx(g[task + "_reward"][i], abs=1e-12)
            o_c, r_c = sim.obs_reward_c(ti)
            np.testing.assert_allclose(o_c, g[task + "_obs"][i], rtol=0, atol=1e-12)
            assert r_c == pytest.approx(g[task + "_reward"][i], abs=1e-12)


OBS_ATOL_F32 = 4e-6     # normalised units; dominated by rounding the state itself to float32 (motor angle up to 40 rad)
REW_ATOL_F32 = 2e-5


def test_device_task_code_on_host_equals_the_reference():
    g = _load("reference_tasks.npz")
    for ti, task in enumerate(TASKS):
        qpos, qvel, tgt = RI.task_states(task)
        t4 = _target4(tgt)
        for i in range(len(qpos)):
            obs, rew = hostsim.obs_reward(qpos[i], qvel[i], ti, t4[i])
            np.testing.assert_allclose(obs, g[task + "_obs"][i], rtol=0, atol=OBS_ATOL_F32)
            assert rew == pytest.approx(g[task + "_reward"][i], abs=REW_ATOL_F32)


# ------------------------------------------------------------------------------------------------ a20 initialize_episode
@pytest.fixture(scope="module")
def packed(compiled_model):
    flat = jbm.flatten_model(compiled_model)
    return hostsim.pack(jbm.fill_struct(flat, double=True), True), hostsim.pack(jbm.fill_struct(flat, double=False), False)


def test_reset_mapping_equals_the_reference(packed):
    """The reference's initialize_episode run with known seeds vs the device reset mapping fed the same uniforms."""
    g = _load("reference_resets.npz")
    Pd, Pf = packed
    for ti, task in enumerate(TASKS):
        for rp in (True, False):
            key = f"{task}_{'pose' if rp else 'nopose'}"
            want = np.concatenate([g[key + "_quat"], g[key + "_target"][:, [0, 1, 3, 6]]], axis=1)
            got = hostsim.reset_from_uniforms(g[key + "_uniforms"], ti, int(rp), Pd)
            np.testing.assert_allclose(got, want, rtol=0, atol=1e-14)
            got32 = hostsim.reset_from_uniforms(g[key + "_uniforms"], ti, int(rp), Pf)
            np.testing.assert_allclose(got32, want, rtol=0, atol=1e-6)
            assert np.allclose(g[key + "_target"][:, 2], 0.035) and np.allclose(g[key + "_target"][:, 4:6], 0)


# ------------------------------------------------------------------------------------------------ a3 factories, a16 tables
def test_factories_and_tables_equal_the_reference():
    g = _load("reference_factories.npz")
    assert tuple(g["constants"]) == (jb.DEFAULT_TIME_LIMIT, jb.DEFAULT_CONTROL_TIMESTEP, jb.TARGET_SPEED)
    np.testing.assert_array_equal(g["norm_all"], jb.Jitterbug._NORM_ALL)
    for task in TASKS:
        np.testing.assert_array_equal(g["norm_" + task], jb.Jitterbug._NORM_TASKS[task])
    mods = {"jitterbug": jb, "augmented_jitterbug": jaug}
    seen = collections.defaultdict(list)
    for modname, name, task, time_limit, control_timestep, tags, random_pose in g["factories"]:
        m = mods[modname]
        seen[modname].append(name)
        assert name == task and name in m.SUITE
        assert sorted(t for t in m.SUITE.tags() if name in m.SUITE.tagged(t)) == tags.split(",")
        sig = inspect.signature(m.SUITE[name])
        assert float(sig.parameters["time_limit"].default) == float(time_limit)
        assert float(sig.parameters["control_timestep"].default) == float(control_timestep)
        assert inspect.signature(jb.Jitterbug.__init__).parameters["random_pose"].default == (random_pose == "True")
    for modname, m in mods.items():
        assert seen[modname] == list(m.SUITE.keys())          # same tasks, same registration order


# ------------------------------------------------------------------------------------------------ a25 augment_Jitterbug
def _numeric_attrs(xml):
    import xml.etree.ElementTree as ET
    root = ET.fromstring(xml)
    out = {}
    for el in root.iter():
        if el.tag in ("geom", "joint", "general") and el.attrib.get("name"):
            for a in ("fromto", "pos", "axis", "density", "size", "gear"):
                if a in el.attrib:
                    out[f"{el.tag}:{el.attrib['name']}:{a}"] = np.array([float(x) for x in el.attrib[a].split()])
    for d in root.findall("default"):
        for geom in d.findall("geom"):
            out["default:geom:density"] = np.array([float(geom.attrib["density"])])
    return out


def test_augment_jitterbug_equals_the_reference():
    """Same numpy stream (np.random.seed(s) == RandomState(s)), same draw order, same perturbed model after compile."""
    g = _load("reference_augment.npz")
    n_cases, reps = [int(x) for x in g["n_cases"]]
    for ci in range(n_cases):
        for r in range(reps):
            key = f"case{ci}_rep{r}"
            kw = {str(k): True for k in g[key + "_kwargs"]}
            xml = jaug.augment_Jitterbug(random=np.random.RandomState(int(g[key + "_seed"])), **kw)
            mine = _numeric_attrs(xml)
            for item in g[key + "_attrs"]:
                name, val = str(item).split("=")
                if name in mine:         # visual-only elements of the reference XML are not in the packaged model
                    np.testing.assert_allclose(mine[name], [float(x) for x in val.split()], rtol=0, atol=1e-15, err_msg=name)
            cm = mjcf.compile_xml(xml)
            for f in ("body_mass", "body_inertia", "body_ipos", "body_iquat", "body_pos", "body_quat", "jnt_pos", "jnt_axis",
                      "body_invweight0", "dof_invweight0"):
                if key + "_" + f in g.files:
                    np.testing.assert_allclose(np.asarray(getattr(cm, f), dtype=np.float64), g[key + "_" + f], rtol=0,
                                               atol=1e-12, err_msg=f"{key} {f}")


# ------------------------------------------------------------------------------------------------ CUDA kernels (C-ABI)
def _need_gpu():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


@pytest.mark.gpu
def test_cuda_heuristic_kernel_equals_the_reference():
    torch = _need_gpu()
    from jitterbug_b200.batched_env import BatchedJitterbugEnv
    g = _load("reference_heuristics.npz")
    for task in TASKS:
        obs = RI.heuristic_observations(task)
        env = BatchedJitterbugEnv(task, num_envs=len(obs))
        act = env.heuristic_actions(torch.as_tensor(obs, device=env.device)).cpu().numpy()
        _check_fp32_actions(task, obs, act, g[task + "_action"])


@pytest.mark.gpu
def test_cuda_observation_and_reward_kernels_equal_the_reference():
    torch = _need_gpu()
    from jitterbug_b200.batched_env import BatchedJitterbugEnv, OBS_DIM
    g = _load("reference_tasks.npz")
    for task in TASKS:
        qpos, qvel, tgt = RI.task_states(task)
        n = len(qpos)
        env = BatchedJitterbugEnv(task, num_envs=n)
        env.reset()
        env.set_state(qpos, qvel, np.zeros((n, 15)), target=_target4(tgt))
        obs, rew = env.observe()
        assert obs.shape == (n, OBS_DIM[task])
        np.testing.assert_allclose(obs.cpu().numpy(), g[task + "_obs"], rtol=0, atol=OBS_ATOL_F32)
        np.testing.assert_allclose(rew.cpu().numpy(), g[task + "_reward"], rtol=0, atol=REW_ATOL_F32)


@pytest.mark.gpu
def test_cuda_mixed_batch_observation_kernel_equals_the_reference():
    """Config 3's layout: five tasks in one batch, observations padded to 19."""
    torch = _need_gpu()
    from jitterbug_b200.batched_env import BatchedJitterbugEnv, OBS_DIM
    g = _load("reference_tasks.npz")
    n = 200
    names, qp, qv, tg, want_o, want_r = [], [], [], [], [], []
    for task in TASKS:
        qpos, qvel, tgt = RI.task_states(task)
        names += [task] * n
        qp.append(qpos[:n]); qv.append(qvel[:n]); tg.append(_target4(tgt[:n]))
        o = np.zeros((n, 19)); o[:, :OBS_DIM[task]] = g[task + "_obs"][:n]
        want_o.append(o); want_r.append(g[task + "_reward"][:n])
    perm = np.random.RandomState(0).permutation(len(names))
    names = [names[i] for i in perm]
    qp, qv, tg = (np.concatenate(x)[perm] for x in (qp, qv, tg))
    env = BatchedJitterbugEnv(names, num_envs=len(names))
    env.reset()
    env.set_state(qp, qv, np.zeros((len(names), 15)), target=tg)
    obs, rew = env.observe()
    np.testing.assert_allclose(obs.cpu().numpy(), np.concatenate(want_o)[perm], rtol=0, atol=OBS_ATOL_F32)
    np.testing.assert_allclose(rew.cpu().numpy(), np.concatenate(want_r)[perm], rtol=0, atol=REW_ATOL_F32)


@pytest.mark.gpu
def test_cuda_reset_distributions_match_the_reference():
    """Philox streams cannot reproduce numpy's (App. C4); the distributions must: two-sample KS on the quantities the
    reference's initialize_episode randomises (400 reference episodes per task vs 8192 device resets)."""
    _need_gpu()
    from scipy import stats
    from jitterbug_b200.batched_env import BatchedJitterbugEnv
    g = _load("reference_resets.npz")
    for task in TASKS:
        env = BatchedJitterbugEnv(task, num_envs=8192, seed=11)
        env.reset()
        q = env.get_state()[0].cpu().numpy().astype(np.float64)[:, 3:7]
        t = env.get_target().cpu().numpy().astype(np.float64)
        rq, rt = g[task + "_pose_quat"], g[task + "_pose_target"]

        def feats(quat, x, y, qw, qz):
            rot = 2 * np.arctan2(np.linalg.norm(quat[:, 1:], axis=1), quat[:, 0])
            ax = quat[:, 1:] / np.linalg.norm(quat[:, 1:], axis=1, keepdims=True)
            return {"rotation_angle": rot, "axis_x": ax[:, 0] / ax[:, 2], "axis_y": ax[:, 1] / ax[:, 2],
                    "target_radius": np.hypot(x, y), "target_angle": np.arctan2(y, x), "target_yaw": 2 * np.arctan2(qz, qw)}
        mine = feats(q, t[:, 0], t[:, 1], t[:, 2], t[:, 3])
        ref = feats(rq, rt[:, 0], rt[:, 1], rt[:, 3], rt[:, 6])
        for k in mine:
            if np.ptp(ref[k]) < 1e-12:                      # not randomised for this task: must be the same constant
                assert np.ptp(mine[k]) < 1e-6 and abs(mine[k][0] - ref[k][0]) < 1e-6, (task, k)
            else:
                assert stats.ks_2samp(mine[k], ref[k]).pvalue > 1e-3, (task, k)


@pytest.mark.gpu
def test_cuda_factories_equal_the_reference():
    _need_gpu()
    g = _load("reference_factories.npz")
    for modname, name, task, time_limit, control_timestep, tags, random_pose in g["factories"]:
        if modname != "jitterbug":
            continue
        env = jb.SUITE[name]()
        assert env.task.task == task and env.task.random_pose == (random_pose == "True")
        assert env._step_limit == float(time_limit) / float(control_timestep)
        assert env.control_timestep() == pytest.approx(float(control_timestep), abs=1e-12)
        ts = env.reset()
        obs = env.task.get_observation(env.physics)         # host twin on the device state == the kernel's observation
        for k in ts.observation:
            np.testing.assert_allclose(obs[k], ts.observation[k], rtol=0, atol=OBS_ATOL_F32)
        env.close()
