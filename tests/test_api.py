"""Host-side mirror of the reference's public interface: suite registry, task validation, TimeStep / specs look-alikes,
heuristic policies, augmented-model generation.  (Environment stepping itself needs the GPU: tests/test_gpu_api.py.)"""
import collections

import numpy as np
import pytest

import hostsim
import jitterbug_b200 as jb
from jitterbug_b200 import augmented_jitterbug, heuristic_policies, mjcf, suite
from jitterbug_b200.batched_env import OBS_DIM, TASKS, observation_layout
from jitterbug_b200.dm import StepType, TaggedTasks, TimeStep
from oracle import oracle as O


def test_suite_registry_matches_reference_tags():
    assert set(suite._DOMAINS) == {"jitterbug", "augmented_jitterbug"}
    assert len(suite.ALL_TASKS) == 10 and suite.ALL_TASKS == suite.BENCHMARKING and suite.EXTRA == ()
    easy = {t for d, t in suite.EASY if d == "jitterbug"}
    hard = {t for d, t in suite.HARD if d == "jitterbug"}
    assert easy == {"move_from_origin", "face_direction", "move_in_direction"}
    assert hard == {"move_to_position", "move_to_pose"}
    assert suite.TASKS_BY_DOMAIN["jitterbug"] == TASKS
    with pytest.raises(ValueError):
        suite.load("jitterbug", "no_such_task")
    with pytest.raises(ValueError):
        suite.load("no_such_domain", "move_from_origin")


def test_task_validation_and_norm_tables():
    with pytest.raises(AssertionError, match="Invalid task"):
        jb.Jitterbug(task="fly")
    t = jb.Jitterbug(random=3, task="move_to_pose", random_pose=False, norm_obs=True)
    assert t.task == "move_to_pose" and t.random_pose is False and t.norm_obs is True
    assert jb.Jitterbug._NORM_ALL.shape == (15, 2)
    assert jb.Jitterbug._norm(np.array([0.05]), 0.0, 0.1)[0] == pytest.approx(0.0)
    assert {k: len(v) for k, v in jb.Jitterbug._NORM_TASKS.items()} == {
        "move_from_origin": 0, "face_direction": 1, "move_in_direction": 4, "move_to_position": 3, "move_to_pose": 4}


def test_observation_layouts():
    for task in TASKS:
        lay = observation_layout(task)
        assert list(lay)[:4] == ["position", "velocity", "motor_position", "motor_velocity"]
        assert max(s.stop for s in lay.values()) == OBS_DIM[task]
    assert list(observation_layout("move_to_pose"))[4:] == ["target_in_jitterbug_frame", "angle_to_target"]
    assert list(observation_layout("move_in_direction"))[4:] == ["angle_to_target", "speed_in_target_frame"]


def test_timestep_and_tagged_tasks():
    ts = TimeStep(StepType.FIRST, None, None, collections.OrderedDict(a=np.zeros(1)))
    assert ts.first() and not ts.mid() and not ts.last()
    assert TimeStep(StepType.LAST, 1.0, 1.0, {}).last()
    s = TaggedTasks()

    @s.add("x", "y")
    def foo():
        return 1
    assert "foo" in s and list(s.tagged("x")) == ["foo"] and list(s.tagged("x", "y")) == ["foo"] and s["foo"]() == 1


def test_heuristic_policies_host_and_device_code_agree_with_reference_port():
    """Three implementations of heuristic_policies.py: the product's host module, the device function (host
    instantiation) and the oracle's independent restatement."""
    rng = np.random.default_rng(0)
    for ti, task in enumerate(TASKS):
        lay = observation_layout(task)
        for _ in range(300):
            obs = rng.uniform(-1, 1, OBS_DIM[task])
            if "target_in_jitterbug_frame" in lay and rng.random() < 0.3:
                obs[lay["target_in_jitterbug_frame"]][:2] *= 0.005
            od = collections.OrderedDict((k, obs[s]) for k, s in lay.items())
            ts = TimeStep(StepType.MID, 0.0, 1.0, od)
            a_host = float(np.ravel(heuristic_policies.POLICIES[task](ts))[0])
            a_ref = float(O.heuristic(task, od))
            a_dev = hostsim.heuristic(ti, obs.astype(np.float32))
            assert a_host == pytest.approx(a_ref, abs=1e-12)
            assert a_dev == pytest.approx(a_ref, abs=2e-6)


def test_augmented_models():
    base = mjcf.compile_xml()
    xml = augmented_jitterbug.augment_Jitterbug(random=0)          # all switches off: only re-serialised
    same = mjcf.compile_xml(xml)
    np.testing.assert_allclose(same.body_mass, base.body_mass, rtol=1e-14)
    models = augmented_jitterbug.make_models(4, seed=1)
    masses = [m.body_mass[1:11].sum() for m in models]
    assert len(set(np.round(masses, 12))) == 4 and all(abs(x - 0.0175165) < 1e-3 for x in masses)
    from jitterbug_b200.model import flatten_model
    f = flatten_model(models[0])
    # recomputed hinge axes are unit and horizontal-ish for hinge 1 (cross(v_upper, z))
    np.testing.assert_allclose(np.linalg.norm(f["a1"], axis=1), 1, atol=1e-12)
    np.testing.assert_allclose(f["a1"][:, 2], 0, atol=1e-12)
    # gear / density switches
    m2 = mjcf.compile_xml(augmented_jitterbug.augment_Jitterbug(modify_gear=True, modify_global_density=True, random=5))
    assert m2.actuator_gear[0] != base.actuator_gear[0] and m2.body_mass[10] != base.body_mass[10]
