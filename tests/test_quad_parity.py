"""The quad (4 lanes per env) algorithm the CUDA kernel runs, emulated on the host with fibers (tests/hostsim/quadsim.cpp),
against the oracle along fresh rollouts (the golden test covers fixed vectors)."""
import numpy as np
import pytest

import hostsim
from helpers import random_state
from jitterbug_b200 import model as jbm
from oracle import oracle as O


@pytest.fixture(scope="module")
def packed(compiled_model):
    flat = jbm.flatten_model(compiled_model)
    return hostsim.pack(jbm.fill_struct(flat, double=True), True), hostsim.pack(jbm.fill_struct(flat, double=False), False)


def test_airborne_states_fp64(compiled_model, packed):
    Pd, _ = packed
    sim = O.OracleSim(compiled_model)
    rng = np.random.default_rng(0)
    for _ in range(6):
        q, v = random_state(compiled_model, rng)
        sim.set_state(q, v, ctrl=-0.7)
        sim.step1()
        sim.step2()
        q1, v1, w1, it, ncon, _, _ = hostsim.quad_substeps(Pd, q, v, np.zeros(15), -0.7)
        assert ncon == 0 and it == 0
        np.testing.assert_allclose(q1, sim.qpos, atol=1e-13)
        np.testing.assert_allclose(v1, sim.qvel, atol=1e-9)


def test_rollout_substeps_fp64(compiled_model, packed):
    Pd, _ = packed
    sim = O.OracleSim(compiled_model)
    rng = np.random.default_rng(21)
    sim.qpos[3:7] = [np.cos(1.0), 0.02 * np.sin(1.0), 0.01 * np.sin(1.0), np.sin(1.0)]
    sim.forward()
    n = 0
    moved0 = hostsim.quad_debug_counts()[2]
    for cs in range(30):
        ctrl = rng.uniform(-1, 1)
        sim.d.ctrl[0] = ctrl
        for k in range(50):
            q, v, w = sim.qpos.copy(), sim.qvel.copy(), sim.warmstart.copy()
            g2, dist, excl, _ = sim.contacts()
            sim.physics_step()
            if any(abs(d) < 1e-9 for d in dist):
                continue
            q1, v1, w1, it, ncon, cons, ovf = hostsim.quad_substeps(Pd, q, v, w, ctrl, 1, False, True)
            assert [c for c, _ in cons] == sorted(int(g) for g, x in zip(g2, excl) if not x)
            np.testing.assert_allclose(q1, sim.qpos, atol=1e-12)
            np.testing.assert_allclose(v1, sim.qvel, atol=5e-9)
            n += 1
    assert n > 1400
    # the rollout includes legs standing on the rim of their lower cylinder (4 contacts in one lane): the spread step
    # that hands such contacts to the quad's other lanes has been exercised
    assert hostsim.quad_debug_counts()[2] - moved0 > 20


def test_forward_only_is_a_zero_length_step(compiled_model, packed):
    Pd, _ = packed
    sim = O.OracleSim(compiled_model)
    sim.qpos[3:7] = [np.cos(0.2), 0, 0.02 * np.sin(0.2), np.sin(0.2)]
    sim.forward()
    for _ in range(20):
        sim.control_step(0.4)
    q, v, w = sim.qpos.copy(), sim.qvel.copy(), sim.warmstart.copy()
    sim.set_state(q, v, np.zeros(15), ctrl=0.0)
    sim.forward()
    q1, v1, w1, *_ = hostsim.quad_substeps(Pd, q, v, np.zeros(15), 0.0, 1, True)
    np.testing.assert_allclose(q1[:3], q[:3], atol=0)
    np.testing.assert_allclose(v1, v, atol=0)
    np.testing.assert_allclose(w1, sim.warmstart, rtol=1e-6, atol=1e-5)


def test_many_contacts_per_lane(compiled_model, packed):
    """Core pressed into the floor: the lanes that own the core boxes carry 4 box corners + foot + lower-leg contacts
    (6 of the 8 slots a lane has)."""
    Pd, Pf = packed
    sim = O.OracleSim(compiled_model)
    rng = np.random.default_rng(5)
    seen = 0
    for trial in range(6):
        q = compiled_model.qpos0.copy()
        q[2] = 0.021 - 0.001 * trial
        ang = 0.02 * rng.standard_normal(3)
        q[3:7] = [1.0, *(0.5 * ang)]
        q[3:7] /= np.linalg.norm(q[3:7])
        v = 0.05 * rng.standard_normal(15)
        sim.set_state(q, v, ctrl=0.3)
        sim.forward()
        g2, dist, excl, _ = sim.contacts()
        if any(abs(d) < 1e-9 for d in dist):
            continue
        w = sim.warmstart.copy()
        active = sorted(int(g) for g, x in zip(g2, excl) if not x)
        sim.physics_step()
        q1, v1, w1, it, ncon, cons, ovf = hostsim.quad_substeps(Pd, q, v, w, 0.3, 1, False, True)
        assert ovf == 0 and [c for c, _ in cons] == active
        assert ncon >= 12                                     # 8 box corners alone
        np.testing.assert_allclose(q1, sim.qpos, atol=1e-12)
        np.testing.assert_allclose(v1, sim.qvel, rtol=1e-7, atol=5e-8)
        qf, vf, *_ = hostsim.quad_substeps(Pf, q, v, w, 0.3, 1, False, False)
        np.testing.assert_allclose(vf, sim.qvel, rtol=2e-3, atol=2e-3)
        seen += 1
    assert seen >= 4
