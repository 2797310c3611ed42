"""The algorithm the CUDA kernel runs (jitterbug_b200/csrc/jb_dynamics.cuh), instantiated on the HOST by the test-only
build in tests/hostsim, against the oracle.  fp64 isolates algorithmic parity (body-frame product-of-exponentials
formulation + arrow L'DL + Newton) from rounding; fp32 shows what the GPU arithmetic can deliver."""
import numpy as np
import pytest

import hostsim
from helpers import random_state
from jitterbug_b200 import model as jbm
from oracle import oracle as O


@pytest.fixture(scope="module")
def structs(compiled_model):
    flat = jbm.flatten_model(compiled_model)
    return jbm.fill_struct(flat, double=True), jbm.fill_struct(flat, double=False)


def _trajectory(compiled_model, seed, control_steps, tilt=True):
    """Yield (qpos, qvel, warm, ctrl, oracle_next_state, oracle) along an oracle random-policy rollout."""
    sim = O.OracleSim(compiled_model)
    rng = np.random.default_rng(seed)
    if tilt:
        a = rng.uniform(0, 2 * np.pi)
        ax = np.array([rng.uniform(-.025, .025), rng.uniform(-.025, .025), 1.0])
        ax /= np.linalg.norm(ax)
        sim.qpos[3:7] = [np.cos(a / 2), *(np.sin(a / 2) * ax)]
    sim.forward()
    for cs in range(control_steps):
        ctrl = rng.uniform(-1, 1)
        sim.d.ctrl[0] = ctrl
        for k in range(50):
            q, v, w = sim.qpos.copy(), sim.qvel.copy(), sim.warmstart.copy()
            g2, dist, excl, _ = sim.contacts()
            sim.physics_step()
            yield cs, k, q, v, w, ctrl, (g2, dist, excl), sim


def test_smooth_dynamics_fp64(compiled_model, structs):
    md, _ = structs
    sim = O.OracleSim(compiled_model)
    rng = np.random.default_rng(0)
    for _ in range(5):
        q, v = random_state(compiled_model, rng)
        sim.set_state(q, v, ctrl=0.3)
        sim.step1()
        Mo, bo = sim.view("qM")[:15, :15].copy(), sim.view("qfrc_bias")[:15].copy()
        sim.step2()
        q1, v1, w1, it, dg = hostsim.substeps(md, q, v, np.zeros(15), 0.3, 1, True, True)
        T = np.eye(15)
        T[:3, :3] = np.array(dg.Rw).reshape(3, 3)
        np.testing.assert_allclose(T @ np.array(dg.M) @ T.T, Mo, atol=1e-15 * np.abs(Mo).max() * 100)
        np.testing.assert_allclose(T @ np.array(dg.bias), bo, atol=1e-12 * np.abs(bo).max())
        np.testing.assert_allclose(q1, sim.qpos, atol=1e-13)
        np.testing.assert_allclose(v1, sim.qvel, atol=1e-10)


def test_contact_substeps_fp64(compiled_model, structs):
    """Per-sub-step parity from identical states along a contact-rich rollout: contact set identical, state 1e-9."""
    md, _ = structs
    n = 0
    for cs, k, q, v, w, ctrl, (g2, dist, excl), sim in _trajectory(compiled_model, 3, 40):
        if cs == 0 and k == 0:
            continue
        q1, v1, w1, it, dg = hostsim.substeps(md, q, v, w, ctrl, 1, True, True)
        active = sorted(int(g) for g, x in zip(g2, excl) if not x)
        assert sorted(dg.geom[:dg.ncon]) == active
        np.testing.assert_allclose(q1, sim.qpos, atol=1e-11)
        np.testing.assert_allclose(v1, sim.qvel, atol=2e-8)
        np.testing.assert_allclose(w1, sim.warmstart, rtol=1e-6, atol=1e-4)
        assert dg.overflow == 0
        n += 1
    assert n > 1900


def test_contact_substeps_fp32(compiled_model, structs):
    """fp32 arithmetic, one sub-step from identical float32 states: 1e-5 relative on qpos, 1e-4 on velocity blocks."""
    _, mf = structs
    worst_q = worst_v = 0.0
    for cs, k, q, v, w, ctrl, (g2, dist, excl), sim in _trajectory(compiled_model, 4, 30):
        if cs == 0 and k == 0:
            continue
        if any(abs(d) < 1e-6 for d in dist):
            continue
        qf, vf, wf = q.astype(np.float32), v.astype(np.float32), w.astype(np.float32)
        ref = O.OracleSim(compiled_model)
        ref.set_state(qf.astype(np.float64), vf.astype(np.float64), wf.astype(np.float64), ctrl=ctrl)
        ref.step1()
        ref.step2()
        q1, v1, w1, it, _ = hostsim.substeps(mf, qf, vf, wf, ctrl, 1, False, False)
        worst_q = max(worst_q, float((np.abs(q1 - ref.qpos) / np.maximum(np.abs(ref.qpos), 1)).max()))
        scale = np.concatenate([np.full(3, max(1e-2, np.abs(ref.qvel[:3]).max())), np.full(3, max(1., np.abs(ref.qvel[3:6]).max())),
                                np.full(8, max(1., np.abs(ref.qvel[6:14]).max())), [max(1., abs(ref.qvel[14]))]])
        worst_v = max(worst_v, float((np.abs(v1 - ref.qvel) / scale).max()))
    assert worst_q < 1e-5, worst_q
    assert worst_v < 1e-4, worst_v


def test_control_step_fp64_matches_oracle(compiled_model, structs):
    """50 fused sub-steps (what one kernel launch does per env) track the oracle over whole control steps."""
    md, _ = structs
    sim = O.OracleSim(compiled_model)
    rng = np.random.default_rng(9)
    sim.qpos[3:7] = [np.cos(0.3), 0.01 * np.sin(0.3), -0.02 * np.sin(0.3), np.sin(0.3)]
    sim.qpos[3:7] /= np.linalg.norm(sim.qpos[3:7])
    sim.forward()
    sim.control_step(0.5)
    for _ in range(20):
        q, v, w = sim.qpos.copy(), sim.qvel.copy(), sim.warmstart.copy()
        ctrl = rng.uniform(-1, 1)
        sim.control_step(ctrl)
        q1, v1, w1, it, _ = hostsim.substeps(md, q, v, w, ctrl, 50, True, False)
        np.testing.assert_allclose(q1, sim.qpos, atol=1e-7)
        np.testing.assert_allclose(v1, sim.qvel, atol=1e-4)
