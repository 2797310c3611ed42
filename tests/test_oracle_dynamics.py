"""Known-answer and independent-derivation checks that anchor the fp64 oracle (parity is otherwise unpinned)."""
import numpy as np
import pytest

from helpers import integrate_pos, numeric_bias, random_state
from jitterbug_b200 import mjcf
from oracle import oracle as O


@pytest.fixture(scope="module")
def sim(compiled_model):
    return O.OracleSim(compiled_model)


def test_mass_matrix_matches_jacobian_form(compiled_model, sim):
    rng = np.random.default_rng(0)
    for _ in range(5):
        q, v = random_state(compiled_model, rng)
        sim.set_state(q, v)
        sim.step1()
        fk = mjcf.forward_kinematics(compiled_model, q)
        M_ref = mjcf.mass_matrix(compiled_model, fk)
        np.testing.assert_allclose(sim.view("qM")[:15, :15], M_ref, atol=1e-18, rtol=1e-12)
        np.testing.assert_allclose(sim.view("xpos")[:12], fk["xpos"], atol=1e-15)
        np.testing.assert_allclose(sim.view("geom_xpos")[:25], fk["geom_xpos"], atol=1e-15)


def test_bias_forces_match_finite_difference_newton_euler(compiled_model, sim):
    """qfrc_bias (RNE, com-based spatial algebra) vs Newton-Euler built from finite differences of FK only."""
    rng = np.random.default_rng(1)
    for _ in range(4):
        q, v = random_state(compiled_model, rng)
        sim.set_state(q, v)
        sim.step1()
        ref = numeric_bias(compiled_model, q, v)
        got = sim.view("qfrc_bias")[:15]
        assert np.abs(got - ref).max() < 5e-5 * np.abs(ref).max()


def test_free_fall(compiled_model, sim):
    q = compiled_model.qpos0.copy()
    q[2] = 1.0
    sim.set_state(q, np.zeros(15))
    sim.forward()
    n = 500
    for _ in range(n):
        sim.physics_step()
    h = compiled_model.timestep
    # semi-implicit Euler: v_n = -g n h ; z_n = z0 - g h^2 n(n+1)/2
    assert sim.qvel[2] == pytest.approx(-9.81 * n * h, rel=1e-9)
    assert sim.qpos[2] == pytest.approx(1.0 - 9.81 * h * h * n * (n + 1) / 2, rel=1e-9)
    # whole-body COM accelerates at -g: no spurious internal motion beyond the sprung legs settling
    assert np.abs(sim.qvel[6:14]).max() < 1.0


def test_motor_terminal_velocity(compiled_model, sim):
    """ctrl = 1 in free flight: motor speed relative to the body -> 1/(0.8*0.00833) = 150.06 rad/s."""
    q = compiled_model.qpos0.copy()
    q[2] = 50.0
    sim.set_state(q, np.zeros(15))
    sim.forward()
    vals = []
    for i in range(160):
        sim.control_step(1.0)
        if i >= 100:
            vals.append(sim.qvel[14])
    # the eccentric rotor makes the free-flying body wobble, so the joint speed ripples around the no-load speed
    assert np.mean(vals) == pytest.approx(1 / (0.8 * 0.00833), rel=1e-2)
    assert np.ptp(vals) < 0.05 * np.mean(vals)


def test_static_equilibrium_normal_force(compiled_model, sim):
    sim.reset()
    sim.forward()
    for _ in range(150):
        sim.control_step(0.0)
    nefc = sim.d.nefc
    f = sim.view("efc_force")[:nefc]
    J = sim.view("efc_J")[:nefc, :15]
    mg = compiled_model.body_mass[1:11].sum() * 9.81
    assert (J.T @ f)[2] == pytest.approx(mg, rel=1e-4)
    g2, dist, excl, _ = sim.contacts()
    assert sorted(g2.tolist()) == [8, 12, 16, 20]
    assert (f >= 0).all()


def test_energy_and_momentum_conservation_in_flight(compiled_model):
    """No gravity, no damping, no actuator: springs only => energy, linear and angular momentum are invariants of the
    continuous dynamics; semi-implicit Euler drifts at first order, so the drift must be small and shrink ~2x when the
    time step is halved (checks M, the Coriolis/centrifugal bias and the passive forces for mutual consistency)."""
    import copy
    rng = np.random.default_rng(5)
    q0, v0 = random_state(compiled_model, rng, vel_scale=0.2, height=5.0)

    def drift(h, nsteps):
        cm = copy.deepcopy(compiled_model)
        cm.gravity = np.zeros(3)
        cm.dof_damping = np.zeros(15)
        cm.actuator_gear = np.zeros(1)
        cm.timestep = h
        sim = O.OracleSim(cm)
        sim.set_state(q0, v0)
        sim.forward()

        def invariants():
            fk = mjcf.forward_kinematics(cm, sim.qpos)
            M = mjcf.mass_matrix(cm, fk)
            T = 0.5 * sim.qvel @ M @ sim.qvel
            V = 0.5 * 0.4 * np.sum(sim.qpos[7:15] ** 2)
            P = np.zeros(3); L = np.zeros(3)
            for b in range(1, 11):
                jp, jr = mjcf.jacobian(cm, fk, fk["xipos"][b], b)
                vc, w = jp @ sim.qvel, jr @ sim.qvel
                iw = fk["ximat"][b] @ np.diag(cm.body_inertia[b]) @ fk["ximat"][b].T
                P += cm.body_mass[b] * vc
                L += iw @ w + np.cross(fk["xipos"][b], cm.body_mass[b] * vc)
            return T + V, P, L

        e0, p0, l0 = invariants()
        for _ in range(nsteps):
            sim.physics_step()
        e1, p1, l1 = invariants()
        return abs(e1 - e0) / abs(e0), np.abs(p1 - p0).max(), np.abs(l1 - l0).max(), np.abs(p0).max(), np.abs(l0).max()

    de1, dp1, dl1, pscale, lscale = drift(2e-4, 200)
    de2, dp2, dl2, _, _ = drift(1e-4, 400)
    assert de1 < 1e-2 and dp1 < 0.1 * pscale and dl1 < 1e-2 * lscale
    for big, small in ((de1, de2), (dp1, dp2), (dl1, dl2)):
        assert 1.4 < big / small < 2.8, (big, small)


def test_tolerance_closed_forms():
    """rewards.tolerance at hand-picked inputs (SURVEY.md section 8c substitute pins)."""
    assert O.tolerance(0.05, bounds=(0, 0), margin=0.05) == pytest.approx(0.1)
    assert O.tolerance(np.pi / 4, bounds=(0, 0), margin=np.pi / 2, value_at_margin=0, sigmoid="cosine") == pytest.approx(0.5)
    assert O.tolerance(0.5, bounds=(1, 1), margin=0.5) == pytest.approx(0.1)
    assert O.tolerance(0.05, bounds=(0.1, float("inf")), margin=0.1, value_at_margin=0, sigmoid="linear") == pytest.approx(0.5)
    assert O.tolerance(0.2, bounds=(0.1, float("inf")), margin=0.1, value_at_margin=0, sigmoid="linear") == 1.0
    assert O.tolerance(-0.1, bounds=(0.1, float("inf")), margin=0.1, value_at_margin=0, sigmoid="linear") == 0.0
    assert O.tolerance(2.0, bounds=(0, 0), margin=np.pi / 2, value_at_margin=0, sigmoid="cosine") == 0.0


def test_task_layer_c_matches_numpy(compiled_model):
    sim = O.OracleSim(compiled_model)
    rng = np.random.default_rng(2)
    for ti, task in enumerate(O.TASKS):
        q, v = random_state(compiled_model, rng, vel_scale=0.05, height=0.04)
        sim.set_state(q, v)
        sim.set_target(rng.uniform(-.2, .2), rng.uniform(-.2, .2), rng.uniform(0, 2 * np.pi))
        sim.step1()
        tk = O.TaskOracle(sim, task)
        o_np = np.concatenate([np.ravel(x) for x in tk.get_observation().values()])
        o_c, r_c = sim.obs_reward_c(ti)
        np.testing.assert_allclose(o_c, o_np, atol=1e-12)
        assert r_c == pytest.approx(tk.get_reward(), abs=1e-12)
        assert len(o_np) == [15, 16, 19, 18, 19][ti]


def test_constraint_solution_satisfies_optimality_conditions(compiled_model):
    """KKT check of the oracle's Newton solution, independent of how it was reached: with f = -D*min(0, J a - aref)
    the stationarity residual M a - qfrc_smooth - J'f vanishes and forces are non-negative (pyramidal rows)."""
    sim = O.OracleSim(compiled_model)
    rng = np.random.default_rng(12)
    sim.qpos[3:7] = [np.cos(0.4), 0.01 * np.sin(0.4), -0.02 * np.sin(0.4), np.sin(0.4)]
    sim.forward()
    checked = 0
    for cs in range(25):
        sim.d.ctrl[0] = rng.uniform(-1, 1)
        for k in range(50):
            sim.physics_step()
            # after physics_step the data holds the position/velocity stage of the NEW state; run the acceleration
            # stage on it without integrating
            q, v, w = sim.qpos.copy(), sim.qvel.copy(), sim.warmstart.copy()
            ctrl = sim.d.ctrl[0]
            probe = O.OracleSim(compiled_model)
            probe.set_state(q, v, w, ctrl=ctrl)
            probe.forward()
            n = probe.d.nefc
            if n == 0:
                continue
            J = probe.view("efc_J")[:n, :15]
            D = probe.view("efc_D")[:n]
            aref = probe.view("efc_aref")[:n]
            a = probe.view("qacc")[:15]
            M = probe.view("qM")[:15, :15]
            jar = J @ a - aref
            f = -D * np.minimum(jar, 0)
            res = M @ a - probe.view("qfrc_smooth")[:15] - J.T @ f
            scale = np.abs(M @ a).max() + np.abs(J.T @ f).max()
            assert np.abs(res).max() < 1e-7 * scale
            assert (f >= 0).all()
            np.testing.assert_allclose(probe.view("efc_force")[:n], f, rtol=1e-9, atol=1e-12)
            checked += 1
    assert checked > 800
