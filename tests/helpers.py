"""Shared helpers for the test-suite (numpy, fp64)."""
import numpy as np

from jitterbug_b200 import mjcf


def integrate_pos(cm, qpos, qvel, h):
    """mj_integratePos for the free + hinge model (exact for constant generalized velocity)."""
    q = np.array(qpos, dtype=np.float64)
    q[0:3] += h * qvel[0:3]
    w = np.asarray(qvel[3:6], dtype=np.float64)
    n = np.linalg.norm(w)
    if n > 0:
        dq = mjcf.axis_angle_quat(w / n, h * n)
        q[3:7] = mjcf.quat_mul(q[3:7] / np.linalg.norm(q[3:7]), dq)
    q[7:] += h * qvel[6:]
    return q


def numeric_bias(cm, qpos, qvel, eps=1e-4):
    """qfrc_bias from finite differences of forward kinematics only (Newton-Euler per body, qacc = 0)."""
    nv = cm.nv

    def body_state(t):
        q = integrate_pos(cm, qpos, qvel, t)
        fk = mjcf.forward_kinematics(cm, q)
        com, L = [], []
        for b in range(cm.nbody):
            jp, jr = mjcf.jacobian(cm, fk, fk["xipos"][b], b)
            w = jr @ qvel
            iw = fk["ximat"][b] @ np.diag(cm.body_inertia[b]) @ fk["ximat"][b].T
            com.append(fk["xipos"][b].copy())
            L.append(iw @ w)
        return fk, np.array(com), np.array(L)

    fk0, c0, _ = body_state(0.0)
    _, cp, Lp = body_state(eps)
    _, cn, Ln = body_state(-eps)
    acc = (cp - 2 * c0 + cn) / eps ** 2
    dL = (Lp - Ln) / (2 * eps)
    bias = np.zeros(nv)
    for b in range(1, cm.nbody):
        if cm.body_dofadr[b] < 0 and cm.body_parentid[b] == 0:
            continue    # static body
        jp, jr = mjcf.jacobian(cm, fk0, fk0["xipos"][b], b)
        bias += jp.T @ (cm.body_mass[b] * (acc[b] - cm.gravity)) + jr.T @ dL[b]
    return bias


def random_state(cm, rng, vel_scale=1.0, height=0.2):
    """A random airborne-ish state with large joint excursions and velocities."""
    q = cm.qpos0.copy()
    q[0:2] = rng.uniform(-1, 1, 2)
    q[2] = height
    quat = rng.normal(size=4)
    q[3:7] = quat / np.linalg.norm(quat)
    q[7:15] = rng.uniform(-0.3, 0.3, 8)
    q[15] = rng.uniform(-np.pi, np.pi)
    v = np.zeros(cm.nv)
    v[0:3] = rng.uniform(-1, 1, 3) * vel_scale
    v[3:6] = rng.uniform(-20, 20, 3) * vel_scale
    v[6:14] = rng.uniform(-30, 30, 8) * vel_scale
    v[14] = rng.uniform(-150, 150) * vel_scale
    return q, v
