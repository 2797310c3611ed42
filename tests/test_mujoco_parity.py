"""Parity against a real MuJoCo -- runs wherever ``import mujoco`` works, skips elsewhere (it is skipped in the build
container and on the GPU box of this project: neither has MuJoCo).

The reference model cannot be loaded by a current MuJoCo (``coordinate="global"`` was removed in 2.3.4), so the tests
load ``jitterbug_b200/assets/jitterbug_local.xml`` -- the same model exported in local coordinates by
``mjcf.to_local_xml`` (the CPU test ``test_mjcf.py::test_local_export_round_trip`` pins the export to the compiled
model).  What a failure would mean, test by test:

* ``test_compiled_model``: the MJCF-subset compiler (geom inertias, global->local frames, mj_setConst) disagrees with
  MuJoCo's compiler -> fix ``jitterbug_b200/mjcf.py`` first, everything else depends on it;
* ``test_forward_contacts``: collision conventions (mjc_Plane* distance / point, contact ordering) differ -> fix
  ``oracle/jb_oracle.c`` and the same routine in ``csrc/jb_quad.cuh``;
* ``test_substep_from_identical_state``: forward dynamics / constraint solver / Euler integration differ -> the oracle
  (the single place MuJoCo semantics live) is wrong; the committed golden substeps then need regenerating;
* ``test_cuda_substep_vs_mujoco`` (``-m gpu``): CUDA vs MuJoCo directly, at the north_star tolerances.

Version caveats (the reference pins MuJoCo 2.0 through dm_control 09b27a0): current MuJoCo keeps the Euler
implicit-in-damping update behind ``eulerdamp`` (on by default) and the same pyramidal-cone regularisation; known
semantic changes since 2.0 that matter here are listed in DESIGN.md section 2.
"""
import os

import numpy as np
import pytest

mujoco = pytest.importorskip("mujoco")

from jitterbug_b200 import mjcf  # noqa: E402
from oracle import oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
LOCAL_XML = os.path.join(HERE, "..", "jitterbug_b200", "assets", "jitterbug_local.xml")


@pytest.fixture(scope="module")
def mj():
    m = mujoco.MjModel.from_xml_path(LOCAL_XML)
    return m, mujoco.MjData(m)


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(HERE, "golden", "substeps.npz"))


def _inertia_in_body_frame(ipos, iquat, mass, diag):
    r = mjcf.quat2mat(iquat)
    return r @ np.diag(diag) @ r.T, ipos, mass


def test_compiled_model(mj, compiled_model):
    m, _ = mj
    cm = compiled_model
    assert (m.nq, m.nv, m.nu, m.nbody, m.njnt, m.ngeom) == (cm.nq, cm.nv, cm.nu, cm.nbody, cm.njnt, cm.ngeom)
    np.testing.assert_allclose(m.body_mass, cm.body_mass, rtol=1e-9, atol=1e-15)
    np.testing.assert_allclose(m.body_pos, cm.body_pos, atol=1e-12)
    np.testing.assert_allclose(m.jnt_pos, cm.jnt_pos, atol=1e-12)
    np.testing.assert_allclose(m.jnt_axis, cm.jnt_axis, atol=1e-12)
    np.testing.assert_allclose(m.qpos0, cm.qpos0, atol=1e-12)
    np.testing.assert_allclose(m.geom_size[:, :3], cm.geom_size, atol=1e-12)
    np.testing.assert_allclose(m.geom_pos, cm.geom_pos, atol=1e-12)
    for b in range(1, cm.nbody):
        # principal axes are only defined up to sign / order: compare the tensors
        ia, pa, _ = _inertia_in_body_frame(m.body_ipos[b], m.body_iquat[b], m.body_mass[b], m.body_inertia[b])
        ib, pb, _ = _inertia_in_body_frame(cm.body_ipos[b], cm.body_iquat[b], cm.body_mass[b], cm.body_inertia[b])
        np.testing.assert_allclose(pa, pb, atol=1e-10)
        np.testing.assert_allclose(ia, ib, rtol=1e-7, atol=1e-16)
    np.testing.assert_allclose(m.body_invweight0, cm.body_invweight0, rtol=1e-6)
    np.testing.assert_allclose(m.dof_invweight0, cm.dof_invweight0, rtol=1e-6)
    np.testing.assert_allclose(m.stat.meaninertia, cm.meaninertia, rtol=1e-6)
    assert m.opt.timestep == cm.timestep


def _set(m, d, qpos, qvel, warm, ctrl):
    mujoco.mj_resetData(m, d)
    d.qpos[:], d.qvel[:], d.qacc_warmstart[:], d.ctrl[:] = qpos, qvel, warm, ctrl


def _contacts(m, d):
    floor = mujoco.mj_name2id(m, mujoco.mjtObj.mjOBJ_GEOM, "floor")
    out = []
    for i in range(d.ncon):
        c = d.contact[i]
        if c.dist < 0:
            out.append((int(c.geom2 if c.geom1 == floor else c.geom1), float(c.dist)))
    return sorted(out)


def test_forward_contacts(mj, gold, compiled_model):
    m, d = mj
    sim = O.OracleSim(compiled_model)
    for i in range(len(gold["ctrl"])):
        _set(m, d, gold["qpos"][i], gold["qvel"][i], gold["warm"][i], gold["ctrl"][i])
        mujoco.mj_forward(m, d)
        got = _contacts(m, d)
        want = [int(c) for c in gold["contacts"][i] if c >= 0]
        assert [g for g, _ in got] == want, i
        sim.set_state(gold["qpos"][i], gold["qvel"][i], gold["warm"][i], ctrl=gold["ctrl"][i])
        sim.step1()
        _, dist, _, _ = sim.contacts()
        np.testing.assert_allclose(sorted(x for _, x in got), sorted(dist), atol=1e-12)
        # no self-collisions in these states (expected for every reachable state)
        assert d.ncon == sum(1 for k in range(d.ncon) if 0 in (d.contact[k].geom1, d.contact[k].geom2))


def test_substep_from_identical_state(mj, gold):
    """north_star: single-step results from identical qpos/qvel/ctrl; here MuJoCo vs the fp64 oracle's golden outputs."""
    m, d = mj
    for i in range(len(gold["ctrl"])):
        _set(m, d, gold["qpos"][i], gold["qvel"][i], gold["warm"][i], gold["ctrl"][i])
        mujoco.mj_step(m, d)
        np.testing.assert_allclose(d.qpos, gold["sub_qpos"][i], rtol=0, atol=1e-9, err_msg=f"qpos {i}")
        sc = np.maximum(np.abs(gold["sub_qvel"][i]), 1.0)
        np.testing.assert_allclose(d.qvel / sc, gold["sub_qvel"][i] / sc, rtol=0, atol=1e-6, err_msg=f"qvel {i}")


def test_control_step_from_identical_state(mj, gold):
    m, d = mj
    for i in range(0, len(gold["ctrl"]), 4):
        _set(m, d, gold["qpos"][i], gold["qvel"][i], gold["warm"][i], gold["ctrl"][i])
        for _ in range(50):
            mujoco.mj_step(m, d)
        np.testing.assert_allclose(d.qpos, gold["ctl_qpos"][i], rtol=0, atol=1e-5)


@pytest.mark.gpu
def test_cuda_substep_vs_mujoco(mj, gold):
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from jitterbug_b200.batched_env import BatchedJitterbugEnv
    m, d = mj
    n = len(gold["ctrl"])
    env = BatchedJitterbugEnv("move_from_origin", num_envs=n, control_timestep=0.0002, time_limit=float("inf"))
    env.reset()
    env.set_state(gold["qpos"], gold["qvel"], gold["warm"], step_count=np.ones(n))
    ncon, geom, _ = [x.cpu().numpy() for x in env.contacts()]
    env.step(torch.as_tensor(gold["ctrl"], dtype=torch.float32, device=env.device))
    q, v, _ = [x.cpu().numpy().astype(np.float64) for x in env.get_state()]
    for i in range(n):
        qf, vf, wf = [gold[k][i].astype(np.float32).astype(np.float64) for k in ("qpos", "qvel", "warm")]
        _set(m, d, qf, vf, wf, float(np.float32(gold["ctrl"][i])))
        mujoco.mj_forward(m, d)
        cons = _contacts(m, d)
        if any(abs(x) < 1e-6 for _, x in cons):
            continue                                   # knife-edge activation in fp32
        assert [g for g, _ in cons] == [int(c) for c in geom[i, :ncon[i]]]
        mujoco.mj_step(m, d)
        assert (np.abs(q[i] - d.qpos) / np.maximum(np.abs(d.qpos), 1)).max() < 1e-5
        assert (np.abs(v[i] - d.qvel) / np.maximum(np.abs(d.qvel).max(), 1)).max() < 1e-4
