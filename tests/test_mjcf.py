"""MJCF-subset compiler: sizes, masses and tree bookkeeping against the hand-derived values of SURVEY.md App. A."""
import os

import numpy as np
import pytest

from jitterbug_b200 import mjcf
from jitterbug_b200.model import flatten_model, fill_struct

REF_XML = "/root/reference/jitterbug_dmc/jitterbug.xml"


def test_sizes_and_names(compiled_model):
    m = compiled_model
    assert (m.nq, m.nv, m.nu, m.nbody, m.njnt, m.ngeom) == (16, 15, 1, 12, 10, 25)
    assert m.body_names[:2] == ["world", "jitterbug"] and m.body_names[10:] == ["mass", "target"]
    assert m.body_names[2:10] == ["leg2upper", "leg2lower", "leg3upper", "leg3lower",
                                  "leg1upper", "leg1lower", "leg4upper", "leg4lower"]
    assert m.timestep == 0.0002
    np.testing.assert_array_equal(m.qpos0, [0, 0, .035, 1, 0, 0, 0] + [0] * 9)
    # sparse-M pattern: nM = 88
    nM = 0
    for d in range(m.nv):
        j = d
        while j >= 0:
            nM += 1
            j = m.dof_parentid[j]
    assert nM == 88
    np.testing.assert_array_equal(m.dof_parentid, [-1, 0, 1, 2, 3, 4, 5, 6, 5, 8, 5, 10, 5, 12, 5])


def test_masses_match_hand_derivation(compiled_model):
    m = compiled_model
    # SURVEY.md App. A.3 (independent hand computation)
    assert m.body_mass[1:11].sum() == pytest.approx(1.75165e-2, rel=1e-5)
    assert m.body_mass[1] == pytest.approx(4.9930e-3, rel=1e-4)
    assert m.body_mass[2] == pytest.approx(4.2354e-4, rel=1e-4)
    assert m.body_mass[6] == pytest.approx(3.9760e-4, rel=1e-4)
    assert m.body_mass[3] == pytest.approx(7.0949e-4, rel=1e-4)
    assert m.body_mass[10] == pytest.approx(8.0433e-3, rel=1e-4)
    g = {n: i for i, n in enumerate(m.geom_names) if n}
    assert m.geom_mass[g["coreBody1"]] == pytest.approx(1.1180e-4, rel=1e-4)
    assert m.geom_mass[g["coreBody2"]] == pytest.approx(2.3654e-3, rel=1e-4)
    assert m.geom_mass[g["screw1"]] == pytest.approx(1.1611e-3, rel=1e-4)
    assert m.geom_mass[g["screw2"]] == pytest.approx(1.3547e-3, rel=1e-4)
    assert m.geom_mass[g["foot1"]] == pytest.approx(1.2441e-4, rel=1e-4)
    assert m.geom_mass[g["mass"]] == pytest.approx(7.7409e-3, rel=1e-4)
    fk = mjcf.forward_kinematics(m, m.qpos0)
    np.testing.assert_allclose(fk["xipos"][1], [0, .00601, .02394], atol=2e-5)
    np.testing.assert_allclose(fk["xipos"][10], [.00414, -.004, .06082], atol=2e-5)
    # feet: front feet start 2 mm above the floor, back feet touch it
    np.testing.assert_allclose(fk["geom_xpos"][[8, 12, 16, 20], 2], [.005, .005, .003, .003], atol=1e-15)


def test_global_to_local_roundtrip(compiled_model):
    """Localised model reproduces the global-coordinate XML numbers at qpos0."""
    m = compiled_model
    fk = mjcf.forward_kinematics(m, m.qpos0)
    j = m.jnt_id("joint2Leg2")
    np.testing.assert_allclose(fk["xanchor"][j], [.0287, -.0308, .068], atol=1e-15)
    ax = np.array([-2.32, -3.68, 0.0])
    np.testing.assert_allclose(fk["xaxis"][j], ax / np.linalg.norm(ax), atol=1e-14)
    np.testing.assert_allclose(fk["xanchor"][m.jnt_id("jointMass")], [0, -.004, .05], atol=1e-15)
    np.testing.assert_allclose(fk["geom_xpos"][m.geom_id("mass")], [.0043, -.004, .061], atol=1e-15)
    np.testing.assert_allclose(fk["geom_xpos"][m.geom_id("targetPointer")], [.01, 0, .035], atol=1e-15)


def test_invweight_and_actuator(compiled_model):
    m = compiled_model
    assert m.body_invweight0[0].tolist() == [0, 0] and m.body_invweight0[11].tolist() == [0, 0]
    assert (m.body_invweight0[1:11] > 0).all()
    # a free body's translational inverse weight is bounded below by 1/total mass
    assert m.body_invweight0[1, 0] > 1 / m.body_mass[1:11].sum()
    assert m.actuator_gear[0] == 0.00833 and m.actuator_biasprm[0, 2] == -0.8
    np.testing.assert_array_equal(m.actuator_ctrlrange[0], [-1, 1])


@pytest.mark.skipif(not os.path.exists(REF_XML), reason="reference checkout not present (GPU box)")
def test_packaged_xml_equals_reference_xml(compiled_model):
    """The packaged physics-only XML compiles to exactly the model of the reference's jitterbug.xml."""
    ref = mjcf.compile_xml(path=REF_XML)
    for f in ("body_mass", "body_pos", "body_quat", "body_inertia", "body_ipos", "jnt_pos", "jnt_axis", "geom_pos",
              "geom_size", "geom_quat", "body_invweight0", "dof_damping", "jnt_stiffness", "geom_contype"):
        np.testing.assert_array_equal(getattr(compiled_model, f), getattr(ref, f), err_msg=f)


def test_flatten_model_structure(compiled_model):
    flat = flatten_model(compiled_model)
    assert flat["mu"].shape == (4,) and flat["p1"].shape == (4, 3)
    total = flat["m0"] + flat["mu"].sum() + flat["ml"].sum() + flat["mm"]
    assert total == pytest.approx(compiled_model.body_mass[1:11].sum(), rel=1e-14)
    # both hinges of a leg are parallel in the base model
    for k in range(4):
        np.testing.assert_allclose(flat["a1"][k], flat["a2"][k], atol=1e-14)
    assert [g["geom_id"] for g in flat["g_foot"]] == [8, 12, 16, 20]
    assert [g["geom_id"] for g in flat["g_lcyl"]] == [7, 11, 15, 19]
    s = fill_struct(flat)
    assert s.gear == pytest.approx(0.00833) and s.timestep == pytest.approx(0.0002)
    # contact reference: K = 1/(dmax^2 tc^2), B = 2/(dmax tc) with tc = .02, dmax = .95
    assert flat["solref_K"] == pytest.approx(1 / (0.95 ** 2 * 0.02 ** 2))
    assert flat["solref_B"] == pytest.approx(2 / (0.95 * 0.02))


def test_local_export_round_trip(compiled_model):
    """mjcf.to_local_xml: the packaged local-coordinate asset is the export of the compiled model, and compiling the
    export again gives the same model (frames composed the other way round) and the same oracle dynamics."""
    import os
    from oracle import oracle as O
    cm = compiled_model
    xml = mjcf.to_local_xml(cm)
    asset = os.path.join(os.path.dirname(mjcf.DEFAULT_XML), "jitterbug_local.xml")
    assert open(asset).read() == xml
    assert 'coordinate="global"' not in xml and "fromto" not in xml
    cm2 = mjcf.compile_xml(xml)
    for f in ("body_mass", "body_pos", "body_ipos", "geom_pos", "geom_size", "jnt_pos", "jnt_axis", "qpos0", "dof_damping",
              "jnt_stiffness", "actuator_gear", "actuator_biasprm", "geom_mass"):
        np.testing.assert_allclose(np.asarray(getattr(cm2, f)), np.asarray(getattr(cm, f)), rtol=0, atol=1e-15, err_msg=f)
    np.testing.assert_allclose(cm2.geom_density[1:], cm.geom_density[1:], rtol=0, atol=0)      # [0] is the massless floor
    np.testing.assert_allclose(np.abs(cm2.body_quat), np.abs(cm.body_quat), atol=1e-15)
    np.testing.assert_allclose(cm2.body_invweight0, cm.body_invweight0, rtol=1e-10)
    np.testing.assert_allclose(cm2.dof_invweight0, cm.dof_invweight0, rtol=1e-10)
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "substeps.npz"))
    sim = O.OracleSim(cm2)
    for i in range(0, len(g["ctrl"]), 5):
        sim.set_state(g["qpos"][i], g["qvel"][i], g["warm"][i], ctrl=g["ctrl"][i])
        sim.step1(); sim.step2()
        np.testing.assert_allclose(sim.qpos, g["sub_qpos"][i], rtol=0, atol=1e-12)
        np.testing.assert_allclose(sim.qvel, g["sub_qvel"][i], rtol=1e-9, atol=1e-9)
