"""On-device model compilation for augmented Jitterbugs: the kernel's algorithm
(csrc/jb_augment.cuh; host build on CPU, the CUDA kernel itself under -m gpu) against the reference route -- perturb the
MJCF text like jitterbug_dmc/augmented_jitterbug.py:95-267, recompile, flatten -- fed the same perturbation vector."""
import ctypes as C

import numpy as np
import pytest

import hostsim
from jitterbug_b200 import augmented_jitterbug as aug
from jitterbug_b200 import mjcf
from jitterbug_b200.model import JbModelD, fill_struct, flatten_model

ALL = dict(modify_legs=True, modify_mass=True, modify_coreBody1=True, modify_coreBody2=True, modify_global_density=True,
           modify_gear=True)


class ScriptedRandom:
    """Stands in for numpy's RNG inside augment_Jitterbug: hands back the prescribed perturbations in draw order."""

    def __init__(self, values):
        self.values = list(values)

    def normal(self, mean, sd):
        return self.values.pop(0)


def draw(rng, n):
    """Perturbation vectors with the reference's distributions (all groups on), layout include/jbsim.h JB_AUG_NDELTA."""
    d = np.zeros((n, 31))
    d[:, 0] = rng.normal(0, 200., n)
    d[:, 1] = rng.normal(0, 10., n)
    d[:, 2] = rng.normal(0, 80., n)
    d[:, 3:27] = rng.normal(0, 1, (n, 24)) * np.tile([0.003, 0.003, 0.002], 8)
    d[:, 27:30] = rng.normal(0, 1, (n, 3)) * [0.0015, 0.002, 0.001]
    d[:, 28:30] = np.clip(d[:, 28:30], -0.001, 1.0)
    d[:, 30] = rng.normal(0, 0.001, n)
    return d


def host_model(delta):
    xml = aug.augment_Jitterbug(random=ScriptedRandom(delta), **ALL)
    return fill_struct(flatten_model(mjcf.compile_xml(xml)), double=True)


def struct_values(m):
    out = {}
    for name, _ in JbModelD._fields_:
        if name == "pad_":
            continue
        v = getattr(m, name)
        if name.startswith("g_"):
            gs = list(v) if hasattr(v, "__len__") else [v]
            for i, g in enumerate(gs):
                for f in ("center", "rot", "size", "rbound"):
                    out[f"{name}[{i}].{f}"] = np.array(getattr(g, f), dtype=np.float64).ravel()
                out[f"{name}[{i}].geom_id"] = np.array([g.geom_id], dtype=np.float64)
        else:
            out[name] = np.array(v, dtype=np.float64).ravel()
    return out


def assert_models_close(got, want, rtol, atol_scale=1e-12):
    a, b = struct_values(got), struct_values(want)
    for k in b:
        scale = max(np.abs(b[k]).max(), 1e-300)
        err = np.abs(a[k] - b[k]).max()
        assert err <= rtol * scale + atol_scale * scale, (k, err, scale, a[k], b[k])


def test_zero_perturbation_of_unaugmented_groups_keeps_the_base_model():
    """Densities / gear untouched, rotor offset zero: root body and rotor must come out as the host compiler made them
    (checks geom inertias, COM / inertia assembly and mj_setConst); legs differ only through the recomputed axes."""
    base = aug.make_aug_base()
    _, _, md = hostsim.augment(base, np.zeros((1, 31)))
    want = host_model(np.zeros(31))
    assert_models_close(md[0], want, rtol=1e-9)


def test_host_build_matches_xml_route():
    base = aug.make_aug_base()
    deltas = draw(np.random.RandomState(0), 12)
    packed, mf, md = hostsim.augment(base, deltas)
    for i in range(len(deltas)):
        want = host_model(deltas[i])
        assert_models_close(md[i], want, rtol=2e-9)
        want_packed = hostsim.pack(fill_struct(flatten_model(mjcf.compile_xml(
            aug.augment_Jitterbug(random=ScriptedRandom(deltas[i]), **ALL)))), double=False)
        np.testing.assert_allclose(packed[i], want_packed, rtol=3e-7, atol=1e-14)
    # the perturbation really changes the dynamics-relevant quantities
    assert abs(md[0].mu[0] - base.model.mu[0]) > 1e-8 and abs(md[0].invw[1] - base.model.invw[1]) > 1e-3 * base.model.invw[1]


def test_augmented_models_simulate_like_their_xml_twins():
    """A few sub-steps of the fp64 algorithm on the device-route model and on the XML-route model agree."""
    base = aug.make_aug_base()
    deltas = draw(np.random.RandomState(1), 2)
    _, _, md = hostsim.augment(base, deltas)
    for i in range(2):
        want = host_model(deltas[i])
        q = np.zeros(16); q[2] = 0.035; q[3] = 1.0
        v, w = np.zeros(15), np.zeros(15)
        qa, va, _, _, _ = hostsim.substeps(md[i], q, v, w, 0.7, nsub=100)
        qb, vb, _, _, _ = hostsim.substeps(want, q, v, w, 0.7, nsub=100)
        assert np.abs(qa - qb).max() < 1e-9 and np.abs(va - vb).max() < 1e-7


@pytest.mark.gpu
def test_device_compile_matches_host_build_and_xml_route():
    import torch
    base = aug.make_aug_base()
    deltas = draw(np.random.RandomState(2), 300)
    packed_d, structs = aug.compile_models_device(torch.from_numpy(deltas).cuda(), base, return_structs=True)
    packed_h, mf, _ = hostsim.augment(base, deltas)
    np.testing.assert_allclose(packed_d.cpu().numpy(), packed_h, rtol=3e-7, atol=1e-14)   # fma contraction only
    for i in (0, 7, 299):
        want = fill_struct(flatten_model(mjcf.compile_xml(aug.augment_Jitterbug(random=ScriptedRandom(deltas[i]), **ALL))))
        a, b = struct_values(structs[i]), struct_values(want)
        for k in b:
            assert np.abs(a[k] - b[k]).max() <= 3e-7 * np.abs(b[k]).max() + 1e-14, k


@pytest.mark.gpu
def test_device_sampler_distribution_and_sharding():
    import torch
    n = 200000
    d = aug.sample_deltas(n, seed=3).cpu().numpy()
    assert np.all(d[:, :3] == 0) and np.all(d[:, 30] == 0)                   # groups that are off stay untouched
    sd = np.tile([0.003, 0.003, 0.002], 8)
    assert np.abs(d[:, 3:27].mean(0)).max() < 3e-5 and np.abs(d[:, 3:27].std(0) / sd - 1).max() < 0.01
    assert abs(d[:, 27].std() / 0.0015 - 1) < 0.01
    assert d[:, 28].min() == -0.001 and d[:, 29].min() == -0.001 and d[:, 28].max() > 0.004   # clip(-0.001, 1)
    assert np.abs(np.corrcoef(d[:, 3:28].T) - np.eye(25)).max() < 0.02
    full = aug.sample_deltas(1000, seed=3, **dict(modify_gear=True, modify_global_density=True))
    shard = aug.sample_deltas(400, seed=3, env_id_offset=600, **dict(modify_gear=True, modify_global_density=True))
    assert torch.equal(full[600:], shard)
    assert float(full[:, 30].std()) > 0 and float(full[:, 0].std()) > 50


@pytest.mark.gpu
def test_env_on_device_compiled_models_matches_env_on_host_compiled_models():
    import torch
    from jitterbug_b200 import BatchedJitterbugEnv
    n = 64
    deltas = aug.sample_deltas(n, seed=11)
    packed = aug.compile_models_device(deltas)
    dh = deltas.cpu().numpy()
    kw = dict(modify_legs=True, modify_mass=True)
    models = []
    for i in range(n):
        vals = list(dh[i, 3:30])
        models.append(mjcf.compile_xml(aug.augment_Jitterbug(random=ScriptedRandom(vals), **kw)))
    a = BatchedJitterbugEnv(task="move_in_direction", num_envs=n, seed=4, packed_models=packed)
    b = BatchedJitterbugEnv(task="move_in_direction", num_envs=n, seed=4, models=models)
    oa, ob = a.reset(), b.reset()
    assert torch.equal(oa, ob)
    act = torch.linspace(-1, 1, n, device="cuda")
    for _ in range(3):
        oa, ra, _ = a.step(act)
        ob, rb, _ = b.step(act)
    # identical up to the last-bit differences of the two model builds, amplified by 150 sub-steps of contact dynamics
    assert (oa - ob).abs().max() < 2e-3 and (ra - rb).abs().max() < 2e-3
    env = aug.make_batched_env("move_from_origin", num_envs=256, seed=1)
    o = env.reset()
    o, r, st = env.step(torch.zeros(256, device="cuda"))
    assert torch.isfinite(o).all() and torch.isfinite(r).all()
