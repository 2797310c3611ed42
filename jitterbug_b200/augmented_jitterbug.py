"""Augmented Jitterbugs: Gaussian perturbations of the model geometry (domain randomisation).

Drop-in for ``jitterbug_dmc/augmented_jitterbug.py``: ``augment_Jitterbug(...)`` has the reference's keyword arguments
and distributions (augmented_jitterbug.py:95-267) and the module carries the same five tagged factories (:313-447).
Differences, all deliberate:

* the reference at HEAD writes ``augmented_jitterbug.xml`` but its factories still load the un-augmented model; here
  the factories simulate the augmented model (the README's intent).  ``augment_Jitterbug`` returns the XML text and
  only writes a file when ``new_file`` is given;
* the ``"mass_geom"`` size/position branch is dead in the reference (the geom is called ``"mass"``), so the rotor
  ellipsoid is left untouched, as the reference's code would leave it;
* hinge axes of perturbed legs are recomputed as ``cross(v_upper, z)`` / ``cross(v_upper, v_lower)`` exactly like the
  reference, which differs from the hand-tuned axes of jitterbug.xml even at zero noise.

For batched simulation there are two routes to one model per env (staged per CTA in shared memory by the kernel):
:func:`make_models` compiles a population on the host (XML -> compile, like the reference) for
``BatchedJitterbugEnv(models=...)``; :func:`sample_deltas` + :func:`compile_models_device` draw the same perturbations
and redo the dependent part of the compile in one CUDA kernel, one thread per env (``jb_augment_compile``), for
``BatchedJitterbugEnv(packed_models=...)`` -- or simply ``make_batched_env(...)``.  1 Mi models take a fraction of a
second on the device against hours of host compiles.
"""

import xml.etree.ElementTree as ET

import numpy as np

from . import jitterbug as _jb
from . import mjcf
from .batched_env import DEFAULT_CONTROL_TIMESTEP, DEFAULT_TIME_LIMIT
from .dm import TaggedTasks

SUITE = TaggedTasks()
TARGET_SPEED = 0.1


def _vec(s):
    return np.array([float(x) for x in s.split()])


def _fmt(a):
    return " ".join(repr(float(x)) for x in np.atleast_1d(a))


def _unit(v):
    return v / np.linalg.norm(v)


def augment_Jitterbug(modify_legs=False, sd_legs=np.array([0.003, 0.003, 0.002]),
                      modify_mass=False, sd_mass_pos=np.array([0.0015, 0.002, 0.001]),
                      sd_mass_size=np.array([0.0007, 0.0007, 0.0007]),
                      modify_coreBody1=False, sd_coreBody1_density=10.,
                      modify_coreBody2=False, sd_coreBody2_density=80.,
                      modify_global_density=False, sd_global_density=200.,
                      modify_gear=False, sd_gear=0.001,
                      original_file=None, new_file=None, random=None):
    """Create an augmented Jitterbug by adding Gaussian noise to key parameters of the model.

    Returns the augmented MJCF text.  ``random``: a ``numpy.random.RandomState`` / seed (the reference draws from the
    global numpy RNG; pass ``None`` for that behaviour).
    """
    rng = np.random if random is None else (random if hasattr(random, "normal") else np.random.RandomState(random))
    root = ET.fromstring(open(original_file or mjcf.DEFAULT_XML, "rb").read())
    if modify_global_density:
        for deflt in root.findall('default'):
            for geom in deflt.findall('geom'):
                geom.attrib['density'] = _fmt(_vec(geom.attrib['density']) + rng.normal(0, sd_global_density))
    body = next(b for b in root.find('worldbody').findall('body') if b.attrib.get('name') == 'jitterbug')
    for geom in body.findall('geom'):
        if geom.attrib.get('name') == "coreBody1" and modify_coreBody1:
            geom.attrib['density'] = _fmt(_vec(geom.attrib['density']) + rng.normal(0, sd_coreBody1_density))
        elif geom.attrib.get('name') == "coreBody2" and modify_coreBody2:
            geom.attrib['density'] = _fmt(_vec(geom.attrib['density']) + rng.normal(0, sd_coreBody2_density))
    leg_names = ["leg{}upper".format(i) for i in range(1, 5)]
    for child in body.findall('body'):
        if child.attrib.get('name') in leg_names and modify_legs:
            d_u = np.array([rng.normal(0, sd_legs[0]), rng.normal(0, sd_legs[1]), rng.normal(0, sd_legs[2])])
            v_u = None
            for geom in child.findall('geom'):
                if geom.attrib['type'] == "cylinder":
                    ft = _vec(geom.attrib['fromto']) + np.concatenate([d_u, np.zeros(3)])
                    geom.attrib['fromto'] = _fmt(ft)
                    v_u = _unit(ft[3:] - ft[:3])
                elif geom.attrib['type'] == "sphere":
                    geom.attrib['pos'] = _fmt(_vec(geom.attrib['pos']) + d_u)
            for joint in child.findall('joint'):
                joint.attrib['axis'] = _fmt(_unit(np.cross(v_u, [0., 0., 1.])))
            for lower in child.findall('body'):
                d_l = np.array([rng.normal(0, sd_legs[0]), rng.normal(0, sd_legs[1]), rng.normal(0, sd_legs[2])])
                v_l = None
                for geom in lower.findall('geom'):
                    if geom.attrib['type'] == "cylinder":
                        ft = _vec(geom.attrib['fromto']) + np.concatenate([d_l, d_u])
                        geom.attrib['fromto'] = _fmt(ft)
                        v_l = _unit(ft[3:] - ft[:3])
                    elif geom.attrib['type'] == "sphere":
                        geom.attrib['pos'] = _fmt(_vec(geom.attrib['pos']) + d_l)
                for joint in lower.findall('joint'):
                    joint.attrib['pos'] = _fmt(_vec(joint.attrib['pos']) + d_u)
                    joint.attrib['axis'] = _fmt(_unit(np.cross(v_u, v_l)))
        elif child.attrib.get('name') == "mass" and modify_mass:
            dp = np.array([rng.normal(0, sd_mass_pos[0]),
                           np.clip(rng.normal(0, sd_mass_pos[1]), -0.001, 1.),
                           np.clip(rng.normal(0, sd_mass_pos[2]), -0.001, 1.)])
            for geom in child.iter('geom'):
                if geom.attrib.get('name') == "threadMass":
                    geom.attrib['fromto'] = _fmt(_vec(geom.attrib['fromto']) + np.array([dp[0], dp[1], 0., *dp]))
                # the reference's "mass_geom" branch never matches (the geom is named "mass"): ellipsoid unchanged
            for joint in child.iter('joint'):
                joint.attrib['pos'] = _fmt(_vec(joint.attrib['pos']) + dp)
    if modify_gear:
        for ac in root.findall('actuator'):
            for ger in ac.findall('general'):
                ger.attrib['gear'] = _fmt(_vec(ger.attrib['gear']) + rng.normal(0, sd_gear))
    xml = ET.tostring(root)
    if new_file:
        with open(new_file, "wb") as f:
            f.write(xml)
    return xml


def make_models(n, seed=0, modify_legs=True, modify_mass=True, **kwargs):
    """Compile ``n`` augmented models (defaults = what the reference's training script turns on, benchmark.py:129-136)."""
    rng = np.random.RandomState(seed)
    return [mjcf.compile_xml(augment_Jitterbug(modify_legs=modify_legs, modify_mass=modify_mass, random=rng, **kwargs))
            for _ in range(n)]


# ---------------------------------------------------------------------------------------------- device-side compile
# order of the perturbation vector (JB_AUG_NDELTA = 31): the order augment_Jitterbug draws its Gaussians in
DELTA_LAYOUT = {"global_density": 0, "coreBody1_density": 1, "coreBody2_density": 2, "legs": slice(3, 27),
                "mass_pos": slice(27, 30), "gear": 30}


def make_aug_base(cm=None):
    """``JbAugBase`` for the C-ABI: the unperturbed model in the device layout (fp64) + per-geom densities."""
    from . import _native
    from .model import fill_struct, flatten_model
    cm = mjcf.compile_xml() if cm is None else cm
    flat = flatten_model(cm)
    base = _native.JbAugBase()
    base.model = fill_struct(flat, double=True)
    slots = [flat["g_box"][0], flat["g_box"][1], flat["g_rcyl"], flat["g_rell"]]
    for k in range(4):
        slots += [flat["g_ucyl"][k], flat["g_knee"][k], flat["g_lcyl"][k], flat["g_foot"][k]]
    slots += [flat["g_mcyl"], flat["g_mell"]]
    for i, g in enumerate(slots):
        base.density[i] = float(cm.geom_density[g["geom_id"]])
        base.follows_default[i] = int(cm.geom_density_default[g["geom_id"]])
    return base


def _spec(modify_legs=True, sd_legs=(0.003, 0.003, 0.002), modify_mass=True, sd_mass_pos=(0.0015, 0.002, 0.001),
          modify_coreBody1=False, sd_coreBody1_density=10., modify_coreBody2=False, sd_coreBody2_density=80.,
          modify_global_density=False, sd_global_density=200., modify_gear=False, sd_gear=0.001):
    from . import _native
    s = _native.JbAugSpec()
    s.modify_legs, s.modify_mass, s.modify_core1, s.modify_core2 = map(int, (modify_legs, modify_mass, modify_coreBody1,
                                                                             modify_coreBody2))
    s.modify_global_density, s.modify_gear = int(modify_global_density), int(modify_gear)
    for i in range(3):
        s.sd_legs[i], s.sd_mass_pos[i] = float(sd_legs[i]), float(sd_mass_pos[i])
    s.sd_core1_density, s.sd_core2_density = float(sd_coreBody1_density), float(sd_coreBody2_density)
    s.sd_global_density, s.sd_gear = float(sd_global_density), float(sd_gear)
    return s


def sample_deltas(n, seed=0, device="cuda", env_id_offset=0, **augment_kwargs):
    """Draw ``n`` perturbation vectors on the device -> float64 ``[n, 31]`` (see ``DELTA_LAYOUT``).  Keyword arguments
    and distributions are ``augment_Jitterbug``'s; the defaults are what the reference's training script turns on
    (legs + mass, benchmark.py:129-136).  Streams are keyed by ``env_id_offset + env`` (sharding-invariant)."""
    import ctypes as C

    import torch

    from . import _native
    lib = _native.load()
    dev = torch.device(device)
    out = torch.empty((n, _native.AUG_NDELTA), dtype=torch.float64, device=dev)
    spec = _spec(**augment_kwargs)
    with torch.cuda.device(dev):
        _native.check(lib.jb_augment_sample(C.byref(spec), int(seed) & (2 ** 64 - 1), int(env_id_offset), n,
                                            out.data_ptr(), torch.cuda.current_stream(dev).cuda_stream))
    return out


def compile_models_device(deltas, base=None, return_structs=False):
    """``deltas [n, 31]`` float64 CUDA tensor -> packed models ``[n, jb_packed_model_size()]`` float32 on the same
    device (input of ``BatchedJitterbugEnv(packed_models=...)``).  ``return_structs``: also return the models as a host
    array of ``JbModelF`` structs (inspection / parity tests)."""
    import ctypes as C

    import torch

    from . import _native
    from .model import JbModelF
    lib = _native.load()
    if deltas.dtype != torch.float64 or deltas.dim() != 2 or deltas.shape[1] != _native.AUG_NDELTA or not deltas.is_cuda:
        raise ValueError("deltas must be a float64 CUDA tensor [n, 31]")
    deltas = deltas.contiguous()
    base = make_aug_base() if base is None else base
    n = deltas.shape[0]
    packed = torch.empty((n, lib.jb_packed_model_size()), dtype=torch.float32, device=deltas.device)
    structs = torch.empty((n, C.sizeof(JbModelF)), dtype=torch.uint8, device=deltas.device) if return_structs else None
    with torch.cuda.device(deltas.device):
        _native.check(lib.jb_augment_compile(C.byref(base), deltas.data_ptr(), n, packed.data_ptr(),
                                             structs.data_ptr() if return_structs else None,
                                             torch.cuda.current_stream(deltas.device).cuda_stream))
    if not return_structs:
        return packed
    host = structs.cpu().numpy()
    arr = (JbModelF * n).from_buffer_copy(host.tobytes())
    return packed, arr


def make_batched_env(task="move_from_origin", num_envs=1, device="cuda:0", seed=0, env_id_offset=0, augment_kwargs=None,
                     **env_kwargs):
    """``BatchedJitterbugEnv`` with one device-compiled augmented model per env (BASELINE config 4)."""
    from .batched_env import BatchedJitterbugEnv
    deltas = sample_deltas(num_envs, seed=seed, device=device, env_id_offset=env_id_offset, **(augment_kwargs or {}))
    packed = compile_models_device(deltas)
    del deltas
    return BatchedJitterbugEnv(task=task, num_envs=num_envs, device=device, seed=seed, env_id_offset=env_id_offset,
                               packed_models=packed, **env_kwargs)


def get_model_and_assets(**augment_kwargs):
    """Returns a tuple containing the (augmented) model XML string and a dict of assets"""
    return augment_Jitterbug(**augment_kwargs), {}


def _make(task_name, time_limit, control_timestep, random, environment_kwargs, augment_kwargs=None, **kwargs):
    augment_kwargs = dict(modify_legs=True, modify_mass=True) if augment_kwargs is None else dict(augment_kwargs)
    augment_kwargs.setdefault("random", random if not isinstance(random, np.random.RandomState) else random)
    model = mjcf.compile_xml(get_model_and_assets(**augment_kwargs)[0])
    return _jb._make(task_name, time_limit, control_timestep, random, environment_kwargs, model=model, **kwargs)


@SUITE.add("benchmarking", "easy")
def move_from_origin(time_limit=DEFAULT_TIME_LIMIT, control_timestep=DEFAULT_CONTROL_TIMESTEP, random=None,
                     environment_kwargs=None, **kwargs):
    """Move the (augmented) Jitterbug away from the origin"""
    return _make("move_from_origin", time_limit, control_timestep, random, environment_kwargs, **kwargs)


@SUITE.add("benchmarking", "easy")
def face_direction(time_limit=DEFAULT_TIME_LIMIT, control_timestep=DEFAULT_CONTROL_TIMESTEP, random=None,
                   environment_kwargs=None, **kwargs):
    """Move the (augmented) Jitterbug to face a certain yaw angle"""
    return _make("face_direction", time_limit, control_timestep, random, environment_kwargs, **kwargs)


@SUITE.add("benchmarking", "easy")
def move_in_direction(time_limit=DEFAULT_TIME_LIMIT, control_timestep=DEFAULT_CONTROL_TIMESTEP, random=None,
                      environment_kwargs=None, **kwargs):
    """Move the (augmented) Jitterbug in a certain direction"""
    return _make("move_in_direction", time_limit, control_timestep, random, environment_kwargs, **kwargs)


@SUITE.add("benchmarking", "hard")
def move_to_position(time_limit=DEFAULT_TIME_LIMIT, control_timestep=DEFAULT_CONTROL_TIMESTEP, random=None,
                     environment_kwargs=None, **kwargs):
    """Move the (augmented) Jitterbug to a certain XYZ position"""
    return _make("move_to_position", time_limit, control_timestep, random, environment_kwargs, **kwargs)


@SUITE.add("benchmarking", "hard")
def move_to_pose(time_limit=DEFAULT_TIME_LIMIT, control_timestep=DEFAULT_CONTROL_TIMESTEP, random=None,
                 environment_kwargs=None, **kwargs):
    """Move the (augmented) Jitterbug to a certain XYZRPY pose"""
    return _make("move_to_pose", time_limit, control_timestep, random, environment_kwargs, **kwargs)
