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

For batched simulation, :func:`make_models` compiles a population of augmented models to hand to
``BatchedJitterbugEnv(models=...)`` (one model per env, staged per CTA in shared memory by the kernel).
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
