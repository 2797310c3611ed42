"""Golden vectors produced by EXECUTING THE REFERENCE'S OWN PYTHON (run in the build container, where /root/reference
exists; the GPU box only sees the committed .npz files).

    python tests/golden/make_reference_golden.py           # writes tests/golden/reference_*.npz

What runs is the reference's code, unmodified, imported from /root/reference:

* ``jitterbug_dmc/heuristic_policies.py`` (pure numpy) -- every policy on >= 10 k random normalised observations;
* ``jitterbug_dmc/jitterbug.py``: the bodies of ``Physics`` (helpers :180-317), ``Jitterbug.get_observation`` (:673-763),
  the reward functions (:840-925), ``Jitterbug.initialize_episode`` (:601-666) and the five task factories (:72-174);
* ``jitterbug_dmc/augmented_jitterbug.py::augment_Jitterbug`` (:95-267).

What is NOT the reference: dm_control / MuJoCo / tensorflow are not installed, so ``sys.modules`` stubs stand in for
them.  The stubs are bookkeeping only (a ``named`` view over state arrays, ``base.Task``'s RandomState handling,
``containers.TaggedTasks``, ``control.Environment`` recording its arguments) with two exceptions that carry arithmetic
and are restated from the upstream projects' published sources: ``mju_quat2Mat`` (MuJoCo engine_util_spatial.c) and
``dm_control.utils.rewards.tolerance`` / ``_sigmoids``.  Two mjData fields the reference reads are forward-kinematics
outputs (``sensordata['jitterbug_framelinvel']`` and the target geom pose); the fake physics computes them from the
state with this repo's compiled model (body COM offset), which is stated in the fixture (`framelinvel_from`).
The HEAD quirk ``use_VAE = True`` (needs ./VAE.pt) is neutralised by a stub ``VAE`` module and switching the flag off
after construction, i.e. README / paper behaviour.
"""
import collections
import contextlib
import importlib
import importlib.util
import io
import os
import sys
import tempfile
import types

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import reference_inputs as RI  # noqa: E402  (shared with the tests: the random inputs are regenerated from their seeds)

REF = os.environ.get("JB_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
TASKS = ("move_from_origin", "face_direction", "move_in_direction", "move_to_position", "move_to_pose")


# ------------------------------------------------------------------------------------------------ upstream restatements
def mju_quat2Mat(res, quat):
    """MuJoCo's mju_quat2Mat: row-major 3x3 from (w, x, y, z)."""
    q0, q1, q2, q3 = [float(x) for x in quat]
    q00, q01, q02, q03 = q0 * q0, q0 * q1, q0 * q2, q0 * q3
    q11, q12, q13, q22, q23, q33 = q1 * q1, q1 * q2, q1 * q3, q2 * q2, q2 * q3, q3 * q3
    res[0] = q00 + q11 - q22 - q33
    res[4] = q00 - q11 + q22 - q33
    res[8] = q00 - q11 - q22 + q33
    res[1] = 2 * (q12 - q03)
    res[2] = 2 * (q13 + q02)
    res[3] = 2 * (q12 + q03)
    res[5] = 2 * (q23 - q01)
    res[6] = 2 * (q13 - q02)
    res[7] = 2 * (q23 + q01)


def _sigmoids(x, value_at_1, sigmoid):
    """dm_control.utils.rewards._sigmoids (the three shapes the reference uses + the default family)."""
    if sigmoid in ('cosine', 'linear', 'quadratic'):
        if not 0 <= value_at_1 < 1:
            raise ValueError('`value_at_1` must be nonnegative and smaller than 1, got {}.'.format(value_at_1))
    else:
        if not 0 < value_at_1 < 1:
            raise ValueError('`value_at_1` must be strictly between 0 and 1, got {}.'.format(value_at_1))
    if sigmoid == 'gaussian':
        scale = np.sqrt(-2 * np.log(value_at_1))
        return np.exp(-0.5 * (x * scale) ** 2)
    elif sigmoid == 'hyperbolic':
        scale = np.arccosh(1 / value_at_1)
        return 1 / np.cosh(x * scale)
    elif sigmoid == 'long_tail':
        scale = np.sqrt(1 / value_at_1 - 1)
        return 1 / ((x * scale) ** 2 + 1)
    elif sigmoid == 'cosine':
        scale = np.arccos(2 * value_at_1 - 1) / np.pi
        scaled_x = x * scale
        return np.where(abs(scaled_x) < 1, (1 + np.cos(np.pi * scaled_x)) / 2, 0.0)
    elif sigmoid == 'linear':
        scale = 1 - value_at_1
        scaled_x = x * scale
        return np.where(abs(scaled_x) < 1, 1 - scaled_x, 0.0)
    elif sigmoid == 'quadratic':
        scale = np.sqrt(1 - value_at_1)
        scaled_x = x * scale
        return np.where(abs(scaled_x) < 1, 1 - scaled_x ** 2, 0.0)
    elif sigmoid == 'tanh_squared':
        scale = np.arctanh(np.sqrt(1 - value_at_1))
        return 1 - np.tanh(x * scale) ** 2
    raise ValueError('Unknown sigmoid type {!r}.'.format(sigmoid))


def tolerance(x, bounds=(0.0, 0.0), margin=0.0, sigmoid='gaussian', value_at_margin=0.1):
    """dm_control.utils.rewards.tolerance."""
    lower, upper = bounds
    if lower > upper:
        raise ValueError('Lower bound must be <= upper bound.')
    if margin < 0:
        raise ValueError('`margin` must be non-negative.')
    in_bounds = np.logical_and(lower <= x, x <= upper)
    if margin == 0:
        value = np.where(in_bounds, 1.0, 0.0)
    else:
        d = np.where(x < lower, lower - x, x - upper) / margin
        value = np.where(in_bounds, 1.0, _sigmoids(d, value_at_margin, sigmoid))
    return float(value) if np.isscalar(x) else value


# ------------------------------------------------------------------------------------------------ stubs (bookkeeping)
class _TaggedTasks:
    def __init__(self):
        self._tasks = collections.OrderedDict()
        self._tags = collections.defaultdict(dict)

    def add(self, *tags):
        def wrap(fn):
            self._tasks[fn.__name__] = fn
            for t in tags:
                self._tags[t][fn.__name__] = fn
            return fn
        return wrap

    def tagged(self, *tags):
        return [n for n in self._tasks if all(n in self._tags[t] for t in tags)]


class _Task:
    """dm_control.suite.base.Task: seeds ``self.random``; ``initialize_episode`` only refreshes reward colours."""

    def __init__(self, random=None):
        if not isinstance(random, np.random.RandomState):
            random = np.random.RandomState(random)
        self._random = random
        self._visualize_reward = False

    @property
    def random(self):
        return self._random

    def initialize_episode(self, physics):
        pass


class _Index:
    """``physics.named.<struct>.<field>[key]`` over plain arrays: ``get(key)`` returns a VIEW, ``set(key, value)`` writes."""

    def __init__(self, get, set=None):
        self._get, self._set = get, set

    def __getitem__(self, key):
        return self._get(key)

    def __setitem__(self, key, value):
        self._set(key, value)


class _FakeMujocoPhysics:
    """Stands in for dm_control.mujoco.Physics: the state arrays plus the few derived mjData fields the reference reads."""

    QPOS = {"root": slice(0, 7), "jointMass": slice(15, 16)}
    QVEL = {"root": slice(0, 6), "jointMass": slice(14, 15)}
    AXIS = {"x": 0, "y": 1, "z": 2}

    def __init__(self, com_offset):
        self.qpos = np.zeros(16)
        self.qpos[:7] = (0, 0, .035, 1, 0, 0, 0)
        self.qvel = np.zeros(15)
        self.target_body_pos = np.array([0., 0., .035])        # jitterbug.xml:114
        self.target_body_quat = np.array([1., 0., 0., 0.])
        self.rgba_alpha = {}
        self.com_offset = np.asarray(com_offset, dtype=np.float64)   # body "jitterbug": xipos - xpos in the body frame
        d, m = types.SimpleNamespace(), types.SimpleNamespace()
        d.qpos = _Index(lambda k: self.qpos[self.QPOS[k]])
        d.qvel = _Index(lambda k: self.qvel[self.QVEL[k]])
        d.geom_xpos = _Index(lambda k: self._target_geom_xpos(k))
        d.xquat = _Index(lambda k: self._target_xquat(k))
        d.sensordata = _Index(lambda k: self._framelinvel(k))
        d.xmat = _Index(lambda k: self._xmat(k))
        m.body_pos = _Index(None, self._set_body_pos)
        m.body_quat = _Index(None, self._set_body_quat)
        m.geom_rgba = _Index(None, lambda k, v: self.rgba_alpha.__setitem__(k, v))
        self.named = types.SimpleNamespace(data=d, model=m)

    # -- derived mjData fields
    def _rot(self):
        R = np.zeros(9)
        mju_quat2Mat(R, self.qpos[3:7] / np.linalg.norm(self.qpos[3:7]))
        return R.reshape(3, 3)

    def _target_geom_xpos(self, k):
        assert k == "target"
        return self.target_body_pos.copy()       # geom "target" sits at its body's origin (jitterbug.xml:115)

    def _target_xquat(self, k):
        assert k == "target"
        return self.target_body_quat / np.linalg.norm(self.target_body_quat)

    def _framelinvel(self, k):
        assert k == "jitterbug_framelinvel"
        # objtype="body": world-frame linear velocity of the body's inertial frame origin
        return self.qvel[0:3] + self._rot() @ np.cross(self.qvel[3:6], self.com_offset)

    def _xmat(self, k):
        assert k == ('jitterbug', 'zz')
        return self._rot()[2, 2]

    def _set_body_pos(self, k, v):
        assert k[0] == "target"
        self.target_body_pos[self.AXIS[k[1]]] = v

    def _set_body_quat(self, k, v):
        assert k == "target"
        self.target_body_quat[:] = v

    @contextlib.contextmanager
    def reset_context(self):
        yield self

    @classmethod
    def from_xml_string(cls, xml, assets=None):
        p = cls(_COM_OFFSET)
        p.xml, p.assets = xml, assets
        return p


class _Environment:
    """dm_control.rl.control.Environment: records what the factories pass."""

    def __init__(self, physics, task, time_limit=float('inf'), control_timestep=None, **kw):
        self.physics, self.task, self.time_limit, self.control_timestep, self.kwargs = physics, task, time_limit, control_timestep, kw


_COM_OFFSET = None


def install_stubs():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class _Anything(types.ModuleType):
        def __getattr__(self, k):
            if k.startswith("__"):
                raise AttributeError(k)
            return _Anything(self.__name__ + "." + k)

        def __call__(self, *a, **k):
            return self

    import importlib.machinery
    for name in ("tensorflow", "denoising_autoencoder"):
        m = _Anything(name)
        m.__spec__ = importlib.machinery.ModuleSpec(name, None)
        sys.modules[name] = m

    class _StubVAE:
        def __init__(self, *a, **k):
            pass

        def to(self, device):
            return self

        def parameters(self):
            import torch
            return [torch.nn.Parameter(torch.zeros(1))]

        def load_autoencoder(self, path):
            pass

    mod("VAE", VAE=_StubVAE)
    mjlib = types.SimpleNamespace(mju_quat2Mat=mju_quat2Mat)
    mod("dm_control")
    mod("dm_control.mujoco", Physics=_FakeMujocoPhysics)
    mod("dm_control.mujoco.wrapper")
    mod("dm_control.mujoco.wrapper.mjbindings", mjlib=mjlib)
    mod("dm_control.rl")
    mod("dm_control.rl.control", Environment=_Environment)
    sys.modules["dm_control.rl"].control = sys.modules["dm_control.rl.control"]
    mod("dm_control.suite")
    mod("dm_control.suite.base", Task=_Task)
    mod("dm_control.suite.common", ASSETS={})
    mod("dm_control.utils")
    mod("dm_control.utils.rewards", tolerance=tolerance)
    mod("dm_control.utils.containers", TaggedTasks=_TaggedTasks)
    mod("dm_control.utils.io", GetResource=lambda p: open(p, "rb").read())
    for parent, child in (("dm_control", "mujoco"), ("dm_control", "rl"), ("dm_control", "suite"), ("dm_control", "utils"),
                          ("dm_control.suite", "base"), ("dm_control.suite", "common"), ("dm_control.utils", "rewards"),
                          ("dm_control.utils", "containers"), ("dm_control.utils", "io")):
        setattr(sys.modules[parent], child, sys.modules[parent + "." + child])
    # the package WITHOUT running its __init__ (which patches dm_control.suite internals)
    pkg = types.ModuleType("jitterbug_dmc")
    pkg.__path__ = [os.path.join(REF, "jitterbug_dmc")]
    sys.modules["jitterbug_dmc"] = pkg


def load_reference():
    global _COM_OFFSET
    sys.path.insert(0, ROOT)
    from jitterbug_b200 import mjcf
    cm = mjcf.compile_xml(open(os.path.join(REF, "jitterbug_dmc", "jitterbug.xml")).read())
    _COM_OFFSET = np.array(cm.body_ipos[1], dtype=np.float64)
    import torch.optim  # noqa: F401  (the reference imports torch; load it before the tensorflow stub exists)
    install_stubs()
    jb = importlib.import_module("jitterbug_dmc.jitterbug")
    aug = importlib.import_module("jitterbug_dmc.augmented_jitterbug")
    spec = importlib.util.spec_from_file_location("ref_heuristic_policies",
                                                  os.path.join(REF, "jitterbug_dmc", "heuristic_policies.py"))
    hp = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(hp)
    return jb, aug, hp, cm


def make_task(jb, task, random=None, random_pose=True):
    with contextlib.redirect_stdout(io.StringIO()):
        t = jb.Jitterbug(random=random, task=task, random_pose=random_pose)
    t.use_VAE = False                     # HEAD hard-codes True (needs ./VAE.pt): canonical behaviour is "all encoders off"
    return t


# ------------------------------------------------------------------------------------------------ generators
def gen_heuristics(hp, n=RI.N_HEURISTIC, seed=RI.SEED_HEURISTIC):
    """Every reference policy on the shared random observations (inputs are regenerated from the seed by the tests)."""
    out = {"n": np.array(n), "seed": np.array(seed)}
    TS = collections.namedtuple("TS", "observation")
    for task in TASKS:
        obs = RI.heuristic_observations(task, n, seed).astype(np.float64)
        act = np.zeros(n)
        for i in range(n):
            d = collections.OrderedDict()
            d["position"], d["velocity"] = obs[i, 0:7], obs[i, 7:13]
            d["motor_position"], d["motor_velocity"] = obs[i, 13:14], obs[i, 14:15]
            c = 15
            for key in RI.EXTRA_KEYS[task]:
                d[key] = obs[i, c:c + RI.KEY_DIMS[key]]
                c += RI.KEY_DIMS[key]
            act[i] = float(np.ravel(getattr(hp, task)(TS(d)))[0])
        out[task + "_action"] = act
        out[task + "_obs_checksum"] = np.array(float(obs.sum()))
    return out


def gen_tasks(jb, n=RI.N_TASK_STATES, seed=RI.SEED_TASK_STATES):
    out = {"framelinvel_from": "v_root + R (w_body x ipos[jitterbug]) with ipos from jitterbug_b200.mjcf.compile_xml",
           "com_offset": _COM_OFFSET, "n": np.array(n), "seed": np.array(seed)}
    for task in TASKS:
        t = make_task(jb, task, random=0)
        ph = jb.Physics(_COM_OFFSET)
        qpos, qvel, tgt = RI.task_states(task, n, seed)
        obs_all, rew, comps, helpers = [], np.zeros(n), np.zeros((n, 4)), np.zeros((n, 11))
        for i in range(n):
            ph.qpos[:] = qpos[i]; ph.qvel[:] = qvel[i]
            ph.target_body_pos[:] = (tgt[i, 0], tgt[i, 1], .035)
            ph.target_body_quat[:] = (np.cos(tgt[i, 2] / 2), 0, 0, np.sin(tgt[i, 2] / 2))
            o = t.get_observation(ph)
            obs_all.append(np.concatenate([np.ravel(v) for v in o.values()]))
            if i == 0:
                out[task + "_keys"] = np.array(list(o.keys()))
                out[task + "_dims"] = np.array([np.size(v) for v in o.values()])
            rew[i] = t.get_reward(ph)
            comps[i] = (t.heading_reward(ph), t.velocity_reward(ph), t.position_reward(ph), t.upright_reward(ph))
            helpers[i] = np.concatenate([[ph.jitterbug_direction_yaw()], np.ravel(ph.motor_position()),
                                         [ph.target_direction_yaw()], ph.target_position_in_jitterbug_frame(),
                                         ph.jitterbug_velocity_in_target_frame(), ph.angle_jitterbug_to_target(),
                                         [ph.named.data.xmat['jitterbug', 'zz']]])
        out[task + "_state_checksum"] = np.array(float(qpos.sum() + qvel.sum() + tgt.sum()))
        out[task + "_obs"], out[task + "_reward"] = np.array(obs_all), rew
        out[task + "_reward_components"], out[task + "_helpers"] = comps, helpers
    return out


def gen_resets(jb, n=400, seed=20260103):
    """The reference's initialize_episode with known RNG seeds, and the uniforms it consumed (replayed from the same seeds)."""
    out = {}
    for task in TASKS:
        for rp in (True, False):
            quat, tg, us = np.zeros((n, 4)), np.zeros((n, 7)), np.zeros((n, 6))
            for i in range(n):
                s_task, s_glob = seed + 7 * i, seed + 7 * i + 3
                t = make_task(jb, task, random=s_task, random_pose=rp)
                ph = jb.Physics(_COM_OFFSET)
                np.random.seed(s_glob)
                t.initialize_episode(ph)
                quat[i] = ph.qpos[3:7]
                tg[i] = np.concatenate([ph.target_body_pos, ph.target_body_quat])
                r1, r2 = np.random.RandomState(s_task), np.random.RandomState(s_glob)
                us[i, 0] = r1.uniform(0, 2 * np.pi) / (2 * np.pi)
                us[i, 1] = (r1.uniform(.05, .2) - .05) / .15
                us[i, 2] = r2.uniform(0, 2 * np.pi) / (2 * np.pi)
                if rp:
                    us[i, 3] = r2.random_sample()
                    us[i, 4:6] = r2.random_sample(2)
            key = f"{task}_{'pose' if rp else 'nopose'}"
            out[key + "_quat"], out[key + "_target"], out[key + "_uniforms"] = quat, tg, us
    return out


def gen_factories(jb, aug):
    rows = []
    for modname, m in (("jitterbug", jb), ("augmented_jitterbug", aug)):
        for name, fn in m.SUITE._tasks.items():
            with contextlib.redirect_stdout(io.StringIO()):
                env = fn()
            env.task.use_VAE = False
            tags = sorted(t for t, d in m.SUITE._tags.items() if name in d)
            rows.append((modname, name, env.task.task, float(env.time_limit), float(env.control_timestep), ",".join(tags),
                         bool(env.task.random_pose)))
    return {"factories": np.array(rows, dtype=object).astype(str),
            "constants": np.array([jb.DEFAULT_TIME_LIMIT, jb.DEFAULT_CONTROL_TIMESTEP, jb.TARGET_SPEED]),
            "norm_all": jb.Jitterbug._NORM_ALL,
            **{"norm_" + k: v for k, v in jb.Jitterbug._NORM_TASKS.items()}}


AUG_CASES = [
    dict(modify_legs=True), dict(modify_mass=True), dict(modify_legs=True, modify_mass=True),
    dict(modify_coreBody1=True, modify_coreBody2=True), dict(modify_global_density=True), dict(modify_gear=True),
    dict(modify_legs=True, modify_mass=True, modify_coreBody1=True, modify_coreBody2=True, modify_global_density=True,
         modify_gear=True),
    dict(),
]


def gen_augment(aug, cm_mod, reps=3, seed=20260104):
    """The reference's augment_Jitterbug on its own jitterbug.xml (global numpy RNG seeded), output compiled by this
    repo's MJCF compiler: stores the perturbed attributes and the compiled arrays."""
    import xml.etree.ElementTree as ET
    out = {"n_cases": np.array([len(AUG_CASES), reps])}
    tmp = tempfile.mkdtemp()
    for ci, kw in enumerate(AUG_CASES):
        for r in range(reps):
            s = seed + 100 * ci + r
            np.random.seed(s)
            path = os.path.join(tmp, f"aug_{ci}_{r}.xml")
            with contextlib.redirect_stdout(io.StringIO()):
                aug.augment_Jitterbug(new_file=path, **kw)
            xml = open(path).read()
            root = ET.fromstring(xml)
            attrs = []
            for el in root.iter():
                if el.tag in ("geom", "joint", "general") and el.attrib.get("name"):
                    for a in ("fromto", "pos", "axis", "density", "size", "gear"):
                        if a in el.attrib:
                            attrs.append(f"{el.tag}:{el.attrib['name']}:{a}={el.attrib[a]}")
            for d in root.findall("default"):
                for g in d.findall("geom"):
                    attrs.append("default:geom:density=" + g.attrib["density"])
            cm = cm_mod.compile_xml(xml)
            key = f"case{ci}_rep{r}"
            out[key + "_seed"] = np.array(s)
            out[key + "_kwargs"] = np.array(sorted(kw))
            out[key + "_attrs"] = np.array(attrs)
            for f in ("body_mass", "body_inertia", "body_ipos", "body_iquat", "body_pos", "body_quat", "jnt_pos", "jnt_axis",
                      "geom_pos", "geom_quat", "geom_size", "body_invweight0", "dof_invweight0"):
                if hasattr(cm, f):
                    out[key + "_" + f] = np.asarray(getattr(cm, f), dtype=np.float64)
            out[key + "_gear"] = np.array(float(getattr(cm, "gear", np.nan)))
    return out


def main():
    if not os.path.isdir(REF):
        raise SystemExit(f"{REF} not found: the reference is only present in the build container")
    jb, aug, hp, cm = load_reference()
    from jitterbug_b200 import mjcf
    np.savez_compressed(os.path.join(HERE, "reference_heuristics.npz"), **gen_heuristics(hp))
    np.savez_compressed(os.path.join(HERE, "reference_tasks.npz"), **gen_tasks(jb))
    np.savez_compressed(os.path.join(HERE, "reference_resets.npz"), **gen_resets(jb))
    np.savez_compressed(os.path.join(HERE, "reference_factories.npz"), **gen_factories(jb, aug))
    np.savez_compressed(os.path.join(HERE, "reference_augment.npz"), **gen_augment(aug, mjcf))
    for f in sorted(os.listdir(HERE)):
        if f.startswith("reference_"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
