"""OpenAI-Gym style adapter -- drop-in for ``jitterbug_dmc/gym_wrapper.py:6-39`` (``JitterbugGymEnv``).

The reference subclasses ``dm2gym.DMControlEnv``: ``step(a) -> (observation, reward, done, {})`` with
``done = timestep.last()``, ``reset() -> observation``.  ``gym`` / ``dm2gym`` are not needed;
spaces are provided through ``gym.spaces`` when it is importable and through tiny ``Box`` / ``Dict`` look-alikes
otherwise.  Like ``dm2gym``, ``observation_space`` describes exactly what ``reset`` / ``step`` return: a ``Dict`` space
keyed like the observation dict (one ``"observations"`` entry for ``flat_observation=True`` environments).  There is no
renderer in the batched simulator: ``render`` keeps the reference's frame counting and is otherwise a no-op.
"""

import numpy as np

try:  # pragma: no cover
    from gym import spaces as _spaces
    _Box = _spaces.Box
except Exception:  # noqa: BLE001
    class _Box:
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.low = np.broadcast_to(np.asarray(low, dtype=dtype), shape).copy()
            self.high = np.broadcast_to(np.asarray(high, dtype=dtype), shape).copy()
            self.shape, self.dtype = tuple(shape), np.dtype(dtype)

        def sample(self):
            return np.random.uniform(self.low, self.high).astype(self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))


try:  # pragma: no cover
    _Dict = _spaces.Dict
except Exception:  # noqa: BLE001
    class _Dict(dict):
        """Minimal ``gym.spaces.Dict``: an ordered mapping of sub-spaces."""

        @property
        def spaces(self):
            return self

        def sample(self):
            return type(self)((k, v.sample()) for k, v in self.items())

        def contains(self, x):
            return set(x) == set(self) and all(self[k].contains(np.asarray(x[k], dtype=self[k].dtype)) for k in self)


class JitterbugGymEnv:
    """A Gym-style env around a Jitterbug ``Environment``.

    Args:
        env: the environment we are wrapping (``suite.load(...)``)
        render_every (int): kept for signature compatibility
    """

    metadata = {"render.modes": ["human", "rgb_array"]}

    def __init__(self, env, *, render_every=1):
        self.num_envs = 1
        self.frame_count = 0
        self.render_every = render_every
        self.env = env
        spec = env.action_spec()
        self.action_space = _Box(spec.minimum.astype(np.float32), spec.maximum.astype(np.float32), shape=spec.shape,
                                 dtype=np.float32)
        import collections
        self.observation_space = _Dict(collections.OrderedDict(
            (k, _Box(-np.inf, np.inf, shape=tuple(v.shape), dtype=np.float64)) for k, v in env.observation_spec().items()))
        self.reward_range = (0.0, 1.0)

    def _obs(self, ts):
        return ts.observation

    def reset(self):
        return self._obs(self.env.reset())

    def step(self, action):
        ts = self.env.step(action)
        reward = 0.0 if ts.reward is None else ts.reward
        return self._obs(ts), reward, ts.last(), {}

    def seed(self, seed=None):
        return [seed]

    def render(self, mode='human', **kwargs):
        self.frame_count += 1
        return None

    def close(self):
        if hasattr(self.env, "close"):
            self.env.close()
