"""Minimal dm_env / dm_control look-alikes so the reference's call sites run unchanged without those packages.

The reference builds ``control.Environment(physics, task, time_limit, control_timestep)`` objects and hands out
``dm_env.TimeStep``s (jitterbug_dmc/jitterbug.py:84-90).  When the real ``dm_env`` is
importable its ``TimeStep`` / ``StepType`` / specs are used, so objects interoperate; otherwise the equivalents below.
"""

from __future__ import annotations

import collections
import enum

import numpy as np

try:  # pragma: no cover - not installed in this image
    import dm_env as _dm_env
    from dm_env import specs as _specs
    StepType = _dm_env.StepType
    TimeStep = _dm_env.TimeStep
    Array = _specs.Array
    BoundedArray = _specs.BoundedArray
    HAVE_DM_ENV = True
except Exception:  # noqa: BLE001
    HAVE_DM_ENV = False

    class StepType(enum.IntEnum):
        FIRST = 0
        MID = 1
        LAST = 2

        def first(self):
            return self is StepType.FIRST

        def mid(self):
            return self is StepType.MID

        def last(self):
            return self is StepType.LAST

    class TimeStep(collections.namedtuple("TimeStep", ["step_type", "reward", "discount", "observation"])):
        __slots__ = ()

        def first(self):
            return self.step_type == StepType.FIRST

        def mid(self):
            return self.step_type == StepType.MID

        def last(self):
            return self.step_type == StepType.LAST

    class Array:
        def __init__(self, shape, dtype, name=None):
            self.shape, self.dtype, self.name = tuple(shape), np.dtype(dtype), name

        def __repr__(self):
            return f"Array(shape={self.shape}, dtype={self.dtype}, name={self.name!r})"

        def validate(self, value):
            value = np.asarray(value)
            if value.shape != self.shape:
                raise ValueError(f"Expected shape {self.shape} but found {value.shape}")
            return value

        def generate_value(self):
            return np.zeros(self.shape, self.dtype)

    class BoundedArray(Array):
        def __init__(self, shape, dtype, minimum, maximum, name=None):
            super().__init__(shape, dtype, name)
            self.minimum = np.broadcast_to(np.asarray(minimum, dtype=self.dtype), self.shape).copy()
            self.maximum = np.broadcast_to(np.asarray(maximum, dtype=self.dtype), self.shape).copy()

        def __repr__(self):
            return (f"BoundedArray(shape={self.shape}, dtype={self.dtype}, name={self.name!r}, "
                    f"minimum={self.minimum}, maximum={self.maximum})")


class TaggedTasks:
    """dm_control.utils.containers.TaggedTasks look-alike (jitterbug.py:53 ``SUITE = containers.TaggedTasks()``)."""

    def __init__(self):
        self._tasks = collections.OrderedDict()
        self._tags = collections.defaultdict(dict)

    def add(self, *tags):
        def wrap(factory_func):
            name = factory_func.__name__
            if name in self._tasks:
                raise ValueError(f"task {name!r} already registered")
            self._tasks[name] = factory_func
            for tag in tags:
                self._tags[tag][name] = factory_func
            return factory_func
        return wrap

    def tagged(self, *tags):
        if not tags:
            return {}
        if tags[0] not in self._tags:
            return {}
        names = set(self._tags[tags[0]])
        for t in tags[1:]:
            names &= set(self._tags.get(t, {}))
        return collections.OrderedDict((n, f) for n, f in self._tasks.items() if n in names)

    def tags(self):
        return list(self._tags.keys())

    def __getitem__(self, k):
        return self._tasks[k]

    def __iter__(self):
        return iter(self._tasks)

    def __len__(self):
        return len(self._tasks)

    def __contains__(self, k):
        return k in self._tasks

    def keys(self):
        return self._tasks.keys()

    def items(self):
        return self._tasks.items()
