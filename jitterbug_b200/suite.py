"""``dm_control.suite``-style registry: ``suite.load(domain_name="jitterbug", task_name=...)``.

The reference patches its two domain modules into ``dm_control.suite._DOMAINS`` and recomputes the task tuples
(jitterbug_dmc/__init__.py:28-43).  This module carries the same names so user code only changes its import:

    from jitterbug_b200 import suite            # instead of: from dm_control import suite; import jitterbug_dmc
    env = suite.load(domain_name="jitterbug", task_name="move_from_origin")

If the real ``dm_control`` is importable, :func:`register_with_dm_control` patches these domains into it exactly like
the reference does.
"""

import collections
import inspect

from . import augmented_jitterbug, jitterbug

_DOMAINS = {name: module for name, module in (("jitterbug", jitterbug), ("augmented_jitterbug", augmented_jitterbug))
            if inspect.ismodule(module) and hasattr(module, "SUITE")}


def _get_tasks(tag):
    result = []
    for domain_name in sorted(_DOMAINS.keys()):
        domain = _DOMAINS[domain_name]
        tasks_in_domain = domain.SUITE if tag is None else domain.SUITE.tagged(tag)
        for task_name in tasks_in_domain.keys():
            result.append((domain_name, task_name))
    return tuple(result)


def _get_tasks_by_domain(tasks):
    result = collections.defaultdict(list)
    for domain_name, task_name in tasks:
        result[domain_name].append(task_name)
    return {k: tuple(v) for k, v in result.items()}


ALL_TASKS = _get_tasks(tag=None)
BENCHMARKING = _get_tasks('benchmarking')
EASY = _get_tasks('easy')
HARD = _get_tasks('hard')
EXTRA = tuple(sorted(set(ALL_TASKS) - set(BENCHMARKING)))
TASKS_BY_DOMAIN = _get_tasks_by_domain(ALL_TASKS)


def build_environment(domain_name, task_name, task_kwargs=None, environment_kwargs=None, visualize_reward=False):
    if domain_name not in _DOMAINS:
        raise ValueError('Domain {!r} does not exist.'.format(domain_name))
    domain = _DOMAINS[domain_name]
    if task_name not in domain.SUITE:
        raise ValueError('Level {!r} does not exist in domain {!r}.'.format(task_name, domain_name))
    task_kwargs = task_kwargs or {}
    if environment_kwargs is not None:
        task_kwargs = dict(task_kwargs, environment_kwargs=environment_kwargs)
    env = domain.SUITE[task_name](**task_kwargs)
    env.task.visualize_reward = visualize_reward
    return env


def load(domain_name, task_name, task_kwargs=None, environment_kwargs=None, visualize_reward=False):
    """Returns an environment from a domain name, task name and optional settings (dm_control.suite.load)."""
    return build_environment(domain_name, task_name, task_kwargs, environment_kwargs, visualize_reward)


def register_with_dm_control():
    """Patch the Jitterbug domains into a real ``dm_control.suite`` (what jitterbug_dmc/__init__.py:31-43 does)."""
    from dm_control import suite as dm_suite  # noqa: WPS433 - optional dependency
    dm_suite._DOMAINS.update(_DOMAINS)
    dm_suite.ALL_TASKS = dm_suite._get_tasks(tag=None)
    dm_suite.BENCHMARKING = dm_suite._get_tasks('benchmarking')
    dm_suite.EASY = dm_suite._get_tasks('easy')
    dm_suite.HARD = dm_suite._get_tasks('hard')
    dm_suite.EXTRA = tuple(sorted(set(dm_suite.ALL_TASKS) - set(dm_suite.BENCHMARKING)))
    dm_suite.TASKS_BY_DOMAIN = dm_suite._get_tasks_by_domain(dm_suite.ALL_TASKS)
    return dm_suite
