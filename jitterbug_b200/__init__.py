"""jitterbug_b200 -- B200-native batched simulator for the Jitterbug dm_control domain.

Drop-in surface of the reference package ``jitterbug_dmc`` (jitterbug_dmc/__init__.py:28-52):

    from jitterbug_b200 import suite
    env = suite.load(domain_name="jitterbug", task_name="move_from_origin")   # dm_control-style single env
    from jitterbug_b200 import BatchedJitterbugEnv, JitterbugGymEnv, heuristic_policies

The physics (MuJoCo's mj_step for this model) runs in hand-written CUDA kernels for sm_100a behind the C-ABI in
``include/jbsim.h``; there is no CPU fallback.
"""

from . import mjcf  # noqa: F401
from .batched_env import BatchedJitterbugEnv, TASKS, observation_layout  # noqa: F401
from . import jitterbug, augmented_jitterbug, suite, heuristic_policies, rollout, distributed, recorder, encoder  # noqa: F401
from .jitterbug import Jitterbug  # noqa: F401
from .gym_wrapper import JitterbugGymEnv  # noqa: F401
from .dm import StepType, TimeStep  # noqa: F401

__all__ = ["BatchedJitterbugEnv", "TASKS", "observation_layout", "mjcf", "suite", "jitterbug", "augmented_jitterbug",
           "heuristic_policies", "Jitterbug", "JitterbugGymEnv", "StepType", "TimeStep"]
