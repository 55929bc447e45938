"""marl_snake_b200 -- B200-native batched Snake-v1 step path behind the marlenv API.

Directory name is `marl-snake_b200/` (repo layout); import it as `marl_snake_b200` through the shim
at the repository root.
"""
from ._lib import LIB_PATH, SnkError, lib                         # noqa: F401  (fails loudly if the .so is missing)
from .env import CoopSnakeEnv, SnakeBatch, SnakeEnv               # noqa: F401
from .registration import make, register_with_gym                # noqa: F401
from .wrappers import (RenderGUI, SingleAgent, SingleMultiAgent,  # noqa: F401
                       VectorSnakeEnv, make_snake)
from .dist import shard_range, allreduce_stats                    # noqa: F401
from .rollout import DeviceReplayBuffer, GraphedSteps, collect, epsilon_greedy, unpack_obs   # noqa: F401

__all__ = ['SnakeBatch', 'SnakeEnv', 'CoopSnakeEnv', 'make', 'make_snake', 'SingleAgent', 'SingleMultiAgent',
           'VectorSnakeEnv', 'RenderGUI', 'shard_range', 'allreduce_stats', 'DeviceReplayBuffer', 'collect',
           'epsilon_greedy', 'GraphedSteps', 'unpack_obs', 'lib', 'LIB_PATH', 'SnkError']
