from marl_snake_b200.env import CoopSnakeEnv, SnakeEnv      # noqa: F401
from . import snake_env                                     # noqa: F401
