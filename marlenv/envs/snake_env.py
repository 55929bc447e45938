from marl_snake_b200.env import SnakeEnv                    # noqa: F401
