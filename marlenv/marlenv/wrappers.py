from marl_snake_b200.wrappers import (RenderGUI, SingleAgent, SingleMultiAgent,  # noqa: F401
                                      VectorSnakeEnv, make_snake)
