"""`marlenv.marlenv.*` spelling used by the reference's scripts (test_env.py:1, train_dqn.py:22)."""
from .. import envs                     # noqa: F401
from . import wrappers                  # noqa: F401
