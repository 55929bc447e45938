from . import async_vector_env, utils  # noqa: F401
