class AsyncVectorEnv:
    """Placeholder: the stub does not provide process vectorisation."""

    def __init__(self, env_fns, **kwargs):
        raise NotImplementedError("gym stub: AsyncVectorEnv is not available")
