"""Environment ids of the reference (marlenv/envs/__init__.py:3-16).  Works without gym; when gym or
gymnasium is importable the ids are registered there as well so `gym.make('Snake-v1', ...)` resolves."""
from .env import CoopSnakeEnv, SnakeEnv

_REGISTRY = {'snake-v1': SnakeEnv, 'snakecoop-v1': CoopSnakeEnv}


def make(env_id, **kwargs):
    """gym.make-compatible: accepts 'Snake-v1' (registered spelling) and 'snake-v1' (README spelling)."""
    kwargs.pop('disable_env_checker', None)
    kwargs.pop('render_mode', None)
    try:
        cls = _REGISTRY[env_id.lower()]
    except KeyError:
        raise KeyError(f'unknown environment id {env_id!r}; known: Snake-v1, SnakeCoop-v1') from None
    return cls(**kwargs)


def register_with_gym():
    done = []
    for modname in ('gym', 'gymnasium'):
        try:
            mod = __import__(modname)
            from importlib import import_module
            reg = import_module(modname + '.envs.registration')
        except Exception:
            continue
        for env_id, cls in (('Snake-v1', 'SnakeEnv'), ('SnakeCoop-v1', 'CoopSnakeEnv')):
            try:
                reg.register(id=env_id, entry_point=f'marl_snake_b200.env:{cls}')
                done.append((mod.__name__, env_id))
            except Exception:
                pass
    return done
