import importlib

registry = {}


def register(id, entry_point, **kwargs):
    registry[id] = (entry_point, kwargs)


def make(id, **kwargs):
    kwargs.pop("disable_env_checker", None)
    entry_point, defaults = registry[id]
    mod_name, cls_name = entry_point.split(":")
    cls = getattr(importlib.import_module(mod_name), cls_name)
    merged = dict(defaults)
    merged.update(kwargs)
    return cls(**merged)
