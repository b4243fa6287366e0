"""Module-level drop-in surface: the registry names the reference's systems/ layer asks for
(models/__init__.py:4-13 of the reference): 'nerf', 'neus', 'volume-density', 'volume-sdf',
'volume-radiance', 'volume-color'.  ``make(name, config)`` accepts OmegaConf nodes, plain dicts or
our ``Config``."""
from ..config import Config, as_config  # noqa: F401

_REGISTRY = {}


def register(name):
    def deco(cls):
        _REGISTRY[name] = cls
        return cls
    return deco


def make(name, config):
    if name not in _REGISTRY:
        raise KeyError(f'unknown model {name!r}; registered: {sorted(_REGISTRY)}')
    return _REGISTRY[name](as_config(config))


from . import fields, nerf_model, neus_model  # noqa: E402,F401
