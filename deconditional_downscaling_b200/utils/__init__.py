"""Host-side helpers with the surface of the reference's `src.utils` (src/utils/__init__.py:1-2):
the builder registry used by the experiment runners and the seeding decorator of the data generators."""
import functools
import random
import types

import numpy as np
import torch

__all__ = ["Registry", "setseed"]


class Registry(dict):
    """name -> builder mapping filled by call or by decorator (src/utils/registry.py:4-66).

    A class is registered through its `build` classmethod, a function as itself; anything else is a
    TypeError; registering a name twice is an error."""

    def register(self, name, module_object=None):
        if module_object is None:                      # decorator form: @REG.register("name")
            def deco(obj):
                self.register(name, obj)
                return obj
            return deco
        if isinstance(module_object, type):
            builder = module_object.build
        elif isinstance(module_object, types.FunctionType):
            builder = module_object
        else:
            raise TypeError("Trying to register unknown data type")
        if name in self:
            raise AssertionError(f"{name!r} is already registered")
        self[name] = builder
        return module_object


_SEEDERS = {"random": random.seed, "numpy": np.random.seed, "torch": torch.manual_seed}


def setseed(*kinds):
    """Decorator for functions taking a `seed=` keyword: seeds the named generators ('random', 'numpy',
    'torch') before the call when a truthy seed is given (src/utils/wrappers.py:7-28)."""
    unknown = set(kinds) - set(_SEEDERS)
    if unknown:
        raise AssertionError("Unkown seeding type")

    def deco(fn):
        @functools.wraps(fn)
        def wrapper(*args, seed=None, **kwargs):
            if seed:
                for k in kinds:
                    _SEEDERS[k](seed)
            return fn(*args, seed=seed, **kwargs)
        return wrapper
    return deco
