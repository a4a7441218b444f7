"""gpytorch.utils.memoize restated (test infrastructure): name-keyed cache on
the object, ignoring call arguments (the only mode the reference uses)."""
import functools
from .errors import CachingError


def cached(method=None, name=None, ignore_args=False):
    if method is None:
        return functools.partial(cached, name=name, ignore_args=ignore_args)

    @functools.wraps(method)
    def g(self, *args, **kwargs):
        cache_name = name if name is not None else method.__name__
        cache = self.__dict__.setdefault("_memoize_cache", {})
        if cache_name not in cache:
            cache[cache_name] = method(self, *args, **kwargs)
        return cache[cache_name]

    return g


def clear_cache_hook(module, *args, **kwargs):
    module.__dict__["_memoize_cache"] = {}


def pop_from_cache_ignore_args(obj, name):
    try:
        return obj.__dict__["_memoize_cache"].pop(name)
    except KeyError:
        raise CachingError(f"Object does not have item {name} stored in cache.")


pop_from_cache = pop_from_cache_ignore_args
