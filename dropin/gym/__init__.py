"""`import gym` for the reference's call sites (gym.make('CrowdSim-v0'), gym.Env, gym.envs.registration.register).
If a real gym / gymnasium is installed further down sys.path it is used; otherwise this minimal registry stands in."""
import importlib
import os
import sys
import types

_here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_real = None
_saved = list(sys.path)
try:
    sys.path[:] = [p for p in sys.path if os.path.abspath(p or ".") != _here]
    del sys.modules[__name__]
    try:
        _real = importlib.import_module("gym")
    except ImportError:
        _real = None
finally:
    sys.path[:] = _saved

if _real is not None:
    sys.modules[__name__] = _real
else:
    _mod = types.ModuleType(__name__)
    _mod.__file__ = __file__
    _mod.__path__ = [os.path.dirname(os.path.abspath(__file__))]
    _registry = {}

    class Env(object):
        metadata = {}

        def reset(self, *a, **k):
            raise NotImplementedError

        def step(self, *a, **k):
            raise NotImplementedError

    def register(id, entry_point, **kw):
        _registry[id] = entry_point

    def make(id, **kw):
        import crowd_sim                                   # noqa: F401  (registers CrowdSim-v0..4 like crowd_sim/__init__.py)
        mod, cls = _registry[id].split(":")
        return getattr(importlib.import_module(mod), cls)(**kw)

    _mod.Env, _mod.make, _mod.register = Env, make, register
    _envs = types.ModuleType(__name__ + ".envs")
    _reg = types.ModuleType(__name__ + ".envs.registration")
    _reg.register = register
    _envs.registration = _reg
    _mod.envs = _envs
    sys.modules[__name__] = _mod
    sys.modules[__name__ + ".envs"] = _envs
    sys.modules[__name__ + ".envs.registration"] = _reg
