"""Alias helper of the drop-in import surface: registers every module of an aemcarl_b200 sub-package under the reference's
top-level name (crowd_sim.*, crowd_nav.*), so `from crowd_nav.policy.policy_factory import policy_factory` resolves to the
B200 implementation.  Opt-in: only a process that puts this directory on sys.path sees these names."""
import importlib
import pkgutil
import sys


def alias_package(real_name, alias_name, into):
    real = importlib.import_module(real_name)
    into.__path__ = list(getattr(real, "__path__", []))      # sub-modules that are imported later resolve there too
    for k, v in vars(real).items():
        if not k.startswith("__"):
            setattr(into, k, v)
    for info in pkgutil.walk_packages(real.__path__, real_name + "."):
        try:
            mod = importlib.import_module(info.name)
        except Exception:                                     # optional pieces (e.g. gym registration) must not break the alias
            continue
        sys.modules[alias_name + info.name[len(real_name):]] = mod
    return real
