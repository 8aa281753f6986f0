"""crowd_sim mirror: the gym ids of crowd_sim/__init__.py:3-26 behind a minimal registry (gym itself is optional)."""
_REGISTRY = {"CrowdSim-v%d" % i: "aemcarl_b200.crowd_sim.envs:CrowdSim" for i in range(5)}

try:                                    # register with gym when it is installed, exactly like the reference
    from gym.envs.registration import register as _register
    for _id, _ep in _REGISTRY.items():
        try:
            _register(id=_id, entry_point=_ep)
        except Exception:
            pass
except Exception:
    pass


def make(env_id):
    """gym.make(env_id) equivalent for the ids the reference registers"""
    import importlib
    mod, cls = _REGISTRY[env_id].split(":")
    return getattr(importlib.import_module(mod), cls)()
