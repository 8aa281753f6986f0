"""ctypes binding of libaemcarl_b200.so (include/aemcarl_b200.h).  No CPU fallback: a missing library or a
failing call raises."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libaemcarl_b200.so")

NUM_WEIGHTS = 113752
MAX_HUMANS = 32
MAX_ACTIONS = 128

INFO_NOTHING, INFO_DANGER, INFO_REACHGOAL, INFO_COLLISION, INFO_TIMEOUT = range(5)

c_double_p = ctypes.POINTER(ctypes.c_double)
c_float_p = ctypes.POINTER(ctypes.c_float)
c_int32_p = ctypes.POINTER(ctypes.c_int32)
c_uint8_p = ctypes.POINTER(ctypes.c_uint8)


class AemError(RuntimeError):
    pass


class EnvParams(ctypes.Structure):
    """aem_env_params (crowd_sim.py:265-299 CrowdSim.configure)."""
    _fields_ = [
        ("time_step", ctypes.c_double),
        ("time_limit", ctypes.c_double),
        ("success_reward", ctypes.c_double),
        ("collision_penalty", ctypes.c_double),
        ("discomfort_dist", ctypes.c_double),
        ("agent_timestep", ctypes.c_double),
        ("human_timestep", ctypes.c_double),
        ("reward_increment", ctypes.c_double),
        ("position_variance", ctypes.c_double),
        ("direction_variance", ctypes.c_double),
        ("robot_visible", ctypes.c_int32),
        ("pad_", ctypes.c_int32),
        ("orca_safety_space", ctypes.c_double),
    ]


class ScenarioParams(ctypes.Structure):
    """aem_scenario_params (env.config [sim] / [humans] / [robot] values that enter the scenario draw)."""
    _fields_ = [("circle_radius", ctypes.c_double), ("square_width", ctypes.c_double), ("discomfort_dist", ctypes.c_double),
                ("robot_radius", ctypes.c_double), ("robot_v_pref", ctypes.c_double), ("human_radius", ctypes.c_double),
                ("human_v_pref", ctypes.c_double), ("randomize_attributes", ctypes.c_int32), ("pad_", ctypes.c_int32)]


# every symbol include/aemcarl_b200.h declares: name -> (restype, argtypes)
_VP = ctypes.c_void_p
SYMBOLS = {
    "aem_version": (ctypes.c_int, []),
    "aem_last_error": (ctypes.c_char_p, []),
    "aem_create": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(_VP)]),
    "aem_destroy": (ctypes.c_int, [_VP]),
    "aem_load_weights": (ctypes.c_int, [_VP, _VP, ctypes.c_size_t]),
    "aem_set_action_space": (ctypes.c_int, [_VP, _VP, ctypes.c_int]),
    "aem_set_env_params": (ctypes.c_int, [_VP, ctypes.POINTER(EnvParams)]),
    "aem_set_act_mode": (ctypes.c_int, [_VP, ctypes.c_int, ctypes.c_int]),
    "aem_set_precision": (ctypes.c_int, [_VP, ctypes.c_int]),
    "aem_orca": (ctypes.c_int, [_VP, ctypes.c_int, ctypes.c_int, _VP, _VP, _VP, _VP]),
    "aem_env_lookahead": (ctypes.c_int, [_VP, ctypes.c_int, ctypes.c_int] + [_VP] * 9),
    "aem_lookahead": (ctypes.c_int, [_VP, ctypes.c_int, ctypes.c_int, _VP, _VP, _VP, ctypes.c_double] + [_VP] * 5),
    "aem_env_apply": (ctypes.c_int, [_VP, ctypes.c_int, ctypes.c_int] + [_VP] * 13),
    "aem_env_prepare": (ctypes.c_int, [_VP, ctypes.c_int, ctypes.c_int] + [_VP] * 9),
    "aem_env_finish": (ctypes.c_int, [_VP, ctypes.c_int, ctypes.c_int] + [_VP] * 15 + [ctypes.c_int, _VP, ctypes.c_int, _VP, _VP]),
    "aem_env_reset_done": (ctypes.c_int, [_VP, ctypes.c_int, ctypes.c_int] + [_VP] * 7 + [ctypes.c_int, _VP, ctypes.c_int, _VP, _VP]),
    "aem_value_forward": (ctypes.c_int, [_VP, ctypes.c_int, ctypes.c_int, _VP, _VP, _VP, _VP]),
    "aem_gru_act": (ctypes.c_int, [_VP, ctypes.c_int, ctypes.c_int, _VP, _VP, _VP, _VP]),
    "aem_decide_step_host": (ctypes.c_int, [_VP, ctypes.c_int, ctypes.c_int, ctypes.c_double] + [_VP] * 8),
    "aem_env_step_xy": (ctypes.c_int, [_VP, ctypes.c_int, ctypes.c_int, _VP, _VP, _VP, _VP, _VP, ctypes.c_double, _VP, _VP, _VP,
                                       _VP, _VP, _VP]),
    "aem_transform": (ctypes.c_int, [_VP, ctypes.c_int, ctypes.c_int, _VP, _VP, _VP, _VP]),
    "aem_generate_mixed_scenarios": (ctypes.c_int, [_VP, ctypes.c_uint32, ctypes.c_int] + [_VP] * 5),
    "aem_generate_scenarios": (ctypes.c_int, [_VP, ctypes.c_int, ctypes.c_uint32, ctypes.c_int, ctypes.c_int,
                                              ctypes.POINTER(ScenarioParams), _VP, _VP, _VP]),
    "aem_set_active_counts": (ctypes.c_int, [_VP, _VP, _VP]),
    "aem_set_train_dropout": (ctypes.c_int, [_VP, ctypes.c_double, ctypes.c_uint64]),
    "aem_train_grad": (ctypes.c_int, [_VP, ctypes.c_int, ctypes.c_int] + [_VP] * 8),
    "aem_sample_batch": (ctypes.c_int, [_VP, ctypes.c_int, ctypes.c_int, ctypes.c_int, _VP, _VP, ctypes.c_uint64, ctypes.c_uint64,
                                        _VP, _VP, _VP, _VP]),
    "aem_adam_step": (ctypes.c_int, [_VP, ctypes.c_size_t, _VP, _VP, _VP, _VP, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                     ctypes.c_double, ctypes.c_int, ctypes.c_double, _VP]),
    "aem_train_step_adam": (ctypes.c_int, [_VP, ctypes.c_int, ctypes.c_int, ctypes.c_int, _VP, _VP, ctypes.c_uint64, _VP, _VP, _VP, _VP,
                                           _VP, _VP, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double, _VP, _VP,
                                           _VP, _VP, _VP, _VP]),
    "aem_sgd_step": (ctypes.c_int, [_VP, ctypes.c_size_t, _VP, _VP, _VP, ctypes.c_double, ctypes.c_double, ctypes.c_int,
                                    ctypes.c_double, _VP]),
    "aem_rmsprop_step": (ctypes.c_int, [_VP, ctypes.c_size_t, _VP, _VP, _VP, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                        ctypes.c_double, _VP]),
    "aem_launch_count": (ctypes.c_longlong, [_VP]),
    "aem_train_scratch_generation": (ctypes.c_longlong, [_VP]),
}

_lib = None


def load():
    """dlopen the in-tree library and type every entry point.  Raises if it has not been built."""
    global _lib
    if _lib is None:
        path = os.environ.get("AEM_LIB_PATH") or LIB_PATH      # AEM_LIB_PATH: an experiment build (build.py --variant)
        if not os.path.exists(path):
            raise AemError("libaemcarl_b200.so is missing (%s); run `python -m aemcarl_b200.build` -- there is no "
                           "CPU fallback" % path)
        lib = ctypes.CDLL(path)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc):
    if rc != 0:
        raise AemError("aemcarl_b200 error %d: %s" % (rc, load().aem_last_error().decode("utf-8", "replace")))
