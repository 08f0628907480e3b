"""ctypes binding of the C ABI declared in include/manette_b200.h (libmanette_b200.so).

There is no CPU fallback: if the library is missing, or a call fails, this raises."""
import ctypes as C
import os

from . import build as _build

_LIB = None


class NativeError(RuntimeError):
    pass


class MnGame(C.Structure):
    _fields_ = [("name", C.c_char_p), ("rom", C.c_char_p), ("rom_size", C.c_int), ("n_envs", C.c_int)]


class MnConfig(C.Structure):
    _fields_ = [("device", C.c_int), ("n_games", C.c_int), ("games", C.POINTER(MnGame)), ("rgb", C.c_int),
                ("single_life_episodes", C.c_int), ("random_start", C.c_int), ("random_seed", C.c_int),
                ("env_id_offset", C.c_int), ("nb_choices", C.c_int), ("tab_rep", C.POINTER(C.c_int)),
                ("envs_per_warp", C.c_int), ("draw_all_frames", C.c_int), ("no_reset_memo", C.c_int), ("history", C.c_int)]


class MnBuffers(C.Structure):
    _fields_ = [("n_envs", C.c_int), ("num_actions", C.c_int), ("nb_choices", C.c_int), ("depth", C.c_int),
                ("states", C.c_void_p), ("rewards", C.c_void_p), ("terminals", C.c_void_p), ("actions", C.c_void_p),
                ("repetitions", C.c_void_p), ("action_idx", C.c_void_p), ("repetition_idx", C.c_void_p),
                ("next_calls", C.c_void_p), ("frames", C.c_void_p), ("ring", C.c_void_p), ("history", C.c_void_p),
                ("history_depth", C.c_int)]


class MnRolloutBuffers(C.Structure):
    _fields_ = [("n_envs", C.c_int), ("max_local_steps", C.c_int), ("num_actions", C.c_int), ("nb_choices", C.c_int),
                ("rewards", C.c_void_p), ("masks", C.c_void_p), ("actions", C.c_void_p), ("repetitions", C.c_void_p),
                ("episode_reward", C.c_void_p), ("episode_steps", C.c_void_p), ("actions_sum", C.c_void_p),
                ("action_rep", C.c_void_p), ("stats", C.c_void_p), ("finished_reward", C.c_void_p),
                ("finished_steps", C.c_void_p), ("finished_count", C.c_void_p)]


# every symbol include/manette_b200.h declares: name -> (restype, argtypes)
_VP, _I, _U32, _U64, _F, _D = C.c_void_p, C.c_int, C.c_uint32, C.c_uint64, C.c_float, C.c_double
SYMBOLS = {
    "mn_last_error": (C.c_char_p, []),
    "mn_create": (_I, [C.POINTER(MnConfig), C.POINTER(_VP)]),
    "mn_destroy": (_I, [_VP]),
    "mn_get_buffers": (_I, [_VP, C.POINTER(MnBuffers)]),
    "mn_set_tab_rep": (_I, [_VP, C.POINTER(_I), _I]),
    "mn_legal_actions": (_I, [_VP, _I, _VP]),
    "mn_reset_all": (_I, [_VP, _VP]),
    "mn_step_async": (_I, [_VP, _I, _VP]),
    "mn_wait": (_I, [_VP]),
    "mn_set_host_states": (_I, [_VP, _VP]),
    "mn_step_host": (_I, [_VP, _VP, _VP, _VP, _VP, _VP, _VP]),
    "mn_env_reset": (_I, [_VP, _I, _VP]),
    "mn_env_next": (_I, [_VP, _I, _I, C.POINTER(_F), C.POINTER(_I), _VP]),
    "mn_get_ram": (_I, [_VP, _I, _VP]),
    "mn_get_screen": (_I, [_VP, _I, _VP]),
    "mn_get_cpu_state": (_I, [_VP, _I, _VP]),
    "mn_get_lives": (_I, [_VP, _I, C.POINTER(_I), C.POINTER(_I), C.POINTER(_I)]),
    "mn_total_next_calls": (_I, [_VP, C.POINTER(C.c_int64)]),
    "mn_memo_stats": (_I, [_VP, C.POINTER(C.c_int64)]),
    "mn_memo_level1_hits": (_I, [_VP, C.POINTER(C.c_int64)]),
    "mn_total_instructions": (_I, [_VP, C.POINTER(C.c_int64)]),
    "mn_redo_count": (_I, [_VP, C.POINTER(C.c_int64)]),
    "mn_palette": (_I, [_VP, _VP]),
    "mn_check_report": (_I, [_VP]),
    "mn_diag_counters": (_I, [_VP, _VP]),
    "mn_start_noops": (_I, [_U32, _U32, _U32]),
    "mn_preprocess": (_I, [_VP, _VP, _I, _I, _VP]),
    "mn_sample_figar": (_I, [_VP, _VP, _I, _I, _I, _I, _F, _U64, _U32, _VP, _VP, _VP, _VP, _VP]),
    "mn_nstep": (_I, [_VP, _VP, _VP, _VP, _D, _I, _I, _I, _VP, _VP, _VP]),
    "mn_history_head": (_I, [_VP, C.POINTER(_I)]),
    "mn_history_gather": (_I, [_VP, _VP, _VP]),
    "mn_rollout_create": (_I, [_I, _I, _I, _I, _I, C.POINTER(_I), C.POINTER(_VP)]),
    "mn_rollout_destroy": (_I, [_VP]),
    "mn_rollout_get_buffers": (_I, [_VP, C.POINTER(MnRolloutBuffers)]),
    "mn_rollout_begin": (_I, [_VP, _VP]),
    "mn_rollout_record": (_I, [_VP, _I, _VP, _VP, _VP, _VP, _I, _VP]),
    "mn_profile_begin": (_I, [_VP]),
    "mn_profile_end": (_I, [_VP, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "mn_launch_count": (_I, [_VP, C.POINTER(C.c_int64)]),
}


def lib_path():
    return _build.LIB_PATH


def load():
    """Returns the loaded library; raises NativeError if it has not been built."""
    global _LIB
    if _LIB is None:
        path = lib_path()
        if not os.path.exists(path):
            raise NativeError("%s is missing: build it with `python -m manette_b200.build` "
                              "(manette_b200 has no CPU fallback)" % path)
        L = C.CDLL(path)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)      # AttributeError here = header / library mismatch
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


def check(rc, what=""):
    if rc < 0:
        msg = load().mn_last_error()
        raise NativeError("%s failed: %s" % (what or "manette_b200 call", msg.decode() if msg else "unknown error"))
    return rc
