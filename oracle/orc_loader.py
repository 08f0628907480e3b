"""ORACLE -- TEST INFRASTRUCTURE ONLY.  ctypes loader for oracle/liborc.so (the CPU
restatement of the ALE/Stella machine).  Imported only by tests/, bench.py's
cpu_baseline / --impl reference legs and __graft_entry__.smoke()."""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "liborc.so")
    srcs = [os.path.join(_HERE, f) for f in ("oracle_capi.cpp", "a2600.hpp", "ale.hpp")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liborc.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_uint32]
        L.orc_console_create.restype = C.c_void_p
        L.orc_console_create.argtypes = [C.c_char_p, C.c_int]
        for f in ("orc_destroy", "orc_reset_game", "orc_console_frame"):
            getattr(L, f).argtypes = [C.c_void_p]
            getattr(L, f).restype = None
        for f in ("orc_game_over", "orc_lives", "orc_num_actions", "orc_frame_number"):
            getattr(L, f).argtypes = [C.c_void_p]
            getattr(L, f).restype = C.c_int
        L.orc_act.argtypes = [C.c_void_p, C.c_int]
        L.orc_act.restype = C.c_int
        L.orc_console_step.argtypes = [C.c_void_p, C.c_int]
        L.orc_set_ram.argtypes = [C.c_void_p, C.c_int, C.c_int]
        for f in ("orc_minimal_actions", "orc_get_ram", "orc_get_screen", "orc_get_both_screens", "orc_get_screen_gray",
                  "orc_get_screen_rgb", "orc_get_cpu"):
            getattr(L, f).argtypes = [C.c_void_p, C.c_void_p]
            getattr(L, f).restype = None
        L.orc_palette.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_hmove_table.argtypes = [C.c_void_p]
        L.orc_pool_step.restype = C.c_long
        L.orc_pool_step.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        _LIB = L
    return _LIB
