"""ORACLE -- TEST INFRASTRUCTURE ONLY.

Stand-in for the third-party `ale_python_interface` module the reference imports
(atari_emulator.py:2, environment_creator.py:25).  It exposes exactly the
`ALEInterface` methods the reference calls and forwards them to the CPU oracle
(oracle/liborc.so), so the reference's own atari_emulator.py / emulator_runner.py /
runners.py can run UNMODIFIED from /root/reference on top of it."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import orc_loader as liborc  # noqa: E402


# act() counters for the bench's reference arm: one int64 slot per ALEInterface, in memory shared with forked
# workers (the reference keeps its own per-emulator step count, but inside the worker processes where the parent
# cannot read it).  Off unless enable_act_counter() was called before the emulators were built.
_ACTS = None
_NEXT_SLOT = [0]


def enable_act_counter(n):
    global _ACTS
    from multiprocessing.sharedctypes import RawArray
    _ACTS = RawArray("q", int(n))
    _NEXT_SLOT[0] = 0
    return _ACTS


class ALEInterface(object):
    def __init__(self):
        self._L = liborc.lib()
        self._slot = None
        if _ACTS is not None and _NEXT_SLOT[0] < len(_ACTS):
            self._slot = _NEXT_SLOT[0]
            _NEXT_SLOT[0] += 1
        self._h = None
        self._ints = {b"random_seed": 0, b"frame_skip": 1}
        self._floats = {b"repeat_action_probability": 0.25}
        self._bools = {b"color_averaging": False}

    def setInt(self, key, value):
        self._ints[key] = int(value)

    def setFloat(self, key, value):
        self._floats[key] = float(value)

    def setBool(self, key, value):
        self._bools[key] = bool(value)

    def loadROM(self, path):
        if isinstance(path, bytes):
            path = path.decode()
        assert self._ints[b"frame_skip"] == 1 and self._floats[b"repeat_action_probability"] == 0.0, \
            "the oracle restates ALE only for the reference's settings (atari_emulator.py:22-26)"
        with open(path, "rb") as f:
            rom = f.read()
        game = os.path.splitext(os.path.basename(path))[0]
        if self._h:
            self._L.orc_destroy(self._h)
        self._h = self._L.orc_create(rom, len(rom), game.encode(), self._ints[b"random_seed"] & 0xFFFFFFFF)

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.orc_destroy(self._h)
            self._h = None

    def getMinimalActionSet(self):
        n = self._L.orc_num_actions(self._h)
        out = np.zeros(n, dtype=np.int32)
        self._L.orc_minimal_actions(self._h, out.ctypes.data)
        return out

    def getScreenDims(self):
        return 160, 210

    def lives(self):
        return self._L.orc_lives(self._h)

    def reset_game(self):
        self._L.orc_reset_game(self._h)

    def act(self, action):
        if self._slot is not None:
            _ACTS[self._slot] += 1
        return self._L.orc_act(self._h, int(action))

    def game_over(self):
        return bool(self._L.orc_game_over(self._h))

    def getScreenGrayscale(self, buf=None):
        if buf is None:
            buf = np.empty((210, 160, 1), dtype=np.uint8)
        self._L.orc_get_screen_gray(self._h, buf.ctypes.data)
        return buf

    def getScreenRGB(self, buf=None):
        if buf is None:
            buf = np.empty((210, 160, 3), dtype=np.uint8)
        self._L.orc_get_screen_rgb(self._h, buf.ctypes.data)
        return buf

    # --- taps beyond the reference's call set (parity tests only)
    def getScreen(self, buf=None):
        if buf is None:
            buf = np.empty((210, 160), dtype=np.uint8)
        self._L.orc_get_screen(self._h, buf.ctypes.data)
        return buf

    def getRAM(self, buf=None):
        if buf is None:
            buf = np.empty(128, dtype=np.uint8)
        self._L.orc_get_ram(self._h, buf.ctypes.data)
        return buf

    def getCPU(self):
        out = np.zeros(10, dtype=np.int32)
        self._L.orc_get_cpu(self._h, out.ctypes.data)
        return out

    def getFrameNumber(self):
        return self._L.orc_frame_number(self._h)
