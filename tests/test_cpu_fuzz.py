"""6502 / bus conformance fuzz: random cartridges (every opcode, documented or not, in every addressing mode;
random bank-switch hot-spot and TIA/RIOT accesses) run on the oracle console and on the DEVICE emulator core
compiled for the host; registers, status, cycle count and RIOT RAM must agree after every burst."""
import ctypes as C
import os

import numpy as np
import pytest

import util
from test_host_emu import libs  # noqa: F401  (fixture: builds + loads both libraries)


def _mk_rom(seed, size, bias):
    rng = np.random.RandomState(seed)
    rom = rng.randint(0, 256, size=size).astype(np.uint8)
    if bias:
        # sprinkle well-formed instruction runs so execution does not only tumble through garbage
        common = [0xA9, 0xA2, 0xA0, 0x85, 0x86, 0x84, 0x65, 0xE5, 0xC9, 0xE6, 0xC6, 0x0A, 0x4A, 0x2A, 0x6A, 0x95, 0xB5,
                  0x69, 0xE9, 0xF8, 0xD8, 0x38, 0x18, 0x48, 0x68, 0x08, 0x28, 0xAA, 0xA8, 0x8A, 0x98, 0x9A, 0xBA, 0xE8, 0xC8,
                  0xCA, 0x88, 0x24, 0x2C, 0xB1, 0x91, 0xA1, 0x81, 0xBD, 0x9D, 0xB9, 0x99, 0xD0, 0xF0, 0x10, 0x30, 0x90, 0xB0]
        for i in range(0, size - 4, 3):
            if rng.rand() < 0.6:
                rom[i] = common[rng.randint(len(common))]
                rom[i + 1] = rng.randint(0x80, 0x100)     # zero-page operands land in RAM
    rom[-4] = 0x00; rom[-3] = 0xF0     # reset vector -> $F000
    rom[-2] = 0x00; rom[-1] = 0xF8     # BRK/IRQ vector -> $F800
    return rom


@pytest.mark.parametrize("fast_mode", [1, 2, 0])
@pytest.mark.parametrize("size,bias", [(2048, True), (4096, True), (4096, False), (8192, True), (16384, True)])
def test_random_cartridges(libs, size, bias, fast_mode):
    """fast_mode: 1 = the kernels' flow (fast tick first), 2 = fast tick checked against the general path on every
    instruction it accepts, 0 = general path only."""
    L, H = libs
    H.he_set_fast_mode(fast_mode)
    L.orc_console_create.restype = C.c_void_p
    H.he_console_create.restype = C.c_void_p
    H.he_console_create.argtypes = [C.c_char_p, C.c_int]
    H.he_console_step.argtypes = [C.c_void_p, C.c_int]
    H.he_set_ram.argtypes = [C.c_void_p, C.c_int, C.c_int]
    for seed in range(12):
        rom = _mk_rom(1000 * size + seed, size, bias).tobytes()
        o = L.orc_console_create(rom, len(rom))
        h = H.he_console_create(rom, len(rom))
        rng = np.random.RandomState(seed)
        for j in range(128):
            v = int(rng.randint(256))
            L.orc_set_ram(o, j, v); H.he_set_ram(h, j, v)
        ro, rh = np.zeros(128, np.uint8), np.zeros(128, np.uint8)
        co, ch = np.zeros(10, np.int32), np.zeros(10, np.int32)
        so, sh = np.zeros(33600, np.uint8), np.zeros(33600, np.uint8)
        for burst in range(40):
            k = int(rng.randint(1, 400))
            L.orc_console_step(o, k); H.he_console_step(h, k)
            L.orc_get_cpu(o, co.ctypes.data); H.he_get_cpu(h, ch.ctypes.data)
            L.orc_get_ram(o, ro.ctypes.data); H.he_get_ram(h, rh.ctypes.data)
            assert np.array_equal(co[:7], ch[:7]) and co[8] == ch[8], (size, seed, burst, co, ch)
            assert np.array_equal(ro, rh), (size, seed, burst)
        L.orc_get_screen(o, so.ctypes.data); H.he_get_screen(h, sh.ctypes.data)
        assert np.array_equal(so, sh), (size, seed, "screen")
        L.orc_destroy(o); H.he_destroy(h)
    H.he_set_fast_mode(1)
