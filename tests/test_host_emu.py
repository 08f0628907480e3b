"""CPU pre-check of the DEVICE emulator core: manette_b200/csrc/emu_core.cuh compiled with the host compiler
(tests/host_emu) against the oracle, so the kernels' logic is diffed on machines without a GPU.
(The GPU parity tests proper are in test_gpu_parity.py.)"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import util
from util import GAMES12, rom_bytes

import orc_loader

HE_DIR = os.path.join(util.ROOT, "tests", "host_emu")


@pytest.fixture(scope="module")
def libs():
    so = os.path.join(HE_DIR, "libhost_emu.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", so,
                           os.path.join(HE_DIR, "host_emu.cpp")])
    H = C.CDLL(so)
    H.he_create.restype = C.c_void_p
    H.he_create.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_uint32, C.c_int]
    H.he_destroy.argtypes = [C.c_void_p]
    H.he_set_drain_at.argtypes = [C.c_void_p, C.c_int]
    H.he_redo_count.argtypes = [C.c_void_p]
    H.he_act.argtypes = [C.c_void_p, C.c_int]
    H.he_next.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int)]
    H.he_reset_game.argtypes = [C.c_void_p, C.c_int, C.c_int]
    for f in ("he_game_over", "he_lives"):
        getattr(H, f).argtypes = [C.c_void_p]
    L = orc_loader.lib()
    for f in ("he_get_ram", "he_get_screen", "he_get_both_screens", "he_get_cpu"):
        getattr(H, f).argtypes = [C.c_void_p, C.c_void_p]
    return L, H


def _taps(L, H, o, h):
    ro, rh = np.zeros(128, np.uint8), np.zeros(128, np.uint8)
    so, sh = np.zeros(33600, np.uint8), np.zeros(33600, np.uint8)
    co, ch = np.zeros(10, np.int32), np.zeros(10, np.int32)
    L.orc_get_ram(o, ro.ctypes.data); H.he_get_ram(h, rh.ctypes.data)
    L.orc_get_screen(o, so.ctypes.data); H.he_get_screen(h, sh.ctypes.data)
    L.orc_get_cpu(o, co.ctypes.data); H.he_get_cpu(h, ch.ctypes.data)
    return (ro, so, co), (rh, sh, ch)


@pytest.mark.parametrize("game", GAMES12)
def test_every_frame_matches_oracle(libs, game):
    """act() by act(): RAM, raw screen, CPU registers, reward, game over -- every frame drawn."""
    L, H = libs
    rom = rom_bytes(game)
    o = L.orc_create(rom, len(rom), game.encode(), 6)
    h = H.he_create(rom, len(rom), game.encode(), 6, 0)
    n = L.orc_num_actions(o)
    acts = np.zeros(18, np.int32)
    L.orc_minimal_actions(o, acts.ctypes.data)
    rng = np.random.RandomState(1)
    for drain_at, frames in ((12, 500), (1, 120), (16, 120)):
        H.he_set_drain_at(h, drain_at)
        for i in range(frames):
            a = int(acts[rng.randint(n)])
            assert L.orc_act(o, a) == H.he_act(h, a), (game, i)
            (ro, so, co), (rh, sh, ch) = _taps(L, H, o, h)
            assert np.array_equal(ro, rh), (game, i, "ram")
            assert np.array_equal(so, sh), (game, i, "screen", int((so != sh).sum()))
            assert np.array_equal(co, ch), (game, i, "cpu")
            assert L.orc_game_over(o) == H.he_game_over(h)
            if L.orc_game_over(o):
                L.orc_reset_game(o); H.he_reset_game(h, 0, 0)
    L.orc_destroy(o); H.he_destroy(h)


@pytest.mark.parametrize("game", GAMES12)
def test_next_with_pixel_less_frames_matches_oracle(libs, game):
    """next() by next() with only the two pooled frames drawn (the product's mode): both frame buffers,
    RAM, reward, terminal must still be exact, including across resets with start no-ops."""
    L, H = libs
    rom = rom_bytes(game)
    o = L.orc_create(rom, len(rom), game.encode(), 9)
    h = H.he_create(rom, len(rom), game.encode(), 9, 1)
    n = L.orc_num_actions(o)
    acts = np.zeros(18, np.int32)
    L.orc_minimal_actions(o, acts.ctypes.data)
    rng = np.random.RandomState(2)
    both = np.zeros((2, 33600), np.uint8)
    oboth = np.zeros((2, 33600), np.uint8)
    for i in range(220):
        a = int(acts[rng.randint(n)])
        want_r, grabs = 0, []
        for f in range(4):
            want_r += L.orc_act(o, a)
            if f >= 2:
                g = np.zeros(33600, np.uint8)
                L.orc_get_screen(o, g.ctypes.data)
                grabs.append(g)
        single = C.c_int()
        got_r = H.he_next(h, a, 1, C.byref(single))
        assert want_r == got_r, (game, i)
        (ro, so, co), (rh, sh, ch) = _taps(L, H, o, h)
        assert np.array_equal(ro, rh) and np.array_equal(co, ch), (game, i)
        assert np.array_equal(so, sh), (game, i, "current screen")
        H.he_get_both_screens(h, both.ctypes.data)
        L.orc_get_both_screens(o, oboth.ctypes.data)
        want_max = np.maximum(grabs[0], grabs[1])
        got_max = sh if single.value else np.maximum(both[0], both[1])
        assert np.array_equal(want_max, got_max), (game, i, "pooled frames")
        if not L.orc_game_over(o):
            assert np.array_equal(both, oboth), (game, i, "both frame buffers")
        assert L.orc_game_over(o) == H.he_game_over(h)
        if L.orc_game_over(o) or i % 70 == 69:
            noops = int(rng.randint(0, 31))
            L.orc_reset_game(o)
            for _ in range(noops):
                L.orc_act(o, 0)
            H.he_reset_game(h, noops, 1)
            (ro, so, co), (rh, sh, ch) = _taps(L, H, o, h)
            assert np.array_equal(ro, rh) and np.array_equal(so, sh) and np.array_equal(co, ch), (game, i, "reset")
            H.he_get_both_screens(h, both.ctypes.data)
            L.orc_get_both_screens(o, oboth.ctypes.data)
            assert np.array_equal(both, oboth), (game, i, "both frame buffers after reset")
    L.orc_destroy(o); H.he_destroy(h)


@pytest.mark.parametrize("flat", [0, 1])
@pytest.mark.parametrize("game", GAMES12)
def test_fast_tick_equals_general_path(libs, game, flat):
    """cpu_fast (the branch-free tick the kernels try first) against cpu_step (the general path, the definition) on
    EVERY instruction the fast tick accepts: registers, status, cycles, data bus, RIOT RAM, the TIA write FIFO and the
    untouched rest of the machine.  The host build aborts on the first difference (he_set_fast_mode(2)).
    flat = 1: the flat-cartridge-window instantiation the kernels use for every cartridge type but E0."""
    L, H = libs
    H.he_set_fast_flat(flat)
    H.he_fast_stats.argtypes = [C.c_void_p]
    rom = rom_bytes(game)
    H.he_set_fast_mode(2)
    try:
        h = H.he_create(rom, len(rom), game.encode(), 11, 1)
        n = L.orc_num_actions(L.orc_create(rom, len(rom), game.encode(), 11))
        acts = np.zeros(18, np.int32)
        o = L.orc_create(rom, len(rom), game.encode(), 11)
        L.orc_minimal_actions(o, acts.ctypes.data)
        L.orc_destroy(o)
        rng = np.random.RandomState(5)
        s0 = np.zeros(2, np.uint64); H.he_fast_stats(s0.ctypes.data)
        for i in range(120):
            H.he_next(h, int(acts[rng.randint(n)]), 1, None)
            if H.he_game_over(h):
                H.he_reset_game(h, int(rng.randint(0, 31)), 1)
        s1 = np.zeros(2, np.uint64); H.he_fast_stats(s1.ctypes.data)
        taken, refused = int(s1[0] - s0[0]), int(s1[1] - s0[1])
        assert taken > 1000000 and refused < 0.08 * taken, (game, taken, refused)   # the fast tick is the common case
        H.he_destroy(h)
    finally:
        H.he_set_fast_mode(1)
        H.he_set_fast_flat(0)


@pytest.mark.parametrize("mode", [0, 2])
def test_general_path_alone_and_checked_fast_tick_match_oracle(libs, mode):
    """The oracle comparison of test_every_frame_matches_oracle with the fast tick switched off (mode 0: the general
    path must stay exact on its own -- it is what refused instructions and the reset probe run) and in checked mode."""
    L, H = libs
    H.he_set_fast_mode(mode)
    try:
        for game in ("breakout", "ms_pacman", "montezuma_revenge"):
            rom = rom_bytes(game)
            o = L.orc_create(rom, len(rom), game.encode(), 4)
            h = H.he_create(rom, len(rom), game.encode(), 4, 0)
            n = L.orc_num_actions(o)
            acts = np.zeros(18, np.int32)
            L.orc_minimal_actions(o, acts.ctypes.data)
            rng = np.random.RandomState(3)
            for i in range(160):
                a = int(acts[rng.randint(n)])
                assert L.orc_act(o, a) == H.he_act(h, a), (game, i)
                (ro, so, co), (rh, sh, ch) = _taps(L, H, o, h)
                assert np.array_equal(ro, rh) and np.array_equal(so, sh) and np.array_equal(co, ch), (game, i, mode)
            L.orc_destroy(o); H.he_destroy(h)
    finally:
        H.he_set_fast_mode(1)
