"""The CPU oracle (oracle/host_path.py over oracle/liborc.so) against the committed golden fixtures, which
were produced by the REFERENCE'S OWN modules (atari_emulator.py / environment.py / emulator_runner.py /
exploration_policy.py imported unmodified, tests/golden/make_golden.py).  Pins the port before it is
trusted as the checker of the CUDA path."""
import numpy as np
import pytest

import util
from util import GAMES12, GOLDEN, OraclePool, host_path


def test_resize_lut_matches_pil_fixture():
    d = np.load(GOLDEN + "/resize_lut.npz")
    assert np.array_equal(d["xmap"], host_path.XMAP) and np.array_equal(d["ymap"], host_path.YMAP)
    # rows are floor((y + 0.5) * 2.5); columns are NOT the closed form at x = 52 and x = 73
    assert np.array_equal(host_path.YMAP, np.floor((np.arange(84) + 0.5) * 2.5).astype(int))
    closed = np.floor((np.arange(84) + 0.5) * 160 / 84).astype(int)
    assert list(np.nonzero(closed != host_path.XMAP)[0]) == [52, 73]


def test_tab_rep_tables():
    d = np.load(GOLDEN + "/tab_rep.npz")
    for key in d.files:
        mr, k = [int(x) for x in key.split("_")]
        assert list(d[key]) == host_path.tab_repetitions(mr, k), key
    assert host_path.tab_repetitions(10, 11) == list(range(11))          # FiGAR10, README.md:135
    assert host_path.tab_repetitions(10, 6) == [0, 2, 4, 6, 8, 10]
    assert host_path.tab_repetitions(0, 1) == [0]                        # plain PAAC


def test_action_sets_and_lives():
    d = np.load(GOLDEN + "/action_sets.npz")
    for g in GAMES12:
        e = host_path.PortAtariEmulator(0, util.args_for(g))
        assert list(e.get_legal_actions()) == list(d[g]), g
        assert e.lives == int(d[g + "_lives"]), g


@pytest.mark.parametrize("game", ["pong", "breakout", "seaquest", "ms_pacman"])
def test_figar_fixture(game):
    """Rewards, terminals, CRC32 of every returned state and the full final states of the reference's
    EmulatorRunner._run loop."""
    d = np.load("%s/figar_%s.npz" % (GOLDEN, game))
    acts, reps = d["actions"], d["repetitions"]
    m, n = acts.shape
    rgb = d["final_states"].shape[-1] == 12
    tab = list(d["tab_rep"])
    ora = OraclePool(game, n, rgb=rgb, nb_choices=len(tab), max_repetition=max(tab))
    assert ora.tab_rep == tab and ora.num_actions == int(d["num_actions"])
    st = ora.initial_states()
    assert [util.crc(st[e]) for e in range(n)] == list(d["init_state_crc"])
    steps = min(m, 60) if game == "breakout" else m
    for t in range(steps):
        st, rw, tm, _ = ora.macro_step(acts[t], reps[t])
        assert np.array_equal(rw, d["rewards"][t]) and np.array_equal(tm, d["terminals"][t]), t
        assert [util.crc(st[e]) for e in range(n)] == list(d["state_crc"][t]), t
    if steps == m:
        assert np.array_equal(st, d["final_states"])


def test_preprocess_from_indices_equals_reference_order():
    """The device keeps raw palette indices and converts late; the reference converts each grabbed screen and
    then takes the max: same result (max is taken in luminance / per RGB channel either way)."""
    rng = np.random.RandomState(0)
    a = (rng.randint(0, 128, size=(210, 160)) * 2).astype(np.uint8)
    b = (rng.randint(0, 128, size=(210, 160)) * 2).astype(np.uint8)
    gray, col = host_path.palettes()
    for rgb in (False, True):
        if rgb:
            pool = np.stack([col[a >> 1], col[b >> 1]])
        else:
            pool = np.stack([gray[a >> 1][..., None], gray[b >> 1][..., None]])
        assert np.array_equal(host_path.process_frame_pool(pool), host_path.preprocess_indices(a, b, rgb))


def test_observation_ring_channel_order():
    """environment.py:58-80: channel c = d * 4 + k, k oldest -> newest."""
    ring = host_path.ObsRing(3)
    for i in range(6):
        ring.push(np.full((84, 84, 3), [10 * i + 1, 10 * i + 2, 10 * i + 3], np.uint8))
    s = ring.stacked()
    assert s.shape == (84, 84, 12)
    assert list(s[0, 0]) == [21, 31, 41, 51, 22, 32, 42, 52, 23, 33, 43, 53]


def test_nstep_closed_forms():
    """paac.py:226-231."""
    T, n, g = 5, 3, 0.99
    r = np.ones((T, n), np.float32)
    z = np.zeros((T, n), np.float32)
    boot = np.full(n, 2.0, np.float32)
    y, adv = host_path.nstep_returns(r, z, z, boot, g)
    for t in range(T):
        want = sum(g ** j for j in range(T - t)) + g ** (T - t) * 2.0
        assert np.allclose(y[t], want, rtol=1e-12)
    term = z.copy(); term[2] = 1.0            # episode ends at t = 2: nothing flows back across it
    y, _ = host_path.nstep_returns(r, term, z, boot, g)
    assert np.allclose(y[2], 1.0) and np.allclose(y[1], 1.0 + g) and np.allclose(y[3], 1 + g * (1 + g * 2.0))
    big = np.full((T, n), 7.0, np.float32)    # clip after summing the repeats (paac.py:180)
    y, _ = host_path.nstep_returns(big, z, z, np.zeros(n, np.float32), 1.0)
    assert np.allclose(y[0], 5.0)
    y, adv = host_path.nstep_returns(-big, z, np.full((T, n), 0.5, np.float32), np.zeros(n, np.float32), 1.0)
    assert np.allclose(y[0], -5.0) and np.allclose(adv[0], -5.5)


def test_philox_known_answers():
    """Random123 known-answer vectors for Philox-4x32-10."""
    out = host_path.philox4x32(0, 0, 0, 0, 0, 0)
    assert [int(x) for x in out] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    out = host_path.philox4x32(0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff)
    assert [int(x) for x in out] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    out = host_path.philox4x32(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0)
    assert [int(x) for x in out] == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_sampler_distribution_and_edges():
    """exploration_policy.py:96-116 restated: multinomial follows the probabilities, argmax / e-greedy edge cases."""
    n = 40000
    p = np.tile(np.array([[0.1, 0.2, 0.3, 0.4]], np.float32), (n, 1))
    q = np.tile(np.array([[0.5, 0.5]], np.float32), (n, 1))
    a, r, ah, rh = host_path.choose_next_actions(p, q, 0, seed=9, step=3)
    freq = np.bincount(a, minlength=4) / n
    assert np.allclose(freq, [0.1, 0.2, 0.3, 0.4], atol=0.01)
    assert ah.shape == (n, 4) and rh.shape == (n, 2) and np.all(ah.sum(1) == 1) and np.all(ah[np.arange(n), a] == 1)
    one = np.tile(np.array([[0.0, 1.0, 0.0]], np.float32), (100, 1))
    a, _, _, _ = host_path.choose_next_actions(one, one, 0, seed=1, step=0)
    assert np.all(a == 1)
    a, r, _, _ = host_path.choose_next_actions(p, q, 2, seed=1, step=0)
    assert np.all(a == 3) and np.all(r == 0)             # argmax, first maximum on ties
    a, _, _, _ = host_path.choose_next_actions(p, q, 1, seed=1, step=0, eps=0.0)
    assert np.all(a == 3)
    a, _, _, _ = host_path.choose_next_actions(p, q, 1, seed=1, step=0, eps=1.0)
    assert np.allclose(np.bincount(a, minlength=4) / n, 0.25, atol=0.01)
