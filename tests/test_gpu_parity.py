"""GPU parity: the CUDA environment pool (through the C ABI) against the CPU oracle on the same ROMs,
seeds and action sequences -- RAM, raw screen, reward, terminal bit-exact; stacked states exact."""
import numpy as np
import pytest

import util
from util import GAMES12, OraclePool, rom_bytes

pytestmark = pytest.mark.gpu


def _device_pool(game, n, **kw):
    import manette_b200 as mb
    return mb.DevicePool([(game, rom_bytes(game), n)], **kw)


@pytest.mark.parametrize("game", GAMES12)
def test_single_env_next_matches_oracle(game):
    """AtariEmulator.next() one call at a time: every tap after every call."""
    n, steps = 3, 40
    ora = OraclePool(game, n)
    pool = _device_pool(game, n)
    try:
        assert list(pool.legal_actions(0)) == list(ora.emus[0].get_legal_actions())
        want = ora.initial_states()
        pool.reset_all()
        got = pool.states.cpu().numpy()
        assert np.array_equal(got, want), "initial states differ"
        rng = np.random.RandomState(7)
        for t in range(steps):
            for e in range(n):
                a = int(rng.randint(ora.num_actions))
                obs, rew, term = ora.emus[e].next(a)
                r2, t2 = pool.env_next(e, a)
                assert (rew, bool(term)) == (r2, t2), (game, t, e)
                assert np.array_equal(pool.ram(e), ora.emus[e].ale.getRAM()), (game, t, e, "ram")
                assert np.array_equal(pool.screen(e), ora.emus[e].ale.getScreen()), (game, t, e, "screen")
                assert np.array_equal(pool.states[e].cpu().numpy(), obs), (game, t, e, "state")
                cpu = ora.emus[e].ale.getCPU()
                assert np.array_equal(pool.cpu_state(e)[:7], cpu[:7]), (game, t, e, "cpu")
                if term:
                    want0 = ora.emus[e].get_initial_state()
                    pool.env_reset(e)
                    assert np.array_equal(pool.states[e].cpu().numpy(), want0)
    finally:
        pool.close()


@pytest.mark.parametrize("game,rgb,k,max_rep,n,steps", [
    ("pong", False, 1, 0, 8, 30),
    ("breakout", False, 11, 10, 16, 25),
    ("seaquest", True, 11, 10, 8, 12),
    ("ms_pacman", False, 6, 10, 8, 12),
])
def test_macro_step_matches_oracle(game, rgb, k, max_rep, n, steps):
    """Runners.update_environments(): FiGAR repeat loop with early exit and in-step reset."""
    ora = OraclePool(game, n, rgb=rgb, nb_choices=k, max_repetition=max_rep)
    pool = _device_pool(game, n, rgb=rgb, tab_rep=ora.tab_rep)
    try:
        import torch
        assert np.array_equal(ora.initial_states(), (pool.reset_all(), pool.states.cpu().numpy())[1])
        acts, reps = util.schedule(99, steps, n, ora.num_actions, k)
        for t in range(steps):
            want_s, want_r, want_t, want_c = ora.macro_step(acts[t], reps[t])
            pool.action_idx.copy_(torch.as_tensor(acts[t].astype(np.int32)))
            pool.repetition_idx.copy_(torch.as_tensor(reps[t].astype(np.int32)))
            pool.step_async(use_indices=True)
            pool.wait()
            assert np.array_equal(pool.rewards.cpu().numpy(), want_r), (game, t)
            assert np.array_equal(pool.terminals.cpu().numpy(), want_t), (game, t)
            assert np.array_equal(pool.next_calls.cpu().numpy(), want_c), (game, t)
            assert np.array_equal(pool.states.cpu().numpy(), want_s), (game, t)
        for e in range(n):
            assert np.array_equal(pool.ram(e), ora.emus[e].ale.getRAM())
            assert np.array_equal(pool.screen(e), ora.emus[e].ale.getScreen())
    finally:
        pool.close()


@pytest.mark.parametrize("game", ["pong", "breakout", "seaquest", "ms_pacman"])
def test_golden_fixture(game):
    """The committed fixtures were produced by the reference's own EmulatorRunner._run / AtariEmulator code
    (tests/golden/make_golden.py); the device path must reproduce them through one-hot inputs."""
    import torch
    d = np.load("%s/figar_%s.npz" % (util.GOLDEN, game))
    acts, reps = d["actions"], d["repetitions"]
    m, n = acts.shape
    rgb = d["final_states"].shape[-1] == 12
    pool = _device_pool(game, n, rgb=rgb, tab_rep=list(d["tab_rep"]))
    try:
        pool.reset_all()
        st = pool.states.cpu().numpy()
        assert [util.crc(st[e]) for e in range(n)] == list(d["init_state_crc"])
        a_n, k_n = int(d["num_actions"]), len(d["tab_rep"])
        assert (pool.num_actions, pool.nb_choices) == (a_n, k_n)
        for t in range(m):
            pool.actions.copy_(torch.as_tensor(np.eye(a_n, dtype=np.float32)[acts[t]]))
            pool.repetitions.copy_(torch.as_tensor(np.eye(k_n, dtype=np.float32)[reps[t]]))
            pool.step_async(use_indices=False)
            pool.wait()
            assert np.array_equal(pool.rewards.cpu().numpy(), d["rewards"][t]), t
            assert np.array_equal(pool.terminals.cpu().numpy(), d["terminals"][t]), t
            st = pool.states.cpu().numpy()
            assert [util.crc(st[e]) for e in range(n)] == list(d["state_crc"][t]), t
        assert np.array_equal(st, d["final_states"])
    finally:
        pool.close()


@pytest.mark.parametrize("game", ["pong", "yars_revenge", "seaquest"])
def test_memoised_resets_match_oracle(game):
    """get_initial_state() restored from the reset memo (same RIOT timer seed seen before) must equal the
    emulated one: many resets of a few envs, every one compared with the oracle, and hits must have happened."""
    n = 4
    ora = OraclePool(game, n)
    pool = _device_pool(game, n)
    try:
        assert np.array_equal(ora.initial_states(), (pool.reset_all(), pool.states.cpu().numpy())[1])
        rng = np.random.RandomState(11)
        for rep in range(45):
            for e in range(n):
                for _ in range(int(rng.randint(0, 3))):
                    a = int(rng.randint(ora.num_actions))
                    obs, rew, term = ora.emus[e].next(a)
                    r2, t2 = pool.env_next(e, a)
                    assert (rew, bool(term)) == (r2, t2)
                    assert np.array_equal(pool.states[e].cpu().numpy(), obs)
                want = ora.emus[e].get_initial_state()
                pool.env_reset(e)
                assert np.array_equal(pool.states[e].cpu().numpy(), want), (game, rep, e, "state")
                assert np.array_equal(pool.ram(e), ora.emus[e].ale.getRAM()), (game, rep, e, "ram")
                assert np.array_equal(pool.screen(e), ora.emus[e].ale.getScreen()), (game, rep, e, "screen")
                assert np.array_equal(pool.cpu_state(e)[:7], ora.emus[e].ale.getCPU()[:7]), (game, rep, e, "cpu")
        hits, misses, stored = pool.memo_stats()
        assert hits + misses == n * 46 and stored >= 1
        if game != "yars_revenge":
            assert hits > 40, (hits, misses, stored)
        else:
            # its four start frames read a pseudo-random byte the game keeps across resets: the whole-segment key almost
            # never recurs, but the reset unit itself (level 1 of the memo) depends on two bytes with a handful of values
            assert pool.memo_level1_hits() > 10, (hits, misses, stored, pool.memo_level1_hits())
    finally:
        pool.close()


def test_memo_off_equals_memo_on():
    """The memo is an optimisation only: identical results with it disabled."""
    import manette_b200 as mb
    import torch
    game, n, k = "breakout", 32, 11
    tab = list(range(k))
    pools = [mb.DevicePool([(game, rom_bytes(game), n)], tab_rep=tab, reset_memo=m, draw_all_frames=not m) for m in (True, False)]
    try:
        for p in pools:
            p.reset_all()
        acts, reps = util.schedule(3, 40, n, 4, k)
        for t in range(40):
            outs = []
            for p in pools:
                p.action_idx.copy_(torch.as_tensor(acts[t].astype(np.int32)))
                p.repetition_idx.copy_(torch.as_tensor(reps[t].astype(np.int32)))
                p.step_async(use_indices=True)
                p.wait()
                outs.append((p.states.cpu().numpy(), p.rewards.cpu().numpy(), p.terminals.cpu().numpy(), p.frames.cpu().numpy()))
            for a, b in zip(outs[0], outs[1]):
                assert np.array_equal(a, b), t
        assert pools[0].memo_stats()[0] > 0 and pools[1].memo_stats()[0] == 0
    finally:
        for p in pools:
            p.close()


@pytest.mark.parametrize("game", ["breakout", "ms_pacman"])
def test_warmed_reset_memo_restores_every_first_reset(game):
    """mn_reset_all warms the reset memo with all 75 timer seeds (envs lent and put back): every get_initial_state of
    the pool -- the first ones included -- is then restored by copy, and must equal the oracle's emulated one."""
    n, k = 80, 11
    ora = util.ParallelOraclePool(game, range(n), nb_choices=k, max_repetition=10)
    pool = _device_pool(game, n, tab_rep=ora.tab_rep)
    try:
        import torch
        pool.reset_all()
        assert np.array_equal(pool.states.cpu().numpy(), ora.initial_states())
        hits, misses, stored = pool.memo_stats()
        assert (hits, misses) == (n, 0) and stored == 75, (hits, misses, stored)
        rng = np.random.RandomState(5)
        for t in range(6):
            acts, reps = rng.randint(0, ora.num_actions, n), rng.randint(0, k, n)
            out = ora.macro_step(acts, reps, taps=True)
            pool.action_idx.copy_(torch.as_tensor(acts.astype(np.int32)))
            pool.repetition_idx.copy_(torch.as_tensor(reps.astype(np.int32)))
            pool.step_async(use_indices=True)
            pool.wait()
            assert np.array_equal(pool.states.cpu().numpy(), out[0]), t
            assert np.array_equal(pool.rewards.cpu().numpy(), out[1]) and np.array_equal(pool.terminals.cpu().numpy(), out[2])
        for e in range(0, n, 7):
            assert np.array_equal(pool.ram(e), out[4][e]) and np.array_equal(pool.screen(e), out[5][e])
    finally:
        pool.close()
        ora.close()


def test_random_start_resets_are_memoised_exactly():
    """random_start: the start no-op count (0..30) is part of the reset memo's key.  Single-life Breakout ends episodes
    every few steps: hundreds of resets, so (timer seed, no-op count) pairs recur and are restored by copy -- every
    state must still equal the oracle's, which emulates every reset."""
    import manette_b200 as mb
    import torch
    game, n, k, seed, steps = "breakout", 48, 11, 3, 70
    ora = OraclePool(game, n, nb_choices=k, max_repetition=10, single_life=True, random_start=True, seed=seed,
                     noops=lambda gid, ep: mb.start_noops(seed, gid, ep))
    pool = mb.DevicePool([(game, rom_bytes(game), n)], tab_rep=ora.tab_rep, single_life_episodes=True, random_start=True,
                         random_seed=seed)
    try:
        pool.reset_all()
        assert np.array_equal(pool.states.cpu().numpy(), ora.initial_states())
        acts, reps = util.schedule(31, steps, n, ora.num_actions, k)
        for t in range(steps):
            ws, wr, wt, _ = ora.macro_step(acts[t], reps[t])
            pool.action_idx.copy_(torch.as_tensor(acts[t].astype(np.int32)))
            pool.repetition_idx.copy_(torch.as_tensor(reps[t].astype(np.int32)))
            pool.step_async(use_indices=True)
            pool.wait()
            assert np.array_equal(pool.rewards.cpu().numpy(), wr) and np.array_equal(pool.terminals.cpu().numpy(), wt), t
            assert np.array_equal(pool.states.cpu().numpy(), ws), t
        hits, misses, stored = pool.memo_stats()
        assert hits + misses > 300 and hits > 5, (hits, misses, stored)
        for e in range(0, n, 5):
            assert np.array_equal(pool.ram(e), ora.emus[e].ale.getRAM())
            assert np.array_equal(pool.screen(e), ora.emus[e].ale.getScreen())
    finally:
        pool.close()
