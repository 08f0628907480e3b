"""GPU parity over LONG runs (the oracle side on all host cores, tests/util.py ParallelOraclePool):
* every README game played under a random FiGAR policy until the pool has seen game over at least twice -- rewards,
  terminals, next() counts and stacked states after every macro step; RAM, raw screen, ALE lives and the AtariEmulator's
  own life counter (mn_get_lives) of every environment every few steps and after every step with a terminal
  (atari_emulator.py:120-121,126-133), with and without single_life_episodes;
* BASELINE config 2 step by step: Breakout FiGAR10, 256 environments, 200 macro steps (emulator_runner.py:19-42);
* 64 environments sampled from a decorrelated 16,384-environment Ms Pacman pool (full 32-lane... 28-lane warps on every SM,
  hand-offs, time-synchronisation slack) followed by the oracle through every macro step."""
import numpy as np
import pytest
import torch

import util
from util import GAMES12, ParallelOraclePool, rom_bytes

pytestmark = pytest.mark.gpu

def _step_async(pool, acts, reps):
    pool.action_idx.copy_(torch.as_tensor(np.asarray(acts, np.int32)))
    pool.repetition_idx.copy_(torch.as_tensor(np.asarray(reps, np.int32)))
    pool.step_async(use_indices=True)


def _step(pool, acts, reps):
    _step_async(pool, acts, reps)
    pool.wait()


def _taps_equal(pool, ids, ram, screen, ale_lives, host_lives, where):
    for j, e in enumerate(ids):
        assert np.array_equal(pool.ram(e), ram[j]), where + (e, "ram")
        assert np.array_equal(pool.screen(e), screen[j]), where + (e, "screen")
        lv, over, _ = pool.lives(e)
        assert lv == int(ale_lives[j]), where + (e, "ale.lives()", lv, int(ale_lives[j]))
        assert not over, where + (e, "game over outside a macro step")   # episodes are reset inside the step


@pytest.mark.parametrize("single_life", [False, True])
def test_games_played_to_game_over(single_life):
    """All games in ONE mixed pool (a round of the emulation kernel costs the same for 8 environments as for 96), the
    oracle side one process group per game; the run goes on until EVERY game has ended at least twice (a random policy
    needs ~550 macro steps to lose a day of Enduro) -- games that end sooner simply keep being compared."""
    import manette_b200 as mb
    per, k, max_rep, want = 8, 11, 10, 2
    # without lives `single_life_episodes` changes nothing: Pong and Enduro are covered by the other variant
    games = [g for g in GAMES12 if not (single_life and g in ("pong", "enduro"))]
    cap = 1500 if not single_life else 400
    oras = [ParallelOraclePool(g, range(i * per, (i + 1) * per), workers=2, nb_choices=k, max_repetition=max_rep,
                               single_life=single_life) for i, g in enumerate(games)]
    pool = mb.DevicePool([(g, rom_bytes(g), per) for g in games], tab_rep=oras[0].tab_rep, single_life_episodes=single_life)
    n = per * len(games)
    try:
        pool.reset_all()
        got = pool.states.cpu().numpy()
        for i, (g, o) in enumerate(zip(games, oras)):
            assert np.array_equal(got[i * per:(i + 1) * per], o.initial_states()), (g, "initial states")
        rng = np.random.RandomState(11)
        n_act = np.repeat([o.num_actions for o in oras], per)
        terminals = np.zeros(len(games), np.int64)
        t = 0
        while t < cap and (terminals < want).any():
            acts, reps = rng.randint(0, 1 << 30, n) % n_act, rng.randint(0, k, n)
            _step_async(pool, acts, reps)                    # the GPU runs while the oracle processes do
            outs = [o.macro_step(acts[i * per:(i + 1) * per], reps[i * per:(i + 1) * per], taps=True) for i, o in enumerate(oras)]
            pool.wait()
            rw, tm, nc, st = (x.cpu().numpy() for x in (pool.rewards, pool.terminals, pool.next_calls, pool.states))
            for i, (g, out) in enumerate(zip(games, outs)):
                sl = slice(i * per, (i + 1) * per)
                assert np.array_equal(rw[sl], out[1]), (g, t, "reward")
                assert np.array_equal(tm[sl], out[2]), (g, t, "terminal")
                assert np.array_equal(nc[sl], out[3]), (g, t, "next() calls")
                assert np.array_equal(st[sl], out[0]), (g, t, "states")
                if t % 16 == 0 or out[2].any():
                    _taps_equal(pool, range(i * per, (i + 1) * per), out[4], out[5], out[6], out[7], (g, t))
                terminals[i] += int(out[2].sum())
            t += 1
        assert (terminals >= want).all(), dict(zip(games, terminals.tolist()))
    finally:
        pool.close()
        for o in oras:
            o.close()


def test_breakout_figar10_256_envs_200_macro_steps():
    """BASELINE.json config 2, step by step."""
    import manette_b200 as mb
    game, n, k, max_rep, steps = "breakout", 256, 11, 10, 200
    ora = ParallelOraclePool(game, range(n), nb_choices=k, max_repetition=max_rep)
    pool = mb.DevicePool([(game, rom_bytes(game), n)], tab_rep=ora.tab_rep)
    try:
        pool.reset_all()
        assert np.array_equal(pool.states.cpu().numpy(), ora.initial_states())
        rng = np.random.RandomState(256)
        terminals = 0
        for t in range(steps):
            acts, reps = rng.randint(0, ora.num_actions, n), rng.randint(0, k, n)
            last = (t == steps - 1)
            out = ora.macro_step(acts, reps, taps=last)
            _step(pool, acts, reps)
            assert np.array_equal(pool.rewards.cpu().numpy(), out[1]), t
            assert np.array_equal(pool.terminals.cpu().numpy(), out[2]), t
            assert np.array_equal(pool.next_calls.cpu().numpy(), out[3]), t
            assert np.array_equal(pool.states.cpu().numpy(), out[0]), t
            terminals += int(out[2].sum())
        _taps_equal(pool, range(n), out[4], out[5], out[6], out[7], (game, steps))
        assert terminals > 50       # the run is long enough to cycle through many episodes
    finally:
        pool.close()
        ora.close()


def test_sampled_envs_of_a_16384_env_pool_follow_the_oracle():
    """Full warps on every SM, decorrelated: 64 sampled environments against the oracle through every macro step."""
    import manette_b200 as mb
    game, n, k, max_rep, steps = "ms_pacman", 16384, 11, 10, 34
    ids = [(i * 16384) // 64 + (i * 7) % 61 for i in range(64)]       # every warp region, varying lanes
    ora = ParallelOraclePool(game, ids, nb_choices=k, max_repetition=max_rep)
    pool = mb.DevicePool([(game, rom_bytes(game), n)], tab_rep=ora.tab_rep)
    try:
        pool.reset_all()
        assert np.array_equal(pool.states[ids].cpu().numpy(), ora.initial_states())
        g = torch.Generator().manual_seed(16384)
        na = ora.num_actions
        for t in range(steps):
            acts = torch.randint(0, na, (n,), generator=g, dtype=torch.int32)
            reps = torch.randint(0, k, (n,), generator=g, dtype=torch.int32)
            last = (t >= steps - 10)
            out = ora.macro_step(acts[ids].numpy(), reps[ids].numpy(), taps=last)
            _step(pool, acts.numpy(), reps.numpy())
            assert np.array_equal(pool.rewards[ids].cpu().numpy(), out[1]), t
            assert np.array_equal(pool.terminals[ids].cpu().numpy(), out[2]), t
            assert np.array_equal(pool.next_calls[ids].cpu().numpy(), out[3]), t
            assert np.array_equal(pool.states[ids].cpu().numpy(), out[0]), t
            if last:
                _taps_equal(pool, ids, out[4], out[5], out[6], out[7], (game, t))
    finally:
        pool.close()
        ora.close()
