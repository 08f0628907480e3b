"""GPU, BASELINE-size pools: size-independent properties where the oracle cannot follow in seconds."""
import numpy as np
import pytest

import util
from util import rom_bytes

pytestmark = pytest.mark.gpu


def _step(pool, acts, reps):
    import torch
    pool.action_idx.copy_(torch.as_tensor(acts.astype(np.int32)))
    pool.repetition_idx.copy_(torch.as_tensor(reps.astype(np.int32)))
    pool.step_async(use_indices=True)
    pool.wait()


def test_replicated_envs_stay_identical_at_full_size():
    """4096 Seaquest RGB envs (BASELINE config 3) fed the same (action, repetition) every step must remain
    bit-identical to each other whatever warp / lane / block they sit in -- and to a 2-env pool driven alike."""
    import manette_b200 as mb
    n, k = 4096, 11
    tab = list(range(k))
    big = mb.DevicePool([("seaquest", rom_bytes("seaquest"), n)], rgb=True, tab_rep=tab, random_seed=0)
    small = mb.DevicePool([("seaquest", rom_bytes("seaquest"), 2)], rgb=True, tab_rep=tab, random_seed=0, draw_all_frames=True,
                          reset_memo=False)
    try:
        big.reset_all(); small.reset_all()
        rng = np.random.RandomState(5)
        for t in range(6):
            a, r = int(rng.randint(18)), int(rng.randint(k))
            _step(big, np.full(n, a), np.full(n, r))
            _step(small, np.full(2, a), np.full(2, r))
            st = big.states
            assert bool((st == st[0:1]).all()), t
            assert bool((big.rewards == big.rewards[0]).all()) and bool((big.next_calls == r + 1).all() or bool(big.terminals.any()))
            assert np.array_equal(st[0].cpu().numpy(), small.states[0].cpu().numpy()), t
            assert float(big.rewards[0]) == float(small.rewards[0])
    finally:
        big.close(); small.close()


def test_sharded_pools_equal_one_pool():
    """Environments shard across GPUs by contiguous global ids with no exchange: two half pools with
    env_id_offset reproduce one whole pool (random_start on, so the per-env schedules matter)."""
    import manette_b200 as mb
    n, k = 64, 11
    tab = list(range(k))
    rom = rom_bytes("breakout")
    whole = mb.DevicePool([("breakout", rom, n)], tab_rep=tab, random_start=True)
    halves = [mb.DevicePool([("breakout", rom, n // 2)], tab_rep=tab, random_start=True, env_id_offset=o) for o in (0, n // 2)]
    try:
        whole.reset_all()
        for h in halves:
            h.reset_all()
        acts, reps = util.schedule(8, 30, n, 4, k)
        for t in range(30):
            _step(whole, acts[t], reps[t])
            for i, h in enumerate(halves):
                sl = slice(i * n // 2, (i + 1) * n // 2)
                _step(h, acts[t][sl], reps[t][sl])
                assert np.array_equal(h.states.cpu().numpy(), whole.states[sl].cpu().numpy()), (t, i)
                assert np.array_equal(h.rewards.cpu().numpy(), whole.rewards[sl].cpu().numpy())
                assert np.array_equal(h.terminals.cpu().numpy(), whole.terminals[sl].cpu().numpy())
    finally:
        whole.close()
        for h in halves:
            h.close()


def test_envs_per_warp_does_not_change_results():
    import manette_b200 as mb
    n, k = 96, 11
    tab = list(range(k))
    rom = rom_bytes("ms_pacman")
    pools = [mb.DevicePool([("ms_pacman", rom, n)], tab_rep=tab, envs_per_warp=e) for e in (1, 4, 32)]
    try:
        for p in pools:
            p.reset_all()
        acts, reps = util.schedule(2, 8, n, 9, k)
        for t in range(8):
            outs = []
            for p in pools:
                _step(p, acts[t], reps[t])
                outs.append((p.states.cpu().numpy(), p.rewards.cpu().numpy(), p.frames.cpu().numpy()))
            for o in outs[1:]:
                assert all(np.array_equal(x, y) for x, y in zip(o, outs[0])), t
    finally:
        for p in pools:
            p.close()


def test_mixed_game_pool_matches_single_game_pools():
    """BASELINE config 5 in miniature: 12 games in one pool, grouped by cartridge."""
    import manette_b200 as mb
    per, k = 3, 11
    tab = list(range(k))
    games = util.GAMES12
    mixed = mb.DevicePool([(g, rom_bytes(g), per) for g in games], tab_rep=tab)
    try:
        mixed.reset_all()
        assert mixed.num_actions == 18
        acts, reps = util.schedule(6, 5, per * len(games), 4, k)   # actions 0..3 are legal everywhere
        singles = []
        for gi, g in enumerate(games):
            p = mb.DevicePool([(g, rom_bytes(g), per)], tab_rep=tab, env_id_offset=gi * per)
            p.reset_all()
            singles.append(p)
        for t in range(5):
            _step(mixed, acts[t], reps[t])
            for gi, p in enumerate(singles):
                sl = slice(gi * per, (gi + 1) * per)
                _step(p, acts[t][sl], reps[t][sl])
                assert np.array_equal(p.states.cpu().numpy(), mixed.states[sl].cpu().numpy()), (games[gi], t)
                assert np.array_equal(p.rewards.cpu().numpy(), mixed.rewards[sl].cpu().numpy())
        for p in singles:
            p.close()
    finally:
        mixed.close()


def test_host_states_mirror_follows_every_publication():
    """mn_set_host_states: a registered pinned array receives every state the pool publishes -- at reset, during the
    macro step as envs finish their repeats (k_emit_early), and at its end (last round, terminals after their reset) --
    and is exactly the device array once the step has been waited for.  Pageable memory is refused; unregistering stops
    the writes."""
    import torch
    import manette_b200 as mb
    n, k = 96, 11
    pool = mb.DevicePool([("breakout", rom_bytes("breakout"), 64), ("pong", rom_bytes("pong"), 32)], tab_rep=list(range(k)))
    try:
        with pytest.raises(ValueError):
            pool.set_host_states(torch.zeros(tuple(pool.states.shape), dtype=torch.uint8))
        pinned = torch.zeros(tuple(pool.states.shape), dtype=torch.uint8).pin_memory()
        pool.set_host_states(pinned)
        pool.reset_all()
        assert np.array_equal(pinned.numpy(), pool.states.cpu().numpy())
        acts, reps = util.schedule(23, 60, n, 4, k)
        terminals = 0
        for t in range(60):
            pinned.fill_(0xA5)
            _step(pool, acts[t], reps[t])
            assert np.array_equal(pinned.numpy(), pool.states.cpu().numpy()), t
            terminals += int(pool.terminals.sum())
        assert terminals > 0
        pool.set_host_states(None)
        pinned.fill_(0xA5)
        _step(pool, acts[0], reps[0])
        assert bool((pinned == 0xA5).all())
    finally:
        pool.close()
