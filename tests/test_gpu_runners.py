"""GPU: the drop-in boundary -- Runners / AtariEmulator / ExplorationPolicy with the reference's signatures --
driven the way paac.py drives it, against the oracle's restatement of the reference's worker pool."""
import numpy as np
import pytest

import util
from util import OraclePool, host_path, rom_bytes

pytestmark = pytest.mark.gpu


def _drive(game, n, workers, rgb=False, k=11, max_rep=10, single_life=False, random_start=False, steps=10, seed=3):
    import manette_b200 as mb
    mb.release_pools()
    args = util.args_for(game, rgb=rgb, max_repetition=max_rep, nb_choices=k, single_life_episodes=single_life,
                         random_start=random_start, random_seed=seed)
    ora = OraclePool(game, n, rgb=rgb, nb_choices=k, max_repetition=max_rep, single_life=single_life,
                     random_start=random_start, seed=seed, noops=lambda gid, ep: mb.start_noops(seed, gid, ep))
    # --- exactly paac.py:97-106
    creator = mb.EnvironmentCreator(args)
    assert creator.num_actions == ora.num_actions
    emulators = np.asarray([creator.create_environment(i) for i in range(n)])
    explo = mb.ExplorationPolicy(args)
    assert explo.tab_rep == ora.tab_rep
    shared_states = np.asarray([e.get_initial_state() for e in emulators], dtype=np.uint8)
    assert np.array_equal(shared_states, ora.initial_states())
    variables = [shared_states, np.zeros(n, np.float32), np.asarray([False] * n, np.float32),
                 np.zeros((n, creator.num_actions), np.float32), np.zeros((n, k), np.float32)]
    runners = mb.Runners(explo.tab_rep, mb.EmulatorRunner, emulators, workers, variables)
    runners.start()
    s_states, s_rewards, s_over, s_actions, s_rep = runners.get_shared_variables()
    assert s_states.dtype == np.uint8 and s_states.shape == (n, 84, 84, 4 * (3 if rgb else 1))
    acts, reps = util.schedule(17, steps, n, creator.num_actions, k)
    for t in range(steps):
        for z in range(n):                                   # paac.py:159-161
            s_actions[z] = np.eye(creator.num_actions)[acts[t][z]]
            s_rep[z] = np.eye(k)[reps[t][z]]
        s_states[...] = 0xA5        # every env's state must be rewritten by the step (the pool's kernels write this array)
        runners.update_environments()
        runners.wait_updated()
        ws, wr, wt, _ = ora.macro_step(acts[t], reps[t])
        assert np.array_equal(s_rewards, wr) and np.array_equal(s_over, wt), t
        assert np.array_equal(s_states, ws), t
    # zero-copy device views alias the same results
    dv = runners.get_device_variables()
    assert np.array_equal(dv[0].cpu().numpy(), s_states) and np.array_equal(dv[1].cpu().numpy(), s_rewards)
    runners.stop()
    mb.release_pools()


def test_pong_paac_default_config():
    """BASELINE config 1 in miniature: Pong, 32 emulators / 8 workers, no repetition head."""
    _drive("pong", 32, 8, k=1, max_rep=0, steps=8)


def test_breakout_figar10():
    _drive("breakout", 16, 4, steps=14)


def test_seaquest_rgb_figar10():
    _drive("seaquest", 8, 2, rgb=True, steps=6)


def test_single_life_episodes():
    _drive("breakout", 8, 2, single_life=True, steps=30)


def test_random_start_schedule():
    _drive("breakout", 8, 2, random_start=True, steps=25)


def test_uneven_worker_split_raises_like_np_split():
    import manette_b200 as mb
    mb.release_pools()
    args = util.args_for("pong")
    emus = [mb.AtariEmulator(i, args) for i in range(6)]
    with pytest.raises(ValueError):
        mb.Runners([0], mb.EmulatorRunner, emus, 4, None)
    mb.release_pools()


def test_test_py_style_single_emulator_loop():
    """test.py:98-109: one emulator, FiGAR loop in the caller."""
    import manette_b200 as mb
    mb.release_pools()
    args = util.args_for("pong", max_repetition=10, nb_choices=11)
    env = mb.AtariEmulator(0, args)
    ora = host_path.PortAtariEmulator(0, args)
    assert list(env.get_legal_actions()) == list(ora.get_legal_actions())
    assert np.array_equal(env.get_initial_state(), ora.get_initial_state())
    rng = np.random.RandomState(4)
    tab = mb.tab_repetitions(10, 11)
    for _ in range(6):
        a = np.eye(6)[rng.randint(6)]
        r = np.eye(11)[rng.randint(11)]
        act = mb.Action(tab, 0, a, r)
        s1, r1, t1 = env.next(act.current_action)
        s2, r2, t2 = ora.next(act.current_action)
        assert np.array_equal(s1, s2) and r1 == r2 and t1 == t2
        while act.is_repeated() and not t1:
            s1, r1, t1 = env.next(act.repeat())
            s2, r2, t2 = ora.next(act.current_action)
            assert np.array_equal(s1, s2) and r1 == r2 and t1 == t2
    assert env.get_noop() == [1.0, 0.0]
    mb.release_pools()


def test_exploration_policy_numpy_in_numpy_out():
    import manette_b200 as mb
    args = util.args_for("pong", max_repetition=10, nb_choices=11)
    explo = mb.ExplorationPolicy(args, seed=42)
    pi = np.random.RandomState(0).dirichlet(np.ones(6), size=64).astype(np.float32)
    rho = np.random.RandomState(1).dirichlet(np.ones(11), size=64).astype(np.float32)
    a, r = explo.choose_next_actions(pi, rho, 6)
    wa, wr, wah, wrh = host_path.choose_next_actions(pi, rho, 0, seed=42, step=0)
    assert a.shape == (64, 6) and r.shape == (64, 11) and a.dtype == np.float64
    assert np.array_equal(a, wah) and np.array_equal(r, wrh)
    assert explo.global_step == 64
    greedy = mb.ExplorationPolicy(args, test=True)
    a, r = greedy.choose_next_actions(pi, rho, 6)
    assert np.array_equal(a.argmax(1), pi.argmax(1)) and np.array_equal(r.argmax(1), rho.argmax(1))


def test_emulator_runner_loop_equals_the_batched_macro_step():
    """EmulatorRunner._run (emulator_runner.py:19-42 restated in process, one AtariEmulator.next() at a time) against
    Runners.update_environments() (every environment at once on the GPU) on twin pools: the same states, rewards and
    terminals after every instruction, in-loop resets included."""
    import queue
    import manette_b200 as mb
    game, n, k, steps = "breakout", 6, 11, 14
    args = util.args_for(game, max_repetition=10, nb_choices=k, single_life_episodes=True)
    mb.release_pools()
    creator = mb.EnvironmentCreator(args)
    explo = mb.ExplorationPolicy(args)
    # pool A: driven through Runners
    emus_a = [creator.create_environment(i) for i in range(n)]
    states_a = np.asarray([e.get_initial_state() for e in emus_a], dtype=np.uint8)
    var_a = [states_a, np.zeros(n, np.float32), np.zeros(n, np.float32), np.zeros((n, creator.num_actions), np.float32),
             np.zeros((n, k), np.float32)]
    runners = mb.Runners(explo.tab_rep, mb.EmulatorRunner, emus_a, 2, var_a)
    runners.start()
    sv = runners.get_shared_variables()
    # pool B: its own device pool, driven by one in-process EmulatorRunner
    pool_b = mb.DevicePool([(game, rom_bytes(game), n)], tab_rep=explo.tab_rep, single_life_episodes=True)
    emus_b = mb.emulators_for_pool(pool_b)
    var_b = [np.asarray([e.get_initial_state() for e in emus_b], dtype=np.uint8), np.zeros(n, np.float32), np.zeros(n, np.float32),
             np.zeros((n, creator.num_actions), np.float32), np.zeros((n, k), np.float32)]
    assert np.array_equal(var_b[0], states_a)
    q, barrier = queue.Queue(), queue.Queue()
    worker = mb.EmulatorRunner(explo.tab_rep, 0, emus_b, var_b, q, barrier)
    acts, reps = util.schedule(41, steps, n, creator.num_actions, k)
    terminals = 0
    for t in range(steps):
        for z in range(n):
            sv[3][z] = var_b[3][z] = np.eye(creator.num_actions)[acts[t][z]]
            sv[4][z] = var_b[4][z] = np.eye(k)[reps[t][z]]
        runners.update_environments()
        runners.wait_updated()
        q.put(True); q.put(None)
        worker._run()                                   # one instruction, then the stop marker
        assert barrier.get_nowait() is True
        assert np.array_equal(var_b[1], sv[1]) and np.array_equal(var_b[2], sv[2]), t
        assert np.array_equal(var_b[0], sv[0]), t
        terminals += int(sv[2].sum())
    assert terminals > 0
    runners.stop()
    pool_b.close()
    mb.release_pools()


def test_step_host_with_pageable_arrays_equals_the_device_step():
    """mn_step_host -- the single call INTEGRATION.md section 2 shows, with ordinary (pageable) numpy arrays -- returns
    exactly what the asynchronous device step leaves in the pool's buffers: twin pools, 40 FiGAR macro steps with
    episodes ending inside them, every env's state rewritten by every call."""
    import torch
    import manette_b200 as mb
    n, k = 48, 11
    tab = list(range(k))
    mk = lambda: mb.DevicePool([("breakout", rom_bytes("breakout"), 32), ("seaquest", rom_bytes("seaquest"), 16)],
                               tab_rep=tab, single_life_episodes=True)
    a, b = mk(), mk()
    try:
        a.reset_all(); b.reset_all()
        A = a.num_actions
        acts, reps = util.schedule(29, 40, n, 4, k)
        hs = np.zeros(tuple(b.states.shape), np.uint8); hr = np.zeros(n, np.float32); ht = np.zeros(n, np.float32)
        terminals = 0
        for t in range(40):
            oa = np.eye(A, dtype=np.float32)[acts[t]]; orr = np.eye(k, dtype=np.float32)[reps[t]]
            a.actions.copy_(torch.as_tensor(oa)); a.repetitions.copy_(torch.as_tensor(orr))
            a.step_async(use_indices=False); a.wait()
            hs[...] = 0xA5
            b.step_host(oa, orr, hs, hr, ht)
            assert np.array_equal(hs, a.states.cpu().numpy()), t
            assert np.array_equal(hr, a.rewards.cpu().numpy()) and np.array_equal(ht, a.terminals.cpu().numpy()), t
            assert np.array_equal(hs, b.states.cpu().numpy()), t
            terminals += int(ht.sum())
        assert terminals > 0
    finally:
        a.close(); b.close()
