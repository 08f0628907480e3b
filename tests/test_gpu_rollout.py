"""GPU: mn_rollout_record (K6) and the observation-history ring against the oracle's restatement of the
reference's rollout bookkeeping (oracle/host_path.py PortRollout; paac.py:79-83,107-205)."""
import numpy as np
import pytest
import torch

import util
from util import OraclePool, host_path, rom_bytes

pytestmark = pytest.mark.gpu


def _onehot(idx, k):
    return np.eye(k)[np.asarray(idx)]


@pytest.mark.parametrize("n,A,K,T", [(7, 4, 11, 5), (1500, 18, 11, 5), (4097, 9, 6, 3)])
def test_record_matches_paac_loop_on_synthetic_steps(n, A, K, T):
    import manette_b200 as mb
    tab = host_path.tab_repetitions(10, K)
    rng = np.random.RandomState(n)
    port = host_path.PortRollout(np.zeros((n, 1, 1, 4), np.uint8), n, A, tab, T)
    ro = mb.Rollout(n, T, A, tab)
    total_r, total_s = [], []
    for rollout in range(3):
        port.begin(); ro.begin()
        for t in range(T):
            a, k = rng.randint(0, A, n), rng.randint(0, K, n)
            rew = (rng.randint(-3, 8, n) * rng.choice([0.0, 0.5, 1.0], n)).astype(np.float32)
            over = (rng.rand(n) < 0.2).astype(np.float32)
            oa, ok = _onehot(a, A), _onehot(k, K)
            port.before_step(t, oa, ok)
            fin = port.after_step(t, oa, ok, None, rew, over)
            ro.record(t, torch.as_tensor(rew).cuda(), torch.as_tensor(over).cuda(), torch.as_tensor(a.astype(np.int32)).cuda(),
                      torch.as_tensor(k.astype(np.int32)).cuda())
            fr, fs = ro.finished()
            assert [float(x[0]) for x in fin] == [float(x) for x in fr] and [int(x[1]) for x in fin] == [int(x) for x in fs], (rollout, t)
            total_r += [x[0] for x in fin]; total_s += [x[1] for x in fin]
        assert np.array_equal(ro.rewards.cpu().numpy(), port.rewards.astype(np.float32))
        assert np.array_equal(ro.masks.cpu().numpy(), port.episodes_over_masks.astype(np.float32))
        assert np.array_equal(ro.actions.cpu().numpy(), port.actions.argmax(-1))
        assert np.array_equal(ro.repetitions.cpu().numpy(), port.repetitions.argmax(-1))
        assert np.array_equal(ro.episode_reward.cpu().numpy(), np.asarray(port.total_episode_rewards, np.float64))
        assert np.array_equal(ro.episode_steps.cpu().numpy(), np.asarray(port.emulator_steps))
        assert np.array_equal(ro.actions_sum.cpu().numpy(), port.actions_sum.astype(np.float32))
        assert np.array_equal(ro.action_rep.cpu().numpy(), port.total_action_rep.astype(np.int64))
    st = ro.stats.cpu().numpy()
    assert st[0] == len(total_r) and st[5] == port.global_step
    # sums in env / completion order, as a Python `sum` over the reference's lists
    acc = 0.0
    for v in total_r:
        acc += float(np.float32(v))
    assert st[1] == acc and st[2] == float(sum(total_s))
    assert st[3] == min(float(np.float32(v)) for v in total_r) and st[4] == max(float(np.float32(v)) for v in total_r)
    # returns from the recorded rows = the reference's recursion on its float64 arrays (paac.py:226-231), 1e-6 relative
    ro.values.copy_(torch.as_tensor(rng.randn(T, n).astype(np.float32)))
    boot = rng.randn(n).astype(np.float32)
    y, adv = ro.returns(torch.as_tensor(boot).cuda(), 0.99)
    R = boot.astype(np.float64)
    vals = ro.values.cpu().numpy().astype(np.float64)
    for t in reversed(range(T)):
        R = port.rewards[t] + 0.99 * R * port.episodes_over_masks[t]
        assert np.allclose(y[t].cpu().numpy(), R, rtol=1e-6, atol=1e-6)
        assert np.allclose(adv[t].cpu().numpy(), R - vals[t], rtol=1e-6, atol=1e-5)
    ro.close()


def test_history_ring_equals_update_memory_through_real_episodes():
    """Breakout with single-life episodes (terminals within a few steps): the pool's H-deep history, read back in
    the reference's order, must equal PAACLearner's `memory` array step by step, wipes included."""
    import manette_b200 as mb
    game, n, k, H, T = "breakout", 12, 11, 5, 4
    ora = OraclePool(game, n, nb_choices=k, max_repetition=10, single_life=True)
    pool = mb.DevicePool([(game, rom_bytes(game), n)], tab_rep=ora.tab_rep, single_life_episodes=True, history=H)
    s0 = ora.initial_states()
    pool.reset_all()
    assert np.array_equal(pool.states.cpu().numpy(), s0)
    port = host_path.PortRollout(s0, n, ora.num_actions, ora.tab_rep, T, lstm=True, n_steps=H)
    assert np.array_equal(pool.history_ordered().cpu().numpy(), port.memory)
    acts, reps = util.schedule(23, 16, n, ora.num_actions, k)
    wipes = 0
    for step in range(16):
        ws, wr, wt, _ = ora.macro_step(acts[step], reps[step])
        pool.action_idx.copy_(torch.as_tensor(acts[step].astype(np.int32)))
        pool.repetition_idx.copy_(torch.as_tensor(reps[step].astype(np.int32)))
        pool.step_async(use_indices=True); pool.wait()
        oa, ok = _onehot(acts[step], ora.num_actions), _onehot(reps[step], k)
        port.before_step(step % T, oa, ok)
        port.after_step(step % T, oa, ok, ws, wr, wt)
        wipes += int(wt.sum())
        got = pool.history_ordered().cpu().numpy()
        assert np.array_equal(got, port.memory), step
        # the ring itself: slot `head` is the newest entry
        assert np.array_equal(pool.history[:, pool.history_head].cpu().numpy(), port.memory[:, -1])
    assert wipes > 0, "the schedule was meant to end some episodes"
    pool.close()


def test_history_gather_is_ordered_with_its_consumer():
    """A large gather (0.3 GB) consumed straight away on torch's current stream must equal a gather taken after a
    full device synchronisation: history_ordered() runs on the caller's stream, after the pool's stream."""
    import manette_b200 as mb
    game, n, k, H = "breakout", 2048, 11, 5
    tab = mb.tab_repetitions(10, k)
    pool = mb.DevicePool([(game, rom_bytes(game), n)], tab_rep=tab, history=H)
    pool.reset_all()
    g = torch.Generator().manual_seed(3)
    for step in range(3):
        pool.action_idx.copy_(torch.randint(0, pool.num_actions, (n,), generator=g, dtype=torch.int32))
        pool.repetition_idx.copy_(torch.randint(0, 3, (n,), generator=g, dtype=torch.int32))
        pool.step_async(use_indices=True)               # no wait: the gather has to order itself after the step
        early = pool.history_ordered().clone()
        torch.cuda.synchronize()
        late = pool.history_ordered()
        torch.cuda.synchronize()
        assert torch.equal(early, late), step
        assert torch.equal(late[:, -1], pool.states)
    pool.close()
