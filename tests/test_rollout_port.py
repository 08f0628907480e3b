"""CPU: the oracle's restatement of the PAAC rollout bookkeeping (oracle/host_path.py PortRollout, paac.py:79-83,
107-205) against hand-derived expectations of the reference's loops."""
import numpy as np

import util  # noqa: F401  (sets up sys.path)
from util import host_path


def _onehot(idx, k):
    return np.eye(k)[np.asarray(idx)]


def test_accumulators_and_episode_end_follow_paac_loop():
    n, A, T = 3, 4, 2
    tab = [0, 2, 4]
    ro = host_path.PortRollout(np.zeros((n, 2, 2, 4), np.uint8), n, A, tab, T)
    acts, reps = _onehot([1, 3, 0], A), _onehot([2, 0, 1], len(tab))
    ro.before_step(0, acts, reps)
    assert ro.nb_actions == 3 + 1 + 2                                   # argmax(rep) + 1 per env (paac.py:157)
    fin = ro.after_step(0, acts, reps, None, np.asarray([2.5, -3.0, 0.5], np.float32), np.asarray([0, 1, 0], np.float32))
    assert np.array_equal(ro.rewards[0], [1.0, -1.0, 0.5])              # clipped (paac.py:180)
    assert np.array_equal(ro.episodes_over_masks[0], [1.0, 0.0, 1.0])
    assert fin == [(-3.0, 1)] and ro.total_rewards == [-3.0] and ro.total_steps == [1]
    assert ro.total_episode_rewards == [2.5, 0, 0.5] and ro.emulator_steps == [5, 0, 3]
    assert ro.global_step == n and ro.total_action_rep[1][2] == 1 and ro.total_action_rep[3][0] == 1
    assert np.array_equal(ro.actions_sum, [[0, 1, 0, 0], [0, 0, 0, 0], [1, 0, 0, 0]])   # env 1 zeroed at its episode end


def test_memory_shifts_and_is_wiped_newest_entry_included():
    n, T, H = 2, 3, 5
    s0 = np.full((n, 2, 2, 4), 7, np.uint8)
    ro = host_path.PortRollout(s0, n, 2, [0], T, lstm=True, n_steps=H)
    assert ro.memory[:, :-1].sum() == 0 and np.array_equal(ro.memory[:, -1], s0)      # paac.py:107-112
    acts, reps = _onehot([0, 1], 2), _onehot([0, 0], 1)
    s1 = np.full((n, 2, 2, 4), 9, np.uint8)
    ro.after_step(0, acts, reps, s1, np.zeros(n, np.float32), np.asarray([0, 1], np.float32))
    assert np.array_equal(ro.whole_memory[0][:, -1], s0)               # the memory the net saw at t = 0
    assert np.array_equal(ro.memory[0, -2:], np.stack([s0[0], s1[0]]))
    assert ro.memory[1].sum() == 0                                      # paac.py:200-201: all H entries, the new state too
