"""GPU: PAACLearner -- network, K4 sampling, pool macro steps, K6 bookkeeping, K5 returns, loss, clip, RMSProp --
drives a real pool without leaving the device."""
import numpy as np
import pytest
import torch

import util
from util import rom_bytes

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("arch,game,n,history", [("NIPS", "pong", 32, 0), ("LSTM", "breakout", 8, 5)])
def test_rollouts_update_the_network(arch, game, n, history):
    import manette_b200 as mb
    from manette_b200.learner import PAACLearner
    torch.manual_seed(0)
    tab = mb.tab_repetitions(10, 11)
    pool = mb.DevicePool([(game, rom_bytes(game), n)], tab_rep=tab, history=history)
    pool.reset_all()
    learner = PAACLearner(pool, arch=arch, seed=5)
    before = [p.detach().clone() for p in learner.network.parameters()]
    calls0 = pool.total_next_calls()
    for i in range(3):
        out = learner.train_rollout()
        assert torch.isfinite(out["loss"]) and torch.isfinite(out["global_norm"])
    assert learner.global_step == 3 * 5 * n and 0.0 < learner.get_lr() < 0.0224
    assert any(not torch.equal(a, b.detach()) for a, b in zip(before, learner.network.parameters()))
    # every env ran 1 + tab_rep[k] next() per macro step unless its episode ended: at least one each
    assert pool.total_next_calls() - calls0 >= 3 * 5 * n
    st = learner.rollout.stats.cpu().numpy()
    assert st[5] == 3 * 5 * n
    # the rollout rows are what the pool published
    assert np.array_equal(learner.rollout.masks[-1].cpu().numpy(), 1.0 - pool.terminals.cpu().numpy())
    if arch == "LSTM":
        assert learner.states.shape == (5, n, 5, 84, 84, 4)
    learner.close(); pool.close()


def test_micro_batched_update_equals_the_whole_batch():
    """The loss is a mean over T x N rows: slices of the batch, weighted by their size, accumulate the same gradient."""
    import manette_b200 as mb
    from manette_b200.learner import PAACLearner
    tab = mb.tab_repetitions(10, 11)
    got = []
    for micro in (1 << 20, 48):
        torch.manual_seed(1)
        pool = mb.DevicePool([("breakout", rom_bytes("breakout"), 40)], tab_rep=tab)
        pool.reset_all()
        learner = PAACLearner(pool, arch="NIPS", seed=9, micro_batch=micro)
        out = learner.train_rollout()
        got.append(([p.detach().clone() for p in learner.network.parameters()], float(out["loss"]), float(out["global_norm"])))
        learner.close(); pool.close()
    (wa, la, ga), (wb, lb, gb) = got
    assert abs(la - lb) < 1e-4 * max(1.0, abs(la)) and abs(ga - gb) < 1e-3 * max(1.0, abs(ga))
    for a, b in zip(wa, wb):
        assert torch.allclose(a, b, rtol=1e-3, atol=1e-5)
