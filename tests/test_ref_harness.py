"""The reference's OWN hot-path modules (imported unmodified from /root/reference through
oracle/ref_harness.py) against the oracle port (oracle/host_path.py).  Only runs where the reference tree
exists (the build container); the GPU box relies on the committed fixtures these modules generated."""
import numpy as np
import pytest

import util
from util import host_path, ref_harness

pytestmark = pytest.mark.skipif(not ref_harness.available(), reason="reference tree not present")


@pytest.fixture(scope="module")
def ref():
    return ref_harness.load()


def test_tab_repetitions(ref):
    for mr, k in [(0, 1), (10, 11), (10, 6), (10, 2), (20, 11), (7, 3)]:
        a = util.args_for("pong", max_repetition=mr, nb_choices=k)
        assert ref.exploration_policy.ExplorationPolicy(a).get_tab_repetitions() == host_path.tab_repetitions(mr, k)
        from manette_b200.pool import tab_repetitions
        assert tab_repetitions(mr, k) == host_path.tab_repetitions(mr, k)


@pytest.mark.parametrize("game,rgb", [("pong", False), ("seaquest", True)])
def test_atari_emulator_next(ref, game, rgb):
    """reference AtariEmulator vs PortAtariEmulator, call by call."""
    a = util.args_for(game, rgb=rgb)
    r = ref.atari_emulator.AtariEmulator(0, a)
    p = host_path.PortAtariEmulator(0, a)
    assert list(r.get_legal_actions()) == list(p.get_legal_actions())
    assert np.array_equal(r.get_initial_state(), p.get_initial_state())
    rng = np.random.RandomState(0)
    for _ in range(25):
        act = int(rng.randint(len(p.get_legal_actions())))
        s1, r1, t1 = r.next(act)
        s2, r2, t2 = p.next(act)
        assert np.array_equal(s1, s2) and r1 == r2 and t1 == t2
    assert r.get_noop() == p.get_noop()


def test_emulator_runner_loop(ref):
    """reference EmulatorRunner._run vs figar_macro_step on the same emulators' twins."""
    game, n, k = "breakout", 3, 11
    a = util.args_for(game, max_repetition=10, nb_choices=k)
    tab = ref.exploration_policy.ExplorationPolicy(a).get_tab_repetitions()
    remus = [ref.atari_emulator.AtariEmulator(i, a) for i in range(n)]
    pemus = [host_path.PortAtariEmulator(i, a) for i in range(n)]
    states = np.asarray([e.get_initial_state() for e in remus], np.uint8)
    for e in pemus:
        e.get_initial_state()
    variables = [states, np.zeros(n, np.float32), np.zeros(n, np.float32), np.zeros((n, 4), np.float32), np.zeros((n, k), np.float32)]
    acts, reps = util.schedule(1, 12, n, 4, k)

    class Q(object):
        def __init__(self, items):
            self.items = list(items)

        def get(self):
            return self.items.pop(0)

        def put(self, _):
            pass

    for t in range(12):
        variables[3][...] = np.eye(4, dtype=np.float32)[acts[t]]
        variables[4][...] = np.eye(k, dtype=np.float32)[reps[t]]
        ref.emulator_runner.EmulatorRunner(tab, 0, remus, variables, Q([True, None]), Q([]))._run()
        for e in range(n):
            s, rew, over, _ = host_path.figar_macro_step(pemus[e], variables[3][e], variables[4][e], tab)
            assert np.array_equal(s, variables[0][e]) and rew == variables[1][e] and float(over) == variables[2][e]


def test_reference_runners_shared_array_quirk(ref):
    """runners.py:9 maps np.uint8 to c_uint: the reference's shared states are uint32 (documented quirk)."""
    v = [np.zeros((2, 84, 84, 4), np.uint8), np.zeros(2, np.float32)]
    r = ref.runners.Runners.__new__(ref.runners.Runners)
    assert r._get_shared(v[0]).dtype == np.uint32 and r._get_shared(v[1]).dtype == np.float32
