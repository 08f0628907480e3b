"""Shared helpers of the test-suite: the oracle-side environments and comparison loops."""
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ROMS = os.path.join(ROOT, "atari_roms")
GOLDEN = os.path.join(ROOT, "tests", "golden")
GAMES12 = ["asterix", "asteroids", "breakout", "enduro", "gopher", "gravitar", "montezuma_revenge", "ms_pacman",
           "pong", "seaquest", "space_invaders", "yars_revenge"]

for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "oracle", "shims")):
    if p not in sys.path:
        sys.path.insert(0, p)

import host_path  # noqa: E402  (oracle: test infrastructure)
import ref_harness  # noqa: E402


def args_for(game, **kw):
    return ref_harness.Args(game, ROMS, **kw)


def rom_bytes(game):
    with open(os.path.join(ROMS, game + ".bin"), "rb") as f:
        return f.read()


def crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes())


def schedule(seed, m, n, num_actions, nb_choices):
    """The action / repetition schedule tests/golden/make_golden.py used."""
    rng = np.random.RandomState(seed)
    return rng.randint(0, num_actions, size=(m, n)), rng.randint(0, nb_choices, size=(m, n))


class OraclePool(object):
    """N oracle-backed PortAtariEmulator objects stepped like emulator_runner.py:19-42."""

    def __init__(self, game, n, rgb=False, nb_choices=1, max_repetition=0, single_life=False, random_start=False,
                 seed=3, env_id_offset=0, noops=None):
        self.args = args_for(game, rgb=rgb, max_repetition=max_repetition, nb_choices=nb_choices,
                             single_life_episodes=single_life, random_start=random_start, random_seed=seed)
        self.emus = []
        for i in range(n):
            sched = None
            if random_start:
                # bind this environment's id now: a bare generator expression would read `gid` when first advanced
                sched = (lambda gid: (noops(gid, ep) for ep in range(1 << 30)))(env_id_offset + i)
            self.emus.append(host_path.PortAtariEmulator(env_id_offset + i, self.args, noop_schedule=sched))
        self.tab_rep = host_path.tab_repetitions(max_repetition, nb_choices)
        self.n = n
        self.num_actions = len(self.emus[0].get_legal_actions())

    def initial_states(self):
        return np.stack([e.get_initial_state() for e in self.emus])

    def macro_step(self, action_idx, rep_idx):
        st, rw, tm, cnt = [], [], [], []
        for e, a, r in zip(self.emus, action_idx, rep_idx):
            ah = np.eye(self.num_actions, dtype=np.float32)[a]
            rh = np.eye(len(self.tab_rep), dtype=np.float32)[r]
            s, rew, over, k = host_path.figar_macro_step(e, ah, rh, self.tab_rep)
            st.append(s); rw.append(rew); tm.append(over); cnt.append(k)
        return np.stack(st), np.asarray(rw, np.float32), np.asarray(tm, np.float32), np.asarray(cnt, np.int32)
