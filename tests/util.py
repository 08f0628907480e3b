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


# ---------------------------------------------------------------------------------------------------------------
# The same oracle environments spread over worker processes (fork), for the long parity runs: episodes played to
# game over, hundreds of macro steps, sampled environments of very large pools.
def _oracle_worker(conn, game, ids, kw):
    args = args_for(game, rgb=kw.get("rgb", False), max_repetition=kw.get("max_repetition", 0),
                    nb_choices=kw.get("nb_choices", 1), single_life_episodes=kw.get("single_life", False),
                    random_seed=kw.get("seed", 3))
    emus = [host_path.PortAtariEmulator(int(i), args) for i in ids]
    tab_rep = host_path.tab_repetitions(kw.get("max_repetition", 0), kw.get("nb_choices", 1))
    na = len(emus[0].get_legal_actions())
    while True:
        msg = conn.recv()
        if msg[0] == "close":
            break
        if msg[0] == "init":
            conn.send(np.stack([e.get_initial_state() for e in emus]))
            continue
        _, acts, reps, taps = msg
        st, rw, tm, cnt = [], [], [], []
        for e, a, r in zip(emus, acts, reps):
            s, rew, over, k = host_path.figar_macro_step(e, np.eye(na, dtype=np.float32)[a],
                                                         np.eye(len(tab_rep), dtype=np.float32)[r], tab_rep)
            st.append(s); rw.append(rew); tm.append(over); cnt.append(k)
        out = [np.stack(st), np.asarray(rw, np.float32), np.asarray(tm, np.float32), np.asarray(cnt, np.int32)]
        if taps:
            out += [np.stack([e.ale.getRAM() for e in emus]), np.stack([e.ale.getScreen() for e in emus]),
                    np.asarray([e.ale.lives() for e in emus], np.int32), np.asarray([e.lives for e in emus], np.int32)]
        conn.send(out)
    conn.close()


class ParallelOraclePool(object):
    """Oracle environments with the given actor ids (ALE seed = random_seed * (id + 1)) of one game, stepped like
    emulator_runner.py:19-42 by `workers` forked processes.  macro_step(..., taps=True) also returns RAM, the
    current raw screen, ale.lives() and AtariEmulator.lives of every environment."""

    def __init__(self, game, ids, workers=None, **kw):
        import multiprocessing as mp
        ids = [int(i) for i in ids]
        workers = max(1, min(len(ids), workers or (os.cpu_count() or 1)))
        self.n = len(ids)
        self.tab_rep = host_path.tab_repetitions(kw.get("max_repetition", 0), kw.get("nb_choices", 1))
        ctx = mp.get_context("fork")
        self._parts = [list(range(w, self.n, workers)) for w in range(workers)]     # positions served by worker w
        self._conns, self._procs = [], []
        for part in self._parts:
            a, b = ctx.Pipe()
            p = ctx.Process(target=_oracle_worker, args=(b, game, [ids[j] for j in part], kw), daemon=True)
            p.start()
            self._conns.append(a); self._procs.append(p)
        probe = host_path.PortAtariEmulator(0, args_for(game))
        self.num_actions = len(probe.get_legal_actions())

    def _gather(self, outs):
        res = []
        for k in range(len(outs[0])):
            full = np.zeros((self.n,) + outs[0][k].shape[1:], outs[0][k].dtype)
            for part, o in zip(self._parts, outs):
                full[part] = o[k]
            res.append(full)
        return res

    def initial_states(self):
        for c in self._conns:
            c.send(("init",))
        return self._gather([[c.recv()] for c in self._conns])[0]

    def macro_step(self, action_idx, rep_idx, taps=False):
        for part, c in zip(self._parts, self._conns):
            c.send(("step", [int(action_idx[j]) for j in part], [int(rep_idx[j]) for j in part], bool(taps)))
        return self._gather([c.recv() for c in self._conns])

    def close(self):
        for c in self._conns:
            try:
                c.send(("close",))
            except Exception:
                pass
        for p in self._procs:
            p.join(timeout=5)
