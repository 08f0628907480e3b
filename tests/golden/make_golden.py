"""Generates the committed golden fixtures under tests/golden/ by running the REFERENCE'S OWN
modules (imported unmodified from /root/reference through oracle/ref_harness.py) on top of the
C++ oracle emulator.  Only runnable where /root/reference exists (the build container):

    python tests/golden/make_golden.py

Fixtures (small .npz files):
  figar_<game>.npz   -- N envs driven for M macro steps through the reference's
                        EmulatorRunner._run loop (emulator_runner.py:19-42) with a fixed
                        action / repetition schedule: rewards, terminals, CRC32 of every
                        returned state, the full final states, per-step next() counts
  resize_lut.npz     -- PIL NEAREST 210x160 -> 84x84 index maps (atari_emulator.py:84)
  tab_rep.npz        -- ExplorationPolicy.get_tab_repetitions for the README configurations
  action_sets.npz    -- minimal action sets / start lives as the reference sees them through ALE
"""
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_harness  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
ROMS = os.path.join(ROOT, "atari_roms")
GAMES12 = ["asterix", "asteroids", "breakout", "enduro", "gopher", "gravitar", "montezuma_revenge", "ms_pacman",
           "pong", "seaquest", "space_invaders", "yars_revenge"]


class _ScriptedQueue(object):
    def __init__(self, n):
        self.left = n

    def get(self):
        if self.left == 0:
            return None
        self.left -= 1
        return True

    def put(self, _):
        pass


def schedule(game_seed, m, n, num_actions, nb_choices):
    rng = np.random.RandomState(game_seed)
    return rng.randint(0, num_actions, size=(m, n)), rng.randint(0, nb_choices, size=(m, n))


def run_reference(ref, game, n, m, rgb, nb_choices, max_rep, seed):
    args = ref_harness.Args(game, ROMS, rgb=rgb, max_repetition=max_rep, nb_choices=nb_choices)
    emus = [ref.atari_emulator.AtariEmulator(i, args) for i in range(n)]
    num_actions = len(emus[0].get_legal_actions())
    tab_rep = ref.exploration_policy.ExplorationPolicy(args).get_tab_repetitions()
    depth = 3 if rgb else 1
    states = np.asarray([e.get_initial_state() for e in emus], dtype=np.uint8)
    variables = [states, np.zeros(n, np.float32), np.zeros(n, np.float32),
                 np.zeros((n, num_actions), np.float32), np.zeros((n, nb_choices), np.float32)]
    acts, reps = schedule(seed, m, n, num_actions, nb_choices)
    rewards = np.zeros((m, n), np.float32)
    terminals = np.zeros((m, n), np.float32)
    crcs = np.zeros((m, n), np.uint32)
    init_crc = np.array([zlib.crc32(states[e].tobytes()) for e in range(n)], np.uint32)
    for t in range(m):
        variables[3][...] = np.eye(num_actions, dtype=np.float32)[acts[t]]
        variables[4][...] = np.eye(nb_choices, dtype=np.float32)[reps[t]]
        runner = ref.emulator_runner.EmulatorRunner(tab_rep, 0, emus, variables, _ScriptedQueue(1), _ScriptedQueue(0))
        runner._run()                       # the reference's loop, in-process
        rewards[t] = variables[1]
        terminals[t] = variables[2]
        for e in range(n):
            crcs[t, e] = zlib.crc32(variables[0][e].tobytes())
    assert variables[0].shape == (n, 84, 84, 4 * depth)
    return dict(actions=acts.astype(np.int32), repetitions=reps.astype(np.int32), rewards=rewards,
                terminals=terminals, state_crc=crcs, init_state_crc=init_crc, final_states=variables[0].copy(),
                tab_rep=np.asarray(tab_rep, np.int32), num_actions=np.int32(num_actions),
                legal_actions=np.asarray(emus[0].get_legal_actions(), np.int32))


def main():
    assert ref_harness.available(), "needs /root/reference"
    ref = ref_harness.load()
    from PIL import Image
    ramp_x = np.tile(np.arange(160, dtype=np.uint8), (210, 1))
    ramp_y = np.tile(np.arange(210, dtype=np.uint8)[:, None], (1, 160))
    xmap = np.asarray(Image.fromarray(ramp_x).resize((84, 84), Image.NEAREST))[0].astype(np.int32)
    ymap = np.asarray(Image.fromarray(ramp_y).resize((84, 84), Image.NEAREST))[:, 0].astype(np.int32)
    np.savez(os.path.join(OUT, "resize_lut.npz"), xmap=xmap, ymap=ymap)

    tabs = {}
    for (mr, k) in [(0, 1), (10, 11), (10, 6), (10, 2), (20, 11)]:
        a = ref_harness.Args("pong", ROMS, max_repetition=mr, nb_choices=k)
        tabs["%d_%d" % (mr, k)] = np.asarray(ref.exploration_policy.ExplorationPolicy(a).get_tab_repetitions(), np.int32)
    np.savez(os.path.join(OUT, "tab_rep.npz"), **tabs)

    sets = {}
    for g in GAMES12:
        e = ref.atari_emulator.AtariEmulator(0, ref_harness.Args(g, ROMS))
        sets[g] = np.asarray(e.get_legal_actions(), np.int32)
        sets[g + "_lives"] = np.int32(e.lives)
    np.savez(os.path.join(OUT, "action_sets.npz"), **sets)

    cases = [("pong", 2, 150, False, 1, 0, 11), ("breakout", 2, 160, False, 11, 10, 12),
             ("seaquest", 2, 40, True, 11, 10, 13), ("ms_pacman", 2, 40, False, 6, 10, 14)]
    for game, n, m, rgb, k, mr, seed in cases:
        d = run_reference(ref, game, n, m, rgb, k, mr, seed)
        np.savez_compressed(os.path.join(OUT, "figar_%s.npz" % game), **d)
        print(game, "terminals:", int(d["terminals"].sum()), "reward sum:", float(d["rewards"].sum()))


if __name__ == "__main__":
    main()
