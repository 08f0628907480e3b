#!/usr/bin/env python
"""Diff the CPU oracle against the REAL Arcade Learning Environment, tap by tap.

The oracle's emulator (oracle/a2600.hpp + ale.hpp) restates ALE / Stella from documentation: nothing in this
repository can prove it equal to the real thing, because ALE is not installable in the build image (DESIGN.md section 2:
"parity unpinned" for the emulator core).  Anyone who HAS `ale_python_interface` (ALE 0.5 / 0.6, what the reference
imports, atari_emulator.py:2) or `ale_py` can run this script: it drives both with the reference's settings
(atari_emulator.py:19-31: seed, repeat_action_probability 0, frame_skip 1, colour averaging off), the same ROMs and
the same action schedule, and prints per game the first frame at which each tap differs -- RAM, raw screen (palette
indices), reward, game_over, lives -- or "identical over N frames".

    python tools/diff_real_ale.py [--games pong breakout ...] [--frames 3000] [--seed 3] [--json out.json]

Exit code 0 = every tap identical (or ALE missing: the script says so and exits 0 so it can sit in any pipeline),
1 = at least one divergence."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
GAMES12 = ["asterix", "asteroids", "breakout", "enduro", "gopher", "gravitar", "montezuma_revenge", "ms_pacman",
           "pong", "seaquest", "space_invaders", "yars_revenge"]


def real_ale():
    """(module name, factory) of a real ALE binding, or (None, None)."""
    try:
        from ale_python_interface import ALEInterface      # the module the reference uses
        return "ale_python_interface", ALEInterface
    except Exception:
        pass
    try:
        from ale_py import ALEInterface
        return "ale_py", ALEInterface
    except Exception:
        return None, None


def make_real(factory, rom_path, seed):
    ale = factory()
    for setter, key, val in (("setInt", "random_seed", seed), ("setFloat", "repeat_action_probability", 0.0),
                             ("setInt", "frame_skip", 1), ("setBool", "color_averaging", False)):
        try:
            getattr(ale, setter)(key.encode(), val)
        except Exception:
            getattr(ale, setter)(key, val)
    try:
        ale.loadROM(rom_path.encode())
    except Exception:
        ale.loadROM(rom_path)
    return ale


def screen_indices(ale):
    """Raw screen as palette indices if the binding exposes them (getScreen), else None."""
    try:
        return np.asarray(ale.getScreen()).reshape(210, 160)
    except Exception:
        return None


def diff_game(game, factory, frames, seed):
    import importlib.util                          # the oracle behind ALE's own call names, under a name of its own
    spec = importlib.util.spec_from_file_location("oracle_ale_shim", os.path.join(ROOT, "oracle", "shims", "ale_python_interface.py"))
    shim = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(shim)
    rom = os.path.join(ROOT, "atari_roms", game + ".bin")
    real = make_real(factory, rom, seed)
    ora = shim.ALEInterface()
    ora.setInt(b"random_seed", seed); ora.setFloat(b"repeat_action_probability", 0.0); ora.setInt(b"frame_skip", 1)
    ora.setBool(b"color_averaging", False); ora.loadROM(rom)
    acts_r, acts_o = list(real.getMinimalActionSet()), list(ora.getMinimalActionSet())
    first = {"action_set": None if [int(a) for a in acts_r] == [int(a) for a in acts_o] else 0}
    if first["action_set"] is not None:
        return first
    for tap in ("ram", "screen", "gray", "reward", "game_over", "lives"):
        first[tap] = None
    real.reset_game(); ora.reset_game()
    rng = np.random.RandomState(seed)
    for f in range(frames):
        a = int(acts_o[rng.randint(len(acts_o))])
        rr, ro = real.act(a), ora.act(a)
        taps = {"reward": (int(rr), int(ro)), "game_over": (bool(real.game_over()), bool(ora.game_over())),
                "lives": (int(real.lives()), int(ora.lives())),
                "ram": (np.asarray(real.getRAM()), ora.getRAM())}
        si = screen_indices(real)
        if si is not None:
            taps["screen"] = (si, ora.getScreen())
        try:
            taps["gray"] = (np.asarray(real.getScreenGrayscale()).reshape(210, 160), ora.getScreenGrayscale().reshape(210, 160))
        except Exception:
            pass
        for tap, (x, y) in taps.items():
            same = np.array_equal(x, y) if isinstance(x, np.ndarray) else x == y
            if not same and first[tap] is None:
                first[tap] = f
        if bool(real.game_over()):
            real.reset_game(); ora.reset_game()
    first["frames"] = frames
    return first


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--games", nargs="*", default=GAMES12)
    ap.add_argument("--frames", type=int, default=3000)
    ap.add_argument("--seed", type=int, default=3)
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    name, factory = real_ale()
    if factory is None:
        print("diff_real_ale: neither ale_python_interface nor ale_py is importable here -- nothing compared.  "
              "The oracle's emulator core stays UNPINNED against ALE (DESIGN.md section 2).")
        return 0
    print("real ALE binding: %s" % name)
    import orc_loader
    orc_loader.build()
    report, bad = {}, False
    for g in a.games:
        r = diff_game(g, factory, a.frames, a.seed)
        report[g] = r
        diverged = {k: v for k, v in r.items() if k != "frames" and v is not None}
        bad = bad or bool(diverged)
        print("%-18s %s" % (g, ("identical over %d frames" % a.frames) if not diverged else
                            "FIRST DIVERGENCE (frame per tap): %s" % diverged))
    if a.json:
        with open(a.json, "w") as f:
            json.dump(report, f, indent=1)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
