#!/usr/bin/env python
"""Small fixed workload for ncu / compute-sanitizer: build a pool, decorrelate it with a few random macro
steps, then run `--steps` FiGAR10 macro steps.  Prints frames/s of the measured steps."""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import manette_b200 as mb  # noqa: E402

ROMS = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "atari_roms")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--game", default="ms_pacman")
    ap.add_argument("--envs", type=int, default=2048)
    ap.add_argument("--rgb", action="store_true")
    ap.add_argument("--decorrelate", type=int, default=4)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--envs-per-warp", type=int, default=0)
    ap.add_argument("--max-rep", type=int, default=10)
    ap.add_argument("--host-mirror", action="store_true", help="register a pinned host copy of the states (mn_set_host_states)")
    a = ap.parse_args()
    k = a.max_rep + 1
    pool = mb.DevicePool([(a.game, mb.load_rom(ROMS, a.game), a.envs)], rgb=a.rgb, tab_rep=list(range(k)),
                         envs_per_warp=a.envs_per_warp)
    if a.host_mirror:
        pool.set_host_states(torch.zeros(tuple(pool.states.shape), dtype=torch.uint8).pin_memory())
    pool.reset_all()
    g = torch.Generator(device="cuda")
    g.manual_seed(7)
    na = len(pool.legal_actions(0))

    def step():
        pool.action_idx.copy_(torch.randint(0, na, (a.envs,), device="cuda", generator=g, dtype=torch.int32))
        pool.repetition_idx.copy_(torch.randint(0, k, (a.envs,), device="cuda", generator=g, dtype=torch.int32))
        torch.cuda.synchronize()
        pool.step_async(use_indices=True)
        pool.wait()

    for _ in range(a.decorrelate):
        step()
    torch.cuda.profiler.start()   # ncu --profile-from-start off captures the measured steps only
    f0 = pool.total_next_calls()
    i0 = pool.total_instructions()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        step()
    dt = time.perf_counter() - t0
    torch.cuda.profiler.stop()
    fr = pool.total_next_calls() - f0
    ins = pool.total_instructions() - i0
    print("game=%s envs=%d epw=%d steps=%d next_calls=%d seconds=%.3f frames_per_s=%.1f instr_per_next=%.0f Minstr_per_s=%.1f redo=%d"
          % (a.game, a.envs, a.envs_per_warp, a.steps, fr, dt, fr / dt, ins / max(fr, 1), ins / dt / 1e6, pool.redo_count()))
    if os.environ.get("MN_DIAG") == "2":
        import ctypes as C
        from manette_b200 import _native
        out = (C.c_ulonglong * 4)()
        _native.load().mn_diag_counters(pool._h, out)
        print("diag: waiting for a free buffer %.1f %%, for blocking hand-offs %.1f %% of the 6502 lanes' time; %d hand-offs"
              % (100.0 * out[0] / max(out[2], 1), 100.0 * out[1] / max(out[2], 1), out[3]))
    pool.close()


if __name__ == "__main__":
    main()
