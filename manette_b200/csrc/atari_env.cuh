// manette_b200 -- device-side restatement of the reference's AtariEmulator layer
// (atari_emulator.py:70-77,90-100,112-130) on top of the machine in emu_core.cuh.
//
// The reference's FramePool (environment.py:42-55) is the emulated TIA's own pair of frame
// buffers here: every next() runs exactly four frames, the TIA flips buffers every frame, so
// after a next() the two buffers ARE the two grabbed screens.  The one exception is the ALE
// freeze (act() emulates nothing once the game is over): if the fourth act() of a next() was
// frozen both grabs saw the current buffer only; `pool_single` reports that.
#pragma once
#include "emu_core.cuh"

namespace mn {

#define MN_IMG 84
#define MN_STACK 4          // atari_emulator.py:11 NR_IMAGES
#define MN_ACTION_REPEAT 4  // atari_emulator.py:12
#define MN_MAX_START_NOOPS 30   // atari_emulator.py:10,75

struct NextOut {
  int32_t reward;
  bool terminal;      // __is_terminal() (atari_emulator.py:126-130)
  bool pool_single;   // both pooled frames are the current screen
};

// what next() returns once its U_ACTS unit (4 acts) has run: atari_emulator.py:118-121
MN_HD MN_INLINE NextOut env_next_result(EnvState& s, const Unit& u, bool single_life) {
  NextOut o;
  o.reward = u.reward;
  o.pool_single = u.frozen_last;
  const bool over = (s.flags & F_TERMINAL) != 0;
  o.terminal = single_life ? (over || s.host_lives > s.lives) : over;
  s.host_lives = s.lives;
  return o;
}

// the start no-op count of episode `episode` of environment `global_env`.  The reference draws it
// from Python's unseeded `random.randint(0, 30)` (atari_emulator.py:75); the boundary makes it a
// reproducible function of (seed, env, episode) so the CPU oracle can be fed the same schedule.
MN_HD MN_INLINE uint32_t start_noops(uint32_t seed, uint32_t global_env, uint32_t episode) {
  uint32_t h = seed * 0x9E3779B1u + global_env * 0x85EBCA77u + episode * 0xC2B2AE3Du + 0x27D4EB2Fu;
  h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12; h *= 0x297A2D39u; h ^= h >> 15;
  return h % (MN_MAX_START_NOOPS + 1);
}

}  // namespace mn
