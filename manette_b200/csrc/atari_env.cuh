// manette_b200 -- device-side restatement of the reference's AtariEmulator layer
// (atari_emulator.py:70-77,90-100,112-130) on top of the machine in emu_core.cuh.
//
// The reference's FramePool (environment.py:42-55) is the emulated TIA's own pair of frame
// buffers here: every next() runs exactly four frames, the TIA flips buffers every frame, so
// after a next() the two buffers ARE the two grabbed screens.  The one exception is the ALE
// freeze (act() emulates nothing once the game is over): if the fourth act() of a next() was
// frozen both grabs saw the current buffer only; env_next() reports that through `pool_single`.
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

// atari_emulator.py:90-100 + :120-121: four act() calls with the same action, then terminal / lives
MN_HD MN_INLINE NextOut env_next(Ctx& c, int ale_action, bool single_life) {
  EnvState& s = *c.s;
  NextOut o;
  o.reward = 0;
  o.pool_single = false;
  for (int f = 0; f < MN_ACTION_REPEAT; ++f) {
    if (f == MN_ACTION_REPEAT - 1) o.pool_single = (s.flags & F_TERMINAL) != 0;
    o.reward += ale_act(c, ale_action);
  }
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

// atari_emulator.py:70-77 __new_game(): reset_game, lives, optional random-start no-ops
MN_HD MN_INLINE void env_new_game(Ctx& c, int noops) {
  EnvState& s = *c.s;
  ale_reset(c);
  s.host_lives = s.lives;
  for (int i = 0; i < noops; ++i) ale_act(c, 0);
}

}  // namespace mn
