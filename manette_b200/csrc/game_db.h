// manette_b200 -- host-side per-game facts the device environment needs: controller wiring,
// minimal action set (what ALE's getMinimalActionSet returns to atari_emulator.py:29), and the
// bank-switching scheme of the cartridge image.
#pragma once
#include <stddef.h>
#include <string.h>
#include "emu_core.cuh"

namespace mn {

struct GameEntry {
  const char* rom_name;     // file stem under --rom_path (train.py:88)
  int ctrl;
  int n_actions;
  unsigned char actions[18];
};

inline const GameEntry& game_db(int g) {
  static const GameEntry db[G_NUM_GAMES] = {
      /* G_GENERIC        */ {"", CTRL_JOYSTICK, 18, {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17}},
      /* G_PONG           */ {"pong", CTRL_PADDLES_SWAPPED, 6, {0, 1, 3, 4, 11, 12}},
      /* G_BREAKOUT       */ {"breakout", CTRL_PADDLES, 4, {0, 1, 3, 4}},
      /* G_SEAQUEST       */ {"seaquest", CTRL_JOYSTICK, 18, {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17}},
      /* G_SPACE_INVADERS */ {"space_invaders", CTRL_JOYSTICK, 6, {0, 1, 3, 4, 11, 12}},
      /* G_MS_PACMAN      */ {"ms_pacman", CTRL_JOYSTICK, 9, {0, 2, 3, 4, 5, 6, 7, 8, 9}},
      /* G_ASTERIX        */ {"asterix", CTRL_JOYSTICK, 9, {0, 2, 3, 4, 5, 6, 7, 8, 9}},
      /* G_ASTEROIDS      */ {"asteroids", CTRL_JOYSTICK, 14, {0, 1, 2, 3, 4, 5, 6, 7, 10, 11, 12, 13, 14, 15}},
      /* G_ENDURO         */ {"enduro", CTRL_JOYSTICK, 9, {0, 1, 3, 4, 5, 8, 9, 11, 12}},
      /* G_GOPHER         */ {"gopher", CTRL_JOYSTICK, 8, {0, 1, 2, 3, 4, 10, 11, 12}},
      /* G_GRAVITAR       */ {"gravitar", CTRL_JOYSTICK, 18, {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17}},
      /* G_MONTEZUMA      */ {"montezuma_revenge", CTRL_JOYSTICK, 18, {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17}},
      /* G_YARS           */ {"yars_revenge", CTRL_JOYSTICK, 18, {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17}},
  };
  return db[(g > 0 && g < G_NUM_GAMES) ? g : 0];
}

inline int game_id_from_name(const char* name) {
  for (int g = 1; g < G_NUM_GAMES; ++g)
    if (strcmp(name, game_db(g).rom_name) == 0) return g;
  return G_GENERIC;
}

// 2K / 4K images are unbanked; 16K = F6; 8K = F8 unless the image carries Parker Brothers' E0
// hot-spot accesses ($1FE0-$1FF7 touched with absolute addressing).
inline int detect_cart(const unsigned char* img, size_t n) {
  if (n <= 2048) return CART_2K;
  if (n <= 4096) return CART_4K;
  if (n >= 16384) return CART_F6;
  static const unsigned char probes[8][3] = {{0x8D, 0xE0, 0x1F}, {0x8D, 0xE0, 0x5F}, {0x8D, 0xE9, 0xFF}, {0x0C, 0xE0, 0x1F},
                                             {0xAD, 0xE0, 0x1F}, {0xAD, 0xE9, 0xFF}, {0xAD, 0xED, 0xFF}, {0xAD, 0xF3, 0xBF}};
  for (int p = 0; p < 8; ++p)
    for (size_t i = 0; i + 3 <= n; ++i)
      if (img[i] == probes[p][0] && img[i + 1] == probes[p][1] && img[i + 2] == probes[p][2]) return CART_E0;
  return CART_F8;
}

}  // namespace mn
