// ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the shipped product path.
//
// CPU restatement of the ALE environment layer the reference drives through
// `ale_python_interface.ALEInterface` (reference call sites atari_emulator.py:19-31,
// 57-61,72-77,94-97,121,128,133; environment_creator.py:25-29): seeded RNG, reset
// sequence (60 NOOP frames + 4 RESET frames + per-game starting actions), act()
// (two RNG draws per frame for the sticky-action test, frozen once terminal),
// joystick/paddle input mapping, per-game RAM-scraped reward/terminal/lives
// ("RomSettings"), minimal action sets and the NTSC palette.
// PARITY UNPINNED: ALE is a third-party dependency absent from /root/reference (no
// pinned version either: atari_emulator.py:21 says ">= ALE 0.5.0"); this restates its
// published behaviour from recollection.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include "a2600.hpp"

namespace orc {

// ----------------------------------------------------------------- games
enum GameId {
  G_GENERIC = 0, G_PONG, G_BREAKOUT, G_SEAQUEST, G_SPACE_INVADERS, G_MS_PACMAN, G_ASTERIX, G_ASTEROIDS,
  G_ENDURO, G_GOPHER, G_GRAVITAR, G_MONTEZUMA, G_YARS, G_NUM_GAMES };

enum AleAction { A_NOOP = 0, A_FIRE, A_UP, A_RIGHT, A_LEFT, A_DOWN, A_UPRIGHT, A_UPLEFT, A_DOWNRIGHT, A_DOWNLEFT,
                 A_UPFIRE, A_RIGHTFIRE, A_LEFTFIRE, A_DOWNFIRE, A_UPRIGHTFIRE, A_UPLEFTFIRE, A_DOWNRIGHTFIRE,
                 A_DOWNLEFTFIRE, A_RESET = 40 };

static const int PADDLE_DELTA = 23000, PADDLE_MIN = 27450, PADDLE_MAX = 790196;
static const int PADDLE_DEFAULT = ((PADDLE_MAX - PADDLE_MIN) / 2 + PADDLE_MIN);
static const int RESET_NOOP_FRAMES = 60, RESET_SWITCH_FRAMES = 4;

struct GameInfo {
  const char* name; int ctrl; int n_actions; uint8_t actions[18]; int n_start; uint8_t start_action; int start_lives;
};

inline const GameInfo& game_info(int g) {
  static const GameInfo info[G_NUM_GAMES] = {
      {"generic", CTRL_JOYSTICK, 18, {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17}, 0, 0, 0},
      {"pong", CTRL_PADDLES_SWAPPED, 6, {0, 1, 3, 4, 11, 12}, 0, 0, 0},
      {"breakout", CTRL_PADDLES, 4, {0, 1, 3, 4}, 0, 0, 5},
      {"seaquest", CTRL_JOYSTICK, 18, {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17}, 0, 0, 4},
      {"space_invaders", CTRL_JOYSTICK, 6, {0, 1, 3, 4, 11, 12}, 0, 0, 3},
      {"ms_pacman", CTRL_JOYSTICK, 9, {0, 2, 3, 4, 5, 6, 7, 8, 9}, 0, 0, 3},
      {"asterix", CTRL_JOYSTICK, 9, {0, 2, 3, 4, 5, 6, 7, 8, 9}, 1, A_FIRE, 3},
      {"asteroids", CTRL_JOYSTICK, 14, {0, 1, 2, 3, 4, 5, 6, 7, 10, 11, 12, 13, 14, 15}, 0, 0, 4},
      {"enduro", CTRL_JOYSTICK, 9, {0, 1, 3, 4, 5, 8, 9, 11, 12}, 0, 0, 0},
      {"gopher", CTRL_JOYSTICK, 8, {0, 1, 2, 3, 4, 10, 11, 12}, 1, A_FIRE, 3},
      {"gravitar", CTRL_JOYSTICK, 18, {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17}, 16, A_FIRE, 6},
      {"montezuma_revenge", CTRL_JOYSTICK, 18, {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17}, 0, 0, 6},
      {"yars_revenge", CTRL_JOYSTICK, 18, {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17}, 1, A_FIRE, 4},
  };
  return info[(g >= 0 && g < G_NUM_GAMES) ? g : 0];
}

inline int game_from_name(const char* n) {
  for (int g = 1; g < G_NUM_GAMES; ++g) if (std::strcmp(n, game_info(g).name) == 0) return g;
  return G_GENERIC;
}

// ----------------------------------------------------------------- RNG (TinyMT32)
struct TinyMT {
  uint32_t s[4];
  static const uint32_t MAT1 = 0x8f7011eeu, MAT2 = 0xfc78ff1fu, TMAT = 0x3793fdffu;
  void next_state() {
    uint32_t y = s[3];
    uint32_t x = (s[0] & 0x7fffffffu) ^ s[1] ^ s[2];
    x ^= (x << 1);
    y ^= (y >> 1) ^ x;
    s[0] = s[1]; s[1] = s[2]; s[2] = x ^ (y << 10); s[3] = y;
    uint32_t mask = uint32_t(-int32_t(y & 1));
    s[1] ^= mask & MAT1; s[2] ^= mask & MAT2;
  }
  uint32_t temper() const {
    uint32_t t0 = s[3];
    uint32_t t1 = s[0] + (s[2] >> 8);
    t0 ^= t1;
    t0 ^= uint32_t(-int32_t(t1 & 1)) & TMAT;
    return t0;
  }
  void seed(uint32_t seed) {
    s[0] = seed; s[1] = MAT1; s[2] = MAT2; s[3] = TMAT;
    for (uint32_t i = 1; i < 8; ++i)
      s[i & 3] ^= i + 1812433253u * (s[(i - 1) & 3] ^ (s[(i - 1) & 3] >> 30));
    if ((s[0] & 0x7fffffffu) == 0 && s[1] == 0 && s[2] == 0 && s[3] == 0) { s[0] = 'T'; s[1] = 'I'; s[2] = 'N'; s[3] = 'Y'; }
    for (int i = 0; i < 8; ++i) next_state();
  }
  uint32_t next() { next_state(); return temper(); }
};

// ----------------------------------------------------------------- palette
inline const uint32_t* ntsc_palette() {   // 128 colours, index = TIA colour byte >> 1
  static const uint32_t p[128] = {
      0x000000, 0x4a4a4a, 0x6f6f6f, 0x8e8e8e, 0xaaaaaa, 0xc0c0c0, 0xd6d6d6, 0xececec,
      0x484800, 0x69690f, 0x86861d, 0xa2a22a, 0xbbbb35, 0xd2d240, 0xe8e84a, 0xfcfc54,
      0x7c2c00, 0x904811, 0xa26221, 0xb47a30, 0xc3903d, 0xd2a44a, 0xdfb755, 0xecc860,
      0x901c00, 0xa33915, 0xb55328, 0xc66c3a, 0xd5824a, 0xe39759, 0xf0aa67, 0xfcbc74,
      0x940000, 0xa71a1a, 0xb83232, 0xc84848, 0xd65c5c, 0xe46f6f, 0xf08080, 0xfc9090,
      0x840064, 0x97197a, 0xa8308f, 0xb846a2, 0xc659b3, 0xd46cc3, 0xe07cd2, 0xec8ce0,
      0x500084, 0x68199a, 0x7d30ad, 0x9246c0, 0xa459d0, 0xb56ce0, 0xc57cee, 0xd48cfc,
      0x140090, 0x331aa3, 0x4e32b5, 0x6848c6, 0x7f5cd5, 0x956fe3, 0xa980f0, 0xbc90fc,
      0x000094, 0x181aa7, 0x2d32b8, 0x4248c8, 0x545cd6, 0x656fe4, 0x7580f0, 0x8490fc,
      0x001c88, 0x183b9d, 0x2d57b0, 0x4272c2, 0x548ad2, 0x65a0e1, 0x75b5ef, 0x84c8fc,
      0x003064, 0x185080, 0x2d6d98, 0x4288b0, 0x54a0c5, 0x65b7d9, 0x75cceb, 0x84e0fc,
      0x004030, 0x18624e, 0x2d8169, 0x429e82, 0x54b899, 0x65d1ae, 0x75e7c2, 0x84fcd4,
      0x004400, 0x1a661a, 0x328432, 0x48a048, 0x5cba5c, 0x6fd26f, 0x80e880, 0x90fc90,
      0x143c00, 0x355f18, 0x527e2d, 0x6e9c42, 0x87b754, 0x9ed065, 0xb4e775, 0xc8fc84,
      0x303800, 0x505916, 0x6d762b, 0x88923e, 0xa0ab4f, 0xb7c25f, 0xccd86e, 0xe0ec7c,
      0x482c00, 0x694d14, 0x866a26, 0xa28638, 0xbb9f47, 0xd2b656, 0xe8cc63, 0xfce070};
  return p;
}
// ALE-ISM (named so that it can be found and flipped): getScreenGrayscale reads the gray value ALE's ColourPalette
// stores beside every NTSC colour, `(uInt8) round(r * 0.2989 + g * 0.5870 + b * 0.1140)` in double arithmetic
// (ColourPalette.cpp convertGrayscale of ALE >= 0.5, the version atari_emulator.py:21 asks for).
// Round 1 of this repo truncated instead (the ALE 0.4 export path); both sides changed together in round 2.
// Unverifiable here -- ALE is not installable -- and inside the +-1 tolerance either way.
#define ORC_ALE_LUMA_ROUND 1
inline uint8_t palette_gray(uint8_t tia_colour) {
  uint32_t px = ntsc_palette()[tia_colour >> 1];
  const double r = (px >> 16) & 0xFF, g = (px >> 8) & 0xFF, b = px & 0xFF;
  const double lum = r * 0.2989 + g * 0.5870 + b * 0.1140;
#if ORC_ALE_LUMA_ROUND
  return uint8_t(std::round(lum));
#else
  return uint8_t(lum);
#endif
}

// ----------------------------------------------------------------- environment
struct AleEnv {
  Console con;
  std::vector<uint8_t> rom;
  int game = G_GENERIC;
  TinyMT rng;
  int32_t left_paddle = PADDLE_DEFAULT, right_paddle = PADDLE_DEFAULT;
  int32_t frame_number = 0, episode_frame_number = 0;
  // RomSettings state
  int32_t score = 0, reward = 0, lives = 0; bool terminal = false, started = false;

  static int detect_cart(const uint8_t* r, size_t n) {
    if (n <= 2048) return CART_2K;
    if (n <= 4096) return CART_4K;
    if (n == 16384) return CART_F6;
    // 8K: Parker Bros E0 carts touch $FE0-$FF7 hot spots with absolute addressing
    static const uint8_t sig[][3] = {{0x8D, 0xE0, 0x1F}, {0x8D, 0xE0, 0x5F}, {0x8D, 0xE9, 0xFF}, {0x0C, 0xE0, 0x1F},
                                     {0xAD, 0xE0, 0x1F}, {0xAD, 0xE9, 0xFF}, {0xAD, 0xED, 0xFF}, {0xAD, 0xF3, 0xBF}};
    for (auto& s : sig)
      for (size_t i = 0; i + 3 <= n; ++i)
        if (r[i] == s[0] && r[i + 1] == s[1] && r[i + 2] == s[2]) return CART_E0;
    return CART_F8;
  }

  void load(const uint8_t* r, size_t n, int game_id, uint32_t seed) {
    rom.assign(r, r + n);
    game = game_id;
    con.rom = rom.data(); con.rom_size = uint32_t(n); con.cart = detect_cart(r, n);
    rng.seed(seed);
    for (int i = 0; i < 128; ++i) con.ram[i] = uint8_t(rng.next());   // M6532 power-on garbage
    reset_game();
  }

  inline uint8_t ram(int offset) const { return con.ram[offset & 0x7F]; }
  static int bcd(int b) { return (b >> 4) * 10 + (b & 15); }
  int dec2(int lo, int hi) const { return bcd(ram(lo)) + 100 * bcd(ram(hi)); }
  int dec3(int lo, int mid, int hi) const { return bcd(ram(lo)) + 100 * bcd(ram(mid)) + 10000 * bcd(ram(hi)); }

  void settings_reset() {
    score = 0; reward = 0; terminal = false; started = false; lives = game_info(game).start_lives;
  }
  void settings_step() {
    int s;
    switch (game) {
      case G_PONG: {
        int x = ram(13), y = ram(14);
        s = y - x; reward = s - score; score = s; terminal = (x == 21 || y == 21);
        break;
      }
      case G_BREAKOUT: {
        int x = ram(77), y = ram(76);
        s = (x & 0x0F) + 10 * ((x & 0xF0) >> 4) + 100 * (y & 0x0F);
        reward = s - score; score = s;
        int b = ram(57);
        if (!started && b == 5) started = true;
        terminal = started && b == 0; lives = b;
        break;
      }
      case G_SEAQUEST:
        s = dec3(0xBA, 0xB9, 0xB8); reward = s - score; score = s;
        terminal = ram(0xA3) != 0; lives = ram(0xBB) + 1;
        break;
      case G_SPACE_INVADERS:
        s = dec2(0xE8, 0xE6); reward = s - score; if (reward < 0) reward = (10000 - score) + s; score = s;
        lives = ram(0xC9); terminal = (ram(0x98) & 0x80) || lives == 0;
        break;
      case G_MS_PACMAN: {
        s = dec3(0xF8, 0xF9, 0xFA); reward = s - score; score = s;
        int lb = ram(0xFB) & 0xF, dt = ram(0xA7);
        terminal = (lb == 0 && dt == 0x53); lives = (lb & 0x7) + 1;
        break;
      }
      case G_ASTERIX: {
        s = dec3(0xE0, 0xDF, 0xDE); reward = s - score; score = s;
        int lv = ram(0xD3) & 0xF, dc = ram(0xC7);
        terminal = (dc == 0x01 && lv == 1); lives = lv;
        break;
      }
      case G_ASTEROIDS: {
        s = dec2(0xBE, 0xBD) * 10; reward = s - score; if (reward < 0) reward += 100000; score = s;
        int b = ram(0xBC); lives = (b - (b & 15)) >> 4; terminal = (lives == 0);
        break;
      }
      case G_ENDURO: {
        s = 0;
        int level = ram(0xAD);
        if (level != 0) {
          int cars = dec2(0xAB, 0xAC);
          if (level == 1) cars = 200 - cars; else cars = 300 - cars;
          if (level >= 2) { s = 200; s += (level - 2) * 300; }
          s += cars;
        }
        reward = s - score; score = s; terminal = (ram(0xAF) == 0xFF);
        break;
      }
      case G_GOPHER: {
        s = dec3(0xB2, 0xB1, 0xB0); reward = s - score; score = s;
        int c = ram(0xB4) & 0x7; terminal = (c == 0);
        lives = (c & 1) + ((c >> 1) & 1) + ((c >> 2) & 1);
        break;
      }
      case G_GRAVITAR: {
        s = dec3(0x09, 0x08, 0x07); reward = s - score; score = s;
        int nl = ram(0x84), scr = ram(0x81);
        terminal = (nl == 0 && scr == 0x01); lives = nl + 1;
        break;
      }
      case G_MONTEZUMA: {
        s = dec3(0x95, 0x94, 0x93); reward = s - score; score = s;
        int nl = ram(0xBA), sb = ram(0xFE);
        terminal = (nl == 0 && sb == 0x60); lives = (nl & 0x7) + 1;
        break;
      }
      case G_YARS: {
        s = dec3(0xE2, 0xE1, 0xE0); reward = s - score; score = s;
        int lb = ram(0x9E) >> 4; terminal = (lb == 0); lives = lb;
        break;
      }
      default: reward = 0; terminal = false; break;
    }
  }

  // latch controller / switch state for one frame
  void apply_action(int a) {
    bool up = false, down = false, left = false, right = false, fire = false, reset = (a == A_RESET);
    switch (a) {
      case A_FIRE: fire = true; break;
      case A_UP: up = true; break;
      case A_RIGHT: right = true; break;
      case A_LEFT: left = true; break;
      case A_DOWN: down = true; break;
      case A_UPRIGHT: up = right = true; break;
      case A_UPLEFT: up = left = true; break;
      case A_DOWNRIGHT: down = right = true; break;
      case A_DOWNLEFT: down = left = true; break;
      case A_UPFIRE: up = fire = true; break;
      case A_RIGHTFIRE: right = fire = true; break;
      case A_LEFTFIRE: left = fire = true; break;
      case A_DOWNFIRE: down = fire = true; break;
      case A_UPRIGHTFIRE: up = right = fire = true; break;
      case A_UPLEFTFIRE: up = left = fire = true; break;
      case A_DOWNRIGHTFIRE: down = right = fire = true; break;
      case A_DOWNLEFTFIRE: down = left = fire = true; break;
      default: break;
    }
    con.swchb = reset ? 0x3E : 0x3F;
    int ctrl = game_info(game).ctrl;
    if (ctrl == CTRL_JOYSTICK) {
      uint8_t v = 0xFF;
      if (up) v &= ~0x10; if (down) v &= ~0x20; if (left) v &= ~0x40; if (right) v &= ~0x80;
      con.swcha = v; con.inpt4_high = !fire; con.inpt5_high = true;
      for (int i = 0; i < 4; ++i) con.analog[i] = RES_MAX;
    } else {
      int delta = right ? -PADDLE_DELTA : left ? PADDLE_DELTA : 0;
      left_paddle += delta;
      if (left_paddle < PADDLE_MIN) left_paddle = PADDLE_MIN;
      if (left_paddle > PADDLE_MAX) left_paddle = PADDLE_MAX;
      bool swap = (ctrl == CTRL_PADDLES_SWAPPED);
      // left jack: INPT0 <- pin Nine, INPT1 <- pin Five; right jack paddles never driven (resistance 0)
      con.analog[0] = swap ? right_paddle : left_paddle;
      con.analog[1] = swap ? left_paddle : right_paddle;
      con.analog[2] = RES_MIN; con.analog[3] = RES_MIN;
      uint8_t v = 0xFF;
      if (fire) v &= swap ? ~0x40 : ~0x80;   // pin Four (bit 7) = paddle 0 button, pin Three (bit 6) = paddle 1
      con.swcha = v; con.inpt4_high = true; con.inpt5_high = true;
    }
  }

  void emulate(int action, int frames) {
    int ctrl = game_info(game).ctrl;
    if (ctrl == CTRL_JOYSTICK) apply_action(action);
    for (int t = 0; t < frames; ++t) {
      if (ctrl != CTRL_JOYSTICK) apply_action(action);   // paddles move a notch every frame
      con.run_frame();
      settings_step();
    }
  }

  void reset_game() {
    episode_frame_number = 0;
    left_paddle = right_paddle = PADDLE_DEFAULT;
    con.system_reset(rng.next());
    emulate(A_NOOP, RESET_NOOP_FRAMES);
    emulate(A_RESET, RESET_SWITCH_FRAMES);
    settings_reset();
    const GameInfo& gi = game_info(game);
    for (int i = 0; i < gi.n_start; ++i) emulate(gi.start_action, 1);
  }

  int act(int action) {
    rng.next(); rng.next();          // sticky-action draws for player A and B (probability 0 -> no effect)
    if (terminal) return 0;
    emulate(action, 1);
    ++frame_number; ++episode_frame_number;
    return reward;
  }
};

}  // namespace orc
