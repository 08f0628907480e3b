// ORACLE -- TEST INFRASTRUCTURE ONLY.  C entry points over the CPU restatement
// (a2600.hpp + ale.hpp) so tests / the `ale_python_interface` shim / the bench's
// cpu_baseline leg can drive it through ctypes.  Never linked into the product.
#include <atomic>
#include <cstdio>
#include <thread>
#include <vector>
#include "ale.hpp"

using namespace orc;

extern "C" {

void* orc_create(const uint8_t* rom, int n, const char* game_name, uint32_t seed) {
  AleEnv* e = new AleEnv();
  e->load(rom, size_t(n), game_from_name(game_name), seed);
  return e;
}
void orc_destroy(void* h) { delete static_cast<AleEnv*>(h); }
int orc_act(void* h, int action) { return static_cast<AleEnv*>(h)->act(action); }
void orc_reset_game(void* h) { static_cast<AleEnv*>(h)->reset_game(); }
int orc_game_over(void* h) { return static_cast<AleEnv*>(h)->terminal ? 1 : 0; }
int orc_lives(void* h) { return static_cast<AleEnv*>(h)->lives; }
int orc_frame_number(void* h) { return static_cast<AleEnv*>(h)->frame_number; }
int orc_num_actions(void* h) { return game_info(static_cast<AleEnv*>(h)->game).n_actions; }
void orc_minimal_actions(void* h, int* out) {
  const GameInfo& gi = game_info(static_cast<AleEnv*>(h)->game);
  for (int i = 0; i < gi.n_actions; ++i) out[i] = gi.actions[i];
}
void orc_get_ram(void* h, uint8_t* out) { std::memcpy(out, static_cast<AleEnv*>(h)->con.ram, 128); }
void orc_set_ram(void* h, int idx, int v) { static_cast<AleEnv*>(h)->con.ram[idx & 127] = uint8_t(v); }
void orc_get_screen(void* h, uint8_t* out) { std::memcpy(out, static_cast<AleEnv*>(h)->con.screen(), SCREEN_W * SCREEN_H); }
// both TIA frame buffers: [0] the current one, [1] the one drawn the frame before
void orc_get_both_screens(void* h, uint8_t* out) {
  Console& c = static_cast<AleEnv*>(h)->con;
  std::memcpy(out, c.fb[c.cur_fb], SCREEN_W * SCREEN_H);
  std::memcpy(out + SCREEN_W * SCREEN_H, c.fb[c.cur_fb ^ 1], SCREEN_W * SCREEN_H);
}
void orc_get_screen_gray(void* h, uint8_t* out) {
  const uint8_t* s = static_cast<AleEnv*>(h)->con.screen();
  for (int i = 0; i < SCREEN_W * SCREEN_H; ++i) out[i] = palette_gray(s[i]);
}
void orc_get_screen_rgb(void* h, uint8_t* out) {
  const uint8_t* s = static_cast<AleEnv*>(h)->con.screen();
  const uint32_t* p = ntsc_palette();
  for (int i = 0; i < SCREEN_W * SCREEN_H; ++i) {
    uint32_t c = p[s[i] >> 1];
    out[3 * i] = uint8_t(c >> 16); out[3 * i + 1] = uint8_t(c >> 8); out[3 * i + 2] = uint8_t(c);
  }
}
// A, X, Y, SP, PC, PS, cycles, scanline-of-last-frame-end, cart bank, timer
void orc_get_cpu(void* h, int32_t* out) {
  Console& c = static_cast<AleEnv*>(h)->con;
  out[0] = c.A; out[1] = c.X; out[2] = c.Y; out[3] = c.SP; out[4] = c.PC; out[5] = c.get_ps(); out[6] = c.cycles;
  out[7] = (c.cycles * 3 - c.clk_frame_start) / 228; out[8] = c.bank; out[9] = c.timer;
}
void orc_palette(uint8_t* gray128, uint8_t* rgb128x3) {
  const uint32_t* p = ntsc_palette();
  for (int i = 0; i < 128; ++i) {
    gray128[i] = palette_gray(uint8_t(i << 1));
    rgb128x3[3 * i] = uint8_t(p[i] >> 16); rgb128x3[3 * i + 1] = uint8_t(p[i] >> 8); rgb128x3[3 * i + 2] = uint8_t(p[i]);
  }
}
void orc_hmove_table(int8_t* out76x16) { std::memcpy(out76x16, tables().motion, 76 * 16); }

// ---- raw-console taps used by the 6502 / TIA conformance tests
void* orc_console_create(const uint8_t* rom, int n) {
  AleEnv* e = new AleEnv();
  e->rom.assign(rom, rom + n);
  e->con.rom = e->rom.data(); e->con.rom_size = uint32_t(n); e->con.cart = AleEnv::detect_cart(rom, size_t(n));
  e->con.system_reset(0);
  e->con.start_frame();            // as the first run_frame() after a reset would
  e->con.partial_frame = true;
  return e;
}
void orc_console_step(void* h, int n_instr) { Console& c = static_cast<AleEnv*>(h)->con; for (int i = 0; i < n_instr; ++i) c.step(); }
void orc_console_frame(void* h) { static_cast<AleEnv*>(h)->con.run_frame(); }

// ---- best-case CPU pool: `n` environments stepped by `threads` host threads; every env runs
// `next()` = 4 act() calls `reps[i]+1` times with early exit + reset on terminal (the FiGAR loop of
// emulator_runner.py:19-42 without the Python).  Returns the number of next() calls executed.
long orc_pool_step(void** envs, int n, const int* actions, const int* reps, int threads, float* rewards, uint8_t* terminals) {
  std::atomic<int> cursor(0);
  std::atomic<long> total(0);
  auto work = [&]() {
    long mine = 0;
    for (;;) {
      int i = cursor.fetch_add(1);
      if (i >= n) break;
      AleEnv* e = static_cast<AleEnv*>(envs[i]);
      int a = game_info(e->game).actions[actions[i]];
      float r = 0; bool over = false;
      for (int k = 0; k <= reps[i] && !over; ++k) {
        for (int f = 0; f < 4; ++f) r += float(e->act(a));
        ++mine;
        over = e->terminal;
        if (over) { e->reset_game(); for (int f = 0; f < 16; ++f) e->act(0); }
      }
      rewards[i] = r; terminals[i] = over;
    }
    total += mine;
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < threads; ++t) pool.emplace_back(work);
  work();
  for (auto& t : pool) t.join();
  return total.load();
}

}  // extern "C"
