// TEST-ONLY: compiles the device emulator core (manette_b200/csrc/emu_core.cuh) with the HOST
// compiler so its logic can be diffed against the oracle on machines without a GPU.
// Never loaded by the product package.
#include <vector>
#include <cstring>
#include "../../manette_b200/csrc/emu_core.cuh"
#include "../../manette_b200/csrc/decode_tables.h"
#include "../../manette_b200/csrc/game_db.h"

using namespace mn;

struct HostEnv {
  EnvState s; Ctx c; Tables tab; std::vector<uint8_t> rom; std::vector<uint8_t> fb; uint8_t ram[128];
};

extern "C" {
void* he_create(const uint8_t* rom, int n, const char* game, uint32_t seed) {
  HostEnv* e = new HostEnv();
  memset(&e->s, 0, sizeof(e->s));
  build_tables(&e->tab);
  e->rom.assign(rom, rom + n);
  e->fb.assign(2 * MN_FRAME_BYTES, 0);
  int g = game_id_from_name(game);
  e->s.game = (uint8_t)g; e->s.cart = (uint8_t)detect_cart(rom, n); e->s.ctrl = (uint8_t)game_db(g).ctrl;
  e->c.s = &e->s; e->c.rom = e->rom.data(); e->c.ram = e->ram; e->c.ram_stride = 4; e->c.fb = e->fb.data(); e->c.tab = &e->tab;
  ale_power_on(e->c, seed);
  return e;
}
void he_destroy(void* h) { delete (HostEnv*)h; }
int he_act(void* h, int a) { return ale_act(((HostEnv*)h)->c, a); }
void he_reset_game(void* h) { ale_reset(((HostEnv*)h)->c); }
int he_game_over(void* h) { return (((HostEnv*)h)->s.flags & F_TERMINAL) ? 1 : 0; }
int he_lives(void* h) { return ((HostEnv*)h)->s.lives; }
void he_get_ram(void* h, uint8_t* out) { memcpy(out, ((HostEnv*)h)->ram, 128); }
void he_get_screen(void* h, uint8_t* out) {
  HostEnv* e = (HostEnv*)h;
  memcpy(out, e->fb.data() + ((e->s.flags & F_CURFB) ? MN_FRAME_BYTES : 0), MN_FRAME_BYTES);
}
void he_get_cpu(void* h, int32_t* out) {
  EnvState& s = ((HostEnv*)h)->s;
  out[0] = s.A; out[1] = s.X; out[2] = s.Y; out[3] = s.SP; out[4] = s.PC; out[5] = (int32_t)pack_ps(s); out[6] = s.cycles;
  out[7] = (s.cycles * 3 - s.clk_frame_start) / 228; out[8] = s.bank; out[9] = s.timer;
}
}
