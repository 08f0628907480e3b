// TEST-ONLY: compiles the device emulator core (manette_b200/csrc/emu_core.cuh) with the HOST
// compiler so its logic can be diffed against the oracle on machines without a GPU.
// Never loaded by the product package.
#include <vector>
#include <cstring>
#include "../../manette_b200/csrc/atari_env.cuh"
#include "../../manette_b200/csrc/decode_tables.h"
#include "../../manette_b200/csrc/game_db.h"

using namespace mn;

struct HostEnv {
  EnvState s; Ctx c; Tables tab; std::vector<uint8_t> rom; std::vector<uint8_t> fb; uint8_t ram[128];
  uint32_t fifo[MN_FIFO_WORDS];
  int drain_at;        // drain when this many writes are pending (tests sweep it)
  int redo_count;      // units that had to be re-run with every frame drawn
  Unit last;
  bool last_tainted, last_alldef64;   // RAM-dependence probe of the last unit
  uint64_t dep_lo, dep_hi;
};

// the flat loop of k_round for a single environment, with the exact fallback
static void run_unit(HostEnv* e, int kind, int action, int count, uint32_t seed, bool all_pixels) {
  EnvState snap = e->s;
  uint8_t ram_snap[128];
  memcpy(ram_snap, e->ram, 128);
  for (int attempt = 0; attempt < 2; ++attempt) {
    e->c.all_pixels = all_pixels || attempt == 1;
    Unit u;
    unit_init(e->c, u, kind, action, count, seed);
    Hot hot;
    hot_init(e->c, u, hot);
    const Mem mm = mem_of(e->c);
    // resets run the RAM-dependence probe (k_round<true>, general path only); everything else the kernels' flow
    while (hot_has_work(hot)) {
      if (kind == U_ACTS) unit_tick<false>(e->c, mm, u, hot); else unit_tick<true>(e->c, mm, u, hot);
      if (MN_FILL(hot.cpu.fifo_n) >= e->drain_at) hot_drain(e->c, hot);
    }
    const bool bad = unit_finish(e->c, hot);
    e->last = u;
    e->last_tainted = hot.tainted; e->last_alldef64 = !hot.obs_bad; e->dep_lo = hot.dep_lo; e->dep_hi = hot.dep_hi;
    if (!bad) break;
    e->redo_count++;
    e->s = snap;
    memcpy(e->ram, ram_snap, 128);
  }
}

// get_initial_state() the way the kernels probe it: the reset unit and the four NOOP next() units with the RAM-dependence
// probe carried from one to the next (k_round<true> RESET, then 4 x INITIAL).  out4 = def_lo def_hi dep_lo dep_hi
static void tracked_initial_state(HostEnv* e, uint32_t rnd, uint64_t* out4) {
  uint64_t def_lo = 0, def_hi = 0, dep_lo = 0, dep_hi = 0; bool tainted = false;
  for (int k = 0; k < 5; ++k) {
    e->c.all_pixels = true;
    Unit u;
    unit_init(e->c, u, k == 0 ? U_RESET : U_ACTS, 0, k == 0 ? 0 : 4, rnd);
    Hot hot;
    hot_init(e->c, u, hot);
    hot.def_lo = def_lo; hot.def_hi = def_hi; hot.dep_lo = dep_lo; hot.dep_hi = dep_hi; hot.tainted = tainted;
    const Mem mm = mem_of(e->c);
    while (hot_has_work(hot)) {
      unit_tick<true>(e->c, mm, u, hot);
      if (MN_FILL(hot.cpu.fifo_n) >= e->drain_at) hot_drain(e->c, hot);
    }
    unit_finish(e->c, hot);
    def_lo = hot.def_lo; def_hi = hot.def_hi; dep_lo = hot.dep_lo; dep_hi = hot.dep_hi; tainted = hot.tainted;
  }
  out4[0] = def_lo; out4[1] = def_hi; out4[2] = dep_lo; out4[3] = dep_hi;
}

extern "C" {
void he_tracked_initial_state(void* h, uint64_t* out4) {
  HostEnv* e = (HostEnv*)h;
  tracked_initial_state(e, rng_next(e->s.rng), out4);
}
void* he_create(const uint8_t* rom, int n, const char* game, uint32_t seed, int skip_frames) {
  HostEnv* e = new HostEnv();
  memset(&e->s, 0, sizeof(e->s));
  build_tables(&e->tab);
  e->rom.assign(rom, rom + n);
  if (n == 2048) e->rom.insert(e->rom.end(), rom, rom + n);   // 2K images twice, as k_round stages them (flat 4 KB window)
  e->rom.resize(e->rom.size() + 16, 0);   // the fast tick fetches operand bytes speculatively (up to 2 bytes past the image)
  e->fb.assign(2 * MN_FRAME_BYTES, 0);
  int g = game_id_from_name(game);
  e->s.game = (uint8_t)g; e->s.cart = (uint8_t)detect_cart(rom, n); e->s.ctrl = (uint8_t)game_db(g).ctrl;
  e->c.s = &e->s; e->c.rom = e->rom.data(); e->c.ram = e->ram; e->c.fb = e->fb.data(); e->c.tab = &e->tab;
  e->c.fifo = e->fifo; e->c.fifo_n = 0; e->c.hseq = 0; e->c.mbox_timeout = false;
  e->drain_at = MN_FIFO_HIGH; e->redo_count = 0;
  run_unit(e, U_POWER_ON, 0, 0, seed, !skip_frames);
  return e;
}
void he_destroy(void* h) { delete (HostEnv*)h; }
void he_set_drain_at(void* h, int n) { ((HostEnv*)h)->drain_at = n < 1 ? 1 : (n > MN_FIFO_CAP ? MN_FIFO_CAP : n); }
int he_redo_count(void* h) { return ((HostEnv*)h)->redo_count; }
// ALE act(): one frame, every frame drawn
int he_act(void* h, int a) { HostEnv* e = (HostEnv*)h; run_unit(e, U_ACTS, a, 1, 0, true); return e->last.reward; }
// AtariEmulator.__action_repeat: 4 acts; with skip_frames only the two pooled frames keep their pixels
int he_next(void* h, int a, int skip_frames, int* pool_single) {
  HostEnv* e = (HostEnv*)h;
  run_unit(e, U_ACTS, a, 4, 0, !skip_frames);
  if (pool_single) *pool_single = e->last.frozen_last ? 1 : 0;
  return e->last.reward;
}
void he_reset_game(void* h, int noops, int skip_frames) {
  HostEnv* e = (HostEnv*)h;
  const uint32_t rnd = rng_next(e->s.rng);   // the draw that seeds the RIOT timer
  run_unit(e, U_RESET, 0, noops, rnd, !skip_frames);
}
void he_reset_dependence(void* h, uint64_t* out2) { HostEnv* e = (HostEnv*)h; out2[0] = e->dep_lo; out2[1] = e->dep_hi; }
// reset with a given RNG draw for the RIOT timer seed, then the four NOOP next() of get_initial_state()
void he_initial_state_with_draw(void* h, uint32_t rnd, int skip_frames) {
  HostEnv* e = (HostEnv*)h;
  run_unit(e, U_RESET, 0, 0, rnd, !skip_frames);
  for (int i = 0; i < 4; ++i) run_unit(e, U_ACTS, 0, 4, 0, !skip_frames);
}
// ---- raw-console taps for the 6502 conformance fuzz (no ALE layer, generic game)
void* he_console_create(const uint8_t* rom, int n) {
  HostEnv* e = new HostEnv();
  memset(&e->s, 0, sizeof(e->s));
  memset(e->ram, 0, 128);
  build_tables(&e->tab);
  e->rom.assign(rom, rom + n);
  if (n == 2048) e->rom.insert(e->rom.end(), rom, rom + n);   // 2K images twice, as k_round stages them (flat 4 KB window)
  e->rom.resize(e->rom.size() + 16, 0);   // the fast tick fetches operand bytes speculatively (up to 2 bytes past the image)
  e->fb.assign(2 * MN_FRAME_BYTES, 0);
  e->s.game = 0; e->s.cart = (uint8_t)detect_cart(rom, n); e->s.ctrl = 0;
  e->c.s = &e->s; e->c.rom = e->rom.data(); e->c.ram = e->ram; e->c.fb = e->fb.data(); e->c.tab = &e->tab;
  e->c.fifo = e->fifo; e->c.fifo_n = 0; e->c.hseq = 0; e->c.mbox_timeout = false; e->c.all_pixels = true;
  e->drain_at = MN_FIFO_HIGH; e->redo_count = 0;
  e->s.swcha = 0xFF; e->s.swchb = 0x3F; e->s.flags = F_INPT4 | F_INPT5;
  for (int i = 0; i < 4; ++i) e->s.analog[i] = MN_RES_MAX;
  console_reset(e->c, 0);
  frame_begin(e->c, true);         // as the first frame job after a reset would
  e->s.flags |= F_PARTIAL;
  return e;
}
void he_set_ram(void* h, int idx, int v) { ((HostEnv*)h)->ram[idx & 127] = (uint8_t)v; }
void he_console_step(void* h, int n_instr) {
  HostEnv* e = (HostEnv*)h;
  Cpu r;
  cpu_load(e->s, r);
  r.fifo_n = e->c.fifo_n;
  const Mem mm = mem_of(e->c);
  for (int i = 0; i < n_instr; ++i) {
    if (!cpu_fast_host(e->c, mm, r)) cpu_step<false>(e->c, mm, r);
    if (MN_FILL(r.fifo_n) >= e->drain_at) { e->c.fifo_n = r.fifo_n; tia_drain(e->c); r.fifo_n = e->c.fifo_n; }
  }
  e->c.fifo_n = r.fifo_n;
  tia_drain(e->c);
  cpu_store(e->s, r);
}
// 0 = general path only, 1 = fast tick first (the kernels' flow), 2 = both on every instruction, abort on a difference
void he_set_fast_mode(int mode) { g_fast_mode = mode; }
void he_set_fast_flat(int on) { g_fast_flat = on; }
void he_fast_stats(uint64_t* out2) { out2[0] = g_fast_taken; out2[1] = g_fast_refused; }
int he_state_size() { return (int)sizeof(EnvState); }
void he_get_state(void* h, uint8_t* out) { memcpy(out, &((HostEnv*)h)->s, sizeof(EnvState)); }
// 1 if the last reset never read a RAM byte before writing it and had written all 128 before the settings reset
int he_reset_was_ram_independent(void* h) { HostEnv* e = (HostEnv*)h; return (!e->last_tainted && e->last_alldef64) ? 1 : 0; }
int he_game_over(void* h) { return (((HostEnv*)h)->s.flags & F_TERMINAL) ? 1 : 0; }
int he_lives(void* h) { return ((HostEnv*)h)->s.lives; }
void he_get_ram(void* h, uint8_t* out) { memcpy(out, ((HostEnv*)h)->ram, 128); }
void he_get_screen(void* h, uint8_t* out) {
  HostEnv* e = (HostEnv*)h;
  memcpy(out, e->fb.data() + ((e->s.pflags & F_CURFB) ? MN_FRAME_BYTES : 0), MN_FRAME_BYTES);
}
// [0] the current frame buffer, [1] the other one
void he_get_both_screens(void* h, uint8_t* out) {
  HostEnv* e = (HostEnv*)h;
  const int cur = (e->s.pflags & F_CURFB) ? 1 : 0;
  memcpy(out, e->fb.data() + cur * MN_FRAME_BYTES, MN_FRAME_BYTES);
  memcpy(out + MN_FRAME_BYTES, e->fb.data() + (cur ^ 1) * MN_FRAME_BYTES, MN_FRAME_BYTES);
}
void he_get_cpu(void* h, int32_t* out) {
  EnvState& s = ((HostEnv*)h)->s;
  out[0] = s.A; out[1] = s.X; out[2] = s.Y; out[3] = s.SP; out[4] = s.PC; out[5] = (int32_t)pack_ps(s); out[6] = s.cycles;
  out[7] = (s.cycles * 3 - s.clk_frame_start) / 228; out[8] = s.bank; out[9] = s.timer;
}
}
