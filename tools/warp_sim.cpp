// TOOL (not product, not a test): replays k_round's flat warp loop on the HOST build of the emulator core for
// L lanes of one game and reports, per tick, which lanes would leave a branch-free "fast tick" and why.  Used to
// decide what the fast tick has to cover.  Build + run: tools/warp_sim.sh <game> [lanes]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <set>
#include <string>
#include <vector>
#include "../manette_b200/csrc/atari_env.cuh"
#include "../manette_b200/csrc/decode_tables.h"
#include "../manette_b200/csrc/game_db.h"

using namespace mn;

struct Lane {
  EnvState s; Ctx c; std::vector<uint8_t> fb; uint8_t ram[128]; uint32_t fifo[MN_FIFO_WORDS];
  Unit u; Hot hot;
};
static Tables g_tab;
static std::vector<uint8_t> g_rom;

static void run_alone(Lane& l, int kind, int action, int count, uint32_t seed) {
  l.c.all_pixels = true;
  unit_init(l.c, l.u, kind, action, count, seed);
  hot_init(l.c, l.u, l.hot);
  const Mem mm = mem_of(l.c);
  while (hot_has_work(l.hot)) {
    unit_tick<false>(l.c, mm, l.u, l.hot);
    if (MN_FILL(l.hot.cpu.fifo_n) >= MN_FIFO_HIGH) hot_drain(l.c, l.hot);
  }
  unit_finish(l.c, l.hot);
}

enum : uint32_t {
  R_CODE = 1u << 0, R_IND_RAM = 1u << 1, R_IND_OTHER = 1u << 2, R_RD_TIA = 1u << 3, R_RD_TIMER = 1u << 4, R_RD_RIOT = 1u << 5,
  R_RD_HOT = 1u << 6, R_WR_WSYNC = 1u << 7, R_WR_TIA_LOW = 1u << 8, R_WR_RIOT = 1u << 9, R_WR_CART = 1u << 10,
  R_FLAG = 1u << 11, R_JMP = 1u << 12, R_JSR = 1u << 13, R_RTS = 1u << 14, R_PHA = 1u << 15, R_PLA = 1u << 16, R_PHP_PLP = 1u << 17,
  R_BIT = 1u << 18, R_DECIMAL = 1u << 19, R_UNDOC = 1u << 20, R_RTI_BRK = 1u << 21, R_FIFO_FULL = 1u << 22, R_STACK_NOT_RAM = 1u << 23,
  R_JOB = 1u << 24 /* tick spent on job begin */, R_JMP_IND = 1u << 25 };
static const char* kNames[] = {"code!rom", "ind(ram ptr)", "ind(other)", "rd TIA", "rd INTIM", "rd RIOT other", "rd hotspot", "wr WSYNC",
                               "wr TIA<4", "wr RIOT", "wr cart", "flag op", "JMP abs", "JSR", "RTS", "PHA", "PLA", "PHP/PLP", "BIT",
                               "decimal", "undoc", "RTI/BRK", "fifo full", "stack!ram", "job begin", "JMP ind"};

static uint32_t peek8(const Lane& l, const Cpu& r, uint32_t addr, bool* ok) {   // side-effect free read of ROM / RAM
  *ok = true;
  if (addr & 0x1000u) {
    if ((addr & 0xFFFu) >= r.hot_lo) { *ok = false; return 0; }
    const uint32_t page = (r.segmap >> (8 * ((addr >> 10) & 3u))) & 0xFFu;
    return g_rom[(page << 10) | (addr & 0x3FFu)];
  }
  if ((addr & 0x0280u) == 0x0080u) return l.ram[addr & 0x7Fu];
  *ok = false;
  return 0;
}

static uint32_t classify(const Lane& l) {
  const Cpu& r = l.hot.cpu;
  const uint32_t pc = r.PC;
  uint32_t why = 0;
  const bool fast_code = (pc & 0x1000u) && ((pc & 0xFFFu) < 0xFDEu) && ((pc & 0x3FFu) < 0x3FEu);
  if (!fast_code) return R_CODE;
  bool ok;
  const uint32_t ir = peek8(l, r, pc, &ok), b1 = peek8(l, r, pc + 1, &ok), b2 = peek8(l, r, pc + 2, &ok);
  const TabEnt t = g_tab.e[ir];
  const uint32_t k = t.k, d = t.d;
  const uint32_t xm = t.x & 0xFFFFu;
  const uint32_t isel = (k >> K_ISEL) & 7u;
  const uint32_t idx = isel == SEL_X ? cpuX(r) : isel == SEL_Y ? cpuY(r) : 0u;
  uint32_t base = (b1 | (b2 << 8)) & xm;
  uint32_t ea = (base + idx) & xm;
  const uint32_t mode = d & 15u, op = (d >> 6) & 63u;
  if (d & D_INDIRECT) {
    if (mode == AM_IND) why |= (op == O_JMP) ? R_JMP_IND : R_IND_OTHER;
    else {
      const uint32_t p0 = (mode == AM_IZX) ? ((b1 + cpuX(r)) & 0xFFu) : b1, p1 = (p0 + 1) & 0xFFu;
      bool ok0, ok1;
      const uint32_t lo = peek8(l, r, p0, &ok0), hi = peek8(l, r, p1, &ok1);
      if (ok0 && ok1) { why |= R_IND_RAM; base = lo | (hi << 8); ea = (mode == AM_IZY) ? ((base + cpuY(r)) & 0xFFFFu) : base; }
      else return why | R_IND_OTHER;
    }
  }
  if (d & D_READ) {
    const bool rom = (ea & 0x1000u) != 0;
    const bool fast = rom ? ((ea & 0xFFFu) < r.hot_lo) : ((ea & 0x0280u) == 0x0080u);
    if (!fast) {
      if (rom) why |= R_RD_HOT;
      else if (ea & 0x80u) why |= ((ea & 0x285u) == 0x284u) ? R_RD_TIMER : R_RD_RIOT;
      else why |= R_RD_TIA;
    }
  }
  const bool generic = (k & K_GENERIC) && !((k & K_DECIMAL) && (r.P & 0x08u));
  if (!generic && !(d & D_BRANCH)) {
    switch (op) {
      case O_FLAG: why |= R_FLAG; break;
      case O_JMP: if (mode != AM_IND) why |= R_JMP; break;
      case O_JSR: why |= R_JSR; if (cpuSP(r) < 0x81) why |= R_STACK_NOT_RAM; break;
      case O_RTS: why |= R_RTS; if (cpuSP(r) < 0x80 || cpuSP(r) > 0xFD) why |= R_STACK_NOT_RAM; break;
      case O_PHA: why |= R_PHA; if (cpuSP(r) < 0x80) why |= R_STACK_NOT_RAM; break;
      case O_PLA: why |= R_PLA; if (cpuSP(r) < 0x7F || cpuSP(r) == 0xFF) why |= R_STACK_NOT_RAM; break;
      case O_PHP: case O_PLP: why |= R_PHP_PLP; break;
      case O_BIT: why |= R_BIT; break;
      case O_ADC: case O_SBC: why |= R_DECIMAL; break;
      case O_RTI: case O_BRK: why |= R_RTI_BRK; break;
      case O_NOP: case O_KIL: break;
      default: why |= R_UNDOC; break;
    }
  }
  if (d & D_WRITE) {
    if ((ea & 0x1280u) != 0x0080u) {
      const uint32_t a6 = ea & 0x3Fu;
      if (ea & 0x1000u) why |= R_WR_CART;
      else if (ea & 0x80u) why |= R_WR_RIOT;
      else if (a6 < 4) why |= (a6 == 2) ? R_WR_WSYNC : R_WR_TIA_LOW;
      else if (MN_FILL(r.fifo_n) >= MN_FIFO_CAP) why |= R_FIFO_FULL;
    }
  }
  return why;
}

int main(int argc, char** argv) {
  const char* game = argc > 1 ? argv[1] : "ms_pacman";
  const int L = argc > 2 ? atoi(argv[2]) : 32;
  const int rounds = argc > 3 ? atoi(argv[3]) : 6;
  const int slack = argc > 4 ? atoi(argv[4]) : 4;
  const int decor = argc > 5 ? atoi(argv[5]) : 150;
  char path[512];
  snprintf(path, sizeof path, "%s/atari_roms/%s.bin", getenv("MN_ROOT") ? getenv("MN_ROOT") : ".", game);
  FILE* f = fopen(path, "rb");
  if (!f) { fprintf(stderr, "no rom %s\n", path); return 1; }
  g_rom.resize(16384);
  const size_t n = fread(g_rom.data(), 1, 16384, f);
  fclose(f);
  g_rom.resize(n);
  build_tables(&g_tab);
  const int g = game_id_from_name(game);
  const GameEntry& ge = game_db(g);
  std::vector<Lane> lanes(L);
  srand(12345);
  for (int i = 0; i < L; ++i) {
    Lane& l = lanes[i];
    memset(&l.s, 0, sizeof l.s);
    l.fb.assign(2 * MN_FRAME_BYTES, 0);
    l.s.game = uint8_t(g); l.s.cart = uint8_t(detect_cart(g_rom.data(), n)); l.s.ctrl = uint8_t(ge.ctrl);
    l.c.s = &l.s; l.c.rom = g_rom.data(); l.c.ram = l.ram; l.c.fb = l.fb.data(); l.c.tab = &g_tab;
    l.c.fifo = l.fifo; l.c.fifo_n = 0; l.c.hseq = 0; l.c.mbox_timeout = false;
    run_alone(l, U_POWER_ON, 0, 0, 3u * uint32_t(i + 1));
    // decorrelate: random macro actions with FiGAR-like repeats, resets on game over
    int done = 0;
    while (done < decor) {
      const int a = ge.actions[rand() % ge.n_actions], rep = 1 + rand() % 11;
      for (int k = 0; k < rep; ++k) {
        run_alone(l, U_ACTS, a, 4, 0);
        ++done;
        if (l.s.flags & F_TERMINAL) { run_alone(l, U_RESET, 0, 0, rng_next(l.s.rng)); for (int q = 0; q < 4; ++q) run_alone(l, U_ACTS, 0, 4, 0); break; }
      }
    }
  }
  // ---- the warp loop of k_round (pool.cu), instrumented
  unsigned long long ticks = 0, lane_ticks = 0, lane_instr = 0, ticks_any_slow[32] = {0}, lanes_slow[32] = {0}, distinct_pc = 0, job_ticks = 0;
  unsigned long long drains = 0;
  // nested coverage sets: each level ADDS categories to the fast tick
  const uint32_t level_adds[] = {
      0,
      R_FLAG,
      R_JMP,
      R_IND_RAM,
      R_BIT,
      R_JSR | R_RTS,
      R_PHA | R_PLA,
      R_RD_TIMER,
      R_WR_WSYNC,
      R_RD_TIA,
      R_DECIMAL,
  };
  const char* level_names[] = {"generic+branch only", "+flag ops", "+JMP abs", "+(zp,X)/(zp),Y via RAM", "+BIT", "+JSR/RTS", "+PHA/PLA", "+INTIM read",
                               "+WSYNC", "+TIA reads", "+decimal"};
  const int n_levels = int(sizeof(level_adds) / sizeof(level_adds[0]));
  std::map<uint32_t, unsigned long long> why_any;   // per single reason bit: ticks in which any lane had it
  unsigned long long why_lane[32] = {0};
  for (int rd = 0; rd < rounds; ++rd) {
    for (int i = 0; i < L; ++i) {
      Lane& l = lanes[i];
      if (l.s.flags & F_TERMINAL) { run_alone(l, U_RESET, 0, 0, rng_next(l.s.rng)); for (int q = 0; q < 4; ++q) run_alone(l, U_ACTS, 0, 4, 0); }
      l.c.all_pixels = false;
      unit_init(l.c, l.u, U_ACTS, ge.actions[rand() % ge.n_actions], 4, 0);
      hot_init(l.c, l.u, l.hot);
    }
    for (;;) {
      int first = 0x7FFFFFFF;
      for (int i = 0; i < L; ++i) if (hot_has_work(lanes[i].hot)) { const int t = hot_time(lanes[i].hot); if (t < first) first = t; }
      if (first == 0x7FFFFFFF) break;
      ++ticks;
      uint32_t any = 0;
      std::set<uint32_t> pcs;
      uint32_t lane_why[64];
      int n_el = 0;
      for (int i = 0; i < L; ++i) {
        Lane& l = lanes[i];
        lane_why[i] = 0xFFFFFFFFu;
        if (!hot_has_work(l.hot) || hot_time(l.hot) - first > slack) continue;
        ++lane_ticks; ++n_el;
        uint32_t why;
        if (!l.hot.in_frame) why = R_JOB;
        else { why = classify(l); pcs.insert(l.hot.cpu.PC); ++lane_instr; }
        lane_why[i] = why;
        any |= why;
        for (int b = 0; b < 26; ++b) if (why & (1u << b)) ++why_lane[b];
      }
      if (any & R_JOB) ++job_ticks;
      distinct_pc += pcs.size();
      for (int b = 0; b < 26; ++b) if (any & (1u << b)) ++why_any[1u << b];
      uint32_t covered = 0;
      for (int lv = 0; lv < n_levels; ++lv) {
        covered |= level_adds[lv];
        bool slow = false; int ns = 0;
        for (int i = 0; i < L; ++i) if (lane_why[i] != 0xFFFFFFFFu && (lane_why[i] & ~covered)) { slow = true; ++ns; }
        if (slow) { ++ticks_any_slow[lv]; lanes_slow[lv] += ns; }
      }
      // execute the tick
      bool drain = false;
      for (int i = 0; i < L; ++i) {
        Lane& l = lanes[i];
        if (lane_why[i] == 0xFFFFFFFFu) continue;
        const Mem mm = mem_of(l.c);
        unit_tick<false>(l.c, mm, l.u, l.hot);
      }
      for (int i = 0; i < L; ++i) if (MN_FILL(lanes[i].hot.cpu.fifo_n) >= MN_FIFO_HIGH) drain = true;
      if (drain) { ++drains; for (int i = 0; i < L; ++i) hot_drain(lanes[i].c, lanes[i].hot); }
    }
    for (int i = 0; i < L; ++i) unit_finish(lanes[i].c, lanes[i].hot);
  }
  printf("game %s lanes %d rounds %d slack %d\n", game, L, rounds, slack);
  printf("ticks/next %.0f  6502 instr/next/lane %.0f  eligible lanes/tick %.2f  distinct PCs/tick %.2f  job ticks %.2f%%  drains/next %.1f\n",
         double(ticks) / rounds, double(lane_instr) / rounds / L, double(lane_ticks) / ticks, double(distinct_pc) / ticks,
         100.0 * job_ticks / ticks, double(drains) / rounds);
  printf("%-26s %12s %14s\n", "reason", "any-lane %tick", "% lane-instr");
  for (int b = 0; b < 26; ++b) if (why_lane[b]) printf("%-26s %11.2f%% %13.2f%%\n", kNames[b], 100.0 * why_any[1u << b] / ticks, 100.0 * why_lane[b] / lane_ticks);
  printf("%-28s %16s %18s\n", "fast tick covers", "ticks w/ slow lane", "slow lanes / such tick");
  for (int lv = 0; lv < n_levels; ++lv)
    printf("%-28s %15.2f%% %18.2f\n", level_names[lv], 100.0 * ticks_any_slow[lv] / ticks, ticks_any_slow[lv] ? double(lanes_slow[lv]) / ticks_any_slow[lv] : 0.0);
  return 0;
}
