// ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the shipped product path.
//
// Scalar CPU restatement of the Atari 2600 machine that sits underneath the
// reference's `ALEInterface.act()` call (reference call sites:
// atari_emulator.py:19-31,57-61,72-77,94-97,121,128,133).  The reference repo
// does not vendor that emulator: it is the third-party Arcade-Learning-
// Environment (ALE 0.5.x/0.6.x, a fork of Stella 2.x) which is absent from
// /root/reference and from this image.  PARITY UNPINNED: this file restates the
// published behaviour of that machine (6502 "low" core with cycles charged up
// front, M6532 RIOT, event-driven/span-rendered TIA, 2K/4K/F8/F6/E0 carts) from
// documentation and recollection; it cannot be diffed against real ALE here.
// It is the single source of truth the CUDA path is checked against.
//
// Style: deliberately simple -- per-pixel loops, lookup tables built at start-up,
// one big struct.  The CUDA kernel (manette_b200/csrc/emu_kernel.cu) is written
// independently with a different formulation (160-bit line masks, packed decode
// descriptors); only the behaviour is shared.
#pragma once
#include <cstdint>
#include <cstring>

namespace orc {

enum CartType { CART_2K = 0, CART_4K = 1, CART_F8 = 2, CART_F6 = 3, CART_E0 = 4 };
enum CtrlType { CTRL_JOYSTICK = 0, CTRL_PADDLES = 1, CTRL_PADDLES_SWAPPED = 2 };

static const int HBLANK = 68;
static const int SCREEN_W = 160;
static const int SCREEN_H = 210;
static const int YSTART = 34;
static const int MAX_SCANLINES = 290;
static const int32_t RES_MIN = 0;            // Controller::minimumResistance
static const int32_t RES_MAX = 0x7FFFFFFF;   // Controller::maximumResistance

// enabled-object bits
enum { P0Bit = 0x01, M0Bit = 0x02, P1Bit = 0x04, M1Bit = 0x08, BLBit = 0x10, PFBit = 0x20,
       ScoreBit = 0x40, PriorityBit = 0x80 };

// ---------------------------------------------------------------- tables
struct Tables {
  uint8_t  player_mask[2][8][160];    // [suppress-first-copy][nusiz&7][distance from POS]
  uint8_t  missile_mask[8][4][160];   // [nusiz&7][size][distance]
  uint8_t  ball_mask[4][160];         // [size][distance]
  uint32_t pf_mask[2][160];           // [reflect][x] -> bit in 20-bit PF register
  int8_t   reset_when[8][160][160];   // [nusiz&7][oldx][newx]: 1 display, -1 delay, 0 neither
  uint16_t collision[64];             // [enabled bits] -> collision latch bits
  uint8_t  priority[2][256];          // [right half][enabled|prio|score] -> colour slot 0..3 (P0,P1,PF,BK)
  int8_t   motion[76][16];            // [cpu cycle of HMOVE][HMxx>>4] -> delta x
  bool     hmove_blank[76];
  int8_t   poke_delay[64];
  uint8_t  reflect[256];
  uint8_t  cycles[256];
  Tables();
};

inline Tables::Tables() {
  std::memset(this, 0, sizeof(*this));
  // --- player mask: bit of GRP shown `d` pixels right of POS (0x80 = leftmost)
  for (int sup = 0; sup < 2; ++sup)
    for (int mode = 0; mode < 8; ++mode)
      for (int x = 0; x < 160; ++x) {
        uint8_t m = 0;
        auto copy = [&](int off, bool first) {
          if (first && sup) return;
          if (x >= off && x < off + 8) m = uint8_t(0x80 >> (x - off));
        };
        switch (mode) {
          case 0: copy(0, true); break;
          case 1: copy(0, true); copy(16, false); break;
          case 2: copy(0, true); copy(32, false); break;
          case 3: copy(0, true); copy(16, false); copy(32, false); break;
          case 4: copy(0, true); copy(64, false); break;
          case 5: if (!sup && x > 0 && x <= 16) m = uint8_t(0x80 >> ((x - 1) / 2)); break;  // 1 px late
          case 6: copy(0, true); copy(32, false); copy(64, false); break;
          case 7: if (!sup && x > 0 && x <= 32) m = uint8_t(0x80 >> ((x - 1) / 4)); break;  // 1 px late
        }
        player_mask[sup][mode][x] = m;
      }
  // --- missile mask
  for (int mode = 0; mode < 8; ++mode)
    for (int size = 0; size < 4; ++size)
      for (int x = 0; x < 160; ++x) {
        int w = 1 << size;
        bool on = (x < w);
        auto copy = [&](int off) { if (x >= off && x < off + w) on = true; };
        switch (mode) {
          case 1: copy(16); break;
          case 2: copy(32); break;
          case 3: copy(16); copy(32); break;
          case 4: copy(64); break;
          case 6: copy(32); copy(64); break;
          default: break;   // 0,5,7: single copy
        }
        missile_mask[mode][size][x] = on;
      }
  for (int size = 0; size < 4; ++size)
    for (int x = 0; x < 160; ++x) ball_mask[size][x] = (x < (1 << size));
  // --- playfield bit selectors
  for (int x = 0; x < 160; ++x) {
    int h = x % 80;
    uint32_t normal, mirrored;
    if (h < 16) normal = 0x00001u << (h / 4);
    else if (h < 48) normal = 0x00800u >> ((h - 16) / 4);
    else normal = 0x01000u << ((h - 48) / 4);
    if (h < 32) mirrored = 0x80000u >> (h / 4);
    else if (h < 64) mirrored = 0x00010u << ((h - 32) / 4);
    else mirrored = 0x00008u >> ((h - 64) / 4);
    pf_mask[0][x] = normal;
    pf_mask[1][x] = (x < 80) ? normal : mirrored;
  }
  // --- where does a RESPx land relative to the copies currently being drawn
  for (int mode = 0; mode < 8; ++mode)
    for (int oldx = 0; oldx < 160; ++oldx) {
      static const int offs[8][3] = {{0, -1, -1}, {0, 16, -1}, {0, 32, -1}, {0, 16, 32},
                                     {0, 64, -1}, {0, -1, -1}, {0, 32, 64}, {0, -1, -1}};
      int width = (mode == 5) ? 16 : (mode == 7) ? 32 : 8;
      for (int newx = 0; newx < 160 + 72 + 5; ++newx)
        for (int c = 0; c < 3; ++c) {
          int off = offs[mode][c];
          if (off < 0) continue;
          if (newx >= oldx + off && newx < oldx + off + 4) reset_when[mode][oldx][newx % 160] = -1;
          if (newx >= oldx + off + 4 && newx < oldx + off + 4 + width) reset_when[mode][oldx][newx % 160] = 1;
        }
    }
  // --- collisions
  for (int e = 0; e < 64; ++e) {
    uint16_t c = 0;
    if ((e & M0Bit) && (e & P1Bit)) c |= 0x0001;
    if ((e & M0Bit) && (e & P0Bit)) c |= 0x0002;
    if ((e & M1Bit) && (e & P0Bit)) c |= 0x0004;
    if ((e & M1Bit) && (e & P1Bit)) c |= 0x0008;
    if ((e & P0Bit) && (e & PFBit)) c |= 0x0010;
    if ((e & P0Bit) && (e & BLBit)) c |= 0x0020;
    if ((e & P1Bit) && (e & PFBit)) c |= 0x0040;
    if ((e & P1Bit) && (e & BLBit)) c |= 0x0080;
    if ((e & M0Bit) && (e & PFBit)) c |= 0x0100;
    if ((e & M0Bit) && (e & BLBit)) c |= 0x0200;
    if ((e & M1Bit) && (e & PFBit)) c |= 0x0400;
    if ((e & M1Bit) && (e & BLBit)) c |= 0x0800;
    if ((e & BLBit) && (e & PFBit)) c |= 0x1000;
    if ((e & P0Bit) && (e & P1Bit)) c |= 0x2000;
    if ((e & M0Bit) && (e & M1Bit)) c |= 0x4000;
    collision[e] = c;
  }
  // --- priority encoder -> colour slot (0 P0, 1 P1, 2 PF, 3 BK)
  for (int half = 0; half < 2; ++half)
    for (int e = 0; e < 256; ++e) {
      uint8_t col = 3;
      if (e & PriorityBit) {
        if (e & M1Bit) col = 1;
        if (e & P1Bit) col = 1;
        if (e & M0Bit) col = 0;
        if (e & P0Bit) col = 0;
        if (e & BLBit) col = 2;
        if (e & PFBit) col = 2;
      } else {
        if (e & BLBit) col = 2;
        if (e & PFBit) col = (e & ScoreBit) ? (half == 0 ? 0 : 1) : 2;
        if (e & M1Bit) col = 1;
        if (e & P1Bit) col = 1;
        if (e & M0Bit) col = 0;
        if (e & P0Bit) col = 0;
      }
      priority[half][e] = col;
    }
  // --- HMOVE motion by CPU cycle of the strobe (derived from the TIA extra-clock model:
  // pulse k (1..15) is sent at colour clock 3x+5+4k and only counts while HBLANK (extended
  // to 76 when the strobe itself fell inside HBLANK) is active)
  for (int x = 0; x < 76; ++x)
    for (int h = 0; h < 16; ++h) {
      int n = h ^ 8;          // number of extra clocks requested
      int mv;
      if (x <= 22) {
        int fit = (70 - 3 * x) / 4; if (fit < 0) fit = 0;
        int cnt = n < fit ? n : fit;
        mv = 8 - cnt;
      } else if (x == 75) {
        mv = 8 - n;
      } else {
        int k0 = (223 - 3 * x + 3) / 4; if (k0 < 1) k0 = 1;   // first pulse inside next HBLANK
        int cnt = n - (k0 - 1); if (cnt < 0) cnt = 0;
        mv = -cnt;
      }
      motion[x][h] = int8_t(mv);
    }
  for (int x = 0; x < 76; ++x) hmove_blank[x] = (x <= 20) || (x == 75);
  static const int8_t pd[64] = {0, 1, 0, 0, 8, 8, 0, 0, 0, 0, 0, 1, 1, -1, -1, -1,
                                0, 0, 8, 8, 0, 0, 0, 0, 0, 0, 0, 1, 1, 0, 0, 0,
                                0, 0, 0, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0,
                                0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  std::memcpy(poke_delay, pd, 64);
  for (int i = 0; i < 256; ++i) {
    uint8_t r = 0;
    for (int b = 0; b < 8; ++b) if (i & (1 << b)) r |= uint8_t(0x80 >> b);
    reflect[i] = r;
  }
  static const uint8_t cyc[256] = {
      7, 6, 2, 8, 3, 3, 5, 5, 3, 2, 2, 2, 4, 4, 6, 6, 2, 5, 2, 8, 4, 4, 6, 6, 2, 4, 2, 7, 4, 4, 7, 7,
      6, 6, 2, 8, 3, 3, 5, 5, 4, 2, 2, 2, 4, 4, 6, 6, 2, 5, 2, 8, 4, 4, 6, 6, 2, 4, 2, 7, 4, 4, 7, 7,
      6, 6, 2, 8, 3, 3, 5, 5, 3, 2, 2, 2, 3, 4, 6, 6, 2, 5, 2, 8, 4, 4, 6, 6, 2, 4, 2, 7, 4, 4, 7, 7,
      6, 6, 2, 8, 3, 3, 5, 5, 4, 2, 2, 2, 5, 4, 6, 6, 2, 5, 2, 8, 4, 4, 6, 6, 2, 4, 2, 7, 4, 4, 7, 7,
      2, 6, 2, 6, 3, 3, 3, 3, 2, 2, 2, 2, 4, 4, 4, 4, 2, 6, 2, 6, 4, 4, 4, 4, 2, 5, 2, 5, 5, 5, 5, 5,
      2, 6, 2, 6, 3, 3, 3, 3, 2, 2, 2, 2, 4, 4, 4, 4, 2, 5, 2, 5, 4, 4, 4, 4, 2, 4, 2, 4, 4, 4, 4, 4,
      2, 6, 2, 8, 3, 3, 5, 5, 2, 2, 2, 2, 4, 4, 6, 6, 2, 5, 2, 8, 4, 4, 6, 6, 2, 4, 2, 7, 4, 4, 7, 7,
      2, 6, 2, 8, 3, 3, 5, 5, 2, 2, 2, 2, 4, 4, 6, 6, 2, 5, 2, 8, 4, 4, 6, 6, 2, 4, 2, 7, 4, 4, 7, 7};
  std::memcpy(cycles, cyc, 256);
}

inline const Tables& tables() { static Tables t; return t; }

// ---------------------------------------------------------------- decode
enum Mode : uint8_t { IMP, ACC, IMM, ZP, ZPX, ZPY, ABS, ABX, ABY, IZX, IZY, REL, IND };
enum Op : uint8_t {
  ADC, AND_, ASL, BCC, BCS, BEQ, BIT, BMI, BNE, BPL, BRK, BVC, BVS, CLC, CLD, CLI, CLV, CMP, CPX, CPY,
  DEC, DEX, DEY, EOR, INC, INX, INY, JMP, JSR, LDA, LDX, LDY, LSR, NOP, ORA, PHA, PHP, PLA, PLP, ROL,
  ROR, RTI, RTS, SBC, SEC, SED, SEI, STA, STX, STY, TAX, TAY, TSX, TXA, TXS, TYA,
  // undocumented
  SLO, RLA, SRE, RRA, SAX, LAX, DCP, ISC, ANC, ALR, ARR, XAA, AXS, AHX, SHY, SHX, TAS, LAS, LXA, KIL };

struct Decoded { Op op; Mode mode; };

inline const Decoded* decode_table() {
  static Decoded t[256];
  static bool init = false;
  if (init) return t;
  for (int i = 0; i < 256; ++i) t[i] = {NOP, IMP};
  auto set = [&](int opc, Op op, Mode m) { t[opc] = {op, m}; };
  // ALU group cc=01
  const Op alu[8] = {ORA, AND_, EOR, ADC, STA, LDA, CMP, SBC};
  const Mode m01[8] = {IZX, ZP, IMM, ABS, IZY, ZPX, ABY, ABX};
  for (int a = 0; a < 8; ++a) for (int b = 0; b < 8; ++b) set(a << 5 | b << 2 | 1, alu[a], m01[b]);
  set(0x89, NOP, IMM);
  // RMW group cc=10
  const Op rmw[8] = {ASL, ROL, LSR, ROR, STX, LDX, DEC, INC};
  for (int a = 0; a < 8; ++a) {
    set(a << 5 | 0x06, rmw[a], ZP);
    set(a << 5 | 0x0E, rmw[a], ABS);
    set(a << 5 | 0x16, rmw[a], (a == 4 || a == 5) ? ZPY : ZPX);
    set(a << 5 | 0x1E, rmw[a], (a == 4 || a == 5) ? ABY : ABX);
  }
  set(0x0A, ASL, ACC); set(0x2A, ROL, ACC); set(0x4A, LSR, ACC); set(0x6A, ROR, ACC);
  set(0xA2, LDX, IMM); set(0x9E, SHX, ABY);
  set(0x8A, TXA, IMP); set(0x9A, TXS, IMP); set(0xAA, TAX, IMP); set(0xBA, TSX, IMP);
  set(0xCA, DEX, IMP); set(0xEA, NOP, IMP);
  // undocumented cc=11 (combination of the two above)
  const Op ill[8] = {SLO, RLA, SRE, RRA, SAX, LAX, DCP, ISC};
  for (int a = 0; a < 8; ++a) {
    set(a << 5 | 0x03, ill[a], IZX);
    set(a << 5 | 0x07, ill[a], ZP);
    set(a << 5 | 0x0F, ill[a], ABS);
    set(a << 5 | 0x13, ill[a], IZY);
    set(a << 5 | 0x17, ill[a], (a == 4 || a == 5) ? ZPY : ZPX);
    set(a << 5 | 0x1B, ill[a], ABY);
    set(a << 5 | 0x1F, ill[a], (a == 4 || a == 5) ? ABY : ABX);
  }
  set(0x0B, ANC, IMM); set(0x2B, ANC, IMM); set(0x4B, ALR, IMM); set(0x6B, ARR, IMM);
  set(0x8B, XAA, IMM); set(0xAB, LXA, IMM); set(0xCB, AXS, IMM); set(0xEB, SBC, IMM);
  set(0x93, AHX, IZY); set(0x9F, AHX, ABY); set(0x9B, TAS, ABY); set(0xBB, LAS, ABY);
  // cc=00
  set(0x00, BRK, IMP); set(0x20, JSR, ABS); set(0x40, RTI, IMP); set(0x60, RTS, IMP);
  set(0x08, PHP, IMP); set(0x28, PLP, IMP); set(0x48, PHA, IMP); set(0x68, PLA, IMP);
  set(0x88, DEY, IMP); set(0xA8, TAY, IMP); set(0xC8, INY, IMP); set(0xE8, INX, IMP);
  set(0x18, CLC, IMP); set(0x38, SEC, IMP); set(0x58, CLI, IMP); set(0x78, SEI, IMP);
  set(0x98, TYA, IMP); set(0xB8, CLV, IMP); set(0xD8, CLD, IMP); set(0xF8, SED, IMP);
  set(0x10, BPL, REL); set(0x30, BMI, REL); set(0x50, BVC, REL); set(0x70, BVS, REL);
  set(0x90, BCC, REL); set(0xB0, BCS, REL); set(0xD0, BNE, REL); set(0xF0, BEQ, REL);
  set(0x24, BIT, ZP); set(0x2C, BIT, ABS); set(0x4C, JMP, ABS); set(0x6C, JMP, IND);
  set(0x84, STY, ZP); set(0x8C, STY, ABS); set(0x94, STY, ZPX); set(0x9C, SHY, ABX);
  set(0xA0, LDY, IMM); set(0xA4, LDY, ZP); set(0xAC, LDY, ABS); set(0xB4, LDY, ZPX); set(0xBC, LDY, ABX);
  set(0xC0, CPY, IMM); set(0xC4, CPY, ZP); set(0xCC, CPY, ABS);
  set(0xE0, CPX, IMM); set(0xE4, CPX, ZP); set(0xEC, CPX, ABS);
  // multi-byte NOPs
  set(0x80, NOP, IMM); set(0x82, NOP, IMM); set(0xC2, NOP, IMM); set(0xE2, NOP, IMM);
  set(0x04, NOP, ZP); set(0x44, NOP, ZP); set(0x64, NOP, ZP);
  set(0x0C, NOP, ABS);
  for (int a : {0x14, 0x34, 0x54, 0x74, 0xD4, 0xF4}) set(a, NOP, ZPX);
  for (int a : {0x1C, 0x3C, 0x5C, 0x7C, 0xDC, 0xFC}) set(a, NOP, ABX);
  for (int a : {0x1A, 0x3A, 0x5A, 0x7A, 0xDA, 0xFA}) set(a, NOP, IMP);
  for (int a : {0x02, 0x12, 0x22, 0x32, 0x42, 0x52, 0x62, 0x72, 0x92, 0xB2, 0xD2, 0xF2}) set(a, KIL, IMP);
  init = true;
  return t;
}

// ---------------------------------------------------------------- the console
struct Console {
  // cartridge
  const uint8_t* rom = nullptr; uint32_t rom_size = 0; int cart = CART_4K;
  uint8_t bank = 0; uint8_t slice[3] = {4, 5, 6};
  // 6502
  uint8_t A = 0, X = 0, Y = 0, SP = 0xFF; uint16_t PC = 0;
  bool N = false, V = false, B = false, D = false, I = false, notZ = true, C = false;
  int32_t cycles = 0; uint8_t dbus = 0; bool stop = false;
  // RIOT
  uint8_t ram[128] = {0};
  uint8_t timer = 0, shift = 6; int32_t timer_set_cycle = 0, irq_reset_cycle = 0; bool read_after_irq = false;
  uint8_t ddra = 0, ddrb = 0;
  // inputs, latched by the environment layer before every frame
  uint8_t swcha = 0xFF, swchb = 0x3F;
  bool inpt4_high = true, inpt5_high = true;
  int32_t analog[4] = {RES_MAX, RES_MAX, RES_MAX, RES_MAX};   // INPT0..3 resistances
  // TIA timing
  int32_t clk_frame_start = 0, clk_start_display = 0, clk_stop_display = 0, clk_last_update = 0;
  int32_t clks_to_eol = 228, vsync_finish_clk = 0x7FFFFFFF;
  bool partial_frame = false;
  // TIA registers
  uint8_t VSYNC = 0, VBLANK = 0, NUSIZ0 = 0, NUSIZ1 = 0, CTRLPF = 0, color[4] = {0, 0, 0, 0}, prio_score = 0;
  bool REFP0 = false, REFP1 = false;
  uint32_t PF = 0;
  uint8_t GRP0 = 0, GRP1 = 0, DGRP0 = 0, DGRP1 = 0, cur_grp0 = 0, cur_grp1 = 0;
  bool ENAM0 = false, ENAM1 = false, ENABL = false, DENABL = false;
  uint8_t HMP0 = 0, HMP1 = 0, HMM0 = 0, HMM1 = 0, HMBL = 0;
  bool VDELP0 = false, VDELP1 = false, VDELBL = false, RESMP0 = false, RESMP1 = false;
  uint16_t collision = 0;
  int16_t POSP0 = 0, POSP1 = 0, POSM0 = 0, POSM1 = 0, POSBL = 0;
  uint8_t p0_suppress = 0, p1_suppress = 0, pf_reflect_cur = 0;
  int32_t last_hmove_clk = 0; bool hmove_blank = false;
  uint8_t enabled = 0;
  bool dump_enabled = false; int32_t dump_disabled_cycle = 0;
  // frame buffers (double buffered like the emulated TIA)
  uint8_t fb[2][SCREEN_W * SCREEN_H]; int cur_fb = 0; int32_t fb_pos = 0;

  Console() { std::memset(fb, 0, sizeof(fb)); }

  // ------------------------------------------------------------ cartridge
  void cart_reset() {
    bank = (cart == CART_F8) ? 1 : 0;   // F8 starts in the upper bank, F6 in bank 0
    slice[0] = 4; slice[1] = 5; slice[2] = 6;
  }
  inline void cart_hotspot(uint16_t a) {   // a = addr & 0x0FFF
    if (cart == CART_F8) { if (a == 0xFF8) bank = 0; else if (a == 0xFF9) bank = 1; }
    else if (cart == CART_F6) { if (a >= 0xFF6 && a <= 0xFF9) bank = uint8_t(a - 0xFF6); }
    else if (cart == CART_E0) {
      if (a >= 0xFE0 && a <= 0xFE7) slice[0] = a & 7;
      else if (a >= 0xFE8 && a <= 0xFEF) slice[1] = a & 7;
      else if (a >= 0xFF0 && a <= 0xFF7) slice[2] = a & 7;
    }
  }
  inline uint8_t cart_peek(uint16_t addr) {
    uint16_t a = addr & 0x0FFF;
    switch (cart) {
      case CART_2K: return rom[a & 0x7FF];
      case CART_4K: return rom[a];
      case CART_F8: case CART_F6: cart_hotspot(a); return rom[(uint32_t(bank) << 12) + a];
      default: {  // E0
        cart_hotspot(a);
        uint32_t seg = a >> 10;
        uint32_t sl = (seg == 3) ? 7 : slice[seg];
        return rom[(sl << 10) + (a & 0x3FF)];
      }
    }
  }
  inline void cart_poke(uint16_t addr) { cart_hotspot(addr & 0x0FFF); }

  // ------------------------------------------------------------ bus
  inline uint8_t peek(uint16_t addr) {
    addr &= 0x1FFF;
    uint8_t v;
    if (addr & 0x1000) v = cart_peek(addr);
    else if (!(addr & 0x0080)) v = tia_peek(addr);
    else if (!(addr & 0x0200)) v = ram[addr & 0x7F];
    else v = riot_peek(addr);
    dbus = v;
    return v;
  }
  inline void poke(uint16_t addr, uint8_t v) {
    addr &= 0x1FFF;
    if (addr & 0x1000) cart_poke(addr);
    else if (!(addr & 0x0080)) tia_poke(addr, v);
    else if (!(addr & 0x0200)) ram[addr & 0x7F] = v;
    else riot_poke(addr, v);
    dbus = v;
  }

  // ------------------------------------------------------------ RIOT (M6532)
  void riot_reset(uint32_t rnd) {
    timer = uint8_t(25 + (rnd % 75)); shift = 6; timer_set_cycle = 0; irq_reset_cycle = 0;
    read_after_irq = false; ddra = 0; ddrb = 0;
  }
  uint8_t riot_peek(uint16_t addr) {
    switch (addr & 0x07) {
      case 0x00: return swcha;
      case 0x01: return ddra;
      case 0x02: return swchb;
      case 0x03: return ddrb;
      case 0x04: case 0x06: {
        uint32_t delta = uint32_t((cycles - 1) - timer_set_cycle);
        int32_t t = int32_t(timer) - int32_t(delta >> shift) - 1;
        if (t >= 0) return uint8_t(t);
        t = int32_t(uint32_t(timer) << shift) - int32_t(delta) - 1;
        if (t <= -2 && !read_after_irq) { read_after_irq = true; irq_reset_cycle = cycles; }
        if (read_after_irq) {
          int32_t offset = irq_reset_cycle - (timer_set_cycle + int32_t(uint32_t(timer) << shift));
          t = int32_t(timer) - int32_t(delta >> shift) - offset;
        }
        return uint8_t(t);
      }
      default: {   // 0x05, 0x07 interrupt flag
        uint32_t delta = uint32_t((cycles - 1) - timer_set_cycle);
        int32_t t = int32_t(timer) - int32_t(delta >> shift) - 1;
        return (t >= 0 || read_after_irq) ? 0x00 : 0x80;
      }
    }
  }
  void riot_poke(uint16_t addr, uint8_t v) {
    if ((addr & 0x07) == 0x01) ddra = v;
    else if ((addr & 0x07) == 0x03) ddrb = v;
    else if ((addr & 0x14) == 0x14) {
      static const uint8_t sh[4] = {0, 3, 6, 10};
      timer = v; shift = sh[addr & 0x03]; timer_set_cycle = cycles; read_after_irq = false;
    }
  }

  // ------------------------------------------------------------ TIA
  void tia_reset() {
    clk_frame_start = 0; clk_start_display = 228 * YSTART; clk_stop_display = clk_start_display + 228 * SCREEN_H;
    clk_last_update = 0; clks_to_eol = 228; vsync_finish_clk = 0x7FFFFFFF; partial_frame = false;
    VSYNC = VBLANK = NUSIZ0 = NUSIZ1 = CTRLPF = 0; color[0] = color[1] = color[2] = color[3] = 0; prio_score = 0;
    REFP0 = REFP1 = false; PF = 0; GRP0 = GRP1 = DGRP0 = DGRP1 = cur_grp0 = cur_grp1 = 0;
    ENAM0 = ENAM1 = ENABL = DENABL = false; HMP0 = HMP1 = HMM0 = HMM1 = HMBL = 0;
    VDELP0 = VDELP1 = VDELBL = RESMP0 = RESMP1 = false; collision = 0;
    POSP0 = POSP1 = POSM0 = POSM1 = POSBL = 0; p0_suppress = p1_suppress = 0; pf_reflect_cur = 0;
    last_hmove_clk = 0; hmove_blank = false; enabled = 0; dump_enabled = false; dump_disabled_cycle = 0;
    std::memset(fb, 0, sizeof(fb)); cur_fb = 0; fb_pos = 0;
  }

  void start_frame() {
    cur_fb ^= 1;
    int32_t clocks = ((cycles * 3) - clk_frame_start) % 228;
    // rebase every cycle-stamped quantity to the new frame (System::resetCycles)
    int32_t c = cycles;
    timer_set_cycle -= c; irq_reset_cycle -= c; dump_disabled_cycle -= c;
    last_hmove_clk -= c * 3;
    if (vsync_finish_clk != 0x7FFFFFFF) vsync_finish_clk -= c * 3;
    cycles = 0;
    clk_frame_start = -clocks;
    clk_start_display = clk_frame_start + 228 * YSTART;
    clk_stop_display = clk_start_display + 228 * SCREEN_H;
    clk_last_update = clk_start_display;
    clks_to_eol = 228;
    fb_pos = 0;
  }

  void render_span(int32_t n, int32_t hpos) {
    const Tables& T = tables();
    uint8_t* out = fb[cur_fb] + fb_pos;
    if (VBLANK & 0x02) { std::memset(out, 0, n); fb_pos += n; return; }
    const uint8_t* p0m = T.player_mask[p0_suppress][NUSIZ0 & 7];
    const uint8_t* p1m = T.player_mask[p1_suppress][NUSIZ1 & 7];
    const uint8_t* m0m = T.missile_mask[NUSIZ0 & 7][(NUSIZ0 >> 4) & 3];
    const uint8_t* m1m = T.missile_mask[NUSIZ1 & 7][(NUSIZ1 >> 4) & 3];
    const uint8_t* blm = T.ball_mask[(CTRLPF >> 4) & 3];
    const uint32_t* pfm = T.pf_mask[pf_reflect_cur];
    for (int32_t i = 0; i < n; ++i, ++hpos) {
      uint8_t e = 0;
      if ((enabled & PFBit) && (PF & pfm[hpos])) e |= PFBit;
      if ((enabled & BLBit) && blm[(hpos - POSBL + 160) % 160]) e |= BLBit;
      if ((enabled & P1Bit) && (cur_grp1 & p1m[(hpos - POSP1 + 160) % 160])) e |= P1Bit;
      if ((enabled & M1Bit) && m1m[(hpos - POSM1 + 160) % 160]) e |= M1Bit;
      if ((enabled & P0Bit) && (cur_grp0 & p0m[(hpos - POSP0 + 160) % 160])) e |= P0Bit;
      if ((enabled & M0Bit) && m0m[(hpos - POSM0 + 160) % 160]) e |= M0Bit;
      collision |= T.collision[e];
      out[i] = color[T.priority[hpos < 80 ? 0 : 1][e | prio_score]];
    }
    fb_pos += n;
  }

  void update_frame(int32_t clock) {
    if (clock < clk_start_display || clk_last_update >= clk_stop_display || clk_last_update >= clock) return;
    if (clock > clk_stop_display) clock = clk_stop_display;
    do {
      int32_t from_sol = 228 - clks_to_eol;
      int32_t n;
      if (clock > clk_last_update + clks_to_eol) { n = clks_to_eol; clks_to_eol = 228; clk_last_update += n; }
      else { n = clock - clk_last_update; clks_to_eol -= n; clk_last_update = clock; }
      if (from_sol < HBLANK) {
        int32_t skip = HBLANK - from_sol; if (skip > n) skip = n;
        from_sol += skip; n -= skip;
      }
      int32_t old_pos = fb_pos;
      if (n != 0) render_span(n, from_sol - HBLANK);
      if (hmove_blank && from_sol < HBLANK + 8) {
        int32_t blanks = (HBLANK + 8) - from_sol;
        int32_t room = SCREEN_W * SCREEN_H - old_pos; if (blanks > room) blanks = room;
        std::memset(fb[cur_fb] + old_pos, 0, blanks);
        if (n + from_sol >= HBLANK + 8) hmove_blank = false;
      }
      if (clks_to_eol == 228) {   // end of scan line
        pf_reflect_cur = CTRLPF & 0x01;
        p0_suppress = 0; p1_suppress = 0;
      }
    } while (clk_last_update < clock);
  }

  uint8_t tia_peek(uint16_t addr) {
    update_frame(cycles * 3);
    uint8_t noise = dbus & 0x3F;
    auto cx = [&](uint16_t hi, uint16_t lo) -> uint8_t {
      return uint8_t(((collision & hi) ? 0x80 : 0) | ((collision & lo) ? 0x40 : 0) | noise);
    };
    auto inpt = [&](int i) -> uint8_t {
      int32_t r = analog[i];
      if (r == RES_MIN) return 0x80 | noise;
      if (r == RES_MAX || dump_enabled) return noise;
      double t = (1.6 * r * 0.01E-6);
      uint32_t needed = uint32_t(t * 1.19E6);
      return (uint32_t(cycles) > uint32_t(dump_disabled_cycle + int32_t(needed))) ? (0x80 | noise) : noise;
    };
    switch (addr & 0x0F) {
      case 0x00: return cx(0x0001, 0x0002);
      case 0x01: return cx(0x0004, 0x0008);
      case 0x02: return cx(0x0010, 0x0020);
      case 0x03: return cx(0x0040, 0x0080);
      case 0x04: return cx(0x0100, 0x0200);
      case 0x05: return cx(0x0400, 0x0800);
      case 0x06: return cx(0x1000, 0);
      case 0x07: return cx(0x2000, 0x4000);
      case 0x08: return inpt(0);
      case 0x09: return inpt(1);
      case 0x0A: return inpt(2);
      case 0x0B: return inpt(3);
      case 0x0C: return inpt4_high ? (0x80 | noise) : noise;
      case 0x0D: return inpt5_high ? (0x80 | noise) : noise;
      default: return noise;
    }
  }

  inline void refresh_grp() {
    const Tables& T = tables();
    uint8_t g0 = VDELP0 ? DGRP0 : GRP0; cur_grp0 = REFP0 ? T.reflect[g0] : g0;
    uint8_t g1 = VDELP1 ? DGRP1 : GRP1; cur_grp1 = REFP1 ? T.reflect[g1] : g1;
    if (cur_grp0) enabled |= P0Bit; else enabled &= ~P0Bit;
    if (cur_grp1) enabled |= P1Bit; else enabled &= ~P1Bit;
  }
  inline void refresh_bl() { if (VDELBL ? DENABL : ENABL) enabled |= BLBit; else enabled &= ~BLBit; }

  void tia_poke(uint16_t addr, uint8_t v) {
    const Tables& T = tables();
    addr &= 0x3F;
    int32_t clock = cycles * 3;
    int32_t delay = T.poke_delay[addr];
    if (delay == -1) {
      static const int32_t d[4] = {4, 5, 2, 3};
      int32_t x = (clock - clk_frame_start) % 228;
      delay = d[(x / 3) & 3];
    }
    update_frame(clock + delay);
    if (((clock - clk_frame_start) / 228) > MAX_SCANLINES) { stop = true; partial_frame = false; }
    int32_t hpos = (clock - clk_frame_start) % 228;
    switch (addr) {
      case 0x00:
        VSYNC = v;
        if (VSYNC & 0x02) vsync_finish_clk = clock + 228;
        else if (clock >= vsync_finish_clk) { vsync_finish_clk = 0x7FFFFFFF; stop = true; partial_frame = false; }
        break;
      case 0x01:
        if (!(VBLANK & 0x80) && (v & 0x80)) dump_enabled = true;
        if ((VBLANK & 0x80) && !(v & 0x80)) { dump_enabled = false; dump_disabled_cycle = cycles; }
        VBLANK = v;
        break;
      case 0x02: {   // WSYNC
        int32_t to_eol = 76 - ((cycles - (clk_frame_start / 3)) % 76);
        if (to_eol < 76) cycles += to_eol;
        break;
      }
      case 0x03: {   // RSYNC
        int32_t to_eol = 76 - ((cycles - (clk_frame_start / 3)) % 76);
        cycles += to_eol - 1;
        break;
      }
      case 0x04: NUSIZ0 = v; p0_suppress = 0; break;
      case 0x05: NUSIZ1 = v; p1_suppress = 0; break;
      case 0x06: color[0] = v & 0xFE; break;
      case 0x07: color[1] = v & 0xFE; break;
      case 0x08: color[2] = v & 0xFE; break;
      case 0x09: color[3] = v & 0xFE; break;
      case 0x0A:
        CTRLPF = v; prio_score = uint8_t((v & 0x06) << 5);
        if (hpos < (68 + 79)) pf_reflect_cur = v & 0x01;
        break;
      case 0x0B: REFP0 = (v & 0x08) != 0; refresh_grp(); break;
      case 0x0C: REFP1 = (v & 0x08) != 0; refresh_grp(); break;
      case 0x0D: PF = (PF & 0x000FFFF0) | ((v >> 4) & 0x0F); if (PF) enabled |= PFBit; else enabled &= ~PFBit; break;
      case 0x0E: PF = (PF & 0x000FF00F) | (uint32_t(v) << 4); if (PF) enabled |= PFBit; else enabled &= ~PFBit; break;
      case 0x0F: PF = (PF & 0x00000FFF) | (uint32_t(v) << 12); if (PF) enabled |= PFBit; else enabled &= ~PFBit; break;
      case 0x10: case 0x11: {   // RESP0 / RESP1
        int16_t newx = int16_t(hpos < HBLANK ? 3 : (((hpos - HBLANK) + 5) % 160));
        bool p1 = (addr == 0x11);
        int8_t when = T.reset_when[(p1 ? NUSIZ1 : NUSIZ0) & 7][p1 ? POSP1 : POSP0][newx];
        if (when == 1) update_frame(clock + 11);
        if (p1) { POSP1 = newx; p1_suppress = (when >= 0); } else { POSP0 = newx; p0_suppress = (when >= 0); }
        break;
      }
      case 0x12: POSM0 = int16_t(hpos < HBLANK ? 2 : (((hpos - HBLANK) + 4) % 160)); break;
      case 0x13: POSM1 = int16_t(hpos < HBLANK ? 2 : (((hpos - HBLANK) + 4) % 160)); break;
      case 0x14: POSBL = int16_t(hpos < HBLANK ? 2 : (((hpos - HBLANK) + 4) % 160)); break;
      case 0x1B: GRP0 = v; DGRP1 = GRP1; refresh_grp(); break;
      case 0x1C: GRP1 = v; DGRP0 = GRP0; DENABL = ENABL; refresh_grp(); refresh_bl(); break;
      case 0x1D: ENAM0 = (v & 0x02) != 0; if (ENAM0 && !RESMP0) enabled |= M0Bit; else enabled &= ~M0Bit; break;
      case 0x1E: ENAM1 = (v & 0x02) != 0; if (ENAM1 && !RESMP1) enabled |= M1Bit; else enabled &= ~M1Bit; break;
      case 0x1F: ENABL = (v & 0x02) != 0; refresh_bl(); break;
      case 0x20: HMP0 = v >> 4; break;
      case 0x21: HMP1 = v >> 4; break;
      case 0x22: HMM0 = v >> 4; break;
      case 0x23: HMM1 = v >> 4; break;
      case 0x24: HMBL = v >> 4; break;
      case 0x25: VDELP0 = v & 1; refresh_grp(); break;
      case 0x26: VDELP1 = v & 1; refresh_grp(); break;
      case 0x27: VDELBL = v & 1; refresh_bl(); break;
      case 0x28: case 0x29: {   // RESMP0 / RESMP1
        bool one = (addr == 0x29);
        bool& res = one ? RESMP1 : RESMP0;
        if (res && !(v & 0x02)) {
          uint8_t ns = (one ? NUSIZ1 : NUSIZ0) & 7;
          int16_t middle = (ns == 5) ? 8 : (ns == 7) ? 16 : 4;
          if (one) POSM1 = int16_t((POSP1 + middle) % 160); else POSM0 = int16_t((POSP0 + middle) % 160);
        }
        res = (v & 0x02) != 0;
        if (one) { if (ENAM1 && !RESMP1) enabled |= M1Bit; else enabled &= ~M1Bit; }
        else { if (ENAM0 && !RESMP0) enabled |= M0Bit; else enabled &= ~M0Bit; }
        break;
      }
      case 0x2A: {   // HMOVE
        int32_t x = hpos / 3;
        if (T.hmove_blank[x]) hmove_blank = true;
        auto mv = [&](int16_t& pos, uint8_t hm) {
          pos = int16_t(pos + T.motion[x][hm]);
          if (pos >= 160) pos -= 160; else if (pos < 0) pos += 160;
        };
        mv(POSP0, HMP0); mv(POSP1, HMP1); mv(POSM0, HMM0); mv(POSM1, HMM1); mv(POSBL, HMBL);
        p0_suppress = 0; p1_suppress = 0;
        last_hmove_clk = clock;
        break;
      }
      case 0x2B: HMP0 = HMP1 = HMM0 = HMM1 = HMBL = 0; break;
      case 0x2C: collision = 0; break;
      default: break;   // audio and unused
    }
  }

  // ------------------------------------------------------------ 6502
  inline uint8_t get_ps() const {
    return uint8_t(0x20 | (N ? 0x80 : 0) | (V ? 0x40 : 0) | (B ? 0x10 : 0) | (D ? 0x08 : 0) | (I ? 0x04 : 0) |
                   (notZ ? 0 : 0x02) | (C ? 0x01 : 0));
  }
  inline void set_ps(uint8_t p) {
    N = p & 0x80; V = p & 0x40; B = p & 0x10; D = p & 0x08; I = p & 0x04; notZ = !(p & 0x02); C = p & 0x01;
  }
  inline void nz(uint8_t v) { notZ = v != 0; N = (v & 0x80) != 0; }
  void cpu_reset() {
    A = X = Y = 0; SP = 0xFF; set_ps(0x20); stop = false;
    PC = uint16_t(peek(0xFFFC)); PC |= uint16_t(peek(0xFFFD)) << 8;
  }
  inline void adc(uint8_t m) {
    uint8_t oldA = A;
    if (!D) {
      int sum = int(int8_t(A)) + int(int8_t(m)) + (C ? 1 : 0);
      V = (sum > 127) || (sum < -128);
      sum = int(A) + int(m) + (C ? 1 : 0);
      A = uint8_t(sum); C = sum > 0xFF; nz(A);
    } else {
      int sum = ((A >> 4) * 10 + (A & 15)) + ((m >> 4) * 10 + (m & 15)) + (C ? 1 : 0);
      C = sum > 99;
      int s = sum & 0xFF;
      A = uint8_t((((s % 100) / 10) << 4) | (s % 10));
      nz(A);
      V = ((oldA ^ A) & 0x80) && ((A ^ m) & 0x80);
    }
  }
  inline void sbc(uint8_t m) {
    uint8_t oldA = A;
    if (!D) {
      uint8_t nm = ~m;
      int diff = int(int8_t(A)) + int(int8_t(nm)) + (C ? 1 : 0);
      V = (diff > 127) || (diff < -128);
      diff = int(A) + int(nm) + (C ? 1 : 0);
      A = uint8_t(diff); C = diff > 0xFF; nz(A);
    } else {
      int diff = ((A >> 4) * 10 + (A & 15)) - ((m >> 4) * 10 + (m & 15)) - (C ? 0 : 1);
      if (diff < 0) diff += 100;
      A = uint8_t((((diff % 100) / 10) << 4) | (diff % 10));
      nz(A);
      C = int(oldA) >= (int(m) + (C ? 0 : 1));
      V = ((oldA ^ A) & 0x80) && ((A ^ m) & 0x80);
    }
  }
  inline void cmp(uint8_t r, uint8_t m) {
    uint16_t v = uint16_t(r) - uint16_t(m);
    notZ = (v & 0xFF) != 0; N = (v & 0x80) != 0; C = !(v & 0x0100);
  }
  inline void push(uint8_t v) { poke(0x0100 | SP, v); --SP; }
  inline uint8_t pull() { ++SP; return peek(0x0100 | SP); }

  void step() {
    const Tables& T = tables();
    const Decoded* DT = decode_table();
    uint8_t ir = peek(PC++);
    cycles += T.cycles[ir];
    Decoded d = DT[ir];
    uint16_t ea = 0;
    bool store_like = false;
    switch (d.op) {
      case STA: case STX: case STY: case SAX: case ASL: case LSR: case ROL: case ROR: case INC: case DEC:
      case SLO: case RLA: case SRE: case RRA: case DCP: case ISC: case AHX: case SHX: case SHY: case TAS:
        store_like = true; break;
      default: break;
    }
    // ---- effective address
    switch (d.mode) {
      case IMP: case ACC: break;
      case IMM: ea = PC++; break;
      case ZP: ea = peek(PC++); break;
      case ZPX: ea = uint8_t(peek(PC++) + X); break;
      case ZPY: ea = uint8_t(peek(PC++) + Y); break;
      case ABS: { uint16_t lo = peek(PC++); uint16_t hi = peek(PC++); ea = lo | (hi << 8); break; }
      case ABX: case ABY: {
        uint16_t lo = peek(PC++); uint16_t hi = peek(PC++);
        uint16_t base = lo | (hi << 8);
        ea = uint16_t(base + (d.mode == ABX ? X : Y));
        if (!store_like && ((base ^ ea) & 0xFF00)) cycles += 1;
        break;
      }
      case IZX: {
        uint8_t p = uint8_t(peek(PC++) + X);
        uint16_t lo = peek(p); uint16_t hi = peek(uint8_t(p + 1));
        ea = lo | (hi << 8);
        break;
      }
      case IZY: {
        uint8_t p = peek(PC++);
        uint16_t lo = peek(p); uint16_t hi = peek(uint8_t(p + 1));
        uint16_t base = lo | (hi << 8);
        ea = uint16_t(base + Y);
        if (!store_like && ((base ^ ea) & 0xFF00)) cycles += 1;
        break;
      }
      case REL: ea = PC++; break;
      case IND: {
        uint16_t lo = peek(PC++); uint16_t hi = peek(PC++);
        uint16_t a = lo | (hi << 8);
        uint16_t a2 = ((a & 0xFF) == 0xFF) ? (a & 0xFF00) : uint16_t(a + 1);
        uint16_t tl = peek(a); uint16_t th = peek(a2);
        ea = tl | (th << 8);
        break;
      }
    }
    auto branch = [&](bool take) {
      int8_t off = int8_t(peek(ea));
      if (take) {
        uint16_t target = uint16_t(PC + off);
        cycles += ((PC ^ target) & 0xFF00) ? 2 : 1;
        PC = target;
      }
    };
    uint8_t m;
    switch (d.op) {
      case ADC: adc(peek(ea)); break;
      case SBC: sbc(peek(ea)); break;
      case AND_: A &= peek(ea); nz(A); break;
      case ORA: A |= peek(ea); nz(A); break;
      case EOR: A ^= peek(ea); nz(A); break;
      case CMP: cmp(A, peek(ea)); break;
      case CPX: cmp(X, peek(ea)); break;
      case CPY: cmp(Y, peek(ea)); break;
      case BIT: m = peek(ea); notZ = (A & m) != 0; N = (m & 0x80) != 0; V = (m & 0x40) != 0; break;
      case LDA: A = peek(ea); nz(A); break;
      case LDX: X = peek(ea); nz(X); break;
      case LDY: Y = peek(ea); nz(Y); break;
      case STA: poke(ea, A); break;
      case STX: poke(ea, X); break;
      case STY: poke(ea, Y); break;
      case ASL: case LSR: case ROL: case ROR: {
        m = (d.mode == ACC) ? A : peek(ea);
        bool oldC = C;
        if (d.op == ASL) { C = (m & 0x80) != 0; m = uint8_t(m << 1); }
        else if (d.op == LSR) { C = m & 1; m >>= 1; }
        else if (d.op == ROL) { C = (m & 0x80) != 0; m = uint8_t((m << 1) | (oldC ? 1 : 0)); }
        else { C = m & 1; m = uint8_t((m >> 1) | (oldC ? 0x80 : 0)); }
        nz(m);
        if (d.mode == ACC) A = m; else poke(ea, m);
        break;
      }
      case INC: m = uint8_t(peek(ea) + 1); poke(ea, m); nz(m); break;
      case DEC: m = uint8_t(peek(ea) - 1); poke(ea, m); nz(m); break;
      case INX: ++X; nz(X); break;
      case INY: ++Y; nz(Y); break;
      case DEX: --X; nz(X); break;
      case DEY: --Y; nz(Y); break;
      case TAX: X = A; nz(X); break;
      case TAY: Y = A; nz(Y); break;
      case TXA: A = X; nz(A); break;
      case TYA: A = Y; nz(A); break;
      case TSX: X = SP; nz(X); break;
      case TXS: SP = X; break;
      case CLC: C = false; break;
      case SEC: C = true; break;
      case CLD: D = false; break;
      case SED: D = true; break;
      case CLI: I = false; break;
      case SEI: I = true; break;
      case CLV: V = false; break;
      case PHA: push(A); break;
      case PHP: push(get_ps() | 0x10); break;
      case PLA: A = pull(); nz(A); break;
      case PLP: set_ps(pull()); break;
      case JMP: PC = ea; break;
      case JSR: {
        // ea was fully fetched (PC now past the operand); the pushed address is the last operand byte
        uint16_t ret = uint16_t(PC - 1);
        push(uint8_t(ret >> 8)); push(uint8_t(ret & 0xFF));
        PC = ea;
        break;
      }
      case RTS: { uint16_t lo = pull(); uint16_t hi = pull(); PC = uint16_t((lo | (hi << 8)) + 1); break; }
      case RTI: { set_ps(pull()); uint16_t lo = pull(); uint16_t hi = pull(); PC = lo | (hi << 8); break; }
      case BRK: {
        peek(PC++); B = true;
        push(uint8_t(PC >> 8)); push(uint8_t(PC & 0xFF)); push(get_ps());
        I = true;
        PC = uint16_t(peek(0xFFFE)); PC |= uint16_t(peek(0xFFFF)) << 8;
        break;
      }
      case BPL: branch(!N); break;
      case BMI: branch(N); break;
      case BVC: branch(!V); break;
      case BVS: branch(V); break;
      case BCC: branch(!C); break;
      case BCS: branch(C); break;
      case BNE: branch(notZ); break;
      case BEQ: branch(!notZ); break;
      case NOP: if (d.mode != IMP) peek(ea); break;
      // ---- undocumented
      case LAX: A = X = peek(ea); nz(A); break;
      case LXA: A = X = uint8_t((A | 0xEE) & peek(ea)); nz(A); break;
      case SAX: poke(ea, A & X); break;
      case DCP: m = uint8_t(peek(ea) - 1); poke(ea, m); cmp(A, m); break;
      case ISC: m = uint8_t(peek(ea) + 1); poke(ea, m); sbc(m); break;
      case SLO: m = peek(ea); C = (m & 0x80) != 0; m = uint8_t(m << 1); poke(ea, m); A |= m; nz(A); break;
      case RLA: { m = peek(ea); bool oc = C; C = (m & 0x80) != 0; m = uint8_t((m << 1) | (oc ? 1 : 0)); poke(ea, m); A &= m; nz(A); break; }
      case SRE: m = peek(ea); C = m & 1; m >>= 1; poke(ea, m); A ^= m; nz(A); break;
      case RRA: { m = peek(ea); bool oc = C; C = m & 1; m = uint8_t((m >> 1) | (oc ? 0x80 : 0)); poke(ea, m); adc(m); break; }
      case ANC: A &= peek(ea); nz(A); C = N; break;
      case ALR: A &= peek(ea); C = A & 1; A >>= 1; nz(A); break;
      case ARR: {
        m = peek(ea);
        A &= m; A = uint8_t(((A >> 1) & 0x7F) | (C ? 0x80 : 0));
        C = (A & 0x40) != 0; V = ((A & 0x40) ^ ((A & 0x20) << 1)) != 0; nz(A);
        break;
      }
      case XAA: A = X & peek(ea); nz(A); break;
      case AXS: { uint16_t v = uint16_t(X & A) - uint16_t(peek(ea)); X = uint8_t(v); nz(X); C = !(v & 0x0100); break; }
      case AHX: poke(ea, A & X & uint8_t((ea >> 8) + 1)); break;
      case SHY: poke(ea, Y & uint8_t((ea >> 8) + 1)); break;
      case SHX: poke(ea, X & uint8_t((ea >> 8) + 1)); break;
      case TAS: SP = A & X; poke(ea, SP & uint8_t((ea >> 8) + 1)); break;
      case LAS: A = X = SP = uint8_t(peek(ea) & SP); nz(A); break;
      case KIL: break;
    }
  }

  // One call of the emulated TIA::update(): run the CPU until the frame ends.
  void run_frame() {
    if (!partial_frame) start_frame();
    partial_frame = true;
    stop = false;
    for (int n = 25000; n > 0 && !stop; --n) step();
  }

  void system_reset(uint32_t rnd_for_timer) {
    cycles = 0;
    riot_reset(rnd_for_timer);
    tia_reset();
    cart_reset();
    cpu_reset();
  }
  const uint8_t* screen() const { return fb[cur_fb]; }
};

}  // namespace orc
