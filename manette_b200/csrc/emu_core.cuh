// manette_b200 -- device-side Atari 2600 + ALE environment core (sm_100a).
//
// One lane owns one environment.  This header holds the machine itself; the kernels that
// drive it (lock-step next(), FiGAR repeat loop, reset) live in kernels.cu.
//
// Replaces, for the reference, the external `ALEInterface.act/reset_game/game_over/lives`
// calls made from atari_emulator.py:72-77,94-97,121,128,133 (ALE = Stella 2.x fork).
// Formulation differs from the CPU oracle on purpose:
//  * the 6502 is decoded from a packed 16-bit descriptor per opcode into {address phase, read phase,
//    operate phase, write phase};
//  * TIA register writes are NOT rendered when they happen: the CPU side appends (colour clock,
//    register, value) to a small per-environment FIFO and only keeps the effects the program can
//    observe (WSYNC/RSYNC stalls, VSYNC frame end, VBLANK input-dump latch).  The picture side drains
//    the FIFO later -- all lanes of a warp together -- so the long, branchy rendering path is not
//    entered by one lane at a time while its 31 neighbours wait;
//  * object graphics are built as 32-pixel words of 160-bit line masks (collisions = word ANDs) and
//    pixels leave as 4-byte groups chosen with a byte permute from the packed colour registers;
//  * of the four frames of a next() only the two pooled ones need pixels: the others run with pixel
//    output off (collision latches still exact).  Frame lengths are tracked so that the rare case in
//    which a skipped frame would have left visible bytes (a shorter frame rendered over it, or the
//    ALE terminal freeze) is detected and the unit is re-run with every frame rendered.
//
// The file is also compilable by a host C++ compiler (MN_HD expands to nothing): tests/ build
// it that way ONLY to pre-check the logic against the oracle on machines without a GPU.
// The product never runs the host build.
#pragma once
#include <stdint.h>
#include <stddef.h>
#if !defined(__CUDACC__)
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#endif
#include "cpu_defs.h"


// Out-of-line functions receive the env record, its RAM and its write queue through ordinary pointers; without a hint
// the compiler addresses them generically (64-bit LD / ST through the address-space check) although they always live
// in shared memory, and the frame buffers likewise although they are global.  ~770 such accesses in k_round.
#if defined(__CUDA_ARCH__)
#define MN_IN_SHARED(p) __builtin_assume(__isShared(p))
#define MN_IN_GLOBAL(p) __builtin_assume(__isGlobal(p))
#else
#define MN_IN_SHARED(p) do { } while (0)
#define MN_IN_GLOBAL(p) do { } while (0)
#endif
#define MN_CTX_SPACES(c) do { MN_IN_SHARED((c).s); MN_IN_SHARED((c).ram); MN_IN_SHARED((c).fifo); MN_IN_SHARED((c).rom); MN_IN_GLOBAL((c).fb); } while (0)

// Inlining of the picture-side call tree.  While the picture side ran between the instructions of the 6502 loop its
// functions were kept out of line for the instruction caches' sake (round 1); on its own warp the whole tree inlined
// into picture_process is fastest (B200, 16,384 Ms Pacman envs, k frames/s: none 694, tia_apply 710, + tia_advance 726,
// + player / missile words 736; the words alone 632).  -DMN_PIC*_OUT put single levels back out of line.
#ifdef MN_PICW_OUT
#define MN_PICW_ATTR MN_NOINLINE
#else
#define MN_PICW_ATTR MN_INLINE
#endif
#ifdef MN_PICADV_OUT
#define MN_PICADV_ATTR MN_NOINLINE
#else
#define MN_PICADV_ATTR MN_INLINE
#endif
#ifdef MN_PICAPP_OUT
#define MN_PICAPP_ATTR MN_NOINLINE
#else
#define MN_PICAPP_ATTR MN_INLINE
#endif

namespace mn {

// ------------------------------------------------------------------ constants
enum { CART_2K = 0, CART_4K = 1, CART_F8 = 2, CART_F6 = 3, CART_E0 = 4 };
enum { CTRL_JOYSTICK = 0, CTRL_PADDLES = 1, CTRL_PADDLES_SWAPPED = 2 };
enum { G_GENERIC = 0, G_PONG, G_BREAKOUT, G_SEAQUEST, G_SPACE_INVADERS, G_MS_PACMAN, G_ASTERIX, G_ASTEROIDS,
       G_ENDURO, G_GOPHER, G_GRAVITAR, G_MONTEZUMA, G_YARS, G_NUM_GAMES };

#define MN_SCREEN_W 160
#define MN_SCREEN_H 210
#define MN_FRAME_BYTES (MN_SCREEN_W * MN_SCREEN_H)
#define MN_HBLANK 68
#define MN_YSTART 34
#define MN_MAX_SCANLINES 290
#define MN_NEVER 0x7FFFFFFF
#define MN_RES_MIN 0
#define MN_RES_MAX 0x7FFFFFFF
#define MN_PADDLE_DELTA 23000
#define MN_PADDLE_MIN 27450
#define MN_PADDLE_MAX 790196
#define MN_PADDLE_DEFAULT (((MN_PADDLE_MAX - MN_PADDLE_MIN) / 2) + MN_PADDLE_MIN)

// object ids / enabled bits
enum { OB_P0 = 0, OB_M0 = 1, OB_P1 = 2, OB_M1 = 3, OB_BL = 4 };
enum { EN_P0 = 0x01, EN_M0 = 0x02, EN_P1 = 0x04, EN_M1 = 0x08, EN_BL = 0x10, EN_PF = 0x20 };

// TIA flag bits
enum : uint32_t {
  F_REFP0 = 1u << 0, F_REFP1 = 1u << 1, F_ENAM0 = 1u << 2, F_ENAM1 = 1u << 3, F_ENABL = 1u << 4, F_DENABL = 1u << 5,
  F_VDELP0 = 1u << 6, F_VDELP1 = 1u << 7, F_VDELBL = 1u << 8, F_RESMP0 = 1u << 9, F_RESMP1 = 1u << 10,
  F_SUP0 = 1u << 11, F_SUP1 = 1u << 12, F_PFREFL = 1u << 13, F_HMBLANK = 1u << 14, F_DUMP = 1u << 15,
  F_PARTIAL = 1u << 16, F_CURFB = 1u << 17, F_STOP = 1u << 18, F_INPT4 = 1u << 19, F_INPT5 = 1u << 20,
  F_TIMER_IRQ_READ = 1u << 21, F_TERMINAL = 1u << 22, F_STARTED = 1u << 23,
  F_PIXELS = 1u << 24 /* picture side: the frame being drawn keeps its pixels */,
  F_ANOMALY = 1u << 25 /* picture side: a pixel-less frame would have been visible -> re-run with pixels */ };

// ------------------------------------------------------------------ per-environment record
struct EnvState {
  // 6502
  uint8_t A, X, Y, SP;
  uint16_t PC;
  uint8_t P;            // C(0x01) I(0x04) D(0x08) B(0x10) V(0x40); N/Z live in nz
  uint8_t dbus;
  uint16_t nz;          // Z <=> (nz & 0xFF)==0 ; N <=> nz & 0x180
  uint8_t bank, slice0, slice1, slice2;
  int32_t cycles;
  // RIOT
  uint8_t timer, tshift, ddra, ddrb;
  int32_t timer_set_cycle, irq_reset_cycle;
  uint8_t swcha, swchb, vblank_cpu /* last VBLANK value written (program side) */, pad1;
  int32_t analog[4];
  // TIA, program side
  int32_t clk_frame_start, vsync_finish_clk, dump_disabled_cycle;
  // TIA, picture side (clocks relative to clk_frame_start of the frame being drawn)
  int32_t clk_last_update, clks_to_eol, fb_pos;
  uint16_t pend_len[2];   // per frame buffer: longest pixel-less frame since the last frame drawn with pixels
  uint32_t pf;
  uint32_t flags;       // program side (the 6502 warp): F_DUMP F_PARTIAL F_STOP F_INPT4/5 F_TIMER_IRQ_READ F_TERMINAL F_STARTED
  uint32_t pflags;      // picture side (the TIA partner warp): every other F_* bit.  Two words because the two sides
                        // update their bits concurrently
  uint16_t collision;
  uint8_t vsync, vblank, nusiz0, nusiz1, ctrlpf, enabled;
  uint8_t col[4];       // P0, P1, PF, BK
  uint8_t grp0, grp1, dgrp0, dgrp1, cur_grp0, cur_grp1;
  uint8_t pos[5];       // P0 M0 P1 M1 BL
  uint8_t hm[5];
  // ALE layer
  uint32_t rng[4];
  int32_t left_paddle, right_paddle;
  int32_t score, reward, lives;
  int32_t frame_number, episode_frame_number;
  // AtariEmulator layer
  int32_t host_lives;   // atari_emulator.py:121 self.lives
  uint8_t ring_head;    // ObservationPool.current_observation_index
  uint8_t game, cart, ctrl;
};   // 172 bytes; the 128 bytes of RIOT RAM live beside it (Ctx::ram)

// per-lane working context (pointers to where the pieces live while a kernel runs)
struct Ctx {
  EnvState* s;          // working copy (local / shared memory)
  const uint8_t* rom;   // cartridge image (shared memory)
  uint8_t* ram;         // the 128 bytes of RIOT RAM (shared memory; lanes are MN_RAM_PITCH bytes apart: an odd
                        // number of words, so the lanes of a warp that touch the same byte hit 32 different banks)
  uint8_t* fb;          // this env's two frame buffers (global memory), 2 * MN_FRAME_BYTES
  const Tables* tab;
  uint32_t* fifo;       // this env's TIA write queue (shared memory): two buffers of MN_FIFO_BUF entries + the mailbox
  int fifo_n;           // write POSITION in the buffers: bits 5..4 = buffer being filled, bits 3..0 = entries in it
  uint32_t hseq;        // hand-offs issued so far (its parity = the buffer being filled)
  bool all_pixels;      // draw every frame with pixels (the exact-fallback mode)
  uint64_t obs_lo, obs_hi;   // RAM bytes the last game_observe() looked at (reset memoisation probe)
  bool mbox_timeout;    // a hand-off wait gave up (protocol error: reported, never a hang)
  uint32_t wait_free, wait_done, n_handoff;   // diagnostics: clocks spent waiting for a free buffer / for a blocking hand-off
};
// The picture side runs on a PARTNER WARP (pool.cu: picture_warp): the 6502 warp fills one buffer while the partner
// renders the other.  A hand-off (tia_handoff) publishes the filled buffer through a four-word mailbox per env --
// per LANE, so that a lane inside a divergent slow path (a collision-latch read) can hand off and wait on its own.
#define MN_FIFO_BUF 16    // entries per buffer (a power of two)
#ifndef MN_FIFO_NBUF
#define MN_FIFO_NBUF 8    // buffers per env (a power of two): the 6502 side may run up to NBUF - 1 hand-offs ahead
#endif                    // (measured, Ms Pacman 16,384 envs: 2 / 4 / 8 buffers = 16 / 7.4 / 6.4 % of the lanes' time waiting for one)
#define MN_FIFO_CAP 15    // usable entries: the fill count must stay below MN_FIFO_BUF
#define MN_FIFO_HIGH 12   // a warp hands off when one of its envs has this many pending writes
#define MN_MBOX (MN_FIFO_NBUF * MN_FIFO_BUF)
#define MN_FIFO_WORDS (MN_MBOX + 2 * MN_FIFO_NBUF + 3)   // buffers + per-buffer {request, sync clock} + {hand, done} + 1: an odd stride
enum { MB_HAND = 0, MB_DONE = 1, MB_REQ = 2 /* + 2 * buffer: request, sync clock */ };
static_assert((MN_FIFO_WORDS & 1) == 1, "odd stride");
enum { PIC_DRAIN = 0, PIC_END = 1 /* + close the frame: the unit is over */, PIC_EXIT = 2 /* the partner lane leaves */ };
#define MN_FILL(pos) ((pos) & (MN_FIFO_BUF - 1))
// FIFO entry: [16:0] colour clock since clk_frame_start, [22:17] register, [30:23] value; bit 31 marks the
// start of a new frame (bit 0 then says whether that frame keeps its pixels)
#define MN_FIFO_FRAME 0x80000000u

#define MN_RAM_PITCH 132
MN_HD MN_INLINE uint8_t& ram_at(const Ctx& c, int j) { return c.ram[j]; }

// ------------------------------------------------------------------ small bit helpers
MN_HD MN_INLINE uint32_t brev32(uint32_t v) {
#ifdef __CUDA_ARCH__
  return __brev(v);
#else
  v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
  v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
  v = ((v >> 4) & 0x0F0F0F0Fu) | ((v & 0x0F0F0F0Fu) << 4);
  v = ((v >> 8) & 0x00FF00FFu) | ((v & 0x00FF00FFu) << 8);
  return (v >> 16) | (v << 16);
#endif
}
MN_HD MN_INLINE uint32_t rev8(uint32_t b) { return brev32(b) >> 24; }
// every bit of an 8-bit value -> 2 / 4 adjacent bits
MN_HD MN_INLINE uint32_t widen2(uint32_t b) {
  b = (b | (b << 4)) & 0x0F0Fu; b = (b | (b << 2)) & 0x3333u; b = (b | (b << 1)) & 0x5555u;
  return b * 3u;
}
MN_HD MN_INLINE uint32_t widen4(uint32_t b) {
  b = (b | (b << 12)) & 0x000F000Fu; b = (b | (b << 6)) & 0x03030303u; b = (b | (b << 3)) & 0x11111111u;
  return b * 15u;
}
// bits of a <=32-pixel pattern (bit 0 = leftmost pixel) whose left edge sits at line position
// `start` (0..159, wraps at 160) that fall into the 32-pixel word `w` of the line
MN_HD MN_INLINE uint32_t place(uint32_t pattern, int start, int w) {
  int rel = start - (w << 5);
  if (rel < 0) rel += 160;
  if (rel < 32) return pattern << rel;
  if (rel > 128) return pattern >> (160 - rel);
  return 0u;
}

// ------------------------------------------------------------------ TIA: line words
MN_HD MN_INLINE uint32_t pf_word(const EnvState& s, int w) {
  uint32_t left = (s.pf & 0xFu) | (rev8((s.pf >> 4) & 0xFFu) << 4) | (s.pf & 0xFF000u);
  uint32_t right = (s.pflags & F_PFREFL) ? (brev32(left) >> 12) : left;
  uint32_t cells = (w < 4) ? (((left | (right << 20)) >> (w << 3)) & 0xFFu) : (right >> 12);
  return widen4(cells);
}
MN_HD MN_INLINE uint32_t copies_word(uint32_t pattern, int pos, int mode, bool skip_first, int w) {
  uint32_t m = skip_first ? 0u : place(pattern, pos, w);
  // NUSIZ copy spacing: 1 -> +16 ; 2 -> +32 ; 3 -> +16,+32 ; 4 -> +64 ; 6 -> +32,+64
  int a = -1, b = -1;
  switch (mode) {
    case 1: a = 16; break;
    case 2: a = 32; break;
    case 3: a = 16; b = 32; break;
    case 4: a = 64; break;
    case 6: a = 32; b = 64; break;
    default: break;
  }
  if (a >= 0) { int p = pos + a; if (p >= 160) p -= 160; m |= place(pattern, p, w); }
  if (b >= 0) { int p = pos + b; if (p >= 160) p -= 160; m |= place(pattern, p, w); }
  return m;
}
MN_HD MN_PICW_ATTR uint32_t player_word(uint32_t grp, int nusiz, int pos, bool suppress, int w) {
  int mode = nusiz & 7;
  if (mode == 5 || mode == 7) {        // double / quad sized single copy, drawn one pixel late
    if (suppress) return 0u;
    uint32_t pattern = (mode == 5) ? widen2(rev8(grp)) : widen4(rev8(grp));
    int p = pos + 1; if (p >= 160) p -= 160;
    return place(pattern, p, w);
  }
  return copies_word(rev8(grp), pos, mode, suppress, w);
}
MN_HD MN_PICW_ATTR uint32_t missile_word(int nusiz, int pos, int w) {
  int mode = nusiz & 7;
  uint32_t pattern = (1u << (1 << ((nusiz >> 4) & 3))) - 1u;
  if (mode == 5 || mode == 7) mode = 0;
  return copies_word(pattern, pos, mode, false, w);
}
MN_HD MN_INLINE uint32_t ball_word(const EnvState& s, int w) {
  uint32_t pattern = (1u << (1 << ((s.ctrlpf >> 4) & 3))) - 1u;
  return place(pattern, s.pos[OB_BL], w);
}

// ------------------------------------------------------------------ TIA: picture side
MN_HD MN_INLINE uint32_t perm4(uint32_t x, uint32_t sel) {   // byte i of the result = byte sel[4i+1:4i] of x
#ifdef __CUDA_ARCH__
  return __byte_perm(x, 0u, sel);
#else
  uint32_t r = 0;
  for (int i = 0; i < 4; ++i) r |= ((x >> (8 * ((sel >> (4 * i)) & 3))) & 0xFFu) << (8 * i);
  return r;
#endif
}
// byte i of the result = byte sel[4i+2:4i] of the 8 bytes {x (0..3), y (4..7)}
MN_HD MN_INLINE uint32_t perm8(uint32_t x, uint32_t y, uint32_t sel) {
#ifdef __CUDA_ARCH__
  return __byte_perm(x, y, sel);
#else
  const uint64_t v = uint64_t(x) | (uint64_t(y) << 32);
  uint32_t r = 0;
  for (int i = 0; i < 4; ++i) r |= uint32_t((v >> (8 * ((sel >> (4 * i)) & 7))) & 0xFFu) << (8 * i);
  return r;
#endif
}
MN_HD MN_INLINE uint32_t byte_of(uint32_t x, uint32_t i) {   // byte i (0..3) of x
#ifdef __CUDA_ARCH__
  return __byte_perm(x, 0u, 0x4440u | i);
#else
  return (x >> (8u * i)) & 0xFFu;
#endif
}
// the 8 bits of b -> bit 0 of 8 consecutive nibbles
MN_HD MN_INLINE uint32_t nibbles8(uint32_t b) {
  b = (b | (b << 12)) & 0x000F000Fu; b = (b | (b << 6)) & 0x03030303u; b = (b | (b << 3)) & 0x11111111u;
  return b;
}
// ---- -DMN_CHECK: every frame-buffer address the picture side forms is checked (range, alignment of the word stores)
// and the first violation is recorded for mn_check_report() instead of being left to fault or to scribble.  Built to
// chase the fault that appeared in round 1 whenever fill_px was compiled out of line (tools/gpu_check_build.sh).
#if defined(MN_CHECK) && defined(__CUDACC__)
__device__ unsigned int g_mn_check[8];
MN_HD MN_INLINE void mn_check_fail(unsigned code, unsigned v0, unsigned v1) {
#ifdef __CUDA_ARCH__
  if (atomicCAS(&g_mn_check[0], 0u, code) == 0u) { g_mn_check[1] = v0; g_mn_check[2] = v1; }
#else
  (void)code; (void)v0; (void)v1;
#endif
}
#define MN_ASSERT(cond, code, v0, v1) do { if (!(cond)) mn_check_fail((unsigned)(code), (unsigned)(v0), (unsigned)(v1)); } while (0)
#else
#define MN_ASSERT(cond, code, v0, v1) do { } while (0)
#endif
#ifdef MN_FILL_NOINLINE
#define MN_FILL_ATTR MN_NOINLINE
#else
#define MN_FILL_ATTR MN_INLINE
#endif
// n bytes of `value` at p (any alignment)
MN_HD MN_FILL_ATTR void fill_px(uint8_t* p, int n, uint32_t value) {
  const uint32_t v4 = value * 0x01010101u;
#pragma unroll 1
  while (n > 0 && (reinterpret_cast<uintptr_t>(p) & 3)) { *p++ = uint8_t(value); --n; }
#pragma unroll 1
  for (; n >= 4; n -= 4, p += 4) *reinterpret_cast<uint32_t*>(p) = v4;
#pragma unroll 1
  while (n > 0) { *p++ = uint8_t(value); --n; }
}

// The picture side runs on its own warp: its call tree no longer sits between the instructions of the 6502 loop, so
// the leaves of the hot chain picture_process -> tia_apply -> tia_advance -> tia_render can be inlined again (call
// overhead and the spills around it were a visible share of the partner's time).  -DMN_PIC_OUTLINE keeps them out.
#ifdef MN_PIC_OUTLINE
#define MN_PIC_INLINE MN_NOINLINE
#else
#define MN_PIC_INLINE MN_INLINE
#endif
// render `n` visible pixels of the current line starting at pixel `hpos`
MN_HD MN_PIC_INLINE void tia_render(Ctx& c, int n, int hpos) {
  MN_CTX_SPACES(c);
  EnvState& s = *c.s;
  const bool pixels = (s.pflags & F_PIXELS) != 0;
  uint8_t* out = c.fb + ((s.pflags & F_CURFB) ? MN_FRAME_BYTES : 0) + s.fb_pos;
  MN_ASSERT(n > 0 && hpos >= 0 && hpos + n <= MN_SCREEN_W, 1, n, hpos);
  MN_ASSERT(s.fb_pos >= 0 && s.fb_pos + n <= MN_FRAME_BYTES, 2, s.fb_pos, n);
  MN_ASSERT(((s.fb_pos - hpos) & 3) == 0, 3, s.fb_pos, hpos);        // the 32-bit pixel stores below rely on this
  s.fb_pos += n;
  const uint32_t en = s.enabled;
  if (s.vblank & 0x02) { if (pixels) fill_px(out, n, 0u); return; }
  if (en == 0 || (!pixels && (en & (en - 1)) == 0)) {   // nothing to draw / a lone object cannot collide
    if (pixels) fill_px(out, n, s.col[3]);
    return;
  }
  const int x0 = hpos, x1 = hpos + n;
  const uint32_t colours = uint32_t(s.col[0]) | (uint32_t(s.col[1]) << 8) | (uint32_t(s.col[2]) << 16) | (uint32_t(s.col[3]) << 24);
  const bool prio = (s.ctrlpf & 0x04) != 0, score = (s.ctrlpf & 0x02) != 0;
  uint32_t cx = 0;
  for (int w = x0 >> 5; w <= ((x1 - 1) >> 5); ++w) {
    const int base = w << 5;
    const int lo = (x0 > base) ? (x0 - base) : 0;
    const int hi = (x1 < base + 32) ? (x1 - base) : 32;
    const uint32_t span = ((hi == 32) ? 0xFFFFFFFFu : ((1u << hi) - 1u)) & ~((1u << lo) - 1u);
    const uint32_t pf = (en & EN_PF) ? (pf_word(s, w) & span) : 0u;
    const uint32_t bl = (en & EN_BL) ? (ball_word(s, w) & span) : 0u;
    const uint32_t p0 = (en & EN_P0) ? (player_word(s.cur_grp0, s.nusiz0, s.pos[OB_P0], (s.pflags & F_SUP0) != 0, w) & span) : 0u;
    const uint32_t m0 = (en & EN_M0) ? (missile_word(s.nusiz0, s.pos[OB_M0], w) & span) : 0u;
    const uint32_t p1 = (en & EN_P1) ? (player_word(s.cur_grp1, s.nusiz1, s.pos[OB_P1], (s.pflags & F_SUP1) != 0, w) & span) : 0u;
    const uint32_t m1 = (en & EN_M1) ? (missile_word(s.nusiz1, s.pos[OB_M1], w) & span) : 0u;
    // collision latches: any common pixel inside the span.  Every pair has a player, missile or ball in it, and most
    // 32-pixel words hold none of them (playfield only): skip the fifteen tests there
    if (p0 | m0 | p1 | m1 | bl) {
    cx |= (m0 & p1) ? 0x0001u : 0u; cx |= (m0 & p0) ? 0x0002u : 0u;
    cx |= (m1 & p0) ? 0x0004u : 0u; cx |= (m1 & p1) ? 0x0008u : 0u;
    cx |= (p0 & pf) ? 0x0010u : 0u; cx |= (p0 & bl) ? 0x0020u : 0u;
    cx |= (p1 & pf) ? 0x0040u : 0u; cx |= (p1 & bl) ? 0x0080u : 0u;
    cx |= (m0 & pf) ? 0x0100u : 0u; cx |= (m0 & bl) ? 0x0200u : 0u;
    cx |= (m1 & pf) ? 0x0400u : 0u; cx |= (m1 & bl) ? 0x0800u : 0u;
    cx |= (bl & pf) ? 0x1000u : 0u; cx |= (p0 & p1) ? 0x2000u : 0u; cx |= (m0 & m1) ? 0x4000u : 0u;
    }
    if (!pixels) continue;
    const uint32_t g0 = p0 | m0, g1 = p1 | m1, gf = pf | bl;
    uint8_t* o = out + (base - x0);
    if ((g0 | g1 | gf) == 0u) { fill_px(o + lo, hi - lo, s.col[3]); continue; }
    // colour slot of every pixel as two bit planes: 0 = P0, 1 = P1, 2 = PF, 3 = BK
    uint32_t is0, is1, is2;
    if (prio) { is2 = gf; is0 = g0 & ~gf; is1 = g1 & ~(gf | g0); }
    else {
      is0 = g0; is1 = g1 & ~g0; is2 = gf & ~(g0 | g1);
      if (score) {   // playfield pixels take the player colour of their half of the line
        const uint32_t spf = is2 & pf;
        const uint32_t left = (base + 31 < 80) ? 0xFFFFFFFFu : (base >= 80) ? 0u : 0x0000FFFFu;
        is0 |= spf & left; is1 |= spf & ~left; is2 &= ~spf;
      }
    }
    const uint32_t bk = span & ~(is0 | is1 | is2);
    const uint32_t plane0 = is1 | bk, plane1 = is2 | bk;
#pragma unroll 1
    for (int b = 0; b < 4; ++b) {   // 8 pixels per round
      const uint32_t sp = (span >> (8 * b)) & 0xFFu;
      if (sp == 0u) continue;
      const uint32_t sel = nibbles8((plane0 >> (8 * b)) & 0xFFu) | (nibbles8((plane1 >> (8 * b)) & 0xFFu) << 1);
      const uint32_t lo4 = perm4(colours, sel), hi4 = perm4(colours, sel >> 16);
      uint8_t* q = o + 8 * b;
      MN_ASSERT((reinterpret_cast<uintptr_t>(q) & 3) == 0 && q >= c.fb && q + 8 <= c.fb + 2 * MN_FRAME_BYTES + 32, 4, q - c.fb, sp);
      if (sp == 0xFFu) { reinterpret_cast<uint32_t*>(q)[0] = lo4; reinterpret_cast<uint32_t*>(q)[1] = hi4; }
      else {
        if ((sp & 0x0Fu) == 0x0Fu) reinterpret_cast<uint32_t*>(q)[0] = lo4;
        else { for (int k = 0; k < 4; ++k) if (sp & (1u << k)) q[k] = uint8_t(lo4 >> (8 * k)); }
        if ((sp & 0xF0u) == 0xF0u) reinterpret_cast<uint32_t*>(q)[1] = hi4;
        else { for (int k = 0; k < 4; ++k) if (sp & (16u << k)) q[4 + k] = uint8_t(hi4 >> (8 * k)); }
      }
    }
  }
  s.collision |= uint16_t(cx);
}

// bring the picture up to colour clock `clock` (relative to the start of the frame)
MN_HD MN_PICADV_ATTR void tia_advance(Ctx& c, int32_t clock) {
  MN_CTX_SPACES(c);
  EnvState& s = *c.s;
  const int32_t start = 228 * MN_YSTART;
  const int32_t stop = start + 228 * MN_SCREEN_H;
  if (clock < start || s.clk_last_update >= stop || s.clk_last_update >= clock) return;
  if (clock > stop) clock = stop;
  do {
    int32_t from_sol = 228 - s.clks_to_eol;
    int32_t n;
    if (clock > s.clk_last_update + s.clks_to_eol) { n = s.clks_to_eol; s.clks_to_eol = 228; s.clk_last_update += n; }
    else { n = clock - s.clk_last_update; s.clks_to_eol -= n; s.clk_last_update = clock; }
    if (from_sol < MN_HBLANK) {
      int32_t skip = MN_HBLANK - from_sol; if (skip > n) skip = n;
      from_sol += skip; n -= skip;
    }
    const int32_t old_pos = s.fb_pos;
    if (n != 0) tia_render(c, n, from_sol - MN_HBLANK);
    if ((s.pflags & F_HMBLANK) && from_sol < MN_HBLANK + 8) {
      int32_t blanks = (MN_HBLANK + 8) - from_sol;
      const int32_t room = MN_FRAME_BYTES - old_pos; if (blanks > room) blanks = room;
      MN_ASSERT(old_pos >= 0 && blanks >= 0 && old_pos + blanks <= MN_FRAME_BYTES, 5, old_pos, blanks);
      if (s.pflags & F_PIXELS) fill_px(c.fb + ((s.pflags & F_CURFB) ? MN_FRAME_BYTES : 0) + old_pos, blanks, 0u);
      if (n + from_sol >= MN_HBLANK + 8) s.pflags &= ~F_HMBLANK;
    }
    if (s.clks_to_eol == 228) {   // line finished: playfield mirror latches, first-copy suppression ends
      s.pflags = (s.pflags & ~(F_SUP0 | F_SUP1 | F_PFREFL)) | ((s.ctrlpf & 1) ? F_PFREFL : 0u);
    }
  } while (s.clk_last_update < clock);
}


MN_HD MN_NOINLINE void tia_refresh_grp(EnvState& s) {
  MN_IN_SHARED(&s);
  uint32_t g0 = (s.pflags & F_VDELP0) ? s.dgrp0 : s.grp0;
  uint32_t g1 = (s.pflags & F_VDELP1) ? s.dgrp1 : s.grp1;
  s.cur_grp0 = uint8_t((s.pflags & F_REFP0) ? rev8(g0) : g0);
  s.cur_grp1 = uint8_t((s.pflags & F_REFP1) ? rev8(g1) : g1);
  s.enabled = uint8_t((s.enabled & ~(EN_P0 | EN_P1)) | (s.cur_grp0 ? EN_P0 : 0) | (s.cur_grp1 ? EN_P1 : 0));
}
MN_HD MN_NOINLINE void tia_refresh_misc(EnvState& s) {
  MN_IN_SHARED(&s);
  bool bl = (s.pflags & F_VDELBL) ? (s.pflags & F_DENABL) != 0 : (s.pflags & F_ENABL) != 0;
  bool m0 = (s.pflags & F_ENAM0) && !(s.pflags & F_RESMP0);
  bool m1 = (s.pflags & F_ENAM1) && !(s.pflags & F_RESMP1);
  s.enabled = uint8_t((s.enabled & ~(EN_BL | EN_M0 | EN_M1 | EN_PF)) | (bl ? EN_BL : 0) | (m0 ? EN_M0 : 0) |
                      (m1 ? EN_M1 : 0) | (s.pf ? EN_PF : 0));
}
MN_HD MN_INLINE void set_flag(EnvState& s, uint32_t f, bool on) { s.flags = on ? (s.flags | f) : (s.flags & ~f); }
MN_HD MN_INLINE void set_pflag(EnvState& s, uint32_t f, bool on) { s.pflags = on ? (s.pflags | f) : (s.pflags & ~f); }

// number of the 15 HMOVE extra clocks that still count, as a movement in pixels (+ = right)
MN_HD MN_INLINE int hmove_delta(int cyc, int hm) {
  const int want = hm ^ 8;
  if (cyc <= 22) {
    int fit = (70 - 3 * cyc) >> 2; if (70 - 3 * cyc < 0) fit = 0;
    return 8 - (want < fit ? want : fit);
  }
  if (cyc == 75) return 8 - want;
  int first = (226 - 3 * cyc) >> 2; if (first < 1) first = 1;
  int cnt = want - (first - 1);
  return cnt > 0 ? -cnt : 0;
}
// where a RESPx strobe falls relative to the copies of the player being drawn: 1 inside a copy,
// -1 in the 4-clock start-up of a copy, 0 elsewhere
MN_HD MN_INLINE int resp_zone(int nusiz, int oldx, int newx) {
  const int mode = nusiz & 7;
  const int width = (mode == 5) ? 16 : (mode == 7) ? 32 : 8;
  int res = 0;
  // candidates newx and newx+160 cover the table's 0..236 sweep (later assignments win)
#pragma unroll 1
  for (int k = 0; k < 2; ++k) {
    const int nx = newx + 160 * k;
    if (nx >= 160 + 72 + 5) break;
#pragma unroll 1
    for (int cidx = 0; cidx < 3; ++cidx) {
      int off;
      switch (mode) {
        case 1: off = (cidx == 0) ? 0 : (cidx == 1) ? 16 : -1; break;
        case 2: off = (cidx == 0) ? 0 : (cidx == 1) ? 32 : -1; break;
        case 3: off = cidx * 16; break;
        case 4: off = (cidx == 0) ? 0 : (cidx == 1) ? 64 : -1; break;
        case 6: off = cidx * 32; break;
        default: off = (cidx == 0) ? 0 : -1; break;
      }
      if (off < 0) continue;
      const int d = nx - (oldx + off);
      if (d >= 0 && d < 4) res = -1;
      else if (d >= 4 && d < 4 + width) res = 1;
    }
  }
  return res;
}

static_assert(OB_P0 == 0 && OB_M0 == 1 && OB_P1 == 2 && OB_M1 == 3 && OB_BL == 4, "tia_apply's packed object-index tables");
static_assert(F_REFP1 == F_REFP0 << 1 && F_ENAM1 == F_ENAM0 << 1 && F_ENABL == F_ENAM0 << 2 && F_VDELP1 == F_VDELP0 << 1,
              "tia_apply shifts these flag bits by the register offset");
// ---- picture side: apply one queued register write (colour clock `rel` since the start of the frame)
MN_HD MN_PICAPP_ATTR void tia_apply(Ctx& c, int32_t rel, uint32_t addr, uint32_t v) {
  MN_CTX_SPACES(c);
  EnvState& s = *c.s;
  const int32_t hpos = rel % 228;
  int32_t delay;
  // colour clocks before a write shows: VBLANK/REFPx/GRPx/HMM1../VDELxx/RESMP0 1, NUSIZx/RESMx 8,
  // PFx 2..5 depending on the phase within the playfield cell, everything else immediate
  if (addr >= 0x0D && addr <= 0x0F) delay = 2 + (((hpos / 3) + 2) & 3);
  else if (addr == 0x04 || addr == 0x05 || addr == 0x12 || addr == 0x13) delay = 8;
  else if (addr == 0x01 || addr == 0x0B || addr == 0x0C || addr == 0x1B || addr == 0x1C || (addr >= 0x23 && addr <= 0x28)) delay = 1;
  else delay = 0;
  tia_advance(c, rel + delay);
  switch (addr) {
    case 0x01: s.vblank = uint8_t(v); break;
    case 0x04: s.nusiz0 = uint8_t(v); s.pflags &= ~F_SUP0; break;
    case 0x05: s.nusiz1 = uint8_t(v); s.pflags &= ~F_SUP1; break;
    case 0x06: case 0x07: case 0x08: case 0x09: s.col[addr - 0x06] = uint8_t(v & 0xFE); break;
    case 0x0A:
      s.ctrlpf = uint8_t(v);
      if (hpos < (68 + 79)) set_pflag(s, F_PFREFL, (v & 1) != 0);
      break;
    // (register groups that differ only in a bit position share one body: the 45-case switch was 9 KB of footprint)
    case 0x0B: case 0x0C: set_pflag(s, F_REFP0 << (addr - 0x0B), (v & 0x08) != 0); tia_refresh_grp(s); break;   // REFP0, REFP1
    case 0x0D: case 0x0E: case 0x0F: {   // PF0 (high nibble), PF1, PF2 -> bits 0..3, 4..11, 12..19
      const uint32_t field = (addr == 0x0D) ? 0x0000Fu : (addr == 0x0E) ? 0x00FF0u : 0xFF000u;
      const uint32_t val = (addr == 0x0D) ? ((v >> 4) & 0x0Fu) : (addr == 0x0E) ? (v << 4) : (v << 12);
      s.pf = (s.pf & 0x000FFFFFu & ~field) | val;
      tia_refresh_misc(s);
      break;
    }
    case 0x10: case 0x11: {
      const int p = (addr == 0x10) ? OB_P0 : OB_P1;
      const int newx = (hpos < MN_HBLANK) ? 3 : ((hpos - MN_HBLANK + 5) % 160);
      const int zone = resp_zone((p == OB_P0) ? s.nusiz0 : s.nusiz1, s.pos[p], newx);
      if (zone == 1) tia_advance(c, rel + 11);
      s.pos[p] = uint8_t(newx);
      set_pflag(s, (p == OB_P0) ? F_SUP0 : F_SUP1, zone >= 0);
      break;
    }
    case 0x12: case 0x13: case 0x14:   // RESM0, RESM1, RESBL -> OB_M0 (1), OB_M1 (3), OB_BL (4)
      s.pos[(0x431u >> (4u * (addr - 0x12))) & 7u] = uint8_t((hpos < MN_HBLANK) ? 2 : ((hpos - MN_HBLANK + 4) % 160));
      break;
    case 0x1B: s.grp0 = uint8_t(v); s.dgrp1 = s.grp1; tia_refresh_grp(s); break;
    case 0x1C:
      s.grp1 = uint8_t(v); s.dgrp0 = s.grp0; set_pflag(s, F_DENABL, (s.pflags & F_ENABL) != 0);
      tia_refresh_grp(s); tia_refresh_misc(s);
      break;
    case 0x1D: case 0x1E: case 0x1F: set_pflag(s, F_ENAM0 << (addr - 0x1D), (v & 2) != 0); tia_refresh_misc(s); break;   // ENAM0, ENAM1, ENABL
    case 0x20: case 0x21: case 0x22: case 0x23: case 0x24:   // HMP0, HMP1, HMM0, HMM1, HMBL -> OB_P0 (0), OB_P1 (2), OB_M0 (1), OB_M1 (3), OB_BL (4)
      s.hm[(0x43120u >> (4u * (addr - 0x20))) & 7u] = uint8_t(v >> 4);
      break;
    case 0x25: case 0x26: set_pflag(s, F_VDELP0 << (addr - 0x25), (v & 1) != 0); tia_refresh_grp(s); break;   // VDELP0, VDELP1
    case 0x27: set_pflag(s, F_VDELBL, (v & 1) != 0); tia_refresh_misc(s); break;
    case 0x28: case 0x29: {
      const bool one = (addr == 0x29);
      const uint32_t f = one ? F_RESMP1 : F_RESMP0;
      if ((s.pflags & f) && !(v & 2)) {
        const int ns = (one ? s.nusiz1 : s.nusiz0) & 7;
        const int middle = (ns == 5) ? 8 : (ns == 7) ? 16 : 4;
        s.pos[one ? OB_M1 : OB_M0] = uint8_t((s.pos[one ? OB_P1 : OB_P0] + middle) % 160);
      }
      set_pflag(s, f, (v & 2) != 0);
      tia_refresh_misc(s);
      break;
    }
    case 0x2A: {
      const int cyc = hpos / 3;
      if (cyc <= 20 || cyc == 75) s.pflags |= F_HMBLANK;
#pragma unroll 1
      for (int k = 0; k < 5; ++k) {   // (rolled: five copies of hmove_delta were 2 KB of instruction footprint)
        int p = int(s.pos[k]) + hmove_delta(cyc, s.hm[k]);
        if (p >= 160) p -= 160; else if (p < 0) p += 160;
        s.pos[k] = uint8_t(p);
      }
      s.pflags &= ~(F_SUP0 | F_SUP1);
      break;
    }
    case 0x2B: for (int k = 0; k < 5; ++k) s.hm[k] = 0; break;
    case 0x2C: s.collision = 0; break;
    default: break;
  }
}

// a frame ends on the picture side: account for what it left in its buffer (see F_ANOMALY)
MN_HD MN_INLINE void picture_close_frame(EnvState& s) {
  const int b = (s.pflags & F_CURFB) ? 1 : 0;
  const uint32_t len = uint32_t(s.fb_pos);
  if (s.pflags & F_PIXELS) { if (s.pend_len[b] > len) s.pflags |= F_ANOMALY; s.pend_len[b] = 0; }
  else if (len > s.pend_len[b]) s.pend_len[b] = uint16_t(len);
}
// ... and the next one starts (what the emulated TIA's frame start does to the picture)
MN_HD MN_INLINE void picture_open_frame(EnvState& s, bool pixels) {
  picture_close_frame(s);
  s.pflags ^= F_CURFB;
  s.pflags = pixels ? (s.pflags | F_PIXELS) : (s.pflags & ~F_PIXELS);
  s.clk_last_update = 228 * MN_YSTART;
  s.clks_to_eol = 228;
  s.fb_pos = 0;
}

// ---- picture side proper: what the partner warp does with one hand-off (the host test build runs it inline).
// `cnt` queued writes of buffer `buf` go through the picture; then the picture is brought up to colour clock
// `sync_clk` if one was asked for (collision-latch reads), and the frame is closed if the unit is over.
MN_HD MN_NOINLINE void picture_process(Ctx& c, int buf, int cnt, int32_t sync_clk, int cmd) {
  MN_CTX_SPACES(c);
  const uint32_t* q = c.fifo + buf * MN_FIFO_BUF;
  MN_ASSERT(cnt >= 0 && cnt <= MN_FIFO_CAP && buf >= 0 && buf < MN_FIFO_NBUF, 6, cnt, buf);
  for (int i = 0; i < cnt; ++i) {
    const uint32_t e = q[i];
    if (e & MN_FIFO_FRAME) picture_open_frame(*c.s, (e & 1u) != 0);
    else tia_apply(c, int32_t(e & 0x1FFFFu), (e >> 17) & 0x3Fu, (e >> 23) & 0xFFu);
  }
  if (sync_clk >= 0) tia_advance(c, sync_clk);
  if (cmd == PIC_END) picture_close_frame(*c.s);
}

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t mbox_load(const uint32_t* p) { return *reinterpret_cast<const volatile uint32_t*>(p); }
__device__ __forceinline__ void mbox_store(uint32_t* p, uint32_t v) { *reinterpret_cast<volatile uint32_t*>(p) = v; }
// bounded spin: a protocol error must end the kernel with garbage and an error code, never hang the GPU
__device__ __forceinline__ bool mbox_wait(const uint32_t* p, uint32_t want) {   // until *p has reached `want`
  for (uint32_t spins = 0; spins < (1u << 24); ++spins) {
    if (int32_t(mbox_load(p) - want) >= 0) return true;
    __nanosleep(20);
  }
  return false;
}
#endif
// ---- program side: hand the buffer being filled to the picture side and go on with the other one.
// `blocking`: wait until the picture side is through with it (the caller needs its results).
MN_HD MN_NOINLINE void tia_handoff(Ctx& c, int32_t sync_clk, int cmd, bool blocking) {
  MN_CTX_SPACES(c);
  const int buf = (c.fifo_n >> 4) & (MN_FIFO_NBUF - 1), cnt = MN_FILL(c.fifo_n);
#ifdef __CUDA_ARCH__
  uint32_t* mb = c.fifo + MN_MBOX;
  const uint32_t h = c.hseq;            // this is hand-off number h, of buffer h % NBUF
  mbox_store(mb + MB_REQ + 2 * buf, uint32_t(cnt) | (uint32_t(cmd) << 8));
  mbox_store(mb + MB_REQ + 2 * buf + 1, uint32_t(sync_clk));
  __threadfence_block();
  mbox_store(mb + MB_HAND, h + 1u);
  // the buffer filled next was handed off NBUF - 1 hand-offs ago: it must have been consumed
  const uint32_t t0 = uint32_t(clock64());
  const uint32_t need = blocking ? h + 1u : h + 2u - MN_FIFO_NBUF;
  if (int32_t(need) > 0) { if (c.mbox_timeout || !mbox_wait(mb + MB_DONE, need)) c.mbox_timeout = true; }   // (after one timeout: no more waiting)
  if (blocking) { __threadfence_block(); c.wait_done += uint32_t(clock64()) - t0; }
  else c.wait_free += uint32_t(clock64()) - t0;
  c.n_handoff++;
#else
  (void)blocking;
  picture_process(c, buf, cnt, sync_clk, cmd);
#endif
  c.hseq++;
  c.fifo_n = ((buf + 1) & (MN_FIFO_NBUF - 1)) << 4;
}
// every queued write goes to the picture side (which catches up on its own time)
MN_HD MN_INLINE void tia_drain(Ctx& c) { if (MN_FILL(c.fifo_n) != 0) tia_handoff(c, -1, PIC_DRAIN, false); }
MN_HD MN_INLINE void fifo_push(Ctx& c, uint32_t e) {
  if (MN_FILL(c.fifo_n) >= MN_FIFO_CAP) tia_drain(c);   // safety net; warps hand off together well before this (MN_FIFO_HIGH)
  c.fifo[c.fifo_n++] = e;
}

// ---- program side: a TIA register write.  Only what the 6502 can observe happens now.
MN_HD MN_INLINE void tia_poke(Ctx& c, uint32_t addr, uint32_t v) {
  MN_CTX_SPACES(c);
  EnvState& s = *c.s;
  addr &= 0x3F;
  const int32_t clock = s.cycles * 3;
  const int32_t rel = clock - s.clk_frame_start;
  if ((rel / 228) > MN_MAX_SCANLINES) s.flags = (s.flags | F_STOP) & ~F_PARTIAL;
  if (addr <= 0x03) {
    if (addr == 0x00) {
      s.vsync = uint8_t(v);
      if (v & 0x02) s.vsync_finish_clk = clock + 228;
      else if (clock >= s.vsync_finish_clk) { s.vsync_finish_clk = MN_NEVER; s.flags = (s.flags | F_STOP) & ~F_PARTIAL; }
      // falls through to the FIFO: VSYNC changes nothing in the picture, but the emulated TIA draws up to the clock
      // of EVERY access, and the VSYNC strobe that ends a frame is the last one -- it fixes where the frame's
      // drawing stops (what stays in the buffer below that point is older)
    } else
    if (addr == 0x01) {
      if (!(s.vblank_cpu & 0x80) && (v & 0x80)) s.flags |= F_DUMP;
      if ((s.vblank_cpu & 0x80) && !(v & 0x80)) { s.flags &= ~F_DUMP; s.dump_disabled_cycle = s.cycles; }
      s.vblank_cpu = uint8_t(v);
    } else {
      const int32_t rest = 76 - ((s.cycles - (s.clk_frame_start / 3)) % 76);
      if (addr == 0x02) { if (rest < 76) s.cycles += rest; }
      else s.cycles += rest - 1;
      return;
    }
  } else if (addr > 0x2C || (addr >= 0x15 && addr <= 0x1A)) return;   // audio, unused
  // 17 bits of clock: far beyond the last drawn line only the position within the line still matters
  const int32_t rel17 = (rel < 570 * 228) ? rel : (570 * 228 + rel % 228);
  fifo_push(c, uint32_t(rel17) | (addr << 17) | ((v & 0xFFu) << 23));
}

MN_HD MN_NOINLINE uint32_t tia_peek(Ctx& c, uint32_t addr) {
  MN_CTX_SPACES(c);
  EnvState& s = *c.s;
  const uint32_t noise = s.dbus & 0x3Fu;
  const uint32_t reg = addr & 0x0F;
  if (reg < 8) {
    // collision latches need the picture up to date: hand off what is queued, ask for the picture to be advanced to
    // now, and wait for it
    tia_handoff(c, s.cycles * 3 - s.clk_frame_start, PIC_DRAIN, true);
    // latch pairs in read order: CXM0P CXM1P CXP0FB CXP1FB CXM0FB CXM1FB CXBLPF CXPPMM
    const uint32_t hi = (reg == 6) ? 0x1000u : (reg == 7) ? 0x2000u : (1u << (2 * reg));
    const uint32_t lo = (reg == 6) ? 0u : (reg == 7) ? 0x4000u : (2u << (2 * reg));
    return ((s.collision & hi) ? 0x80u : 0u) | ((s.collision & lo) ? 0x40u : 0u) | noise;
  }
  if (reg < 12) {
    const int32_t r = s.analog[reg - 8];
    if (r == MN_RES_MIN) return 0x80u | noise;
    if (r == MN_RES_MAX || (s.flags & F_DUMP)) return noise;
#ifdef __CUDA_ARCH__
    const double t = __dmul_rn(__dmul_rn(1.6, double(r)), 0.01E-6);
    const uint32_t needed = uint32_t(__dmul_rn(t, 1.19E6));
#else
    const double t = (1.6 * r * 0.01E-6);
    const uint32_t needed = uint32_t(t * 1.19E6);
#endif
    return (uint32_t(s.cycles) > uint32_t(s.dump_disabled_cycle + int32_t(needed))) ? (0x80u | noise) : noise;
  }
  if (reg == 12) return ((s.flags & F_INPT4) ? 0x80u : 0u) | noise;
  if (reg == 13) return ((s.flags & F_INPT5) ? 0x80u : 0u) | noise;
  return noise;
}


// ------------------------------------------------------------------ RIOT
MN_HD MN_NOINLINE uint32_t riot_peek(Ctx& c, uint32_t addr) {
  MN_CTX_SPACES(c);
  EnvState& s = *c.s;
  switch (addr & 7) {
    case 0: return s.swcha;
    case 1: return s.ddra;
    case 2: return s.swchb;
    case 3: return s.ddrb;
    default: break;
  }
  const uint32_t delta = uint32_t((s.cycles - 1) - s.timer_set_cycle);
  int32_t t = int32_t(s.timer) - int32_t(delta >> s.tshift) - 1;
  if (addr & 1) return (t >= 0 || (s.flags & F_TIMER_IRQ_READ)) ? 0x00u : 0x80u;   // interrupt flag
  if (t >= 0) return uint32_t(t) & 0xFFu;
  t = int32_t(uint32_t(s.timer) << s.tshift) - int32_t(delta) - 1;
  if (t <= -2 && !(s.flags & F_TIMER_IRQ_READ)) { s.flags |= F_TIMER_IRQ_READ; s.irq_reset_cycle = s.cycles; }
  if (s.flags & F_TIMER_IRQ_READ) {
    const int32_t offset = s.irq_reset_cycle - (s.timer_set_cycle + int32_t(uint32_t(s.timer) << s.tshift));
    t = int32_t(s.timer) - int32_t(delta >> s.tshift) - offset;
  }
  return uint32_t(t) & 0xFFu;
}
MN_HD MN_NOINLINE void riot_poke(Ctx& c, uint32_t addr, uint32_t v) {
  MN_CTX_SPACES(c);
  EnvState& s = *c.s;
  if ((addr & 7) == 1) s.ddra = uint8_t(v);
  else if ((addr & 7) == 3) s.ddrb = uint8_t(v);
  else if ((addr & 0x14) == 0x14) {
    const uint32_t sel = addr & 3;
    s.timer = uint8_t(v); s.tshift = uint8_t(sel == 0 ? 0 : sel == 1 ? 3 : sel == 2 ? 6 : 10);
    s.timer_set_cycle = s.cycles; s.flags &= ~F_TIMER_IRQ_READ;
  }
}

// ------------------------------------------------------------------ bus
MN_HD MN_NOINLINE void cart_touch(EnvState& s, uint32_t a) {
  MN_IN_SHARED(&s);   // a = addr & 0xFFF, banked carts only
  if (s.cart == CART_F8) { if (a == 0xFF8) s.bank = 0; else if (a == 0xFF9) s.bank = 1; }
  else if (s.cart == CART_F6) { if (a >= 0xFF6 && a <= 0xFF9) s.bank = uint8_t(a - 0xFF6); }
  else if (s.cart == CART_E0) {
    if (a >= 0xFE0 && a <= 0xFF7) {
      const uint8_t sl = uint8_t(a & 7);
      if (a < 0xFE8) s.slice0 = sl; else if (a < 0xFF0) s.slice1 = sl; else s.slice2 = sl;
    }
  }
}

// ---- the memory the hot path reads and writes: cartridge ROM, RIOT RAM, decode tables, TIA write FIFO.
// On the GPU all four live in the block's dynamic shared memory and are addressed by 32-bit offsets into it
// (shared-space loads and stores, no generic 64-bit pointer arithmetic); the host test build uses plain addresses.
#ifdef __CUDACC__
typedef uint32_t maddr;
extern __shared__ __align__(16) uint8_t mn_smem[];
#else
typedef uintptr_t maddr;
#endif
// (device: `maddr` is an address of the .shared window itself -- __cvta_generic_to_shared -- and the accessors are
// ld.shared / st.shared with that register as the address: indexing mn_smem[] instead made the compiler rebuild the
// window base and add it to every offset, ~9 instructions of every tick.  The "memory" clobber keeps their order
// against the ordinary C++ accesses the slow paths make to the same bytes.)
MN_HD MN_INLINE uint32_t m8(maddr a) {
#if defined(__CUDA_ARCH__)
  uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v;
#elif defined(__CUDACC__)
  (void)a; return 0u;   // host pass of nvcc: never executed
#else
  return *reinterpret_cast<const uint8_t*>(a);
#endif
}
MN_HD MN_INLINE void m8w(maddr a, uint32_t v) {
#if defined(__CUDA_ARCH__)
  asm volatile("st.shared.u8 [%0], %1;" :: "r"(a), "r"(v) : "memory");
#elif defined(__CUDACC__)
  (void)a; (void)v;
#else
  *reinterpret_cast<uint8_t*>(a) = uint8_t(v);
#endif
}
MN_HD MN_INLINE void m32w(maddr a, uint32_t v) {
#if defined(__CUDA_ARCH__)
  asm volatile("st.shared.u32 [%0], %1;" :: "r"(a), "r"(v) : "memory");
#elif defined(__CUDACC__)
  (void)a; (void)v;
#else
  *reinterpret_cast<uint32_t*>(a) = v;
#endif
}
MN_HD MN_INLINE TabEnt tab_entry(maddr tab, uint32_t ir) {
#if defined(__CUDA_ARCH__)
  TabEnt t;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(t.k), "=r"(t.d), "=r"(t.x), "=r"(t.dm) : "r"(tab + ir * 16u) : "memory");
  return t;
#elif defined(__CUDACC__)
  (void)tab; (void)ir; TabEnt t; t.k = t.d = t.x = t.dm = 0u; return t;
#else
  return reinterpret_cast<const TabEnt*>(tab)[ir];
#endif
}
MN_HD MN_INLINE FastEnt fast_entry(maddr tab, uint32_t ir) {   // Tables::f follows Tables::e
#if defined(__CUDA_ARCH__)
  const uint32_t q = tab + 4096u + ir * 96u;
  FastEnt t;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(t.sel_a), "=r"(t.sel_b), "=r"(t.sel_idx), "=r"(t.sel_fn) : "r"(q) : "memory");
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4+16];" : "=r"(t.sel_cout), "=r"(t.xm), "=r"(t.binv), "=r"(t.pm) : "r"(q) : "memory");
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4+32];" : "=r"(t.dm), "=r"(t.spd), "=r"(t.cyc), "=r"(t.seqinc) : "r"(q) : "memory");
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4+48];" : "=r"(t.cmask), "=r"(t.cconst), "=r"(t.rotmask), "=r"(t.nzmask) : "r"(q) : "memory");
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4+64];" : "=r"(t.g), "=r"(t.bm_nz), "=r"(t.bm_p), "=r"(t.pclr) : "r"(q) : "memory");
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4+80];" : "=r"(t.pset), "=r"(t.sel_pb), "=r"(t.sel_padd), "=r"(t.penbit) : "r"(q) : "memory");
  return t;
#elif defined(__CUDACC__)
  (void)tab; (void)ir; FastEnt t = {}; return t;
#else
  return reinterpret_cast<const Tables*>(tab)->f[ir];
#endif
}
MN_HD MN_INLINE uint32_t m32(maddr a) {
#if defined(__CUDA_ARCH__)
  uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v;
#elif defined(__CUDACC__)
  (void)a; return 0u;
#else
  return *reinterpret_cast<const uint32_t*>(a);
#endif
}
MN_HD MN_INLINE maddr maddr_of(const void* p) {
#if defined(__CUDA_ARCH__)
  return maddr(__cvta_generic_to_shared(p));
#elif defined(__CUDACC__)
  (void)p; return 0u;
#else
  return reinterpret_cast<maddr>(p);
#endif
}

// The 6502 and what it needs on every instruction, register resident while a frame runs.
struct Cpu {
  uint32_t axys;     // A | X << 8 | Y << 16 | SP << 24: one byte permute picks any of them as an ALU operand
  uint32_t PC;
  uint32_t P;        // C(0x01) I(0x04) D(0x08) B(0x10) V(0x40); N/Z live in nz
  uint32_t nz;       // Z <=> (nz & 0xFF) == 0 ; N <=> nz & 0x180
  uint32_t dbus;     // last value on the data bus (the undriven bits of TIA reads)
  uint32_t segmap;   // cartridge window: byte k = 1K ROM page visible at $1000 + k * $400
  uint32_t romw;     // offset of the window in the ROM image when its four pages are consecutive (every cartridge type but E0)
  uint32_t hot_lo;   // first cartridge offset that may be a bank-switch hot spot ($1000 = none)
  int32_t cycles;
  int32_t clk0;      // EnvState::clk_frame_start (changes only between frames)
  int32_t cyc0;      // clk0 / 3: the CPU cycle the frame's first scan line began at (WSYNC)
  int32_t fifo_n;    // pending TIA writes; Ctx::fifo_n is brought up to date around every out-of-line call
  bool stop;
  // RAM-dependence probe (TRACK instantiations only): which RIOT RAM bytes the program has written since
  // the console reset, and whether it ever read one before writing it
  uint64_t def_lo, def_hi, dep_lo, dep_hi;   // dep: the bytes read before written
  bool tainted;
};
// the read-mostly addresses of the fast paths, passed by value so they stay in registers
struct Mem { maddr rom, ram, tab, fifo, core; };
MN_HD MN_INLINE Mem mem_of(const Ctx& c) {
  Mem m; m.rom = maddr_of(c.rom); m.ram = maddr_of(c.ram); m.tab = maddr_of(c.tab); m.fifo = maddr_of(c.fifo);
  m.core = maddr_of(c.s); return m;
}
MN_HD MN_INLINE uint32_t cpuA(const Cpu& r) { return r.axys & 0xFFu; }
MN_HD MN_INLINE uint32_t cpuX(const Cpu& r) { return (r.axys >> 8) & 0xFFu; }
MN_HD MN_INLINE uint32_t cpuY(const Cpu& r) { return (r.axys >> 16) & 0xFFu; }
MN_HD MN_INLINE uint32_t cpuSP(const Cpu& r) { return r.axys >> 24; }
MN_HD MN_INLINE void setA(Cpu& r, uint32_t v) { r.axys = (r.axys & 0xFFFFFF00u) | (v & 0xFFu); }
MN_HD MN_INLINE void setX(Cpu& r, uint32_t v) { r.axys = (r.axys & 0xFFFF00FFu) | ((v & 0xFFu) << 8); }
MN_HD MN_INLINE void setSP(Cpu& r, uint32_t v) { r.axys = (r.axys & 0x00FFFFFFu) | (v << 24); }
MN_HD MN_NOINLINE uint32_t make_segmap(const EnvState& s) {
  MN_IN_SHARED(&s);
  // 4K: pages 0..3; F8 / F6: the four pages of the selected 4K bank; 2K: its two pages twice; E0: three 1K slices + page 7
  const uint32_t banked = uint32_t(s.bank) * 0x04040404u + 0x03020100u;   // bank is 0 for 4K
  const uint32_t sliced = uint32_t(s.slice0) | (uint32_t(s.slice1) << 8) | (uint32_t(s.slice2) << 16) | (7u << 24);
  return (s.cart == CART_E0) ? sliced : (s.cart == CART_2K) ? 0x01000100u : banked;
}
// first window offset that may be a bank-switch hot spot: F8 $FF8-$FF9, F6 $FF6-$FF9, E0 $FE0-$FF7 (cart_touch); the stubs
// that do the switching sit right below them and run on the fast paths like any other code
#ifdef MN_OLD_HOTLO
MN_HD MN_INLINE uint32_t cart_hot_lo(uint32_t cart) { return cart > CART_4K ? 0xFE0u : 0x1000u; }
#else
MN_HD MN_INLINE uint32_t cart_hot_lo(uint32_t cart) { return cart == CART_F8 ? 0xFF8u : cart == CART_F6 ? 0xFF6u : cart == CART_E0 ? 0xFE0u : 0x1000u; }
#endif
// does not touch fifo_n: that one is carried by the flat loop across frames
MN_HD MN_INLINE void cpu_load(const EnvState& s, Cpu& r) {
  r.axys = uint32_t(s.A) | (uint32_t(s.X) << 8) | (uint32_t(s.Y) << 16) | (uint32_t(s.SP) << 24); r.PC = s.PC; r.P = s.P; r.nz = s.nz; r.dbus = s.dbus;
  r.cycles = s.cycles; r.clk0 = s.clk_frame_start; r.cyc0 = s.clk_frame_start / 3; r.segmap = make_segmap(s); r.romw = (r.segmap & 0xFFu) << 10; r.hot_lo = cart_hot_lo(s.cart);
  r.stop = (s.flags & F_STOP) != 0;
  r.def_lo = r.def_hi = r.dep_lo = r.dep_hi = 0; r.tainted = false;
}
MN_HD MN_INLINE void cpu_store(EnvState& s, const Cpu& r) {
  s.A = uint8_t(r.axys); s.X = uint8_t(r.axys >> 8); s.Y = uint8_t(r.axys >> 16); s.SP = uint8_t(r.axys >> 24); s.PC = uint16_t(r.PC);
  s.P = uint8_t(r.P); s.nz = uint16_t(r.nz); s.dbus = uint8_t(r.dbus); s.cycles = r.cycles;
}

// where a cartridge-window / RIOT RAM address lives
MN_HD MN_INLINE maddr rom_addr(const Mem& mm, uint32_t segmap, uint32_t addr) {
  const uint32_t page = byte_of(segmap, (addr >> 10) & 3u);
  return mm.rom + ((page << 10) | (addr & 0x3FFu));
}
MN_HD MN_INLINE maddr ram_addr(const Mem& mm, uint32_t addr) {
  return mm.ram + (addr & 0x7Fu);
}
// the uncommon reads: bank-switch hot spots (the switch happens before the read), TIA, RIOT
MN_HD MN_NOINLINE uint32_t rd_slow(Ctx& c, uint32_t addr, int32_t cycles, uint32_t dbus) {
  MN_CTX_SPACES(c);
  EnvState& s = *c.s;
  if (addr & 0x1000u) { cart_touch(s, addr & 0xFFFu); return m8(rom_addr(mem_of(c), make_segmap(s), addr)); }
  s.cycles = cycles; s.dbus = uint8_t(dbus);
  return (addr & 0x80u) ? riot_peek(c, addr) : tia_peek(c, addr);
}
template <bool TRACK>
MN_HD MN_INLINE uint32_t rd(Ctx& c, const Mem& mm, Cpu& r, uint32_t addr) {
  const bool rom = (addr & 0x1000u) != 0;
  const bool fast = rom ? ((addr & 0xFFFu) < r.hot_lo) : ((addr & 0x0280u) == 0x0080u);
  uint32_t v;
  if (TRACK && !rom && fast) { const uint32_t j = addr & 0x7Fu; if (!((((j & 64u) ? r.def_hi : r.def_lo) >> (j & 63u)) & 1ull)) { r.tainted = true; if (j & 64u) r.dep_hi |= 1ull << (j & 63u); else r.dep_lo |= 1ull << (j & 63u); } }
  if (fast) v = m8(rom ? rom_addr(mm, r.segmap, addr) : ram_addr(mm, addr));
  else {
    c.fifo_n = r.fifo_n;
    v = rd_slow(c, addr, r.cycles, r.dbus);
    r.fifo_n = c.fifo_n;
    if (rom) { r.segmap = make_segmap(*c.s); r.romw = (r.segmap & 0xFFu) << 10; }
  }
  r.dbus = v;
  return v;
}
// the writes that land neither in RIOT RAM nor (the common case) in the TIA write FIFO;
// returns the CPU cycle count (WSYNC / RSYNC stall the 6502)
MN_HD MN_NOINLINE int32_t wr_slow(Ctx& c, uint32_t addr, uint32_t v, int32_t cycles) {
  MN_CTX_SPACES(c);
  EnvState& s = *c.s;
  if (addr & 0x1000u) { if (s.cart > CART_4K) cart_touch(s, addr & 0xFFFu); return cycles; }
  s.cycles = cycles;
  if (addr & 0x80u) riot_poke(c, addr, v); else tia_poke(c, addr, v);
  return s.cycles;
}
template <bool TRACK>
MN_HD MN_INLINE void wr(Ctx& c, const Mem& mm, Cpu& r, uint32_t addr, uint32_t v) {
  v &= 0xFFu;
  r.dbus = v;
  if ((addr & 0x1280u) == 0x0080u) {
    m8w(ram_addr(mm, addr), v);
    if (TRACK) { const uint32_t j = addr & 0x7Fu; if (j & 64u) r.def_hi |= 1ull << (j & 63u); else r.def_lo |= 1ull << (j & 63u); }
    return;
  }
  // TIA registers $04..$2C before the scan-line overflow point: nothing the 6502 can observe, only a FIFO entry
  // (same arithmetic as tia_poke, which stays the reference for everything else)
  const uint32_t a6 = addr & 0x3Fu;
  const int32_t rel = r.cycles * 3 - r.clk0;
  if (!(addr & 0x1080u) && a6 >= 0x04u && rel < 228 * (MN_MAX_SCANLINES + 1) && MN_FILL(r.fifo_n) < MN_FIFO_CAP) {
    if (a6 <= 0x2Cu && !(a6 >= 0x15u && a6 <= 0x1Au)) { m32w(mm.fifo + uint32_t(r.fifo_n) * 4u, uint32_t(rel) | (a6 << 17) | (v << 23)); r.fifo_n++; }
    return;
  }
  c.fifo_n = r.fifo_n;
  r.cycles = wr_slow(c, addr, v, r.cycles);
  r.fifo_n = c.fifo_n;
  if (addr & 0x1000u) { r.segmap = make_segmap(*c.s); r.romw = (r.segmap & 0xFFu) << 10; }
  else r.stop = (c.s->flags & F_STOP) != 0;
}

// ------------------------------------------------------------------ 6502
MN_HD MN_INLINE uint32_t pack_ps(uint32_t P, uint32_t nz) {
  return 0x20u | (P & 0x5Du) | ((nz & 0x180u) ? 0x80u : 0u) | ((nz & 0xFFu) ? 0u : 0x02u);
}
MN_HD MN_INLINE uint32_t pack_ps(const EnvState& s) { return pack_ps(s.P, s.nz); }
MN_HD MN_INLINE void unpack_ps(Cpu& r, uint32_t p) { r.P = p & 0x5Du; r.nz = ((p & 0x80u) << 1) | ((p & 0x02u) ? 0u : 1u); }
MN_HD MN_INLINE void unpack_ps(EnvState& s, uint32_t p) {
  s.P = uint8_t(p & 0x5D);
  s.nz = uint16_t(((p & 0x80) << 1) | ((p & 0x02) ? 0 : 1));
}
MN_HD MN_INLINE uint32_t bcd_bin(uint32_t v) { return (v >> 4) * 10 + (v & 15); }
MN_HD MN_INLINE void op_adc(Cpu& r, uint32_t m) {
  const uint32_t a = cpuA(r), cin = r.P & 1;
  if (!(r.P & 0x08)) {
    const uint32_t sum = a + m + cin;
    const bool v = ((~(a ^ m)) & (a ^ sum) & 0x80) != 0;
    setA(r, sum); r.nz = sum & 0xFF;
    r.P = (r.P & ~0x41u) | (sum > 0xFF ? 1u : 0u) | (v ? 0x40u : 0u);
  } else {
    const uint32_t sum = bcd_bin(a) + bcd_bin(m) + cin;
    const uint32_t low = sum & 0xFF;
    const uint32_t res = (((low % 100) / 10) << 4) | (low % 10);
    const bool v = ((a ^ res) & 0x80) && ((res ^ m) & 0x80);
    setA(r, res); r.nz = res & 0xFF;
    r.P = (r.P & ~0x41u) | (sum > 99 ? 1u : 0u) | (v ? 0x40u : 0u);
  }
}
MN_HD MN_INLINE void op_sbc(Cpu& r, uint32_t m) {
  const uint32_t a = cpuA(r), cin = r.P & 1;
  if (!(r.P & 0x08)) {
    const uint32_t nm = (~m) & 0xFF;
    const uint32_t sum = a + nm + cin;
    const bool v = ((~(a ^ nm)) & (a ^ sum) & 0x80) != 0;
    setA(r, sum); r.nz = sum & 0xFF;
    r.P = (r.P & ~0x41u) | (sum > 0xFF ? 1u : 0u) | (v ? 0x40u : 0u);
  } else {
    int32_t diff = int32_t(bcd_bin(a)) - int32_t(bcd_bin(m)) - int32_t(1 - cin);
    if (diff < 0) diff += 100;
    // signed arithmetic on purpose: with non-BCD operands diff can still be negative here and the
    // oracle's int expression is the behaviour to match
    const uint32_t res = uint32_t((((diff % 100) / 10) << 4) | (diff % 10)) & 0xFFu;
    const bool carry = a >= (m + (1 - cin));
    const bool v = ((a ^ res) & 0x80) && ((res ^ m) & 0x80);
    setA(r, res); r.nz = res & 0xFF;
    r.P = (r.P & ~0x41u) | (carry ? 1u : 0u) | (v ? 0x40u : 0u);
  }
}
MN_HD MN_INLINE void op_cmp(Cpu& r, uint32_t reg, uint32_t m) {
  const uint32_t d = (reg - m) & 0x1FF;
  r.nz = d & 0xFF;
  r.P = (r.P & ~1u) | ((d & 0x100) ? 0u : 1u);
}
template <bool TRACK>
MN_HD MN_INLINE void stk_push(Ctx& c, const Mem& mm, Cpu& r, uint32_t v) { wr<TRACK>(c, mm, r, 0x0100u | cpuSP(r), v); r.axys -= 0x01000000u; }
template <bool TRACK>
MN_HD MN_INLINE uint32_t stk_pull(Ctx& c, const Mem& mm, Cpu& r) { r.axys += 0x01000000u; return rd<TRACK>(c, mm, r, 0x0100u | cpuSP(r)); }

// The opcodes outside the table-driven datapath (stack / flow / flag ops, BIT, decimal ADC/SBC, undocumented).
// Returns the value of the write phase for the write / read-modify-write classes.
template <bool TRACK>
MN_HD MN_NOINLINE_DEV uint32_t cpu_special(Ctx& c, const Mem& mm, Cpu& r, uint32_t ax, uint32_t op, uint32_t m, uint32_t ea) {
  uint32_t w = 0;
  // Stack traffic of the flow / stack opcodes goes through ONE inlined pull and ONE inlined push (the loop keeps
  // each byte's bus access and their order; what differs per opcode is only how many bytes and what they mean):
  // the 14 separate stk_pull / stk_push expansions this replaces were 4 KB of the flat loop's instruction footprint.
  uint32_t pulled = 0, pushv = 0;
  int npull = 0, npush = 0;
  switch (op) {
    case O_RTS: npull = 2; break;
    case O_RTI: npull = 3; break;
    case O_PLA: case O_PLP: npull = 1; break;
    case O_JSR: { const uint32_t ret = (r.PC - 1) & 0xFFFF; pushv = (ret >> 8) | ((ret & 0xFF) << 8); npush = 2; break; }
    case O_PHA: pushv = cpuA(r); npush = 1; break;
    case O_PHP: pushv = pack_ps(r.P, r.nz) | 0x10; npush = 1; break;
    case O_BRK:
      rd<TRACK>(c, mm, r, r.PC); r.PC = (r.PC + 1) & 0xFFFF; r.P |= 0x10;
      pushv = (r.PC >> 8) | ((r.PC & 0xFF) << 8) | (pack_ps(r.P, r.nz) << 16); npush = 3;
      break;
    default: break;
  }
#pragma unroll 1
  for (int i = 0; i < npull; ++i) pulled |= stk_pull<TRACK>(c, mm, r) << (8 * i);
#pragma unroll 1
  for (int i = 0; i < npush; ++i) stk_push<TRACK>(c, mm, r, (pushv >> (8 * i)) & 0xFFu);
  switch (op) {
    case O_ADC: op_adc(r, m); break;
    case O_SBC: op_sbc(r, m); break;
    case O_BIT: r.nz = ((m & 0x80) << 1) | (cpuA(r) & m); r.P = (r.P & ~0x40u) | (m & 0x40); break;   // (Z from A & M, N = M's bit 7: the fast tick forms the same word)
    case O_LXA: { const uint32_t v = (cpuA(r) | 0xEE) & m; setA(r, v); setX(r, v); r.nz = v; break; }
    case O_ANC: { const uint32_t v = cpuA(r) & m; setA(r, v); r.nz = v; r.P = (r.P & ~1u) | (v >> 7); break; }
    case O_ALR: { uint32_t v = cpuA(r) & m; r.P = (r.P & ~1u) | (v & 1); v >>= 1; setA(r, v); r.nz = v; break; }
    case O_ARR: {
      uint32_t a = cpuA(r) & m; a = ((a >> 1) & 0x7F) | ((r.P & 1) << 7);
      setA(r, a); r.nz = a;
      r.P = (r.P & ~0x41u) | ((a >> 6) & 1) | ((((a >> 6) ^ (a >> 5)) & 1) ? 0x40u : 0u);
      break;
    }
    case O_XAA: { const uint32_t v = cpuX(r) & m; setA(r, v); r.nz = v; break; }
    case O_AXS: { const uint32_t dd = ((cpuX(r) & cpuA(r)) - m) & 0x1FF; setX(r, dd); r.nz = dd & 0xFF; r.P = (r.P & ~1u) | ((dd & 0x100) ? 0u : 1u); break; }
    case O_LAS: { const uint32_t v = m & cpuSP(r); setA(r, v); setX(r, v); setSP(r, v); r.nz = v; break; }
    case O_SAX: w = cpuA(r) & cpuX(r); break;
    case O_AHX: w = cpuA(r) & cpuX(r) & (((ea >> 8) + 1) & 0xFF); break;
    case O_SHY: w = cpuY(r) & (((ea >> 8) + 1) & 0xFF); break;
    case O_SHX: w = cpuX(r) & (((ea >> 8) + 1) & 0xFF); break;
    case O_TAS: { const uint32_t v = cpuA(r) & cpuX(r); setSP(r, v); w = v & (((ea >> 8) + 1) & 0xFF); break; }
    case O_SLO: r.P = (r.P & ~1u) | (m >> 7); w = (m << 1) & 0xFF; { const uint32_t v = cpuA(r) | w; setA(r, v); r.nz = v; } break;
    case O_SRE: r.P = (r.P & ~1u) | (m & 1); w = m >> 1; { const uint32_t v = cpuA(r) ^ w; setA(r, v); r.nz = v; } break;
    case O_RLA: { const uint32_t cin = r.P & 1; r.P = (r.P & ~1u) | (m >> 7); w = ((m << 1) | cin) & 0xFF; const uint32_t v = cpuA(r) & w; setA(r, v); r.nz = v; break; }
    case O_RRA: { const uint32_t cin = r.P & 1; r.P = (r.P & ~1u) | (m & 1); w = (m >> 1) | (cin << 7); op_adc(r, w); break; }
    case O_DCP: w = (m - 1) & 0xFF; op_cmp(r, cpuA(r), w); break;
    case O_ISC: w = (m + 1) & 0xFF; op_sbc(r, w); break;
    case O_JMP: r.PC = ea; break;
    case O_JSR: r.PC = ea; break;
    case O_RTS: r.PC = ((pulled & 0xFFFFu) + 1) & 0xFFFF; break;
    case O_RTI: unpack_ps(r, pulled & 0xFFu); r.PC = (pulled >> 8) & 0xFFFFu; break;
    case O_BRK: {
      r.P |= 0x04;
      const uint32_t lo = rd<TRACK>(c, mm, r, 0xFFFE); r.PC = lo | (rd<TRACK>(c, mm, r, 0xFFFF) << 8);
      break;
    }
    case O_PLA: setA(r, pulled); r.nz = pulled; break;
    case O_PLP: unpack_ps(r, pulled); break;
    case O_FLAG: { const uint32_t mask = 1u << (ax >> 1); r.P = (ax & 1) ? (r.P | mask) : (r.P & ~mask); break; }
    default: break;   // O_KIL, O_NOP
  }
  return w;
}

// Everything of one instruction after the opcode fetch, for a decode entry: the table-driven datapath.  Every
// opcode runs the same instruction sequence, and inside it nothing is chosen by a branch: operands, function and
// carry source are byte-permute selections, results are committed through masks (cpu_defs.h).
template <bool TRACK>
MN_HD MN_INLINE void cpu_exec(Ctx& c, const Mem& mm, Cpu& r, const uint32_t pc, const uint32_t ir, uint32_t b1, uint32_t b2,
                              const bool fast_code, const TabEnt t) {
  const uint32_t k = t.k, d = t.d;
  const uint32_t len1 = (k >> K_LEN) & 3u;   // length - 1
  // the whole base cycle count is charged right after the opcode fetch (operand fetches from the RIOT see it)
  r.cycles += int32_t((d >> 12) & 15u);
  if (!fast_code) {   // operand bytes over the bus, one inlined read (rolled: footprint)
    uint32_t ops = 0;
#pragma unroll 1
    for (uint32_t i = 1; i <= len1; ++i) ops |= rd<TRACK>(c, mm, r, (pc + i) & 0xFFFFu) << (8u * i);
    b1 = (ops >> 8) & 0xFFu; b2 = (ops >> 16) & 0xFFu;
  } else r.dbus = byte_of(ir | (b1 << 8) | (b2 << 16), len1);
  r.PC = (pc + len1 + 1u) & 0xFFFFu;
  // ---- address phase: zero-page / absolute, optionally indexed, from the operand mask of the entry
  const uint32_t xm = t.x & 0xFFFFu;
  const uint32_t idx = perm8(r.axys, 0u, ((k >> K_ISEL) & 7u) | 0x7770u);
  uint32_t base = (b1 | (b2 << 8)) & xm;
  uint32_t ea = (base + idx) & xm;
  if (d & D_INDIRECT) {   // (zp,X)  (zp),Y  (abs)
    const uint32_t mode = d & 15u;
    uint32_t p0, p1;
    if (mode == AM_IND) { p0 = b1 | (b2 << 8); p1 = ((p0 & 0xFF) == 0xFF) ? (p0 & 0xFF00u) : ((p0 + 1) & 0xFFFFu); }
    else { p0 = (mode == AM_IZX) ? ((b1 + cpuX(r)) & 0xFFu) : b1; p1 = (p0 + 1) & 0xFFu; }
    base = 0;
#pragma unroll 1
    for (int i = 0; i < 2; ++i) base |= rd<TRACK>(c, mm, r, i ? p1 : p0) << (8 * i);   // pointer low, then high
    ea = (mode == AM_IZY) ? ((base + cpuY(r)) & 0xFFFFu) : base;
  }
  if ((d & D_PAGEPEN) && ((base ^ ea) & 0xFF00u)) r.cycles += 1;
  // ---- read phase
  uint32_t m = b1;
  if (d & D_READ) m = rd<TRACK>(c, mm, r, ea);
  // ---- operate phase.  The datapath runs for every opcode; what it may change is in the masks of the entry, which
  // are all zero for the opcodes it cannot express (and are ignored for ADC / SBC in decimal mode).
  const bool generic = (k & K_GENERIC) && !((k & K_DECIMAL) && (r.P & 0x08u));
  uint32_t w;
  {
    const uint32_t live = generic ? 0xFFFFFFFFu : 0u;
    const uint32_t s2 = m | 0x00FF0100u;                                   // bytes 4..7 of the operand pool: M, 1, 0xFF, 0
    const uint32_t a = perm8(r.axys, s2, (k & 7u) | 0x7770u);
    const uint32_t b = perm8(r.axys, s2, ((k >> K_BSEL) & 7u) | 0x7770u) ^ ((t.x >> 16) & 0xFFu);
    const uint32_t carry = r.P & 1u;
    const uint32_t cin = ((k >> K_CSEL) & 1u) | ((k >> (K_CSEL + 1)) & carry);
    const uint32_t sum = a + b + cin;                                      // bit 8 = carry out
    const uint32_t rot = (k >> 11) & carry;                                // K_ROT
    const uint32_t left = (a << 1) | rot;                                  // bit 8 = carry out
    const uint32_t right = ((a | (rot << 8)) >> 1) | ((a & 1u) << 8);      // bit 8 = carry out
    const uint32_t lo_sum_or = perm8(sum, a | b, 0x7740u), lo_and_xor = perm8(a & b, a ^ b, 0x7740u);
    const uint32_t fn_pool0 = perm8(lo_sum_or, lo_and_xor, 0x5410u);      // sum, or, and, xor
    const uint32_t fn_pool1 = perm8(left, right, 0x7740u);                // left, right, 0, 0
    const uint32_t res = perm8(fn_pool0, fn_pool1, ((k >> K_FN) & 7u) | 0x7770u);
    const uint32_t c_pool = perm8(perm8(sum, left, 0x7751u), right, 0x7510u);   // carry out of sum, left, right
    const uint32_t cout = perm8(c_pool, 0u, ((k >> K_CSRC) & 3u) | 0x4440u);
    const uint32_t vbit = ((~(a ^ b)) & (a ^ sum) & 0x80u) >> 1;
    const uint32_t pm = (t.x >> 24) & live;                                // bits of P to rewrite
    r.P = (r.P & ~pm) | ((cout | vbit) & pm);
    if (k & live & K_NZ) r.nz = res;
    const uint32_t dm = t.dm & live;
    r.axys = (r.axys & ~dm) | ((res * 0x01010101u) & dm);
    w = res;
  }
  if (!generic) {
    if (d & D_BRANCH) {
      const uint32_t ax = (d >> 16) & 0xFFu;
      // N V C Z as bits 0..3
      const uint32_t fl = ((r.nz & 0x180u) ? 1u : 0u) | ((r.P >> 5) & 2u) | ((r.P & 1u) << 2) | ((r.nz & 0xFFu) ? 0u : 8u);
      if (((fl >> (ax >> 6)) & 1u) == (ax & 1u)) {
        const uint32_t target = (r.PC + uint32_t(int32_t(int8_t(b1)))) & 0xFFFFu;
        r.cycles += ((r.PC ^ target) & 0xFF00u) ? 2 : 1;
        r.PC = target;
      }
    } else w = cpu_special<TRACK>(c, mm, r, (d >> 16) & 0xFFu, (d >> 6) & 63u, m, ea);
    // (Everything here stays inline although it makes the loop ~30 KB of code.  Measured on B200: cpu_special out of
    // line, only its stack / undocumented cases out of line, or stack and pointer accesses through out-of-line bus
    // functions all LOSE 5-15 %: call sites in the loop cost more registers / spills than the code costs in fetch.)
  }
  // ---- write phase
  if (d & D_WRITE) wr<TRACK>(c, mm, r, ea, w);
}

// one instruction
template <bool TRACK>
MN_HD MN_INLINE void cpu_step(Ctx& c, const Mem& mm, Cpu& r) {
  const uint32_t pc = r.PC;
  // ---- fetch: code almost always runs from cartridge ROM, away from the bank-switch hot spots and not across
  // a 1K page of the cartridge window (pages need not be contiguous in the ROM image)
  const bool fast_code = (pc & 0x1000u) && ((pc & 0xFFFu) < 0xFDEu) && ((pc & 0x3FFu) < 0x3FEu);
  uint32_t ir, b1 = 0, b2 = 0;
  if (fast_code) { const maddr a = rom_addr(mm, r.segmap, pc); ir = m8(a); b1 = m8(a + 1); b2 = m8(a + 2); }
  else ir = rd<TRACK>(c, mm, r, pc);
  // (A switch over per-opcode bodies folded from this same source was measured: 2-3x slower. Even with lanes
  // kept together in time, its ~600 KB of code thrashes the instruction cache of an SM that runs 4 warps.)
  cpu_exec<TRACK>(c, mm, r, pc, ir, b1, b2, fast_code, tab_entry(mm.tab, ir));
}

// ------------------------------------------------------------------ the fast tick
// One 6502 instruction WITHOUT A BRANCH IN IT.  ncu on the general path above (profiles/r1_k_round_*): with one warp
// per SM sub-partition every conditional branch of the instruction stream costs 14-40 cycles (predicate -> branch
// latency, refetch at the target, reconvergence), and cpu_step / cpu_exec carry ~14 of them per instruction even
// when no lane takes any -- 930 of the 2,370 cycles of an average tick.  cpu_fast computes the whole instruction
// speculatively, in straight-line code, for the cases tools/warp_sim.cpp found to matter (code in cartridge ROM; the
// table-driven datapath; branches; flag ops; JMP; BIT; (zp,X) / (zp),Y through pointers in RIOT RAM; JSR / RTS / PHA /
// PLA / PHP / PLP on a stack in RIOT RAM; operands in ROM or RAM; RIOT timer reads that have not underflowed; writes to
// RAM, to the TIA write FIFO, and the WSYNC strobe), decides at the end whether every assumption held, and only then
// commits.  Loads are always issued (their addresses are safe whatever the lane's state), stores are predicated.
// A lane for which an assumption failed changes nothing and takes cpu_step: the general path is the definition, and
// tests/ check on the host build that both agree on every instruction of every game (he_set_fast_mode(2)).
// FLAT: the cartridge window is 4 KB of consecutive ROM (2K images are staged twice, 4K / F8 / F6 banks are contiguous):
// its address is one add.  E0 cartridges (three switchable 1K slices) go through the page map.
template <bool FLAT>
MN_HD MN_INLINE maddr fast_rom_addr(const Mem& mm, const Cpu& r, uint32_t addr) {
  return FLAT ? (mm.rom + r.romw + (addr & 0xFFFu)) : rom_addr(mm, r.segmap, addr);
}
#ifdef MN_OLD_HOTLO
#define MN_CODE_HI(r) 0xFE0u
#else
#define MN_CODE_HI(r) (r).hot_lo
#endif
template <bool TRACK, bool FLAT>
MN_HD MN_INLINE bool cpu_fast(const Mem& mm, Cpu& r, const bool go) {
  if (TRACK) return false;   // the RAM-dependence probe instruments the general path only
  // (conditions are 0 / 1 words combined with & and |: `&&` / `||` invite the compiler to branch; every field of the
  // decode entry arrives in the form it is used in, see FastEnt)
  const uint32_t pc = r.PC;
  // (the three bytes must not straddle a 1K page when pages need not be consecutive)
  const uint32_t code_ok = uint32_t((pc & 0x1000u) != 0u) & uint32_t((pc & 0xFFFu) + 2u < MN_CODE_HI(r)) & (FLAT ? 1u : uint32_t((pc & 0x3FFu) < 0x3FEu));
  const maddr ca = fast_rom_addr<FLAT>(mm, r, pc);
  const uint32_t ir = m8(ca), b1 = m8(ca + 1u), b2 = m8(ca + 2u);
  const FastEnt t = fast_entry(mm.tab, ir);
  const uint32_t g = t.g;
  int32_t cyc = r.cycles + int32_t(t.cyc);
  const uint32_t seq = (pc + t.seqinc) & 0xFFFFu;
  const uint32_t sp = r.axys >> 24;
  // ---- a byte pair from RIOT RAM: the pointer of (zp,X) / (zp),Y, or the return address RTS pulls; PLA reads the
  // first byte of the same address.  address = {operand | SP} + {0 | X | 1}, both picked by table selectors
  const uint32_t p0 = (perm8(r.axys, b1, t.sel_pb) + perm8(r.axys, 0x00000100u, t.sel_padd)) & 0xFFu;
  const uint32_t p1 = (p0 + 1u) & 0xFFu;
  const uint32_t pair_ok = uint32_t((g & G_NEEDPAIR) == 0u) | ((p0 & p1) >> 7);
  const uint32_t phi = m8(ram_addr(mm, p1));
  const uint32_t pair = m8(ram_addr(mm, p0)) | (phi << 8);
  // ---- address phase
  const uint32_t idx = perm8(r.axys, 0u, t.sel_idx);
  const uint32_t base = (g & G_PTR) ? pair : ((b1 | (b2 << 8)) & t.xm);
  const uint32_t ea = (base + idx) & t.xm;
  cyc += int32_t(((base ^ ea) & t.penbit & 0xFF00u) != 0u);
  // ---- read phase: cartridge ROM away from the hot spots, RIOT RAM, or the RIOT timer before it underflows
  const uint32_t ra = (g & G_RA_P0) ? (0x100u | p0) : ea;
  const bool r_rom = (ra & 0x1000u) != 0u;
  const uint32_t r_mem = r_rom ? uint32_t((ra & 0xFFFu) < r.hot_lo) : uint32_t((ra & 0x0280u) == 0x0080u);
  const bool r_tim = (ra & 0x1285u) == 0x0284u;
  const uint32_t mv = m8(r_rom ? fast_rom_addr<FLAT>(mm, r, ra) : ram_addr(mm, ra));
  const uint32_t tw = m32(mm.core + uint32_t(offsetof(EnvState, timer)));   // timer | tshift << 8 (riot_peek)
  const int32_t tsc = int32_t(m32(mm.core + uint32_t(offsetof(EnvState, timer_set_cycle))));
  const uint32_t delta = uint32_t((cyc - 1) - tsc);
  const int32_t tv = int32_t(tw & 0xFFu) - int32_t(delta >> ((tw >> 8) & 0xFFu)) - 1;
  const bool has_read = (g & G_READ) != 0u;
  const uint32_t r_ok = uint32_t(!has_read) | r_mem | (uint32_t(r_tim) & uint32_t(tv >= 0));
  const uint32_t m = has_read ? (r_tim ? (uint32_t(tv) & 0xFFu) : mv) : b1;
  // ---- operate phase: the datapath of cpu_exec (nothing is masked off here: the entries of opcodes it does not
  // serve have empty commit masks)
  const uint32_t s2 = m | 0x00FF0100u;
  const uint32_t a = perm8(r.axys, s2, t.sel_a);
  const uint32_t b = perm8(r.axys, s2, t.sel_b) ^ t.binv;
  const uint32_t cin = (r.P & t.cmask) | t.cconst;
  const uint32_t sum = a + b + cin;
  const uint32_t rot = r.P & t.rotmask;
  const uint32_t left = (a << 1) | rot;
  const uint32_t right = ((a | (rot << 8)) >> 1) | ((a & 1u) << 8);
  const uint32_t lo_sum_or = perm8(sum, a | b, 0x7740u), lo_and_xor = perm8(a & b, a ^ b, 0x7740u);
  const uint32_t fn_pool0 = perm8(lo_sum_or, lo_and_xor, 0x5410u);
  const uint32_t fn_pool1 = perm8(left, right, 0x7740u);
  const uint32_t res = perm8(fn_pool0, fn_pool1, t.sel_fn);
  const uint32_t c_pool = perm8(perm8(sum, left, 0x7751u), right, 0x7510u);
  const uint32_t cout = perm8(c_pool, 0u, t.sel_cout);
  const uint32_t vbit = ((~(a ^ b)) & (a ^ sum) & 0x80u) >> 1;
  uint32_t P = (r.P & ~t.pm) | ((cout | vbit) & t.pm);
  uint32_t nz = (res & t.nzmask) | (r.nz & ~t.nzmask);
  // BIT (the datapath computed A & M for Z): N and V are the operand's bits 7 and 6 (penbit = 0xC0 for BIT, else 0); flag ops
  nz |= (m & t.penbit & 0x80u) << 1;
  const uint32_t vsel = t.penbit & 0x40u;
  P = (P & ~vsel) | (m & vsel);
  P = (P & ~t.pclr) | t.pset;
  const uint32_t axys = ((r.axys & ~t.dm) | ((res * 0x01010101u) & t.dm)) + t.spd;
  // ---- branches: taken <=> ((nz & mask) | (P & mask)) != 0, possibly inverted (Z: nz[7:0], N: nz[8:7], C, V in P)
  const uint32_t taken = uint32_t(((r.nz & t.bm_nz) | (r.P & t.bm_p)) != 0u) ^ uint32_t((g & G_INV) != 0u);
  const uint32_t target = (seq + uint32_t(int32_t(int8_t(b1)))) & 0xFFFFu;
  // ---- write phase: RIOT RAM, the TIA write queue (registers $04-$2C before the scan-line overflow point), WSYNC
  const uint32_t wa = (g & G_PUSH) ? (0x100u | sp) : ea;
  const uint32_t ret = (seq - 1u) & 0xFFFFu;
  const uint32_t wv = (g & G_W_RET) ? (ret >> 8) : (res & 0xFFu);
  const bool has_write = (g & G_WRITE) != 0u;
  const uint32_t w_ram = uint32_t((wa & 0x1280u) == 0x0080u);
  const uint32_t w_tia = uint32_t((wa & 0x1080u) == 0u);
  const uint32_t a6 = wa & 0x3Fu;
  const int32_t rel = cyc * 3 - r.clk0;
  const uint32_t rel_ok = uint32_t(rel < 228 * (MN_MAX_SCANLINES + 1));
  const uint32_t w_fifo = w_tia & uint32_t(a6 >= 0x04u) & rel_ok & uint32_t(MN_FILL(r.fifo_n) < MN_FIFO_CAP);
  const uint32_t w_sync = w_tia & uint32_t(a6 == 0x02u) & rel_ok;
  const uint32_t w_ok = uint32_t(!has_write) | w_ram | w_fifo | w_sync;
  const uint32_t w2_ok = uint32_t((g & G_PUSH2) == 0u) | ((sp - 1u) >> 7 & 1u);
  const int32_t rest = 76 - ((cyc - r.cyc0) % 76);   // tia_poke: WSYNC holds the 6502 until the end of the scan line
  // ---- all assumptions held?
  const uint32_t dec_ok = uint32_t((g & G_DECIMAL) == 0u) | uint32_t((r.P & 0x08u) == 0u);
  const bool fast = (uint32_t(go) & code_ok & g & dec_ok & pair_ok & r_ok & w_ok & w2_ok & 1u) != 0u;   // (g & 1: G_VALID)
  // ---- commit (the stores are predicated; the rest is register moves)
  const bool do_write = fast & has_write;
  if (do_write & (w_ram != 0u)) m8w(ram_addr(mm, wa), wv);
  if (do_write & ((g & G_PUSH2) != 0u)) m8w(ram_addr(mm, 0x100u | ((sp - 1u) & 0xFFu)), ret & 0xFFu);
  const bool queued = do_write & (w_ram == 0u) & (w_fifo != 0u) & (a6 <= 0x2Cu) & !((a6 >= 0x15u) & (a6 <= 0x1Au));
  if (queued) m32w(mm.fifo + uint32_t(r.fifo_n) * 4u, uint32_t(rel) | (a6 << 17) | (wv << 23));
  if (fast) {
    uint32_t dbus = byte_of(ir | (b1 << 8) | (b2 << 16), t.seqinc - 1u);
    dbus = (g & G_NEEDPAIR) ? phi : dbus;
    dbus = has_read ? m : dbus;
    dbus = has_write ? ((g & G_PUSH2) ? (ret & 0xFFu) : wv) : dbus;
    cyc += (has_write & (w_ram == 0u) & (w_sync != 0u) & (rest < 76)) ? rest : 0;
    cyc += taken ? (((seq ^ target) & 0xFF00u) ? 2 : 1) : 0;
    uint32_t npc = taken ? target : seq;
    npc = (g & G_PC_EA) ? ea : npc;
    npc = (g & G_PC_PAIR) ? ((pair + 1u) & 0xFFFFu) : npc;
    r.PC = npc; r.cycles = cyc; r.P = P; r.nz = nz; r.axys = axys; r.dbus = dbus;
    r.fifo_n += queued ? 1 : 0;
  }
  return fast;
}


// ------------------------------------------------------------------ frame
// program side of the emulated TIA's frame start: rebase every cycle-stamped quantity, tell the picture
MN_HD MN_INLINE void frame_begin(Ctx& c, bool pixels) {
  MN_CTX_SPACES(c);
  EnvState& s = *c.s;
  const int32_t clocks = ((s.cycles * 3) - s.clk_frame_start) % 228;
  const int32_t cy = s.cycles;
  s.timer_set_cycle -= cy; s.irq_reset_cycle -= cy; s.dump_disabled_cycle -= cy;
  if (s.vsync_finish_clk != MN_NEVER) s.vsync_finish_clk -= cy * 3;
  s.cycles = 0;
  s.clk_frame_start = -clocks;
  fifo_push(c, MN_FIFO_FRAME | (pixels ? 1u : 0u));
}


// ------------------------------------------------------------------ ALE layer
MN_HD MN_INLINE void rng_advance(uint32_t* r) {
  uint32_t y = r[3];
  uint32_t x = (r[0] & 0x7FFFFFFFu) ^ r[1] ^ r[2];
  x ^= (x << 1);
  y ^= (y >> 1) ^ x;
  r[0] = r[1]; r[1] = r[2]; r[2] = x ^ (y << 10); r[3] = y;
  if (y & 1) { r[1] ^= 0x8f7011eeu; r[2] ^= 0xfc78ff1fu; }
}
MN_HD MN_INLINE uint32_t rng_next(uint32_t* r) {
  rng_advance(r);
  uint32_t t0 = r[3];
  const uint32_t t1 = r[0] + (r[2] >> 8);
  t0 ^= t1;
  if (t1 & 1) t0 ^= 0x3793fdffu;
  return t0;
}
MN_HD MN_INLINE void rng_seed(uint32_t* r, uint32_t seed) {
  r[0] = seed; r[1] = 0x8f7011eeu; r[2] = 0xfc78ff1fu; r[3] = 0x3793fdffu;
  for (uint32_t i = 1; i < 8; ++i) r[i & 3] ^= i + 1812433253u * (r[(i - 1) & 3] ^ (r[(i - 1) & 3] >> 30));
  if ((r[0] & 0x7FFFFFFFu) == 0 && r[1] == 0 && r[2] == 0 && r[3] == 0) { r[0] = 'T'; r[1] = 'I'; r[2] = 'N'; r[3] = 'Y'; }
  for (int i = 0; i < 8; ++i) rng_advance(r);
}

MN_HD MN_INLINE uint32_t ram_seen(Ctx& c, int off) {
  const uint32_t j = uint32_t(off) & 0x7Fu;
  if (j & 64u) c.obs_hi |= 1ull << (j & 63u); else c.obs_lo |= 1ull << (j & 63u);
  return ram_at(c, int(j));
}
MN_HD MN_INLINE int32_t ram_bcd(Ctx& c, int off) { const uint32_t b = ram_seen(c, off); return int32_t((b >> 4) * 10 + (b & 15)); }

// per-game reward / terminal / lives from RAM after every frame
MN_HD MN_NOINLINE void game_observe(Ctx& c) {
  MN_CTX_SPACES(c);
  EnvState& s = *c.s;
  int32_t sc = s.score;
  bool term = false;
#define RB(o) int32_t(ram_seen(c, (o)))
  c.obs_lo = c.obs_hi = 0;
  switch (s.game) {
    case G_PONG: { const int32_t x = RB(13), y = RB(14); sc = y - x; term = (x == 21 || y == 21); break; }
    case G_BREAKOUT: {
      const int32_t x = RB(77), y = RB(76), b = RB(57);
      sc = (x & 15) + 10 * (x >> 4) + 100 * (y & 15);
      if (!(s.flags & F_STARTED) && b == 5) s.flags |= F_STARTED;
      term = (s.flags & F_STARTED) && b == 0; s.lives = b;
      break;
    }
    case G_SEAQUEST: sc = ram_bcd(c, 0xBA) + 100 * ram_bcd(c, 0xB9) + 10000 * ram_bcd(c, 0xB8); term = RB(0xA3) != 0; s.lives = RB(0xBB) + 1; break;
    case G_SPACE_INVADERS: sc = ram_bcd(c, 0xE8) + 100 * ram_bcd(c, 0xE6); s.lives = RB(0xC9); term = (RB(0x98) & 0x80) || s.lives == 0; break;
    case G_MS_PACMAN: {
      sc = ram_bcd(c, 0xF8) + 100 * ram_bcd(c, 0xF9) + 10000 * ram_bcd(c, 0xFA);
      const int32_t lb = RB(0xFB) & 15; term = (lb == 0 && RB(0xA7) == 0x53); s.lives = (lb & 7) + 1;
      break;
    }
    case G_ASTERIX: {
      sc = ram_bcd(c, 0xE0) + 100 * ram_bcd(c, 0xDF) + 10000 * ram_bcd(c, 0xDE);
      const int32_t lv = RB(0xD3) & 15; term = (RB(0xC7) == 1 && lv == 1); s.lives = lv;
      break;
    }
    case G_ASTEROIDS: sc = (ram_bcd(c, 0xBE) + 100 * ram_bcd(c, 0xBD)) * 10; s.lives = RB(0xBC) >> 4; term = (s.lives == 0); break;
    case G_ENDURO: {
      sc = 0;
      const int32_t level = RB(0xAD);
      if (level != 0) {
        int32_t cars = ram_bcd(c, 0xAB) + 100 * ram_bcd(c, 0xAC);
        cars = ((level == 1) ? 200 : 300) - cars;
        if (level >= 2) sc = 200 + (level - 2) * 300;
        sc += cars;
      }
      term = (RB(0xAF) == 0xFF);
      break;
    }
    case G_GOPHER: {
      sc = ram_bcd(c, 0xB2) + 100 * ram_bcd(c, 0xB1) + 10000 * ram_bcd(c, 0xB0);
      const int32_t cb = RB(0xB4) & 7; term = (cb == 0); s.lives = (cb & 1) + ((cb >> 1) & 1) + (cb >> 2);
      break;
    }
    case G_GRAVITAR: {
      sc = ram_bcd(c, 0x09) + 100 * ram_bcd(c, 0x08) + 10000 * ram_bcd(c, 0x07);
      const int32_t nl = RB(0x84); term = (nl == 0 && RB(0x81) == 1); s.lives = nl + 1;
      break;
    }
    case G_MONTEZUMA: {
      sc = ram_bcd(c, 0x95) + 100 * ram_bcd(c, 0x94) + 10000 * ram_bcd(c, 0x93);
      const int32_t nl = RB(0xBA); term = (nl == 0 && RB(0xFE) == 0x60); s.lives = (nl & 7) + 1;
      break;
    }
    case G_YARS: {
      sc = ram_bcd(c, 0xE2) + 100 * ram_bcd(c, 0xE1) + 10000 * ram_bcd(c, 0xE0);
      const int32_t lb = RB(0x9E) >> 4; term = (lb == 0); s.lives = lb;
      break;
    }
    default: s.reward = 0; set_flag(s, F_TERMINAL, false); return;
  }
#undef RB
  int32_t r = sc - s.score;
  if (s.game == G_SPACE_INVADERS && r < 0) r = (10000 - s.score) + sc;
  if (s.game == G_ASTEROIDS && r < 0) r += 100000;
  s.reward = r; s.score = sc;
  set_flag(s, F_TERMINAL, term);
}

MN_HD MN_INLINE int game_start_lives(int g) {
  switch (g) { case G_BREAKOUT: return 5; case G_SEAQUEST: return 4; case G_SPACE_INVADERS: return 3; case G_MS_PACMAN: return 3;
    case G_ASTERIX: return 3; case G_ASTEROIDS: return 4; case G_GOPHER: return 3; case G_GRAVITAR: return 6;
    case G_MONTEZUMA: return 6; case G_YARS: return 4; default: return 0; }
}
MN_HD MN_INLINE int game_start_actions(int g) {   // all are FIRE
  switch (g) { case G_ASTERIX: return 1; case G_GOPHER: return 1; case G_GRAVITAR: return 16; case G_YARS: return 1; default: return 0; }
}

// ALE action enum -> stick bits: 1 up, 2 down, 4 left, 8 right, 16 fire (actions 0..17), 32 = console reset (40)
MN_HD MN_INLINE uint32_t action_bits(int a) {
  if (a == 40) return 32u;
  if (a < 0 || a > 17) return 0u;
  // per action: nibble-free table packed as 18 x 5 bits
  const uint8_t t[18] = {0, 16, 1, 8, 4, 2, 9, 5, 10, 6, 17, 24, 20, 18, 25, 21, 26, 22};
  return t[a];
}
MN_HD MN_INLINE void latch_inputs(EnvState& s, int action) {
  const uint32_t b = action_bits(action);
  s.swchb = (b & 32) ? 0x3E : 0x3F;
  if (s.ctrl == CTRL_JOYSTICK) {
    s.swcha = uint8_t(0xFF & ~(((b & 1) ? 0x10 : 0) | ((b & 2) ? 0x20 : 0) | ((b & 4) ? 0x40 : 0) | ((b & 8) ? 0x80 : 0)));
    s.flags = (s.flags | F_INPT5 | F_INPT4) & ~((b & 16) ? F_INPT4 : 0u);
    s.analog[0] = s.analog[1] = s.analog[2] = s.analog[3] = MN_RES_MAX;
  } else {
    int32_t p = s.left_paddle + ((b & 8) ? -MN_PADDLE_DELTA : (b & 4) ? MN_PADDLE_DELTA : 0);
    p = p < MN_PADDLE_MIN ? MN_PADDLE_MIN : p > MN_PADDLE_MAX ? MN_PADDLE_MAX : p;
    s.left_paddle = p;
    const bool swap = (s.ctrl == CTRL_PADDLES_SWAPPED);
    s.analog[0] = swap ? s.right_paddle : s.left_paddle;
    s.analog[1] = swap ? s.left_paddle : s.right_paddle;
    s.analog[2] = MN_RES_MIN; s.analog[3] = MN_RES_MIN;
    s.swcha = uint8_t((b & 16) ? (swap ? 0xBF : 0x7F) : 0xFF);
    s.flags |= F_INPT4 | F_INPT5;
  }
}
MN_HD MN_INLINE void console_reset(Ctx& c, uint32_t rnd) {
  MN_CTX_SPACES(c);
  EnvState& s = *c.s;
  s.cycles = 0;
  // RIOT
  s.timer = uint8_t(25 + (rnd % 75)); s.tshift = 6; s.timer_set_cycle = 0; s.irq_reset_cycle = 0; s.ddra = 0; s.ddrb = 0;
  // TIA (the write FIFO is empty here: units always end drained)
  s.clk_frame_start = 0; s.clk_last_update = 0; s.clks_to_eol = 228; s.vsync_finish_clk = MN_NEVER; s.fb_pos = 0;
  s.dump_disabled_cycle = 0; s.pf = 0; s.collision = 0; s.pend_len[0] = s.pend_len[1] = 0;
  s.flags &= (F_TERMINAL | F_STARTED | F_INPT4 | F_INPT5); s.pflags = 0;
  s.vsync = s.vblank = s.vblank_cpu = s.nusiz0 = s.nusiz1 = s.ctrlpf = s.enabled = 0;
  for (int k = 0; k < 4; ++k) s.col[k] = 0;
  s.grp0 = s.grp1 = s.dgrp0 = s.dgrp1 = s.cur_grp0 = s.cur_grp1 = 0;
  for (int k = 0; k < 5; ++k) { s.pos[k] = 0; s.hm[k] = 0; }
  {
    uint32_t* p = reinterpret_cast<uint32_t*>(c.fb);
    for (int i = 0; i < 2 * MN_FRAME_BYTES / 4; ++i) p[i] = 0u;
  }
  // cartridge
  s.bank = (s.cart == CART_F8) ? 1 : 0; s.slice0 = 4; s.slice1 = 5; s.slice2 = 6;
  // CPU
  s.A = s.X = s.Y = 0; s.SP = 0xFF; unpack_ps(s, 0x20);
  {   // reset vector (a hot-spot free cartridge read)
    Cpu r;
    cpu_load(s, r);
    r.fifo_n = c.fifo_n;
    const Mem mm = mem_of(c);
    const uint32_t lo = rd<false>(c, mm, r, 0xFFFC);
    s.PC = uint16_t(lo | (rd<false>(c, mm, r, 0xFFFD) << 8));
    s.dbus = uint8_t(r.dbus);
  }
}

// ------------------------------------------------------------------ units of work
// Everything a kernel launch asks of one environment is a UNIT: a fixed sequence of single-frame jobs,
// advanced one 6502 instruction per tick so that all lanes of a warp stay in one flat loop.
//   U_ACTS   `total` x ALE act(action): two RNG draws, frozen (nothing emulated, reward 0) once the episode is
//            over, else one frame + the per-game RAM scrape.  next() = 4 acts  (atari_emulator.py:90-100)
//   U_RESET  ALE reset_game(): console reset, 60 NOOP frames, 4 frames of the RESET switch, settings reset,
//            the game's start actions; then `noops` act(NOOP) calls          (atari_emulator.py:70-77)
//   U_POWER_ON  ALE construction + loadROM: seed the RNG, fill RIOT RAM with its garbage, then U_RESET with no no-ops
enum { U_ACTS = 0, U_RESET = 1, U_POWER_ON = 2 };
struct Unit {
  int kind, idx, total, action, nstart, budget;
  int32_t reward;
  bool in_frame, frozen_last, job_is_act;
};
// the part of a running unit the flat loop keeps in registers
struct Hot {
  Cpu cpu; int budget; bool in_frame, more;
  int jobs;                       // jobs begun so far (lane time = job, CPU cycle within its frame)
  uint32_t instr;                 // 6502 instructions executed
  uint64_t def_lo, def_hi, dep_lo, dep_hi;   // RAM-dependence probe, carried from frame to frame (see Cpu)
  bool tainted, obs_bad;          // obs_bad: the RAM scrape looked at a byte the program had not written yet
};

MN_HD MN_INLINE void unit_idle(Unit& u) { u.kind = U_ACTS; u.idx = u.total = 0; u.in_frame = false; u.reward = 0; u.frozen_last = false; }

// `seed`: U_POWER_ON: the ALE seed; U_RESET: the RNG draw (already taken) that seeds the RIOT timer
MN_HD MN_INLINE void unit_init(Ctx& c, Unit& u, int kind, int action, int count, uint32_t seed) {
  MN_CTX_SPACES(c);
  EnvState& s = *c.s;
  u.kind = kind; u.idx = 0; u.action = action; u.reward = 0; u.in_frame = false; u.frozen_last = false; u.budget = 0;
  u.job_is_act = false; u.nstart = 0;
  // (the write queue is empty here -- units end drained -- and c.fifo_n already points at the buffer to fill)
  s.pflags &= ~F_ANOMALY;
  if (kind == U_ACTS) { u.total = count; return; }
  if (kind == U_POWER_ON) {
    rng_seed(s.rng, seed);
    for (int i = 0; i < 128; ++i) ram_at(c, i) = uint8_t(rng_next(s.rng));
    s.flags = 0; s.pflags = 0; s.frame_number = 0; s.ring_head = 0; s.pend_len[0] = s.pend_len[1] = 0;
  }
  s.episode_frame_number = 0;
  s.left_paddle = s.right_paddle = MN_PADDLE_DEFAULT;
  console_reset(c, (kind == U_POWER_ON) ? rng_next(s.rng) : seed);
  u.nstart = game_start_actions(s.game);
  u.total = 64 + u.nstart + count;
}

MN_HD MN_NOINLINE void unit_job_done(Ctx& c, Unit& u) {
  MN_CTX_SPACES(c);
  EnvState& s = *c.s;
  u.in_frame = false;
  game_observe(c);
  if (u.job_is_act) { s.frame_number++; s.episode_frame_number++; u.reward += s.reward; }
  if (u.kind != U_ACTS && u.idx == 64) {   // RomSettings::reset() after the RESET-switch frames
    s.score = 0; s.reward = 0; s.flags &= ~(F_TERMINAL | F_STARTED); s.lives = game_start_lives(s.game);
  }
  // AtariEmulator reads ale.lives() right after loadROM / reset_game (atari_emulator.py:31,73)
  if (u.kind != U_ACTS && u.idx == 64 + u.nstart) s.host_lives = s.lives;
}

// start the next job of the unit
MN_HD MN_NOINLINE void unit_job_begin(Ctx& c, Unit& u) {
  MN_CTX_SPACES(c);
  EnvState& s = *c.s;
  {
    int action;
    if (u.kind == U_ACTS) { action = u.action; u.job_is_act = true; }
    else if (u.idx < 60) { action = 0; u.job_is_act = false; }
    else if (u.idx < 64) { action = 40; u.job_is_act = false; }
    else if (u.idx < 64 + u.nstart) { action = 1; u.job_is_act = false; }
    else { action = 0; u.job_is_act = true; }
    const bool pixels = c.all_pixels || (u.idx >= u.total - 2);
    u.idx++;
    if (u.job_is_act) {
      rng_advance(s.rng); rng_advance(s.rng);
      if (u.kind == U_ACTS && u.idx == u.total) u.frozen_last = (s.flags & F_TERMINAL) != 0;
      if (s.flags & F_TERMINAL) return;   // act() of a finished episode: nothing is emulated
    }
    latch_inputs(s, action);
    if (!(s.flags & F_PARTIAL)) frame_begin(c, pixels);
    s.flags = (s.flags | F_PARTIAL) & ~F_STOP;
    u.in_frame = true;
  }
}
MN_HD MN_INLINE void hot_init(const Ctx& c, const Unit& u, Hot& h) {
  h.in_frame = false; h.more = u.idx < u.total; h.budget = 0; h.cpu.stop = false; h.cpu.fifo_n = c.fifo_n; h.instr = 0; h.jobs = 0;
  // cpu_fast issues its loads before it knows whether the lane runs at all: the registers they are addressed from
  // must be valid (any cartridge page, any PC) from the first tick on, not only after the first cpu_load
  h.cpu.axys = 0; h.cpu.PC = 0x1000u; h.cpu.P = 0; h.cpu.nz = 0; h.cpu.dbus = 0; h.cpu.segmap = 0; h.cpu.romw = 0; h.cpu.hot_lo = 0x1000u;
  h.cpu.cycles = 0; h.cpu.clk0 = 0; h.cpu.cyc0 = 0;
  h.cpu.def_lo = h.cpu.def_hi = h.cpu.dep_lo = h.cpu.dep_hi = 0; h.cpu.tainted = false;
  h.def_lo = h.def_hi = h.dep_lo = h.dep_hi = 0; h.tainted = false; h.obs_bad = false;
}
MN_HD MN_INLINE bool hot_has_work(const Hot& h) { return h.in_frame || h.more; }
// Emulated time of a lane, comparable across the lanes of a warp: (job, CPU cycle since the frame of that job began).
// The flat loop only advances the lanes that are not ahead of the slowest one by more than a few cycles: programs
// are scan-line structured (WSYNC), so lanes kept together in time mostly sit at the same program counter.
MN_HD MN_INLINE int32_t hot_time(const Hot& h) { return h.in_frame ? ((h.jobs << 16) + h.cpu.cycles) : ((h.jobs + 1) << 16); }
// one tick: start the next job, or run one instruction of the frame in progress
#if !defined(__CUDACC__)
// host test build only: 0 = general path only, 1 = fast tick first (what the kernels do), 2 = run both on every
// instruction the fast tick accepts and abort on the first difference
static int g_fast_mode = 1;
// host test build: use the flat-window instantiation of the fast tick (what the kernels run for every cartridge type
// but E0); the harness stages 2K images twice, as the kernels do
static int g_fast_flat = 0;
static unsigned long long g_fast_taken = 0, g_fast_refused = 0;
// one instruction the way the kernels run it (mode 1), or with the fast tick checked against the general path (mode 2);
// returns false if the general path has to run it
static inline bool cpu_fast_host(Ctx& c, const Mem& mm, Cpu& r) {
  if (g_fast_mode == 0) return false;
  const bool flat = g_fast_flat != 0 && c.s->cart != CART_E0;
  if (g_fast_mode == 1) return flat ? cpu_fast<false, true>(mm, r, true) : cpu_fast<false, false>(mm, r, true);
  const Cpu before = r;
  const EnvState s_before = *c.s;
  uint8_t ram0[128], ram1[128]; uint32_t fifo0[MN_MBOX], fifo1[MN_MBOX];
  for (int j = 0; j < 128; ++j) ram0[j] = ram_at(c, j);
  for (int j = 0; j < MN_MBOX; ++j) fifo0[j] = c.fifo[j];
  if (!(flat ? cpu_fast<false, true>(mm, r, true) : cpu_fast<false, false>(mm, r, true))) { ++g_fast_refused; return false; }
  const Cpu fast = r;
  for (int j = 0; j < 128; ++j) { ram1[j] = ram_at(c, j); ram_at(c, j) = ram0[j]; }
  for (int j = 0; j < MN_MBOX; ++j) { fifo1[j] = c.fifo[j]; c.fifo[j] = fifo0[j]; }
  r = before;
  cpu_step<false>(c, mm, r);
  const Cpu& g = r;
  bool same = g.axys == fast.axys && g.PC == fast.PC && g.P == fast.P && g.nz == fast.nz && g.dbus == fast.dbus && g.cycles == fast.cycles &&
              g.fifo_n == fast.fifo_n && g.segmap == fast.segmap && g.stop == before.stop;
  for (int j = 0; j < 128; ++j) same = same && ram_at(c, j) == ram1[j];
  for (int j = 0; j < MN_MBOX; ++j) same = same && c.fifo[j] == fifo1[j];
  EnvState sa = *c.s, sb = s_before;
  sa.cycles = sb.cycles = 0; sa.dbus = sb.dbus = 0;   // scratch copies the slow bus functions leave behind
  same = same && memcmp(&sa, &sb, sizeof(EnvState)) == 0;
  if (!same) {
    fprintf(stderr, "cpu_fast != cpu_step at PC %04x: axys %08x/%08x PC %04x/%04x P %02x/%02x nz %03x/%03x dbus %02x/%02x cycles %d/%d fifo %d/%d\n",
            before.PC, fast.axys, g.axys, fast.PC, g.PC, fast.P, g.P, fast.nz, g.nz, fast.dbus, g.dbus, fast.cycles, g.cycles, fast.fifo_n, g.fifo_n);
    abort();
  }
  ++g_fast_taken;
  return true;
}
#endif
template <bool TRACK, bool FLAT = false>
MN_HD MN_INLINE void unit_tick(Ctx& c, const Mem& mm, Unit& u, Hot& h, const bool elig = true) {
#if !defined(__CUDACC__)
  if (!TRACK && elig && h.in_frame && h.budget > 1 && cpu_fast_host(c, mm, h.cpu)) { --h.budget; return; }
#else
  if (cpu_fast<TRACK, FLAT>(mm, h.cpu, elig && h.in_frame && h.budget > 1)) { --h.budget; return; }
#endif
  if (!elig) return;
  if (!h.in_frame) {
    c.fifo_n = h.cpu.fifo_n;
    unit_job_begin(c, u);
    h.cpu.fifo_n = c.fifo_n;
    h.jobs++;
    h.in_frame = u.in_frame; h.more = u.idx < u.total;
    if (h.in_frame) {
      cpu_load(*c.s, h.cpu); h.cpu.stop = false; h.budget = 25000;
      if (TRACK) { h.cpu.def_lo = h.def_lo; h.cpu.def_hi = h.def_hi; h.cpu.dep_lo = h.dep_lo; h.cpu.dep_hi = h.dep_hi; h.cpu.tainted = h.tainted; }
    }
    return;
  }
  cpu_step<TRACK>(c, mm, h.cpu);
  if (h.cpu.stop || --h.budget == 0) {
    h.instr += uint32_t(25000 - h.budget);
    cpu_store(*c.s, h.cpu);
    if (TRACK) {
      h.def_lo = h.cpu.def_lo; h.def_hi = h.cpu.def_hi; h.dep_lo = h.cpu.dep_lo; h.dep_hi = h.cpu.dep_hi; h.tainted = h.cpu.tainted;
    }
    unit_job_done(c, u);
    h.in_frame = false;
    // Scrapes after RomSettings::reset() (job 64 on) shape the episode.  A byte they look at before the program has
    // written it still holds its value from before the reset: the result depends on it exactly as it does on a byte
    // the program itself reads before writing, so it joins the dependence set (the memo key).  Yars' Revenge is the
    // README game this matters for: its lives byte is scraped before the program rewrites it.
    if (TRACK && (u.kind == U_ACTS || u.idx > 64)) {
      const uint64_t lo = c.obs_lo & ~h.def_lo, hi = c.obs_hi & ~h.def_hi;
      if (lo | hi) { h.dep_lo |= lo; h.dep_hi |= hi; h.tainted = true; }
    }
  }
}

// the flat loop's drain: every queued write of this env goes through the picture
MN_HD MN_INLINE void hot_drain(Ctx& c, Hot& h) { c.fifo_n = h.cpu.fifo_n; tia_drain(c); h.cpu.fifo_n = c.fifo_n; }

// after the unit's last tick: flush the picture and report whether the pixel-less frames were harmless
MN_HD MN_INLINE bool unit_finish(Ctx& c, Hot& h) {
  MN_CTX_SPACES(c);
  EnvState& s = *c.s;
  c.fifo_n = h.cpu.fifo_n;
  tia_handoff(c, -1, PIC_END, true);   // the rest of the queue, then the frame is closed; wait for the verdict
  h.cpu.fifo_n = c.fifo_n;
  const bool bad = (s.pflags & F_ANOMALY) || s.pend_len[0] != 0 || s.pend_len[1] != 0;
  s.pflags &= ~F_ANOMALY;
  return bad;
}

}  // namespace mn
