// manette_b200 -- the B200-native environment pool: kernels + C ABI (include/manette_b200.h).
//
// Replaces the reference's multi-process ALE worker pool (runners.py:7-50, emulator_runner.py:19-42,
// atari_emulator.py:17-136) with device-resident emulator state.  One macro step
// (Runners.update_environments) is a short, fixed sequence of launches on the caller's stream:
//
//   k_begin_step     decode (action, repetition) of every env, build the round-0 work list
//   repeat max(tab_rep)+1 times:
//     k_round        one next() (= 4 emulated frames) for every env still on the work list; envs that
//                    still have repetitions left and are not terminal are COMPACTED into the next
//                    round's list, terminal ones into the reset list          (emulator_runner.py:26-40)
//     k_push_frames  K3: max of the two pooled frames -> luminance/RGB -> 84x84 nearest -> ring slot
//   k_round(reset)   reset_game + start no-ops for the envs on the reset list   (atari_emulator.py:70-77)
//   4 x { k_round(initial), k_push_frames }                                     (atari_emulator.py:105-107)
//   k_emit           ring -> stacked NHWC states (uint8x16 stores), rewards, terminals
//
// Work lists are per game (one cartridge image per thread block, staged in shared memory) and are
// spread evenly over a fixed number of warps, so later FiGAR rounds -- fewer live envs -- run with
// fewer envs per warp (less opcode divergence) instead of with idle warps.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>

#include "../../include/manette_b200.h"
#include "atari_env.cuh"
#include "decode_tables.h"
#include "game_db.h"

namespace mn {

#define MN_MAX_GAMES 16
#define MN_WARPS_PER_BLOCK 4                  // 6502 warps per block, one per SM sub-partition ...
#ifndef MN_PARTNERS
#define MN_PARTNERS 1                         // ... each with this many picture-side partner warps (warps w + 4, w + 8 ...),
#endif
#define MN_THREADS ((1 + MN_PARTNERS) * MN_WARPS_PER_BLOCK * 32)   // which share the 6502 warp's environments evenly
#define MN_CORE_WORDS 43   // EnvState is 43 words: an odd stride, conflict-free across slots
#define MN_PLANE (MN_IMG * MN_IMG)

static_assert(sizeof(EnvState) == 172 && MN_CORE_WORDS * 4 == 172, "EnvState layout changed: update MN_CORE_WORDS (odd)");

// PIL Image.resize((84,84), NEAREST) column map for a 160-wide source (atari_emulator.py:84); rows
// are floor((y + 0.5) * 2.5).  Golden copy: tests/golden/resize_lut.npz.
__constant__ uint8_t c_xmap[MN_IMG] = {
    0,  2,  4,  6,  8,  10, 12, 14, 16, 18, 20,  21,  23,  25,  27,  29,  31,  33,  35,  37,  39,
    40, 42, 44, 46, 48, 50, 52, 54, 56, 58, 60,  61,  63,  65,  67,  69,  71,  73,  75,  77,  79,
    80, 82, 84, 86, 88, 90, 92, 94, 96, 98, 99,  101, 103, 105, 107, 109, 111, 113, 115, 117, 119,
    120, 122, 124, 126, 128, 130, 132, 134, 136, 138, 139, 141, 143, 145, 147, 149, 151, 153, 155, 157, 159};

// NTSC palette of the emulated TIA (ALE getScreenRGB), index = colour byte >> 1
static const uint32_t h_ntsc[128] = {
    0x000000, 0x4a4a4a, 0x6f6f6f, 0x8e8e8e, 0xaaaaaa, 0xc0c0c0, 0xd6d6d6, 0xececec, 0x484800, 0x69690f, 0x86861d,
    0xa2a22a, 0xbbbb35, 0xd2d240, 0xe8e84a, 0xfcfc54, 0x7c2c00, 0x904811, 0xa26221, 0xb47a30, 0xc3903d, 0xd2a44a,
    0xdfb755, 0xecc860, 0x901c00, 0xa33915, 0xb55328, 0xc66c3a, 0xd5824a, 0xe39759, 0xf0aa67, 0xfcbc74, 0x940000,
    0xa71a1a, 0xb83232, 0xc84848, 0xd65c5c, 0xe46f6f, 0xf08080, 0xfc9090, 0x840064, 0x97197a, 0xa8308f, 0xb846a2,
    0xc659b3, 0xd46cc3, 0xe07cd2, 0xec8ce0, 0x500084, 0x68199a, 0x7d30ad, 0x9246c0, 0xa459d0, 0xb56ce0, 0xc57cee,
    0xd48cfc, 0x140090, 0x331aa3, 0x4e32b5, 0x6848c6, 0x7f5cd5, 0x956fe3, 0xa980f0, 0xbc90fc, 0x000094, 0x181aa7,
    0x2d32b8, 0x4248c8, 0x545cd6, 0x656fe4, 0x7580f0, 0x8490fc, 0x001c88, 0x183b9d, 0x2d57b0, 0x4272c2, 0x548ad2,
    0x65a0e1, 0x75b5ef, 0x84c8fc, 0x003064, 0x185080, 0x2d6d98, 0x4288b0, 0x54a0c5, 0x65b7d9, 0x75cceb, 0x84e0fc,
    0x004030, 0x18624e, 0x2d8169, 0x429e82, 0x54b899, 0x65d1ae, 0x75e7c2, 0x84fcd4, 0x004400, 0x1a661a, 0x328432,
    0x48a048, 0x5cba5c, 0x6fd26f, 0x80e880, 0x90fc90, 0x143c00, 0x355f18, 0x527e2d, 0x6e9c42, 0x87b754, 0x9ed065,
    0xb4e775, 0xc8fc84, 0x303800, 0x505916, 0x6d762b, 0x88923e, 0xa0ab4f, 0xb7c25f, 0xccd86e, 0xe0ec7c, 0x482c00,
    0x694d14, 0x866a26, 0xa28638, 0xbb9f47, 0xd2b656, 0xe8cc63, 0xfce070};
// packed per palette entry: byte0 luminance, byte1 R, byte2 G, byte3 B
__constant__ uint32_t c_pal[128];

static void host_palette(uint8_t* gray, uint8_t* rgb) {
  for (int i = 0; i < 128; ++i) {
    const uint8_t r = (h_ntsc[i] >> 16) & 0xFF, g = (h_ntsc[i] >> 8) & 0xFF, b = h_ntsc[i] & 0xFF;
    // ALE's luminance (ColourPalette::convertGrayscale of ALE >= 0.5): (uInt8) round(r * 0.2989 + g * 0.5870 + b * 0.1140)
    // in double arithmetic -- MN_ALE_LUMA_ROUND 0 restores round 1's truncation (see oracle/ale.hpp)
#ifndef MN_ALE_LUMA_ROUND
#define MN_ALE_LUMA_ROUND 1
#endif
    const double lum = double(r) * 0.2989 + double(g) * 0.5870 + double(b) * 0.1140;
    gray[i] = MN_ALE_LUMA_ROUND ? uint8_t(round(lum)) : uint8_t(lum);
    rgb[3 * i] = r; rgb[3 * i + 1] = g; rgb[3 * i + 2] = b;
  }
}

struct GameDev {
  int32_t rom_off, rom_size, game_id, cart, ctrl;
  int32_t env0, n_envs;      // env ids [env0, env0 + n_envs)
  int32_t blk0, n_blks;      // blocks of a k_round launch that serve this game
  int32_t n_actions;
  uint8_t actions[20];
};

struct PoolDev {             // kernel argument block (by value)
  GameDev games[MN_MAX_GAMES];
  int32_t n_games, n_envs, slots /* env slots per warp */, depth, num_actions, nb_choices;
  int32_t single_life, random_start, seed, env_id_offset, draw_all_frames;
  int32_t sync_slack;        // lanes of a warp stay within this many CPU cycles of the slowest one (see hot_time)
  int32_t fifo_high;         // a warp hands its TIA write queues off when one of them holds this many entries
  int32_t diag;              // MN_DIAG (measurement only, results are wrong): 1 = the picture side acknowledges without rendering
  int32_t tab_rep[32];
  const uint8_t* roms;
  EnvState* env;
  uint8_t* ram;              // (N,128)
  uint8_t* frames;           // (N,2,210,160)
  uint8_t* ring;             // (N,4,84,84,D)
  uint8_t* states;           // (N,84,84,4D)
  uint8_t* states_host;      // the caller's pinned host copy of `states` as the device sees it (mn_set_host_states), or null
  float *rewards, *terminals, *actions, *repetitions;
  int32_t *action_idx, *repetition_idx, *next_calls;
  // per-step bookkeeping
  int32_t* cur_action;       // ALE action enum of the running macro action
  int32_t* rep_left;
  int32_t* reward_acc;
  uint8_t* over;
  uint8_t* push_info;        // bits 0-1 ring slot, bits 2-3 pool mode (0 both, 1 buffer 0 only, 2 buffer 1 only)
  uint32_t* episode;         // episodes started so far (start no-op schedule)
  uint8_t* env_game;
  int32_t* lists;            // 4 x N : work list A, work list B, reset list, level-1 memo hits   (regions per game = env id ranges)
  int32_t* counts;           // 4 x MN_MAX_GAMES
  int32_t* error;            // sticky: 1 = episode over right after reset (atari_emulator.py:108-109)
  unsigned long long* total_next;
  // reset memoisation (see k_reset_prepare)
  uint8_t* memo;             // n_games x 75 timer seeds x memo_slots entries of memo_entry_bytes
  int32_t memo_entry_bytes, memo_enabled, memo_slots;
  // level 1 of the memo: the machine right after the reset UNIT (before the four start frames), see k_reset_prepare
  uint8_t* memo1; int32_t memo1_entry_bytes; int32_t* memo1_busy;
  uint32_t* reset_rnd;       // (N,) the RNG draw that seeds the RIOT timer of the reset in flight
  uint8_t* pre_ram;          // (N,128) RIOT RAM as it was before the reset in flight (memo key of a miss)
  int32_t* memo_hit;         // (N,) entry index the reset in flight is restored from
  int32_t* memo_busy;        // per (game, timer seed) bucket: one insertion at a time
  unsigned long long* track; // (N,5) def_lo def_hi dep_lo dep_hi flags : RAM-dependence probe of the reset in flight
  unsigned long long* memo_stats;   // hits, misses, inserts
  unsigned long long* total_instr;  // emulated 6502 instructions
  unsigned long long* diag_out;     // MN_DIAG=2 counters
  // reset-memo warm-up (mn_reset_all): scratch for the envs that lend themselves to it
  EnvState* warm_env; uint8_t* warm_ram; uint32_t* warm_episode;
  unsigned long long* redo_count;   // units re-run with every frame drawn (exact fallback of the pixel-less frames)
  uint8_t* history;          // (N, H, 84, 84, 4D) ring of the last H published states (paac.py:79-83,107-112), or null
  int32_t history_depth;
  const Tables* tables;
};

enum { ROUND_FIGAR = 0, ROUND_RESET = 1, ROUND_INITIAL = 2, ROUND_POWER_ON = 3, ROUND_SINGLE = 4 };

// ---- reset memoisation
// get_initial_state() = reset_game (64+ frames) + 16 NOOP frames is a pure function of (a) the RNG draw that
// seeds the RIOT timer, 25 + draw % 75, and (b) the RIOT RAM bytes the program READS BEFORE IT WRITES them
// after the console reset -- for 11 of the 12 README games there are none.  The first reset with a given key
// is emulated with a probe (k_round<true>) that records exactly that read-before-write set and which bytes got
// written; its result is stored.  Later resets whose RAM agrees on the entry's read-before-write bytes are
// bit-identical by determinism and are restored by copy (bytes the segment never wrote keep their own value).
#define MN_MEMO_SLOTS 8
#define MN_TIMER_SEEDS 75
struct MemoHdr {
  int32_t state;             // 0 empty, 1 being written, 2 valid
  int32_t n_acts;            // ALE act() calls inside the segment (2 RNG advances and one frame count each)
  int32_t noops;             // start no-ops the segment ran (random_start; part of the key)
  int32_t pad0;
  unsigned long long dep_lo, dep_hi, def_lo, def_hi;
  uint8_t dep_val[128];      // RAM before the reset (only the dep bytes matter)
  uint8_t ram[128];          // RAM after the segment (only the def bytes matter)
  EnvState env;              // after the segment; RNG / frame counter / ring head are the env's own on restore
  uint32_t ring_head;
  uint32_t pad;
};
__host__ __device__ inline size_t memo_hdr_bytes() { return (sizeof(MemoHdr) + 15) & ~size_t(15); }

// ----------------------------------------------------------------------------- kernels
__global__ void k_begin_step(PoolDev p, int use_indices) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < MN_MAX_GAMES * 3) {
    // list 0 starts full, list 1 and the reset list empty
    const int which = e / MN_MAX_GAMES, g = e % MN_MAX_GAMES;
    p.counts[e] = (which == 0 && g < p.n_games) ? p.games[g].n_envs : 0;
  }
  if (e >= p.n_envs) return;
  int a = 0, r = 0;
  if (use_indices) { a = p.action_idx[e]; r = p.repetition_idx[e]; }
  else {   // np.argmax of the one-hot rows (exploration_policy.py:18-19): first maximum
    const float* pa = p.actions + size_t(e) * p.num_actions;
    float best = pa[0];
    for (int i = 1; i < p.num_actions; ++i) if (pa[i] > best) { best = pa[i]; a = i; }
    const float* pr = p.repetitions + size_t(e) * p.nb_choices;
    best = pr[0];
    for (int i = 1; i < p.nb_choices; ++i) if (pr[i] > best) { best = pr[i]; r = i; }
  }
  const GameDev& g = p.games[p.env_game[e]];
  if (a < 0) a = 0; if (a >= g.n_actions) a = g.n_actions - 1;
  if (r < 0) r = 0; if (r >= p.nb_choices) r = p.nb_choices - 1;
  p.cur_action[e] = g.actions[a];
  p.rep_left[e] = p.tab_rep[r];
  p.reward_acc[e] = 0;
  p.over[e] = 0;
  p.next_calls[e] = 0;
  p.lists[e] = e;
}

// fills a list with every env (reset_all / power-on)
__global__ void k_fill_list(PoolDev p, int which) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < MN_MAX_GAMES * 3) {
    const int w = e / MN_MAX_GAMES, g = e % MN_MAX_GAMES;
    p.counts[e] = (w == which && g < p.n_games) ? p.games[g].n_envs : 0;
  }
  if (e >= p.n_envs) return;
  p.lists[size_t(which) * p.n_envs + e] = e;
  p.reward_acc[e] = 0; p.over[e] = 0; p.next_calls[e] = 0; p.rep_left[e] = 0; p.cur_action[e] = 0;
}
__global__ void k_single_list(PoolDev p, int which, int env, int ale_action) {
  if (threadIdx.x < MN_MAX_GAMES * 3) p.counts[threadIdx.x] = 0;
  __syncthreads();
  if (threadIdx.x == 0) {
    const int g = p.env_game[env];
    p.counts[which * MN_MAX_GAMES + g] = 1;
    p.lists[size_t(which) * p.n_envs + p.games[g].env0] = env;
    p.cur_action[env] = ale_action; p.rep_left[env] = 0; p.reward_acc[env] = 0; p.over[env] = 0; p.next_calls[env] = 0;
  }
}

// The picture side of the envs of one 6502 warp, on its partner warp: lane l serves the env of lane l.  It waits for
// hand-offs (emu_core.cuh tia_handoff) and renders them while the 6502 warp runs on.  Requests published by one
// instruction of the 6502 warp (the warp-wide hand-off of k_round's loop) are seen together and rendered by all
// lanes in step -- the long branchy rendering code is entered by the whole warp, as it was when it ran inline.
__device__ __forceinline__ void picture_warp(Ctx& c, unsigned wmask, int diag) {
  uint32_t* mb = c.fifo + MN_MBOX;
  uint32_t pseq = 0;
  bool fin = false;
  int idle = 0;
  for (;;) {
    const bool req = !fin && mbox_load(mb + MB_HAND) != pseq;
    if (!__any_sync(wmask, req)) {
      if (__all_sync(wmask, fin)) break;
      // An idle partner shares its SM sub-partition's issue slots with the 6502 warp it serves: poll rarely.  Hand-offs
      // come ~50 us apart and only two kinds are waited for (collision-latch reads, the end of a unit).  Measured alone
      // (tools/sleep_probe.cu): __nanosleep(20) = 115 cycles, (200) = 502, (512...1000) = 2,008 -- so the counted run below
      // sleeps 1 us after the first empty poll and up to 6 us once nothing has arrived for a while.  (The per-line
      // instruction counts ncu attributes to this loop are inflated: its instrumented passes stretch the kernel, not the
      // sleeps.)  A finer first back-off (8 polls of __nanosleep(20)) measured no gain, Pong included.
      if (idle < 24) idle += 4;
#pragma unroll 1
      for (int k = 0; k < idle; ++k) __nanosleep(200);
      continue;
    }
    idle = 0;
    if (req) {
      __threadfence_block();
      const int buf = int(pseq & (MN_FIFO_NBUF - 1));
      const uint32_t rq = mbox_load(mb + MB_REQ + 2 * buf);
      const int32_t sync_clk = int32_t(mbox_load(mb + MB_REQ + 2 * buf + 1));
      const int cmd = int(rq >> 8);
      if (cmd == PIC_EXIT) fin = true;
      else if (!(diag & 1)) picture_process(c, buf, int(rq & 0xFFu), sync_clk, cmd);
      __threadfence_block();
      mbox_store(mb + MB_DONE, ++pseq);
    }
  }
}

// The flat warp loop of k_round.  One warp-wide reduction per tick carries all three decisions: the emulated time of
// the lane furthest behind (lanes within sync_slack of it run), "some lane's TIA write queue is nearly full" (-1: every
// lane hands off to the picture side) and "no lane has work left" (INT_MAX).
template <bool TRACK, bool FLAT>
__device__ __forceinline__ void run_units(Ctx& c, const Mem& mm, Unit& u, Hot& hot, unsigned wmask, int sync_slack, int fifo_high) {
  for (;;) {
    const bool work = hot_has_work(hot);
    int now = work ? hot_time(hot) : 0x7FFFFFFF;
    if (MN_FILL(hot.cpu.fifo_n) >= fifo_high) now = -1;
    const int first = __reduce_min_sync(wmask, now);
    if (((uint32_t(first) + 1u) & 0x7FFFFFFFu) == 0u) {   // -1 or INT_MAX: one test (one branch) on the path every tick takes
      if (first < 0) { hot_drain(c, hot); continue; }
      break;
    }
    unit_tick<TRACK, FLAT>(c, mm, u, hot, work && now - first <= sync_slack);
  }
}

// One round of emulation for the envs on list `in`.  Dynamic shared memory:
//   [rom | tables | core slots (4 warps x slots x 43 words) | ram (4 warps x slots x 132 B: 128 used, odd word pitch)
//    | TIA write queues (4 warps x slots x MN_FIFO_WORDS: NBUF buffers of 16 + the hand-off mailbox)]
template <bool TRACK>
__global__ void __launch_bounds__(MN_THREADS, 1) k_round(PoolDev p, int mode, int in, int out) {
  extern __shared__ __align__(16) uint8_t smem[];
  // ---- which game does this block serve
  int gi = 0;
  while (gi + 1 < p.n_games && int(blockIdx.x) >= p.games[gi + 1].blk0) ++gi;
  const GameDev& G = p.games[gi];
  const int count = p.counts[in * MN_MAX_GAMES + gi];
  const int j = blockIdx.x - G.blk0;
  const int lo = int((long long)count * j / G.n_blks), hi = int((long long)count * (j + 1) / G.n_blks);
  if (hi <= lo) return;   // whole block idle (uniform)
  const int rom_bytes = (G.rom_size == 2048) ? 4096 : ((G.rom_size + 15) & ~15);
  const int nslots = MN_WARPS_PER_BLOCK * p.slots;
  uint8_t* s_rom = smem;
  Tables* s_tab = reinterpret_cast<Tables*>(smem + rom_bytes);
  uint32_t* s_core = reinterpret_cast<uint32_t*>(smem + rom_bytes + sizeof(Tables));
  uint8_t* s_ram = reinterpret_cast<uint8_t*>(s_core + nslots * MN_CORE_WORDS);
  uint32_t* s_fifo = reinterpret_cast<uint32_t*>(s_ram + nslots * MN_RAM_PITCH);
  {
    const uint4* src = reinterpret_cast<const uint4*>(p.roms + G.rom_off);
    uint4* dst = reinterpret_cast<uint4*>(s_rom);
    for (int i = threadIdx.x; i < ((G.rom_size + 15) & ~15) / 16; i += blockDim.x) dst[i] = src[i];   // (the image, not the window: a 2K image is 2 KB in HBM)
    // a 2K image a second time behind itself: the cartridge window is then 4 KB of consecutive bytes like any other
    if (G.rom_size == 2048) for (int i = threadIdx.x; i < 2048 / 16; i += blockDim.x) dst[2048 / 16 + i] = src[i];
    const uint32_t* ts = reinterpret_cast<const uint32_t*>(p.tables);
    uint32_t* td = reinterpret_cast<uint32_t*>(s_tab);
    for (int i = threadIdx.x; i < int(sizeof(Tables) / 4); i += blockDim.x) td[i] = ts[i];
    for (int i = threadIdx.x; i < nslots; i += blockDim.x) {   // mailboxes: nothing handed off, nothing done
      uint32_t* mb = s_fifo + i * MN_FIFO_WORDS + MN_MBOX;
      mb[MB_HAND] = 0u; mb[MB_DONE] = 0u;
    }
  }
  __syncthreads();
  // ---- my env (the same for a 6502 lane and its picture-side partner)
  const int warp = (threadIdx.x >> 5) & (MN_WARPS_PER_BLOCK - 1);
  const int role = threadIdx.x / (MN_WARPS_PER_BLOCK * 32);   // 0: 6502 warp, 1..MN_PARTNERS: picture-side partner
  const bool picture_side = role != 0;
  const int m = hi - lo;
  const int wlo = lo + m * warp / MN_WARPS_PER_BLOCK, whi = lo + m * (warp + 1) / MN_WARPS_PER_BLOCK;
  // a partner warp serves a contiguous share of the 6502 warp's lanes with its first lanes (the picture side is the
  // divergent half of the machine: fewer environments per warp cost it less than they would cost the 6502 side)
  const int per = (whi - wlo + MN_PARTNERS - 1) / MN_PARTNERS;
  const int lane = picture_side ? (role - 1) * per + int(threadIdx.x & 31) : int(threadIdx.x & 31);
  const bool active = lane < whi - wlo && (!picture_side || int(threadIdx.x & 31) < per);
  const unsigned wmask = __ballot_sync(0xFFFFFFFFu, active);   // the lanes that vote in the loops below
  if (!active) return;
  const int e = p.lists[size_t(in) * p.n_envs + G.env0 + wlo + lane];
  const int slot = warp * p.slots + lane;
  EnvState* s = reinterpret_cast<EnvState*>(s_core + slot * MN_CORE_WORDS);
  Ctx c;
  c.s = s; c.rom = s_rom; c.tab = s_tab;
  c.ram = s_ram + slot * MN_RAM_PITCH;
  c.fb = p.frames + size_t(e) * (2 * MN_FRAME_BYTES);
  c.fifo = s_fifo + slot * MN_FIFO_WORDS;
  c.fifo_n = 0; c.hseq = 0; c.mbox_timeout = false; c.wait_free = c.wait_done = c.n_handoff = 0;
  const long long t_begin = clock64();
  if (picture_side) { picture_warp(c, wmask, p.diag); return; }
  // what this launch asks of the env
  int kind = U_ACTS, action = 0, ucount = MN_ACTION_REPEAT;
  uint32_t seed = 0;
  if (mode == ROUND_POWER_ON) {
    // atari_emulator.py:20: ALE seed = random_seed * (actor_id + 1); loadROM resets once
    kind = U_POWER_ON; ucount = 0; seed = uint32_t(p.seed) * uint32_t(p.env_id_offset + e + 1);
    p.episode[e] = 0;
  } else if (mode == ROUND_RESET) {
    const uint32_t ep = p.episode[e];
    kind = U_RESET;
    ucount = p.random_start ? int(start_noops(uint32_t(p.seed), uint32_t(p.env_id_offset + e), ep)) : 0;
    seed = p.reset_rnd[e];   // drawn by k_reset_prepare
    p.episode[e] = ep + 1;
  } else action = (mode == ROUND_INITIAL) ? int(G.actions[0]) : p.cur_action[e];

  // First pass: only the pooled frames keep their pixels.  If that turns out to have been visible
  // (unit_finish), the env is reloaded and the unit re-run with every frame drawn.
  Unit res;
  bool bad = false;
  for (int attempt = 0; attempt < 2; ++attempt) {
    const bool mine = (attempt == 0) || bad;
    Unit u;
    unit_idle(u);
    if (mine) {
      const uint32_t* src = reinterpret_cast<const uint32_t*>(p.env + e);
      uint32_t* dst = reinterpret_cast<uint32_t*>(s);
#pragma unroll 6
      for (int i = 0; i < int(sizeof(EnvState) / 4); ++i) dst[i] = src[i];
      const uint32_t* rsrc = reinterpret_cast<const uint32_t*>(p.ram + size_t(e) * 128);
      for (int i = 0; i < 32; ++i) *reinterpret_cast<uint32_t*>(c.ram + i * 4) = rsrc[i];
      if (mode == ROUND_POWER_ON) { s->game = uint8_t(G.game_id); s->cart = uint8_t(G.cart); s->ctrl = uint8_t(G.ctrl); s->host_lives = 0; }
      c.all_pixels = (attempt == 1) || (p.draw_all_frames != 0);
      unit_init(c, u, kind, action, ucount, seed);
    }
    Hot hot;
    hot_init(c, u, hot);
    if (TRACK && mine && mode == ROUND_INITIAL) {   // the probe continues from the reset round
      const unsigned long long* t = p.track + size_t(e) * 5;
      hot.def_lo = t[0]; hot.def_hi = t[1]; hot.dep_lo = t[2]; hot.dep_hi = t[3];
      hot.tainted = (t[4] & 1ull) != 0; hot.obs_bad = (t[4] & 2ull) != 0;
    }
    const Mem mm = mem_of(c);
    // (block-uniform: a block serves one cartridge)
    if (G.cart != CART_E0) run_units<TRACK, true>(c, mm, u, hot, wmask, p.sync_slack, p.fifo_high);
    else run_units<TRACK, false>(c, mm, u, hot, wmask, p.sync_slack, p.fifo_high);
    if (mine) {
      bad = unit_finish(c, hot); res = u; atomicAdd(p.total_instr, (unsigned long long)hot.instr);
      // an env whose game ended inside this macro step is reset before anyone can look at its frames
      if (mode == ROUND_FIGAR && (s->flags & F_TERMINAL)) bad = false;
      if (TRACK) {
        unsigned long long* t = p.track + size_t(e) * 5;
        t[0] = hot.def_lo; t[1] = hot.def_hi; t[2] = hot.dep_lo; t[3] = hot.dep_hi;
        t[4] = (hot.tainted ? 1ull : 0ull) | (hot.obs_bad ? 2ull : 0ull);
      }
    }
    if (!__any_sync(wmask, bad)) break;
    if (attempt == 0 && bad) atomicAdd(p.redo_count, 1ull);
  }
  tia_handoff(c, -1, PIC_EXIT, false);   // the partner lane leaves
  if (c.mbox_timeout) atomicExch(p.error, 2);
  if (p.diag & 2) {   // MN_DIAG=2: where the 6502 lanes waited (lane-clocks, summed over the launches)
    atomicAdd(p.diag_out + 0, (unsigned long long)c.wait_free); atomicAdd(p.diag_out + 1, (unsigned long long)c.wait_done);
    atomicAdd(p.diag_out + 2, (unsigned long long)(clock64() - t_begin)); atomicAdd(p.diag_out + 3, (unsigned long long)c.n_handoff);
  }

  const bool single_life = p.single_life != 0;
  if (mode != ROUND_POWER_ON && mode != ROUND_RESET) {
    const NextOut o = env_next_result(*s, res, single_life);
    const int head = s->ring_head;
    s->ring_head = uint8_t((head + 1) & (MN_STACK - 1));
    const int pool_mode = o.pool_single ? ((s->pflags & F_CURFB) ? 2 : 1) : 0;
    p.push_info[e] = uint8_t(head | (pool_mode << 2));
    if (mode == ROUND_INITIAL) {
      if (out == -1 && o.terminal) atomicExch(p.error, 1);   // last of the four start frames: 'This should never happen.'
    } else {
      p.reward_acc[e] += o.reward;
      p.next_calls[e] += 1;
      atomicAdd(p.total_next, 1ull);
      if (o.terminal) {
        p.over[e] = 1;
        if (mode == ROUND_FIGAR) {
          const int k = atomicAdd(&p.counts[2 * MN_MAX_GAMES + gi], 1);
          p.lists[size_t(2) * p.n_envs + G.env0 + k] = e;
        }
      } else if (mode == ROUND_FIGAR) {
        const int left = p.rep_left[e];
        if (left > 0) {
          p.rep_left[e] = left - 1;
          const int k = atomicAdd(&p.counts[out * MN_MAX_GAMES + gi], 1);
          p.lists[size_t(out) * p.n_envs + G.env0 + k] = e;
        } else {
          p.rep_left[e] = -p.next_calls[e];   // done, not terminal: -(round + 1) tells k_emit_early which round it was
        }
      }
    }
  }
  // ---- write the machine back
  {
    uint32_t* dst = reinterpret_cast<uint32_t*>(p.env + e);
    const uint32_t* src = reinterpret_cast<const uint32_t*>(s);
#pragma unroll 6
    for (int i = 0; i < int(sizeof(EnvState) / 4); ++i) dst[i] = src[i];
    uint32_t* rd = reinterpret_cast<uint32_t*>(p.ram + size_t(e) * 128);
    for (int i = 0; i < 32; ++i) rd[i] = *reinterpret_cast<uint32_t*>(c.ram + i * 4);
  }
}

__global__ void k_clear_counts(PoolDev p, int which) {
  if (threadIdx.x < MN_MAX_GAMES) p.counts[which * MN_MAX_GAMES + threadIdx.x] = 0;
}

// ---- reset-memo warm-up.  A reset that misses the memo puts ~80 emulated frames on the critical path of the macro step
// it falls into.  For the games whose reset depends on nothing but the RIOT timer seed (11 of the 12 README games) all
// 75 possible results can be produced once, up front: the first min(n, 75) envs of every game lend themselves -- their
// machine, RAM and episode counter are saved, they run the reset with a FORCED timer seed through the ordinary probe /
// insert path, and are put back.  `pass` p covers seeds p*W .. p*W + W - 1 (W = envs lent by the game).
__global__ void k_warm_begin(PoolDev p, int pass) {
  const int gi = blockIdx.x, i = threadIdx.x;            // one block per game, one thread per lent env
  const GameDev& G = p.games[gi];
  const int W = G.n_envs < MN_TIMER_SEEDS ? G.n_envs : MN_TIMER_SEEDS;
  __shared__ int s_count;
  if (i == 0) s_count = 0;
  __syncthreads();
  const int seed = pass * W + i;
  if (i < W && seed < MN_TIMER_SEEDS) {
    const int e = G.env0 + i;
    const int slot = gi * MN_TIMER_SEEDS + i;
    p.warm_env[slot] = p.env[e];
    p.warm_episode[slot] = p.episode[e];
    for (int j = 0; j < 128; ++j) { const uint8_t v = p.ram[size_t(e) * 128 + j]; p.warm_ram[size_t(slot) * 128 + j] = v; p.pre_ram[size_t(e) * 128 + j] = v; }
    p.reset_rnd[e] = uint32_t(seed);                     // console_reset: timer = 25 + rnd % 75
    p.memo_hit[e] = -1;
    const int k = atomicAdd(&s_count, 1);
    p.lists[size_t(1) * p.n_envs + G.env0 + k] = e;
  }
  __syncthreads();
  if (i == 0) { p.counts[1 * MN_MAX_GAMES + gi] = s_count; p.counts[0 * MN_MAX_GAMES + gi] = 0; }
}
__global__ void k_warm_end(PoolDev p, int pass) {
  const int gi = blockIdx.x, i = threadIdx.x;
  const GameDev& G = p.games[gi];
  const int W = G.n_envs < MN_TIMER_SEEDS ? G.n_envs : MN_TIMER_SEEDS;
  if (i < W && pass * W + i < MN_TIMER_SEEDS) {
    const int e = G.env0 + i;
    const int slot = gi * MN_TIMER_SEEDS + i;
    p.env[e] = p.warm_env[slot];
    p.episode[e] = p.warm_episode[slot];
    for (int j = 0; j < 128; ++j) p.ram[size_t(e) * 128 + j] = p.warm_ram[size_t(slot) * 128 + j];
  }
}

// one bucket of a memo table: the entry whose key (start no-ops, values of its dependence bytes) the env matches, or -1
__device__ __forceinline__ int memo_find(const uint8_t* table, int entry_bytes, int slots, size_t bucket, int noops,
                                         const unsigned long long* ram8) {
  for (int sl = 0; sl < slots; ++sl) {
    const MemoHdr* m = reinterpret_cast<const MemoHdr*>(table + (bucket + sl) * size_t(entry_bytes));
    if (*reinterpret_cast<const volatile int32_t*>(&m->state) != 2) continue;
    bool same = (m->noops == noops);
    if (m->dep_lo | m->dep_hi) {
      const unsigned long long* want = reinterpret_cast<const unsigned long long*>(m->dep_val);
      for (int w = 0; w < 16 && same; ++w) {
        const unsigned long long bits = ((w < 8) ? m->dep_lo : m->dep_hi) >> ((w & 7) * 8);
        unsigned long long mask = 0;   // one 0xFF per dependent byte of this 8-byte word
        for (int b = 0; b < 8; ++b) if ((bits >> b) & 1ull) mask |= 0xFFull << (8 * b);
        same = ((ram8[w] ^ want[w]) & mask) == 0ull;
      }
    }
    if (same) return int(bucket + sl);
  }
  return -1;
}

// Every env on the reset list draws the RNG value that seeds its RIOT timer (ALE's System::reset) and looks for a memo
// entry it may be restored from.  Two levels: (2) the whole of get_initial_state() -- hits go to list 0 and are restored by
// copy; (1) the reset UNIT alone (reset_game and the start no-ops, 64+ of the 80 frames) -- hits go to list 3, are restored
// to that point and only emulate the four start frames.  Level 1 exists for games whose start frames read RAM the reset
// unit does not (Yars' Revenge: a pseudo-random byte, so level 2 almost never recurs, while the reset unit depends on
// two bytes with a handful of values).  Misses (list 1) are emulated with the probe and stored at both levels.
__global__ void k_reset_prepare(PoolDev p) {
  const int pos = blockIdx.x * blockDim.x + threadIdx.x;
  if (pos >= p.n_envs) return;
  int gi = 0;
  while (gi + 1 < p.n_games && pos >= p.games[gi + 1].env0) ++gi;
  if (pos - p.games[gi].env0 >= p.counts[2 * MN_MAX_GAMES + gi]) return;
  const int e = p.lists[size_t(2) * p.n_envs + pos];
  const uint32_t rnd = rng_next(p.env[e].rng);
  p.reset_rnd[e] = rnd;
  int hit = -1, which = 1;
  const unsigned long long* ram8 = reinterpret_cast<const unsigned long long*>(p.ram + size_t(e) * 128);
  // the start no-ops the reset will run (k_round<RESET> computes the same): part of the key under random_start
  const int noops = p.random_start ? int(start_noops(uint32_t(p.seed), uint32_t(p.env_id_offset + e), p.episode[e])) : 0;
  if (p.memo_enabled) {
    const size_t bucket = (size_t(gi) * MN_TIMER_SEEDS + rnd % MN_TIMER_SEEDS) * size_t(p.memo_slots);
    hit = memo_find(p.memo, p.memo_entry_bytes, p.memo_slots, bucket, noops, ram8);
    if (hit >= 0) which = 0;
    else {
      hit = memo_find(p.memo1, p.memo1_entry_bytes, p.memo_slots, bucket, noops, ram8);
      if (hit >= 0) which = 3;
    }
  }
  p.memo_hit[e] = hit;
  const int k = atomicAdd(&p.counts[which * MN_MAX_GAMES + gi], 1);
  p.lists[size_t(which) * p.n_envs + p.games[gi].env0 + k] = e;
  atomicAdd(p.memo_stats + (which == 0 ? 0 : 1), 1ull);
  if (which == 3) atomicAdd(p.memo_stats + 3, 1ull);
  if (which != 0) {
    unsigned long long* dst = reinterpret_cast<unsigned long long*>(p.pre_ram + size_t(e) * 128);
    for (int w = 0; w < 16; ++w) dst[w] = ram8[w];
  }
}

// restores the envs on list 0 from their memo entries: machine, RAM, both frame buffers, the four ring planes
template <int D>
__global__ void __launch_bounds__(256) k_reset_restore(PoolDev p) {
  const int pos = blockIdx.x;
  int gi = 0;
  while (gi + 1 < p.n_games && pos >= p.games[gi + 1].env0) ++gi;
  if (pos - p.games[gi].env0 >= p.counts[0 * MN_MAX_GAMES + gi]) return;
  const int e = p.lists[size_t(0) * p.n_envs + pos];
  const uint8_t* entry = p.memo + size_t(p.memo_hit[e]) * size_t(p.memo_entry_bytes);
  const MemoHdr* m = reinterpret_cast<const MemoHdr*>(entry);
  const uint4* src = reinterpret_cast<const uint4*>(entry + memo_hdr_bytes());
  uint4* fb = reinterpret_cast<uint4*>(p.frames + size_t(e) * (2 * MN_FRAME_BYTES));
  for (int i = threadIdx.x; i < 2 * MN_FRAME_BYTES / 16; i += blockDim.x) fb[i] = src[i];
  src += 2 * MN_FRAME_BYTES / 16;
  const int head = p.env[e].ring_head;   // unchanged by four pushes
  for (int k = 0; k < MN_STACK; ++k) {   // memo plane k = k-th oldest
    uint4* dst = reinterpret_cast<uint4*>(p.ring + (size_t(e) * MN_STACK + ((head + k) & 3)) * (MN_PLANE * D));
    for (int i = threadIdx.x; i < MN_PLANE * D / 16; i += blockDim.x) dst[i] = src[size_t(k) * (MN_PLANE * D / 16) + i];
  }
  if (threadIdx.x < 128) {
    const int j = threadIdx.x;
    const bool def = (((j & 64) ? m->def_hi : m->def_lo) >> (j & 63)) & 1ull;
    if (def) p.ram[size_t(e) * 128 + j] = m->ram[j];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    EnvState own = p.env[e];
    EnvState s = m->env;
    for (int i = 0; i < 4; ++i) s.rng[i] = own.rng[i];
    for (int i = 0; i < 2 * m->n_acts; ++i) rng_advance(s.rng);
    s.frame_number = own.frame_number + m->n_acts;
    s.ring_head = own.ring_head;
    p.env[e] = s;
    p.episode[e] += 1;
  }
}

// level-1 hits (list 3): the machine, RAM and frame buffers as the reset unit left them, the probe's sets as they stood
// there; the envs then join list 1 for the four start frames (and are stored at level 2 like the emulated ones)
__global__ void __launch_bounds__(256) k_reset_restore_l1(PoolDev p) {
  const int pos = blockIdx.x;
  int gi = 0;
  while (gi + 1 < p.n_games && pos >= p.games[gi + 1].env0) ++gi;
  if (pos - p.games[gi].env0 >= p.counts[3 * MN_MAX_GAMES + gi]) return;
  const int e = p.lists[size_t(3) * p.n_envs + pos];
  const uint8_t* entry = p.memo1 + size_t(p.memo_hit[e]) * size_t(p.memo1_entry_bytes);
  const MemoHdr* m = reinterpret_cast<const MemoHdr*>(entry);
  const uint4* src = reinterpret_cast<const uint4*>(entry + memo_hdr_bytes());
  uint4* fb = reinterpret_cast<uint4*>(p.frames + size_t(e) * (2 * MN_FRAME_BYTES));
  for (int i = threadIdx.x; i < 2 * MN_FRAME_BYTES / 16; i += blockDim.x) fb[i] = src[i];
  if (threadIdx.x < 128) {
    const int j = threadIdx.x;
    const bool def = (((j & 64) ? m->def_hi : m->def_lo) >> (j & 63)) & 1ull;
    if (def) p.ram[size_t(e) * 128 + j] = m->ram[j];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    EnvState own = p.env[e];
    EnvState s = m->env;
    for (int i = 0; i < 4; ++i) s.rng[i] = own.rng[i];
    for (int i = 0; i < 2 * m->n_acts; ++i) rng_advance(s.rng);     // the start no-ops are act() calls
    s.frame_number = own.frame_number + m->n_acts;
    s.ring_head = own.ring_head;
    p.env[e] = s;
    p.episode[e] += 1;
    unsigned long long* t = p.track + size_t(e) * 5;
    t[0] = m->def_lo; t[1] = m->def_hi; t[2] = m->dep_lo; t[3] = m->dep_hi; t[4] = (m->dep_lo | m->dep_hi) ? 1ull : 0ull;
    const int k = atomicAdd(&p.counts[1 * MN_MAX_GAMES + gi], 1);
    p.lists[size_t(1) * p.n_envs + p.games[gi].env0 + k] = e;
  }
}

// after the probe emulated the misses (list 1): store what it found.  LEVEL 2: after the four start frames (machine, RAM,
// frame buffers, ring planes); LEVEL 1: right after the reset unit (machine, RAM, frame buffers)
template <int D, int LEVEL>
__global__ void __launch_bounds__(256) k_memo_insert(PoolDev p) {
  uint8_t* const table = (LEVEL == 2) ? p.memo : p.memo1;
  const int entry_bytes = (LEVEL == 2) ? p.memo_entry_bytes : p.memo1_entry_bytes;
  int32_t* const busy = (LEVEL == 2) ? p.memo_busy : p.memo1_busy;
  __shared__ int s_slot;
  const int pos = blockIdx.x;
  int gi = 0;
  while (gi + 1 < p.n_games && pos >= p.games[gi + 1].env0) ++gi;
  if (pos - p.games[gi].env0 >= p.counts[1 * MN_MAX_GAMES + gi]) return;
  const int e = p.lists[size_t(1) * p.n_envs + pos];
  const unsigned long long* t = p.track + size_t(e) * 5;
  const int bucket_id = gi * MN_TIMER_SEEDS + int(p.reset_rnd[e] % MN_TIMER_SEEDS);
  // (the reset round has moved the episode counter on: the no-ops it ran were those of episode - 1)
  const int my_noops = p.random_start ? int(start_noops(uint32_t(p.seed), uint32_t(p.env_id_offset + e), p.episode[e] - 1u)) : 0;
  if (threadIdx.x == 0) {
    int slot = -1;
    // the RAM scrape must only have seen written bytes; one insertion per bucket at a time
    if (!(t[4] & 2ull) && atomicCAS(&busy[bucket_id], 0, 1) == 0) {
      const size_t bucket = size_t(bucket_id) * size_t(p.memo_slots);
      const uint8_t* pre = p.pre_ram + size_t(e) * 128;
      bool known = false;
      for (int sl = 0; sl < p.memo_slots && !known; ++sl) {   // stored since k_reset_prepare looked?
        const MemoHdr* m = reinterpret_cast<const MemoHdr*>(table + (bucket + sl) * size_t(entry_bytes));
        if (*reinterpret_cast<const volatile int32_t*>(&m->state) != 2) continue;
        bool same = (m->noops == my_noops);
        for (int j = 0; j < 128 && same; ++j)
          if ((((j & 64) ? m->dep_hi : m->dep_lo) >> (j & 63)) & 1ull) same = (m->dep_val[j] == pre[j]);
        known = same;
      }
      for (int sl = 0; sl < p.memo_slots && !known && slot < 0; ++sl) {
        MemoHdr* m = reinterpret_cast<MemoHdr*>(table + (bucket + sl) * size_t(entry_bytes));
        if (atomicCAS(&m->state, 0, 1) == 0) slot = int(bucket + sl);
      }
      if (slot < 0) atomicExch(&busy[bucket_id], 0);
    }
    s_slot = slot;
  }
  __syncthreads();
  if (s_slot < 0) return;
  uint8_t* entry = table + size_t(s_slot) * size_t(entry_bytes);
  MemoHdr* m = reinterpret_cast<MemoHdr*>(entry);
  uint4* dst = reinterpret_cast<uint4*>(entry + memo_hdr_bytes());
  const uint4* fb = reinterpret_cast<const uint4*>(p.frames + size_t(e) * (2 * MN_FRAME_BYTES));
  for (int i = threadIdx.x; i < 2 * MN_FRAME_BYTES / 16; i += blockDim.x) dst[i] = fb[i];
  dst += 2 * MN_FRAME_BYTES / 16;
  const int head = p.env[e].ring_head;
  if (LEVEL == 2) for (int k = 0; k < MN_STACK; ++k) {
    const uint4* src = reinterpret_cast<const uint4*>(p.ring + (size_t(e) * MN_STACK + ((head + k) & 3)) * (MN_PLANE * D));
    for (int i = threadIdx.x; i < MN_PLANE * D / 16; i += blockDim.x) dst[size_t(k) * (MN_PLANE * D / 16) + i] = src[i];
  }
  if (threadIdx.x < 128) { m->dep_val[threadIdx.x] = p.pre_ram[size_t(e) * 128 + threadIdx.x]; m->ram[threadIdx.x] = p.ram[size_t(e) * 128 + threadIdx.x]; }
  if (threadIdx.x == 0) {
    m->def_lo = t[0]; m->def_hi = t[1]; m->dep_lo = t[2]; m->dep_hi = t[3];
    m->env = p.env[e];
    m->n_acts = (LEVEL == 2 ? MN_STACK * MN_ACTION_REPEAT : 0) + my_noops;
    m->noops = my_noops;
    m->ring_head = uint32_t(head);
    if (LEVEL == 2) atomicAdd(p.memo_stats + 2, 1ull);
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    *reinterpret_cast<volatile int32_t*>(&m->state) = 2;
    __threadfence();
    atomicExch(&busy[bucket_id], 0);
  }
}

// K3 core, one block of 256 threads per environment.  a, b: the two raw palette-index screens (210x160);
// mode 0: max of both, 1: a only, 2: b only.  Only the 84 source rows the nearest-neighbour map keeps are read, each
// as whole 160-byte rows (8-byte loads, one warp per row), mapped through a shared-memory palette (128 luminance
// bytes fill the 32 banks exactly once: conflict-free; the RGB table is 128 words), reduced with max, and gathered
// into the output plane in shared memory; the plane leaves with 16-byte stores.  Ends with a barrier (callers loop).
template <int D>
__device__ __forceinline__ void preprocess_plane(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, int mode,
                                                 uint8_t* __restrict__ plane) {
  __shared__ __align__(16) uint32_t s_pal[D == 1 ? 32 : 128];
  __shared__ __align__(16) uint32_t s_row[8][D == 1 ? MN_SCREEN_W / 4 : MN_SCREEN_W];
  __shared__ __align__(16) uint32_t s_out[MN_PLANE * D / 4];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (mode == 1) b = a; else if (mode == 2) a = b;
  if (tid < 128) {
    if (D == 1) reinterpret_cast<uint8_t*>(s_pal)[tid] = uint8_t(c_pal[tid]);
    else s_pal[tid] = c_pal[tid] >> 8;                                     // 0x00BBGGRR
  }
  // output word `lane` of a row = these four source columns (PIL's nearest map, the same for every row)
  int x0 = 0, x1 = 0, x2 = 0, x3 = 0;
  if (lane < MN_IMG / 4) { x0 = c_xmap[4 * lane]; x1 = c_xmap[4 * lane + 1]; x2 = c_xmap[4 * lane + 2]; x3 = c_xmap[4 * lane + 3]; }
  __syncthreads();
  for (int y = warp; y < MN_IMG; y += 8) {
    const int row = ((2 * y + 1) * 5) >> 2;   // floor((y + 0.5) * 2.5)
    const uint8_t* ra = a + row * MN_SCREEN_W;
    const uint8_t* rb = b + row * MN_SCREEN_W;
    if (D == 1) {
      if (lane < MN_SCREEN_W / 8) {
        const uint2 va = *reinterpret_cast<const uint2*>(ra + lane * 8), vb = *reinterpret_cast<const uint2*>(rb + lane * 8);
        const uint8_t* pal = reinterpret_cast<const uint8_t*>(s_pal);
        uint32_t lo = 0, hi = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t la = pal[(va.x >> (8 * k + 1)) & 0x7Fu], lb = pal[(vb.x >> (8 * k + 1)) & 0x7Fu];
          const uint32_t ha = pal[(va.y >> (8 * k + 1)) & 0x7Fu], hb = pal[(vb.y >> (8 * k + 1)) & 0x7Fu];
          lo |= (la > lb ? la : lb) << (8 * k);
          hi |= (ha > hb ? ha : hb) << (8 * k);
        }
        *reinterpret_cast<uint2*>(&s_row[warp][lane * 2]) = make_uint2(lo, hi);
      }
      __syncwarp();
      if (lane < MN_IMG / 4) {
        const uint8_t* r = reinterpret_cast<const uint8_t*>(s_row[warp]);
        s_out[y * (MN_IMG / 4) + lane] = uint32_t(r[x0]) | (uint32_t(r[x1]) << 8) | (uint32_t(r[x2]) << 16) | (uint32_t(r[x3]) << 24);
      }
    } else {
#pragma unroll
      for (int j = 0; j < MN_SCREEN_W / 32; ++j) {
        const int x = lane + 32 * j;
        s_row[warp][x] = __vmaxu4(s_pal[ra[x] >> 1], s_pal[rb[x] >> 1]);
      }
      __syncwarp();
      if (lane < MN_IMG / 4) {
        const uint32_t p0 = s_row[warp][x0], p1 = s_row[warp][x1], p2 = s_row[warp][x2], p3 = s_row[warp][x3];
        uint32_t* o = s_out + y * (3 * MN_IMG / 4) + 3 * lane;   // bytes R0 G0 B0 R1 | G1 B1 R2 G2 | B2 R3 G3 B3
        o[0] = p0 | (p1 << 24);
        o[1] = (p1 >> 8) | (p2 << 16);
        o[2] = (p2 >> 16) | (p3 << 8);
      }
    }
    __syncwarp();
  }
  __syncthreads();
  uint4* dst = reinterpret_cast<uint4*>(plane);
  const uint4* src = reinterpret_cast<const uint4*>(s_out);
  for (int i = tid; i < MN_PLANE * D / 16; i += 256) dst[i] = src[i];
  __syncthreads();
}

// K3 over a work list: one block per (game, list entry)
template <int D>
__global__ void __launch_bounds__(256) k_push_frames(PoolDev p, int in) {
  // block -> env through the per-game list regions: blockIdx.x is an env-id-like position
  const int pos = blockIdx.x;
  int gi = 0;
  while (gi + 1 < p.n_games && pos >= p.games[gi + 1].env0) ++gi;
  const int k = pos - p.games[gi].env0;
  if (k >= p.counts[in * MN_MAX_GAMES + gi]) return;
  const int e = p.lists[size_t(in) * p.n_envs + pos];
  const uint8_t info = p.push_info[e];
  const uint8_t* fb = p.frames + size_t(e) * (2 * MN_FRAME_BYTES);
  uint8_t* plane = p.ring + (size_t(e) * MN_STACK + (info & 3)) * (MN_PLANE * D);
  preprocess_plane<D>(fb, fb + MN_FRAME_BYTES, info >> 2, plane);
}

// K3 stand-alone (mn_preprocess)
template <int D>
__global__ void __launch_bounds__(256) k_preprocess(const uint8_t* __restrict__ frames, uint8_t* __restrict__ planes, int n) {
  for (int e = blockIdx.x; e < n; e += gridDim.x) {
    const uint8_t* fb = frames + size_t(e) * (2 * MN_FRAME_BYTES);
    preprocess_plane<D>(fb, fb + MN_FRAME_BYTES, 0, planes + size_t(e) * (MN_PLANE * D));
  }
}

// ring -> stacked observation (environment.py:73-76): states[e][y][x][d*4 + k] = ring[(head + k) & 3][y][x][d],
// 4 pixels per thread, 16-byte stores.  Also publishes rewards / terminals.
// `hist_slot` >= 0: also the learner's observation history (PAACLearner.update_memory, paac.py:79-83): the new state
// goes into ring slot `hist_slot` of the env's H-deep history; an env whose episode ended in this step has its whole
// history zeroed AFTER that -- newest entry included -- exactly as paac.py:200-201 does.
//
// With a host copy registered (mn_set_host_states) every 16-byte store is made twice: into `states` in HBM and, through the
// mapped pinned allocation, straight into the caller's host array -- so the states of environments that are done early
// in a macro step cross PCIe while the remaining FiGAR rounds still run (k_emit_early) instead of in one 462 MB copy
// at the end of the step.
template <int D>
__device__ __forceinline__ void emit_env(const PoolDev& p, int e, int publish, int hist_slot) {
  const int head = p.env[e].ring_head;
  const uint8_t* ring = p.ring + size_t(e) * MN_STACK * (MN_PLANE * D);
  uint4* out = reinterpret_cast<uint4*>(p.states + size_t(e) * (MN_PLANE * D * MN_STACK));
  uint4* hostout = p.states_host ? reinterpret_cast<uint4*>(p.states_host + size_t(e) * (MN_PLANE * D * MN_STACK)) : nullptr;
  const bool hist = hist_slot >= 0 && p.history != nullptr;
  const bool wipe = hist && publish && p.over[e];
  const size_t state_q = size_t(MN_PLANE) * D * MN_STACK / 16;   // uint4 per stacked state
  uint4* hout = hist ? reinterpret_cast<uint4*>(p.history) + (size_t(e) * p.history_depth + hist_slot) * state_q : nullptr;
  for (int q = threadIdx.x; q < MN_PLANE / 4; q += blockDim.x) {
    uint32_t in[MN_STACK][D];   // [k][word]: 4 pixels x D bytes of ring plane (head + k)
#pragma unroll
    for (int k = 0; k < MN_STACK; ++k) {
      const uint32_t* src = reinterpret_cast<const uint32_t*>(ring + size_t((head + k) & 3) * (MN_PLANE * D)) + q * D;
#pragma unroll
      for (int d = 0; d < D; ++d) in[k][d] = src[d];
    }
    if (D == 1) {
      // 4x4 byte transpose: out word j = pixel j's 4 time steps
      const uint32_t a = in[0][0], b = in[1][0], c2 = in[2][0], d2 = in[3][0];
      const uint32_t ab_lo = __byte_perm(a, b, 0x5140), ab_hi = __byte_perm(a, b, 0x7362);   // a0 b0 a1 b1 | a2 b2 a3 b3
      const uint32_t cd_lo = __byte_perm(c2, d2, 0x5140), cd_hi = __byte_perm(c2, d2, 0x7362);
      uint4 o;
      o.x = __byte_perm(ab_lo, cd_lo, 0x5410); o.y = __byte_perm(ab_lo, cd_lo, 0x7632);
      o.z = __byte_perm(ab_hi, cd_hi, 0x5410); o.w = __byte_perm(ab_hi, cd_hi, 0x7632);
      out[q] = o;
      if (hostout) hostout[q] = o;
      if (hist && !wipe) hout[q] = o;
    } else {
      // 4 pixels x (3 colours x 4 steps) = 48 bytes = 3 x uint4
      uint8_t bytes[MN_STACK][12];
#pragma unroll
      for (int k = 0; k < MN_STACK; ++k)
#pragma unroll
        for (int i = 0; i < 12; ++i) bytes[k][i] = uint8_t(in[k][i >> 2] >> (8 * (i & 3)));
      uint32_t w[12];
#pragma unroll
      for (int px = 0; px < 4; ++px)
#pragma unroll
        for (int d = 0; d < 3; ++d)
          w[px * 3 + d] = uint32_t(bytes[0][px * 3 + d]) | (uint32_t(bytes[1][px * 3 + d]) << 8) |
                          (uint32_t(bytes[2][px * 3 + d]) << 16) | (uint32_t(bytes[3][px * 3 + d]) << 24);
      const uint4 o0 = make_uint4(w[0], w[1], w[2], w[3]), o1 = make_uint4(w[4], w[5], w[6], w[7]), o2 = make_uint4(w[8], w[9], w[10], w[11]);
      out[3 * q + 0] = o0; out[3 * q + 1] = o1; out[3 * q + 2] = o2;
      if (hostout) { hostout[3 * q + 0] = o0; hostout[3 * q + 1] = o1; hostout[3 * q + 2] = o2; }
      if (hist && !wipe) { hout[3 * q + 0] = o0; hout[3 * q + 1] = o1; hout[3 * q + 2] = o2; }
    }
  }
  if (wipe) {
    uint4* all = reinterpret_cast<uint4*>(p.history) + size_t(e) * p.history_depth * state_q;
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    for (size_t i = threadIdx.x; i < state_q * p.history_depth; i += blockDim.x) all[i] = z;
  }
  if (publish && threadIdx.x == 0) {
    p.rewards[e] = float(p.reward_acc[e]);
    p.terminals[e] = p.over[e] ? 1.0f : 0.0f;
  }
}

// An env whose repeats ran out in FiGAR round r < `last_tag` - 1 without a terminal (k_round left rep_left = -(r + 1)) is
// final from then on -- no later round and no reset touches it -- and is published by k_emit_early<r>; k_emit publishes
// the rest at the end of the step (last round's, and the terminal ones after their get_initial_state()).
__device__ __forceinline__ bool emitted_early(const PoolDev& p, int e, int last_tag) {
  const int left = p.rep_left[e];
  return !p.over[e] && left < 0 && -left < last_tag;
}
// one block per env of [env_lo, env_hi); `last_tag` = 0: every env, else only those k_emit_early did not publish
template <int D>
__global__ void __launch_bounds__(256) k_emit(PoolDev p, int env_lo, int env_hi, int publish, int hist_slot, int last_tag) {
  const int e = env_lo + blockIdx.x;
  if (e >= env_hi) return;
  if (last_tag > 0 && emitted_early(p, e, last_tag)) return;
  emit_env<D>(p, e, publish, hist_slot);
}
// runs on the pool's side stream beside the next FiGAR round: the envs that finished in the round tagged `tag`
template <int D>
__global__ void __launch_bounds__(256) k_emit_early(PoolDev p, int tag, int hist_slot) {
  for (int e = blockIdx.x; e < p.n_envs; e += gridDim.x) {
    if (p.over[e] || p.rep_left[e] != -tag) continue;   // block-uniform
    emit_env<D>(p, e, 1, hist_slot);
  }
}

// the history ring in the reference's order (oldest -> newest): out (N, H, 84, 84, 4D)
__global__ void __launch_bounds__(256) k_history_gather(const uint8_t* __restrict__ hist, uint8_t* __restrict__ out, int n, int depth,
                                                        int head, size_t state_q) {
  const int e = blockIdx.x / depth, j = blockIdx.x % depth;
  if (e >= n) return;
  const int slot = (head + 1 + j) % depth;
  const uint4* src = reinterpret_cast<const uint4*>(hist) + (size_t(e) * depth + slot) * state_q;
  uint4* dst = reinterpret_cast<uint4*>(out) + (size_t(e) * depth + j) * state_q;
  for (size_t i = threadIdx.x; i < state_q; i += blockDim.x) dst[i] = src[i];
}

// ----------------------------------------------------------------------------- K6: rollout bookkeeping
// paac.py:173-205 for one local step t of every environment: clipped reward and mask rows of the rollout, per-env
// episode accumulators, the action x repetition histogram, and -- in ENVIRONMENT ORDER, like the reference's Python
// loop -- the log and running statistics of the episodes that ended in this step.  One block: the work is a few
// bytes per env; what matters is that the order of the log and of the double sums is fixed.
struct RolloutDev {
  int32_t n, T, A, K;
  float *rewards, *masks;            // (T,N)
  int32_t *actions, *repetitions;    // (T,N)
  double* ep_reward;                 // (N,)
  int32_t* ep_steps;                 // (N,)
  float* actions_sum;                // (N,A)
  unsigned long long* action_rep;    // (A,K)
  double* stats;                     // count, sum reward, sum length, min reward, max reward, global_step
  float* fin_reward; int32_t* fin_steps; int32_t* fin_count;   // (N,) (N,) (1,): the episodes that ended in the last recorded step
  int32_t tab_rep[32];
};
__global__ void __launch_bounds__(1024) k_rollout_record(RolloutDev r, int t, const float* __restrict__ rewards,
                                                         const float* __restrict__ terminals, const int32_t* __restrict__ a_idx,
                                                         const int32_t* __restrict__ r_idx, int clip) {
  __shared__ int s_warp_cnt[32];
  __shared__ int s_base;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();
  for (int e0 = 0; e0 < r.n; e0 += blockDim.x) {
    const int e = e0 + threadIdx.x;
    bool over = false;
    double tot = 0.0; int steps = 0;
    if (e < r.n) {
      const float raw = rewards[e];
      over = terminals[e] != 0.0f;
      int a = a_idx[e], k = r_idx[e];
      a = a < 0 ? 0 : (a >= r.A ? r.A - 1 : a);
      k = k < 0 ? 0 : (k >= r.K ? r.K - 1 : k);
      tot = r.ep_reward[e] + double(raw);                                     // total_episode_rewards[e] += actual_reward
      float c = raw;
      if (clip) c = c > 1.0f ? 1.0f : (c < -1.0f ? -1.0f : c);                // rescale_reward (actor_learner.py:108-114)
      const size_t i = size_t(t) * r.n + e;
      r.rewards[i] = c;
      r.masks[i] = 1.0f - (over ? 1.0f : 0.0f);                               // paac.py:176
      r.actions[i] = a; r.repetitions[i] = k;
      steps = r.ep_steps[e] + r.tab_rep[k] + 1;                               // emulator_steps[e] += tab_rep[argmax] + 1
      atomicAdd(&r.action_rep[size_t(a) * r.K + k], 1ull);                    // total_action_rep[a][r] += 1 (integer: order-free)
      float* as = r.actions_sum + size_t(e) * r.A;
      if (over) { for (int j = 0; j < r.A; ++j) as[j] = 0.f; }                // actions_sum[e] = zeros (after += new_actions)
      else as[a] += 1.0f;
      r.ep_reward[e] = over ? 0.0 : tot;
      r.ep_steps[e] = over ? 0 : steps;
    }
    // finished episodes of this chunk, compacted in env order
    const unsigned bal = __ballot_sync(0xFFFFFFFFu, over);
    if (lane == 0) s_warp_cnt[warp] = __popc(bal);
    __syncthreads();
    int before = s_base;
    for (int w = 0; w < warp; ++w) before += s_warp_cnt[w];
    if (over) {
      const int pos = before + __popc(bal & ((1u << lane) - 1u));
      r.fin_reward[pos] = float(tot); r.fin_steps[pos] = steps;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int add = 0;
      for (int w = 0; w < nwarps; ++w) add += s_warp_cnt[w];
      // running statistics in env order (fixed order of the double sums)
      double cnt = r.stats[0], sr = r.stats[1], sl = r.stats[2], mn = r.stats[3], mx = r.stats[4];
      for (int j = 0; j < add; ++j) {
        const int pos = s_base + j;
        const double v = double(r.fin_reward[pos]);
        sr += v; sl += double(r.fin_steps[pos]);
        mn = (cnt == 0.0 || v < mn) ? v : mn; mx = (cnt == 0.0 || v > mx) ? v : mx;
        cnt += 1.0;
      }
      r.stats[0] = cnt; r.stats[1] = sr; r.stats[2] = sl; r.stats[3] = mn; r.stats[4] = mx;
      s_base += add;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) { *r.fin_count = s_base; r.stats[5] += double(r.n); }   // self.global_step += 1 per env
}

// ----------------------------------------------------------------------------- K4: FiGAR sampling
struct Philox { uint32_t v[4]; };
__device__ __forceinline__ Philox philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  Philox o; o.v[0] = c0; o.v[1] = c1; o.v[2] = c2; o.v[3] = c3;
  return o;
}
__device__ __forceinline__ float u01(uint32_t x) { return float(x >> 8) * (1.0f / 16777216.0f); }

// exploration_policy.py:108-116: np.random.multinomial(1, p - epsneg) is an inverse-CDF draw
__device__ __forceinline__ int draw_multinomial(const float* p, int k, float u) {
  const float eps = 5.9604645e-08f;   // np.finfo(np.float32).epsneg
  float cum = 0.f;
  for (int j = 0; j < k; ++j) {
    cum = __fadd_rn(cum, __fsub_rn(p[j], eps));
    if (u < cum) return j;
  }
  return k - 1;
}
__device__ __forceinline__ int draw_argmax(const float* p, int k) {
  int best = 0; float bv = p[0];
  for (int j = 1; j < k; ++j) if (p[j] > bv) { bv = p[j]; best = j; }
  return best;
}
__device__ __forceinline__ int draw_egreedy(const float* p, int k, float u_test, float u_pick, float eps) {
  if (u_test < eps) { int i = int(u_pick * float(k)); return i < k ? i : k - 1; }
  return draw_argmax(p, k);
}
__global__ void k_sample_figar(const float* __restrict__ pi, const float* __restrict__ rho, int n, int a, int k, int mode,
                               float eps, uint32_t seed_lo, uint32_t seed_hi, uint32_t step, int32_t* a_idx, int32_t* r_idx,
                               float* a_hot, float* r_hot) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const Philox x = philox4x32_10(uint32_t(e), step, 0u, 0u, seed_lo, seed_hi);
  const float* pa = pi + size_t(e) * a;
  const float* pr = rho + size_t(e) * k;
  int ia, ir;
  if (mode == 0) { ia = draw_multinomial(pa, a, u01(x.v[0])); ir = draw_multinomial(pr, k, u01(x.v[1])); }
  else if (mode == 1) { ia = draw_egreedy(pa, a, u01(x.v[0]), u01(x.v[2]), eps); ir = draw_egreedy(pr, k, u01(x.v[1]), u01(x.v[3]), eps); }
  else { ia = draw_argmax(pa, a); ir = draw_argmax(pr, k); }
  if (a_idx) a_idx[e] = ia;
  if (r_idx) r_idx[e] = ir;
  if (a_hot) for (int j = 0; j < a; ++j) a_hot[size_t(e) * a + j] = (j == ia) ? 1.f : 0.f;
  if (r_hot) for (int j = 0; j < k; ++j) r_hot[size_t(e) * k + j] = (j == ir) ? 1.f : 0.f;
}

// ----------------------------------------------------------------------------- K5: n-step returns
// paac.py:176,180,226-231 + actor_learner.py:108-114.  The reference runs this recursion in float64
// and feeds float32; so does this kernel (T sequential steps per env, one thread per env).
__global__ void k_nstep(const float* __restrict__ rewards, const float* __restrict__ terminals, const float* __restrict__ values,
                        const float* __restrict__ bootstrap, double gamma, int clip, int t_max, int n, float* __restrict__ y,
                        float* __restrict__ adv) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  double R = double(bootstrap[e]);
  for (int t = t_max - 1; t >= 0; --t) {
    const size_t i = size_t(t) * n + e;
    double r = double(rewards[i]);
    if (clip) r = r > 1.0 ? 1.0 : (r < -1.0 ? -1.0 : r);
    const double mask = double(1.0f - terminals[i]);
    R = r + gamma * R * mask;
    y[i] = float(R);
    adv[i] = float(R - double(values[i]));
  }
}

}  // namespace mn

// ============================================================================= host side / C ABI
using namespace mn;

static thread_local std::string g_err;
static int fail(const std::string& m) { g_err = m; return -1; }
#define CU(x) do { cudaError_t _e = (x); if (_e != cudaSuccess) return fail(std::string(#x) + ": " + cudaGetErrorString(_e)); } while (0)

struct mn_pool {
  PoolDev d;
  int device;
  int max_rep;
  size_t round_smem;
  int round_grid;
  cudaEvent_t done;
  bool pending;
  // early publication (k_emit_early): a side stream, one event per FiGAR round, the join
  cudaStream_t side;
  cudaEvent_t ev_round[32], ev_side;
  int early_emit, early_grid;
  int64_t launches;
  std::vector<void*> allocs;
  uint8_t* pin_ram;
  bool memo_warm;            // the reset memo has been warmed up (first mn_reset_all)
  int warm_passes;           // passes the warm-up takes (0 = off)
  int* err_host;             // pinned: the device error word as of the last completed step (copied before `done` is recorded)
  std::vector<int> env_game_host;
  GameDev games_host[MN_MAX_GAMES];
  Tables* tables_dev;
  int hist_head;             // ring slot of the newest state of the observation history
  // optional per-kernel timing (mn_profile_begin / mn_profile_end)
  bool profiling;
  std::vector<cudaEvent_t> ev_pool;
  size_t ev_used;
  std::vector<int> ev_kind;   // kind of the launch bracketed by events (2i, 2i+1)
};
enum { PK_ROUND = 0, PK_PUSH = 1, PK_EMIT = 2, PK_OTHER = 3, PK_COUNT = 4 };

static void prof_mark(mn_pool* h, int kind, cudaStream_t st, bool begin) {
  if (!h->profiling) return;
  if (h->ev_used == h->ev_pool.size()) { cudaEvent_t e; cudaEventCreate(&e); h->ev_pool.push_back(e); }
  cudaEventRecord(h->ev_pool[h->ev_used++], st);
  if (begin) h->ev_kind.push_back(kind);
}

static bool g_const_ready[64] = {false};

static int upload_constants(int device) {
  if (device >= 0 && device < 64 && g_const_ready[device]) return 0;
  uint8_t gray[128], rgb[384];
  host_palette(gray, rgb);
  uint32_t pal[128];
  for (int i = 0; i < 128; ++i) pal[i] = gray[i] | (uint32_t(rgb[3 * i]) << 8) | (uint32_t(rgb[3 * i + 1]) << 16) | (uint32_t(rgb[3 * i + 2]) << 24);
  CU(cudaMemcpyToSymbol(c_pal, pal, sizeof(pal)));
  if (device >= 0 && device < 64) g_const_ready[device] = true;
  return 0;
}

template <typename T>
static int dev_alloc(mn_pool* h, T** out, size_t count) {
  void* ptr = nullptr;
  cudaError_t e = cudaMalloc(&ptr, count * sizeof(T) > 0 ? count * sizeof(T) : 16);
  if (e != cudaSuccess) return fail(std::string("cudaMalloc: ") + cudaGetErrorString(e));
  e = cudaMemset(ptr, 0, count * sizeof(T));
  if (e != cudaSuccess) return fail(std::string("cudaMemset: ") + cudaGetErrorString(e));
  h->allocs.push_back(ptr);
  *out = static_cast<T*>(ptr);
  return 0;
}

extern "C" {

const char* mn_last_error(void) { return g_err.c_str(); }

// -DMN_CHECK builds only: the first recorded violation {code, value, value} (code 0 = none); -1 in a normal build
int mn_check_report(unsigned int* out3) {
#ifdef MN_CHECK
  unsigned int v[8];
  if (cudaMemcpyFromSymbol(v, g_mn_check, sizeof(v)) != cudaSuccess) return fail("mn_check_report: cudaMemcpyFromSymbol");
  out3[0] = v[0]; out3[1] = v[1]; out3[2] = v[2];
  return 0;
#else
  (void)out3; return -1;
#endif
}

// MN_DIAG=2 runs: {lane-clocks waiting for a free buffer, waiting for blocking hand-offs, total lane-clocks, hand-offs}
int mn_diag_counters(mn_handle h, unsigned long long* out4) {
  if (!h || !out4) return fail("mn_diag_counters: null argument");
  CU(cudaMemcpy(out4, h->d.diag_out, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  return 0;
}

int mn_palette(uint8_t* gray128_host, uint8_t* rgb128x3_host) { host_palette(gray128_host, rgb128x3_host); return 0; }

int mn_start_noops(uint32_t seed, uint32_t global_env, uint32_t episode) { return int(start_noops(seed, global_env, episode)); }

int mn_destroy(mn_handle h) {
  if (!h) return 0;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  for (void* p : h->allocs) cudaFree(p);
  for (cudaEvent_t e : h->ev_pool) cudaEventDestroy(e);
  if (h->done) cudaEventDestroy(h->done);
  if (h->ev_side) cudaEventDestroy(h->ev_side);
  for (cudaEvent_t e : h->ev_round) if (e) cudaEventDestroy(e);
  if (h->side) cudaStreamDestroy(h->side);
  if (h->err_host) cudaFreeHost(h->err_host);
  delete h;
  return 0;
}

int mn_create(const mn_config* cfg, mn_handle* out) {
  if (!cfg || !out) return fail("mn_create: null argument");
  if (cfg->n_games < 1 || cfg->n_games > MN_MAX_GAMES) return fail("mn_create: n_games must be 1..16");
  if (cfg->nb_choices < 0 || cfg->nb_choices > 32 || (cfg->nb_choices > 0 && !cfg->tab_rep)) return fail("mn_create: nb_choices must be 0..32 (with a tab_rep when > 0)");
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev == 0) return fail("mn_create: no CUDA device (the product has no CPU fallback)");
  if (cfg->device < 0 || cfg->device >= ndev) return fail("mn_create: bad device ordinal");
  CU(cudaSetDevice(cfg->device));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, cfg->device));
  if (prop.major < 10) return fail("mn_create: this library is built for sm_100a (B200) only");
  if (upload_constants(cfg->device)) return -1;

  mn_pool* h = new mn_pool();
  memset(&h->d, 0, sizeof(h->d));
  h->device = cfg->device; h->done = nullptr; h->pending = false; h->launches = 0; h->pin_ram = nullptr; h->err_host = nullptr;
  h->profiling = false; h->ev_used = 0; h->hist_head = 0;
  PoolDev& d = h->d;
  int n = 0, max_actions = 0;
  size_t rom_total = 0, max_rom = 0;
  for (int g = 0; g < cfg->n_games; ++g) {
    const mn_game& mg = cfg->games[g];
    if (!mg.rom || mg.rom_size < 2048 || mg.rom_size > 16384 || mg.n_envs < 1) { delete h; return fail("mn_create: bad game entry (rom 2K..16K, n_envs >= 1)"); }
    GameDev& G = d.games[g];
    G.game_id = game_id_from_name(mg.name ? mg.name : "");
    const GameEntry& ge = game_db(G.game_id);
    G.cart = detect_cart(mg.rom, size_t(mg.rom_size));
    G.ctrl = ge.ctrl;
    G.n_actions = ge.n_actions;
    for (int i = 0; i < 18; ++i) G.actions[i] = ge.actions[i];
    G.rom_off = int(rom_total); G.rom_size = mg.rom_size;
    rom_total += (size_t(mg.rom_size) + 15) & ~size_t(15);
    if (size_t(mg.rom_size) > max_rom) max_rom = mg.rom_size;
    if (max_rom < 4096) max_rom = 4096;   // 2K images are staged twice (k_round)
    G.env0 = n; G.n_envs = mg.n_envs;
    n += mg.n_envs;
    if (G.n_actions > max_actions) max_actions = G.n_actions;
  }
  d.n_games = cfg->n_games; d.n_envs = n; d.depth = cfg->rgb ? 3 : 1; d.num_actions = max_actions;
  d.nb_choices = cfg->nb_choices > 0 ? cfg->nb_choices : 1;
  d.single_life = cfg->single_life_episodes; d.random_start = cfg->random_start; d.seed = cfg->random_seed; d.env_id_offset = cfg->env_id_offset;
  d.draw_all_frames = cfg->draw_all_frames;
  d.sync_slack = 4;
  if (const char* ev = getenv("MN_SYNC_SLACK")) d.sync_slack = atoi(ev) < 0 ? 0x3FFFFFFF : atoi(ev);
  d.diag = 0;
  if (const char* ev = getenv("MN_DIAG")) d.diag = atoi(ev);
  d.fifo_high = MN_FIFO_HIGH;
  if (const char* ev = getenv("MN_FIFO_HIGH")) { const int v = atoi(ev); if (v >= 1 && v < MN_FIFO_CAP) d.fifo_high = v; }
  h->max_rep = 0;
  for (int i = 0; i < cfg->nb_choices; ++i) {
    if (cfg->tab_rep[i] < 0 || cfg->tab_rep[i] > 1000) { delete h; return fail("mn_create: tab_rep entries must be 0..1000"); }
    d.tab_rep[i] = cfg->tab_rep[i];
    if (cfg->tab_rep[i] > h->max_rep) h->max_rep = cfg->tab_rep[i];
  }
  // env slots per warp.  Measured on B200 (profiles/): the emulation loop is bound by the latency of ONE warp's
  // instruction stream, and a second 6502 warp on an SM sub-partition slows the first one down more than it adds.
  // So: the fewest lanes per warp for which the whole pool still fits one block (4 6502 warps + their 4 picture-side
  // partners) per SM -- thin warps while the pool is small, and e.g. 28 lanes on 147 of the 148 SMs for 16,384
  // environments of one game.  Beyond 32 x 4 x SMs environments the warps are full and blocks queue up.
  int slots = cfg->envs_per_warp;
  if (slots <= 0) {
    slots = 32;
    for (int sl = 1; sl <= 32; ++sl) {
      int blocks = 0;
      for (int g = 0; g < d.n_games; ++g) blocks += (d.games[g].n_envs + MN_WARPS_PER_BLOCK * sl - 1) / (MN_WARPS_PER_BLOCK * sl);
      if (blocks <= prop.multiProcessorCount) { slots = sl; break; }
    }
  }
  if (slots < 1 || slots > 32) { delete h; return fail("mn_create: envs_per_warp must be 1..32"); }
  d.slots = slots;
  int blk = 0;
  for (int g = 0; g < d.n_games; ++g) {
    GameDev& G = d.games[g];
    G.blk0 = blk;
    G.n_blks = (G.n_envs + MN_WARPS_PER_BLOCK * slots - 1) / (MN_WARPS_PER_BLOCK * slots);
    blk += G.n_blks;
  }
  h->round_grid = blk;
  h->round_smem = ((max_rom + 15) & ~size_t(15)) + sizeof(Tables) +
                  size_t(MN_WARPS_PER_BLOCK) * slots * (MN_CORE_WORDS * 4 + MN_RAM_PITCH + MN_FIFO_WORDS * 4);
  if (h->round_smem > size_t(prop.sharedMemPerBlockOptin)) { delete h; return fail("mn_create: shared memory budget exceeded"); }
  // A pool that fits one block per SM must GET one block per SM: below half of the SM's shared memory two blocks fit one
  // SM, and whenever the block scheduler finds some SMs busy at launch (k_emit_early runs beside the rounds) it doubles
  // blocks up on the others -- measured: every round 36 % slower, whatever the size of the other kernel.  Asking for
  // more than half makes the placement independent of what else is resident.
  if (blk <= prop.multiProcessorCount) {
    const size_t more_than_half = size_t(prop.sharedMemPerMultiprocessor) / 2 + 1024;
    if (h->round_smem < more_than_half && more_than_half <= size_t(prop.sharedMemPerBlockOptin)) h->round_smem = more_than_half;
  }
  // the attribute belongs to the function, not to this pool: several pools with different needs may coexist
  CU(cudaFuncSetAttribute(k_round<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(prop.sharedMemPerBlockOptin)));
  CU(cudaFuncSetAttribute(k_round<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(prop.sharedMemPerBlockOptin)));
  // the side-stream kernel shares SMs with k_round blocks: same carve-out, so no SM has to drain to be reconfigured
  CU(cudaFuncSetAttribute(k_emit_early<1>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  CU(cudaFuncSetAttribute(k_emit_early<3>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));

  const size_t N = size_t(n), D = size_t(d.depth);
  uint8_t* roms = nullptr;
  int rc = 0;
  rc |= dev_alloc(h, &roms, rom_total);
  rc |= dev_alloc(h, &d.env, N);
  rc |= dev_alloc(h, &d.ram, N * 128);
  rc |= dev_alloc(h, &d.frames, N * 2 * MN_FRAME_BYTES);
  rc |= dev_alloc(h, &d.ring, N * MN_STACK * MN_PLANE * D);
  rc |= dev_alloc(h, &d.states, N * MN_STACK * MN_PLANE * D);
  rc |= dev_alloc(h, &d.rewards, N);
  rc |= dev_alloc(h, &d.terminals, N);
  rc |= dev_alloc(h, &d.actions, N * size_t(d.num_actions));
  rc |= dev_alloc(h, &d.repetitions, N * size_t(32));
  rc |= dev_alloc(h, &d.action_idx, N);
  rc |= dev_alloc(h, &d.repetition_idx, N);
  rc |= dev_alloc(h, &d.next_calls, N);
  rc |= dev_alloc(h, &d.cur_action, N);
  rc |= dev_alloc(h, &d.rep_left, N);
  rc |= dev_alloc(h, &d.reward_acc, N);
  rc |= dev_alloc(h, &d.over, N);
  rc |= dev_alloc(h, &d.push_info, N);
  rc |= dev_alloc(h, &d.episode, N);
  rc |= dev_alloc(h, &d.env_game, N);
  rc |= dev_alloc(h, &d.lists, 4 * N);
  rc |= dev_alloc(h, &d.counts, size_t(4 * MN_MAX_GAMES));
  rc |= dev_alloc(h, &d.error, size_t(1));
  rc |= dev_alloc(h, &d.total_next, size_t(1));
  rc |= dev_alloc(h, &d.redo_count, size_t(1));
  rc |= dev_alloc(h, &d.diag_out, size_t(8));
  rc |= dev_alloc(h, &d.warm_env, size_t(d.n_games) * MN_TIMER_SEEDS);
  rc |= dev_alloc(h, &d.warm_ram, size_t(d.n_games) * MN_TIMER_SEEDS * 128);
  rc |= dev_alloc(h, &d.warm_episode, size_t(d.n_games) * MN_TIMER_SEEDS);
  rc |= dev_alloc(h, &d.total_instr, size_t(1));
  // random_start: the start no-op count (0..30) joins the key, so a bucket needs room for 31 counts (x RAM-dependence
  // variants): 40 slots instead of 8 (gray: 75 x 40 x 96 KB = 290 MB per game)
  d.memo_enabled = (cfg->no_reset_memo == 0) ? 1 : 0;
  d.memo_slots = cfg->random_start ? 40 : MN_MEMO_SLOTS;
  // warm-up: on when every game lends at least 38 envs (at most two extra reset passes at creation); MN_WARM_MEMO=0/1 overrides
  h->memo_warm = false; h->warm_passes = 0;
  if (d.memo_enabled && !cfg->random_start) {   // (under random_start the memo fills as episodes end: 75 x 31 keys per game)
    int min_envs = 1 << 30;
    for (int g = 0; g < d.n_games; ++g) if (d.games[g].n_envs < min_envs) min_envs = d.games[g].n_envs;
    const int W = min_envs < MN_TIMER_SEEDS ? min_envs : MN_TIMER_SEEDS;
    const int passes = (MN_TIMER_SEEDS + W - 1) / W;
    bool on = passes <= 2;
    if (const char* ev = getenv("MN_WARM_MEMO")) on = atoi(ev) != 0;
    h->warm_passes = on ? passes : 0;
  }
  d.memo_entry_bytes = int32_t(memo_hdr_bytes() + 2 * MN_FRAME_BYTES + size_t(MN_STACK) * MN_PLANE * D);
  rc |= dev_alloc(h, &d.memo, d.memo_enabled ? size_t(d.n_games) * MN_TIMER_SEEDS * size_t(d.memo_slots) * size_t(d.memo_entry_bytes) : size_t(16));
  rc |= dev_alloc(h, &d.reset_rnd, N);
  rc |= dev_alloc(h, &d.pre_ram, N * 128);
  rc |= dev_alloc(h, &d.memo_hit, N);
  rc |= dev_alloc(h, &d.memo_busy, size_t(d.n_games) * MN_TIMER_SEEDS);
  d.memo1_entry_bytes = int32_t(memo_hdr_bytes() + 2 * MN_FRAME_BYTES);
  rc |= dev_alloc(h, &d.memo1, d.memo_enabled ? size_t(d.n_games) * MN_TIMER_SEEDS * size_t(d.memo_slots) * size_t(d.memo1_entry_bytes) : size_t(16));
  rc |= dev_alloc(h, &d.memo1_busy, size_t(d.n_games) * MN_TIMER_SEEDS);
  rc |= dev_alloc(h, &d.track, N * 5);
  rc |= dev_alloc(h, &d.memo_stats, size_t(4));
  rc |= dev_alloc(h, &h->tables_dev, size_t(1));
  d.history_depth = cfg->history > 0 ? cfg->history : 0;
  if (d.history_depth > 64) { mn_destroy(h); return fail("mn_create: history must be 0..64"); }
  if (d.history_depth) rc |= dev_alloc(h, &d.history, N * size_t(d.history_depth) * MN_STACK * MN_PLANE * D);
  if (rc) { mn_destroy(h); return -1; }
  d.roms = roms;
  d.tables = h->tables_dev;
  {
    Tables t;
    build_tables(&t);
    CU(cudaMemcpy(h->tables_dev, &t, sizeof(t), cudaMemcpyHostToDevice));
    std::vector<uint8_t> eg(N);
    h->env_game_host.resize(N);
    for (int g = 0; g < d.n_games; ++g) {
      CU(cudaMemcpy(roms + d.games[g].rom_off, cfg->games[g].rom, size_t(cfg->games[g].rom_size), cudaMemcpyHostToDevice));
      for (int i = 0; i < d.games[g].n_envs; ++i) { eg[size_t(d.games[g].env0 + i)] = uint8_t(g); h->env_game_host[size_t(d.games[g].env0 + i)] = g; }
    }
    CU(cudaMemcpy(d.env_game, eg.data(), N, cudaMemcpyHostToDevice));
  }
  memcpy(h->games_host, d.games, sizeof(d.games));
  CU(cudaEventCreateWithFlags(&h->done, cudaEventDisableTiming));
  CU(cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking));
  CU(cudaEventCreateWithFlags(&h->ev_side, cudaEventDisableTiming));
  for (cudaEvent_t& e : h->ev_round) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  h->early_emit = 1;
  h->early_grid = 16;   // measured (e2e, 16,384 envs, k frames/s): 8 / 16 / 32 / 74 / 148 / 296 blocks = 747 / 748 / 740 / 738 / 722 / 714 (off: 705): PCIe-bound, must not crowd the rounds
  if (const char* ev = getenv("MN_EARLY_EMIT")) h->early_emit = atoi(ev) != 0;
  if (const char* ev = getenv("MN_EARLY_GRID")) { const int v = atoi(ev); if (v >= 1 && v <= 65535) h->early_grid = v; }
  CU(cudaMallocHost(&h->err_host, sizeof(int)));
  *h->err_host = 0;
  // construct every AtariEmulator (atari_emulator.py:18-31): seed, RAM garbage, loadROM's reset
  {
    const int tb = 256, gb = (n + tb - 1) / tb > 1 ? (n + tb - 1) / tb : 1;
    k_fill_list<<<gb, tb>>>(d, 0);
    k_round<false><<<h->round_grid, MN_THREADS, h->round_smem>>>(d, ROUND_POWER_ON, 0, -1);
    h->launches += 2;
    CU(cudaGetLastError());
    CU(cudaDeviceSynchronize());
  }
  *out = h;
  return 0;
}

int mn_get_buffers(mn_handle h, mn_buffers* out) {
  if (!h || !out) return fail("mn_get_buffers: null argument");
  const PoolDev& d = h->d;
  out->n_envs = d.n_envs; out->num_actions = d.num_actions; out->nb_choices = d.nb_choices; out->depth = d.depth;
  out->states = d.states; out->rewards = d.rewards; out->terminals = d.terminals; out->actions = d.actions;
  out->repetitions = d.repetitions; out->action_idx = d.action_idx; out->repetition_idx = d.repetition_idx;
  out->next_calls = d.next_calls; out->frames = d.frames; out->ring = d.ring;
  out->history = d.history; out->history_depth = d.history_depth;
  return 0;
}

int mn_set_tab_rep(mn_handle h, const int* tab_rep, int nb_choices) {
  if (!h || !tab_rep || nb_choices < 1 || nb_choices > 32) return fail("mn_set_tab_rep: nb_choices must be 1..32");
  for (int i = 0; i < nb_choices; ++i) if (tab_rep[i] < 0 || tab_rep[i] > 1000) return fail("mn_set_tab_rep: entries must be 0..1000");
  if (h->pending) return fail("mn_set_tab_rep: a macro step is in flight");
  h->max_rep = 0;
  for (int i = 0; i < nb_choices; ++i) { h->d.tab_rep[i] = tab_rep[i]; if (tab_rep[i] > h->max_rep) h->max_rep = tab_rep[i]; }
  h->d.nb_choices = nb_choices;
  return 0;
}

int mn_legal_actions(mn_handle h, int env, int32_t* out) {
  if (!h || env < 0 || env >= h->d.n_envs) return fail("mn_legal_actions: bad environment id");
  const GameDev& G = h->games_host[h->env_game_host[size_t(env)]];
  if (out) for (int i = 0; i < G.n_actions; ++i) out[i] = G.actions[i];
  return G.n_actions;
}

}  // extern "C"

// ---- launch helpers (host)
static void launch_push(mn_pool* h, int in, cudaStream_t st) {
  prof_mark(h, PK_PUSH, st, true);
  if (h->d.depth == 1) k_push_frames<1><<<h->d.n_envs, 256, 0, st>>>(h->d, in);
  else k_push_frames<3><<<h->d.n_envs, 256, 0, st>>>(h->d, in);
  prof_mark(h, PK_PUSH, st, false);
  h->launches++;
}
static void launch_round(mn_pool* h, int mode, int in, int out, cudaStream_t st, bool track = false) {
  prof_mark(h, PK_ROUND, st, true);
  if (track) k_round<true><<<h->round_grid, MN_THREADS, h->round_smem, st>>>(h->d, mode, in, out);
  else k_round<false><<<h->round_grid, MN_THREADS, h->round_smem, st>>>(h->d, mode, in, out);
  prof_mark(h, PK_ROUND, st, false);
  h->launches++;
}
static void launch_emit(mn_pool* h, int lo, int hi, int publish, cudaStream_t st, int hist_slot = -1, int last_tag = 0) {
  prof_mark(h, PK_EMIT, st, true);
  if (h->d.depth == 1) k_emit<1><<<hi - lo, 256, 0, st>>>(h->d, lo, hi, publish, hist_slot, last_tag);
  else k_emit<3><<<hi - lo, 256, 0, st>>>(h->d, lo, hi, publish, hist_slot, last_tag);
  prof_mark(h, PK_EMIT, st, false);
  h->launches++;
}
// after FiGAR round `r` and its K3 on `st`: publish the envs that finished in it, on the side stream
static int launch_emit_early(mn_pool* h, int r, cudaStream_t st, int hist_slot) {
  CU(cudaEventRecord(h->ev_round[r], st));
  CU(cudaStreamWaitEvent(h->side, h->ev_round[r], 0));
  if (h->d.depth == 1) k_emit_early<1><<<h->early_grid, 256, 0, h->side>>>(h->d, r + 1, hist_slot);
  else k_emit_early<3><<<h->early_grid, 256, 0, h->side>>>(h->d, r + 1, hist_slot);
  h->launches++;
  return 0;
}
// get_initial_state() for the envs on the reset list (list 2) (atari_emulator.py:102-110): memo hits are restored
// by copy, misses are emulated with the RAM-dependence probe and then stored
static void launch_initial_state(mn_pool* h, cudaStream_t st) {
  const int n = h->d.n_envs;
  const bool memo = h->d.memo_enabled != 0;
  prof_mark(h, PK_OTHER, st, true);
  k_clear_counts<<<1, 32, 0, st>>>(h->d, 0);
  k_clear_counts<<<1, 32, 0, st>>>(h->d, 1);
  k_clear_counts<<<1, 32, 0, st>>>(h->d, 3);
  k_reset_prepare<<<(n + 255) / 256, 256, 0, st>>>(h->d);
  if (memo) {
    if (h->d.depth == 1) k_reset_restore<1><<<n, 256, 0, st>>>(h->d); else k_reset_restore<3><<<n, 256, 0, st>>>(h->d);
    h->launches++;
  }
  prof_mark(h, PK_OTHER, st, false);
  h->launches += 4;
  launch_round(h, ROUND_RESET, 1, -1, st, memo);
  if (memo) {   // level 1: what the reset unit left (stored for the misses, restored for the level-1 hits, which join list 1)
    prof_mark(h, PK_OTHER, st, true);
    if (h->d.depth == 1) k_memo_insert<1, 1><<<n, 256, 0, st>>>(h->d); else k_memo_insert<3, 1><<<n, 256, 0, st>>>(h->d);
    k_reset_restore_l1<<<n, 256, 0, st>>>(h->d);
    prof_mark(h, PK_OTHER, st, false);
    h->launches += 2;
  }
  for (int i = 0; i < MN_STACK; ++i) {
    launch_round(h, ROUND_INITIAL, 1, (i == MN_STACK - 1) ? -1 : -2, st, memo);
    launch_push(h, 1, st);
  }
  if (memo) {
    prof_mark(h, PK_OTHER, st, true);
    if (h->d.depth == 1) k_memo_insert<1, 2><<<n, 256, 0, st>>>(h->d); else k_memo_insert<3, 2><<<n, 256, 0, st>>>(h->d);
    prof_mark(h, PK_OTHER, st, false);
    h->launches++;
  }
}

extern "C" {

int mn_reset_all(mn_handle h, void* stream) {
  if (!h) return fail("mn_reset_all: null handle");
  CU(cudaSetDevice(h->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int n = h->d.n_envs, tb = 256, gb = (n + tb - 1) / tb;
  if (!h->memo_warm && h->warm_passes > 0) {   // once: all 75 timer seeds of every game into the reset memo
    for (int pass = 0; pass < h->warm_passes; ++pass) {
      k_warm_begin<<<h->d.n_games, 96, 0, st>>>(h->d, pass);
      launch_round(h, ROUND_RESET, 1, -1, st, true);
      for (int i = 0; i < MN_STACK; ++i) {
        launch_round(h, ROUND_INITIAL, 1, (i == MN_STACK - 1) ? -1 : -2, st, true);
        launch_push(h, 1, st);
      }
      if (h->d.depth == 1) k_memo_insert<1, 2><<<n, 256, 0, st>>>(h->d); else k_memo_insert<3, 2><<<n, 256, 0, st>>>(h->d);
      k_warm_end<<<h->d.n_games, 96, 0, st>>>(h->d, pass);
      h->launches += 3;
    }
    h->memo_warm = true;
  }
  k_fill_list<<<gb, tb, 0, st>>>(h->d, 2);
  h->launches++;
  launch_initial_state(h, st);
  if (h->d.history) {   // paac.py:107-112: memory = zeros, memory[e, -1] = shared_states[e]
    CU(cudaMemsetAsync(h->d.history, 0, size_t(n) * h->d.history_depth * MN_STACK * MN_PLANE * h->d.depth, st));
    h->hist_head = 0;
  }
  // over[] is zero here (k_fill_list), so nothing is wiped
  launch_emit(h, 0, n, 1, st, h->d.history ? h->hist_head : -1);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(h->err_host, h->d.error, sizeof(int), cudaMemcpyDeviceToHost, st));
  CU(cudaEventRecord(h->done, st));
  h->pending = true;
  return 0;
}

int mn_step_async(mn_handle h, int use_indices, void* stream) {
  if (!h) return fail("mn_step_async: null handle");
  CU(cudaSetDevice(h->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int n = h->d.n_envs, tb = 256, gb = (n + tb - 1) / tb;
  k_begin_step<<<gb, tb, 0, st>>>(h->d, use_indices);
  h->launches++;
  if (h->d.history) h->hist_head = (h->hist_head + 1) % h->d.history_depth;   // update_memory: shift, newest last
  const int hist_slot = h->d.history ? h->hist_head : -1;
  const bool early = h->early_emit && h->max_rep > 0 && h->max_rep <= 32;
  int in = 0;
  for (int r = 0; r <= h->max_rep; ++r) {
    const int out = in ^ 1;
    launch_round(h, ROUND_FIGAR, in, out, st);
    launch_push(h, in, st);
    if (r < h->max_rep) { k_clear_counts<<<1, 32, 0, st>>>(h->d, in); h->launches++; }
    if (early && r < h->max_rep && launch_emit_early(h, r, st, hist_slot)) return -1;
    in = out;
  }
  launch_initial_state(h, st);
  if (early) {   // the rest, once the side stream is through
    CU(cudaEventRecord(h->ev_side, h->side));
    CU(cudaStreamWaitEvent(st, h->ev_side, 0));
  }
  launch_emit(h, 0, n, 1, st, hist_slot, early ? h->max_rep + 1 : 0);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(h->err_host, h->d.error, sizeof(int), cudaMemcpyDeviceToHost, st));
  CU(cudaEventRecord(h->done, st));
  h->pending = true;
  return 0;
}

int mn_set_host_states(mn_handle h, uint8_t* states_pinned_host) {
  if (!h) return fail("mn_set_host_states: null handle");
  CU(cudaSetDevice(h->device));
  if (h->pending) { CU(cudaEventSynchronize(h->done)); }   // no step in flight may still write through the old pointer
  if (!states_pinned_host) { h->d.states_host = nullptr; return 0; }
  void* dev = nullptr;
  if (cudaHostGetDevicePointer(&dev, states_pinned_host, 0) != cudaSuccess || !dev) {
    cudaGetLastError();
    return fail("mn_set_host_states: not page-locked, device-mapped host memory (cudaHostAlloc / cudaHostRegister / torch pin_memory)");
  }
  h->d.states_host = static_cast<uint8_t*>(dev);
  return 0;
}

int mn_wait(mn_handle h) {
  if (!h) return fail("mn_wait: null handle");
  CU(cudaSetDevice(h->device));
  if (h->pending) { CU(cudaEventSynchronize(h->done)); h->pending = false; }
  CU(cudaGetLastError());
  const int err = *h->err_host;   // travelled with the step: no blocking copy here
  if (err == 2) return fail("internal error: a 6502 lane gave up waiting for its picture-side partner (hand-off protocol)");
  if (err) return fail("episode over right after reset ('This should never happen.', atari_emulator.py:108-109)");
  return 0;
}

int mn_step_host(mn_handle h, const float* actions_host, const float* repetitions_host, uint8_t* states_host,
                 float* rewards_host, float* terminals_host, void* stream) {
  if (!h) return fail("mn_step_host: null handle");
  CU(cudaSetDevice(h->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const PoolDev& d = h->d;
  const size_t N = size_t(d.n_envs);
  CU(cudaMemcpyAsync(d.actions, actions_host, N * d.num_actions * sizeof(float), cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(d.repetitions, repetitions_host, N * d.nb_choices * sizeof(float), cudaMemcpyHostToDevice, st));
  if (mn_step_async(h, 0, stream)) return -1;
  CU(cudaMemcpyAsync(states_host, d.states, N * MN_STACK * MN_PLANE * d.depth, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(rewards_host, d.rewards, N * sizeof(float), cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(terminals_host, d.terminals, N * sizeof(float), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  return mn_wait(h);
}

int mn_env_reset(mn_handle h, int env, void* stream) {
  if (!h || env < 0 || env >= h->d.n_envs) return fail("mn_env_reset: bad environment id");
  CU(cudaSetDevice(h->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  k_single_list<<<1, 64, 0, st>>>(h->d, 2, env, 0);
  h->launches++;
  launch_initial_state(h, st);
  launch_emit(h, env, env + 1, 1, st);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(h->err_host, h->d.error, sizeof(int), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  return mn_wait(h);
}

int mn_env_next(mn_handle h, int env, int action_index, float* reward, int* terminal, void* stream) {
  if (!h || env < 0 || env >= h->d.n_envs) return fail("mn_env_next: bad environment id");
  const GameDev& G = h->games_host[h->env_game_host[size_t(env)]];
  if (action_index < 0 || action_index >= G.n_actions) return fail("mn_env_next: action index outside the legal action set");
  CU(cudaSetDevice(h->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  k_single_list<<<1, 64, 0, st>>>(h->d, 0, env, int(G.actions[action_index]));
  h->launches++;
  launch_round(h, ROUND_SINGLE, 0, 1, st);
  launch_push(h, 0, st);
  launch_emit(h, env, env + 1, 1, st);
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(st));
  float r = 0.f, t = 0.f;
  CU(cudaMemcpy(&r, h->d.rewards + env, sizeof(float), cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(&t, h->d.terminals + env, sizeof(float), cudaMemcpyDeviceToHost));
  if (reward) *reward = r;
  if (terminal) *terminal = (t != 0.f);
  return 0;
}

int mn_get_ram(mn_handle h, int env, uint8_t* out) {
  if (!h || env < 0 || env >= h->d.n_envs) return fail("mn_get_ram: bad environment id");
  CU(cudaSetDevice(h->device));
  CU(cudaMemcpy(out, h->d.ram + size_t(env) * 128, 128, cudaMemcpyDeviceToHost));
  return 0;
}

int mn_get_screen(mn_handle h, int env, uint8_t* out) {
  if (!h || env < 0 || env >= h->d.n_envs) return fail("mn_get_screen: bad environment id");
  CU(cudaSetDevice(h->device));
  EnvState s;
  CU(cudaMemcpy(&s, h->d.env + env, sizeof(s), cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(out, h->d.frames + size_t(env) * 2 * MN_FRAME_BYTES + ((s.pflags & F_CURFB) ? MN_FRAME_BYTES : 0), MN_FRAME_BYTES,
                cudaMemcpyDeviceToHost));
  return 0;
}

int mn_get_cpu_state(mn_handle h, int env, int32_t* out) {
  if (!h || env < 0 || env >= h->d.n_envs) return fail("mn_get_cpu_state: bad environment id");
  CU(cudaSetDevice(h->device));
  EnvState s;
  CU(cudaMemcpy(&s, h->d.env + env, sizeof(s), cudaMemcpyDeviceToHost));
  out[0] = s.A; out[1] = s.X; out[2] = s.Y; out[3] = s.SP; out[4] = s.PC; out[5] = int32_t(pack_ps(s)); out[6] = s.cycles;
  out[7] = (s.cycles * 3 - s.clk_frame_start) / 228; out[8] = s.bank; out[9] = s.timer;
  return 0;
}

int mn_get_lives(mn_handle h, int env, int* lives, int* game_over, int* frame_number) {
  if (!h || env < 0 || env >= h->d.n_envs) return fail("mn_get_lives: bad environment id");
  CU(cudaSetDevice(h->device));
  EnvState s;
  CU(cudaMemcpy(&s, h->d.env + env, sizeof(s), cudaMemcpyDeviceToHost));
  if (lives) *lives = s.lives;
  if (game_over) *game_over = (s.flags & F_TERMINAL) ? 1 : 0;
  if (frame_number) *frame_number = s.frame_number;
  return 0;
}

int mn_total_next_calls(mn_handle h, int64_t* out) {
  if (!h || !out) return fail("mn_total_next_calls: null argument");
  CU(cudaSetDevice(h->device));
  unsigned long long v = 0;
  CU(cudaMemcpy(&v, h->d.total_next, sizeof(v), cudaMemcpyDeviceToHost));
  *out = int64_t(v);
  return 0;
}

int mn_profile_begin(mn_handle h) {
  if (!h) return fail("mn_profile_begin: null handle");
  h->profiling = true; h->ev_used = 0; h->ev_kind.clear();
  return 0;
}

int mn_profile_end(mn_handle h, double* ms_by_kind4, int64_t* launches_by_kind4) {
  if (!h) return fail("mn_profile_end: null handle");
  CU(cudaSetDevice(h->device));
  CU(cudaDeviceSynchronize());
  for (int k = 0; k < PK_COUNT; ++k) { ms_by_kind4[k] = 0.0; launches_by_kind4[k] = 0; }
  for (size_t i = 0; i < h->ev_kind.size(); ++i) {
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, h->ev_pool[2 * i], h->ev_pool[2 * i + 1]));
    ms_by_kind4[h->ev_kind[i]] += double(ms);
    launches_by_kind4[h->ev_kind[i]] += 1;
  }
  h->profiling = false; h->ev_used = 0; h->ev_kind.clear();
  return 0;
}

int mn_redo_count(mn_handle h, int64_t* out) {
  if (!h || !out) return fail("mn_redo_count: null argument");
  CU(cudaSetDevice(h->device));
  unsigned long long v = 0;
  CU(cudaMemcpy(&v, h->d.redo_count, sizeof(v), cudaMemcpyDeviceToHost));
  *out = int64_t(v);
  return 0;
}

int mn_memo_stats(mn_handle h, int64_t* out3) {
  if (!h || !out3) return fail("mn_memo_stats: null argument");
  CU(cudaSetDevice(h->device));
  unsigned long long v[4] = {0, 0, 0, 0};
  CU(cudaMemcpy(v, h->d.memo_stats, sizeof(v), cudaMemcpyDeviceToHost));
  out3[0] = int64_t(v[0]); out3[1] = int64_t(v[1]); out3[2] = int64_t(v[2]);
  return 0;
}

int mn_memo_level1_hits(mn_handle h, int64_t* out) {
  if (!h || !out) return fail("mn_memo_level1_hits: null argument");
  CU(cudaSetDevice(h->device));
  unsigned long long v = 0;
  CU(cudaMemcpy(&v, h->d.memo_stats + 3, sizeof(v), cudaMemcpyDeviceToHost));
  *out = int64_t(v);
  return 0;
}

int mn_total_instructions(mn_handle h, int64_t* out) {
  if (!h || !out) return fail("mn_total_instructions: null argument");
  CU(cudaSetDevice(h->device));
  unsigned long long v = 0;
  CU(cudaMemcpy(&v, h->d.total_instr, sizeof(v), cudaMemcpyDeviceToHost));
  *out = int64_t(v);
  return 0;
}

int mn_launch_count(mn_handle h, int64_t* out) {
  if (!h || !out) return fail("mn_launch_count: null argument");
  *out = h->launches;
  return 0;
}

int mn_preprocess(const uint8_t* frames_dev, uint8_t* planes_dev, int n, int rgb, void* stream) {
  if (n <= 0) return 0;
  int dev = 0;
  CU(cudaGetDevice(&dev));
  if (upload_constants(dev)) return -1;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = n < 148 * 16 ? n : 148 * 16;
  if (rgb) k_preprocess<3><<<grid, 256, 0, st>>>(frames_dev, planes_dev, n);
  else k_preprocess<1><<<grid, 256, 0, st>>>(frames_dev, planes_dev, n);
  CU(cudaGetLastError());
  return 0;
}

int mn_sample_figar(const float* pi_dev, const float* rho_dev, int n, int a, int k, int mode, float epsilon, uint64_t seed,
                    uint32_t step, int32_t* action_idx_dev, int32_t* rep_idx_dev, float* action_onehot_dev,
                    float* rep_onehot_dev, void* stream) {
  if (n <= 0) return 0;
  if (a < 1 || k < 1 || mode < 0 || mode > 2) return fail("mn_sample_figar: bad shape or mode");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  k_sample_figar<<<(n + 255) / 256, 256, 0, st>>>(pi_dev, rho_dev, n, a, k, mode, epsilon, uint32_t(seed), uint32_t(seed >> 32),
                                                   step, action_idx_dev, rep_idx_dev, action_onehot_dev, rep_onehot_dev);
  CU(cudaGetLastError());
  return 0;
}

int mn_nstep(const float* rewards_dev, const float* terminals_dev, const float* values_dev, const float* bootstrap_dev,
             double gamma, int clip, int t, int n, float* y_dev, float* adv_dev, void* stream) {
  if (n <= 0 || t <= 0) return 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  k_nstep<<<(n + 255) / 256, 256, 0, st>>>(rewards_dev, terminals_dev, values_dev, bootstrap_dev, gamma, clip, t, n, y_dev, adv_dev);
  CU(cudaGetLastError());
  return 0;
}

int mn_history_head(mn_handle h, int* head) {
  if (!h || !head) return fail("mn_history_head: null argument");
  if (!h->d.history) return fail("mn_history_head: the pool was created without an observation history");
  *head = h->hist_head;
  return 0;
}

int mn_history_gather(mn_handle h, uint8_t* out_dev, void* stream) {
  if (!h || !out_dev) return fail("mn_history_gather: null argument");
  if (!h->d.history) return fail("mn_history_gather: the pool was created without an observation history");
  CU(cudaSetDevice(h->device));
  const size_t state_q = size_t(MN_PLANE) * h->d.depth * MN_STACK / 16;
  k_history_gather<<<h->d.n_envs * h->d.history_depth, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      h->d.history, out_dev, h->d.n_envs, h->d.history_depth, h->hist_head, state_q);
  h->launches++;
  CU(cudaGetLastError());
  return 0;
}

}  // extern "C"

// ---- K6: rollout bookkeeping (paac.py:140-205)
struct mn_rollout {
  RolloutDev d;
  int device;
  std::vector<void*> allocs;
};
template <typename T>
static int ro_alloc(mn_rollout* r, T** out, size_t count) {
  void* ptr = nullptr;
  cudaError_t e = cudaMalloc(&ptr, count * sizeof(T) > 0 ? count * sizeof(T) : 16);
  if (e != cudaSuccess) return fail(std::string("cudaMalloc: ") + cudaGetErrorString(e));
  e = cudaMemset(ptr, 0, count * sizeof(T));
  if (e != cudaSuccess) return fail(std::string("cudaMemset: ") + cudaGetErrorString(e));
  r->allocs.push_back(ptr);
  *out = static_cast<T*>(ptr);
  return 0;
}

extern "C" {

int mn_rollout_destroy(mn_rollout_handle r) {
  if (!r) return 0;
  cudaSetDevice(r->device);
  cudaDeviceSynchronize();
  for (void* p : r->allocs) cudaFree(p);
  delete r;
  return 0;
}

int mn_rollout_create(int device, int n_envs, int max_local_steps, int num_actions, int nb_choices, const int* tab_rep,
                      mn_rollout_handle* out) {
  if (!out || !tab_rep) return fail("mn_rollout_create: null argument");
  if (n_envs < 1 || max_local_steps < 1 || num_actions < 1 || nb_choices < 1 || nb_choices > 32)
    return fail("mn_rollout_create: bad shape (n_envs, max_local_steps, num_actions >= 1, nb_choices 1..32)");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail("mn_rollout_create: no CUDA device (the product has no CPU fallback)");
  if (device < 0 || device >= ndev) return fail("mn_rollout_create: bad device ordinal");
  CU(cudaSetDevice(device));
  mn_rollout* r = new mn_rollout();
  memset(&r->d, 0, sizeof(r->d));
  r->device = device;
  RolloutDev& d = r->d;
  d.n = n_envs; d.T = max_local_steps; d.A = num_actions; d.K = nb_choices;
  for (int i = 0; i < nb_choices; ++i) d.tab_rep[i] = tab_rep[i];
  const size_t N = size_t(n_envs), T = size_t(max_local_steps);
  int rc = 0;
  rc |= ro_alloc(r, &d.rewards, T * N);
  rc |= ro_alloc(r, &d.masks, T * N);
  rc |= ro_alloc(r, &d.actions, T * N);
  rc |= ro_alloc(r, &d.repetitions, T * N);
  rc |= ro_alloc(r, &d.ep_reward, N);
  rc |= ro_alloc(r, &d.ep_steps, N);
  rc |= ro_alloc(r, &d.actions_sum, N * size_t(num_actions));
  rc |= ro_alloc(r, &d.action_rep, size_t(num_actions) * nb_choices);
  rc |= ro_alloc(r, &d.stats, size_t(6));
  rc |= ro_alloc(r, &d.fin_reward, N);
  rc |= ro_alloc(r, &d.fin_steps, N);
  rc |= ro_alloc(r, &d.fin_count, size_t(1));
  if (rc) { mn_rollout_destroy(r); return -1; }
  *out = r;
  return 0;
}

int mn_rollout_get_buffers(mn_rollout_handle r, mn_rollout_buffers* out) {
  if (!r || !out) return fail("mn_rollout_get_buffers: null argument");
  const RolloutDev& d = r->d;
  out->n_envs = d.n; out->max_local_steps = d.T; out->num_actions = d.A; out->nb_choices = d.K;
  out->rewards = d.rewards; out->masks = d.masks; out->actions = d.actions; out->repetitions = d.repetitions;
  out->episode_reward = d.ep_reward; out->episode_steps = d.ep_steps; out->actions_sum = d.actions_sum;
  out->action_rep = reinterpret_cast<uint64_t*>(d.action_rep); out->stats = d.stats;
  out->finished_reward = d.fin_reward; out->finished_steps = d.fin_steps; out->finished_count = d.fin_count;
  return 0;
}

int mn_rollout_begin(mn_rollout_handle r, void* stream) {
  if (!r) return fail("mn_rollout_begin: null handle");
  CU(cudaSetDevice(r->device));
  CU(cudaMemsetAsync(r->d.action_rep, 0, size_t(r->d.A) * r->d.K * sizeof(unsigned long long), static_cast<cudaStream_t>(stream)));
  return 0;
}

int mn_rollout_record(mn_rollout_handle r, int t, const float* rewards_dev, const float* terminals_dev,
                      const int32_t* action_idx_dev, const int32_t* rep_idx_dev, int clip, void* stream) {
  if (!r || !rewards_dev || !terminals_dev || !action_idx_dev || !rep_idx_dev) return fail("mn_rollout_record: null argument");
  if (t < 0 || t >= r->d.T) return fail("mn_rollout_record: t outside 0..max_local_steps-1");
  CU(cudaSetDevice(r->device));
  k_rollout_record<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(r->d, t, rewards_dev, terminals_dev, action_idx_dev, rep_idx_dev, clip);
  CU(cudaGetLastError());
  return 0;
}

}  // extern "C"
