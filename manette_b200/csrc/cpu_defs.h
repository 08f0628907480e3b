// manette_b200 -- 6502 decode definitions shared by the emulator core and its table builder:
// addressing modes, operation ids, the control word of the table-driven datapath, and the decode entry of
// every opcode as a COMPILE-TIME constant (decode_entry).  The source of truth is the opcode matrix below
// (mnemonic + addressing mode + base cycles), the same information the reference's emulator dependency
// (ALE / Stella M6502Low) keeps in its instruction tables; undocumented opcodes included.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define MN_HD __host__ __device__
#define MN_NOINLINE __noinline__
#define MN_NOINLINE_DEV __forceinline__
#define MN_INLINE __forceinline__
#else
#define MN_HD
#define MN_NOINLINE
#define MN_NOINLINE_DEV inline
#define MN_INLINE inline
#endif

namespace mn {

// addressing modes / operation classes of the packed decode descriptor
enum { AM_IMP = 0, AM_ACC, AM_IMM, AM_ZP, AM_ZPX, AM_ZPY, AM_ABS, AM_ABX, AM_ABY, AM_IZX, AM_IZY, AM_REL, AM_IND };
enum { OC_NONE = 0, OC_READ = 1, OC_WRITE = 2, OC_RMW = 3 };   // what the operate phase needs from memory
enum {  // operations
  O_NOP = 0, O_ORA, O_AND, O_EOR, O_ADC, O_SBC, O_CMP, O_CPX, O_CPY, O_BIT, O_LDA, O_LDX, O_LDY, O_LAX, O_LXA, O_ANC,
  O_ALR, O_ARR, O_XAA, O_AXS, O_LAS,                                   // read class
  O_STA, O_STX, O_STY, O_SAX, O_AHX, O_SHY, O_SHX, O_TAS,              // write class
  O_ASL, O_LSR, O_ROL, O_ROR, O_INC, O_DEC, O_SLO, O_RLA, O_SRE, O_RRA, O_DCP, O_ISC,   // rmw class
  O_BRANCH, O_JMP, O_JSR, O_RTS, O_RTI, O_BRK, O_PHA, O_PHP, O_PLA, O_PLP,
  O_TAX, O_TAY, O_TXA, O_TYA, O_TSX, O_TXS, O_INX, O_INY, O_DEX, O_DEY, O_FLAG, O_KIL };
// descriptor: [3:0] mode  [5:4] class  [11:6] op  [15:12] base cycles (2..8)
// branches keep their condition in the aux byte of the table entry; flag ops likewise.
#define MN_DESC(mode, cls, op, cyc) uint32_t((mode) | ((cls) << 4) | ((op) << 6) | ((cyc) << 12))

// One 16-byte entry per opcode, read with a single shared-memory load:
//   k  control word of the table-driven datapath (K_* fields, see cpu_step)
//   d  [15:0] descriptor (MN_DESC, cycles in 4 bits)  [23:16] aux (branch: [7:6]=flag selector 0 N,1 V,2 C,3 Z,
//      [0]=wanted value ; flag op: [7:1]=bit index in P, [0]=set)  [31:24] D_* phase flags
//   x  operand mask: 0xFF for the zero-page forms, 0xFFFF for the absolute ones (0 when there is no operand address)
struct alignas(16) TabEnt { uint32_t k, d, x, dm; };
enum : uint32_t {
  D_EA = 1u << 24       /* has an effective address (mode >= zp, not relative) */,
  D_INDIRECT = 1u << 25 /* (zp,X) (zp),Y (abs) */,
  D_READ = 1u << 26     /* read phase: read / read-modify-write class with an effective address */,
  D_WRITE = 1u << 27    /* write phase: write class, or read-modify-write on memory */,
  D_PAGEPEN = 1u << 28  /* read class, indexed: +1 cycle when the index crosses a page */,
  D_BRANCH = 1u << 29 };
// Decode entry of the FAST TICK (cpu_fast in emu_core.cuh): the same four words (for the opcodes the fast tick covers
// that the general datapath does not -- flag ops, JMP, BIT, JSR/RTS, PHA/PLA/PHP/PLP -- k/x/dm are programmed so that
// the datapath does the right thing, or nothing) plus `f`:
//   [7:0]  bits of P a flag op rewrites   [15:8] their new value   [18:16] stack pointer change + 2
//   [20:19] next PC: 0 sequential / branch, 1 effective address (JMP, JSR), 2 pulled pair + 1 (RTS)
//   FX_* bits below; pad0: branch test (mask over the flag word | invert << 16), zero for everything else
struct alignas(16) FastCompact { uint32_t k, d, x, dm, f, pad0, pad1, pad2; };
enum : uint32_t {
  FX_SPD = 16, FX_PCS = 19,
  FX_PULL = 1u << 21    /* the read phase reads 0x100 | (SP + 1) */,
  FX_PUSH = 1u << 22    /* the write phase writes 0x100 | SP */,
  FX_PAIR_STACK = 1u << 23 /* the byte pair is pulled from the stack (RTS) instead of read through a zero-page pointer */,
  FX_BIT = 1u << 24, FX_W_RET = 1u << 25 /* write value = return address (JSR: high byte, then low) */,
  FX_W_PS = 1u << 26    /* write value = packed status | B (PHP) */,
  FX_PLP = 1u << 27, FX_VALID = 1u << 28, FX_PAIR = 1u << 29 /* reads a byte pair from RIOT RAM: (zp,X) (zp),Y RTS */,
  FX_PAIR_X = 1u << 30  /* pointer address = operand + X (zp,X) */, FX_PUSH2 = 1u << 31 /* second push (JSR) */ };
// What cpu_fast actually loads: the compact entry spread over 24 words (six 16-byte shared-memory loads), every field
// in the exact form the instruction stream consumes it -- byte-permute selectors ready made, masks as whole words,
// the stack-pointer change as the word to add to the packed register file -- so that no field costs a shift-and-mask
// on the serial path of the one warp that runs it (profiles/: that path is bound by its instruction count).
struct alignas(16) FastEnt {
  uint32_t sel_a, sel_b, sel_idx, sel_fn;        // operand / index / function selectors (perm8)
  uint32_t sel_cout, xm, binv, pm;               // carry-out selector, operand mask, inversion of operand b, P bits rewritten
  uint32_t dm, spd, cyc, seqinc;                 // register-file mask of the result, SP change << 24, base cycles, length
  uint32_t cmask, cconst, rotmask, nzmask;       // carry in = (C & cmask) | cconst; rotate-in = C & rotmask; nz <- result?
  uint32_t g, bm_nz, bm_p, pclr;                 // G_* flags; branch test over nz / over P; flag op: bits of P cleared ...
  uint32_t pset, sel_pb, sel_padd, penbit;       // ... and set; the byte pair's address = (byte sel_pb) + (byte sel_padd);
                                                 // penbit: 0xFF00 for indexed reads (+1 cycle across a page) | 0xC0 for BIT (N, V <- M)
};
enum : uint32_t {
  G_VALID = 1u << 0, G_NEEDPAIR = 1u << 1 /* both bytes of the pair must lie in RIOT RAM */, G_PTR = 1u << 2 /* base = the pair */,
  G_RA_P0 = 1u << 3 /* read address = 0x100 | pair address (PLA) */, G_READ = 1u << 4, G_WRITE = 1u << 5,
  G_PUSH = 1u << 6 /* write address = 0x100 | SP */, G_PUSH2 = 1u << 7 /* second push (JSR) */, G_W_RET = 1u << 8,
  G_BIT = 1u << 9, G_INV = 1u << 10 /* branch test inverted */, G_PC_EA = 1u << 11, G_PC_PAIR = 1u << 12,
  G_DECIMAL = 1u << 13, G_PAGEPEN = 1u << 14 };
struct Tables {          // read-only, staged in shared memory by the kernels
  TabEnt e[256];
  FastEnt f[256];
};
static_assert(sizeof(TabEnt) == 16 && sizeof(FastEnt) == 96, "fast_entry() addresses Tables::f at byte 4096 + 96 * opcode");

// Control word of the table-driven datapath (TabEnt::k), see cpu_exec.  Operand selectors are byte-permute
// selectors over the 8 bytes {A, X, Y, SP | M, 0x01, 0xFF, 0x00}, the function selector one over the bytes
// {sum, or, and, xor | left, right}: choosing is one instruction, never a branch.
enum : uint32_t {
  K_ASEL = 0 /* 3 bits */, K_BSEL = 3 /* 3 bits */, K_CSEL = 6 /* 2 bits: carry in = 0, 1, or the C flag (2) */,
  K_FN = 8 /* 3 bits */, K_ROT = 1u << 11 /* shifts rotate through the C flag */, K_NZ = 1u << 12, K_GENERIC = 1u << 15,
  K_DECIMAL = 1u << 16, K_ISEL = 17 /* 3 bits, index register selector */, K_LEN = 20 /* 2 bits, length - 1 */,
  K_CSRC = 22 /* 2 bits: carry out comes from the sum (0), the left shift (1), the right shift (2) */ };
enum { FN_ADD = 0, FN_OR, FN_AND, FN_EOR, FN_LEFT, FN_RIGHT };
enum { SEL_A = 0, SEL_X, SEL_Y, SEL_SP, SEL_M, SEL_ONE, SEL_FF, SEL_ZERO };
// TabEnt::x  [15:0] operand mask  [23:16] inversion mask of the second ALU operand (0xFF: SBC and the compares)
//            [31:24] the bits of P the datapath rewrites (C 0x01, V 0x40)
// TabEnt::dm byte mask of the packed register file (A | X << 8 | Y << 16 | SP << 24) that takes the result

// 16 x 16 opcode matrix, row = high nibble
static constexpr const char* kOpcodeMatrix[256] = {
    "BRK imp 7", "ORA izx 6", "KIL imp 2", "SLO izx 8", "NOP zp 3", "ORA zp 3", "ASL zp 5", "SLO zp 5",
    "PHP imp 3", "ORA imm 2", "ASL acc 2", "ANC imm 2", "NOP abs 4", "ORA abs 4", "ASL abs 6", "SLO abs 6",
    "BPL rel 2", "ORA izy 5", "KIL imp 2", "SLO izy 8", "NOP zpx 4", "ORA zpx 4", "ASL zpx 6", "SLO zpx 6",
    "CLC imp 2", "ORA aby 4", "NOP imp 2", "SLO aby 7", "NOP abx 4", "ORA abx 4", "ASL abx 7", "SLO abx 7",
    "JSR abs 6", "AND izx 6", "KIL imp 2", "RLA izx 8", "BIT zp 3", "AND zp 3", "ROL zp 5", "RLA zp 5",
    "PLP imp 4", "AND imm 2", "ROL acc 2", "ANC imm 2", "BIT abs 4", "AND abs 4", "ROL abs 6", "RLA abs 6",
    "BMI rel 2", "AND izy 5", "KIL imp 2", "RLA izy 8", "NOP zpx 4", "AND zpx 4", "ROL zpx 6", "RLA zpx 6",
    "SEC imp 2", "AND aby 4", "NOP imp 2", "RLA aby 7", "NOP abx 4", "AND abx 4", "ROL abx 7", "RLA abx 7",
    "RTI imp 6", "EOR izx 6", "KIL imp 2", "SRE izx 8", "NOP zp 3", "EOR zp 3", "LSR zp 5", "SRE zp 5",
    "PHA imp 3", "EOR imm 2", "LSR acc 2", "ALR imm 2", "JMP abs 3", "EOR abs 4", "LSR abs 6", "SRE abs 6",
    "BVC rel 2", "EOR izy 5", "KIL imp 2", "SRE izy 8", "NOP zpx 4", "EOR zpx 4", "LSR zpx 6", "SRE zpx 6",
    "CLI imp 2", "EOR aby 4", "NOP imp 2", "SRE aby 7", "NOP abx 4", "EOR abx 4", "LSR abx 7", "SRE abx 7",
    "RTS imp 6", "ADC izx 6", "KIL imp 2", "RRA izx 8", "NOP zp 3", "ADC zp 3", "ROR zp 5", "RRA zp 5",
    "PLA imp 4", "ADC imm 2", "ROR acc 2", "ARR imm 2", "JMP ind 5", "ADC abs 4", "ROR abs 6", "RRA abs 6",
    "BVS rel 2", "ADC izy 5", "KIL imp 2", "RRA izy 8", "NOP zpx 4", "ADC zpx 4", "ROR zpx 6", "RRA zpx 6",
    "SEI imp 2", "ADC aby 4", "NOP imp 2", "RRA aby 7", "NOP abx 4", "ADC abx 4", "ROR abx 7", "RRA abx 7",
    "NOP imm 2", "STA izx 6", "NOP imm 2", "SAX izx 6", "STY zp 3", "STA zp 3", "STX zp 3", "SAX zp 3",
    "DEY imp 2", "NOP imm 2", "TXA imp 2", "XAA imm 2", "STY abs 4", "STA abs 4", "STX abs 4", "SAX abs 4",
    "BCC rel 2", "STA izy 6", "KIL imp 2", "AHX izy 6", "STY zpx 4", "STA zpx 4", "STX zpy 4", "SAX zpy 4",
    "TYA imp 2", "STA aby 5", "TXS imp 2", "TAS aby 5", "SHY abx 5", "STA abx 5", "SHX aby 5", "AHX aby 5",
    "LDY imm 2", "LDA izx 6", "LDX imm 2", "LAX izx 6", "LDY zp 3", "LDA zp 3", "LDX zp 3", "LAX zp 3",
    "TAY imp 2", "LDA imm 2", "TAX imp 2", "LXA imm 2", "LDY abs 4", "LDA abs 4", "LDX abs 4", "LAX abs 4",
    "BCS rel 2", "LDA izy 5", "KIL imp 2", "LAX izy 5", "LDY zpx 4", "LDA zpx 4", "LDX zpy 4", "LAX zpy 4",
    "CLV imp 2", "LDA aby 4", "TSX imp 2", "LAS aby 4", "LDY abx 4", "LDA abx 4", "LDX aby 4", "LAX aby 4",
    "CPY imm 2", "CMP izx 6", "NOP imm 2", "DCP izx 8", "CPY zp 3", "CMP zp 3", "DEC zp 5", "DCP zp 5",
    "INY imp 2", "CMP imm 2", "DEX imp 2", "AXS imm 2", "CPY abs 4", "CMP abs 4", "DEC abs 6", "DCP abs 6",
    "BNE rel 2", "CMP izy 5", "KIL imp 2", "DCP izy 8", "NOP zpx 4", "CMP zpx 4", "DEC zpx 6", "DCP zpx 6",
    "CLD imp 2", "CMP aby 4", "NOP imp 2", "DCP aby 7", "NOP abx 4", "CMP abx 4", "DEC abx 7", "DCP abx 7",
    "CPX imm 2", "SBC izx 6", "NOP imm 2", "ISC izx 8", "CPX zp 3", "SBC zp 3", "INC zp 5", "ISC zp 5",
    "INX imp 2", "SBC imm 2", "NOP imp 2", "SBC imm 2", "CPX abs 4", "SBC abs 4", "INC abs 6", "ISC abs 6",
    "BEQ rel 2", "SBC izy 5", "KIL imp 2", "ISC izy 8", "NOP zpx 4", "SBC zpx 4", "INC zpx 6", "ISC zpx 6",
    "SED imp 2", "SBC aby 4", "NOP imp 2", "ISC aby 7", "NOP abx 4", "SBC abx 4", "INC abx 7", "ISC abx 7",
};


struct MnemonicInfo { const char* name; int op; int cls; int aux; };

static constexpr MnemonicInfo kMnemonics[] = {
    {"NOP", O_NOP, OC_READ, 0},   {"ORA", O_ORA, OC_READ, 0},   {"AND", O_AND, OC_READ, 0},   {"EOR", O_EOR, OC_READ, 0},
    {"ADC", O_ADC, OC_READ, 0},   {"SBC", O_SBC, OC_READ, 0},   {"CMP", O_CMP, OC_READ, 0},   {"CPX", O_CPX, OC_READ, 0},
    {"CPY", O_CPY, OC_READ, 0},   {"BIT", O_BIT, OC_READ, 0},   {"LDA", O_LDA, OC_READ, 0},   {"LDX", O_LDX, OC_READ, 0},
    {"LDY", O_LDY, OC_READ, 0},   {"LAX", O_LAX, OC_READ, 0},   {"LXA", O_LXA, OC_READ, 0},   {"ANC", O_ANC, OC_READ, 0},
    {"ALR", O_ALR, OC_READ, 0},   {"ARR", O_ARR, OC_READ, 0},   {"XAA", O_XAA, OC_READ, 0},   {"AXS", O_AXS, OC_READ, 0},
    {"LAS", O_LAS, OC_READ, 0},
    {"STA", O_STA, OC_WRITE, 0},  {"STX", O_STX, OC_WRITE, 0},  {"STY", O_STY, OC_WRITE, 0},  {"SAX", O_SAX, OC_WRITE, 0},
    {"AHX", O_AHX, OC_WRITE, 0},  {"SHY", O_SHY, OC_WRITE, 0},  {"SHX", O_SHX, OC_WRITE, 0},  {"TAS", O_TAS, OC_WRITE, 0},
    {"ASL", O_ASL, OC_RMW, 0},    {"LSR", O_LSR, OC_RMW, 0},    {"ROL", O_ROL, OC_RMW, 0},    {"ROR", O_ROR, OC_RMW, 0},
    {"INC", O_INC, OC_RMW, 0},    {"DEC", O_DEC, OC_RMW, 0},    {"SLO", O_SLO, OC_RMW, 0},    {"RLA", O_RLA, OC_RMW, 0},
    {"SRE", O_SRE, OC_RMW, 0},    {"RRA", O_RRA, OC_RMW, 0},    {"DCP", O_DCP, OC_RMW, 0},    {"ISC", O_ISC, OC_RMW, 0},
    // branches: aux = selector<<6 | wanted   (0 N, 1 V, 2 C, 3 Z)
    {"BPL", O_BRANCH, OC_NONE, (0 << 6) | 0}, {"BMI", O_BRANCH, OC_NONE, (0 << 6) | 1},
    {"BVC", O_BRANCH, OC_NONE, (1 << 6) | 0}, {"BVS", O_BRANCH, OC_NONE, (1 << 6) | 1},
    {"BCC", O_BRANCH, OC_NONE, (2 << 6) | 0}, {"BCS", O_BRANCH, OC_NONE, (2 << 6) | 1},
    {"BNE", O_BRANCH, OC_NONE, (3 << 6) | 0}, {"BEQ", O_BRANCH, OC_NONE, (3 << 6) | 1},
    {"JMP", O_JMP, OC_NONE, 0},   {"JSR", O_JSR, OC_NONE, 0},   {"RTS", O_RTS, OC_NONE, 0},   {"RTI", O_RTI, OC_NONE, 0},
    {"BRK", O_BRK, OC_NONE, 0},   {"PHA", O_PHA, OC_NONE, 0},   {"PHP", O_PHP, OC_NONE, 0},   {"PLA", O_PLA, OC_NONE, 0},
    {"PLP", O_PLP, OC_NONE, 0},   {"TAX", O_TAX, OC_NONE, 0},   {"TAY", O_TAY, OC_NONE, 0},   {"TXA", O_TXA, OC_NONE, 0},
    {"TYA", O_TYA, OC_NONE, 0},   {"TSX", O_TSX, OC_NONE, 0},   {"TXS", O_TXS, OC_NONE, 0},   {"INX", O_INX, OC_NONE, 0},
    {"INY", O_INY, OC_NONE, 0},   {"DEX", O_DEX, OC_NONE, 0},   {"DEY", O_DEY, OC_NONE, 0},
    // flag ops: aux = bit index in P << 1 | set      (C bit 0, I bit 2, D bit 3, V bit 6)
    {"CLC", O_FLAG, OC_NONE, (0 << 1) | 0}, {"SEC", O_FLAG, OC_NONE, (0 << 1) | 1},
    {"CLI", O_FLAG, OC_NONE, (2 << 1) | 0}, {"SEI", O_FLAG, OC_NONE, (2 << 1) | 1},
    {"CLD", O_FLAG, OC_NONE, (3 << 1) | 0}, {"SED", O_FLAG, OC_NONE, (3 << 1) | 1},
    {"CLV", O_FLAG, OC_NONE, (6 << 1) | 0},
    {"KIL", O_KIL, OC_NONE, 0},
};


static constexpr const char* kModeNames[13] = {"imp", "acc", "imm", "zp", "zpx", "zpy", "abs", "abx", "aby", "izx", "izy", "rel", "ind"};

// Datapath programming of one opcode: which register / constant feeds each ALU input, the function, which
// flags and registers take the result.  Opcodes it cannot express (stack, flow, flag ops, BIT, undocumented
// combinations) are left to cpu_special().
struct Datapath { uint32_t k, binv, pmask, dm; };
enum : uint32_t { DM_A = 0x000000FFu, DM_X = 0x0000FF00u, DM_Y = 0x00FF0000u, DM_SP = 0xFF000000u };
constexpr Datapath dp_make(uint32_t k, uint32_t asel, uint32_t bsel, uint32_t binv, uint32_t csel, uint32_t fn, uint32_t csrc,
                           uint32_t kflags, uint32_t pmask, uint32_t dm) {
  return Datapath{k | K_GENERIC | (asel << K_ASEL) | (bsel << K_BSEL) | (csel << K_CSEL) | (fn << K_FN) | (csrc << K_CSRC) | kflags,
                  binv, pmask, dm};
}
constexpr Datapath datapath_control(int op, int mode) {
  const uint32_t len = (mode <= AM_ACC) ? 1u : (mode == AM_ABS || mode == AM_ABX || mode == AM_ABY || mode == AM_IND) ? 3u : 2u;
  const uint32_t isel = (mode == AM_ZPX || mode == AM_ABX) ? uint32_t(SEL_X) : (mode == AM_ZPY || mode == AM_ABY) ? uint32_t(SEL_Y) : uint32_t(SEL_ZERO);
  const uint32_t k = ((len - 1) << K_LEN) | (isel << K_ISEL);
  const uint32_t sh_src = (mode == AM_ACC) ? SEL_A : SEL_M, sh_dst = (mode == AM_ACC) ? DM_A : 0u;
  const uint32_t Z = SEL_ZERO, NZ = K_NZ, C = 0x01u, V = 0x40u;
  switch (op) {
    //                         a       b        ~b    cin fn       cout  flags        P     registers
    case O_NOP: return dp_make(k, Z,      Z,       0,    0, FN_ADD,   0, 0,           0,    0);
    case O_LDA: return dp_make(k, SEL_M,  Z,       0,    0, FN_ADD,   0, NZ,          0,    DM_A);
    case O_LDX: return dp_make(k, SEL_M,  Z,       0,    0, FN_ADD,   0, NZ,          0,    DM_X);
    case O_LDY: return dp_make(k, SEL_M,  Z,       0,    0, FN_ADD,   0, NZ,          0,    DM_Y);
    case O_LAX: return dp_make(k, SEL_M,  Z,       0,    0, FN_ADD,   0, NZ,          0,    DM_A | DM_X);
    case O_STA: return dp_make(k, SEL_A,  Z,       0,    0, FN_ADD,   0, 0,           0,    0);
    case O_STX: return dp_make(k, SEL_X,  Z,       0,    0, FN_ADD,   0, 0,           0,    0);
    case O_STY: return dp_make(k, SEL_Y,  Z,       0,    0, FN_ADD,   0, 0,           0,    0);
    case O_ORA: return dp_make(k, SEL_A,  SEL_M,   0,    0, FN_OR,    0, NZ,          0,    DM_A);
    case O_AND: return dp_make(k, SEL_A,  SEL_M,   0,    0, FN_AND,   0, NZ,          0,    DM_A);
    case O_EOR: return dp_make(k, SEL_A,  SEL_M,   0,    0, FN_EOR,   0, NZ,          0,    DM_A);
    case O_ADC: return dp_make(k, SEL_A,  SEL_M,   0,    2, FN_ADD,   0, NZ | K_DECIMAL, C | V, DM_A);
    case O_SBC: return dp_make(k, SEL_A,  SEL_M,   0xFF, 2, FN_ADD,   0, NZ | K_DECIMAL, C | V, DM_A);
    case O_CMP: return dp_make(k, SEL_A,  SEL_M,   0xFF, 1, FN_ADD,   0, NZ,          C,    0);
    case O_CPX: return dp_make(k, SEL_X,  SEL_M,   0xFF, 1, FN_ADD,   0, NZ,          C,    0);
    case O_CPY: return dp_make(k, SEL_Y,  SEL_M,   0xFF, 1, FN_ADD,   0, NZ,          C,    0);
    case O_INC: return dp_make(k, SEL_M,  SEL_ONE, 0,    0, FN_ADD,   0, NZ,          0,    0);
    case O_DEC: return dp_make(k, SEL_M,  SEL_FF,  0,    0, FN_ADD,   0, NZ,          0,    0);
    case O_INX: return dp_make(k, SEL_X,  SEL_ONE, 0,    0, FN_ADD,   0, NZ,          0,    DM_X);
    case O_INY: return dp_make(k, SEL_Y,  SEL_ONE, 0,    0, FN_ADD,   0, NZ,          0,    DM_Y);
    case O_DEX: return dp_make(k, SEL_X,  SEL_FF,  0,    0, FN_ADD,   0, NZ,          0,    DM_X);
    case O_DEY: return dp_make(k, SEL_Y,  SEL_FF,  0,    0, FN_ADD,   0, NZ,          0,    DM_Y);
    case O_TAX: return dp_make(k, SEL_A,  Z,       0,    0, FN_ADD,   0, NZ,          0,    DM_X);
    case O_TAY: return dp_make(k, SEL_A,  Z,       0,    0, FN_ADD,   0, NZ,          0,    DM_Y);
    case O_TXA: return dp_make(k, SEL_X,  Z,       0,    0, FN_ADD,   0, NZ,          0,    DM_A);
    case O_TYA: return dp_make(k, SEL_Y,  Z,       0,    0, FN_ADD,   0, NZ,          0,    DM_A);
    case O_TSX: return dp_make(k, SEL_SP, Z,       0,    0, FN_ADD,   0, NZ,          0,    DM_X);
    case O_TXS: return dp_make(k, SEL_X,  Z,       0,    0, FN_ADD,   0, 0,           0,    DM_SP);
    case O_ASL: return dp_make(k, sh_src, Z,       0,    0, FN_LEFT,  1, NZ,          C,    sh_dst);
    case O_LSR: return dp_make(k, sh_src, Z,       0,    0, FN_RIGHT, 2, NZ,          C,    sh_dst);
    case O_ROL: return dp_make(k, sh_src, Z,       0,    0, FN_LEFT,  1, NZ | K_ROT,  C,    sh_dst);
    case O_ROR: return dp_make(k, sh_src, Z,       0,    0, FN_RIGHT, 2, NZ | K_ROT,  C,    sh_dst);
    default: return Datapath{k, 0u, 0u, 0u};
  }
}

constexpr bool txt_eq(const char* a, const char* b, int n) {   // first n characters equal and b ends there
  for (int i = 0; i < n; ++i) if (a[i] != b[i] || b[i] == 0) return false;
  return b[n] == 0;
}
// the decode entry of one opcode, from its row of the matrix
constexpr TabEnt decode_entry(int opc) {
  const char* row = kOpcodeMatrix[opc];
  const char* mode_txt = row + 4;
  int mlen = 0;
  while (mode_txt[mlen] && mode_txt[mlen] != ' ') ++mlen;
  int mode = 0;
  for (int m = 0; m < 13; ++m) if (txt_eq(mode_txt, kModeNames[m], mlen)) mode = m;
  const int cyc = mode_txt[mlen + 1] - '0';
  int mi = 0;
  for (int k = 0; k < int(sizeof(kMnemonics) / sizeof(kMnemonics[0])); ++k) if (txt_eq(row, kMnemonics[k].name, 3)) mi = k;
  const int op = kMnemonics[mi].op;
  int cls = kMnemonics[mi].cls;
  if (op == O_NOP && mode == AM_IMP) cls = OC_NONE;   // only the multi-byte NOPs touch memory
  const bool has_ea = mode >= AM_ZP && mode != AM_REL;
  const bool wide = mode == AM_ABS || mode == AM_ABX || mode == AM_ABY || mode == AM_IND;
  uint32_t d = MN_DESC(uint32_t(mode), uint32_t(cls), uint32_t(op), uint32_t(cyc)) | (uint32_t(kMnemonics[mi].aux & 0xFF) << 16);
  if (has_ea) d |= D_EA;
  if (mode >= AM_IZX && mode != AM_REL) d |= D_INDIRECT;
  if (has_ea && (cls == OC_READ || cls == OC_RMW)) d |= D_READ;
  if (cls == OC_WRITE || (cls == OC_RMW && mode != AM_ACC)) d |= D_WRITE;
  if (has_ea && cls == OC_READ) d |= D_PAGEPEN;
  if (op == O_BRANCH) d |= D_BRANCH;
  const Datapath dp = datapath_control(op, mode);
  TabEnt t = {dp.k, d, (has_ea ? (wide ? 0xFFFFu : 0xFFu) : 0u) | (dp.binv << 16) | (dp.pmask << 24), dp.dm};
  return t;
}

// the fast-tick entry of one opcode (FX_VALID clear: the opcode always takes the general path)
constexpr FastCompact fast_compact_entry(int opc) {
  const TabEnt g = decode_entry(opc);
  const uint32_t mode = g.d & 15u, op = (g.d >> 6) & 63u;
  const uint32_t len1 = (g.k >> K_LEN) & 3u;
  FastCompact t = {g.k, g.d, g.x, g.dm, 2u << FX_SPD, 0u, 0u, 0u};
  const uint32_t lenbits = len1 << K_LEN;
  // what the datapath does for a plain register store / load, borrowed for the stack forms
  const Datapath sta = datapath_control(O_STA, AM_IMP), lda = datapath_control(O_LDA, AM_IMP);
  if (g.k & K_GENERIC) {
    if (mode == AM_IND) return t;
    t.f |= FX_VALID;
    if (g.d & D_INDIRECT) {
      t.f |= FX_PAIR | (mode == AM_IZX ? uint32_t(FX_PAIR_X) : 0u);
      // (zp),Y: the index is added to the pointer that was read, not to the operand
      t.k = (t.k & ~(7u << K_ISEL)) | (uint32_t(mode == AM_IZY ? SEL_Y : SEL_ZERO) << K_ISEL);
      t.x = (t.x & 0xFFFF0000u) | 0xFFFFu;
    }
    return t;
  }
  if (g.d & D_BRANCH) {
    // taken <=> ((flag word & mask) != 0) ^ invert; flag word = nz[8:0] | C << 9 | V << 15 (cpu_fast)
    const uint32_t ax = (g.d >> 16) & 0xFFu, sel = ax >> 6, wanted = ax & 1u;
    const uint32_t mask = sel == 0u ? 0x180u : sel == 1u ? 0x8000u : sel == 2u ? 0x200u : 0xFFu;
    // N, V, C: the test reads the flag itself; Z: the test reads "not zero"
    t.pad0 = mask | ((sel == 3u ? wanted : (wanted ^ 1u)) << 16);
    t.f |= FX_VALID;
    return t;
  }
  // the rest: neutral datapath unless set below
  t.k = lenbits | (uint32_t(SEL_ZERO) << K_ISEL) | (uint32_t(SEL_ZERO) << K_ASEL) | (uint32_t(SEL_ZERO) << K_BSEL);
  t.x = g.x & 0xFFFFu; t.dm = 0u;
  switch (op) {
    case O_FLAG: {
      const uint32_t ax = (g.d >> 16) & 0xFFu, mask = 1u << (ax >> 1);
      t.f |= FX_VALID | mask | ((ax & 1u) ? (mask << 8) : 0u);
      break;
    }
    case O_JMP: if (mode == AM_ABS) t.f |= FX_VALID | (1u << FX_PCS); break;
    case O_BIT: t.f |= FX_VALID | FX_BIT; break;
    case O_JSR: t.f = (0u << FX_SPD) | FX_VALID | (1u << FX_PCS) | FX_PUSH | FX_PUSH2 | FX_W_RET; t.d |= D_WRITE; break;
    case O_RTS: t.f = (4u << FX_SPD) | FX_VALID | (2u << FX_PCS) | FX_PAIR | FX_PAIR_STACK; break;
    case O_PHA: t.f = (1u << FX_SPD) | FX_VALID | FX_PUSH; t.d |= D_WRITE; t.k = sta.k | lenbits | (uint32_t(SEL_ZERO) << K_ISEL); break;
    // (PHP / PLP take the general path: rare, and serving them cost every tick a dozen instructions)
    case O_PLA: t.f = (3u << FX_SPD) | FX_VALID | FX_PULL; t.d |= D_READ; t.k = lda.k | lenbits | (uint32_t(SEL_ZERO) << K_ISEL); t.dm = lda.dm; break;
    default: break;
  }
  return t;
}

// the wide entry of one opcode
constexpr FastEnt fast_decode_entry(int opc) {
  const FastCompact c = fast_compact_entry(opc);
  const uint32_t k = c.k, d = c.d, f = c.f;
  FastEnt t = {};
  t.sel_a = (k & 7u) | 0x7770u;
  t.sel_b = ((k >> K_BSEL) & 7u) | 0x7770u;
  t.sel_idx = ((k >> K_ISEL) & 7u) | 0x7770u;
  t.sel_fn = ((k >> K_FN) & 7u) | 0x7770u;
  t.sel_cout = ((k >> K_CSRC) & 3u) | 0x4440u;
  t.xm = c.x & 0xFFFFu; t.binv = (c.x >> 16) & 0xFFu; t.pm = c.x >> 24;
  t.dm = c.dm;
  t.spd = (((f >> FX_SPD) & 7u) - 2u) << 24;
  t.cyc = (d >> 12) & 15u;
  t.seqinc = ((k >> K_LEN) & 3u) + 1u;
  t.cconst = (k >> K_CSEL) & 1u; t.cmask = (k >> (K_CSEL + 1)) & 1u;
  t.rotmask = (k >> 11) & 1u;
  t.nzmask = (k & K_NZ) ? 0xFFFFFFFFu : 0u;
  t.pclr = f & 0xFFu; t.pset = (f >> 8) & 0xFFu;
  // branch test: flag word of the compact form = nz[8:0] | C << 9 | V << 15; here nz and P are tested where they are
  t.bm_nz = c.pad0 & 0x1FFu; t.bm_p = (c.pad0 >> 9) & 0x41u;
  // pair address: (zp,X): operand + X; (zp),Y: operand; RTS / PLA: SP + 1.  perm8 pools {A X Y SP | operand 0 0 0}
  // and {A X Y SP | 0 1 0 0}
  const bool from_stack = (f & (FX_PAIR_STACK | FX_PULL)) != 0u;
  t.sel_pb = from_stack ? 0x5553u : 0x5554u;
  t.sel_padd = from_stack ? 0x4445u : (f & FX_PAIR_X) ? 0x4441u : 0x4444u;
  const uint32_t pcs = (f >> FX_PCS) & 3u;
  // BIT runs through the datapath as A AND M without a destination: Z comes out of nz = A & M, N and V are copied from
  // the operand by two masked moves (penbit low byte)
  if (f & FX_BIT) { t.sel_a = uint32_t(SEL_A) | 0x7770u; t.sel_b = uint32_t(SEL_M) | 0x7770u; t.sel_fn = uint32_t(FN_AND) | 0x7770u; t.nzmask = 0xFFFFFFFFu; }
  t.penbit = ((d & D_PAGEPEN) ? 0xFF00u : 0u) | ((f & FX_BIT) ? 0xC0u : 0u);
  t.g = ((f & FX_VALID) ? G_VALID : 0u) | ((f & FX_PAIR) ? G_NEEDPAIR : 0u) |
        (((f & (FX_PAIR | FX_PAIR_STACK)) == FX_PAIR) ? G_PTR : 0u) | ((f & FX_PULL) ? G_RA_P0 : 0u) |
        ((d & D_READ) ? G_READ : 0u) | ((d & D_WRITE) ? G_WRITE : 0u) | ((f & FX_PUSH) ? G_PUSH : 0u) |
        ((f & FX_PUSH2) ? G_PUSH2 : 0u) | ((f & FX_W_RET) ? G_W_RET : 0u) | ((f & FX_BIT) ? G_BIT : 0u) |
        ((c.pad0 >> 16) ? G_INV : 0u) | (pcs == 1u ? G_PC_EA : 0u) | (pcs == 2u ? G_PC_PAIR : 0u) |
        ((k & K_DECIMAL) ? G_DECIMAL : 0u) | ((d & D_PAGEPEN) ? G_PAGEPEN : 0u);
  return t;
}

}  // namespace mn
