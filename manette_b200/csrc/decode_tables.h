// manette_b200 -- host-side builder of the packed 6502 decode descriptors staged into shared
// memory by the emulator kernels (see emu_core.cuh: Tables / MN_DESC).
// Source of truth is the opcode matrix below (mnemonic + addressing mode + base cycles), the
// same information the reference's emulator dependency (ALE/Stella M6502Low) keeps in its
// instruction tables; undocumented opcodes included.
#pragma once
#include <string.h>
#include "emu_core.cuh"

namespace mn {

struct OpcodeRow { const char* text; };   // "MNE mode cycles"

// 16 x 16 opcode matrix, row = high nibble
static const char* const kOpcodeMatrix[256] = {
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

static const MnemonicInfo kMnemonics[] = {
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

static const char* const kModeNames[] = {"imp", "acc", "imm", "zp", "zpx", "zpy", "abs", "abx", "aby", "izx", "izy", "rel", "ind"};

// Control word of cpu_step's table-driven datapath for one opcode: which register / constant feeds each
// ALU input, the function, which flags and registers take the result.  Opcodes it cannot express (stack,
// flow, flag ops, BIT, undocumented read-modify-write combinations) are left to cpu_special().
inline uint32_t datapath_control(int op, int mode) {
  uint32_t len = (mode <= AM_ACC) ? 1u : (mode == AM_ABS || mode == AM_ABX || mode == AM_ABY || mode == AM_IND) ? 3u : 2u;
  uint32_t isel = (mode == AM_ZPX || mode == AM_ABX) ? 1u : (mode == AM_ZPY || mode == AM_ABY) ? 2u : 0u;
  uint32_t k = ((len - 1) << K_LEN) | (isel << K_ISEL);
  const uint32_t G = K_GENERIC;
  auto mk = [&](uint32_t asel, uint32_t bsel, uint32_t csel, uint32_t fn, uint32_t flags) {
    return k | G | (asel << K_ASEL) | (bsel << K_BSEL) | (csel << K_CSEL) | (fn << K_FN) | flags;
  };
  const uint32_t shift_src = (mode == AM_ACC) ? AS_A : AS_M, shift_dst = (mode == AM_ACC) ? K_DA : 0u;
  switch (op) {
    case O_NOP: return mk(AS_ZERO, BS_ZERO, 0, FN_ADD, 0);
    case O_LDA: return mk(AS_M, BS_ZERO, 0, FN_ADD, K_NZ | K_DA);
    case O_LDX: return mk(AS_M, BS_ZERO, 0, FN_ADD, K_NZ | K_DX);
    case O_LDY: return mk(AS_M, BS_ZERO, 0, FN_ADD, K_NZ | K_DY);
    case O_LAX: return mk(AS_M, BS_ZERO, 0, FN_ADD, K_NZ | K_DA | K_DX);
    case O_STA: return mk(AS_A, BS_ZERO, 0, FN_ADD, 0);
    case O_STX: return mk(AS_X, BS_ZERO, 0, FN_ADD, 0);
    case O_STY: return mk(AS_Y, BS_ZERO, 0, FN_ADD, 0);
    case O_SAX: return mk(AS_AX, BS_ZERO, 0, FN_ADD, 0);
    case O_ORA: return mk(AS_A, BS_M, 0, FN_OR, K_NZ | K_DA);
    case O_AND: return mk(AS_A, BS_M, 0, FN_AND, K_NZ | K_DA);
    case O_EOR: return mk(AS_A, BS_M, 0, FN_EOR, K_NZ | K_DA);
    case O_ADC: return mk(AS_A, BS_M, 2, FN_ADD, K_NZ | K_C | K_V | K_DA | K_DECIMAL);
    case O_SBC: return mk(AS_A, BS_M, 2, FN_ADD, K_NZ | K_C | K_V | K_DA | K_DECIMAL | K_BINV);
    case O_CMP: return mk(AS_A, BS_M, 1, FN_ADD, K_NZ | K_C | K_BINV);
    case O_CPX: return mk(AS_X, BS_M, 1, FN_ADD, K_NZ | K_C | K_BINV);
    case O_CPY: return mk(AS_Y, BS_M, 1, FN_ADD, K_NZ | K_C | K_BINV);
    case O_INC: return mk(AS_M, BS_ONE, 0, FN_ADD, K_NZ);
    case O_DEC: return mk(AS_M, BS_FF, 0, FN_ADD, K_NZ);
    case O_INX: return mk(AS_X, BS_ONE, 0, FN_ADD, K_NZ | K_DX);
    case O_INY: return mk(AS_Y, BS_ONE, 0, FN_ADD, K_NZ | K_DY);
    case O_DEX: return mk(AS_X, BS_FF, 0, FN_ADD, K_NZ | K_DX);
    case O_DEY: return mk(AS_Y, BS_FF, 0, FN_ADD, K_NZ | K_DY);
    case O_TAX: return mk(AS_A, BS_ZERO, 0, FN_ADD, K_NZ | K_DX);
    case O_TAY: return mk(AS_A, BS_ZERO, 0, FN_ADD, K_NZ | K_DY);
    case O_TXA: return mk(AS_X, BS_ZERO, 0, FN_ADD, K_NZ | K_DA);
    case O_TYA: return mk(AS_Y, BS_ZERO, 0, FN_ADD, K_NZ | K_DA);
    case O_TSX: return mk(AS_SP, BS_ZERO, 0, FN_ADD, K_NZ | K_DX);
    case O_TXS: return mk(AS_X, BS_ZERO, 0, FN_ADD, K_DSP);
    case O_ASL: return mk(shift_src, BS_ZERO, 0, FN_ASL, K_NZ | K_C | shift_dst);
    case O_LSR: return mk(shift_src, BS_ZERO, 0, FN_LSR, K_NZ | K_C | shift_dst);
    case O_ROL: return mk(shift_src, BS_ZERO, 0, FN_ROL, K_NZ | K_C | shift_dst);
    case O_ROR: return mk(shift_src, BS_ZERO, 0, FN_ROR, K_NZ | K_C | shift_dst);
    default: return k;
  }
}

inline void build_tables(Tables* t) {
  memset(t, 0, sizeof(*t));
  for (int opc = 0; opc < 256; ++opc) {
    const char* row = kOpcodeMatrix[opc];
    char mne[4] = {row[0], row[1], row[2], 0};
    const char* mode_txt = row + 4;
    int mode = -1, mlen = 0;
    while (mode_txt[mlen] && mode_txt[mlen] != ' ') ++mlen;
    for (int m = 0; m < 13; ++m)
      if ((int)strlen(kModeNames[m]) == mlen && strncmp(kModeNames[m], mode_txt, mlen) == 0) mode = m;
    int cyc = mode_txt[mlen + 1] - '0';
    const MnemonicInfo* mi = 0;
    for (size_t k = 0; k < sizeof(kMnemonics) / sizeof(kMnemonics[0]); ++k)
      if (strcmp(kMnemonics[k].name, mne) == 0) mi = &kMnemonics[k];
    int cls = mi->cls;
    if (mi->op == O_NOP && mode == AM_IMP) cls = OC_NONE;   // only the multi-byte NOPs touch memory
    const bool has_ea = mode >= AM_ZP && mode != AM_REL;
    const bool wide = mode == AM_ABS || mode == AM_ABX || mode == AM_ABY || mode == AM_IND;
    uint32_t d = MN_DESC(mode, cls, mi->op, cyc) | (uint32_t(mi->aux & 0xFF) << 16);
    if (has_ea) d |= D_EA;
    if (mode >= AM_IZX && mode != AM_REL) d |= D_INDIRECT;
    if (has_ea && (cls == OC_READ || cls == OC_RMW)) d |= D_READ;
    if (cls == OC_WRITE || (cls == OC_RMW && mode != AM_ACC)) d |= D_WRITE;
    if (has_ea && cls == OC_READ) d |= D_PAGEPEN;
    if (mi->op == O_BRANCH) d |= D_BRANCH;
    t->e[opc].k = datapath_control(mi->op, mode);
    t->e[opc].d = d;
    t->e[opc].x = has_ea ? (wide ? 0xFFFFu : 0xFFu) : 0u;
    t->e[opc].pad = 0;
  }
}
inline int desc_cycles(uint32_t d) { return int((d >> 12) & 15); }

}  // namespace mn
