// manette_b200 -- host-side builder of the 6502 decode table staged into shared memory by the emulator
// kernels: a plain copy of the compile-time entries of cpu_defs.h (decode_entry).
#pragma once
#include <string.h>
#include "cpu_defs.h"

namespace mn {

inline void build_tables(Tables* t) {
  memset(t, 0, sizeof(*t));
  for (int opc = 0; opc < 256; ++opc) { t->e[opc] = decode_entry(opc); t->f[opc] = fast_decode_entry(opc); }
}
inline int desc_cycles(uint32_t d) { return int((d >> 12) & 15); }

}  // namespace mn
