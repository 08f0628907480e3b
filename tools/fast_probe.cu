// TOOL (not product): the fast tick alone in a loop, to count its SASS instructions off line:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -cubin -o /tmp/probe/p.cubin tools/fast_probe.cu
//   cuobjdump -sass /tmp/probe/p.cubin | grep -c '^\s*/\*[0-9a-f]*\*/'
#include "../manette_b200/csrc/emu_core.cuh"
using namespace mn;
template <bool FLAT>
__global__ void probe(uint32_t* io, int n, maddr rom, maddr ram, maddr tab, maddr fifo, maddr core) {
  Cpu r;
  r.axys = io[0]; r.PC = io[1]; r.P = io[2]; r.nz = io[3]; r.dbus = io[4]; r.segmap = io[5]; r.romw = io[12]; r.hot_lo = io[6];
  r.cycles = int(io[7]); r.clk0 = int(io[8]); r.cyc0 = int(io[9]); r.fifo_n = int(io[10]); r.stop = false;
  r.def_lo = r.def_hi = r.dep_lo = r.dep_hi = 0; r.tainted = false;
  Mem mm; mm.rom = rom; mm.ram = ram; mm.tab = tab; mm.fifo = fifo; mm.core = core;
  int acc = 0;
#pragma unroll 1
  for (int i = 0; i < n; ++i) acc += cpu_fast<false, FLAT>(mm, r, (i & 7) != 7) ? 1 : 0;
  io[0] = r.axys; io[1] = r.PC; io[2] = r.P; io[3] = r.nz; io[4] = r.dbus; io[7] = uint32_t(r.cycles); io[10] = uint32_t(r.fifo_n); io[11] = uint32_t(acc);
}
template __global__ void probe<true>(uint32_t*, int, maddr, maddr, maddr, maddr, maddr);
template __global__ void probe<false>(uint32_t*, int, maddr, maddr, maddr, maddr, maddr);
