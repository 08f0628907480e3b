// Measurement only: what does an idle warp cost on sm_100a?  Cycles per __nanosleep(t) (alone, and beside a busy warp on
// the same SM sub-partition), per dependent L2 load (ld.volatile.global chain), and per mbarrier.try_wait that times out.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/sleep_probe tools/sleep_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void k_sleep(unsigned ns, int iters, int busy_warps, long long* out, unsigned* sink) {
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    const long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) __nanosleep(ns);
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = (t1 - t0) / iters;
  } else if (warp <= busy_warps * 4 && (warp & 3) == 0) {   // warps 4, 8 ... share sub-partition 0 with warp 0
    unsigned x = threadIdx.x;
#pragma unroll 1
    for (int i = 0; i < iters * 64; ++i) x = x * 1664525u + 1013904223u;
    sink[threadIdx.x] = x;
  }
}

__global__ void k_chase(const unsigned* zero, int iters, long long* out, unsigned* sink) {
  unsigned long long a = reinterpret_cast<unsigned long long>(zero);
  unsigned v = 0;
  const long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(a) : "memory");
    a += v;
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) { out[0] = (t1 - t0) / iters; sink[0] = v; }
}

__global__ void k_trywait(unsigned hint_ns, int iters, long long* out) {
  __shared__ unsigned long long bar;
  if (threadIdx.x == 0) {
    unsigned a = unsigned(__cvta_generic_to_shared(&bar));
    asm volatile("mbarrier.init.shared.b64 [%0], 1;" :: "r"(a));
  }
  __syncthreads();
  const unsigned a = unsigned(__cvta_generic_to_shared(&bar));
  unsigned ok = 0;
  const long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared.b64 p, [%1], %2, %3; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(a), "r"(0u), "r"(hint_ns) : "memory");
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) { out[0] = (t1 - t0) / iters; out[1] = ok; }
}

int main() {
  long long* out; unsigned* sink; unsigned* zero;
  cudaMallocManaged(&out, 64); cudaMalloc(&sink, 4096); cudaMalloc(&zero, 4); cudaMemset(zero, 0, 4);
  const unsigned ns[] = {20, 200, 512, 1000, 4000, 20000, 100000};
  for (int busy = 0; busy <= 1; ++busy)
    for (unsigned t : ns) {
      const int iters = t >= 20000 ? 200 : 4000;
      k_sleep<<<1, 256>>>(t, iters, busy, out, sink);
      cudaDeviceSynchronize();
      printf("nanosleep(%u) busy_warps=%d : %lld cycles per call\n", t, busy, out[0]);
    }
  for (int lanes : {1, 32}) {
    k_chase<<<1, lanes>>>(zero, 4000, out, sink);
    cudaDeviceSynchronize();
    printf("ld.volatile.global chain, %d lane(s): %lld cycles per load\n", lanes, out[0]);
  }
  for (unsigned h : {0u, 1000u, 20000u, 1000000u}) {
    k_trywait<<<1, 32>>>(h, 200, out);
    cudaDeviceSynchronize();
    printf("mbarrier.try_wait (never completes) hint %u ns: %lld cycles per call, ok=%lld\n", h, out[0], out[1]);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
