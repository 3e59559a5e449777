// Integer-pipe throughput probe for the SHA-512 (FSS) kernel: SHF / LOP3 / IADD3 / IMAD / IMAD.WIDE and mixes.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o int_pipes int_pipes.cu ; prints warp-instructions / clk / SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 4096
template <int MODE>
__global__ void k(uint32_t* out, uint32_t seed, long long* cyc) {
  uint32_t a[8], b[8];
  uint64_t w[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = seed + threadIdx.x * 7 + i; b[i] = seed * 3 + i + threadIdx.x * 13; w[i] = a[i]; }
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("shf.r.wrap.b32 %0, %0, %1, 7;" : "+r"(a[i]) : "r"(b[i]));
      if (MODE == 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(seed));
      if (MODE == 2) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i]));
      if (MODE == 3) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(seed));
      if (MODE == 4) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(b[i]), "r"(seed));
      if (MODE == 5) { asm volatile("shf.r.wrap.b32 %0, %0, %1, 7;" : "+r"(a[i]) : "r"(b[i]));
                       asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(b[i]), "r"(seed)); }
      if (MODE == 6) { asm volatile("shf.r.wrap.b32 %0, %0, %1, 7;" : "+r"(a[i]) : "r"(b[i]));
                       asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b[i]) : "r"(seed), "r"(seed)); }
      if (MODE == 7) asm volatile("add.u64 %0, %0, %1;" : "+l"(w[i]) : "l"((uint64_t)b[i] << 32 | seed));
      if (MODE == 8) { asm volatile("shf.r.wrap.b32 %0, %0, %1, 7;" : "+r"(a[i]) : "r"(b[i]));
                       asm volatile("shf.r.wrap.b32 %0, %0, %1, 9;" : "+r"(b[i]) : "r"(seed));
                       asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(seed), "r"(seed)); }
    }
  }
  long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i] + b[i] + (uint32_t)w[i] + (uint32_t)(w[i] >> 32);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int MODE>
void run(const char* name, int per_iter) {
  uint32_t* out; long long* cyc; long long h;
  cudaMalloc(&out, 148 * 4 * 1024 * 4); cudaMalloc(&cyc, 8);
  const int threads = 512;  // 16 warps / SM, 4 per SMSP
  k<MODE><<<148, threads>>>(out, 12345, cyc);
  k<MODE><<<148, threads>>>(out, 12345, cyc);
  cudaDeviceSynchronize();
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  double winstr = (double)ITERS * 8 * per_iter * (threads / 32);
  printf("%-28s %8.2f warp-instr/clk/SM  (%lld clk)\n", name, winstr / h, h);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<0>("shf", 1); run<1>("lop3", 1); run<2>("iadd", 1); run<3>("imad.lo", 1); run<4>("imad.wide.u32", 1);
  run<5>("shf + imad.wide", 2); run<6>("shf + imad.lo", 2); run<7>("add.u64", 1); run<8>("2 shf + imad.wide", 3);
  return 0;
}
