// Probe: sustained tcgen05.mma.cta_group::1.kind::f16 rate with both operands in shared memory (SS), per N and per
// A-descriptor alignment, one CTA per SM, no TMA traffic during the timed region.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 --expt-relaxed-constexpr -o umma_rate umma_rate.cu ; ./umma_rate
// Prints clk per MMA instruction (M=128, K=16) and the fraction of the 128*N/256-clk floor.
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tWAIT_LOOP:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra WAIT_DONE;\n\tbra WAIT_LOOP;\n\tWAIT_DONE:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }

// mode bit0: A start shifted by 3 rows (384 B); bit1: rotate over 4 A tiles / 2 B tiles (distinct smem addresses)
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
// mode bit2: tcgen05.fence::after_thread_sync every 8 MMAs; bit3: tcgen05.commit to a scratch mbarrier every 8 MMAs;
// bit4: warps 2-5 keep reading the other accumulator with tcgen05.ld; bit5: warp 1 keeps 4 TMA loads of 16 KB in flight
template <int N>
__global__ void __launch_bounds__(192, 1) rate(const __grid_constant__ CUtensorMap tm, int iters, int mode, long long* out, int delay, int group) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  // A: 4 tiles of (128+8 rows) x 128 B ; B: 2 tiles of N rows x 128 B
  const uint32_t a_s = smem_u32(smem), b_s = a_s + 4 * 17408;
  const uint32_t ring = b_s + 2 * N * 128;  // 4 x 16 KB TMA landing zone
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 4 * 17408 + 2 * N * 128 + 4 * 16384);
  const uint32_t done = smem_u32(bars), scratch = done + 8, tb0 = done + 16, scratch2 = done + 48, scratch3 = done + 56;
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 8);
  volatile int* stop = reinterpret_cast<volatile int*>(bars + 9);
  const int tid = threadIdx.x, warp = tid >> 5;
  // mode bit6: pseudo-random bf16 operands in [-2, 2) instead of zeros (data-dependent power)
  for (int i = tid; i < (4 * 17408 + 2 * N * 128) / 4; i += 192) {
    uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u;
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    const uint32_t lo = 0x3F80u | (h & 0x807Fu), hi = 0x3F80u | ((h >> 16) & 0x807Fu);  // +-[1,2)
    reinterpret_cast<uint32_t*>(smem)[i] = (mode & 64) ? (lo | (hi << 16)) : 0u;
  }
  if (tid == 0) {
    mbar_init(done, 1); mbar_init(scratch, 1); mbar_init(scratch2, 1); mbar_init(scratch3, 1);
    for (int i = 0; i < 4; ++i) mbar_init(tb0 + 8 * i, 1);
    *stop = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot;
  if (tid == 0) {
    constexpr uint32_t idesc = make_idesc(128, N);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const int ta = (mode & 2) ? (it & 3) : 0, tb = (mode & 2) ? (it & 1) : 0;
      const uint64_t adesc = make_desc(a_s + ta * 17408 + ((mode & 1) ? 384 : 0), 16, 1024);
      const uint64_t bdesc = make_desc(b_s + tb * N * 128, 16, 1024);
#pragma unroll
      if (mode & 1024) { const long long tt = clock64(); while (clock64() - tt < 256) { } continue; }
      for (int k = 0; k < 4; ++k) {
        const uint32_t acc = 1;
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(adesc + (uint64_t)(k * 2)), "l"(bdesc + (uint64_t)(k * 2)), "r"(idesc), "r"(acc) : "memory");
      }
      if ((it + 1) % group == 0 && delay) {
        if (mode & 512) mbar_wait(scratch3, 1);  // a fresh barrier: the wait on parity 1 returns at the first poll
        const long long tt = clock64();
        while (clock64() - tt < delay) { }
      }
      if (it & 1) {
        if (mode & 8) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(scratch) : "memory");
        if (mode & 4) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      }
    }
    const long long t1 = clock64();
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(done) : "memory");
    mbar_wait(done, 0);
    const long long t2 = clock64();
    out[blockIdx.x * 2] = t1 - t0;
    out[blockIdx.x * 2 + 1] = t2 - t0;
    *stop = 1;
  } else if (tid == 32 && (mode & 32)) {
    int it = 0;
    while (!*stop) {
      const int s = it & 3;
      if (it >= 4) mbar_wait(tb0 + 8 * s, ((it >> 2) - 1) & 1);
      mbar_expect_tx(tb0 + 8 * s, 16384);
      tma_load_2d(ring + s * 16384, &tm, tb0 + 8 * s, 0, (it * 128) & 32767);
      ++it;
    }
    for (int j = 0; j < 4 && j < it; ++j) { const int k = it - 1 - j; mbar_wait(tb0 + 8 * (k & 3), (k >> 2) & 1); }
    out[blockIdx.x * 2 + 296] = it;
  } else if (warp >= 2 && (mode & 384)) {
    // bit7: warps 2-5 spin on mbarrier.try_wait of a barrier that never completes (what idle epilogue warps do);
    // bit8: same, with a nanosleep back-off between polls
    int polls = 0;
    while (true) {
      uint32_t ok;
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(scratch2), "r"(0) : "memory");
      if (mode & 256) __nanosleep(500);
      if ((++polls & 15) == 0 && *stop) break;
    }
    if (tid == 64) out[blockIdx.x * 2 + 297] = polls;
  } else if (warp >= 2 && (mode & 16)) {
    uint32_t sink = 0;
    long long nld = 0;
    while (!*stop) {
      ++nld;
      uint32_t v[32];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
            "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
            "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
            "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
          : "r"(tmem + 256 + ((uint32_t)((warp & 3) * 32) << 16)));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      sink += v[0] + v[31];
    }
    if (sink == 0x12345678) out[0] = sink;
    if (tid == 64) out[blockIdx.x * 2 + 297] = nld;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

template <int N>
void run(int grid, long long* d_out, const CUtensorMap& tm) {
  const int smem = 4 * 17408 + 2 * N * 128 + 4 * 16384 + 1024 + 128;
  cudaFuncSetAttribute(rate<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 4096;
  for (int cfg = 0; cfg < 2; ++cfg) {
    const int mode = cfg == 0 ? 64 + 16 : 64 + 16 + 1024;   // tcgen05.ld loop in 4 warps, with / without the MMA stream
    const int delay = 0, group = 2;
    cudaMemset(d_out, 0, sizeof(long long) * 4 * 148);
    rate<N><<<grid, 192, smem>>>(tm, iters, mode, d_out, delay, group);
    rate<N><<<grid, 192, smem>>>(tm, iters, mode, d_out, delay, group);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return; }
    long long h[4 * 148];
    cudaMemcpy(h, d_out, sizeof(long long) * 4 * 148, cudaMemcpyDeviceToHost);
    double issue = 0, total = 0;
    for (int i = 0; i < grid; ++i) { issue += h[2 * i]; total += h[2 * i + 1]; }
    issue /= grid; total /= grid;
    const double per = total / (iters * 4.0), floor_clk = 128.0 * N / 256.0;
    printf("N=%3d grid=%3d random=%d fence=%d commit=%d tmem_ld=%d tma=%d: issue %.1f clk/MMA, complete %.1f clk/MMA, floor %.0f -> %.1f %% of peak (tma loads/CTA %lld = %.1f B/clk) no_mma=%d: tcgen05.ld.32x32b.x32+wait per warp: %lld in %.0f clk = %.1f clk each\n", N, grid, (mode >> 6) & 1, (mode >> 2) & 1, (mode >> 3) & 1, (mode >> 4) & 1, (mode >> 5) & 1, issue / (iters * 4.0), per, floor_clk, 100.0 * floor_clk / per, h[296], h[296] * 16384.0 / total, (mode >> 10) & 1, h[297], total, total / (double)(h[297] ? h[297] : 1));
  }
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, sizeof(long long) * 4 * 148);
  void* gbuf;
  cudaMalloc(&gbuf, 32768 * 128 * 2 + (1 << 20));
  cudaMemset(gbuf, 0, 32768 * 128 * 2);
  cudaDriverEntryPointQueryResult q;
  void* f = nullptr;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q);
  typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  CUtensorMap tm;
  cuuint64_t dims[2] = {64, 65536};
  cuuint64_t strides[1] = {128};
  cuuint32_t box[2] = {64, 128};
  cuuint32_t estr[2] = {1, 1};
  if (((EncodeTiledFn)f)(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, gbuf, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("tensor map failed\n"); return 1; }
  for (int grid : {148}) {
    run<64>(grid, d_out, tm);
    run<128>(grid, d_out, tm);
    run<256>(grid, d_out, tm);
  }
  return 0;
}
