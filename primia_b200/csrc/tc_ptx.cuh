// Thin inline-PTX wrappers for the sm_100a tensor-core path: mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma /
// commit / ld), shared-memory matrix descriptors.  Used by conv_halo.cu.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace tcx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
               "l"(tm), "r"(bar), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::
                   "r"(dst),
               "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// One lane of a converged warp.  Producer / MMA roles run with the WHOLE warp converged and issue inside
// `if (elect_one()) { ... }`: ptxas recognises elect.sync and keeps descriptors, barrier addresses and TMEM addresses in uniform
// registers, so tcgen05.mma / cp.async.bulk.tensor compile to back-to-back UTCHMMA / UTMALDG.  A single-thread role written as
// `if (threadIdx.x == X)` instead gets an ELECT / R2UR.BROADCAST / BRA.U.ANY "waterfall" around EVERY such instruction (~10
// dependent instructions each), and the issuing thread -- not the tensor pipe -- paces the kernel (measured: 100-150 clk per
// 128x128x16 MMA instead of 64).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(pred));
  return pred != 0;
}
// descriptor from its two 32-bit halves (the low half carries the 16-byte-unit start address: add offsets there)
__device__ __forceinline__ uint64_t desc_pack(uint32_t lo, uint32_t hi) { return (uint64_t)lo | ((uint64_t)hi << 32); }
constexpr uint32_t DESC_SW128_HI = (1024u >> 4) | (1u << 14) | (2u << 29);  // SBO = 1024 B, bit 46, SWIZZLE_128B
constexpr uint32_t DESC_SW128_LO = (16u >> 4) << 16;                         // LBO = 16 B (ignored for K-major SW128)

// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread = TMEM lane = accumulator row)
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t v[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// K-major / MN-major SWIZZLE_128B shared-memory matrix descriptor.  The swizzle XOR is applied by the hardware on absolute
// shared-memory address bits, so `saddr` may be any 128-byte row of a TMA-written tile (probe: scripts/ubench/umma_shift.cu).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::f16 instruction descriptor: D fp32, A/B bf16, a_mn / b_mn = 1 for MN-major operands
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

// ---- epilogue of one [32 rows x 32 columns] accumulator chunk held as tcgen05.ld 32x32b registers (lane = row).
// Writing each lane's row straight to global memory costs one 16-byte store per 128-byte line and lane (32 LSU wavefronts per
// instruction): the epilogue, not the MMA, then paces N = 64 layers.  Instead the warp stages the bf16 rows in a 2 KB
// shared-memory scratch (64-byte rows, 16-byte chunks XOR-swizzled by (row >> 1) & 3: conflict-free for the row-wise writes,
// the 4-lanes-per-row reads and the column reads) and (1) stores 8 rows x 64 contiguous bytes per instruction, (2) optionally
// adds into the existing bf16 values (dgrad accumulate), (3) optionally accumulates the per-column sum / sum of squares of the
// ROUNDED values (BatchNorm batch statistics) into per-lane registers, flushed with stats_flush32.
//   valid: this lane's row exists ; out: its global address at the chunk's first column (ignored when !valid).
__device__ __forceinline__ void epilogue_chunk32(const uint32_t (&v)[32], bool valid, __nv_bfloat16* out, bool accumulate,
                                                 bool do_stats, uint8_t* scr /* 2048 B, this warp's */, int lane,
                                                 float (&st)[4] /* this lane's running partial sums (see stats_flush32) */) {
  {
    const uint32_t swz = (uint32_t)(lane >> 1) & 3u;
    uint8_t* my = scr + lane * 64;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint4 pk;
      __nv_bfloat162* ph = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
      for (int e = 0; e < 4; ++e) ph[e] = __floats2bfloat162_rn(__uint_as_float(v[q * 8 + 2 * e]), __uint_as_float(v[q * 8 + 2 * e + 1]));
      if (!valid) pk = make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4*>(my + (((uint32_t)q ^ swz) << 4)) = pk;
    }
  }
  __syncwarp();
  {
    // 4 lanes per row, 8 rows per instruction: all shuffles, then all shared-memory reads, then the (predicated) global
    // accesses -- independent instruction streams instead of four dependent shuffle -> load -> store chains
    const int sub = lane & 3, r0 = lane >> 2;
    const unsigned long long mine = valid ? (unsigned long long)(uintptr_t)out : 0ull;
    unsigned long long p[4];
    uint4 nv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) p[i] = __shfl_sync(0xffffffffu, mine, i * 8 + r0);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = i * 8 + r0;
      nv[i] = *reinterpret_cast<const uint4*>(scr + r * 64 + ((((uint32_t)sub) ^ ((uint32_t)(r >> 1) & 3u)) << 4));
    }
    if (accumulate) {
      uint4 old[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) old[i] = p[i] ? *(reinterpret_cast<const uint4*>((uintptr_t)p[i]) + sub) : make_uint4(0, 0, 0, 0);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const __nv_bfloat162* oh = reinterpret_cast<const __nv_bfloat162*>(&old[i]);
        __nv_bfloat162* nh = reinterpret_cast<__nv_bfloat162*>(&nv[i]);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 a = __bfloat1622float2(nh[e]), b = __bfloat1622float2(oh[e]);
          nh[e] = __floats2bfloat162_rn(a.x + b.x, a.y + b.y);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (p[i]) *(reinterpret_cast<uint4*>((uintptr_t)p[i]) + sub) = nv[i];
  }
  if (do_stats) {
    // lane = (column pair cp, row parity): partial sums of its 16 rows stay in registers across tiles (shared-memory float
    // atomics are CAS loops: ~1000 clk per chunk when 8 warps hit the same 64 addresses)
    const int cp = lane & 15, par = lane >> 4;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int r = 2 * i + par;
      const uint32_t w = *reinterpret_cast<const uint32_t*>(scr + r * 64 + ((((uint32_t)(cp >> 2)) ^ ((uint32_t)(r >> 1) & 3u)) << 4) + (cp & 3) * 4);
      const float x0 = __uint_as_float(w << 16), x1 = __uint_as_float(w & 0xffff0000u);
      st[0] += x0; st[1] += x1;
      st[2] = fmaf(x0, x0, st[2]); st[3] = fmaf(x1, x1, st[3]);
    }
  }
  __syncwarp();  // the scratch is reused by the next chunk
}
// adds the warp's partial sums of one 32-column chunk into s1[32] / s2[32] and clears them.  s1 / s2 point into THIS WARP's
// private shared-memory slot (no atomics): every CTA-level sum is then formed in a fixed order, so the statistics -- and with
// them the whole bf16 forward -- do not depend on warp scheduling (the cross-CTA step is a double-precision atomic per CTA and
// channel: a 1e-16 relative order effect, far below fp32 resolution).
__device__ __forceinline__ void stats_flush32(float (&st)[4], int lane, float* s1, float* s2) {
#pragma unroll
  for (int k = 0; k < 4; ++k) st[k] += __shfl_xor_sync(0xffffffffu, st[k], 16);
  if (lane < 16) {
    s1[2 * lane] += st[0]; s1[2 * lane + 1] += st[1];
    s2[2 * lane] += st[2]; s2[2 * lane + 1] += st[3];
  }
  __syncwarp();
#pragma unroll
  for (int k = 0; k < 4; ++k) st[k] = 0.f;
}

}  // namespace tcx
