// Path T, throughput mode, TMA-fed variant of the tcgen05 implicit-GEMM convolutions (conv_tc.cu holds the
// cp.async-gather variant, which remains the fallback for the shapes TMA im2col cannot express here: the
// 3(->8)-channel stem and stride-2 dgrad).
//
//   warp 0 (one lane): TMA producer -- the activation operand is fetched with im2col-mode tensor maps
//           (cp.async.bulk.tensor.4d...im2col: [128 output pixels] x [64 channels] of one filter tap per
//           instruction, zero fill for padding/out-of-range pixels), the dense operand with tiled 2D maps;
//           both land in SWIZZLE_128B shared-memory tiles and signal an mbarrier by transaction bytes;
//   warp 1 (one lane): tcgen05.mma issue (M = 128, fp32 accumulators in TMEM) + tcgen05.commit;
//   warps 2-5: epilogue (tcgen05.ld -> registers -> bf16 / fp32 global).
#include "common.cuh"
#include <cuda.h>
#include <algorithm>

namespace tma {

typedef __nv_bfloat16 bf16;
constexpr int NTHREADS = 192;
constexpr int TILE_BYTES = 128 * 128;

struct Geo {
  int B, H, W, C, K, R, S, stride, pad, Ho, Wo;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
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
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
               "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_load_im2col(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c, int w, int h, int n,
                                                uint16_t off_w, uint16_t off_h) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};" ::
          "r"(dst),
      "l"(tm), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h)
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
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t v[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

// ------------------------------------------------------------------------------------------------ fwd / dgrad(stride 1)
// PERSISTENT: one CTA per SM loops over output tiles (n fastest, so co-scheduled CTAs share the activation tile in L2).
// D[128 pixels, BN] ; A = im2col(src) via tmA ; B = dense [N][Ktot] via tmB.  FLIP: dgrad (taps mirrored).
// Two TMEM accumulator stages: the epilogue of tile i overlaps the TMA/MMA main loop of tile i+1.
// Barriers: full/empty[STAGES] (TMA <-> MMA), tmem_full/tmem_empty[2] (MMA <-> epilogue).
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// RB ("resident B"): when the whole dense operand of this CTA's n-tile fits in shared memory (N == BN and
// R*S*C*BN*2 bytes <= RB_MAX_BYTES: the 64-channel layers and the stem), it is loaded ONCE per CTA and only the activation
// tiles stream through the stage ring -- these layers are bound by L2->SM bandwidth, not by the tensor pipe.
constexpr int RB_MAX_BYTES = 80 * 1024;

template <int BN, int STAGES, bool FLIP, bool RB>
__global__ void __launch_bounds__(NTHREADS, 1)
conv_tma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, Geo p, bf16* __restrict__ dst,
                int accumulate, double* __restrict__ stats) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment as an OFFSET from the __shared__ array: the pointer keeps its address space, so the epilogue scratch
  // compiles to STS / LDS (a round trip through uintptr_t made nvcc emit generic ST.E / LD.E)
  uint8_t* smem = smem_raw + ((1024u - ((uint32_t)__cvta_generic_to_shared(smem_raw) & 1023u)) & 1023u);
  constexpr int A_BYTES = TILE_BYTES, B_BYTES = BN * 128, STAGE_BYTES = RB ? A_BYTES : A_BYTES + B_BYTES;
  constexpr int TMEM_COLS = 2 * BN;  // two accumulator stages (power of two: 128 or 256)
  const uint32_t s_base = smem_u32(smem);
  const uint32_t b_res = s_base + STAGES * STAGE_BYTES;  // RB: [nkb][BN rows][128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + (RB ? RB_MAX_BYTES : 0));
  const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * STAGES;
  const uint32_t tfull0 = full0 + 16 * STAGES, tempty0 = tfull0 + 16, bfull = tempty0 + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 5);
  float* scratch_all = reinterpret_cast<float*>(bars + 2 * STAGES + 6);  // 4 x [32][33] transposition scratch
  float* cta_stats = scratch_all + 4 * 32 * 33;                         // [4 warps][2][N]: one private slot per epilogue warp

  const int tid = threadIdx.x, warp = tid >> 5;
  const int OH = FLIP ? p.H : p.Ho, OW = FLIP ? p.W : p.Wo;
  const int CR = FLIP ? p.K : p.C;
  const int N = FLIP ? p.C : p.K;
  const int M = p.B * OH * OW;
  const int cblocks = CR / 64;
  const int nkb = p.R * p.S * cblocks;
  const int ntn = N / BN;
  const int ntiles = ((M + 127) / 128) * ntn;
  const bool do_stats = !FLIP && stats != nullptr;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull0 + 8 * a, 1);
      mbar_init(tempty0 + 8 * a, 4);  // one arrival per epilogue warp
    }
    mbar_init(bfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
  }
  if (do_stats)
    for (int i = tid; i < 4 * 2 * N; i += NTHREADS) cta_stats[i] = 0.f;
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pm_pdl_sync();  // programmatic dependent launch (common.cuh): CTA-local prologue above, first global access below

  if (tid == 0) {
    // ------------------------------------------------------------ TMA producer
    int it = 0;
    if (RB && (int)blockIdx.x < ntiles) {  // the whole dense operand, once (ntn == 1 => n0 == 0)
      mbar_expect_tx(bfull, (uint32_t)(nkb * B_BYTES));
      for (int kb = 0; kb < nkb; ++kb) tma_load_2d(b_res + kb * B_BYTES, &tmB, bfull, kb * 64, 0);
    }
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int m0 = (tile / ntn) * 128, n0 = (tile % ntn) * BN;
      const int nb = m0 / (OH * OW);
      const int rem = m0 - nb * OH * OW;
      const int py = rem / OW, px = rem - py * OW;
      const int bw = FLIP ? px + p.pad - (p.S - 1) : px * p.stride - p.pad;
      const int bh = FLIP ? py + p.pad - (p.R - 1) : py * p.stride - p.pad;
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const int s = it % STAGES;
        mbar_wait(empty0 + 8 * s, ((it / STAGES) & 1) ^ 1);
        const uint32_t a_tile = s_base + s * STAGE_BYTES, b_tile = a_tile + A_BYTES;
        const int tap = kb / cblocks, c0 = (kb - tap * cblocks) * 64;
        const int r = tap / p.S, sx = tap - r * p.S;
        mbar_expect_tx(full0 + 8 * s, STAGE_BYTES);
        tma_load_im2col(a_tile, &tmA, full0 + 8 * s, c0, bw, bh, nb, (uint16_t)(FLIP ? p.S - 1 - sx : sx),
                        (uint16_t)(FLIP ? p.R - 1 - r : r));
        if (!RB) tma_load_2d(b_tile, &tmB, full0 + 8 * s, kb * 64, n0);
      }
    }
  } else if (tid == 32) {
    // ------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = make_idesc(128, BN, 0, 0);
    int it = 0, lt = 0;
    if (RB && (int)blockIdx.x < ntiles) mbar_wait(bfull, 0);
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++lt) {
      const int as = lt & 1;
      mbar_wait(tempty0 + 8 * as, ((lt >> 1) & 1) ^ 1);  // epilogue has drained this accumulator stage
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + as * BN;
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const int s = it % STAGES;
        mbar_wait(full0 + 8 * s, (it / STAGES) & 1);
        tc_fence_after();
        const uint32_t a_tile = s_base + s * STAGE_BYTES, b_tile = RB ? b_res + kb * B_BYTES : a_tile + A_BYTES;
        const uint64_t adesc = make_desc(a_tile, 16, 1024), bdesc = make_desc(b_tile, 16, 1024);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem_d, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0);
        umma_commit(empty0 + 8 * s);
      }
      umma_commit(tfull0 + 8 * as);
    }
  } else if (warp >= 2) {
    // ------------------------------------------------------------ epilogue (4 warps)
    const int quad = warp & 3, lane = tid & 31;
    float* scratch = scratch_all + quad * (32 * 33);
    int lt = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++lt) {
      const int as = lt & 1;
      const int m0 = (tile / ntn) * 128, n0 = (tile % ntn) * BN;
      mbar_wait(tfull0 + 8 * as, (lt >> 1) & 1);
      tc_fence_after();
      const int row = m0 + quad * 32 + lane;
      bf16* out = dst + (size_t)row * N + n0;
#pragma unroll 1
      for (int cc = 0; cc < BN / 32; ++cc) {
        uint32_t v[32];
        tmem_ld32(tmem_base + as * BN + ((uint32_t)(quad * 32) << 16) + cc * 32, v);
        if (do_stats) {
          // BatchNorm batch statistics fused here: per-channel sum / sum of squares of the bf16-rounded outputs (what a
          // separate bn_stats pass would read back from HBM).  Each warp transposes its 32x32 chunk through smem, every
          // lane reduces one channel and adds it to the CTA-wide partial sums; one double atomic per channel per CTA at
          // the end of the kernel.
#pragma unroll
          for (int e = 0; e < 32; ++e)
            scratch[lane * 33 + e] = row < M ? __bfloat162float(__float2bfloat16_rn(__uint_as_float(v[e]))) : 0.f;
          __syncwarp();
          float s1 = 0.f, s2 = 0.f;
#pragma unroll
          for (int r = 0; r < 32; ++r) {
            const float t = scratch[r * 33 + lane];
            s1 += t;
            s2 = fmaf(t, t, s2);
          }
          cta_stats[quad * 2 * N + n0 + cc * 32 + lane] += s1;   // this warp's slot: no atomics, fixed summation order
          cta_stats[quad * 2 * N + N + n0 + cc * 32 + lane] += s2;
          __syncwarp();
        }
        if (row < M) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float f[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[q * 8 + e]);
            uint4* o = reinterpret_cast<uint4*>(out + cc * 32 + q * 8);
            if (accumulate) {
              const uint4 old = *o;
              const __nv_bfloat162* oh = reinterpret_cast<const __nv_bfloat162*>(&old);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 t = __bfloat1622float2(oh[e]);
                f[2 * e] += t.x; f[2 * e + 1] += t.y;
              }
            }
            uint4 pk;
            __nv_bfloat162* ph = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
            for (int e = 0; e < 4; ++e) ph[e] = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
            *o = pk;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty0 + 8 * as);  // this warp is done reading the accumulator stage
    }
  }
  __syncthreads();
  if (do_stats)
    for (int i = tid; i < 2 * N; i += NTHREADS) {
      double v = 0.0;
#pragma unroll
      for (int w = 0; w < 4; ++w) v += (double)cta_stats[w * 2 * N + i];
      if (v != 0.0) atomicAdd(stats + i, v);
    }
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------ dgrad, stride 2
// Decomposed by output parity class (a,b) = (h & 1, w & 1) -> blockIdx.z: dx[2i+a, 2j+b] only receives the taps with
// r = (a+pad) mod 2, s = (b+pad) mod 2 (1+2+2+4 = 9 taps for 3x3/pad 1, 1 tap for 1x1/pad 0: no zero-stuffed work), and
// each class is a plain stride-1 correlation of dy with row offset (a+pad-r)/2 in {0,1} -- an im2col TMA load.
template <int BN, int STAGES>
__global__ void __launch_bounds__(NTHREADS, 1)
dgrad_s2_tma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, Geo p, bf16* __restrict__ dst,
                    int accumulate) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment as an OFFSET from the __shared__ array: the pointer keeps its address space, so the epilogue scratch
  // compiles to STS / LDS (a round trip through uintptr_t made nvcc emit generic ST.E / LD.E)
  uint8_t* smem = smem_raw + ((1024u - ((uint32_t)__cvta_generic_to_shared(smem_raw) & 1023u)) & 1023u);
  constexpr int A_BYTES = TILE_BYTES, B_BYTES = BN * 128, STAGE_BYTES = A_BYTES + B_BYTES;
  const uint32_t s_base = smem_u32(smem);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * STAGES, accum_bar = full0 + 16 * STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int ca = blockIdx.z >> 1, cb = blockIdx.z & 1;       // parity class of (h, w)
  const int H2 = p.H / 2, W2 = p.W / 2;                      // class-pixel grid (== Ho x Wo)
  const int M = p.B * H2 * W2;
  const int N = p.C;
  const int cblocks = p.K / 64;
  const int r_first = (ca + p.pad) & 1, s_first = (cb + p.pad) & 1;
  const int nr = r_first < p.R ? (p.R - r_first + 1) / 2 : 0, ns = s_first < p.S ? (p.S - s_first + 1) / 2 : 0;
  const int nkb = nr * ns * cblocks;
  const int m0 = blockIdx.x * 128, n0 = blockIdx.y * BN;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;

  if (tid == 0) {
    const int nb = m0 / (H2 * W2);
    const int rem = m0 - nb * H2 * W2;
    const int pi = rem / W2, pj = rem - pi * W2;
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % STAGES;
      mbar_wait(empty0 + 8 * s, ((kb / STAGES) & 1) ^ 1);
      const uint32_t a_tile = s_base + s * STAGE_BYTES, b_tile = a_tile + A_BYTES;
      const int ti = kb / cblocks, k0 = (kb - ti * cblocks) * 64;
      const int ri = ti / ns, si = ti - ri * ns;
      const int r = r_first + 2 * ri, sx = s_first + 2 * si;
      const int off_h = (ca + p.pad - r) / 2, off_w = (cb + p.pad - sx) / 2;  // in {0, 1} for 3x3/p1 and 1x1/p0
      mbar_expect_tx(full0 + 8 * s, STAGE_BYTES);
      tma_load_im2col(a_tile, &tmA, full0 + 8 * s, k0, pj, pi, nb, (uint16_t)off_w, (uint16_t)off_h);
      tma_load_2d(b_tile, &tmB, full0 + 8 * s, (r * p.S + sx) * p.K + k0, n0);
    }
  } else if (tid == 32) {
    constexpr uint32_t idesc = make_idesc(128, BN, 0, 0);
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % STAGES;
      mbar_wait(full0 + 8 * s, (kb / STAGES) & 1);
      tc_fence_after();
      const uint32_t a_tile = s_base + s * STAGE_BYTES, b_tile = a_tile + A_BYTES;
      const uint64_t adesc = make_desc(a_tile, 16, 1024), bdesc = make_desc(b_tile, 16, 1024);
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16(tmem_d, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0);
      umma_commit(empty0 + 8 * s);
    }
    if (nkb > 0) umma_commit(accum_bar);
  } else if (warp >= 2) {
    const int quad = warp & 3;
    const int m = m0 + quad * 32 + (tid & 31);
    const bool row_ok = m < M;
    const int mm = row_ok ? m : 0;
    const int nb = mm / (H2 * W2);
    const int rem = mm - nb * H2 * W2;
    const int pi = rem / W2, pj = rem - pi * W2;
    bf16* out = dst + (((size_t)nb * p.H + 2 * pi + ca) * p.W + 2 * pj + cb) * N + n0;
    if (nkb == 0) {
      if (!accumulate && row_ok) {  // class receives no tap (1x1 / stride 2): gradient is exactly zero there
        for (int q = 0; q < BN / 8; ++q) reinterpret_cast<uint4*>(out)[q] = make_uint4(0, 0, 0, 0);
      }
    } else {
      mbar_wait(accum_bar, 0);
      tc_fence_after();
#pragma unroll 1
      for (int cc = 0; cc < BN / 32; ++cc) {
        uint32_t v[32];
        tmem_ld32(tmem_d + ((uint32_t)(quad * 32) << 16) + cc * 32, v);
        if (row_ok) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float f[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[q * 8 + e]);
            uint4* o = reinterpret_cast<uint4*>(out + cc * 32 + q * 8);
            if (accumulate) {
              const uint4 old = *o;
              const __nv_bfloat162* oh = reinterpret_cast<const __nv_bfloat162*>(&old);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 t = __bfloat1622float2(oh[e]);
                f[2 * e] += t.x; f[2 * e + 1] += t.y;
              }
            }
            uint4 pk;
            __nv_bfloat162* ph = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
            for (int e = 0; e < 4; ++e) ph[e] = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
            *o = pk;
          }
        }
      }
      tc_fence_before();
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_d, BN);
  }
}

// ------------------------------------------------------------------------------------------------ wgrad
// D[kg (128 rows = two 64-wide (tap,c) blocks), n (BN couts)] += sum over the split's pixels.
// A = im2col(x) (MN-major), B = dy [M][K] (MN-major).  grid: (ceil(Kg/128), K/BN, splits)
// PS ("per sample", DP-SGD): blockIdx.z is the SAMPLE; the CTA reduces over that image's pixels only and STORES its tile to
// dw[sample][K][Kg] (no atomics, nothing to clear).  tmB is then a 3-D map {K, Ho*Wo, B} so that the rows past the image's last
// pixel are out-of-bounds zero fill: the im2col operand may run on into the next image, its partner dy rows are zero.
template <int BN, int STAGES, bool PS = false>
__global__ void __launch_bounds__(NTHREADS, 1)
wgrad_tma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, Geo p, float* __restrict__ dw,
                 int pix_per_split, double* __restrict__ norm2 = nullptr) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment as an OFFSET from the __shared__ array: the pointer keeps its address space, so the epilogue scratch
  // compiles to STS / LDS (a round trip through uintptr_t made nvcc emit generic ST.E / LD.E)
  uint8_t* smem = smem_raw + ((1024u - ((uint32_t)__cvta_generic_to_shared(smem_raw) & 1023u)) & 1023u);
  constexpr int A_BYTES = 2 * TILE_BYTES, B_SUB = BN / 64, B_BYTES = B_SUB * TILE_BYTES, STAGE_BYTES = A_BYTES + B_BYTES;
  const uint32_t s_base = smem_u32(smem);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * STAGES, accum_bar = full0 + 16 * STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int M = p.B * p.Ho * p.Wo;
  const int Kg = p.R * p.S * p.C;
  const int kg0 = blockIdx.x * 128, n0 = blockIdx.y * BN;
  const int pm0 = blockIdx.z * (PS ? p.Ho * p.Wo : pix_per_split);
  const int pm1 = PS ? pm0 + p.Ho * p.Wo : min(M, pm0 + pix_per_split);
  const int nsteps = pm1 > pm0 ? (pm1 - pm0 + 127) / 128 : 0;
  const bool second = kg0 + 64 < Kg;  // the upper 64 rows of the tile exist

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;

  if (tid == 0) {
    const int tap0 = kg0 / p.C, c00 = kg0 - tap0 * p.C;
    const int tap1 = (kg0 + 64) / p.C, c01 = (kg0 + 64) - tap1 * p.C;
    const int r0 = tap0 / p.S, s0 = tap0 - r0 * p.S, r1 = tap1 / p.S, s1 = tap1 - r1 * p.S;
    const uint32_t tx = (second ? 2 : 1) * TILE_BYTES + B_BYTES;
    for (int st = 0; st < nsteps; ++st) {
      const int s = st % STAGES;
      mbar_wait(empty0 + 8 * s, ((st / STAGES) & 1) ^ 1);
      const uint32_t a_tile = s_base + s * STAGE_BYTES, b_tile = a_tile + A_BYTES;
      const int m = pm0 + st * 128;
      const int nb = m / (p.Ho * p.Wo);
      const int rem = m - nb * p.Ho * p.Wo;
      const int py = rem / p.Wo, px = rem - py * p.Wo;
      const int bw = px * p.stride - p.pad, bh = py * p.stride - p.pad;
      mbar_expect_tx(full0 + 8 * s, tx);
      tma_load_im2col(a_tile, &tmA, full0 + 8 * s, c00, bw, bh, nb, (uint16_t)s0, (uint16_t)r0);
      if (second) tma_load_im2col(a_tile + TILE_BYTES, &tmA, full0 + 8 * s, c01, bw, bh, nb, (uint16_t)s1, (uint16_t)r1);
#pragma unroll
      for (int sb = 0; sb < B_SUB; ++sb) {
        if (PS) tma_load_3d(b_tile + sb * TILE_BYTES, &tmB, full0 + 8 * s, n0 + sb * 64, st * 128, blockIdx.z);
        else tma_load_2d(b_tile + sb * TILE_BYTES, &tmB, full0 + 8 * s, n0 + sb * 64, m);
      }
    }
  } else if (tid == 32) {
    constexpr uint32_t idesc = make_idesc(128, BN, 1, 1);
    for (int st = 0; st < nsteps; ++st) {
      const int s = st % STAGES;
      mbar_wait(full0 + 8 * s, (st / STAGES) & 1);
      tc_fence_after();
      const uint32_t a_tile = s_base + s * STAGE_BYTES, b_tile = a_tile + A_BYTES;
      const uint64_t adesc = make_desc(a_tile, TILE_BYTES, 1024), bdesc = make_desc(b_tile, TILE_BYTES, 1024);
#pragma unroll
      for (int k = 0; k < 8; ++k) umma_bf16(tmem_d, adesc + (uint64_t)(k * 128), bdesc + (uint64_t)(k * 128), idesc, (st | k) != 0);
      umma_commit(empty0 + 8 * s);
    }
    if (nsteps > 0) umma_commit(accum_bar);
  } else if (warp >= 2 && nsteps > 0) {
    const int quad = warp & 3;
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    const int kg = kg0 + quad * 32 + (tid & 31);
    float sq = 0.f;  // PS: this thread's share of |dw_sample|^2 (DP-SGD per-sample norm), reduced per warp below
#pragma unroll 1
    for (int cc = 0; cc < BN / 32; ++cc) {
      uint32_t v[32];
      tmem_ld32(tmem_d + ((uint32_t)(quad * 32) << 16) + cc * 32, v);
      if (kg < Kg) {
        if (PS) {
          float* o = dw + (size_t)blockIdx.z * Kg * p.K;
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const float f = __uint_as_float(v[e]);
            o[(size_t)(n0 + cc * 32 + e) * Kg + kg] = f;
            sq = fmaf(f, f, sq);
          }
        } else {
#pragma unroll
          for (int e = 0; e < 32; ++e) atomicAdd(dw + (size_t)(n0 + cc * 32 + e) * Kg + kg, __uint_as_float(v[e]));
        }
      }
    }
    if (PS && norm2) {
      double d = (double)sq;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
      if ((tid & 31) == 0 && d != 0.0) atomicAdd(norm2 + blockIdx.z, d);
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_d, BN);
  }
}

// ------------------------------------------------------------------------------------------------ host: tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_tiled = nullptr;
static EncodeIm2colFn g_im2col = nullptr;

static bool load_driver() {
  if (g_tiled && g_im2col) return true;
  cudaDriverEntryPointQueryResult q;
  void* f = nullptr;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || !f) return false;
  g_tiled = (EncodeTiledFn)f;
  f = nullptr;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &f, cudaEnableDefault, &q) != cudaSuccess || !f) return false;
  g_im2col = (EncodeIm2colFn)f;
  return true;
}

// dense row-major bf16 matrix [rows][cols] -> box {64 cols, box_rows}, SWIZZLE_128B
static bool map_dense(CUtensorMap* tm, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  return g_tiled(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// dy [B][Ho*Wo][K] bf16 as a 3-D tensor: box {64 channels, 128 pixels of ONE image} (rows past the image: zero fill)
static bool map_rows_per_image(CUtensorMap* tm, const void* base, uint64_t B, uint64_t pix, uint64_t K) {
  cuuint64_t dims[3] = {K, pix, B};
  cuuint64_t strides[2] = {K * 2, pix * K * 2};
  cuuint32_t box[3] = {64, 128, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  return g_tiled(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// NHWC bf16 activation [B][H][W][C] as an im2col source: 128 pixels x 64 channels per load
static bool map_im2col(CUtensorMap* tm, const void* base, int B, int H, int W, int C, int lower_w, int lower_h, int upper_w,
                       int upper_h, int stride) {
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  int lower[2] = {lower_w, lower_h};
  int upper[2] = {upper_w, upper_h};
  cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
  if (g_im2col(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, lower, upper, 64, 128, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return false;
  // Known driver issue (<= 13.1) with im2col descriptors of tensors smaller than 128 KiB: the same bit CUTLASS clears
  // in cute/atom/copy_traits_sm90_im2col.hpp after cuTensorMapEncodeIm2col.
  int drv = 0;
  cudaDriverGetVersion(&drv);
  if (drv <= 13010 && (size_t)B * H * W * C * 2 < 131072) reinterpret_cast<uint64_t*>(tm)[1] &= ~(1ull << 21);
  return true;
}

template <typename Kern>
static bool set_smem(Kern k, int bytes) {
  return cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) == cudaSuccess;
}

constexpr int FSTAGES = 5;
constexpr int RSTAGES = 6;  // resident-B variant: stages hold only the 16 KB activation tiles
constexpr int smem_conv(int BN, int stages) { return stages * (TILE_BYTES + BN * 128) + 1024 + 256 + 4 * 32 * 33 * 4 + 4 * 2 * 512 * 4; }
constexpr int smem_conv_rb(int stages) { return stages * TILE_BYTES + RB_MAX_BYTES + 1024 + 256 + 4 * 32 * 33 * 4 + 4 * 2 * 512 * 4; }
constexpr int smem_wg(int BN, int stages) { return stages * (2 * TILE_BYTES + (BN / 64) * TILE_BYTES) + 1024 + 256; }

static Geo geo(const pm_conv_t* p) { return Geo{p->B, p->H, p->W, p->C, p->K, p->R, p->S, p->stride, p->pad, p->Ho, p->Wo}; }

}  // namespace tma

// Entry points used by conv_tc.cu's dispatcher.  Return 0 on success, 1 if this shape is not eligible
// (caller falls back to the cp.async variant), 2 on a CUDA/driver error.
int pm_tma_conv_fwd(const pm_conv_t* p, const void* x, const void* w, void* y, double* stats, cudaStream_t st) {
  using namespace tma;
  if (p->C % 64 != 0 || p->K % 64 != 0 || !load_driver()) return 1;
  CUtensorMap tmA, tmB;
  const int Ktot = p->R * p->S * p->C;
  if (!map_im2col(&tmA, x, p->B, p->H, p->W, p->C, -p->pad, -p->pad, p->pad - (p->S - 1), p->pad - (p->R - 1), p->stride)) return 2;
  const int M = p->B * p->Ho * p->Wo;
  if (p->K % 128 == 0) {
    if (!map_dense(&tmB, w, p->K, Ktot, 128)) return 2;
    if (!set_smem(conv_tma_kernel<128, FSTAGES, false, false>, smem_conv(128, FSTAGES))) return 2;
    dim3 grid(std::min(pm_num_sms(), ((M + 127) / 128) * (p->K / 128)));
    if (pm_launch(conv_tma_kernel<128, FSTAGES, false, false>, grid, dim3(NTHREADS), (size_t)smem_conv(128, FSTAGES), st, tmA, tmB, geo(p), (bf16*)y, 0, stats) != cudaSuccess) return 2;
  } else if (p->K == 64 && Ktot * 64 * 2 <= RB_MAX_BYTES) {
    if (!map_dense(&tmB, w, p->K, Ktot, 64)) return 2;
    if (!set_smem(conv_tma_kernel<64, RSTAGES, false, true>, smem_conv_rb(RSTAGES))) return 2;
    dim3 grid(std::min(pm_num_sms(), (M + 127) / 128));
    if (pm_launch(conv_tma_kernel<64, RSTAGES, false, true>, grid, dim3(NTHREADS), (size_t)smem_conv_rb(RSTAGES), st, tmA, tmB, geo(p), (bf16*)y, 0, stats) != cudaSuccess) return 2;
  } else {
    if (!map_dense(&tmB, w, p->K, Ktot, 64)) return 2;
    if (!set_smem(conv_tma_kernel<64, FSTAGES, false, false>, smem_conv(64, FSTAGES))) return 2;
    dim3 grid(std::min(pm_num_sms(), ((M + 127) / 128) * (p->K / 64)));
    if (pm_launch(conv_tma_kernel<64, FSTAGES, false, false>, grid, dim3(NTHREADS), (size_t)smem_conv(64, FSTAGES), st, tmA, tmB, geo(p), (bf16*)y, 0, stats) != cudaSuccess) return 2;
  }
  return 0;
}

int pm_tma_conv_dgrad(const pm_conv_t* p, const void* dy, const void* wt, void* dx, int accumulate, cudaStream_t st) {
  using namespace tma;
  if (p->C % 64 != 0 || p->K % 64 != 0 || !load_driver()) return 1;
  CUtensorMap tmA, tmB;
  const int Ktot = p->R * p->S * p->K;
  if (p->stride == 2) {
    const bool geom_ok = p->H % 2 == 0 && p->W % 2 == 0 && p->Ho == p->H / 2 && p->Wo == p->W / 2 &&
                         ((p->R == 3 && p->S == 3 && p->pad == 1) || (p->R == 1 && p->S == 1 && p->pad == 0));
    if (!geom_ok) return 1;
    if (!map_im2col(&tmA, dy, p->B, p->Ho, p->Wo, p->K, 0, 0, 0, 0, 1)) return 2;
    const int Mc = p->B * p->Ho * p->Wo;
    if (p->C % 128 == 0) {
      if (!map_dense(&tmB, wt, p->C, Ktot, 128)) return 2;
      if (!set_smem(dgrad_s2_tma_kernel<128, FSTAGES>, smem_conv(128, FSTAGES))) return 2;
      dim3 grid((Mc + 127) / 128, p->C / 128, 4);
      dgrad_s2_tma_kernel<128, FSTAGES><<<grid, NTHREADS, smem_conv(128, FSTAGES), st>>>(tmA, tmB, geo(p), (bf16*)dx, accumulate);
    } else {
      if (!map_dense(&tmB, wt, p->C, Ktot, 64)) return 2;
      if (!set_smem(dgrad_s2_tma_kernel<64, FSTAGES>, smem_conv(64, FSTAGES))) return 2;
      dim3 grid((Mc + 127) / 128, p->C / 64, 4);
      dgrad_s2_tma_kernel<64, FSTAGES><<<grid, NTHREADS, smem_conv(64, FSTAGES), st>>>(tmA, tmB, geo(p), (bf16*)dx, accumulate);
    }
    return 0;
  }
  if (p->stride != 1) return 1;
  // dgrad with stride 1 == correlation of dy with the mirrored filter: window base = out + pad - (R-1)
  const int lw = p->pad - (p->S - 1), lh = p->pad - (p->R - 1);
  const int uw = lw + (p->W - p->Wo), uh = lh + (p->H - p->Ho);
  if (!map_im2col(&tmA, dy, p->B, p->Ho, p->Wo, p->K, lw, lh, uw, uh, 1)) return 2;
  const int M = p->B * p->H * p->W;
  if (p->C % 128 == 0) {
    if (!map_dense(&tmB, wt, p->C, Ktot, 128)) return 2;
    if (!set_smem(conv_tma_kernel<128, FSTAGES, true, false>, smem_conv(128, FSTAGES))) return 2;
    dim3 grid(std::min(pm_num_sms(), ((M + 127) / 128) * (p->C / 128)));
    if (pm_launch(conv_tma_kernel<128, FSTAGES, true, false>, grid, dim3(NTHREADS), (size_t)smem_conv(128, FSTAGES), st, tmA, tmB, geo(p), (bf16*)dx, accumulate, (double*)nullptr) != cudaSuccess) return 2;
  } else if (p->C == 64 && Ktot * 64 * 2 <= RB_MAX_BYTES) {
    if (!map_dense(&tmB, wt, p->C, Ktot, 64)) return 2;
    if (!set_smem(conv_tma_kernel<64, RSTAGES, true, true>, smem_conv_rb(RSTAGES))) return 2;
    dim3 grid(std::min(pm_num_sms(), (M + 127) / 128));
    if (pm_launch(conv_tma_kernel<64, RSTAGES, true, true>, grid, dim3(NTHREADS), (size_t)smem_conv_rb(RSTAGES), st, tmA, tmB, geo(p), (bf16*)dx, accumulate, (double*)nullptr) != cudaSuccess) return 2;
  } else {
    if (!map_dense(&tmB, wt, p->C, Ktot, 64)) return 2;
    if (!set_smem(conv_tma_kernel<64, FSTAGES, true, false>, smem_conv(64, FSTAGES))) return 2;
    dim3 grid(std::min(pm_num_sms(), ((M + 127) / 128) * (p->C / 64)));
    if (pm_launch(conv_tma_kernel<64, FSTAGES, true, false>, grid, dim3(NTHREADS), (size_t)smem_conv(64, FSTAGES), st, tmA, tmB, geo(p), (bf16*)dx, accumulate, (double*)nullptr) != cudaSuccess) return 2;
  }
  return 0;
}

int pm_tma_conv_wgrad(const pm_conv_t* p, const void* x, const void* dy, float* dw, cudaStream_t st) {
  using namespace tma;
  if (p->C % 64 != 0 || p->K % 64 != 0 || !load_driver()) return 1;
  CUtensorMap tmA, tmB;
  const int M = p->B * p->Ho * p->Wo;
  const int Kg = p->R * p->S * p->C;
  if (!map_im2col(&tmA, x, p->B, p->H, p->W, p->C, -p->pad, -p->pad, p->pad - (p->S - 1), p->pad - (p->R - 1), p->stride)) return 2;
  if (!map_dense(&tmB, dy, (uint64_t)M, p->K, 128)) return 2;
  const int BN = p->K % 128 == 0 ? 128 : 64;
  const long tiles = ((Kg + 127) / 128) * (long)(p->K / BN);
  long splits = (2L * pm_num_sms() + tiles - 1) / tiles;
  const long max_splits = (M + 511) / 512;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  const int pps = (int)(((M + splits - 1) / splits + 127) / 128 * 128);
  dim3 grid((Kg + 127) / 128, p->K / BN, (unsigned)splits);
  if (BN == 128) {
    if (!set_smem(wgrad_tma_kernel<128, 3>, smem_wg(128, 3))) return 2;
    wgrad_tma_kernel<128, 3><<<grid, NTHREADS, smem_wg(128, 3), st>>>(tmA, tmB, geo(p), dw, pps);
  } else {
    if (!set_smem(wgrad_tma_kernel<64, 4>, smem_wg(64, 4))) return 2;
    wgrad_tma_kernel<64, 4><<<grid, NTHREADS, smem_wg(64, 4), st>>>(tmA, tmB, geo(p), dw, pps);
  }
  return 0;
}

// DP-SGD: per-sample weight gradients dw[b][K][R*S*C] = sum over image b's pixels (written, not accumulated).
int pm_tma_conv_wgrad_persample(const pm_conv_t* p, const void* x, const void* dy, float* dw, double* norm2, cudaStream_t st) {
  using namespace tma;
  if (p->C % 64 != 0 || p->K % 64 != 0 || !load_driver()) return 1;
  CUtensorMap tmA, tmB;
  const int Kg = p->R * p->S * p->C;
  if (!map_im2col(&tmA, x, p->B, p->H, p->W, p->C, -p->pad, -p->pad, p->pad - (p->S - 1), p->pad - (p->R - 1), p->stride)) return 2;
  if (!map_rows_per_image(&tmB, dy, (uint64_t)p->B, (uint64_t)p->Ho * p->Wo, (uint64_t)p->K)) return 2;
  const int BN = p->K % 128 == 0 ? 128 : 64;
  if (p->B > 65535) return 1;
  dim3 grid((Kg + 127) / 128, p->K / BN, (unsigned)p->B);
  if (BN == 128) {
    if (!set_smem(wgrad_tma_kernel<128, 3, true>, smem_wg(128, 3))) return 2;
    wgrad_tma_kernel<128, 3, true><<<grid, NTHREADS, smem_wg(128, 3), st>>>(tmA, tmB, geo(p), dw, 0, norm2);
  } else {
    if (!set_smem(wgrad_tma_kernel<64, 4, true>, smem_wg(64, 4))) return 2;
    wgrad_tma_kernel<64, 4, true><<<grid, NTHREADS, smem_wg(64, 4), st>>>(tmA, tmB, geo(p), dw, 0, norm2);
  }
  return 0;
}
