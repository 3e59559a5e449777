// Path T, throughput mode: bf16 implicit-GEMM convolutions on the 5th-generation tensor cores.
//
//   * operands are staged in shared memory in the canonical UMMA SWIZZLE_128B layout
//     ([rows][64 bf16 = 128 B], 16-byte chunk index XOR (row & 7), 1024-byte aligned tiles),
//   * tcgen05.mma (cta_group::1, kind::f16, M = 128) is issued by one elected thread, fp32 accumulators
//     live in TMEM and are read back with tcgen05.ld for the epilogue,
//   * a 4-stage mbarrier ring overlaps the im2col gather (cp.async, zero-fill for padding) with the MMAs;
//     tcgen05.commit releases smem stages and publishes the accumulator.
//
//   fwd  : D[pix, cout]  = A[pix, (r,s,c)]      (gather, K-major)  x  W[cout, (r,s,c)]   (dense, K-major)
//   dgrad: D[ipix, cin]  = A[ipix, (r,s,k)]     (gather of dy)     x  Wt[cin, (r,s,k)]   (dense, K-major)
//   wgrad: D[(r,s,c), k] = A[pix, (r,s,c)]^T    (gather, MN-major) x  dy[pix, k]         (dense, MN-major)
//          reduction over pixels, split across CTAs, fp32 red.add into dW.
#include "common.cuh"
#include <stdlib.h>

namespace {

typedef __nv_bfloat16 bf16;

constexpr int STAGES = 4;
constexpr int LAG = 2;                 // producer arrives on full[] this many k-steps late (cp.async groups in flight)
constexpr int NPROD = 128;             // warps 0-3: producers, then epilogue
constexpr int NTHREADS = 160;          // + warp 4: TMEM alloc + MMA issue
constexpr int TILE_BYTES = 128 * 128;  // one [128 rows][128 B] sub-tile

struct TcP {
  int B, H, W, C, K, R, S, stride, pad, Ho, Wo;  // conv geometry (C = padded input channels)
};

// ------------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
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
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const uint32_t sz = valid ? 16u : 0u;  // src-size 0 => 16 bytes of zero fill, no global read
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
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

// UMMA shared-memory descriptor, SWIZZLE_128B (cute::UMMA::SmemDescriptor): start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), layout_type=2 [61,64)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D=f32 [4,6)=1, A=bf16 [7,10)=1, B=bf16 [10,13)=1,
// a_major bit 15, b_major bit 16 (1 = MN-major), N>>3 [17,23), M>>4 [24,29)
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ uint32_t swz(int row, int chunk) { return (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4)); }

// ------------------------------------------------------------------------------------------------ loaders
// Every producer thread owns chunk column j = tid & 7 and rows (tid >> 3) + 16*i, i = 0..7 of a [128][128B] sub-tile.

// dense row-major bf16 matrix [nrows][ld]: tile rows r0.., element columns c0 + j*8 .. +8
__device__ __forceinline__ void load_dense(uint32_t tile, const bf16* __restrict__ base, int ld, int nrows, int ncols, int r0,
                                           int c0, int tid, int rows_in_tile) {
  const int j = tid & 7, rb = tid >> 3;
  const int col = c0 + j * 8;
  const bool col_ok = col < ncols;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = rb + 16 * i;
    if (r >= rows_in_tile) break;
    const int gr = r0 + r;
    const bool ok = col_ok && gr < nrows;
    const bf16* src = ok ? base + (size_t)gr * ld + col : base;
    cp_async16(tile + swz(r, j), src, ok);
  }
}

struct Pix {  // decoded output-space pixel of one tile row
  int b, y, x;
  bool ok;
};

// MODE 0 (fwd, wgrad): src = x [B,H,W,C], rows are output pixels, tap (r,s) reads (y*st - pad + r, x*st - pad + s)
// MODE 1 (dgrad)     : src = dy [B,Ho,Wo,K], rows are input pixels, tap reads ((y + pad - r)/st, (x + pad - s)/st)
template <int MODE>
__device__ __forceinline__ void load_gather(uint32_t tile, const bf16* __restrict__ src, const TcP& p, const Pix pix[8],
                                            int ke0, int Ktot, int tid) {
  const int j = tid & 7, rb = tid >> 3;
  const int CR = MODE == 1 ? p.K : p.C;  // channels of the gathered tensor
  const int SH = MODE == 1 ? p.Ho : p.H, SW = MODE == 1 ? p.Wo : p.W;
  const int ke = ke0 + j * 8;
  const bool k_ok = ke < Ktot;
  const int tap = ke / CR, c = ke - tap * CR;
  const int r = tap / p.S, s = tap - r * p.S;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = rb + 16 * i;
    bool ok = k_ok && pix[i].ok;
    int sy, sx;
    if (MODE == 0) {
      sy = pix[i].y * p.stride - p.pad + r;
      sx = pix[i].x * p.stride - p.pad + s;
      ok = ok && sy >= 0 && sy < SH && sx >= 0 && sx < SW;
    } else {
      const int ty = pix[i].y + p.pad - r, tx = pix[i].x + p.pad - s;
      ok = ok && ty >= 0 && tx >= 0;
      if (p.stride == 1) {
        sy = ty; sx = tx;
      } else {
        ok = ok && (ty % p.stride == 0) && (tx % p.stride == 0);
        sy = ty / p.stride; sx = tx / p.stride;
      }
      ok = ok && sy < SH && sx < SW;
    }
    const bf16* g = ok ? src + (((size_t)pix[i].b * SH + sy) * SW + sx) * CR + c : src;
    cp_async16(tile + swz(row, j), g, ok);
  }
}

// ------------------------------------------------------------------------------------------------ fwd / dgrad kernel
// grid: (ceil(M/128), N/BN).  MODE 0 fwd, 1 dgrad.
template <int MODE, int BN>
__global__ void __launch_bounds__(NTHREADS, 1)
conv_tc_kernel(TcP p, const bf16* __restrict__ src, const bf16* __restrict__ wmat, bf16* __restrict__ dst, int accumulate) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment as an OFFSET from the __shared__ array: the pointer keeps its address space, so the epilogue scratch
  // compiles to STS / LDS (a round trip through uintptr_t made nvcc emit generic ST.E / LD.E)
  uint8_t* smem = smem_raw + ((1024u - ((uint32_t)__cvta_generic_to_shared(smem_raw) & 1023u)) & 1023u);
  constexpr int A_BYTES = TILE_BYTES, B_BYTES = BN * 128;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  const uint32_t s_base = smem_u32(smem);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * STAGES, accum_bar = full0 + 16 * STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int OH = MODE == 0 ? p.Ho : p.H, OW = MODE == 0 ? p.Wo : p.W;
  const int CR = MODE == 0 ? p.C : p.K;
  const int N = MODE == 0 ? p.K : p.C;
  const int M = p.B * OH * OW;
  const int Ktot = p.R * p.S * CR;
  const int nkb = (Ktot + 63) / 64;
  const int m0 = blockIdx.x * 128, n0 = blockIdx.y * BN;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full0 + 8 * s, NPROD);
      mbar_init(empty0 + 8 * s, 1);
    }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) tmem_alloc(smem_u32(tmem_slot), BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;

  if (warp < 4) {
    // ------------------------------------------------------------ producers
    Pix pix[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int m = m0 + (tid >> 3) + 16 * i;
      pix[i].ok = m < M;
      const int mm = pix[i].ok ? m : 0;
      pix[i].b = mm / (OH * OW);
      const int rem = mm - pix[i].b * OH * OW;
      pix[i].y = rem / OW;
      pix[i].x = rem - pix[i].y * OW;
    }
    for (int kb = 0; kb < nkb + LAG; ++kb) {
      if (kb < nkb) {
        const int s = kb % STAGES;
        mbar_wait(empty0 + 8 * s, ((kb / STAGES) & 1) ^ 1);
        const uint32_t a_tile = s_base + s * STAGE_BYTES, b_tile = a_tile + A_BYTES;
        load_gather<MODE>(a_tile, src, p, pix, kb * 64, Ktot, tid);
        load_dense(b_tile, wmat, Ktot, N, Ktot, n0, kb * 64, tid, BN);
      }
      cp_async_commit();
      if (kb >= LAG) {
        cp_async_wait<LAG>();
        fence_proxy_async();
        mbar_arrive(full0 + 8 * ((kb - LAG) % STAGES));
      }
    }
    // ------------------------------------------------------------ epilogue
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    const int row = m0 + warp * 32 + (tid & 31);
    bf16* out = dst + (size_t)row * N + n0;
#pragma unroll 1
    for (int cc = 0; cc < BN / 32; ++cc) {
      uint32_t v[32];
      tmem_ld32(tmem_d + ((uint32_t)(warp * 32) << 16) + cc * 32, v);
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
  } else if (tid == 128) {
    // ------------------------------------------------------------ MMA issuer (one thread)
    constexpr uint32_t idesc = make_idesc(128, BN, 0, 0);
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % STAGES;
      mbar_wait(full0 + 8 * s, (kb / STAGES) & 1);
      tc_fence_after();
      const uint32_t a_tile = s_base + s * STAGE_BYTES, b_tile = a_tile + A_BYTES;
      const uint64_t adesc = make_desc(a_tile, 16, 1024), bdesc = make_desc(b_tile, 16, 1024);
#pragma unroll
      for (int k = 0; k < 4; ++k)  // UMMA_K = 16 bf16 = 32 bytes inside the 128-byte swizzle atom
        umma_bf16(tmem_d, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0);
      umma_commit(empty0 + 8 * s);  // frees the smem stage once these MMAs have read it
    }
    umma_commit(accum_bar);
  }
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_d, BN);
  }
}

// ------------------------------------------------------------------------------------------------ wgrad kernel
// D[kg (128 rows), n (BN cols)] += sum over a pixel range.  grid: (ceil(Kg/128), K/BN, splits)
template <int BN, int WSTAGES>
__global__ void __launch_bounds__(NTHREADS, 1)
wgrad_tc_kernel(TcP p, const bf16* __restrict__ x, const bf16* __restrict__ dy, float* __restrict__ dw, int pix_per_split) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment as an OFFSET from the __shared__ array: the pointer keeps its address space, so the epilogue scratch
  // compiles to STS / LDS (a round trip through uintptr_t made nvcc emit generic ST.E / LD.E)
  uint8_t* smem = smem_raw + ((1024u - ((uint32_t)__cvta_generic_to_shared(smem_raw) & 1023u)) & 1023u);
  constexpr int A_BYTES = 2 * TILE_BYTES, B_SUB = BN / 64, B_BYTES = B_SUB * TILE_BYTES;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  const uint32_t s_base = smem_u32(smem);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + WSTAGES * STAGE_BYTES);
  const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * WSTAGES, accum_bar = full0 + 16 * WSTAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * WSTAGES + 1);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int M = p.B * p.Ho * p.Wo;
  const int Kg = p.R * p.S * p.C;
  const int kg0 = blockIdx.x * 128, n0 = blockIdx.y * BN;
  const int pm0 = blockIdx.z * pix_per_split;
  const int pm1 = min(M, pm0 + pix_per_split);
  const int nsteps = pm1 > pm0 ? (pm1 - pm0 + 127) / 128 : 0;

  if (tid == 0) {
    for (int s = 0; s < WSTAGES; ++s) {
      mbar_init(full0 + 8 * s, NPROD);
      mbar_init(empty0 + 8 * s, 1);
    }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) tmem_alloc(smem_u32(tmem_slot), BN < 32 ? 32 : BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;

  if (warp < 4) {
    for (int st = 0; st < nsteps + LAG; ++st) {
      if (st < nsteps) {
        const int s = st % WSTAGES;
        mbar_wait(empty0 + 8 * s, ((st / WSTAGES) & 1) ^ 1);
        const uint32_t a_tile = s_base + s * STAGE_BYTES, b_tile = a_tile + A_BYTES;
        Pix pix[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int m = pm0 + st * 128 + (tid >> 3) + 16 * i;
          pix[i].ok = m < pm1;
          const int mm = pix[i].ok ? m : 0;
          pix[i].b = mm / (p.Ho * p.Wo);
          const int rem = mm - pix[i].b * p.Ho * p.Wo;
          pix[i].y = rem / p.Wo;
          pix[i].x = rem - pix[i].y * p.Wo;
        }
        load_gather<0>(a_tile, x, p, pix, kg0, Kg, tid);                    // kg0 .. kg0+63
        load_gather<0>(a_tile + TILE_BYTES, x, p, pix, kg0 + 64, Kg, tid);  // kg0+64 .. kg0+127
#pragma unroll
        for (int sb = 0; sb < B_SUB; ++sb)
          load_dense(b_tile + sb * TILE_BYTES, dy, p.K, pm1, p.K, pm0 + st * 128, n0 + sb * 64, tid, 128);
      }
      cp_async_commit();
      if (st >= LAG) {
        cp_async_wait<LAG>();
        fence_proxy_async();
        mbar_arrive(full0 + 8 * ((st - LAG) % WSTAGES));
      }
    }
    if (nsteps > 0) {
      mbar_wait(accum_bar, 0);
      tc_fence_after();
      const int kg = kg0 + warp * 32 + (tid & 31);
#pragma unroll 1
      for (int cc = 0; cc < BN / 32; ++cc) {
        uint32_t v[32];
        tmem_ld32(tmem_d + ((uint32_t)(warp * 32) << 16) + cc * 32, v);
        if (kg < Kg) {
#pragma unroll
          for (int e = 0; e < 32; ++e) atomicAdd(dw + (size_t)(n0 + cc * 32 + e) * Kg + kg, __uint_as_float(v[e]));
        }
      }
      tc_fence_before();
    }
  } else if (tid == 128) {
    constexpr uint32_t idesc = make_idesc(128, BN, 1, 1);  // both operands MN-major
    for (int st = 0; st < nsteps; ++st) {
      const int s = st % WSTAGES;
      mbar_wait(full0 + 8 * s, (st / WSTAGES) & 1);
      tc_fence_after();
      const uint32_t a_tile = s_base + s * STAGE_BYTES, b_tile = a_tile + A_BYTES;
      // MN-major SW128: 64 MN-elements per 128-byte row, LBO = distance between 64-element MN blocks (one sub-tile),
      // SBO = 8 k-rows = 1024 bytes; one MMA consumes 16 k-rows (pixels) = 2048 bytes.
      const uint64_t adesc = make_desc(a_tile, TILE_BYTES, 1024), bdesc = make_desc(b_tile, TILE_BYTES, 1024);
#pragma unroll
      for (int k = 0; k < 8; ++k)
        umma_bf16(tmem_d, adesc + (uint64_t)(k * 128), bdesc + (uint64_t)(k * 128), idesc, (st | k) != 0);
      umma_commit(empty0 + 8 * s);
    }
    if (nsteps > 0) umma_commit(accum_bar);
  }
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_d, BN < 32 ? 32 : BN);
  }
}

TcP to_tc(const pm_conv_t* p) { return TcP{p->B, p->H, p->W, p->C, p->K, p->R, p->S, p->stride, p->pad, p->Ho, p->Wo}; }

bool tc_ok(const pm_conv_t* p) {
  return p && p->B > 0 && p->C % 8 == 0 && p->K % 64 == 0 && p->stride > 0 && p->pad >= 0 &&
         p->Ho == (p->H + 2 * p->pad - p->R) / p->stride + 1 && p->Wo == (p->W + 2 * p->pad - p->S) / p->stride + 1;
}

template <typename Kern>
int set_smem(Kern k, int bytes) {
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) return pm_set_err(__FILE__, __LINE__, cudaGetErrorString(e));
  return 0;
}

constexpr int smem_fwd(int BN) { return STAGES * (TILE_BYTES + BN * 128) + 1024 + 256; }
constexpr int wgrad_stages(int BN) { return BN == 128 ? 3 : 4; }
constexpr int smem_wgrad(int BN) { return wgrad_stages(BN) * (2 * TILE_BYTES + (BN / 64) * TILE_BYTES) + 1024 + 256; }

template <int MODE>
int launch_conv(const pm_conv_t* p, const void* src, const void* w, void* dst, int accumulate, cudaStream_t st) {
  const int OHW = MODE == 0 ? p->Ho * p->Wo : p->H * p->W;
  const int N = MODE == 0 ? p->K : p->C;
  const int M = p->B * OHW;
  if (N % 128 == 0) {
    if (set_smem(conv_tc_kernel<MODE, 128>, smem_fwd(128))) return PM_ECUDA;
    dim3 grid((M + 127) / 128, N / 128);
    conv_tc_kernel<MODE, 128><<<grid, NTHREADS, smem_fwd(128), st>>>(to_tc(p), (const bf16*)src, (const bf16*)w, (bf16*)dst, accumulate);
  } else {
    if (set_smem(conv_tc_kernel<MODE, 64>, smem_fwd(64))) return PM_ECUDA;
    dim3 grid((M + 127) / 128, N / 64);
    conv_tc_kernel<MODE, 64><<<grid, NTHREADS, smem_fwd(64), st>>>(to_tc(p), (const bf16*)src, (const bf16*)w, (bf16*)dst, accumulate);
  }
  return 0;
}

int tc_wgrad_splits(const pm_conv_t* p, int BN) {
  const long M = (long)p->B * p->Ho * p->Wo;
  const long Kg = (long)p->R * p->S * p->C;
  const long tiles = ((Kg + 127) / 128) * (p->K / BN);
  long splits = (2L * pm_num_sms() + tiles - 1) / tiles;
  const long max_splits = (M + 511) / 512;  // >= 4 steps of 128 pixels per split
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  return (int)splits;
}

}  // namespace

// TMA-fed variants (conv_tma.cu): 0 = launched, 1 = shape not eligible (use the cp.async variant), 2 = error
int pm_tma_conv_fwd(const pm_conv_t* p, const void* x, const void* w, void* y, double* stats, cudaStream_t st);
int pm_tma_conv_dgrad(const pm_conv_t* p, const void* dy, const void* wt, void* dx, int accumulate, cudaStream_t st);
int pm_tma_conv_wgrad(const pm_conv_t* p, const void* x, const void* dy, float* dw, cudaStream_t st);
// halo-strip variant for 3x3 / stride 1 / pad 1 (conv_halo.cu), same return convention
int pm_halo_conv(const pm_conv_t* p, const void* src, const void* wmat, void* dst, int accumulate, double* stats, int flip,
                 cudaStream_t st);
// persistent stride-2 data gradient (conv_s2.cu), same return convention
int pm_s2p_conv_dgrad(const pm_conv_t* p, const void* dy, const void* wt, void* dx, int accumulate, cudaStream_t st);
// experimental halo-strip weight gradient (wgrad_halo.cu; only when PRIMIA_HALO_WGRAD=1), same return convention
int pm_halo_conv_wgrad(const pm_conv_t* p, const void* x, const void* dy, float* dw, cudaStream_t st);
int pm_tma_conv_wgrad_persample(const pm_conv_t* p, const void* x, const void* dy, float* dw, double* norm2, cudaStream_t st);
static bool use_tma() {
  const char* e = getenv("PRIMIA_NO_TMA");
  return !(e && e[0] == '1');
}

extern "C" {

int pm_conv_fwd_bf16(const pm_conv_t* p, const void* x, const void* w, void* y, double* stats, pm_stream_t s) {
  PM_CHECK_ARG(tc_ok(p) && x && w && y);
  if (use_tma()) {
    const int rh = pm_halo_conv(p, x, w, y, 0, stats, 0, S(s));
    if (rh == 2) return pm_set_err(__FILE__, __LINE__, "halo conv fwd setup failed");
    if (rh == 0) PM_LAUNCH_OK();
    const int r = pm_tma_conv_fwd(p, x, w, y, stats, S(s));  // statistics fused into the epilogue
    if (r == 2) return pm_set_err(__FILE__, __LINE__, "TMA conv fwd setup failed");
    if (r == 0) PM_LAUNCH_OK();
  }
  if (launch_conv<0>(p, x, w, y, 0, S(s))) return PM_ECUDA;
  if (stats) return pm_bn_stats_bf16(y, (size_t)p->B * p->Ho * p->Wo, p->K, stats, s);  // cp.async variant: separate pass
  PM_LAUNCH_OK();
}

int pm_conv_dgrad_bf16(const pm_conv_t* p, const void* dy, const void* wt, void* dx, int accumulate, pm_stream_t s) {
  PM_CHECK_ARG(tc_ok(p) && dy && wt && dx && p->C % 64 == 0);
  if (use_tma()) {
    const int rh = pm_halo_conv(p, dy, wt, dx, accumulate, nullptr, 1, S(s));
    if (rh == 2) return pm_set_err(__FILE__, __LINE__, "halo conv dgrad setup failed");
    if (rh == 0) PM_LAUNCH_OK();
    const int rs = pm_s2p_conv_dgrad(p, dy, wt, dx, accumulate, S(s));
    if (rs == 2) return pm_set_err(__FILE__, __LINE__, "stride-2 conv dgrad setup failed");
    if (rs == 0) PM_LAUNCH_OK();
    const int r = pm_tma_conv_dgrad(p, dy, wt, dx, accumulate, S(s));
    if (r == 2) return pm_set_err(__FILE__, __LINE__, "TMA conv dgrad setup failed");
    if (r == 0) PM_LAUNCH_OK();
  }
  if (launch_conv<1>(p, dy, wt, dx, accumulate, S(s))) return PM_ECUDA;
  PM_LAUNCH_OK();
}

int pm_conv_wgrad_bf16(const pm_conv_t* p, const void* x, const void* dy, float* dw, void* ws, pm_stream_t s) {
  PM_CHECK_ARG(tc_ok(p) && x && dy && dw);
  (void)ws;
  if (use_tma()) {
    const int rh = pm_halo_conv_wgrad(p, x, dy, dw, S(s));
    if (rh == 2) return pm_set_err(__FILE__, __LINE__, "halo conv wgrad setup failed");
    if (rh == 0) PM_LAUNCH_OK();
    const int r = pm_tma_conv_wgrad(p, x, dy, dw, S(s));
    if (r == 2) return pm_set_err(__FILE__, __LINE__, "TMA conv wgrad setup failed");
    if (r == 0) PM_LAUNCH_OK();
  }
  const int M = p->B * p->Ho * p->Wo;
  const int Kg = p->R * p->S * p->C;
  if (p->K % 128 == 0) {
    const int splits = tc_wgrad_splits(p, 128);
    int pps = ((M + splits - 1) / splits + 127) / 128 * 128;
    if (set_smem(wgrad_tc_kernel<128, 3>, smem_wgrad(128))) return PM_ECUDA;
    dim3 grid((Kg + 127) / 128, p->K / 128, splits);
    wgrad_tc_kernel<128, 3><<<grid, NTHREADS, smem_wgrad(128), S(s)>>>(to_tc(p), (const bf16*)x, (const bf16*)dy, dw, pps);
  } else {
    const int splits = tc_wgrad_splits(p, 64);
    int pps = ((M + splits - 1) / splits + 127) / 128 * 128;
    if (set_smem(wgrad_tc_kernel<64, 4>, smem_wgrad(64))) return PM_ECUDA;
    dim3 grid((Kg + 127) / 128, p->K / 64, splits);
    wgrad_tc_kernel<64, 4><<<grid, NTHREADS, smem_wgrad(64), S(s)>>>(to_tc(p), (const bf16*)x, (const bf16*)dy, dw, pps);
  }
  PM_LAUNCH_OK();
}

/* DP-SGD: per-sample weight gradients (TMA / tcgen05 path only: there is no fallback kernel for other channel counts) */
int pm_conv_wgrad_persample_bf16(const pm_conv_t* p, const void* x, const void* dy, float* dw, double* norm2, pm_stream_t s) {
  PM_CHECK_ARG(tc_ok(p) && x && dy && dw);
  const int r = pm_tma_conv_wgrad_persample(p, x, dy, dw, norm2, S(s));
  if (r == 1) return pm_set_err(__FILE__, __LINE__, "per-sample wgrad: channel counts must be multiples of 64");
  if (r == 2) return pm_set_err(__FILE__, __LINE__, "per-sample wgrad setup failed");
  PM_LAUNCH_OK();
}

}  // extern "C"
