// Path E: the Beaver-combine GEMM over Z_2^64 on the INT8 tensor cores of sm_100a (tcgen05.mma.kind::i8), exact.
//
// There is no 64-bit integer MMA.  Every int64 is 8 unsigned byte limbs, x = sum_l x_l 2^(8l), so mod 2^64
//     (A @ B)[m,n] = sum_{i+j<=7} 2^(8(i+j)) * sum_k A_i[m,k] * B_j[k,n]
// i.e. 36 u8 x u8 -> s32 GEMMs on limb planes (pairs with i+j >= 8 vanish mod 2^64).  A pair sum over K <= 33 025 stays
// below 2^31 (255^2 K), so it is exact in the s32 accumulator.  Diagonals d = i+j < 4 need more than 32 bits of the sum
// after the shift, so their 10 pairs get their own TMEM accumulators; for d >= 4 only the low 64-8d <= 32 bits matter and
// all pairs of a diagonal share one wrapping accumulator: 14 accumulators x 32 columns = 448 TMEM columns per 128x32 tile.
//
// Layout: operands are staged as byte PLANES in HBM ([8][rows][Kp] u8, Kp = K rounded up to 128, zero padded) by
// planarize kernels; the GEMM kernel streams them with 3D TMA (SWIZZLE_128B) -- the 8 B-side planes of a 128-deep k-chunk
// are resident (32 KB, double buffered), the A-side planes flow through a ring of 16 KB stages -- and issues
// 144 UTCIMMA (M=128, N=32, K=32) per k-chunk.  Epilogue: 14 accumulators -> shifts/adds in 64-bit registers -> int64.
// Split-K uses 64-bit integer atomics (exact, order independent).
#include "common.cuh"
#include <cuda.h>
#include <algorithm>

namespace ri8 {

typedef unsigned long long u64;
constexpr int NTHREADS = 192;
constexpr int A_TILE = 128 * 128;  // one plane: 128 rows x 128 k-bytes
constexpr int BN = 32;
constexpr int B_TILE = BN * 128;   // one plane: 32 rows x 128 k-bytes
constexpr int ASTAGES = 6;
constexpr int NACC = 14;

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
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
               "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_i8(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t v[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr));
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {  // K-major, SWIZZLE_128B, SBO = 1024
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// InstrDescriptor: c_format [4,6) = 2 (S32), a_format [7,10) = 0 (u8), b_format [10,13) = 0 (u8), K-major, N>>3 [17,23), M>>4 [24,29)
constexpr uint32_t IDESC = (2u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

// accumulator slot of limb pair (i,j)
__device__ __forceinline__ int acc_slot(int i, int j) {
  const int d = i + j;
  return d < 4 ? d * (d + 1) / 2 + i : 10 + (d - 4);
}

// C[rows, N] (+)= Cinit + A1 @ B1 + A2 @ B2 ;  planes: tmA1/tmA2 [8][R][Kp], tmB1/tmB2 [8][N][Kp]
__global__ void __launch_bounds__(NTHREADS, 1)
ring_gemm_i8_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB1,
                    const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2, int nseg, int R, int N, int Kp,
                    int splitK, const u64* __restrict__ Cinit, u64* __restrict__ Cout) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment as an OFFSET from the __shared__ array: the pointer keeps its address space, so the epilogue scratch
  // compiles to STS / LDS (a round trip through uintptr_t made nvcc emit generic ST.E / LD.E)
  uint8_t* smem = smem_raw + ((1024u - ((uint32_t)__cvta_generic_to_shared(smem_raw) & 1023u)) & 1023u);
  const uint32_t s_base = smem_u32(smem);
  const uint32_t b_base = s_base + ASTAGES * A_TILE;  // 2 x 8 x B_TILE
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ASTAGES * A_TILE + 2 * 8 * B_TILE);
  const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * ASTAGES, bfull0 = empty0 + 8 * ASTAGES, bempty0 = bfull0 + 16,
                 accum_bar = bempty0 + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * ASTAGES + 5);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int m0 = blockIdx.x * 128, n0 = blockIdx.y * BN, split = blockIdx.z;
  const int nk = Kp / 128, total = nk * nseg;
  const int per = (total + splitK - 1) / splitK;
  const int c_begin = split * per, c_end = min(total, c_begin + per);
  const int nchunks = max(0, c_end - c_begin);

  if (tid == 0) {
    for (int s = 0; s < ASTAGES; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bfull0 + 8 * b, 1);
      mbar_init(bempty0 + 8 * b, 1);
    }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (tid == 0) {
    // ------------------------------------------------------------ TMA producer
    int it = 0;
    for (int cc = 0; cc < nchunks; ++cc) {
      const int chunk = c_begin + cc;
      const int seg = chunk / nk, k0 = (chunk - seg * nk) * 128;
      const CUtensorMap* ta = seg == 0 ? &tmA1 : &tmA2;
      const CUtensorMap* tb = seg == 0 ? &tmB1 : &tmB2;
      const int bb = cc & 1;
      mbar_wait(bempty0 + 8 * bb, ((cc >> 1) & 1) ^ 1);
      mbar_expect_tx(bfull0 + 8 * bb, 8 * B_TILE);
      for (int j = 0; j < 8; ++j) tma_load_3d(b_base + (bb * 8 + j) * B_TILE, tb, bfull0 + 8 * bb, k0, n0, j);
      for (int i = 0; i < 8; ++i, ++it) {
        const int s = it % ASTAGES;
        mbar_wait(empty0 + 8 * s, ((it / ASTAGES) & 1) ^ 1);
        mbar_expect_tx(full0 + 8 * s, A_TILE);
        tma_load_3d(s_base + s * A_TILE, ta, full0 + 8 * s, k0, m0, i);
      }
    }
  } else if (tid == 32) {
    // ------------------------------------------------------------ MMA issuer
    int it = 0;
    uint32_t touched = 0;  // bit a: accumulator slot a already holds data
    for (int cc = 0; cc < nchunks; ++cc) {
      const int bb = cc & 1;
      mbar_wait(bfull0 + 8 * bb, (cc >> 1) & 1);
      for (int i = 0; i < 8; ++i, ++it) {
        const int s = it % ASTAGES;
        mbar_wait(full0 + 8 * s, (it / ASTAGES) & 1);
        tc_fence_after();
        const uint64_t adesc = make_desc(s_base + s * A_TILE);
        for (int j = 0; j < 8 - i; ++j) {
          const int slot = acc_slot(i, j);
          const uint64_t bdesc = make_desc(b_base + (bb * 8 + j) * B_TILE);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            umma_i8(tmem + slot * BN, adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), IDESC, (touched >> slot) & 1u);
            touched |= 1u << slot;
          }
        }
        umma_commit(empty0 + 8 * s);
      }
      umma_commit(bempty0 + 8 * bb);
    }
    if (nchunks > 0) umma_commit(accum_bar);
  } else if (warp >= 2 && nchunks > 0) {
    // ------------------------------------------------------------ epilogue: limb recombination
    const int quad = warp & 3;
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    const int row = m0 + quad * 32 + (tid & 31);
    const uint32_t lane_addr = tmem + ((uint32_t)(quad * 32) << 16);
#pragma unroll 1
    for (int c8 = 0; c8 < BN / 8; ++c8) {
      u64 z[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) z[e] = 0;
#pragma unroll
      for (int slot = 0; slot < NACC; ++slot) {
        uint32_t v[8];
        tmem_ld8(lane_addr + slot * BN + c8 * 8, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        // shift of the slot: individual pairs d = 0,1,1,2,2,2,3,3,3,3 ; shared diagonals d = 4..7
        const int d = slot < 1 ? 0 : slot < 3 ? 1 : slot < 6 ? 2 : slot < 10 ? 3 : slot - 6;
#pragma unroll
        for (int e = 0; e < 8; ++e) z[e] += (u64)v[e] << (8 * d);
      }
      if (row < R) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const size_t o = (size_t)row * N + n0 + c8 * 8 + e;
          u64 val = z[e];
          if (split == 0 && Cinit) val += Cinit[o];
          if (splitK > 1) atomicAdd(&Cout[o], val);
          else Cout[o] = val;
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------- limb planes
// x [R][K] int64 (row-major) -> planes [8][R][Kp] u8.  One thread = one row x 16 consecutive k (one 16-byte store per plane).
// peer != NULL: the OPENING is fused -- the planes of (x + peer) mod 2^64, peer being the other party's masked share (possibly a
// peer-mapped pointer into another GPU); the opened int64 operand is never materialised.
__global__ void planarize_rows_kernel(const u64* __restrict__ x, const u64* __restrict__ peer, int R, int K, int Kp,
                                      uint8_t* __restrict__ out) {
  const int kv = Kp / 16;
  const size_t total = (size_t)R * kv;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int kc = (int)(i % kv);
    const size_t r = i / kv;
    __align__(16) uint8_t b[8][16];
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      const int k = kc * 16 + e;
      u64 v = k < K ? x[r * K + k] : 0ull;
      if (peer != nullptr && k < K) v += peer[r * K + k];
#pragma unroll
      for (int l = 0; l < 8; ++l) b[l][e] = (uint8_t)(v >> (8 * l));
    }
#pragma unroll
    for (int l = 0; l < 8; ++l)
      *reinterpret_cast<uint4*>(out + ((size_t)l * R + r) * Kp + kc * 16) = *reinterpret_cast<const uint4*>(b[l]);
  }
}

// x [K][N] int64 (row-major, the reference's right operand) -> planes of its transpose [8][N][Kp] u8 (32x32 smem transpose)
__global__ void planarize_cols_kernel(const u64* __restrict__ x, int K, int N, int Kp, uint8_t* __restrict__ out) {
  __shared__ u64 tile[32][33];
  const int k0 = blockIdx.x * 32, n0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int k = k0 + r, n = n0 + threadIdx.x;
    tile[r][threadIdx.x] = (k < K && n < N) ? x[(size_t)k * N + n] : 0ull;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int n = n0 + r, k = k0 + threadIdx.x;
    if (n < N && k < Kp) {
      const u64 v = tile[threadIdx.x][r];
#pragma unroll
      for (int l = 0; l < 8; ++l) out[((size_t)l * N + n) * Kp + k] = (uint8_t)(v >> (8 * l));
    }
  }
}

// ---------------------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_tiled = nullptr;
static bool load_driver() {
  if (g_tiled) return true;
  cudaDriverEntryPointQueryResult q;
  void* f = nullptr;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || !f) return false;
  g_tiled = (EncodeTiledFn)f;
  return true;
}
static bool map_planes(CUtensorMap* tm, const void* base, uint64_t rows, uint64_t Kp, uint32_t box_rows) {
  cuuint64_t dims[3] = {Kp, rows, 8};
  cuuint64_t strides[2] = {Kp, rows * Kp};
  cuuint32_t box[3] = {128, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  return g_tiled(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
constexpr int SMEM_BYTES = ASTAGES * A_TILE + 2 * 8 * B_TILE + 1024 + 256;

}  // namespace ri8

extern "C" {

size_t pm_ring_tc_ws_bytes(int rows, int K, int N, int nseg) {
  const size_t Kp = ((size_t)K + 127) / 128 * 128;
  return (size_t)nseg * 8 * Kp * ((size_t)rows + (size_t)N);
}

int pm_ring_tc_supported(int rows, int K, int N) { return N % 32 == 0 && N >= 32 && K >= 1 && rows >= 1 && 2L * K <= 33000; }

// ---- the limb-plane GEMM in pieces, so that operands which do not depend on the image (the triple's a and b, the opened
// weight mask eps) are planarised ONCE in the offline phase and the online phase runs: open+planarise(delta) -> GEMM.
size_t pm_ring_planes_bytes(int n, int K) { return (size_t)8 * (size_t)n * (((size_t)K + 127) / 128 * 128); }

// planes [8][rows][Kp] of A (+ peer): left operand.  peer may be NULL.
int pm_ring_planarize_rows_i64(const int64_t* A, const int64_t* peer, int rows, int K, void* planes, pm_stream_t s) {
  using namespace ri8;
  PM_CHECK_ARG(A && planes && rows >= 1 && K >= 1);
  const int Kp = (K + 127) / 128 * 128;
  const size_t tot = (size_t)rows * (Kp / 16);
  planarize_rows_kernel<<<pm_grid(tot, 128, 1, 16), 128, 0, S(s)>>>((const u64*)A, (const u64*)peer, rows, K, Kp, (uint8_t*)planes);
  PM_LAUNCH_OK();
}

// planes [8][N][Kp] of B^T, B [K,N]: right operand.
int pm_ring_planarize_cols_i64(const int64_t* B, int K, int N, void* planes, pm_stream_t s) {
  using namespace ri8;
  PM_CHECK_ARG(B && planes && N >= 1 && K >= 1);
  const int Kp = (K + 127) / 128 * 128;
  dim3 gb((Kp + 31) / 32, (N + 31) / 32), bb(32, 8);
  planarize_cols_kernel<<<gb, bb, 0, S(s)>>>((const u64*)B, K, N, Kp, (uint8_t*)planes);
  PM_LAUNCH_OK();
}

// C[rows,N] = Cinit + A1 @ B1 (+ A2 @ B2) from limb planes (pm_ring_planarize_*).  pa2/pb2 may be NULL.
int pm_ring_gemm_planes_i64(const void* pa1, const void* pb1, const void* pa2, const void* pb2, const int64_t* Cinit, int rows, int K,
                            int N, int64_t* C, pm_stream_t s) {
  using namespace ri8;
  PM_CHECK_ARG(pa1 && pb1 && C && pm_ring_tc_supported(rows, K, N) && ((pa2 == nullptr) == (pb2 == nullptr)));
  if (!load_driver()) return pm_set_err(__FILE__, __LINE__, "cuTensorMapEncodeTiled unavailable");
  const int nseg = pa2 ? 2 : 1;
  const int Kp = (K + 127) / 128 * 128;
  const void* pa[2] = {pa1, pa2};
  const void* pb[2] = {pb1, pb2};
  CUtensorMap tA[2], tB[2];
  for (int g = 0; g < 2; ++g) {
    const int h = g < nseg ? g : 0;
    if (!map_planes(&tA[g], pa[h], rows, Kp, 128) || !map_planes(&tB[g], pb[h], N, Kp, BN))
      return pm_set_err(__FILE__, __LINE__, "tensor map encode failed");
  }
  const int tm = (rows + 127) / 128, tn = N / BN;
  const int total = (Kp / 128) * nseg;
  int splitK = 1;
  const long tiles = (long)tm * tn;
  if (tiles < pm_num_sms()) splitK = (int)std::min<long>(total, (pm_num_sms() + tiles - 1) / tiles);
  if (splitK > 1) PM_CUDA(cudaMemsetAsync(C, 0, (size_t)rows * N * sizeof(int64_t), S(s)));
  PM_CUDA(cudaFuncSetAttribute(ring_gemm_i8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  dim3 grid(tm, tn, splitK);
  ring_gemm_i8_kernel<<<grid, NTHREADS, SMEM_BYTES, S(s)>>>(tA[0], tB[0], tA[1], tB[1], nseg, rows, N, Kp, splitK, (const u64*)Cinit, (u64*)C);
  PM_LAUNCH_OK();
}

// C[rows,N] = Cinit + A1[rows,K] @ B1[K,N] (+ A2 @ B2) over Z_2^64 on the int8 tensor cores.  A2/B2 may be NULL.
// ws: pm_ring_tc_ws_bytes(rows, K, N, nseg) bytes of scratch for the limb planes.
int pm_ring_gemm2_tc_i64(const int64_t* A1, const int64_t* B1, const int64_t* A2, const int64_t* B2, const int64_t* Cinit, int rows,
                         int K, int N, void* ws, int64_t* C, pm_stream_t s) {
  PM_CHECK_ARG(A1 && B1 && C && ws && pm_ring_tc_supported(rows, K, N) && ((A2 == nullptr) == (B2 == nullptr)));
  const int nseg = A2 ? 2 : 1;
  uint8_t* w = (uint8_t*)ws;
  uint8_t* pa[2] = {nullptr, nullptr};
  uint8_t* pb[2] = {nullptr, nullptr};
  const int64_t* As[2] = {A1, A2};
  const int64_t* Bs[2] = {B1, B2};
  for (int g = 0; g < nseg; ++g) {
    pa[g] = w; w += pm_ring_planes_bytes(rows, K);
    pb[g] = w; w += pm_ring_planes_bytes(N, K);
    if (int rc = pm_ring_planarize_rows_i64(As[g], nullptr, rows, K, pa[g], s)) return rc;
    if (int rc = pm_ring_planarize_cols_i64(Bs[g], K, N, pb[g], s)) return rc;
  }
  return pm_ring_gemm_planes_i64(pa[0], pb[0], pa[1], pb[1], Cinit, rows, K, N, C, s);
}

}  // extern "C"
