// Path T, throughput mode: data gradient of the stride-2 convolutions (3x3 / pad 1 and 1x1 / pad 0), persistent version.
//
// Same decomposition as dgrad_s2_tma_kernel (conv_tma.cu): dx[2i+a, 2j+b] only receives the taps with r = (a+pad) mod 2,
// s = (b+pad) mod 2, so each output parity class (a, b) is a plain stride-1 correlation of dy over 1, 2, 2 or 4 taps -- an
// im2col-mode TMA load per (tap, 64-channel block).  The one-tile-per-CTA kernel spent most of its time in prologues (barrier
// init, TMEM allocation, descriptor fetch: 1568 CTAs of 2-8 k-blocks each for the 56x56 layer) and in 16-byte-per-line
// strided stores.  Here one CTA per SM walks (class, m-tile, n-tile) work items with the warp-converged elect.sync issue,
// double-buffered TMEM accumulators and the shared-memory staged epilogue of conv_halo.cu.
#include "common.cuh"
#include "tc_ptx.cuh"
#include <algorithm>
#include <stdlib.h>

namespace s2p {
using namespace tcx;
typedef __nv_bfloat16 bf16;

struct Geo {
  int B, H, W, C, K, R, S, pad, Ho, Wo;
};

__device__ __forceinline__ void tma_load_im2col(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c, int w, int h, int n,
                                                uint16_t off_w, uint16_t off_h) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};" ::
          "r"(dst),
      "l"(tm), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h)
      : "memory");
}

constexpr int NTHR = 256;   // warp 0 producer, warp 1 MMA, warps 4-7 epilogue
constexpr int TILE_BYTES = 128 * 128;

template <int BN, int STAGES>
__global__ void __launch_bounds__(NTHR, 1)
dgrad_s2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, Geo p, bf16* __restrict__ dst,
                int accumulate) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment as an OFFSET from the __shared__ array: the pointer keeps its address space, so the epilogue scratch
  // compiles to STS / LDS (a round trip through uintptr_t made nvcc emit generic ST.E / LD.E)
  uint8_t* smem = smem_raw + ((1024u - ((uint32_t)__cvta_generic_to_shared(smem_raw) & 1023u)) & 1023u);
  constexpr int A_BYTES = TILE_BYTES, B_BYTES = BN * 128, STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int TMEM_COLS = 2 * BN;
  const uint32_t s_base = smem_u32(smem);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * STAGES, tfull0 = empty0 + 8 * STAGES, tempty0 = tfull0 + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  uint8_t* epi_scr_all = reinterpret_cast<uint8_t*>(bars + 2 * STAGES + 6);  // [4 warps][2048]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int H2 = p.H / 2, W2 = p.W / 2;  // class-pixel grid (== Ho x Wo)
  const int Mc = p.B * H2 * W2;
  const int N = p.C;
  const int cblocks = p.K / 64;
  const int ntm = (Mc + 127) / 128, ntn = N / BN;
  const int ntiles = 4 * ntm * ntn;  // work item = (m-tile, n-tile, class); class fastest: the four classes share dy tiles

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull0 + 8 * a, 1);
      mbar_init(tempty0 + 8 * a, 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pm_pdl_sync();  // CTA-local prologue above, first global access below

  // per work item: class (ca, cb), taps r = r_first + 2*ri (nr of them), s = s_first + 2*si (ns of them)
#define S2P_DECODE(tile)                                                                                     \
  const int cls = (tile) & 3, rest = (tile) >> 2;                                                           \
  const int ca = cls >> 1, cb = cls & 1;                                                                    \
  const int m0 = (rest / ntn) * 128, n0 = (rest % ntn) * BN;                                                \
  const int r_first = (ca + p.pad) & 1, s_first = (cb + p.pad) & 1;                                         \
  const int nr = r_first < p.R ? (p.R - r_first + 1) / 2 : 0, ns = s_first < p.S ? (p.S - s_first + 1) / 2 : 0; \
  const int nkb = nr * ns * cblocks;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    uint32_t ss = 0, ph = 1;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      S2P_DECODE(tile)
      (void)n0;
      const int nb = m0 / (H2 * W2);
      const int rem = m0 - nb * H2 * W2;
      const int pi = rem / W2, pj = rem - pi * W2;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(empty0 + 8 * ss, ph);
        if (elect_one()) {
          const uint32_t a_tile = s_base + ss * STAGE_BYTES, b_tile = a_tile + A_BYTES;
          const int ti = kb / cblocks, k0 = (kb - ti * cblocks) * 64;
          const int ri = ti / ns, si = ti - ri * ns;
          const int r = r_first + 2 * ri, sx = s_first + 2 * si;
          const int off_h = (ca + p.pad - r) / 2, off_w = (cb + p.pad - sx) / 2;  // in {0, 1} for 3x3/p1 and 1x1/p0
          mbar_expect_tx(full0 + 8 * ss, STAGE_BYTES);
          tma_load_im2col(a_tile, &tmA, full0 + 8 * ss, k0, pj, pi, nb, (uint16_t)off_w, (uint16_t)off_h);
          tma_load_2d(b_tile, &tmB, full0 + 8 * ss, (r * p.S + sx) * p.K + k0, n0);
        }
        __syncwarp();
        if (++ss == STAGES) { ss = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = make_idesc(128, BN, 0, 0);
    uint32_t ss = 0, ph = 0;
    int lt = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      S2P_DECODE(tile)
      (void)m0; (void)n0;
      if (nkb == 0) continue;
      const int as = lt & 1;
      mbar_wait(tempty0 + 8 * as, ((lt >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + as * BN;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(full0 + 8 * ss, ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_lo = DESC_SW128_LO + ((s_base + ss * STAGE_BYTES) >> 4), b_lo = a_lo + (A_BYTES >> 4);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_d, desc_pack(a_lo + k * 2, DESC_SW128_HI), desc_pack(b_lo + k * 2, DESC_SW128_HI), idesc, (kb | k) != 0);
          umma_commit(empty0 + 8 * ss);
          if (kb == nkb - 1) umma_commit(tfull0 + 8 * as);
        }
        __syncwarp();
        if (++ss == STAGES) { ss = 0; ph ^= 1; }
      }
      ++lt;
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue
    const int quad = warp & 3;
    uint8_t* epi_scr = epi_scr_all + quad * 2048;
    float st[4] = {0.f, 0.f, 0.f, 0.f};
    int lt = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      S2P_DECODE(tile)
      const int m = m0 + quad * 32 + lane;
      const bool row_ok = m < Mc;
      const int mm = row_ok ? m : 0;
      const int nb = mm / (H2 * W2);
      const int rem = mm - nb * H2 * W2;
      const int pi = rem / W2, pj = rem - pi * W2;
      bf16* out = dst + (((size_t)nb * p.H + 2 * pi + ca) * p.W + 2 * pj + cb) * N + n0;
      if (nkb == 0) {
        // this class receives no tap (1x1 / stride 2): its gradient is exactly zero
        if (!accumulate && row_ok)
          for (int q = 0; q < BN / 8; ++q) reinterpret_cast<uint4*>(out)[q] = make_uint4(0, 0, 0, 0);
        continue;
      }
      const int as = lt & 1;
      mbar_wait(tfull0 + 8 * as, (lt >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int cc = 0; cc < BN / 32; ++cc) {
        uint32_t v[32];
        tmem_ld32_nowait(tmem_base + as * BN + cc * 32 + ((uint32_t)(quad * 32) << 16), v);
        tmem_ld_wait();
        epilogue_chunk32(v, row_ok, out + cc * 32, accumulate != 0, false, epi_scr, lane, st);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty0 + 8 * as);
      ++lt;
    }
  }
#undef S2P_DECODE
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------ host
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
static bool map_dense(CUtensorMap* tm, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  return g_tiled(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// dy [B][Ho][Wo][K] as an im2col source with window offsets in {0, 1}: 128 pixels x 64 channels per load
static bool map_im2col(CUtensorMap* tm, const void* base, int B, int H, int W, int C) {
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  int lower[2] = {0, 0};
  int upper[2] = {0, 0};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  if (g_im2col(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, lower, upper, 64, 128, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return false;
  // same driver workaround as conv_tma.cu (im2col descriptors of tensors smaller than 128 KiB, driver <= 13.1)
  int drv = 0;
  cudaDriverGetVersion(&drv);
  if (drv <= 13010 && (size_t)B * H * W * C * 2 < 131072) reinterpret_cast<uint64_t*>(tm)[1] &= ~(1ull << 21);
  return true;
}
constexpr int STAGES = 5;
constexpr int smem_bytes(int BN) { return STAGES * (TILE_BYTES + BN * 128) + (2 * STAGES + 6) * 8 + 4 * 2048 + 1024; }

}  // namespace s2p

// 0 = launched, 1 = not eligible (caller uses dgrad_s2_tma_kernel), 2 = CUDA / driver error
int pm_s2p_conv_dgrad(const pm_conv_t* p, const void* dy, const void* wt, void* dx, int accumulate, cudaStream_t st) {
  using namespace s2p;
  const char* e = getenv("PRIMIA_NO_S2P");
  if (e && e[0] == '1') return 1;
  if (p->stride != 2 || p->C % 64 != 0 || p->K % 64 != 0) return 1;
  const bool geom_ok = p->H % 2 == 0 && p->W % 2 == 0 && p->Ho == p->H / 2 && p->Wo == p->W / 2 &&
                       ((p->R == 3 && p->S == 3 && p->pad == 1) || (p->R == 1 && p->S == 1 && p->pad == 0));
  if (!geom_ok || !load_driver()) return 1;
  CUtensorMap tmA, tmB;
  const int Ktot = p->R * p->S * p->K;
  if (!map_im2col(&tmA, dy, p->B, p->Ho, p->Wo, p->K)) return 2;
  Geo g{p->B, p->H, p->W, p->C, p->K, p->R, p->S, p->pad, p->Ho, p->Wo};
  const int Mc = p->B * p->Ho * p->Wo;
  if (p->C % 128 == 0) {
    if (!map_dense(&tmB, wt, p->C, Ktot, 128)) return 2;
    if (cudaFuncSetAttribute(dgrad_s2_kernel<128, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes(128)) != cudaSuccess) return 2;
    const int ntiles = 4 * ((Mc + 127) / 128) * (p->C / 128);
    if (pm_launch(dgrad_s2_kernel<128, STAGES>, dim3(std::min(pm_num_sms(), ntiles)), dim3(NTHR), (size_t)smem_bytes(128), st, tmA, tmB, g,
                  (bf16*)dx, accumulate) != cudaSuccess) return 2;
  } else {
    if (!map_dense(&tmB, wt, p->C, Ktot, 64)) return 2;
    if (cudaFuncSetAttribute(dgrad_s2_kernel<64, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes(64)) != cudaSuccess) return 2;
    const int ntiles = 4 * ((Mc + 127) / 128) * (p->C / 64);
    if (pm_launch(dgrad_s2_kernel<64, STAGES>, dim3(std::min(pm_num_sms(), ntiles)), dim3(NTHR), (size_t)smem_bytes(64), st, tmA, tmB, g,
                  (bf16*)dx, accumulate) != cudaSuccess) return 2;
  }
  return 0;
}
