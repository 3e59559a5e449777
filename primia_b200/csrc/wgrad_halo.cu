// Path T, throughput mode: weight gradient of the 3x3 / stride-1 / pad-1 convolutions as a HALO-STRIP implicit GEMM.
// Default path for the 128-multiple channel counts since round 2 (PRIMIA_HALO_WGRAD=0 falls back to wgrad_tma_kernel):
// weight-gradient family 0.565 -> 0.519 ms per step at B = 64 (tests/test_conv_tc_gpu.py covers it).
//
//   dw[n][r][s][c] = sum over pixels  dy[pix][n] * x[pix + (r-1, s-1)][c]
//
// wgrad_tma_kernel (conv_tma.cu) gives every (tap, 128-channel) tile its own CTA, which re-fetches an im2col copy of x and the
// dy tile for each 128-pixel step: 64 KB per 8 MMAs = 128 B/clk/SM against an L2->SM cap of ~54 B/clk/SM.  Here a CTA owns
// one filter ROW r (three taps s = 0..2) x 128 input channels x 128 output channels: three [128 c][128 n] fp32 accumulators
// (384 TMEM columns), and per 128-pixel step it loads ONE strip of x and ONE strip of dy, both on the virtual pixel grid of
// conv_halo.cu (rows padded to Wp = W+1, images to Hp = H+1 by TMA out-of-bounds zero fill), so that
//   * tap (r, s) of x is the row shift r*Wp + s of an MN-major SWIZZLE_128B descriptor into the x strip,
//   * virtual pixels (w == W or h == H) have dy == 0 and contribute nothing.
// 24 MMAs (M=128 channels, N=128, K=16 pixels) per ~89 KB loaded: ~58 B/clk/SM.
//
//   warp 0: x-strip producer (one TMA box {64 c, Wp, 1, 1} per padded row and 64-channel half)
//   warp 2: dy-strip producer (same boxes over dy)
//   warp 1: MMA issuer, warps 4-7: final epilogue (fp32 atomics into dw, KRSC)
#include "common.cuh"
#include "tc_ptx.cuh"
#include <algorithm>
#include <stdlib.h>

namespace wgh {
using namespace tcx;
typedef __nv_bfloat16 bf16;

struct WGeo {
  int B, H, W, C, K;      // x [B,H,W,C], dy [B,H,W,K]
  int Wp, Hp, V;          // padded row / image pitch, virtual pixels
  int steps_per_split;    // 128-pixel steps per blockIdx.z
  int half_bytes;         // one 64-wide half of a strip stage (max rows * Wp * 128, rounded to 1024)
};

constexpr int STAGES = 2;
constexpr int NTHR = 256;

__global__ void __launch_bounds__(NTHR, 1)
wgrad_halo_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmD, WGeo g, float* __restrict__ dw) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment as an OFFSET from the __shared__ array: the pointer keeps its address space, so the epilogue scratch
  // compiles to STS / LDS (a round trip through uintptr_t made nvcc emit generic ST.E / LD.E)
  uint8_t* smem = smem_raw + ((1024u - ((uint32_t)__cvta_generic_to_shared(smem_raw) & 1023u)) & 1023u);
  // stage = [x half 0][x half 1][dy half 0][dy half 1], each half_bytes
  const uint32_t s_base = smem_u32(smem);
  const uint32_t stage_bytes = 4u * (uint32_t)g.half_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * stage_bytes);
  const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * STAGES, accum_bar = empty0 + 8 * STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cgroups = g.C / 128;
  const int r = blockIdx.x / cgroups;                 // filter row of this CTA
  const int c0 = (blockIdx.x - r * cgroups) * 128;    // its 128 input channels
  const int n0 = blockIdx.y * 128;                    // its 128 output channels
  const int nsteps_total = (g.V + 127) / 128;
  const int st_begin = blockIdx.z * g.steps_per_split;
  const int st_end = min(nsteps_total, st_begin + g.steps_per_split);
  const int nsteps = st_end > st_begin ? st_end - st_begin : 0;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full0 + 8 * s, 2);   // the x producer and the dy producer each arrive (with their byte counts)
      mbar_init(empty0 + 8 * s, 1);
    }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmD) : "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t row_bytes = (uint32_t)g.Wp * 128u;

  if (warp == 0 || warp == 2) {
    // ------------------------------------------------------------ strip producers (warp 0: x, warp 2: dy)
    const bool is_x = warp == 0;
    const CUtensorMap* tm = is_x ? &tmX : &tmD;
    const int ch0 = is_x ? c0 : n0;
    uint32_t ss = 0, ph = 1;
    for (int st = 0; st < nsteps; ++st) {
      const int q0 = (st_begin + st) * 128;
      // x: padded positions q0 + r*Wp .. q0 + r*Wp + 127 + 2 ; dy: virtual pixels q0 .. q0 + 127
      const int first = is_x ? q0 + r * g.Wp : q0;
      const int last = is_x ? first + 127 + 2 : first + 127;
      const int R0 = first / g.Wp;
      const int nrows = last / g.Wp - R0 + 1;
      mbar_wait(empty0 + 8 * ss, ph);
      if (elect_one()) {
        const uint32_t full = full0 + 8 * ss;
        mbar_expect_tx(full, 2u * (uint32_t)nrows * row_bytes);
        const uint32_t dst0 = s_base + ss * stage_bytes + (is_x ? 0u : 2u * (uint32_t)g.half_bytes);
        int img = R0 / g.Hp, hh = R0 - img * g.Hp;
        uint32_t off = 0;
        for (int i = 0; i < nrows; ++i, off += row_bytes) {
          // x: padded row hh <-> image row hh-1, padded column 0 <-> image column -1 (zero fill on both borders)
          // dy: virtual row hh <-> image row hh (hh == H is the zero row), virtual column w <-> image column w (w == W zero)
          const int cw = is_x ? -1 : 0, chh = is_x ? hh - 1 : hh;
          tma_load_4d(dst0 + off, tm, full, ch0, cw, chh, img);
          tma_load_4d(dst0 + (uint32_t)g.half_bytes + off, tm, full, ch0 + 64, cw, chh, img);
          if (++hh == g.Hp) { hh = 0; ++img; }
        }
      }
      __syncwarp();
      if (++ss == STAGES) { ss = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = make_idesc(128, 128, 1, 1);  // both operands MN-major
    constexpr uint32_t HI = (1024u >> 4) | (1u << 14) | (2u << 29);
    const uint32_t LO = ((uint32_t)g.half_bytes >> 4) << 16;  // LBO = distance between the two 64-wide halves
    uint32_t ss = 0, ph = 0;
    for (int st = 0; st < nsteps; ++st) {
      const int q0 = (st_begin + st) * 128;
      const int xfirst = q0 + r * g.Wp;
      const uint32_t xoff = (uint32_t)(xfirst - (xfirst / g.Wp) * g.Wp);  // row offset of the step's first pixel in each strip
      const uint32_t doff = (uint32_t)(q0 - (q0 / g.Wp) * g.Wp);
      mbar_wait(full0 + 8 * ss, ph);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t xs = LO + ((s_base + ss * stage_bytes) >> 4) + xoff * 8u;
        const uint32_t ds = LO + ((s_base + ss * stage_bytes + 2u * (uint32_t)g.half_bytes) >> 4) + doff * 8u;
#pragma unroll
        for (int s = 0; s < 3; ++s)
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)  // 16 pixels (rows of 128 B) per K step
            umma_bf16(tmem_base + s * 128, desc_pack(xs + (uint32_t)(s * 8 + ks * 128), HI), desc_pack(ds + (uint32_t)(ks * 128), HI), idesc,
                      (st | ks) != 0);
        umma_commit(empty0 + 8 * ss);
        if (st == nsteps - 1) umma_commit(accum_bar);
      }
      __syncwarp();
      if (++ss == STAGES) { ss = 0; ph ^= 1; }
    }
  } else if (warp >= 4 && nsteps > 0) {
    // ------------------------------------------------------------ epilogue: dw[n][r][s][c] += acc_s[c][n]
    const int quad = warp & 3;
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    const int c = c0 + quad * 32 + lane;
    const size_t Kg = (size_t)9 * g.C;
#pragma unroll 1
    for (int s = 0; s < 3; ++s) {
      const size_t kg = (size_t)(r * 3 + s) * g.C + c;
#pragma unroll 1
      for (int cc = 0; cc < 4; ++cc) {
        uint32_t v[32];
        tmem_ld32_nowait(tmem_base + s * 128 + cc * 32 + ((uint32_t)(quad * 32) << 16), v);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e) atomicAdd(dw + (size_t)(n0 + cc * 32 + e) * Kg + kg, __uint_as_float(v[e]));
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

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
// NHWC bf16 [B][H][W][C]: one padded row of W+1 pixels x 64 channels per load
static bool map_rows(CUtensorMap* tm, const void* base, int B, int H, int W, int C) {
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)(W + 1), 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  return g_tiled(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace wgh

// 0 = launched, 1 = not eligible / not enabled (caller uses wgrad_tma_kernel), 2 = CUDA / driver error.  dw accumulates.
int pm_halo_conv_wgrad(const pm_conv_t* p, const void* x, const void* dy, float* dw, cudaStream_t st) {
  using namespace wgh;
  const char* e = getenv("PRIMIA_HALO_WGRAD");   // default on since round 2 (1.94 vs 1.97 ms per step); "0" switches it off
  if (e && e[0] == '0') return 1;
  if (p->R != 3 || p->S != 3 || p->stride != 1 || p->pad != 1 || p->Ho != p->H || p->Wo != p->W) return 1;
  if (p->C % 128 != 0 || p->K % 128 != 0 || p->W + 1 > 256 || !load_driver()) return 1;
  WGeo g;
  g.B = p->B; g.H = p->H; g.W = p->W; g.C = p->C; g.K = p->K;
  g.Wp = p->W + 1; g.Hp = p->H + 1;
  const long V = (long)g.B * g.Hp * g.Wp;
  if (V > 0x7fffffffL - 1024) return 1;
  g.V = (int)V;
  // rows of one strip: the x strip spans 128 + 2 positions starting anywhere in a row
  const int max_rows = (g.Wp - 1 + 129) / g.Wp + 1;
  g.half_bytes = (max_rows * g.Wp * 128 + 1023) / 1024 * 1024;
  const int smem = STAGES * 4 * g.half_bytes + (2 * STAGES + 2) * 8 + 1024;
  if (smem > 227 * 1024) return 1;
  const int tiles = 3 * (g.C / 128) * (g.K / 128);
  const int nsteps = (g.V + 127) / 128;
  int splits = std::max(1, pm_num_sms() / tiles);
  if (splits > nsteps) splits = nsteps;
  g.steps_per_split = (nsteps + splits - 1) / splits;
  splits = (nsteps + g.steps_per_split - 1) / g.steps_per_split;
  CUtensorMap tmX, tmD;
  if (!map_rows(&tmX, x, g.B, g.H, g.W, g.C) || !map_rows(&tmD, dy, g.B, g.H, g.W, g.K)) return 2;
  if (cudaFuncSetAttribute(wgrad_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return 2;
  dim3 grid(3 * (g.C / 128), g.K / 128, splits);
  wgrad_halo_kernel<<<grid, NTHR, smem, st>>>(tmX, tmD, g, dw);
  return 0;
}
