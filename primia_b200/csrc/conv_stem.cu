// Path T, throughput mode: the ResNet stem (conv 7x7 / stride 2 / pad 3, 3 -> 64 channels; models.py:379, 468) as a direct
// tcgen05 implicit GEMM over the fp32 NCHW input -- forward and weight gradient -- with NO materialised im2col matrix.
//
// Why: the im2col + dense-GEMM route writes and re-reads a [B*Ho*Wo, 192] bf16 matrix (308 MB at B = 64, three passes per
// step).  Here producer warps build each [128 output pixels x 192 k] operand tile directly in shared memory, in the
// SWIZZLE_128B image the tensor core expects, from the fp32 input patch (3 channels x 7 rows x 264 pixels) that TMA drops into
// shared memory (tiled boxes straight over the NCHW tensor, out-of-image pixels zero-filled by the hardware).
//
// k order: k = (r*3 + c)*8 + s with s in 0..7 (s = 7 is a zero weight), 21 groups of 8 = 168, padded to 192 = 3 blocks of 64.
// A 16-byte operand chunk (8 consecutive k) is then 8 CONSECUTIVE input pixels of one (channel, row): for output pixel m the
// chunk of group (r, c) is x[c][2*oh-3+r][2*m-3 .. 2*m+4] -- four aligned 64-bit shared-memory loads, four packed fp32->bf16
// conversions and one 128-bit store.
//
// Tile = one output row segment of 128 pixels (b, oh, ow0).  Persistent CTA, 640 threads:
//   warp 0      : MMA issuer (11 k-steps of M=128, N=64, K=16 per tile; accumulators double-buffered in TMEM)
//   warp 1      : loads the [64][192] weight matrix once (TMA)            [wgrad: streams the dy tiles]
//   warp 2      : input patches: two TMA boxes {136 px, 7 rows, 3 ch} per tile (pixels 0..135 and 128..263), double-buffered
//   warps 4-11  : producers: fp32 patch -> swizzled bf16 operand tile (double-buffered)
//   warps 12-19 : epilogue, one [32 rows x 32 columns] chunk per warp: tcgen05.ld -> bf16 NHWC store + BatchNorm batch
//                 statistics   [wgrad: final fp32 atomics]
// The weight-gradient variant reuses the same operand image MN-major (M = k, K = pixels): D[192(+64 pad)][64] accumulates
// over all of the CTA's tiles in two TMEM accumulators and is added to dw[64][7][7][3] with fp32 atomics at the end.
#include "common.cuh"
#include "tc_ptx.cuh"
#include <algorithm>

namespace stem {
using namespace tcx;
typedef __nv_bfloat16 bf16;

constexpr int NTHR = 640;
constexpr int PBW = 136;                // one patch box: 136 input pixels (pixels 2m .. 2m+7 of 64 output pixels, + slack)
constexpr int PROWS = 21;               // (c, r) rows
constexpr int PBOX_TX = PROWS * PBW * 4;             // bytes one box delivers (11424)
constexpr int PBOX_BYTES = 11520;                    // box stride in shared memory (TMA destinations are 128-byte aligned)
constexpr int PATCH_BYTES = 2 * PBOX_BYTES;          // both boxes of one tile
constexpr int BLK = 16384;              // one 64-wide k block of the operand tile: 128 rows x 128 B

struct SGeo { int B, H, W, Ho, Wo, tiles_per_row, ntiles; };

// the (b, oh, ow0) of a tile
__device__ __forceinline__ void tile_coords(const SGeo& g, int tile, int& b, int& oh, int& ow0) {
  const int row = tile / g.tiles_per_row;
  ow0 = (tile - row * g.tiles_per_row) * 128;
  b = row / g.Ho;
  oh = row - b * g.Ho;
}

// producer: operand tile from the fp32 patch.  Thread -> pixel m = pt & 127, chunk groups kc = pt >> 7, +2, ... < 21.
// Pixels m < 64 read box 0, m >= 64 box 1 (which starts 128 input pixels further right).
template <bool WGRAD>
__device__ __forceinline__ void build_tile(const uint8_t* patch, uint8_t* a_tile, int pt, int valid_pixels) {
  const int m = pt & 127;
  const bool zero = WGRAD && m >= valid_pixels;  // wgrad: pixels past the row end must not meet the next row's dy
  const uint32_t row_off = (uint32_t)m * 128u, sw = (uint32_t)(m & 7);
  // the box starts at input pixel 2*ow0 - 4 (TMA needs a 16-byte aligned start: coordinate % 4 == 0 for fp32), so tap s of
  // pixel m is patch pixel 2m + 1 + s: one 32-bit, three 64-bit and one 32-bit load
  const uint8_t* base = patch + (m >> 6) * PBOX_BYTES + (m & 63) * 8 + 4;
#pragma unroll
  for (int kc = pt >> 7; kc < PROWS; kc += 2) {
    const int r = kc / 3, c = kc - r * 3;
    const uint8_t* src = base + (c * 7 + r) * (PBW * 4);
    const float e0 = *reinterpret_cast<const float*>(src);
    const float2 f1 = *reinterpret_cast<const float2*>(src + 4), f2 = *reinterpret_cast<const float2*>(src + 12),
                 f3 = *reinterpret_cast<const float2*>(src + 20);
    const float e7 = *reinterpret_cast<const float*>(src + 28);
    uint4 q;
    __nv_bfloat162* qh = reinterpret_cast<__nv_bfloat162*>(&q);
    qh[0] = __floats2bfloat162_rn(e0, f1.x); qh[1] = __floats2bfloat162_rn(f1.y, f2.x);
    qh[2] = __floats2bfloat162_rn(f2.y, f3.x); qh[3] = __floats2bfloat162_rn(f3.y, e7);
    if (zero) q = make_uint4(0, 0, 0, 0);
    *reinterpret_cast<uint4*>(a_tile + (kc >> 3) * BLK + row_off + ((((uint32_t)kc & 7u) ^ sw) << 4)) = q;
  }
}

template <bool WGRAD>
__global__ void __launch_bounds__(NTHR, 1)
stem_kernel(const __grid_constant__ CUtensorMap tmW /* fwd: weights [64][192]; wgrad: dy [P][64] */,
            const __grid_constant__ CUtensorMap tmX /* x fp32 NCHW as {W, H, 3, B}, box {136, 7, 3, 1} */, SGeo g, bf16* __restrict__ y,
            double* __restrict__ stats, float* __restrict__ dw) {
  constexpr int NBLK = WGRAD ? 4 : 3;            // wgrad pads M to 256 = 4 blocks (the 4th stays zero)
  constexpr int A_BYTES = NBLK * BLK;
  constexpr int W_BYTES = WGRAD ? 2 * BLK : 3 * 8192;  // wgrad: two dy tiles [128 px][64] ; fwd: 3 blocks [64 n][64 k]
  constexpr int TMEM_COLS = 128;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment as an OFFSET from the __shared__ array: the pointer keeps its address space, so the epilogue scratch
  // compiles to STS / LDS (a round trip through uintptr_t made nvcc emit generic ST.E / LD.E)
  uint8_t* smem = smem_raw + ((1024u - ((uint32_t)__cvta_generic_to_shared(smem_raw) & 1023u)) & 1023u);
  uint8_t* a_buf = smem;                        // [2][A_BYTES]
  uint8_t* w_buf = smem + 2 * A_BYTES;          // weights / dy tiles
  uint8_t* patch = w_buf + W_BYTES;             // [2][PATCH_BYTES] fp32
  uint64_t* bars = reinterpret_cast<uint64_t*>(patch + 2 * PATCH_BYTES);
  const uint32_t afull0 = smem_u32(bars), aempty0 = afull0 + 16, tfull0 = afull0 + 32, tempty0 = afull0 + 48, wfull0 = afull0 + 64,
                 wempty0 = afull0 + 80, pfull0 = afull0 + 96, pempty0 = afull0 + 112;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
  float* cta_stats = reinterpret_cast<float*>(bars + 18);  // [8 warps][2][64]: one private slot per epilogue warp (16-byte aligned)
  uint8_t* epi_scr_all = reinterpret_cast<uint8_t*>(cta_stats + 8 * 128);  // [8 warps][2048]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(afull0 + 8 * i, 8);    // one arrival per producer warp
      mbar_init(aempty0 + 8 * i, 1);
      mbar_init(tfull0 + 8 * i, 1);
      mbar_init(tempty0 + 8 * i, 8);
      mbar_init(wfull0 + 8 * i, 1);
      mbar_init(wempty0 + 8 * i, 1);
      mbar_init(pfull0 + 8 * i, 1);
      mbar_init(pempty0 + 8 * i, 8);   // one arrival per producer warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
  }
  // the operand tiles start as zeros: chunk groups 21..23 (and wgrad's 4th block) are never written again
  for (int i = tid; i < 2 * A_BYTES / 16; i += NTHR) reinterpret_cast<uint4*>(a_buf)[i] = make_uint4(0, 0, 0, 0);
  for (int i = tid; i < 8 * 128; i += NTHR) cta_stats[i] = 0.f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pm_pdl_sync();  // the prologue above only touched shared memory / TMEM
  const int my_tiles = (int)blockIdx.x < g.ntiles ? (g.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (warp == 1) {
    // ------------------------------------------------------------ weights (fwd, once) / dy tiles (wgrad, per tile)
    if (!WGRAD) {
      if (my_tiles > 0 && elect_one()) {
        mbar_expect_tx(wfull0, 3 * 8192);
        for (int kb = 0; kb < 3; ++kb) tma_load_2d(smem_u32(w_buf) + kb * 8192, &tmW, wfull0, kb * 64, 0);
      }
      __syncwarp();
    } else {
      for (int it = 0; it < my_tiles; ++it) {
        const int tile = blockIdx.x + it * gridDim.x;
        int b, oh, ow0;
        tile_coords(g, tile, b, oh, ow0);
        const int buf = it & 1;
        mbar_wait(wempty0 + 8 * buf, ((it >> 1) & 1) ^ 1);
        if (elect_one()) {
          mbar_expect_tx(wfull0 + 8 * buf, BLK);
          tma_load_2d(smem_u32(w_buf) + buf * BLK, &tmW, wfull0 + 8 * buf, 0, (b * g.Ho + oh) * g.Wo + ow0);
        }
        __syncwarp();
      }
    }
  } else if (warp == 0) {
    // ------------------------------------------------------------ MMA issuer
    const uint32_t a0 = DESC_SW128_LO + (smem_u32(a_buf) >> 4), w0 = DESC_SW128_LO + (smem_u32(w_buf) >> 4);
    if (!WGRAD) {
      constexpr uint32_t idesc = make_idesc(128, 64, 0, 0);
      if (my_tiles > 0) mbar_wait(wfull0, 0);
      for (int it = 0; it < my_tiles; ++it) {
        const int buf = it & 1, as = it & 1;
        mbar_wait(tempty0 + 8 * as, ((it >> 1) & 1) ^ 1);
        mbar_wait(afull0 + 8 * buf, (it >> 1) & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t ab = a0 + buf * (A_BYTES >> 4);
#pragma unroll
          for (int ks = 0; ks < 11; ++ks)  // k = 16*ks .. : block ks/4, 32 B per step inside the 128-byte swizzle row
            umma_bf16(tmem_base + as * 64, desc_pack(ab + (ks >> 2) * (BLK >> 4) + (ks & 3) * 2, DESC_SW128_HI),
                      desc_pack(w0 + (ks >> 2) * (8192 >> 4) + (ks & 3) * 2, DESC_SW128_HI), idesc, ks != 0);
          umma_commit(aempty0 + 8 * buf);
          umma_commit(tfull0 + 8 * as);
        }
        __syncwarp();
      }
    } else {
      // MN-major operands: LBO = distance between 64-wide M (or N) blocks, SBO = 1024 between 8-row K groups
      constexpr uint32_t idesc = make_idesc(128, 64, 1, 1);
      constexpr uint32_t HI = (1024u >> 4) | (1u << 14) | (2u << 29);
      constexpr uint32_t LO = (uint32_t)(BLK >> 4) << 16;
      const uint32_t a0m = LO + (smem_u32(a_buf) >> 4), w0m = LO + (smem_u32(w_buf) >> 4);
      for (int it = 0; it < my_tiles; ++it) {
        const int buf = it & 1;
        mbar_wait(afull0 + 8 * buf, (it >> 1) & 1);
        mbar_wait(wfull0 + 8 * buf, (it >> 1) & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t ab = a0m + buf * (A_BYTES >> 4), wb = w0m + buf * (BLK >> 4);
#pragma unroll
          for (int half = 0; half < 2; ++half)
#pragma unroll
            for (int ks = 0; ks < 8; ++ks)  // 16 pixels per step = 2048 B along K
              umma_bf16(tmem_base + half * 64, desc_pack(ab + half * (2 * BLK >> 4) + ks * 128, HI), desc_pack(wb + ks * 128, HI),
                        idesc, (it | ks) != 0);
          umma_commit(aempty0 + 8 * buf);
          umma_commit(wempty0 + 8 * buf);
          if (it == my_tiles - 1) umma_commit(tfull0);
        }
        __syncwarp();
      }
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------ input patches (TMA, zero fill outside the image)
    for (int it = 0; it < my_tiles; ++it) {
      const int tile = blockIdx.x + it * gridDim.x;
      int b, oh, ow0;
      tile_coords(g, tile, b, oh, ow0);
      const int buf = it & 1;
      mbar_wait(pempty0 + 8 * buf, ((it >> 1) & 1) ^ 1);
      if (elect_one()) {
        const uint32_t dst = smem_u32(patch) + buf * PATCH_BYTES, full = pfull0 + 8 * buf;
        mbar_expect_tx(full, 2 * PBOX_TX);
        tma_load_4d(dst, &tmX, full, 2 * ow0 - 4, 2 * oh - 3, 0, b);
        tma_load_4d(dst + PBOX_BYTES, &tmX, full, 2 * ow0 - 4 + 128, 2 * oh - 3, 0, b);
      }
      __syncwarp();
    }
  } else if (warp >= 4 && warp < 12) {
    // ------------------------------------------------------------ producers
    const int pt = tid - 128;
    for (int it = 0; it < my_tiles; ++it) {
      const int tile = blockIdx.x + it * gridDim.x;
      const int buf = it & 1;
      int b, oh, ow0;
      tile_coords(g, tile, b, oh, ow0);
      mbar_wait(pfull0 + 8 * buf, (it >> 1) & 1);
      mbar_wait(aempty0 + 8 * buf, ((it >> 1) & 1) ^ 1);
      build_tile<WGRAD>(patch + buf * PATCH_BYTES, a_buf + buf * A_BYTES, pt, g.Wo - ow0);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(afull0 + 8 * buf);
        mbar_arrive(pempty0 + 8 * buf);
      }
    }
  } else if (warp >= 12) {
    // ------------------------------------------------------------ epilogue
    const int quad = warp & 3, ch = (warp - 12) >> 2;  // TMEM lane quarter, 32-column chunk (fwd) / accumulator half (wgrad)
    uint8_t* epi_scr = epi_scr_all + (warp - 12) * 2048;
    if (!WGRAD) {
      float st[4] = {0.f, 0.f, 0.f, 0.f};
      for (int it = 0; it < my_tiles; ++it) {
        const int tile = blockIdx.x + it * gridDim.x;
        const int as = it & 1;
        int b, oh, ow0;
        tile_coords(g, tile, b, oh, ow0);
        const int m = quad * 32 + lane;
        const bool valid = ow0 + m < g.Wo;
        bf16* out = y + (((size_t)b * g.Ho + oh) * g.Wo + ow0 + m) * 64;
        mbar_wait(tfull0 + 8 * as, (it >> 1) & 1);
        tc_fence_after();
        {
          uint32_t acc[32];
          tmem_ld32_nowait(tmem_base + as * 64 + ch * 32 + ((uint32_t)(quad * 32) << 16), acc);
          tmem_ld_wait();
          epilogue_chunk32(acc, valid, out + ch * 32, false, stats != nullptr, epi_scr, lane, st);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty0 + 8 * as);
      }
      if (stats) stats_flush32(st, lane, cta_stats + (warp - 12) * 128 + ch * 32, cta_stats + (warp - 12) * 128 + 64 + ch * 32);
    } else if (my_tiles > 0) {
      mbar_wait(tfull0, 0);
      tc_fence_after();
      {
        const int half = ch;
        const int kk = half * 128 + quad * 32 + lane;   // D row = k = (r*3 + c)*8 + s
        const int kc = kk >> 3, s = kk & 7;
        const int r = kc / 3, c = kc - r * 3;
        const bool ok = kc < PROWS && s < 7;
#pragma unroll 1
        for (int cc = 0; cc < 2; ++cc) {
          uint32_t acc[32];
          tmem_ld32_nowait(tmem_base + half * 64 + cc * 32 + ((uint32_t)(quad * 32) << 16), acc);
          tmem_ld_wait();
          if (ok) {
#pragma unroll
            for (int e = 0; e < 32; ++e) atomicAdd(dw + ((size_t)((cc * 32 + e) * 7 + r) * 7 + s) * 3 + c, __uint_as_float(acc[e]));
          }
        }
      }
      tc_fence_before();
    }
  }
  __syncthreads();
  if (!WGRAD && stats && tid < 128) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += (double)cta_stats[w * 128 + tid];  // fixed order
    if (t != 0.0) atomicAdd(stats + tid, t);
  }
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// conv1.weight fp32 KRSC [64][7][7][3] -> bf16 [64][192], k = (r*3 + c)*8 + s, zero for s == 7 and k >= 168
__global__ void stem_prep_w_kernel(const float* __restrict__ w, bf16* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 64 * 192) return;
  const int n = i / 192, k = i - n * 192;
  const int kc = k >> 3, s = k & 7;
  const int r = kc / 3, c = kc - r * 3;
  const float v = (kc < PROWS && s < 7) ? w[((n * 7 + r) * 7 + s) * 3 + c] : 0.f;
  out[i] = __float2bfloat16_rn(v);
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
static bool map_dense(CUtensorMap* tm, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  return g_tiled(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// x fp32 NCHW [B][3][H][W] -> box {136 pixels, 7 rows, 3 channels, 1 image}, no swizzle, zero fill out of bounds
static bool map_input(CUtensorMap* tm, const void* x, int B, int H, int W) {
  cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, 3, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4, (cuuint64_t)3 * H * W * 4};
  cuuint32_t box[4] = {PBW, 7, 3, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  return g_tiled(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                 CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
static SGeo geo(int B, int H, int W) {
  SGeo g;
  g.B = B; g.H = H; g.W = W;
  g.Ho = (H + 6 - 7) / 2 + 1;
  g.Wo = (W + 6 - 7) / 2 + 1;
  g.tiles_per_row = (g.Wo + 127) / 128;
  g.ntiles = B * g.Ho * g.tiles_per_row;
  return g;
}
template <bool WGRAD>
constexpr int smem_bytes() { return 2 * (WGRAD ? 4 : 3) * BLK + (WGRAD ? 2 * BLK : 3 * 8192) + 2 * PATCH_BYTES + 18 * 8 + 8 * 128 * 4 + 8 * 2048 + 1024; }

}  // namespace stem

extern "C" {

int pm_stem_prep_w_bf16(const float* w_krsc, void* w192, pm_stream_t s) {
  PM_CHECK_ARG(w_krsc && w192);
  stem::stem_prep_w_kernel<<<(64 * 192 + 255) / 256, 256, 0, S(s)>>>(w_krsc, (stem::bf16*)w192);
  PM_LAUNCH_OK();
}

int pm_stem_conv_fwd_bf16(const float* x_nchw, const void* w192, int B, int H, int W, void* y, double* stats, pm_stream_t s) {
  using namespace stem;
  PM_CHECK_ARG(x_nchw && w192 && y && B > 0 && H >= 7 && W >= 7 && W % 4 == 0 && ((uintptr_t)x_nchw & 15) == 0);
  if (!load_driver()) return pm_set_err(__FILE__, __LINE__, "cuTensorMapEncodeTiled unavailable");
  const SGeo g = geo(B, H, W);
  CUtensorMap tm, tmx;
  if (!map_input(&tmx, x_nchw, B, H, W)) return pm_set_err(__FILE__, __LINE__, "stem input tensor map failed");
  if (!map_dense(&tm, w192, 64, 192, 64)) return pm_set_err(__FILE__, __LINE__, "stem weight tensor map failed");
  PM_CUDA(cudaFuncSetAttribute(stem_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes<false>()));
  PM_CUDA(pm_launch(stem_kernel<false>, dim3(std::min(pm_num_sms(), g.ntiles)), dim3(NTHR), (size_t)smem_bytes<false>(), S(s), tm, tmx, g,
                    (bf16*)y, stats, (float*)nullptr));
  PM_LAUNCH_OK();
}

int pm_stem_conv_wgrad_bf16(const float* x_nchw, const void* dy, int B, int H, int W, float* dw_krsc, pm_stream_t s) {
  using namespace stem;
  PM_CHECK_ARG(x_nchw && dy && dw_krsc && B > 0 && H >= 7 && W >= 7 && W % 4 == 0 && ((uintptr_t)x_nchw & 15) == 0);
  if (!load_driver()) return pm_set_err(__FILE__, __LINE__, "cuTensorMapEncodeTiled unavailable");
  const SGeo g = geo(B, H, W);
  CUtensorMap tm, tmx;
  if (!map_input(&tmx, x_nchw, B, H, W)) return pm_set_err(__FILE__, __LINE__, "stem input tensor map failed");
  if (!map_dense(&tm, dy, (uint64_t)B * g.Ho * g.Wo, 64, 128)) return pm_set_err(__FILE__, __LINE__, "stem dy tensor map failed");
  PM_CUDA(cudaFuncSetAttribute(stem_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes<true>()));
  PM_CUDA(pm_launch(stem_kernel<true>, dim3(std::min(pm_num_sms(), g.ntiles)), dim3(NTHR), (size_t)smem_bytes<true>(), S(s), tm, tmx, g,
                    (bf16*)nullptr, (double*)nullptr, dw_krsc));
  PM_LAUNCH_OK();
}

}  // extern "C"
