// Path T, throughput mode: 3x3 / stride-1 / pad-1 convolutions (forward and data gradient) as HALO-STRIP implicit GEMMs.
//
// Why: the im2col-TMA kernels (conv_tma.cu) fetch every filter tap's [128 pixels x 64 channels] operand tile separately
// from L2 -- 9 fetches of (almost) the same pixels -- and sit on the L2->SM throughput cap (~43 B/clk/SM), not on the
// tensor pipe.  Here the activation strip that covers a tile's pixels plus its one-pixel halo is loaded ONCE per 64-channel
// block and the 9 taps are 9 row-shifted views of that strip: the K-major SWIZZLE_128B shared-memory descriptor may start at
// any 128-byte row (the swizzle XOR works on absolute address bits; scripts/ubench/umma_shift.cu).
//
// Virtual pixel grid: every image row is stored in shared memory as [0, x_0 .. x_{W-1}] (Wp = W+1 pixels, the leading zero
// is TMA out-of-bounds fill) and every image as [zero row, row_0 .. row_{H-1}] (Hp = H+1 rows): the right/bottom padding of
// one row/image is the left/top padding of the next.  With p = (img*Hp + hh)*Wp + ww the linear padded index, output (img,h,w)
// has virtual index q = (img*Hp + h)*Wp + w and its tap (r,s) reads p = q + r*Wp + s: a constant row shift.  Virtual pixels
// with w == W or h == H are computed and discarded (efficiency H*W / (Hp*Wp): 96.5 % at 56x56 ... 76.6 % at 7x7).
//
// CTA (persistent, one per SM): tile = MT (128 | 256) virtual pixels x BN (64 | 128) output channels.
//   warp 0 : strip producer  -- one TMA (tiled 4-D box {64 ch, Wp, 1, 1}) per padded row, SSTAGES strip ring;
//   warp 1 : tcgen05.mma issuer (M = 128 per instruction, MT/128 accumulators in TMEM, double-buffered across tiles);
//   warp 2 : weight-tile producer ({64 k, BN} boxes, BSTAGES ring) -- or the whole weight matrix once when it fits (RB);
//   warps 4.. : epilogue, 4 warps per 128-row accumulator: tcgen05.ld -> bf16 store (+ accumulate) and, for the forward,
//            the BatchNorm batch statistics (sum, sum of squares of the bf16-rounded outputs); rows are staged through a per-warp
//            shared-memory scratch so that global stores are 64-byte contiguous per row (tc_ptx.cuh: epilogue_chunk32).
#include "common.cuh"
#include "tc_ptx.cuh"
#include <algorithm>
#include <stdlib.h>

namespace halo {
using namespace tcx;
typedef __nv_bfloat16 bf16;

struct HGeo {
  int B, H, W, Cin, N;   // Cin: reduction channels, N: output channels
  int Wp, Hp, V;         // padded row / image pitch, number of virtual pixels
  int strip_bytes;       // one strip stage (max rows * Wp * 128, rounded to 1024)
  int b_bytes;           // weight region: BSTAGES * BN * 128, or the whole matrix (RB)
  int prof;              // != 0: per-CTA wait/busy cycle counters into halo_prof (diagnostics, scripts/conv_bench.py)
};

// [cta][8]: 0 kernel, 1 mma wait strip, 2 mma wait weights, 3 mma wait tmem-empty, 4 epilogue wait tmem-full, 5 epilogue busy,
//           6 strip producer wait empty, 7 weight producer wait empty   (clock64 ticks, summed over the CTA's tiles)
__device__ long long halo_prof[160 * 8];
__device__ long long halo_epi_prof[160 * 4];  // warp 4: tcgen05.ld+wait, epilogue_chunk32, (unused), chunks
// CTA 0 event trace: [2*i] = event code, [2*i+1] = clock64 ; codes: 1 mma tile begin (after tmem-empty), 2 mma strip ready,
// 3 mma tile issued, 4 epi(warp 4) accumulator ready, 5 epi done, 6 strip producer: slot free, 7 strip producer: rows issued
__device__ long long halo_trace[2 * 256];
__device__ int halo_trace_n;
#define TRACE(code) if (g.prof && blockIdx.x == 0) { const int _i = atomicAdd(&halo_trace_n, 1); if (_i < 256) { halo_trace[2 * _i] = (code); halo_trace[2 * _i + 1] = clock64(); } }
#define PROF_T(var) const long long var = g.prof ? clock64() : 0
#define PROF_ADD(acc, t0) if (g.prof) acc += clock64() - (t0)

constexpr int BSTAGES = 4;

// SSTAGES: depth of the activation-strip ring (3 when shared memory allows: the producer then runs two strips ahead of the MMA
// warp across tile boundaries); ACC: the data gradient adds into dst (identity branch) -- a template parameter so that the
// unpack / add / repack of the old values is not even compiled into the forward and the plain data gradients.
template <int MT, int BN, bool FLIP, bool RB, int SSTAGES, bool ACC>
__global__ void __launch_bounds__(128 + MT, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, HGeo g, bf16* __restrict__ dst,
                 double* __restrict__ stats) {
  constexpr int NH = MT / 128;
  constexpr int NTHR = 128 + MT;
  constexpr int B_BYTES = BN * 128;
  constexpr int ACC_COLS = NH * BN;
  constexpr int TMEM_COLS = 2 * ACC_COLS;
  static_assert(TMEM_COLS == 128 || TMEM_COLS == 256 || TMEM_COLS == 512, "TMEM columns must be a power of two <= 512");
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment as an OFFSET from the __shared__ array: the pointer keeps its address space, so the epilogue scratch
  // compiles to STS / LDS (a round trip through uintptr_t made nvcc emit generic ST.E / LD.E)
  uint8_t* smem = smem_raw + ((1024u - ((uint32_t)__cvta_generic_to_shared(smem_raw) & 1023u)) & 1023u);
  const uint32_t strip0 = smem_u32(smem);
  const uint32_t b_base = strip0 + SSTAGES * g.strip_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SSTAGES * g.strip_bytes + g.b_bytes);
  const uint32_t sfull0 = smem_u32(bars), sempty0 = sfull0 + 8 * SSTAGES;
  const uint32_t bfull0 = sempty0 + 8 * SSTAGES, bempty0 = bfull0 + 8 * BSTAGES;
  const uint32_t tfull0 = bempty0 + 8 * BSTAGES, tempty0 = tfull0 + 16, bres = tempty0 + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * SSTAGES + 2 * BSTAGES + 5);
  constexpr int NEW = 4 * NH;                                                          // epilogue warps
  float* cta_stats = reinterpret_cast<float*>(bars + 2 * SSTAGES + 2 * BSTAGES + 6);  // [NEW warps][2][N]: one private slot per warp
  uint8_t* epi_scr_all = reinterpret_cast<uint8_t*>(cta_stats + NEW * 2 * g.N);        // [NEW][2048]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  PROF_T(t_kernel);
  long long pw0 = 0, pw1 = 0, pw2 = 0, pe0 = 0, pe1 = 0, pe3 = 0;
  const int cblocks = g.Cin / 64;
  const int nkb = 9 * cblocks;
  const int ntn = g.N / BN;
  const int ntiles = ((g.V + MT - 1) / MT) * ntn;
  const bool do_stats = !FLIP && stats != nullptr;

  if (tid == 0) {
    for (int s = 0; s < SSTAGES; ++s) {
      mbar_init(sfull0 + 8 * s, 1);
      mbar_init(sempty0 + 8 * s, 1);
    }
    for (int s = 0; s < BSTAGES; ++s) {
      mbar_init(bfull0 + 8 * s, 1);
      mbar_init(bempty0 + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull0 + 8 * a, 1);
      mbar_init(tempty0 + 8 * a, 4 * NH);  // one arrival per epilogue warp
    }
    mbar_init(bres, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
  }
  if (do_stats)
    for (int i = tid; i < NEW * 2 * g.N; i += NTHR) cta_stats[i] = 0.f;
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pm_pdl_sync();  // everything above is CTA-local; the first global access (TMA, stores) is below

  if (warp == 0) {
    // ------------------------------------------------------------ strip producer: one TMA per padded row
    const uint32_t row_bytes = (uint32_t)g.Wp * 128u;
    uint32_t ss = 0, sph = 1;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int q0 = (tile / ntn) * MT;
      const int R0 = q0 / g.Wp, o = q0 - R0 * g.Wp;
      const int nrows = (o + MT + 2 * g.Wp + 1) / g.Wp + 1;
      const int img0 = R0 / g.Hp, hh0 = R0 - img0 * g.Hp;
      for (int cb = 0; cb < cblocks; ++cb) {
        PROF_T(t0);
        mbar_wait(sempty0 + 8 * ss, sph);
        PROF_ADD(pw0, t0);
        if (elect_one()) {
          const uint32_t full = sfull0 + 8 * ss;
          mbar_expect_tx(full, (uint32_t)nrows * row_bytes);
          uint32_t dstrow = strip0 + ss * g.strip_bytes;
          int img = img0, hh = hh0;
          for (int i = 0; i < nrows; ++i, dstrow += row_bytes) {
            // hh == 0 is the shared zero row (coordinate -1: entirely out of bounds -> zero fill), as is img >= B
            tma_load_4d(dstrow, &tmA, full, cb * 64, -1, hh - 1, img);
            if (++hh == g.Hp) { hh = 0; ++img; }
          }
        }
        __syncwarp();
        if (++ss == SSTAGES) { ss = 0; sph ^= 1; }
      }
    }
    if (g.prof && lane == 0) halo_prof[blockIdx.x * 8 + 6] = pw0;
  } else if (warp == 2) {
    // ------------------------------------------------------------ weight producer
    if (RB) {
      if ((int)blockIdx.x < ntiles && elect_one()) {
        mbar_expect_tx(bres, (uint32_t)(nkb * B_BYTES));
        for (int kb = 0; kb < nkb; ++kb) tma_load_2d(b_base + kb * B_BYTES, &tmB, bres, kb * 64, 0);
      }
      __syncwarp();
    } else {
      uint32_t bs = 0, bph = 1;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int n0 = (tile % ntn) * BN;
        for (int cb = 0; cb < cblocks; ++cb)
          for (int tap = 0; tap < 9; ++tap) {
            PROF_T(t0);
            mbar_wait(bempty0 + 8 * bs, bph);
            PROF_ADD(pw0, t0);
            if (elect_one()) {
              mbar_expect_tx(bfull0 + 8 * bs, B_BYTES);
              tma_load_2d(b_base + bs * B_BYTES, &tmB, bfull0 + 8 * bs, (tap * cblocks + cb) * 64, n0);
            }
            __syncwarp();
            if (++bs == BSTAGES) { bs = 0; bph ^= 1; }
          }
      }
    }
    if (g.prof && lane == 0) halo_prof[blockIdx.x * 8 + 7] = pw0;
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (whole warp converged, one elected lane issues)
    constexpr uint32_t idesc = make_idesc(128, BN, 0, 0);
    uint32_t tapoff[9];  // row shift of tap (r, s) in 16-byte descriptor units
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      int r = tap / 3, sx = tap - r * 3;
      if (FLIP) { r = 2 - r; sx = 2 - sx; }
      tapoff[tap] = (uint32_t)(r * g.Wp + sx) * 8u;
    }
    uint32_t ss = 0, sph = 0, bs = 0, bph = 0;
    int lt = 0;
    if (RB && (int)blockIdx.x < ntiles) mbar_wait(bres, 0);
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++lt) {
      const int as = lt & 1;
      const int q0 = (tile / ntn) * MT;
      const int o = q0 % g.Wp;
      PROF_T(t2);
      mbar_wait(tempty0 + 8 * as, ((lt >> 1) & 1) ^ 1);  // the epilogue has drained this accumulator stage
      PROF_ADD(pw2, t2);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + as * ACC_COLS;
      for (int cb = 0; cb < cblocks; ++cb) {
        PROF_T(t0);
        mbar_wait(sfull0 + 8 * ss, sph);
        PROF_ADD(pw0, t0);
        tc_fence_after();
        const uint32_t a0 = DESC_SW128_LO + ((strip0 + ss * g.strip_bytes) >> 4) + (uint32_t)o * 8u;
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          uint32_t b_lo;
          if (RB) {
            b_lo = DESC_SW128_LO + ((b_base + (tap * cblocks + cb) * B_BYTES) >> 4);
          } else {
            PROF_T(t1);
            mbar_wait(bfull0 + 8 * bs, bph);
            PROF_ADD(pw1, t1);
            tc_fence_after();
            b_lo = DESC_SW128_LO + ((b_base + bs * B_BYTES) >> 4);
          }
          if (elect_one()) {
#pragma unroll
            for (int j = 0; j < NH; ++j) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16(tmem_d + j * BN, desc_pack(a0 + tapoff[tap] + (uint32_t)(j * 1024 + k * 2), DESC_SW128_HI),
                          desc_pack(b_lo + (uint32_t)(k * 2), DESC_SW128_HI), idesc, (tap | k) != 0 ? 1u : (uint32_t)cb);
            }
            if (!RB) umma_commit(bempty0 + 8 * bs);
            if (tap == 8) umma_commit(sempty0 + 8 * ss);
          }
          __syncwarp();
          if (!RB) {
            if (++bs == BSTAGES) { bs = 0; bph ^= 1; }
          }
        }
        if (++ss == SSTAGES) { ss = 0; sph ^= 1; }
      }
      if (elect_one()) umma_commit(tfull0 + 8 * as);
      __syncwarp();
    }
    if (g.prof && lane == 0) {
      halo_prof[blockIdx.x * 8 + 1] = pw0;
      halo_prof[blockIdx.x * 8 + 2] = pw1;
      halo_prof[blockIdx.x * 8 + 3] = pw2;
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue (4 warps per 128-row accumulator)
    const int j = (warp - 4) >> 2, quad = warp & 3;
    uint8_t* epi_scr = epi_scr_all + (warp - 4) * 2048;
    float* my_stats = cta_stats + (warp - 4) * 2 * g.N;
    float st[BN / 32][4];  // BatchNorm partial sums of this warp's rows, per 32-column chunk, kept across tiles
#pragma unroll
    for (int cc = 0; cc < BN / 32; ++cc)
#pragma unroll
      for (int k = 0; k < 4; ++k) st[cc][k] = 0.f;
    int st_n0 = -1;
    int lt = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++lt) {
      const int as = lt & 1;
      const int n0 = (tile % ntn) * BN;
      if (do_stats && n0 != st_n0) {  // the column range changes between this CTA's tiles only if gridDim % ntn != 0
        if (st_n0 >= 0) {
#pragma unroll
          for (int cc = 0; cc < BN / 32; ++cc) stats_flush32(st[cc], lane, my_stats + st_n0 + cc * 32, my_stats + g.N + st_n0 + cc * 32);
        }
        st_n0 = n0;
      }
      const int q = (tile / ntn) * MT + j * 128 + quad * 32 + lane;
      const int Rr = q / g.Wp, w = q - Rr * g.Wp;
      const int img = Rr / g.Hp, h = Rr - img * g.Hp;
      const bool valid = w < g.W && h < g.H && img < g.B;
      bf16* out = dst + (((size_t)img * g.H + h) * g.W + w) * g.N + n0;
      PROF_T(t0);
      mbar_wait(tfull0 + 8 * as, (lt >> 1) & 1);
      PROF_ADD(pw0, t0);
      PROF_T(t1);
      tc_fence_after();
      const uint32_t tsrc = tmem_base + as * ACC_COLS + j * BN + ((uint32_t)(quad * 32) << 16);
#pragma unroll
      for (int cc = 0; cc < BN / 32; ++cc) {
        uint32_t v[32];
        PROF_T(e0);
        tmem_ld32_nowait(tsrc + cc * 32, v);
        tmem_ld_wait();
        PROF_ADD(pe0, e0);
        PROF_T(e1);
        epilogue_chunk32(v, valid, out + cc * 32, ACC, do_stats, epi_scr, lane, st[cc]);
        PROF_ADD(pe1, e1);
        ++pe3;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty0 + 8 * as);  // this warp is done reading the accumulator stage
      PROF_ADD(pw1, t1);
    }
    if (do_stats && st_n0 >= 0) {
#pragma unroll
      for (int cc = 0; cc < BN / 32; ++cc) stats_flush32(st[cc], lane, my_stats + st_n0 + cc * 32, my_stats + g.N + st_n0 + cc * 32);
    }
    if (g.prof && tid == 128) {
      halo_prof[blockIdx.x * 8 + 4] = pw0;
      halo_prof[blockIdx.x * 8 + 5] = pw1;
      halo_epi_prof[blockIdx.x * 4 + 0] = pe0;
      halo_epi_prof[blockIdx.x * 4 + 1] = pe1;
      halo_epi_prof[blockIdx.x * 4 + 3] = pe3;
    }
  }
  __syncthreads();
  if (g.prof && tid == 0) {
    halo_prof[blockIdx.x * 8 + 0] = clock64() - t_kernel;
    if (blockIdx.x == 0) { const int _i = atomicAdd(&halo_trace_n, 1); if (_i < 256) { halo_trace[2 * _i] = 0; halo_trace[2 * _i + 1] = t_kernel; } }
  }
  if (do_stats)
    for (int i = tid; i < 2 * g.N; i += NTHR) {
      double v = 0.0;
#pragma unroll
      for (int w = 0; w < NEW; ++w) v += (double)cta_stats[w * 2 * g.N + i];  // fixed order
      if (v != 0.0) atomicAdd(stats + i, v);
    }
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------ host
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
// NHWC bf16 activation [B][H][W][C]: one padded row [col -1 .. W-1] x 64 channels per load
static bool map_rows(CUtensorMap* tm, const void* base, int B, int H, int W, int C) {
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)(W + 1), 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  return g_tiled(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static int strip_stages() {
  static int v = 0;
  if (!v) {
    const char* e = getenv("PRIMIA_HALO_SSTAGES");
    v = (e && e[0] == '2') ? 2 : 3;
  }
  return v;
}

template <int MT, int BN, bool FLIP, bool RB, int SSTAGES, bool ACC>
static int launch_ss(const CUtensorMap& tmA, const CUtensorMap& tmB, HGeo g, void* dst, double* stats, cudaStream_t st) {
  const int max_rows = (g.Wp - 1 + MT + 2 * g.Wp + 1) / g.Wp + 1;
  g.strip_bytes = (max_rows * g.Wp * 128 + 1023) / 1024 * 1024;
  g.b_bytes = RB ? 9 * (g.Cin / 64) * BN * 128 : BSTAGES * BN * 128;
  const int smem = SSTAGES * g.strip_bytes + g.b_bytes + 256 + (MT / 32) * 2 * g.N * 4 + (MT / 32) * 2048 + 1024;
  if (smem > 227 * 1024) return 1;
  auto kern = conv_halo_kernel<MT, BN, FLIP, RB, SSTAGES, ACC>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return 2;
  const int ntiles = ((g.V + MT - 1) / MT) * (g.N / BN);
  dim3 grid(std::min(pm_num_sms(), ntiles));
  if (pm_launch(kern, grid, dim3(128 + MT), (size_t)smem, st, tmA, tmB, g, (bf16*)dst, stats) != cudaSuccess) return 2;
  return 0;
}

template <int MT, int BN, bool FLIP, bool RB>
static int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, HGeo g, void* dst, int accumulate, double* stats, cudaStream_t st) {
  // deepest strip ring that fits; the forward never accumulates
  if (FLIP && accumulate) {
    if (strip_stages() == 3) {
      const int r = launch_ss<MT, BN, FLIP, RB, 3, FLIP>(tmA, tmB, g, dst, stats, st);
      if (r != 1) return r;
    }
    return launch_ss<MT, BN, FLIP, RB, 2, FLIP>(tmA, tmB, g, dst, stats, st);
  }
  if (strip_stages() == 3) {
    const int r = launch_ss<MT, BN, FLIP, RB, 3, false>(tmA, tmB, g, dst, stats, st);
    if (r != 1) return r;
  }
  return launch_ss<MT, BN, FLIP, RB, 2, false>(tmA, tmB, g, dst, stats, st);
}

}  // namespace halo

static int g_halo_prof = 0;
extern "C" int pm_halo_prof(int enable, int64_t* out_host /* [160*8 + 1 + 512] or NULL */) {
  g_halo_prof = enable;
  if (out_host) {
    if (cudaMemcpyFromSymbol(out_host, halo::halo_prof, sizeof(long long) * 160 * 8) != cudaSuccess) return 1;
    int n = 0;
    if (cudaMemcpyFromSymbol(&n, halo::halo_trace_n, sizeof(int)) != cudaSuccess) return 1;
    out_host[160 * 8] = n < 256 ? n : 256;
    if (cudaMemcpyFromSymbol(out_host + 160 * 8 + 1, halo::halo_epi_prof, sizeof(long long) * 512) != cudaSuccess) return 1;
  }
  if (enable) {
    const int zero = 0;
    if (cudaMemcpyToSymbol(halo::halo_trace_n, &zero, sizeof(int)) != cudaSuccess) return 1;
    static long long zeros[160 * 8];
    if (cudaMemcpyToSymbol(halo::halo_prof, zeros, sizeof(zeros)) != cudaSuccess) return 1;
  }
  return 0;
}

static bool use_halo() {
  const char* e = getenv("PRIMIA_NO_HALO");
  return !(e && e[0] == '1');
}

// 0 = launched, 1 = shape not eligible (caller falls back to the im2col-TMA kernels), 2 = CUDA / driver error.
// flip = 0: forward  (src = x [B,H,W,C], wmat = [K][9*C], dst = y [B,H,W,K]);
// flip = 1: data gradient (src = dy [B,H,W,K], wmat = [C][9*K] (taps mirrored by the kernel), dst = dx [B,H,W,C]).
int pm_halo_conv(const pm_conv_t* p, const void* src, const void* wmat, void* dst, int accumulate, double* stats, int flip,
                 cudaStream_t st) {
  using namespace halo;
  if (!use_halo()) return 1;
  if (p->R != 3 || p->S != 3 || p->stride != 1 || p->pad != 1 || p->Ho != p->H || p->Wo != p->W) return 1;
  if (p->C % 64 != 0 || p->K % 64 != 0 || p->W + 1 > 256 || !load_driver()) return 1;
  HGeo g;
  g.B = p->B; g.H = p->H; g.W = p->W;
  g.Cin = flip ? p->K : p->C;
  g.N = flip ? p->C : p->K;
  g.Wp = p->W + 1; g.Hp = p->H + 1;
  const long V = (long)g.B * g.Hp * g.Wp;
  if (V > 0x7fffffffL - 1024) return 1;
  g.V = (int)V;
  g.strip_bytes = g.b_bytes = 0;
  g.prof = g_halo_prof;
  CUtensorMap tmA, tmB;
  if (!map_rows(&tmA, src, g.B, g.H, g.W, g.Cin)) return 2;
  const int Ktot = 9 * g.Cin;
  const bool rb = g.N == 64 && g.Cin == 64;
  const int BN = g.N % 128 == 0 ? 128 : 64;
  if (!map_dense(&tmB, wmat, g.N, Ktot, BN)) return 2;
  // 256-pixel tiles halve the weight traffic per MAC; 128-pixel tiles when 256 would leave most SMs idle
  const long tiles256 = ((V + 255) / 256) * (g.N / BN);
  const bool big = tiles256 * 4 >= (long)pm_num_sms() * 3;
  int r;
  if (rb) {
    r = flip ? launch<256, 64, true, true>(tmA, tmB, g, dst, accumulate, stats, st)
             : launch<256, 64, false, true>(tmA, tmB, g, dst, accumulate, stats, st);
  } else if (BN == 128) {
    if (big) r = flip ? launch<256, 128, true, false>(tmA, tmB, g, dst, accumulate, stats, st)
                      : launch<256, 128, false, false>(tmA, tmB, g, dst, accumulate, stats, st);
    else r = flip ? launch<128, 128, true, false>(tmA, tmB, g, dst, accumulate, stats, st)
                  : launch<128, 128, false, false>(tmA, tmB, g, dst, accumulate, stats, st);
  } else {
    r = flip ? launch<256, 64, true, false>(tmA, tmB, g, dst, accumulate, stats, st)
             : launch<256, 64, false, false>(tmA, tmB, g, dst, accumulate, stats, st);
  }
  return r;
}
