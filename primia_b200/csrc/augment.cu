// Training-side image front end on the GPU: the reference augments every image on the CPU with PIL +
// albumentations (torchlib/dataloader.py:138-217 create_albu_transform) before the batch is sent to the hospital's worker; here
// the raw uint8 images of a batch are uploaded once and ONE kernel produces the normalised fp32 NCHW batch the stem reads:
//
//   RandomAffine (PIL ImagingTransform AFFINE, NEAREST, fill 0)   dataloader.py:139-145
//   Resize(inference_resolution)  = cv::resize 8U INTER_LINEAR    :147
//   RandomCrop(train_resolution)                                  :148
//   VerticalFlip, GaussNoise                                      :157,199
//   ToFloat(255), Normalize(mean, std)                            :200-203
//
// One thread = one output pixel.  Nothing intermediate is materialised: an output pixel is the bilinear blend (OpenCV's 11-bit
// fixed-point arithmetic, bit for bit) of four pixels of the affine-warped image, each of which is one nearest-neighbour fetch
// from the source (PIL's 16.16 fixed-point arithmetic, bit for bit).  The random parameters are drawn on the host
// (primia_b200/train/augment.py) and arrive as one descriptor per sample; the per-axis resize tables depend on the source size
// only.  Bound: HBM / L2 gather -- 4 source bytes read and 4 output bytes written per pixel and channel.
#include "common.cuh"

static_assert(sizeof(pm_aug_sample_t) == 80, "pm_aug_sample_t is mirrored field by field in primia_b200/_lib.py (AugSample)");

namespace {

// Philox4x32-10 -> one N(0,1) value per (pixel, channel) counter
__device__ __forceinline__ float philox_normal(uint64_t seed, uint64_t ctr) {
  uint32_t c[4] = {(uint32_t)ctr, (uint32_t)(ctr >> 32), 0x41554721u, 0u};
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  const float u1 = ((float)c[0] + 1.0f) * 2.3283064365386963e-10f;   // (0, 1]
  const float u2 = (float)c[1] * 2.3283064365386963e-10f;            // [0, 1)
  return sqrtf(-2.0f * __logf(u1)) * __cosf(6.283185307179586f * u2);
}

// one pixel of the affine-warped image (PIL Geometry.c affine_fixed): nearest source pixel or the fill value 0
__device__ __forceinline__ int warped(const uint8_t* __restrict__ img, const pm_aug_sample_t& s, int r, int c, int ch) {
  const long long xx = (long long)s.fix[2] + (long long)s.fix[1] * r + (long long)s.fix[0] * c;
  const long long yy = (long long)s.fix[5] + (long long)s.fix[4] * r + (long long)s.fix[3] * c;
  const long long xin = xx >> 16, yin = yy >> 16;
  if (xin < 0 || xin >= s.Ws || yin < 0 || yin >= s.Hs) return 0;
  return img[((size_t)yin * s.Ws + (size_t)xin) * s.C + ch];
}

__global__ void __launch_bounds__(256)
augment_kernel(const uint8_t* __restrict__ src, const pm_aug_sample_t* __restrict__ samples, const int32_t* __restrict__ tables,
               int B, int R, int T, int Cout, float m0, float m1, float m2, float r0, float r1, float r2,
               float* __restrict__ out, uint8_t* __restrict__ out_u8) {
  const size_t total = (size_t)B * T * T;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % T);
    const int y = (int)((i / T) % T);
    const int b = (int)(i / ((size_t)T * T));
    const pm_aug_sample_t s = samples[b];
    const uint8_t* img = src + s.src_off;
    const int ry = (s.flip ? T - 1 - y : y) + s.cy, rx = x + s.cx;     // position in the resized R x R image
    const int32_t* tab = tables + s.tab_off;                            // [8][R]: sx0 sx1 ax0 ax1 sy0 sy1 by0 by1
    int sx0 = 0, sx1 = 0, ax0 = 0, ax1 = 0, sy0 = 0, sy1 = 0, by0 = 0, by1 = 0;
    if (!s.area2) {
      sx0 = tab[rx]; sx1 = tab[R + rx]; ax0 = tab[2 * R + rx]; ax1 = tab[3 * R + rx];
      sy0 = tab[4 * R + ry]; sy1 = tab[5 * R + ry]; by0 = tab[6 * R + ry]; by1 = tab[7 * R + ry];
    }
    for (int co = 0; co < Cout; ++co) {
      const int ch = co < s.C ? co : s.C - 1;                           // a one-channel source is replicated (the RGB loader)
      int v;
      if (s.area2) {   // cv::resize turns an exact 2x INTER_LINEAR down-scale into the 2x2 box mean
        v = (warped(img, s, 2 * ry, 2 * rx, ch) + warped(img, s, 2 * ry, 2 * rx + 1, ch) + warped(img, s, 2 * ry + 1, 2 * rx, ch) +
             warped(img, s, 2 * ry + 1, 2 * rx + 1, ch) + 2) >> 2;
      } else {         // HResizeLinear (int32, x 2048) then VResizeLinear<uchar>
        const int h0 = warped(img, s, sy0, sx0, ch) * ax0 + warped(img, s, sy0, sx1, ch) * ax1;
        const int h1 = warped(img, s, sy1, sx0, ch) * ax0 + warped(img, s, sy1, sx1, ch) * ax1;
        v = (((by0 * (h0 >> 4)) >> 16) + ((by1 * (h1 >> 4)) >> 16) + 2) >> 2;
      }
      if (s.noise_sigma > 0.f) {   // albumentations gauss_noise on uint8: float add, clip to [0, 255], cast (truncation)
        const float f = (float)v + s.noise_sigma * philox_normal(s.noise_seed, ((uint64_t)co * T + y) * T + x);
        v = (int)fminf(fmaxf(f, 0.f), 255.f);
      }
      const size_t o = (((size_t)b * Cout + co) * T + y) * T + x;
      if (out_u8) out_u8[o] = (uint8_t)v;
      const float mean = co == 0 ? m0 : co == 1 ? m1 : m2, rstd = co == 0 ? r0 : co == 1 ? r1 : r2;
      // ToFloat: x / 255 (float32 division); Normalize: (x - mean) * reciprocal(std) -- separate roundings, no contraction
      if (out) out[o] = __fmul_rn(__fsub_rn(__fdiv_rn((float)v, 255.0f), mean), rstd);
    }
  }
}

// ---- CLAHE (dataloader.py:150-156: a.CLAHE(clip_limit=(1, 1)) between the crop and the flip/noise group) -- cv::CLAHE_Impl
// restated (clahe.cpp CLAHE_CalcLut_Body / CLAHE_Interpolation_Body), bit for bit.
// Pass 1: one block per (tile, image): histogram of the tile in shared memory, clip at `clip`, spread the excess (every bin gets
// clipped / 256, the first `residual` bins at stride max(256 / residual, 1) one more), inclusive scan, lut = cvRound(sum * 255/area).
// pre: optional 256-entry table applied to every pixel first (3-channel recipe: grey -> L of cv2.COLOR_RGB2LAB).
__global__ void __launch_bounds__(256)
clahe_lut_kernel(const uint8_t* __restrict__ img, const uint8_t* __restrict__ pre, int T, int tiles, int clip, float lut_scale,
                 uint8_t* __restrict__ luts) {
  __shared__ int hist[256];
  __shared__ int scan[256];
  __shared__ int s_clipped;
  const int tile = blockIdx.x, b = blockIdx.y;
  const int ty = tile / tiles, tx = tile % tiles, ts = T / tiles;
  const uint8_t* base = img + (size_t)b * T * T + (size_t)ty * ts * T + (size_t)tx * ts;
  hist[threadIdx.x] = 0;
  if (threadIdx.x == 0) s_clipped = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < ts * ts; i += blockDim.x) {
    int v = base[(size_t)(i / ts) * T + (i % ts)];
    if (pre) v = pre[v];
    atomicAdd(&hist[v], 1);
  }
  __syncthreads();
  int h = hist[threadIdx.x];
  if (clip > 0) {
    if (h > clip) { atomicAdd(&s_clipped, h - clip); h = clip; }
    __syncthreads();
    const int clipped = s_clipped;
    const int batch = clipped / 256;
    int residual = clipped - batch * 256;
    h += batch;
    if (residual != 0) {
      const int step = max(256 / residual, 1);
      // for (i = 0; i < 256 && residual > 0; i += step, --residual) ++hist[i]
      if (threadIdx.x % step == 0 && threadIdx.x / step < residual) ++h;
    }
  }
  scan[threadIdx.x] = h;
  __syncthreads();
  for (int off = 1; off < 256; off <<= 1) {     // Hillis-Steele inclusive scan over the 256 bins
    const int add = threadIdx.x >= off ? scan[threadIdx.x - off] : 0;
    __syncthreads();
    scan[threadIdx.x] += add;
    __syncthreads();
  }
  const int r = __float2int_rn(__fmul_rn((float)scan[threadIdx.x], lut_scale));   // saturate_cast<uchar>(float) = cvRound, clamp
  luts[((size_t)b * tiles * tiles + tile) * 256 + threadIdx.x] = (uint8_t)min(max(r, 0), 255);
}

// Pass 2: bilinear blend of the four neighbouring tiles' LUTs in float32 (each product and sum rounded separately, as the
// scalar loop of CLAHE_Interpolation_Body), then the tail of the pipeline: VerticalFlip, GaussNoise, ToFloat, Normalize.
// post: optional [256][3] table (3-channel recipe: (L', 128, 128) -> RGB of cv2.COLOR_LAB2RGB).
__global__ void __launch_bounds__(256)
clahe_finish_kernel(const uint8_t* __restrict__ img, const uint8_t* __restrict__ pre, const uint8_t* __restrict__ luts,
                    const uint8_t* __restrict__ post, const pm_aug_sample_t* __restrict__ samples, int B, int T, int tiles, int Cout,
                    float m0, float m1, float m2, float r0, float r1, float r2, float* __restrict__ out, uint8_t* __restrict__ out_u8) {
  const size_t total = (size_t)B * T * T;
  const int ts = T / tiles;
  const float inv_t = __fdiv_rn(1.0f, (float)ts);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % T);
    const int y = (int)((i / T) % T);
    const int b = (int)(i / ((size_t)T * T));
    const pm_aug_sample_t s = samples[b];
    const int sy = s.flip ? T - 1 - y : y;                       // CLAHE ran on the unflipped crop
    int v = img[((size_t)b * T + sy) * T + x];
    if (pre) v = pre[v];
    const float txf = __fsub_rn(__fmul_rn((float)x, inv_t), 0.5f), tyf = __fsub_rn(__fmul_rn((float)sy, inv_t), 0.5f);
    int tx1 = (int)floorf(txf), ty1 = (int)floorf(tyf);
    const float xa = __fsub_rn(txf, (float)tx1), ya = __fsub_rn(tyf, (float)ty1);
    const float xa1 = __fsub_rn(1.0f, xa), ya1 = __fsub_rn(1.0f, ya);
    const int tx2 = min(tx1 + 1, tiles - 1), ty2 = min(ty1 + 1, tiles - 1);
    tx1 = max(tx1, 0); ty1 = max(ty1, 0);
    const uint8_t* L = luts + (size_t)b * tiles * tiles * 256;
    const float l11 = L[(ty1 * tiles + tx1) * 256 + v], l12 = L[(ty1 * tiles + tx2) * 256 + v];
    const float l21 = L[(ty2 * tiles + tx1) * 256 + v], l22 = L[(ty2 * tiles + tx2) * 256 + v];
    const float top = __fmul_rn(__fadd_rn(__fmul_rn(l11, xa1), __fmul_rn(l12, xa)), ya1);
    const float bot = __fmul_rn(__fadd_rn(__fmul_rn(l21, xa1), __fmul_rn(l22, xa)), ya);
    const int lp = min(max(__float2int_rn(__fadd_rn(top, bot)), 0), 255);
    for (int co = 0; co < Cout; ++co) {
      int w = post ? post[lp * 3 + co] : lp;
      if (s.noise_sigma > 0.f) {
        const float f = (float)w + s.noise_sigma * philox_normal(s.noise_seed, ((uint64_t)co * T + y) * T + x);
        w = (int)fminf(fmaxf(f, 0.f), 255.f);
      }
      const size_t o = (((size_t)b * Cout + co) * T + y) * T + x;
      if (out_u8) out_u8[o] = (uint8_t)w;
      const float mean = co == 0 ? m0 : co == 1 ? m1 : m2, rstd = co == 0 ? r0 : co == 1 ? r1 : r2;
      out[o] = __fmul_rn(__fsub_rn(__fdiv_rn((float)w, 255.0f), mean), rstd);
    }
  }
}

}  // namespace

extern "C" int pm_augment_batch_u8_f32(const uint8_t* src, const pm_aug_sample_t* samples, const int32_t* tables, int B, int R, int T,
                                       int Cout, const float* mean, const float* rstd, float* out, uint8_t* out_u8, pm_stream_t s) {
  PM_CHECK_ARG(src && samples && tables && mean && rstd && (out || out_u8) && B >= 1 && R >= T && T >= 1 && Cout >= 1 && Cout <= 3);
  const size_t total = (size_t)B * T * T;
  augment_kernel<<<pm_grid(total, 256, 1, 16), 256, 0, S(s)>>>(src, samples, tables, B, R, T, Cout, mean[0], mean[Cout > 1 ? 1 : 0],
                                                               mean[Cout > 2 ? 2 : 0], rstd[0], rstd[Cout > 1 ? 1 : 0],
                                                               rstd[Cout > 2 ? 2 : 0], out, out_u8);
  PM_LAUNCH_OK();
}

extern "C" int pm_clahe_luts_u8(const uint8_t* img, const uint8_t* pre, int B, int T, int tiles, float clip_limit, uint8_t* luts,
                                pm_stream_t s) {
  PM_CHECK_ARG(img && luts && B >= 1 && B <= 65535 && tiles >= 1 && T >= tiles && T % tiles == 0 && clip_limit >= 0.f);
  const int area = (T / tiles) * (T / tiles);
  int clip = 0;
  if (clip_limit > 0.f) {   // CLAHE_Impl::apply: static_cast<int>(clipLimit_ * tileSizeTotal / histSize), at least 1 (double arithmetic)
    clip = (int)((double)clip_limit * area / 256);
    if (clip < 1) clip = 1;
  }
  const float lut_scale = 255.0f / (float)area;   // static_cast<float>(histSize - 1) / tileSizeTotal
  clahe_lut_kernel<<<dim3(tiles * tiles, B), 256, 0, S(s)>>>(img, pre, T, tiles, clip, lut_scale, luts);
  PM_LAUNCH_OK();
}

extern "C" int pm_augment_clahe_finish_f32(const uint8_t* img, const uint8_t* pre, const uint8_t* luts, const uint8_t* post,
                                           const pm_aug_sample_t* samples, int B, int T, int tiles, int Cout, const float* mean,
                                           const float* rstd, float* out, uint8_t* out_u8, pm_stream_t s) {
  PM_CHECK_ARG(img && luts && samples && mean && rstd && out && B >= 1 && tiles >= 1 && T % tiles == 0 && Cout >= 1 && Cout <= 3 &&
               (Cout == 1 || post));
  const size_t total = (size_t)B * T * T;
  clahe_finish_kernel<<<pm_grid(total, 256, 1, 16), 256, 0, S(s)>>>(img, pre, luts, post, samples, B, T, tiles, Cout, mean[0],
                                                                    mean[Cout > 1 ? 1 : 0], mean[Cout > 2 ? 2 : 0], rstd[0],
                                                                    rstd[Cout > 1 ? 1 : 0], rstd[Cout > 2 ? 2 : 0], out, out_u8);
  PM_LAUNCH_OK();
}
