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
      out[o] = __fmul_rn(__fsub_rn(__fdiv_rn((float)v, 255.0f), mean), rstd);
    }
  }
}

}  // namespace

extern "C" int pm_augment_batch_u8_f32(const uint8_t* src, const pm_aug_sample_t* samples, const int32_t* tables, int B, int R, int T,
                                       int Cout, const float* mean, const float* rstd, float* out, uint8_t* out_u8, pm_stream_t s) {
  PM_CHECK_ARG(src && samples && tables && mean && rstd && out && B >= 1 && R >= T && T >= 1 && Cout >= 1 && Cout <= 3);
  const size_t total = (size_t)B * T * T;
  augment_kernel<<<pm_grid(total, 256, 1, 16), 256, 0, S(s)>>>(src, samples, tables, B, R, T, Cout, mean[0], mean[Cout > 1 ? 1 : 0],
                                                               mean[Cout > 2 ? 2 : 0], rstd[0], rstd[Cout > 1 ? 1 : 0],
                                                               rstd[Cout > 2 ? 2 : 0], out, out_u8);
  PM_LAUNCH_OK();
}
