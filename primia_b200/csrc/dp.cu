// Path T, DP-SGD (train.py:304-334 attaches pytorch-dp's PrivacyEngine: per-sample clipping to max_grad_norm, Gaussian noise
// noise_multiplier * max_grad_norm on the summed gradient, division by the batch size).  The per-sample gradients are
// MATERIALISED per layer ([B][n] fp32: 5.7 GB for ResNet-18 at B = 128 -- a small fraction of 180 GB of HBM) by the
// per-sample weight-gradient kernels (conv_tma.cu: wgrad_tma_kernel<.., PS>), then
//   dp_sqnorm_kernel        norm2[b] += sum_j g[b][j]^2                              (one pass, HBM bound)
//   dp_fc_sqnorm_kernel     the Linear layer's share without materialising it:      |dlogits_b|^2 * (|feat_b|^2 + 1)
//   dp_clip_factor_kernel   c[b] = min(1, C / (scale * sqrt(norm2[b]) + 1e-6))
//   dp_weighted_sum_kernel  out[j] = sum_b c[b] * g[b][j]                            (second pass, HBM bound)
//   dp_noise_kernel         out[j] += std * N(0,1)  (Philox4x32-10 + Box-Muller), or an explicit noise tensor (tests)
//   bn_eval_bwd / bn_persample_pgrad   BatchNorm as a frozen per-channel affine map (DP needs per-sample gradients, which batch
//                           statistics do not have: the reference refuses BatchNorm models under DP, train.py:306-310)
#include "common.cuh"

namespace {

typedef __nv_bfloat16 bf16;

__global__ void dp_sqnorm_kernel(const float* __restrict__ g, size_t n, double* __restrict__ norm2) {
  // grid (chunks, B): every block reduces one chunk of one sample's gradient
  const float* row = g + (size_t)blockIdx.y * n;
  double acc = 0.0;
  const size_t n4 = n / 4;
  if ((reinterpret_cast<uintptr_t>(row) & 15) == 0) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
      const float4 v = reinterpret_cast<const float4*>(row)[i];
      acc += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
    }
    for (size_t i = n4 * 4 + blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
      acc += (double)row[i] * row[i];
  } else {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
      acc += (double)row[i] * row[i];
  }
  __shared__ double sh[8];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    double t = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.0;
    t = warp_sum(t);
    if (threadIdx.x == 0 && t != 0.0) atomicAdd(norm2 + blockIdx.y, t);
  }
}

// Linear(F, ncls): grad_sample W_b = dlogits_b (x) feat_b, bias_b = dlogits_b  ->  |.|^2 = |dlogits_b|^2 (|feat_b|^2 + 1)
__global__ void dp_fc_sqnorm_kernel(const float* __restrict__ dlogits, int ld, float dl_scale, const float* __restrict__ feat, int F,
                                    int ncls, double* __restrict__ norm2) {
  const int b = blockIdx.x;
  double f2 = 0.0;
  for (int i = threadIdx.x; i < F; i += blockDim.x) f2 += (double)feat[(size_t)b * F + i] * feat[(size_t)b * F + i];
  __shared__ double sh[8];
  f2 = warp_sum(f2);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = f2;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (blockDim.x >> 5); ++w) t += sh[w];
    double d2 = 0.0;
    for (int c = 0; c < ncls; ++c) {
      const double d = (double)dlogits[(size_t)b * ld + c] * dl_scale;
      d2 += d * d;
    }
    norm2[b] += d2 * (t + 1.0);
  }
}

__global__ void dp_clip_factor_kernel(const double* __restrict__ norm2, int B, double scale, double max_norm, float* __restrict__ c,
                                      float* __restrict__ norms_out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const double nrm = scale * sqrt(norm2[b]);
  double f = max_norm / (nrm + 1e-6);   // ConstantFlatClipper: flat_value / (norm + 1e-6), clamped to 1
  if (f > 1.0) f = 1.0;
  c[b] = (float)f;
  if (norms_out) norms_out[b] = (float)nrm;
}

__global__ void dp_weighted_sum_kernel(const float* __restrict__ g, const float* __restrict__ c, int B, size_t n, float* __restrict__ out,
                                       int accumulate) {
  for (size_t j = blockIdx.x * (size_t)blockDim.x + threadIdx.x; j < n; j += (size_t)gridDim.x * blockDim.x) {
    float acc = accumulate ? out[j] : 0.f;
    int b = 0;
    for (; b + 8 <= B; b += 8) {   // 8 independent loads in flight; the accumulation order over b stays fixed (deterministic)
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = g[(size_t)(b + u) * n + j];
#pragma unroll
      for (int u = 0; u < 8; ++u) acc = fmaf(c[b + u], v[u], acc);
    }
    for (; b < B; ++b) acc = fmaf(c[b], g[(size_t)b * n + j], acc);
    out[j] = acc;
  }
}

// dW[c][f] = sum_b c[b] dlogits[b][c] feat[b][f], db[c] = sum_b c[b] dlogits[b][c]
__global__ void dp_fc_weighted_kernel(const float* __restrict__ dlogits, int ld, float dl_scale, const float* __restrict__ feat,
                                      const float* __restrict__ c, int B, int F, int ncls, float* __restrict__ dW, float* __restrict__ db) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < ncls * F) {
    const int cls = i / F, f = i - cls * F;
    float acc = 0.f;
    for (int b = 0; b < B; ++b) acc = fmaf(c[b] * dl_scale * dlogits[(size_t)b * ld + cls], feat[(size_t)b * F + f], acc);
    dW[i] = acc;
  } else if (i < ncls * F + ncls) {
    const int cls = i - ncls * F;
    float acc = 0.f;
    for (int b = 0; b < B; ++b) acc = fmaf(c[b] * dl_scale, dlogits[(size_t)b * ld + cls], acc);
    db[cls] = acc;
  }
}

// Philox4x32-10 (same generator as ring.cu's shares), two Box-Muller pairs per counter
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
__global__ void dp_counter_bump_kernel(unsigned long long* c) { *c += 1ull; }
__global__ void dp_noise_kernel(float* __restrict__ out, size_t n, float stddev, uint64_t seed, uint64_t offset,
                                const unsigned long long* __restrict__ counter_dev) {
  if (counter_dev) offset += *counter_dev;   // device-side step counter: a captured CUDA graph draws fresh noise at every replay
  for (size_t q = blockIdx.x * (size_t)blockDim.x + threadIdx.x; q * 4 < n; q += (size_t)gridDim.x * blockDim.x) {
    uint32_t c[4] = {(uint32_t)q, (uint32_t)(q >> 32), (uint32_t)offset, (uint32_t)(offset >> 32)};
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      philox_round(c, k0, k1);
      k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    float z[4];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float u1 = ((float)c[2 * h] + 1.0f) * 2.3283064365386963e-10f;   // (0, 1]
      const float u2 = (float)c[2 * h + 1] * 2.3283064365386963e-10f;        // [0, 1)
      const float rad = sqrtf(-2.0f * __logf(u1));
      float sn, cs;
      __sincosf(6.283185307179586f * u2, &sn, &cs);
      z[2 * h] = rad * cs; z[2 * h + 1] = rad * sn;
    }
#pragma unroll
    for (int e = 0; e < 4; ++e)
      if (q * 4 + e < n) out[q * 4 + e] += stddev * z[e];
  }
}

__global__ void dp_axpy_kernel(float* __restrict__ out, const float* __restrict__ x, float a, float post, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = (out[i] + a * x[i]) * post;
}

// ---- BatchNorm as a frozen affine map y = (x - mean) * invstd * gamma + beta (eval statistics), backward:
//   g = dy * (y_out > 0 if masked) ; dx = g * gamma * invstd ; per-sample dgamma_b[c] = sum_pix g * xhat, dbeta_b[c] = sum_pix g
template <typename T>
__global__ void bn_eval_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ y_out, const float* __restrict__ gamma,
                                   const float* __restrict__ invstd, size_t total, int C, T* __restrict__ g_out, T* __restrict__ dx) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    float g = to_f<T>(dy[i]);
    if (y_out && !(to_f<T>(y_out[i]) > 0.f)) g = 0.f;
    if (g_out) g_out[i] = from_f<T>(g);
    dx[i] = from_f<T>(g * gamma[c] * invstd[c]);
  }
}

// grid (ceil(C/32), pixel chunks, B), block (32 channels, 8 pixel lanes)
template <typename T>
__global__ void bn_persample_pgrad_kernel(const T* __restrict__ g, const T* __restrict__ x, const float* __restrict__ mean,
                                          const float* __restrict__ invstd, int HW, int C, float* __restrict__ dgamma,
                                          float* __restrict__ dbeta) {
  const int c = blockIdx.x * 32 + threadIdx.x;
  const int b = blockIdx.z;
  float sg = 0.f, sb = 0.f;
  if (c < C) {
    const float m = mean[c], is = invstd[c];
    const size_t base = (size_t)b * HW * C + c;
    for (int p = blockIdx.y * blockDim.y + threadIdx.y; p < HW; p += gridDim.y * blockDim.y) {
      const float gv = to_f<T>(g[base + (size_t)p * C]);
      sb += gv;
      sg = fmaf(gv, (to_f<T>(x[base + (size_t)p * C]) - m) * is, sg);
    }
  }
  __shared__ float s1[8][33], s2[8][33];
  s1[threadIdx.y][threadIdx.x] = sg;
  s2[threadIdx.y][threadIdx.x] = sb;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    float a = 0.f, d = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) { a += s1[k][threadIdx.x]; d += s2[k][threadIdx.x]; }
    atomicAdd(dgamma + (size_t)b * C + c, a);
    atomicAdd(dbeta + (size_t)b * C + c, d);
  }
}

template <typename T>
int bn_eval_bwd_t(const T* dy, const T* y_out, const float* gamma, const float* invstd, size_t P, int C, T* g_out, T* dx, pm_stream_t s) {
  PM_CHECK_ARG(dy && gamma && invstd && dx && C > 0);
  const size_t total = P * (size_t)C;
  if (!total) return PM_OK;
  bn_eval_bwd_kernel<T><<<pm_grid(total, 256, 1, 16), 256, 0, S(s)>>>(dy, y_out, gamma, invstd, total, C, g_out, dx);
  PM_LAUNCH_OK();
}
template <typename T>
int bn_persample_t(const T* g, const T* x, const float* mean, const float* invstd, int B, int HW, int C, float* dgamma, float* dbeta,
                   pm_stream_t s) {
  PM_CHECK_ARG(g && x && mean && invstd && dgamma && dbeta && B > 0 && B <= 65535 && HW > 0 && C > 0);
  PM_CUDA(cudaMemsetAsync(dgamma, 0, (size_t)B * C * 4, S(s)));
  PM_CUDA(cudaMemsetAsync(dbeta, 0, (size_t)B * C * 4, S(s)));
  int chunks = (HW + 63) / 64;
  if (chunks > 32) chunks = 32;
  bn_persample_pgrad_kernel<T><<<dim3((C + 31) / 32, chunks, B), dim3(32, 8), 0, S(s)>>>(g, x, mean, invstd, HW, C, dgamma, dbeta);
  PM_LAUNCH_OK();
}

}  // namespace

extern "C" {

int pm_dp_sqnorm_f32(const float* g, int B, size_t n, double* norm2, pm_stream_t s) {
  PM_CHECK_ARG(g && norm2 && B > 0 && B <= 65535);
  if (!n) return PM_OK;
  size_t chunks = (n + 256 * 16 - 1) / (256 * 16);
  const size_t cap = (size_t)pm_num_sms() * 8 / (size_t)B + 1;
  if (chunks > cap) chunks = cap;
  dp_sqnorm_kernel<<<dim3((unsigned)chunks, B), 256, 0, S(s)>>>(g, n, norm2);
  PM_LAUNCH_OK();
}
int pm_dp_fc_sqnorm_f32(const float* dlogits, int ld, float dl_scale, const float* feat, int B, int F, int ncls, double* norm2,
                        pm_stream_t s) {
  PM_CHECK_ARG(dlogits && feat && norm2 && B > 0 && ld >= ncls);
  dp_fc_sqnorm_kernel<<<B, 128, 0, S(s)>>>(dlogits, ld, dl_scale, feat, F, ncls, norm2);
  PM_LAUNCH_OK();
}
int pm_dp_clip_factors(const double* norm2, int B, double scale, double max_grad_norm, float* factors, float* norms_out, pm_stream_t s) {
  PM_CHECK_ARG(norm2 && factors && B > 0 && max_grad_norm > 0);
  dp_clip_factor_kernel<<<(B + 127) / 128, 128, 0, S(s)>>>(norm2, B, scale, max_grad_norm, factors, norms_out);
  PM_LAUNCH_OK();
}
int pm_dp_weighted_sum_f32(const float* g, const float* factors, int B, size_t n, float* out, int accumulate, pm_stream_t s) {
  PM_CHECK_ARG(g && factors && out && B > 0);
  if (!n) return PM_OK;
  dp_weighted_sum_kernel<<<pm_grid(n, 256, 1, 8), 256, 0, S(s)>>>(g, factors, B, n, out, accumulate);
  PM_LAUNCH_OK();
}
int pm_dp_fc_weighted_f32(const float* dlogits, int ld, float dl_scale, const float* feat, const float* factors, int B, int F, int ncls,
                          float* dW, float* db, pm_stream_t s) {
  PM_CHECK_ARG(dlogits && feat && factors && dW && db && ld >= ncls);
  const int n = ncls * F + ncls;
  dp_fc_weighted_kernel<<<(n + 127) / 128, 128, 0, S(s)>>>(dlogits, ld, dl_scale, feat, factors, B, F, ncls, dW, db);
  PM_LAUNCH_OK();
}
int pm_dp_add_noise_f32(float* g, size_t n, float stddev, uint64_t seed, uint64_t offset, uint64_t* counter_dev, pm_stream_t s) {
  PM_CHECK_ARG(g);
  if (!n || stddev == 0.f) return PM_OK;
  dp_noise_kernel<<<pm_grid((n + 3) / 4, 256), 256, 0, S(s)>>>(g, n, stddev, seed, offset, (const unsigned long long*)counter_dev);
  if (counter_dev) dp_counter_bump_kernel<<<1, 1, 0, S(s)>>>((unsigned long long*)counter_dev);
  PM_LAUNCH_OK();
}
int pm_dp_axpy_scale_f32(float* g, const float* x, float a, float post, size_t n, pm_stream_t s) {
  PM_CHECK_ARG(g && x);
  if (!n) return PM_OK;
  dp_axpy_kernel<<<pm_grid(n, 256), 256, 0, S(s)>>>(g, x, a, post, n);
  PM_LAUNCH_OK();
}
int pm_bn_eval_bwd_f32(const float* dy, const float* y_out, const float* gamma, const float* invstd, size_t P, int C, float* g_out,
                       float* dx, pm_stream_t s) {
  return bn_eval_bwd_t<float>(dy, y_out, gamma, invstd, P, C, g_out, dx, s);
}
int pm_bn_eval_bwd_bf16(const void* dy, const void* y_out, const float* gamma, const float* invstd, size_t P, int C, void* g_out,
                        void* dx, pm_stream_t s) {
  return bn_eval_bwd_t<bf16>((const bf16*)dy, (const bf16*)y_out, gamma, invstd, P, C, (bf16*)g_out, (bf16*)dx, s);
}
int pm_bn_persample_param_grads_f32(const float* g, const float* x, const float* mean, const float* invstd, int B, int HW, int C,
                                    float* dgamma, float* dbeta, pm_stream_t s) {
  return bn_persample_t<float>(g, x, mean, invstd, B, HW, C, dgamma, dbeta, s);
}
int pm_bn_persample_param_grads_bf16(const void* g, const void* x, const float* mean, const float* invstd, int B, int HW, int C,
                                     float* dgamma, float* dbeta, pm_stream_t s) {
  return bn_persample_t<bf16>((const bf16*)g, (const bf16*)x, mean, invstd, B, HW, C, dgamma, dbeta, s);
}

}  // extern "C"
