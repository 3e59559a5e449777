// Path T: BatchNorm2d (training) forward/backward, ReLU/residual fusion, max-pool and global average pool
// on NHWC activations.  HBM-bound kernels: 16-byte vector accesses along the channel axis, per-thread
// double accumulation for the statistics (the CPU reference accumulates in double too), grids sized as a
// multiple of the SM count.  T = float (parity mode) or __nv_bfloat16 (throughput mode).
#include "common.cuh"
#include <stdlib.h>
#include <type_traits>

namespace {

template <typename T>
struct Vec;
template <>
struct Vec<float> {
  static constexpr int N = 4;
  using raw = float4;
  __device__ static void load(const float* p, float v[4]) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  __device__ static void store(float* p, const float v[4]) { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
};
template <>
struct Vec<__nv_bfloat16> {
  static constexpr int N = 8;
  __device__ static void load(const __nv_bfloat16* p, float v[8]) {
    const uint4 t = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __bfloat1622float2(h[i]);
      v[2 * i] = f.x; v[2 * i + 1] = f.y;
    }
  }
  __device__ static void store(__nv_bfloat16* p, const float v[8]) {
    uint4 t;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = t;
  }
};

constexpr int BT = 256;
template <typename T>
bool chan_ok(int C);

// ---- per-channel sums: out[0..C) += sum f0 , out[C..2C) += sum f1   (f0,f1 produced by Functor per element)
// Each thread owns one 16-byte channel group (column v) and walks rows, U rows in flight; partial sums are kept in
// double in fp32 (parity) mode -- ATen's CPU kernels accumulate in double -- and in float in bf16 mode.
template <typename T, typename F>
__global__ void __launch_bounds__(BT) chan_reduce_kernel(size_t P, int C, double* __restrict__ out, F f) {
  constexpr int V = Vec<T>::N;
  constexpr int U = 4;
  using acc_t = typename std::conditional<sizeof(T) == 4, double, float>::type;
  const int vr = C / V;                 // vectors per row (divides BT)
  const int v = threadIdx.x % vr;
  const int rpb = BT / vr;              // rows per block pass
  const int r0 = threadIdx.x / vr;
  acc_t s0[V], s1[V];
#pragma unroll
  for (int i = 0; i < V; ++i) s0[i] = s1[i] = 0;
  const size_t stride = (size_t)gridDim.x * rpb;
  for (size_t row = (size_t)blockIdx.x * rpb + r0; row < P; row += stride * U) {
    float a[U][V], b[U][V];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t r = row + u * stride;
      if (r < P) f(r, v * V, a[u], b[u]);
      else {
#pragma unroll
        for (int i = 0; i < V; ++i) a[u][i] = b[u][i] = 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int i = 0; i < V; ++i) { s0[i] += (acc_t)a[u][i]; s1[i] += (acc_t)b[u][i]; }
  }
  extern __shared__ double dyn[];       // [2][BT*V]
  double* d0 = dyn;
  double* d1 = dyn + BT * V;
#pragma unroll
  for (int i = 0; i < V; ++i) { d0[threadIdx.x * V + i] = (double)s0[i]; d1[threadIdx.x * V + i] = (double)s1[i]; }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += BT) {
    const int vv = c / V, ii = c % V;
    double t0 = 0.0, t1 = 0.0;
    for (int r = 0; r < rpb; ++r) { t0 += d0[(r * vr + vv) * V + ii]; t1 += d1[(r * vr + vv) * V + ii]; }
    atomicAdd(&out[c], t0);
    atomicAdd(&out[C + c], t1);
  }
}

template <typename T>
struct StatsF {
  const T* x; int C;
  __device__ void operator()(size_t row, int c, float* a, float* b) const {
    Vec<T>::load(x + row * C + c, a);
#pragma unroll
    for (int i = 0; i < Vec<T>::N; ++i) b[i] = a[i] * a[i];
  }
};

// g = dy * (y_out > 0) ; a = g ; b = g * xhat ; optionally writes g
template <typename T>
struct BwdF {
  const T* dy; const T* y_out; const T* x; const float* mean; const float* invstd; T* g_out; int C;
  __device__ void operator()(size_t row, int c, float* a, float* b) const {
    constexpr int V = Vec<T>::N;
    float xv[V];
    Vec<T>::load(dy + row * C + c, a);
    Vec<T>::load(x + row * C + c, xv);
    if (y_out) {
      float yv[V];
      Vec<T>::load(y_out + row * C + c, yv);
#pragma unroll
      for (int i = 0; i < V; ++i) a[i] = yv[i] > 0.f ? a[i] : 0.f;
    }
    if (g_out) Vec<T>::store(g_out + row * C + c, a);
#pragma unroll
    for (int i = 0; i < V; ++i) b[i] = a[i] * ((xv[i] - mean[c + i]) * invstd[c + i]);
  }
};

__global__ void bn_finalize_kernel(const double* __restrict__ stats, double P, int C, float eps, float momentum,
                                   float* __restrict__ mean, float* __restrict__ invstd, float* __restrict__ rm,
                                   float* __restrict__ rv) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double m = stats[c] / P;
  double var = stats[C + c] / P - m * m;
  if (var < 0.0) var = 0.0;
  mean[c] = (float)m;
  invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
  if (rm) {
    const double unbiased = P > 1.0 ? var * P / (P - 1.0) : var;
    rm[c] = (float)((1.0 - momentum) * (double)rm[c] + momentum * m);
    rv[c] = (float)((1.0 - momentum) * (double)rv[c] + momentum * unbiased);
  }
}

// FUSED = true: mean / invstd are derived from the batch statistics here (replaces bn_finalize + bn_apply).
// The per-channel constants are computed once per block by C threads into shared memory (mu, scale, shift), so the
// per-thread prologue is a few LDS instead of 4*V global loads and V double-precision evaluations.
template <typename T, bool FUSED>
__global__ void __launch_bounds__(BT, 3)
bn_apply_kernel(const T* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ invstd,
                const double* __restrict__ stats, double Pd, float eps, float momentum, float* __restrict__ mean_out,
                float* __restrict__ invstd_out, float* __restrict__ rm, float* __restrict__ rv,
                const float* __restrict__ gamma, const float* __restrict__ beta, const T* __restrict__ res, int relu,
                size_t P, int C, T* __restrict__ y) {
  constexpr int V = Vec<T>::N;
  constexpr int U = 2;
  extern __shared__ float sp[];  // [4][C]: mu, invstd (or invstd*gamma), gamma, beta
  const double invPd = FUSED ? 1.0 / Pd : 0.0;
  for (int ch = threadIdx.x; ch < C; ch += BT) {
    float m, is;
    if (FUSED) {
      // no double division / sqrt (software sequences): one double multiply each for mean and E[x^2]-mean^2,
      // invstd = rsqrt in float + one Newton step (<= 1 ulp)
      const double md = stats[ch] * invPd;
      double var = stats[C + ch] * invPd - md * md;
      if (var < 0.0) var = 0.0;
      m = (float)md;
      const float ve = (float)(var + (double)eps);
      is = rsqrtf(ve);
      is = is * (1.5f - 0.5f * ve * is * is);
      if (blockIdx.x == 0) {
        mean_out[ch] = m;
        invstd_out[ch] = is;
        if (rm) {
          const double unbiased = Pd > 1.0 ? var * Pd / (Pd - 1.0) : var;
          rm[ch] = (float)((1.0 - momentum) * (double)rm[ch] + momentum * md);
          rv[ch] = (float)((1.0 - momentum) * (double)rv[ch] + momentum * unbiased);
        }
      }
    } else {
      m = mean[ch];
      is = invstd[ch];
    }
    const float ga = gamma[ch];
    sp[ch] = m;
    sp[C + ch] = sizeof(T) == 4 ? is : is * ga;  // parity mode keeps the (x-mean)*invstd*gamma+beta evaluation order
    sp[2 * C + ch] = ga;
    sp[3 * C + ch] = beta[ch];
  }
  __syncthreads();
  const int vr = C / V, v = threadIdx.x % vr, rpb = BT / vr, r0 = threadIdx.x / vr;
  const int c = v * V;
  float mu[V], sc[V], ga[V], be[V];
#pragma unroll
  for (int k = 0; k < V; ++k) { mu[k] = sp[c + k]; sc[k] = sp[C + c + k]; ga[k] = sp[2 * C + c + k]; be[k] = sp[3 * C + c + k]; }
  const size_t stride = (size_t)gridDim.x * rpb;
  for (size_t row = (size_t)blockIdx.x * rpb + r0; row < P; row += stride * U) {
    float xv[U][V], rvv[U][V];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t r = row + u * stride;
      if (r < P) {
        Vec<T>::load(x + r * C + c, xv[u]);
        if (res) Vec<T>::load(res + r * C + c, rvv[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t r = row + u * stride;
      if (r < P) {
#pragma unroll
        for (int k = 0; k < V; ++k) {
          float t;
          if (sizeof(T) == 4) t = (xv[u][k] - mu[k]) * sc[k] * ga[k] + be[k];
          else t = (xv[u][k] - mu[k]) * sc[k] + be[k];
          if (res) t += rvv[u][k];
          if (relu) t = t > 0.f ? t : 0.f;
          xv[u][k] = t;
        }
        Vec<T>::store(y + r * C + c, xv[u]);
      }
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(BT, sizeof(T) == 4 ? 1 : 3)
bn_bwd_apply_kernel(const T* __restrict__ dy, const T* __restrict__ y_out, const T* __restrict__ x,
                    const float* __restrict__ mean, const float* __restrict__ invstd, const float* __restrict__ gamma,
                    const double* __restrict__ sums, double invP, size_t P, int C, T* __restrict__ dx,
                    float* __restrict__ dgamma, float* __restrict__ dbeta) {
  constexpr int V = Vec<T>::N;
  constexpr int U = 2;
  using par_t = typename std::conditional<sizeof(T) == 4, double, float>::type;
  const int vr = C / V, v = threadIdx.x % vr, rpb = BT / vr, r0 = threadIdx.x / vr;
  const int c = v * V;
  float mu[V], is[V], gi[V];  // gi = gamma * invstd
  par_t mg[V], mgx[V];
#pragma unroll
  for (int k = 0; k < V; ++k) {
    mu[k] = mean[c + k]; is[k] = invstd[c + k]; gi[k] = gamma[c + k];
    mg[k] = (par_t)(sums[c + k] * invP); mgx[k] = (par_t)(sums[C + c + k] * invP);
    if (dgamma && blockIdx.x == 0 && r0 == 0) {
      dbeta[c + k] = (float)sums[c + k];
      dgamma[c + k] = (float)sums[C + c + k];
    }
    if (sizeof(T) != 4) gi[k] *= is[k];
  }
  const size_t stride = (size_t)gridDim.x * rpb;
  for (size_t row = (size_t)blockIdx.x * rpb + r0; row < P; row += stride * U) {
    float g[U][V], xv[U][V], yv[U][V];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t r = row + u * stride;
      if (r < P) {
        Vec<T>::load(dy + r * C + c, g[u]);
        Vec<T>::load(x + r * C + c, xv[u]);
        if (y_out) Vec<T>::load(y_out + r * C + c, yv[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t r = row + u * stride;
      if (r < P) {
#pragma unroll
        for (int k = 0; k < V; ++k) {
          float gg = g[u][k];
          if (y_out) gg = yv[u][k] > 0.f ? gg : 0.f;
          if (sizeof(T) == 4) {
            // parity mode: ATen's CPU kernel evaluates dy - mean(dy) - xhat*mean(dy*xhat) in double (acc_type<float>); the
            // per-channel common mode of dy can exceed its fluctuation by 1e3-1e4, so fp32 here would cost 1e-4 relative.
            const double isd = (double)is[k];
            const double xhat = ((double)xv[u][k] - (double)mu[k]) * isd;
            g[u][k] = (float)((double)gi[k] * isd * ((double)gg - (double)mg[k] - xhat * (double)mgx[k]));
          } else {
            const float xhat = (xv[u][k] - mu[k]) * is[k];
            g[u][k] = gi[k] * (gg - (float)mg[k] - xhat * (float)mgx[k]);
          }
        }
        Vec<T>::store(dx + r * C + c, g[u]);
      }
    }
  }
}

// ---- BatchNorm backward in ONE launch: reduce -> grid barrier -> distributed total -> grid barrier -> apply.
// All blocks are co-resident (grid <= 2 blocks/SM), so a sense-reversing barrier on two words of device memory is safe,
// needs no host-side epoch and survives CUDA-graph replay.  No floating-point atomics: per-block partial sums are
// combined in a fixed order (deterministic).  For the mid-size layers the second pass over dy / x / y_out hits L2.
__device__ __forceinline__ void grid_barrier(unsigned* count, volatile unsigned* gen) {
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned my_gen = *gen;
    __threadfence();
    if (atomicAdd(count, 1u) == gridDim.x - 1) {
      *count = 0;
      __threadfence();
      atomicAdd((unsigned*)gen, 1u);
    } else {
      while (*gen == my_gen) { }
    }
    __threadfence();
  }
  __syncthreads();
}

// gradient of MaxPool2d(3,2,1) gathered on the fly: element (b,h,w,c..c+V) of the pool INPUT gradient from the pooled
// gradient dpool [B,Ho,Wo,C] and the stored argmax (one byte per pooled element) -- lets the stem's BN backward read the
// 4x smaller pooled tensors instead of a materialised full-resolution gradient.
struct PoolGeo { int H, W, Ho, Wo; };
template <typename T>
__device__ __forceinline__ void load_pool_grad(const T* __restrict__ dpool, const uint8_t* __restrict__ idx, PoolGeo pg, size_t r,
                                               int C, int c, float* g) {
  constexpr int V = Vec<T>::N;
  const int w = (int)(r % pg.W);
  const size_t t = r / pg.W;
  const int h = (int)(t % pg.H);
  const size_t b = t / pg.H;
#pragma unroll
  for (int k = 0; k < V; ++k) g[k] = 0.f;
  for (int oh = h / 2; oh <= (h + 1) / 2; ++oh) {
    if (oh >= pg.Ho) continue;
    const int rr = h - (oh * 2 - 1);
    if (rr < 0 || rr > 2) continue;
    for (int ow = w / 2; ow <= (w + 1) / 2; ++ow) {
      if (ow >= pg.Wo) continue;
      const int ss = w - (ow * 2 - 1);
      if (ss < 0 || ss > 2) continue;
      const size_t o = ((b * pg.Ho + oh) * pg.Wo + ow) * C + c;
      float d[V];
      Vec<T>::load(dpool + o, d);
      uint8_t id[V];
      if (V == 8) *reinterpret_cast<uint2*>(id) = *reinterpret_cast<const uint2*>(idx + o);
      else *reinterpret_cast<uint32_t*>(id) = *reinterpret_cast<const uint32_t*>(idx + o);
      const int want = rr * 3 + ss;
#pragma unroll
      for (int k = 0; k < V; ++k) g[k] += id[k] == want ? d[k] : 0.f;
    }
  }
}

// ATOMIC = true (bf16 / throughput mode): the per-block sums go straight into a ping-pong totals buffer with double
// atomics and only ONE grid barrier is needed (parity = barrier generation & 1; the buffer of the other parity is cleared
// for the next launch).  ATOMIC = false (fp32 / parity mode): per-block partials + fixed-order tree, two barriers,
// bit-reproducible.
template <typename T, bool POOL, bool ATOMIC>
__global__ void __launch_bounds__(BT, 2)
bn_bwd_fused_kernel(const T* __restrict__ dy, const uint8_t* __restrict__ pool_idx, PoolGeo pg, const T* __restrict__ y_out,
                    const T* __restrict__ x,
                    const float* __restrict__ mean, const float* __restrict__ invstd, const float* __restrict__ gamma,
                    const float* __restrict__ beta_mask /* != NULL: ReLU decision recomputed from x (no y_out read) */,
                    size_t P, int C, double invP, double* __restrict__ partial /* [grid][2C] */,
                    double* __restrict__ totals /* [2C] */, unsigned* __restrict__ sync /* [2] */, T* __restrict__ g_out,
                    T* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  constexpr int V = Vec<T>::N;
  constexpr int U = 2;
  using acc_t = typename std::conditional<sizeof(T) == 4, double, float>::type;
  extern __shared__ double dyn[];  // [2][BT*V]
  const int vr = C / V, v = threadIdx.x % vr, rpb = BT / vr, r0 = threadIdx.x / vr;
  const int c = v * V;
  // contiguous row range per block (keeps the second pass of mid-size layers inside L2 and DRAM pages sequential)
  const size_t rows_per_block = ((P + gridDim.x - 1) / gridDim.x + rpb - 1) / rpb * rpb;
  const size_t row_begin = (size_t)blockIdx.x * rows_per_block;
  const size_t row_end = row_begin + rows_per_block < P ? row_begin + rows_per_block : P;
  float mu[V], is[V], msc[V], mbe[V];
#pragma unroll
  for (int k = 0; k < V; ++k) {
    mu[k] = mean[c + k]; is[k] = invstd[c + k];
    // the forward's bf16-mode evaluation (bn_apply_kernel): t = (x - mu) * (invstd * gamma) + beta ; y > 0 <=> t > 0
    msc[k] = beta_mask ? is[k] * gamma[c + k] : 0.f;
    mbe[k] = beta_mask ? beta_mask[c + k] : 0.f;
  }

  // ---- phase 1: per-block partial sums of g and g * xhat (g = dy masked by the ReLU decision)
  acc_t s0[V], s1[V];
#pragma unroll
  for (int k = 0; k < V; ++k) s0[k] = s1[k] = 0;
  for (size_t row = row_begin + r0; row < row_end; row += (size_t)rpb * U) {
    float g[U][V], xv[U][V], yv[U][V];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t r = row + (size_t)u * rpb;
      if (r < row_end) {
        if (POOL) load_pool_grad<T>(dy, pool_idx, pg, r, C, c, g[u]);
        else Vec<T>::load(dy + r * C + c, g[u]);
        Vec<T>::load(x + r * C + c, xv[u]);
        if (y_out) Vec<T>::load(y_out + r * C + c, yv[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t r = row + (size_t)u * rpb;
      if (r < row_end) {
#pragma unroll
        for (int k = 0; k < V; ++k) {
          float gg = g[u][k];
          if (y_out) gg = yv[u][k] > 0.f ? gg : 0.f;
          else if (beta_mask) gg = ((xv[u][k] - mu[k]) * msc[k] + mbe[k]) > 0.f ? gg : 0.f;
          g[u][k] = gg;
          s0[k] += (acc_t)gg;
          s1[k] += (acc_t)(gg * ((xv[u][k] - mu[k]) * is[k]));
        }
        if (g_out) Vec<T>::store(g_out + r * C + c, g[u]);
      }
    }
  }
  double* d0 = dyn;
  double* d1 = dyn + BT * V;
#pragma unroll
  for (int k = 0; k < V; ++k) { d0[threadIdx.x * V + k] = (double)s0[k]; d1[threadIdx.x * V + k] = (double)s1[k]; }
  __syncthreads();
  const unsigned parity = ATOMIC ? (*(volatile unsigned*)(sync + 1)) & 1u : 0u;  // generation before this launch's barrier
  constexpr int CMAX = 512;  // fixed ping-pong stride so that launches with different C share one workspace safely
  double* tot = ATOMIC ? totals + (size_t)parity * 2 * CMAX : totals;
  const int TS = ATOMIC ? CMAX : C;  // offset of the second statistic inside `tot`
  for (int ch = threadIdx.x; ch < C; ch += BT) {
    const int vv = ch / V, ii = ch % V;
    double t0 = 0.0, t1 = 0.0;
    for (int r = 0; r < rpb; ++r) { t0 += d0[(r * vr + vv) * V + ii]; t1 += d1[(r * vr + vv) * V + ii]; }
    if (ATOMIC) {
      atomicAdd(tot + ch, t0);
      atomicAdd(tot + TS + ch, t1);
    } else {
      partial[(size_t)blockIdx.x * 2 * C + ch] = t0;
      partial[(size_t)blockIdx.x * 2 * C + C + ch] = t1;
    }
  }
  if (ATOMIC && blockIdx.x == 0)  // clear the WHOLE other-parity buffer for the next launch (its readers finished a launch ago)
    for (int i = threadIdx.x; i < 2 * CMAX; i += BT) totals[(size_t)(parity ^ 1u) * 2 * CMAX + i] = 0.0;
  grid_barrier(sync, sync + 1);

  if (!ATOMIC) {
    // ---- phase 1b: the 2C totals are spread over all warps of the grid; each warp sums one column of the per-block
    // partials (lanes stride over blocks, fixed-order shuffle tree => deterministic)
    const int lane = threadIdx.x & 31, gw = blockIdx.x * (BT / 32) + (threadIdx.x >> 5), nw = gridDim.x * (BT / 32);
    for (int q = gw; q < 2 * C; q += nw) {
      double t = 0.0;
      for (unsigned b = lane; b < gridDim.x; b += 32) t += __ldcg(partial + (size_t)b * 2 * C + q);
      t = warp_sum(t);
      if (lane == 0) {
        totals[q] = t;
        if (dgamma) {
          if (q < C) dbeta[q] = (float)t;
          else dgamma[q - C] = (float)t;
        }
      }
    }
    grid_barrier(sync, sync + 1);
  } else if (dgamma && blockIdx.x == 0) {
    for (int ch = threadIdx.x; ch < C; ch += BT) {
      dbeta[ch] = (float)__ldcg(tot + ch);
      dgamma[ch] = (float)__ldcg(tot + TS + ch);
    }
  }

  // ---- phase 2: dx = gamma * invstd * (g - mean(g) - xhat * mean(g * xhat)); per-channel constants via shared memory
  using par_t = typename std::conditional<sizeof(T) == 4, double, float>::type;
  double* sm_mg = dyn;        // [C]
  double* sm_mgx = dyn + C;   // [C]
  __syncthreads();
  for (int ch = threadIdx.x; ch < C; ch += BT) {
    sm_mg[ch] = __ldcg(tot + ch) * invP;
    sm_mgx[ch] = __ldcg(tot + TS + ch) * invP;
  }
  __syncthreads();
  par_t mg[V], mgx[V];
  float gi[V];
#pragma unroll
  for (int k = 0; k < V; ++k) {
    mg[k] = (par_t)sm_mg[c + k];
    mgx[k] = (par_t)sm_mgx[c + k];
    gi[k] = gamma[c + k];
    if (sizeof(T) != 4) gi[k] *= is[k];
  }
  const T* gsrc = g_out ? g_out : dy;   // the masked gradient was materialised in phase 1 when g_out != NULL
  const T* ymask = g_out ? nullptr : y_out;
  const bool xmask = beta_mask != nullptr && g_out == nullptr && y_out == nullptr;
  for (size_t row = row_begin + r0; row < row_end; row += (size_t)rpb * U) {
    float g[U][V], xv[U][V], yv[U][V];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t r = row + (size_t)u * rpb;
      if (r < row_end) {
        // POOL: the gathered gradient is either re-read from g_out (materialised in phase 1) or gathered again
        if (POOL && !g_out) load_pool_grad<T>(dy, pool_idx, pg, r, C, c, g[u]);
        else Vec<T>::load(gsrc + r * C + c, g[u]);
        Vec<T>::load(x + r * C + c, xv[u]);
        if (ymask) Vec<T>::load(ymask + r * C + c, yv[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t r = row + (size_t)u * rpb;
      if (r < row_end) {
#pragma unroll
        for (int k = 0; k < V; ++k) {
          float gg = g[u][k];
          if (ymask) gg = yv[u][k] > 0.f ? gg : 0.f;
          else if (xmask) gg = ((xv[u][k] - mu[k]) * msc[k] + mbe[k]) > 0.f ? gg : 0.f;
          if (sizeof(T) == 4) {
            const double isd = (double)is[k];
            const double xhat = ((double)xv[u][k] - (double)mu[k]) * isd;
            g[u][k] = (float)((double)gi[k] * isd * ((double)gg - (double)mg[k] - xhat * (double)mgx[k]));
          } else {
            const float xhat = (xv[u][k] - mu[k]) * is[k];
            g[u][k] = gi[k] * (gg - (float)mg[k] - xhat * (float)mgx[k]);
          }
        }
        Vec<T>::store(dx + r * C + c, g[u]);
      }
    }
  }
}

// ---- bf16 (throughput-mode) BatchNorm backward, same one-launch protocol as the ATOMIC variant above (per-block sums ->
// double atomics into the ping-pong totals -> ONE grid barrier -> apply) but written for memory-level parallelism: four rows
// per thread are in flight as raw 16-byte vectors (the generic kernel keeps two rows as 8 floats each and sits on its
// 128-register cap), per-channel constants of the apply phase are folded into dx = A*g + B*x + C.
// MASK: 0 none (BN without ReLU: the downsample branch), 1 ReLU decision from y_out, 2 recomputed from x as the forward did.
__device__ __forceinline__ void bf8_unpack(const uint4& t, float (&v)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x; v[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ uint4 bf8_pack(const float (&v)[8]) {
  uint4 t;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  return t;
}

// bf16 BatchNorm forward apply (training; mean / invstd finalised from the batch statistics in the prologue exactly as
// bn_apply_kernel<T, true> does): y = act((x - mu) * (invstd * gamma) + beta (+ residual)), four raw 16-byte rows in flight.
__global__ void __launch_bounds__(BT, 3)
bn_fwd_bf16_kernel(const __nv_bfloat16* __restrict__ x, const double* __restrict__ stats, double Pd, float eps, float momentum,
                   float* __restrict__ mean_out, float* __restrict__ invstd_out, float* __restrict__ rm, float* __restrict__ rv,
                   const float* __restrict__ gamma, const float* __restrict__ beta, const __nv_bfloat16* __restrict__ res, int relu,
                   size_t P, int C, __nv_bfloat16* __restrict__ y) {
  constexpr int V = 8, U = 4;
  extern __shared__ float sp[];  // [3][C]: mu, invstd*gamma, beta
  pm_pdl_sync();
  const double invPd = 1.0 / Pd;
  for (int ch = threadIdx.x; ch < C; ch += BT) {
    const double md = stats[ch] * invPd;
    double var = stats[C + ch] * invPd - md * md;
    if (var < 0.0) var = 0.0;
    const float m = (float)md;
    const float ve = (float)(var + (double)eps);
    float is = rsqrtf(ve);
    is = is * (1.5f - 0.5f * ve * is * is);
    if (blockIdx.x == 0) {
      mean_out[ch] = m;
      invstd_out[ch] = is;
      if (rm) {
        const double unbiased = Pd > 1.0 ? var * Pd / (Pd - 1.0) : var;
        rm[ch] = (float)((1.0 - momentum) * (double)rm[ch] + momentum * md);
        rv[ch] = (float)((1.0 - momentum) * (double)rv[ch] + momentum * unbiased);
      }
    }
    sp[ch] = m;
    sp[C + ch] = is * gamma[ch];
    sp[2 * C + ch] = beta[ch];
  }
  __syncthreads();
  const int vr = C / V, v = threadIdx.x % vr, rpb = BT / vr, r0 = threadIdx.x / vr;
  const int c = v * V;
  float mu[V], sc[V], be[V];
#pragma unroll
  for (int k = 0; k < V; ++k) { mu[k] = sp[c + k]; sc[k] = sp[C + c + k]; be[k] = sp[2 * C + c + k]; }
  const size_t stride = (size_t)gridDim.x * rpb;
  for (size_t row = (size_t)blockIdx.x * rpb + r0; row < P; row += stride * U) {
    uint4 X[U], R[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t r = row + u * stride;
      if (r < P) {
        X[u] = *reinterpret_cast<const uint4*>(x + r * C + c);
        if (res) R[u] = *reinterpret_cast<const uint4*>(res + r * C + c);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t r = row + u * stride;
      if (r < P) {
        float xv[V], rvv[V];
        bf8_unpack(X[u], xv);
        if (res) bf8_unpack(R[u], rvv);
#pragma unroll
        for (int k = 0; k < V; ++k) {
          float t = (xv[k] - mu[k]) * sc[k] + be[k];
          if (res) t += rvv[k];
          if (relu) t = t > 0.f ? t : 0.f;
          xv[k] = t;
        }
        *reinterpret_cast<uint4*>(y + r * C + c) = bf8_pack(xv);
      }
    }
  }
}

template <int MASK, bool GOUT>
__global__ void __launch_bounds__(BT, 2)
bn_bwd_bf16_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ y_out, const __nv_bfloat16* __restrict__ x,
                   const float* __restrict__ mean, const float* __restrict__ invstd, const float* __restrict__ gamma,
                   const float* __restrict__ beta, size_t P, int C, double invP, double* __restrict__ totals,
                   unsigned* __restrict__ sync, __nv_bfloat16* __restrict__ g_out, __nv_bfloat16* __restrict__ dx,
                   float* __restrict__ dgamma, float* __restrict__ dbeta) {
  constexpr int V = 8, U = 4, CMAX = 512;
  extern __shared__ double dyn[];  // phase 1: float [2][BT*V] block reduction ; phase 2: float [3][C] constants
  float* red = reinterpret_cast<float*>(dyn);
  pm_pdl_sync();
  const int vr = C / V, v = threadIdx.x % vr, rpb = BT / vr, r0 = threadIdx.x / vr;
  const int c = v * V;
  const size_t rows_per_block = ((P + gridDim.x - 1) / gridDim.x + rpb - 1) / rpb * rpb;
  const size_t row_begin = (size_t)blockIdx.x * rows_per_block;
  const size_t row_end = row_begin + rows_per_block < P ? row_begin + rows_per_block : P;
  float mu[V], is[V], msc[V], mbe[V];
#pragma unroll
  for (int k = 0; k < V; ++k) {
    mu[k] = mean[c + k]; is[k] = invstd[c + k];
    msc[k] = MASK == 2 ? is[k] * gamma[c + k] : 0.f;  // the forward: t = (x - mu) * (invstd * gamma) + beta ; y > 0 <=> t > 0
    mbe[k] = MASK == 2 ? beta[c + k] : 0.f;
  }
  // ---- phase 1
  float s0[V], s1[V];
#pragma unroll
  for (int k = 0; k < V; ++k) s0[k] = s1[k] = 0.f;
  for (size_t row = row_begin + r0; row < row_end; row += (size_t)rpb * U) {
    uint4 G[U], X[U], Y[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t r = row + (size_t)u * rpb;
      if (r < row_end) {
        G[u] = *reinterpret_cast<const uint4*>(dy + r * C + c);
        X[u] = *reinterpret_cast<const uint4*>(x + r * C + c);
        if (MASK == 1) Y[u] = *reinterpret_cast<const uint4*>(y_out + r * C + c);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t r = row + (size_t)u * rpb;
      if (r < row_end) {
        float g[V], xv[V], yv[V];
        bf8_unpack(G[u], g);
        bf8_unpack(X[u], xv);
        if (MASK == 1) bf8_unpack(Y[u], yv);
#pragma unroll
        for (int k = 0; k < V; ++k) {
          float gg = g[k];
          if (MASK == 1) gg = yv[k] > 0.f ? gg : 0.f;
          if (MASK == 2) gg = ((xv[k] - mu[k]) * msc[k] + mbe[k]) > 0.f ? gg : 0.f;
          g[k] = gg;
          s0[k] += gg;
          s1[k] = fmaf(gg, (xv[k] - mu[k]) * is[k], s1[k]);
        }
        if (GOUT) *reinterpret_cast<uint4*>(g_out + r * C + c) = bf8_pack(g);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < V; ++k) { red[threadIdx.x * V + k] = s0[k]; red[BT * V + threadIdx.x * V + k] = s1[k]; }
  __syncthreads();
  const unsigned parity = (*(volatile unsigned*)(sync + 1)) & 1u;  // barrier generation before this launch's barrier
  double* tot = totals + (size_t)parity * 2 * CMAX;
  for (int ch = threadIdx.x; ch < C; ch += BT) {
    const int vv = ch / V, ii = ch % V;
    double t0 = 0.0, t1 = 0.0;
    for (int r = 0; r < rpb; ++r) { t0 += (double)red[(r * vr + vv) * V + ii]; t1 += (double)red[BT * V + (r * vr + vv) * V + ii]; }
    atomicAdd(tot + ch, t0);
    atomicAdd(tot + CMAX + ch, t1);
  }
  if (blockIdx.x == 0)  // clear the other-parity buffer for the next launch (its readers finished a launch ago)
    for (int i = threadIdx.x; i < 2 * CMAX; i += BT) totals[(size_t)(parity ^ 1u) * 2 * CMAX + i] = 0.0;
  grid_barrier(sync, sync + 1);
  if (dgamma && blockIdx.x == 0) {
    for (int ch = threadIdx.x; ch < C; ch += BT) {
      dbeta[ch] = (float)__ldcg(tot + ch);
      dgamma[ch] = (float)__ldcg(tot + CMAX + ch);
    }
  }
  // ---- phase 2: dx = gamma*invstd * (g - mean(g) - xhat*mean(g*xhat)) = A*g + B*x + Cc per channel
  float* cA = red;
  float* cB = red + C;
  float* cC = red + 2 * C;
  __syncthreads();
  for (int ch = threadIdx.x; ch < C; ch += BT) {
    const float mg = (float)(__ldcg(tot + ch) * invP), mgx = (float)(__ldcg(tot + CMAX + ch) * invP);
    const float isc = invstd[ch], a = gamma[ch] * isc;
    cA[ch] = a;
    cB[ch] = -a * mgx * isc;
    cC[ch] = a * (mgx * isc * mean[ch] - mg);
  }
  __syncthreads();
  float A[V], Bc[V], Cc[V];
#pragma unroll
  for (int k = 0; k < V; ++k) { A[k] = cA[c + k]; Bc[k] = cB[c + k]; Cc[k] = cC[c + k]; }
  const __nv_bfloat16* gsrc = GOUT ? g_out : dy;  // the masked gradient was materialised in phase 1 when GOUT
  for (size_t row = row_begin + r0; row < row_end; row += (size_t)rpb * U) {
    uint4 G[U], X[U], Y[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t r = row + (size_t)u * rpb;
      if (r < row_end) {
        G[u] = *reinterpret_cast<const uint4*>(gsrc + r * C + c);
        X[u] = *reinterpret_cast<const uint4*>(x + r * C + c);
        if (MASK == 1 && !GOUT) Y[u] = *reinterpret_cast<const uint4*>(y_out + r * C + c);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t r = row + (size_t)u * rpb;
      if (r < row_end) {
        float g[V], xv[V], yv[V];
        bf8_unpack(G[u], g);
        bf8_unpack(X[u], xv);
        if (MASK == 1 && !GOUT) bf8_unpack(Y[u], yv);
#pragma unroll
        for (int k = 0; k < V; ++k) {
          float gg = g[k];
          if (MASK == 1 && !GOUT) gg = yv[k] > 0.f ? gg : 0.f;
          if (MASK == 2) gg = ((xv[k] - mu[k]) * msc[k] + mbe[k]) > 0.f ? gg : 0.f;
          g[k] = fmaf(A[k], gg, fmaf(Bc[k], xv[k], Cc[k]));
        }
        *reinterpret_cast<uint4*>(dx + r * C + c) = bf8_pack(g);
      }
    }
  }
}

static bool bn_bwd_generic() {  // PRIMIA_BN_BWD_GENERIC=1: the generic two-rows-in-flight kernel also in bf16 mode (cross-check)
  const char* e = getenv("PRIMIA_BN_BWD_GENERIC");
  return e && e[0] == '1';
}

template <typename T>
int bn_bwd_fused_t(const T* dy, const T* y_out, const T* x, const float* mean, const float* invstd, const float* gamma, size_t P,
                   int C, double* ws, T* g_out, T* dx, float* dgamma, float* dbeta, pm_stream_t s,
                   const uint8_t* pool_idx = nullptr, PoolGeo pg = PoolGeo{0, 0, 0, 0}, const float* beta_mask = nullptr) {
  PM_CHECK_ARG(dy && x && mean && invstd && gamma && ws && dx && P > 0 && chan_ok<T>(C) && ((dgamma == nullptr) == (dbeta == nullptr)));
  PM_CHECK_ARG(g_out != dx);
  const int rpb = BT / (C / Vec<T>::N);
  size_t blocks = (P + (size_t)rpb * 8 - 1) / ((size_t)rpb * 8);
  const size_t cap = (size_t)pm_num_sms() * 2;  // all blocks must be co-resident (grid barrier)
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  // ws layout: [2 x unsigned sync | pad to 16 B][2C totals][grid x 2C partials]
  unsigned* sync = reinterpret_cast<unsigned*>(ws);
  PM_CHECK_ARG(C <= 512);
  double* totals = ws + 2;          // [2][2*512] (ping-pong in ATOMIC mode)
  double* partial = totals + 4 * 512;
  const size_t smem = 2 * BT * Vec<T>::N * sizeof(double);
  PM_CHECK_ARG(!(beta_mask && y_out));  // the ReLU decision comes from y_out OR is recomputed from x, not both
  if constexpr (sizeof(T) == 2) {
    if (!pool_idx && !bn_bwd_generic()) {
      typedef __nv_bfloat16 b16;
      const b16* d = (const b16*)dy; const b16* yo = (const b16*)y_out; const b16* xx = (const b16*)x;
      b16* go = (b16*)g_out; b16* dxx = (b16*)dx;
      const size_t sm = 2 * BT * 8 * sizeof(float);
      const double invP = 1.0 / (double)P;
#define PM_BNB(MASK, GO) PM_CUDA(pm_launch(bn_bwd_bf16_kernel<MASK, GO>, dim3((unsigned)blocks), dim3(BT), sm, S(s), d, yo, xx, mean, invstd, gamma, beta_mask, P, C, invP, totals, sync, go, dxx, dgamma, dbeta))
      if (beta_mask && !g_out) PM_BNB(2, false);
      else if (beta_mask) PM_BNB(2, true);
      else if (y_out && g_out) PM_BNB(1, true);
      else if (y_out) PM_BNB(1, false);
      else if (g_out) PM_BNB(0, true);
      else PM_BNB(0, false);
#undef PM_BNB
      PM_LAUNCH_OK();
    }
  }
  if (pool_idx) {
    // g_out: optional scratch for the gathered gradient (written in phase 1, re-read in phase 2); NULL = gather twice
    bn_bwd_fused_kernel<T, true, sizeof(T) != 4><<<(int)blocks, BT, smem, S(s)>>>(
        dy, pool_idx, pg, y_out, x, mean, invstd, gamma, beta_mask, P, C, 1.0 / (double)P, partial, totals, sync, g_out, dx,
        dgamma, dbeta);
  } else {
    bn_bwd_fused_kernel<T, false, sizeof(T) != 4><<<(int)blocks, BT, smem, S(s)>>>(
        dy, nullptr, pg, y_out, x, mean, invstd, gamma, beta_mask, P, C, 1.0 / (double)P, partial, totals, sync, g_out, dx,
        dgamma, dbeta);
  }
  PM_LAUNCH_OK();
}

// ---- max pool 3x3 s2 p1 (first max wins, like ATen's CPU kernel); one thread = one 16-byte channel group
template <typename T>
__global__ void maxpool_fwd_kernel(const T* __restrict__ x, int H, int W, int C, int Ho, int Wo, size_t total,
                                   T* __restrict__ y, uint8_t* __restrict__ idx) {
  constexpr int V = Vec<T>::N;
  const int CV = C / V;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int cv = (int)(i % CV);
    size_t t = i / CV;
    const int ow = (int)(t % Wo); t /= Wo;
    const int oh = (int)(t % Ho);
    const size_t b = t / Ho;
    float best[V];
    int bi[V];
    bool first = true;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int ih = oh * 2 - 1 + r;
      if (ih < 0 || ih >= H) continue;
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const int iw = ow * 2 - 1 + s;
        if (iw < 0 || iw >= W) continue;
        float v[V];
        Vec<T>::load(x + ((b * H + ih) * W + iw) * C + cv * V, v);
#pragma unroll
        for (int k = 0; k < V; ++k)
          if (first || v[k] > best[k] || v[k] != v[k]) { best[k] = v[k]; bi[k] = r * 3 + s; }
        first = false;
      }
    }
    const size_t o = i * V;
    Vec<T>::store(y + o, best);
#pragma unroll
    for (int k = 0; k < V; ++k) idx[o + k] = (uint8_t)bi[k];
  }
}

// ---- stem forward in one pass (bf16 throughput mode): BatchNorm (batch statistics finalised here) + ReLU + MaxPool2d(3,2,1).
// The full-resolution activation is never written: one thread = one pooled output x one 16-byte channel group reads its
// 3x3 window of conv outputs (each element is touched by <= 4 windows: L1/L2 hits).  BN followed by ReLU and the bf16
// rounding is monotonic per channel (increasing for gamma*invstd >= 0, decreasing otherwise), so the window maximum of the
// activation is the activation of the window max (min) of the RAW conv output: the 9-tap scan is a packed bf16x2
// compare-select on sign-adjusted raw values and BN is evaluated once per output instead of once per tap.  The pooled
// VALUE is exactly what bn_apply -> max-pool produce; the argmax is the first strict maximum of the raw values (ATen breaks
// ties of the rounded activations by position; both are valid subgradients).  idx = argmax, or 255 when the maximum is
// not positive: ReLU passes no gradient there, so the backward never needs the activation.
__global__ void __launch_bounds__(BT)
bn_relu_maxpool_fwd_kernel(const __nv_bfloat16* __restrict__ x, const double* __restrict__ stats, double Pd, float eps,
                           float momentum, float* __restrict__ mean_out, float* __restrict__ invstd_out, float* __restrict__ rm,
                           float* __restrict__ rv, const float* __restrict__ gamma, const float* __restrict__ beta, int H, int W,
                           int C, int Ho, int Wo, size_t total, __nv_bfloat16* __restrict__ y, uint8_t* __restrict__ idx,
                           __nv_bfloat16* __restrict__ xmax) {
  constexpr int V = 8;
  extern __shared__ float sp[];  // [3][C]: mu, invstd*gamma, beta
  pm_pdl_sync();
  const double invPd = 1.0 / Pd;
  for (int ch = threadIdx.x; ch < C; ch += BT) {
    const double md = stats[ch] * invPd;
    double var = stats[C + ch] * invPd - md * md;
    if (var < 0.0) var = 0.0;
    const float m = (float)md;
    const float ve = (float)(var + (double)eps);
    float is = rsqrtf(ve);
    is = is * (1.5f - 0.5f * ve * is * is);
    if (blockIdx.x == 0) {
      mean_out[ch] = m;
      invstd_out[ch] = is;
      if (rm) {
        const double unbiased = Pd > 1.0 ? var * Pd / (Pd - 1.0) : var;
        rm[ch] = (float)((1.0 - momentum) * (double)rm[ch] + momentum * md);
        rv[ch] = (float)((1.0 - momentum) * (double)rv[ch] + momentum * unbiased);
      }
    }
    sp[ch] = m;
    sp[C + ch] = is * gamma[ch];
    sp[2 * C + ch] = beta[ch];
  }
  __syncthreads();
  const int CV = C / V;
  // blockDim * gridDim is a multiple of CV (= C/8 <= 64 | 256): the channel group of a thread never changes
  const int cv = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) % CV);
  float mu[V], sc[V], be[V];
  uint32_t flip[4];
#pragma unroll
  for (int k = 0; k < V; ++k) { mu[k] = sp[cv * V + k]; sc[k] = sp[C + cv * V + k]; be[k] = sp[2 * C + cv * V + k]; }
#pragma unroll
  for (int q = 0; q < 4; ++q) flip[q] = (sc[2 * q] < 0.f ? 0x8000u : 0u) | (sc[2 * q + 1] < 0.f ? 0x80000000u : 0u);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    size_t t = i / CV;
    const int ow = (int)(t % Wo); t /= Wo;
    const int oh = (int)(t % Ho);
    const size_t b = t / Ho;
    // all nine taps are loaded before the first compare (nine independent 16-byte loads in flight per thread); a tap outside
    // the image is -inf after the sign adjustment, so it never wins a strict comparison and the scan order -- hence the
    // "first strict maximum" tie rule -- is unchanged
    uint4 raw[9];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int ih = oh * 2 - 1 + r;
#pragma unroll
      for (int s2 = 0; s2 < 3; ++s2) {
        const int iw = ow * 2 - 1 + s2;
        const bool ok = ih >= 0 && ih < H && iw >= 0 && iw < W;
        raw[r * 3 + s2] = ok ? *reinterpret_cast<const uint4*>(x + ((b * H + ih) * W + iw) * C + cv * V)
                             : make_uint4(0xFF80FF80u ^ flip[0], 0xFF80FF80u ^ flip[1], 0xFF80FF80u ^ flip[2], 0xFF80FF80u ^ flip[3]);
      }
    }
    uint32_t best[4] = {raw[0].x ^ flip[0], raw[0].y ^ flip[1], raw[0].z ^ flip[2], raw[0].w ^ flip[3]};
    uint32_t bi[4] = {0u, 0u, 0u, 0u};
#pragma unroll
    for (int tp = 1; tp < 9; ++tp) {
      const uint32_t v[4] = {raw[tp].x ^ flip[0], raw[tp].y ^ flip[1], raw[tp].z ^ flip[2], raw[tp].w ^ flip[3]};
      const uint32_t tap2 = (uint32_t)tp * 0x10001u;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t m = __hgt2_mask(*reinterpret_cast<const __nv_bfloat162*>(&v[q]), *reinterpret_cast<const __nv_bfloat162*>(&best[q]));
        best[q] = (v[q] & m) | (best[q] & ~m);
        bi[q] = (tap2 & m) | (bi[q] & ~m);
      }
    }
    float out[V];
    uint8_t id[V];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint32_t xm = best[q] ^ flip[q];
      const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&xm));
      float a0 = (f.x - mu[2 * q]) * sc[2 * q] + be[2 * q];
      float a1 = (f.y - mu[2 * q + 1]) * sc[2 * q + 1] + be[2 * q + 1];
      a0 = a0 > 0.f ? a0 : 0.f;
      a1 = a1 > 0.f ? a1 : 0.f;
      out[2 * q] = a0;
      out[2 * q + 1] = a1;
      // the decision must be taken on the value as stored (bf16): a tiny positive that rounds to +0 passes no gradient
      id[2 * q] = __bfloat162float(__float2bfloat16_rn(a0)) > 0.f ? (uint8_t)(bi[q] & 0xFF) : (uint8_t)255;
      id[2 * q + 1] = __bfloat162float(__float2bfloat16_rn(a1)) > 0.f ? (uint8_t)((bi[q] >> 16) & 0xFF) : (uint8_t)255;
    }
    const size_t o = i * V;
    Vec<__nv_bfloat16>::store(y + o, out);
    *reinterpret_cast<uint2*>(idx + o) = *reinterpret_cast<const uint2*>(id);
    // the RAW conv output at the window's argmax: with it the backward's batch sums (sum g, sum g*xhat) become a reduction over
    // the POOLED tensors only (each pooled gradient reaches exactly one input position -- the one whose raw value this is)
    if (xmax) *reinterpret_cast<uint4*>(xmax + o) = make_uint4(best[0] ^ flip[0], best[1] ^ flip[1], best[2] ^ flip[2], best[3] ^ flip[3]);
  }
}

// PHASE 0 of the stem backward when the forward kept `xmax`: sum g and sum g*xhat over the pooled positions (g = dpool where the
// ReLU is open, idx != 255; xhat from the raw value at the argmax).  Reads 64 MB instead of the 141 MB of the routed scan.
__global__ void __launch_bounds__(BT)
stem_pool_reduce_kernel(const __nv_bfloat16* __restrict__ dpool, const uint8_t* __restrict__ idx, const __nv_bfloat16* __restrict__ xmax,
                        const float* __restrict__ mean, const float* __restrict__ invstd, int C, size_t total, double* __restrict__ sums) {
  constexpr int V = 8;
  const int CV = C / V;
  const int cv = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) % CV);
  pm_pdl_sync();
  float mu[V], is[V], s0[V], s1[V];
#pragma unroll
  for (int k = 0; k < V; ++k) { mu[k] = mean[cv * V + k]; is[k] = invstd[cv * V + k]; s0[k] = s1[k] = 0.f; }
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += 2 * stride) {
    uint4 D[2], X[2];
    uint2 I[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const size_t o = (i + u * stride) * V;
      if (i + u * stride < total) {
        D[u] = *reinterpret_cast<const uint4*>(dpool + o);
        X[u] = *reinterpret_cast<const uint4*>(xmax + o);
        I[u] = *reinterpret_cast<const uint2*>(idx + o);
      } else {
        D[u] = X[u] = make_uint4(0, 0, 0, 0);
        I[u] = make_uint2(0xffffffffu, 0xffffffffu);
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      float d[V], x[V];
      bf8_unpack(D[u], d);
      bf8_unpack(X[u], x);
#pragma unroll
      for (int k = 0; k < V; ++k) {
        const uint32_t byte = ((k < 4 ? I[u].x : I[u].y) >> (8 * (k & 3))) & 0xffu;
        const float g = byte != 255u ? d[k] : 0.f;
        s0[k] += g;
        s1[k] = fmaf(g, (x[k] - mu[k]) * is[k], s1[k]);
      }
    }
  }
  __shared__ float red[2][BT][V + 1];
#pragma unroll
  for (int k = 0; k < V; ++k) { red[0][threadIdx.x][k] = s0[k]; red[1][threadIdx.x][k] = s1[k]; }
  __syncthreads();
  for (int q = threadIdx.x; q < 2 * C; q += BT) {
    const int which = q / C, ch = q % C, vv = ch / V, kk = ch % V;
    double acc = 0.0;
    for (int tdx = vv; tdx < BT; tdx += CV) acc += (double)red[which][tdx][kk];
    atomicAdd(sums + q, acc);
  }
}

// ---- stem backward (bf16 throughput mode): max-pool backward gathered from (dpool, idx) + BatchNorm backward, as a
// reduce launch and an apply launch over 2x2 input blocks.  For MaxPool2d(3,2,1) on an even-sized input the 2x2 block at
// (2k, 2j) receives gradient from at most four windows -- (k,j), (k,j+1), (k+1,j), (k+1,j+1) -- through nine fixed
// (window, tap) pairs, so a thread loads four pooled vectors once and serves four input positions; idx == 255 (ReLU closed)
// never matches a tap.  PHASE 0: s0 += g, s1 += g * xhat (double atomics per block).  PHASE 1: dx = gamma*invstd *
// (g - mean(g) - xhat * mean(g*xhat)).
template <int PHASE>
__global__ void __launch_bounds__(BT, 3)
stem_pool_bn_bwd_kernel(const __nv_bfloat16* __restrict__ dpool, const uint8_t* __restrict__ idx,
                        const __nv_bfloat16* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ invstd,
                        const float* __restrict__ gamma, double* __restrict__ sums, double invP, int H, int W, int C, size_t total,
                        __nv_bfloat16* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  // The twelve vectors of one 2x2 block stay RAW (16-byte bf16 / 8-byte index words, 40 registers) and are unpacked one
  // channel pair at a time: two resident blocks per SM instead of one, every load of the block issued before the first use.
  constexpr int V = 8;
  const int CV = C / V, Ho = H / 2, Wo = W / 2;
  const int cv = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) % CV);
  const int c = cv * V;
  __shared__ float cst[5][64 * 8];  // per-channel constants (C <= 512): mu, is, and for the apply phase A, B, Cc
  pm_pdl_sync();
  for (int ch = threadIdx.x; ch < C; ch += BT) {
    const float m = mean[ch], isd = invstd[ch];
    cst[0][ch] = m;
    cst[1][ch] = isd;
    if (PHASE == 1) {
      // dx = gamma*invstd * (g - mean(g) - xhat * mean(g*xhat)) = A*g + B*x + Cc
      const float a = gamma[ch] * isd;
      const float mg = (float)(sums[ch] * invP), mgx = (float)(sums[C + ch] * invP);
      cst[2][ch] = a;
      cst[3][ch] = -a * mgx * isd;
      cst[4][ch] = a * (mgx * isd * m - mg);
    }
  }
  if (PHASE == 1 && dgamma && blockIdx.x == 0)
    for (int ch = threadIdx.x; ch < C; ch += BT) { dbeta[ch] = (float)sums[ch]; dgamma[ch] = (float)sums[C + ch]; }
  __syncthreads();
  float s0[V], s1[V];
#pragma unroll
  for (int k = 0; k < V; ++k) s0[k] = s1[k] = 0.f;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    size_t t = i / CV;
    const int j = (int)(t % Wo); t /= Wo;
    const int k2 = (int)(t % Ho);
    const size_t b = t / Ho;
    uint4 D[4], X[4];
    uint2 I[4];
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const int ok = k2 + (w >> 1), oj = j + (w & 1);
      D[w] = make_uint4(0, 0, 0, 0);
      I[w] = make_uint2(0xffffffffu, 0xffffffffu);
      if (ok < Ho && oj < Wo) {
        const size_t o = ((b * Ho + ok) * Wo + oj) * C + c;
        D[w] = *reinterpret_cast<const uint4*>(dpool + o);
        I[w] = *reinterpret_cast<const uint2*>(idx + o);
      }
    }
#pragma unroll
    for (int p = 0; p < 4; ++p)
      X[p] = *reinterpret_cast<const uint4*>(x + ((b * H + 2 * k2 + (p >> 1)) * W + 2 * j + (p & 1)) * C + c);
    // Routing with byte-SIMD compares: the eight argmax bytes of a window are tested against a tap with two __vcmpeq4, each
    // byte mask is widened to a bf16x2 lane mask with one PRMT and ANDed onto the packed gradients; the (at most four)
    // contributions of a position are added as packed bf16 (the unfused path rounds the routed gradient to bf16 too).  Nine
    // (window, tap) pairs: 0->(w0,4) | 1->(w0,5),(w1,3) | 2->(w0,7),(w2,1) | 3->(w0,8),(w1,6),(w2,2),(w3,0)
    uint32_t G[4][4];  // [position][channel pair] packed bf16x2
    {
      auto masked = [&](int w, uint32_t tap, uint32_t (&out)[4]) {
        const uint32_t t4 = tap * 0x01010101u;
        const uint32_t mlo = __vcmpeq4(I[w].x, t4), mhi = __vcmpeq4(I[w].y, t4);  // 0xFF per matching channel byte
        out[0] = D[w].x & __byte_perm(mlo, 0, 0x1100);
        out[1] = D[w].y & __byte_perm(mlo, 0, 0x3322);
        out[2] = D[w].z & __byte_perm(mhi, 0, 0x1100);
        out[3] = D[w].w & __byte_perm(mhi, 0, 0x3322);
      };
      auto add2 = [](uint32_t a, uint32_t b2) {
        const __nv_bfloat162 r = __hadd2(*reinterpret_cast<const __nv_bfloat162*>(&a), *reinterpret_cast<const __nv_bfloat162*>(&b2));
        return *reinterpret_cast<const uint32_t*>(&r);
      };
      uint32_t t[4];
      masked(0, 4, G[0]);
      masked(0, 5, G[1]); masked(1, 3, t);
#pragma unroll
      for (int q = 0; q < 4; ++q) G[1][q] = add2(G[1][q], t[q]);
      masked(0, 7, G[2]); masked(2, 1, t);
#pragma unroll
      for (int q = 0; q < 4; ++q) G[2][q] = add2(G[2][q], t[q]);
      masked(0, 8, G[3]); masked(1, 6, t);
#pragma unroll
      for (int q = 0; q < 4; ++q) G[3][q] = add2(G[3][q], t[q]);
      masked(2, 2, t);
#pragma unroll
      for (int q = 0; q < 4; ++q) G[3][q] = add2(G[3][q], t[q]);
      masked(3, 0, t);
#pragma unroll
      for (int q = 0; q < 4; ++q) G[3][q] = add2(G[3][q], t[q]);
    }
    uint4 O[4];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {  // channel pair (2kk, 2kk+1)
      float2 xv[4];
      float g[4][2];
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        const uint32_t xw = kk == 0 ? X[w].x : kk == 1 ? X[w].y : kk == 2 ? X[w].z : X[w].w;
        xv[w] = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&xw));
        const float2 gf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&G[w][kk]));
        g[w][0] = gf.x; g[w][1] = gf.y;
      }
      if (PHASE == 0) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int ch = c + 2 * kk + h;
          const float m = cst[0][ch], isd = cst[1][ch];
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            const float xx = h ? xv[p].y : xv[p].x;
            s0[2 * kk + h] += g[p][h];
            s1[2 * kk + h] = fmaf(g[p][h], (xx - m) * isd, s1[2 * kk + h]);
          }
        }
      } else {
        const int ch = c + 2 * kk;
        const float a0 = cst[2][ch], b0 = cst[3][ch], c0 = cst[4][ch], a1 = cst[2][ch + 1], b1 = cst[3][ch + 1], c1 = cst[4][ch + 1];
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          const __nv_bfloat162 o2 = __floats2bfloat162_rn(fmaf(a0, g[p][0], fmaf(b0, xv[p].x, c0)), fmaf(a1, g[p][1], fmaf(b1, xv[p].y, c1)));
          const uint32_t ow = *reinterpret_cast<const uint32_t*>(&o2);
          if (kk == 0) O[p].x = ow; else if (kk == 1) O[p].y = ow; else if (kk == 2) O[p].z = ow; else O[p].w = ow;
        }
      }
    }
    if (PHASE == 1) {
#pragma unroll
      for (int p = 0; p < 4; ++p)
        *reinterpret_cast<uint4*>(dx + ((b * H + 2 * k2 + (p >> 1)) * W + 2 * j + (p & 1)) * C + c) = O[p];
    }
  }
  if (PHASE == 0) {
    // block reduction over the threads that share a channel group, then one double atomic per channel per block
    __shared__ float red[2][BT][V + 1];
#pragma unroll
    for (int k = 0; k < V; ++k) { red[0][threadIdx.x][k] = s0[k]; red[1][threadIdx.x][k] = s1[k]; }
    __syncthreads();
    for (int q = threadIdx.x; q < 2 * C; q += BT) {
      const int which = q / C, ch = q % C, vv = ch / V, kk = ch % V;
      double acc = 0.0;
      // threads with (tid % CV) == vv hold this channel (the block's first thread index is a multiple of CV)
      for (int tdx = vv; tdx < BT; tdx += CV) acc += (double)red[which][tdx][kk];
      atomicAdd(sums + q, acc);
    }
  }
}

template <typename T>
__global__ void maxpool_bwd_kernel(const T* __restrict__ dy, const uint8_t* __restrict__ idx, int H, int W, int C,
                                   int Ho, int Wo, size_t total, T* __restrict__ dx) {
  constexpr int V = Vec<T>::N;
  const int CV = C / V;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int cv = (int)(i % CV);
    size_t t = i / CV;
    const int w = (int)(t % W); t /= W;
    const int h = (int)(t % H);
    const size_t b = t / H;
    float acc[V];
#pragma unroll
    for (int k = 0; k < V; ++k) acc[k] = 0.f;
    for (int oh = h / 2; oh <= (h + 1) / 2; ++oh) {
      if (oh >= Ho) continue;
      const int r = h - (oh * 2 - 1);
      if (r < 0 || r > 2) continue;
      for (int ow = w / 2; ow <= (w + 1) / 2; ++ow) {
        if (ow >= Wo) continue;
        const int s = w - (ow * 2 - 1);
        if (s < 0 || s > 2) continue;
        const size_t o = (((b * Ho + oh) * Wo + ow) * CV + cv) * V;
        float g[V];
        Vec<T>::load(dy + o, g);
        uint8_t id[V];
        if (V == 8) *reinterpret_cast<uint2*>(id) = *reinterpret_cast<const uint2*>(idx + o);
        else *reinterpret_cast<uint32_t*>(id) = *reinterpret_cast<const uint32_t*>(idx + o);
        const int want = r * 3 + s;
#pragma unroll
        for (int k = 0; k < V; ++k) acc[k] += id[k] == want ? g[k] : 0.f;
      }
    }
    Vec<T>::store(dx + i * V, acc);
  }
}

template <typename T>
__global__ void gap_fwd_kernel(const T* __restrict__ x, int HW, int C, size_t total, float* __restrict__ y) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const size_t b = i / C;
    float s = 0.f;
    for (int p = 0; p < HW; ++p) s += to_f<T>(x[(b * HW + p) * C + c]);
    y[i] = s / (float)HW;
  }
}

template <typename T>
__global__ void gap_bwd_kernel(const float* __restrict__ dy, int HW, int C, size_t total, T* __restrict__ dx) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const size_t b = i / ((size_t)HW * C);
    dx[i] = from_f<T>(dy[b * C + c] / (float)HW);
  }
}

template <typename T>
bool chan_ok(int C) { return C > 0 && C % Vec<T>::N == 0 && BT % (C / Vec<T>::N) == 0; }

// grid for the row-walking kernels: every thread gets >= `min_rows` rows, at most 8 resident blocks per SM
template <typename T>
int row_grid(size_t P, int C, int min_rows) {
  const int rpb = BT / (C / Vec<T>::N);
  size_t blocks = (P + (size_t)rpb * min_rows - 1) / ((size_t)rpb * min_rows);
  const size_t cap = (size_t)pm_num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

template <typename T, typename F>
int launch_reduce(size_t P, int C, double* out, F f, cudaStream_t st) {
  const int rpb = BT / (C / Vec<T>::N);
  size_t blocks = (P + (size_t)rpb * 8 - 1) / ((size_t)rpb * 8);  // >= 8 rows per thread
  const size_t cap = (size_t)pm_num_sms() * 4;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  const size_t smem = 2 * BT * Vec<T>::N * sizeof(double);
  chan_reduce_kernel<T, F><<<(int)blocks, BT, smem, st>>>(P, C, out, f);
  return 0;
}

template <typename T>
int bn_stats_t(const T* x, size_t P, int C, double* stats, pm_stream_t s) {
  PM_CHECK_ARG(x && stats && P > 0 && chan_ok<T>(C));
  launch_reduce<T>(P, C, stats, StatsF<T>{x, C}, S(s));
  PM_LAUNCH_OK();
}
template <typename T>
int bn_apply_t(const T* x, const float* mean, const float* invstd, const float* gamma, const float* beta, const T* res,
               int relu, size_t P, int C, T* y, pm_stream_t s) {
  PM_CHECK_ARG(x && mean && invstd && gamma && beta && y && C % Vec<T>::N == 0);
  PM_CHECK_ARG(chan_ok<T>(C));
  bn_apply_kernel<T, false><<<row_grid<T>(P, C, 4), BT, 4 * C * sizeof(float), S(s)>>>(x, mean, invstd, nullptr, 0.0, 0.f, 0.f, nullptr, nullptr, nullptr,
                                                                   nullptr, gamma, beta, res, relu, P, C, y);
  PM_LAUNCH_OK();
}
template <typename T>
int bn_fwd_fused_t(const T* x, const double* stats, size_t P, int C, float eps, float momentum, const float* gamma,
                   const float* beta, const T* res, int relu, T* y, float* mean, float* invstd, float* rm, float* rv,
                   pm_stream_t s) {
  PM_CHECK_ARG(x && stats && gamma && beta && y && mean && invstd && P > 0 && chan_ok<T>(C) && ((rm == nullptr) == (rv == nullptr)));
  if constexpr (sizeof(T) == 2) {
    if (!bn_bwd_generic()) {
      typedef __nv_bfloat16 b16;
      int grid = row_grid<T>(P, C, 8);
      if (grid > 3 * pm_num_sms()) grid = 3 * pm_num_sms();   // three resident blocks per SM (<= 85 registers)
      PM_CUDA(pm_launch(bn_fwd_bf16_kernel, dim3(grid), dim3(BT), 3 * C * sizeof(float), S(s), (const b16*)x, stats, (double)P, eps, momentum,
                        mean, invstd, rm, rv, gamma, beta, (const b16*)res, relu, P, C, (b16*)y));
      PM_LAUNCH_OK();
    }
  }
  bn_apply_kernel<T, true><<<row_grid<T>(P, C, 4), BT, 4 * C * sizeof(float), S(s)>>>(x, nullptr, nullptr, stats, (double)P, eps, momentum, mean, invstd,
                                                                  rm, rv, gamma, beta, res, relu, P, C, y);
  PM_LAUNCH_OK();
}
template <typename T>
int bn_bwd_reduce_t(const T* dy, const T* y_out, const T* x, const float* mean, const float* invstd, size_t P, int C,
                    double* sums, T* g_out, pm_stream_t s) {
  PM_CHECK_ARG(dy && x && mean && invstd && sums && P > 0 && chan_ok<T>(C));
  launch_reduce<T>(P, C, sums, BwdF<T>{dy, y_out, x, mean, invstd, g_out, C}, S(s));
  PM_LAUNCH_OK();
}
template <typename T>
int bn_bwd_apply_t(const T* dy, const T* y_out, const T* x, const float* mean, const float* invstd, const float* gamma,
                   const double* sums, size_t P, int C, T* dx, float* dgamma, float* dbeta, pm_stream_t s) {
  PM_CHECK_ARG(dy && x && mean && invstd && gamma && sums && dx && C % Vec<T>::N == 0);
  PM_CHECK_ARG(chan_ok<T>(C) && ((dgamma == nullptr) == (dbeta == nullptr)));
  bn_bwd_apply_kernel<T><<<row_grid<T>(P, C, 2), BT, 0, S(s)>>>(dy, y_out, x, mean, invstd, gamma, sums, 1.0 / (double)P, P, C, dx,
                                                                dgamma, dbeta);
  PM_LAUNCH_OK();
}

// ---- avg pool 3x3 s2 p1 (nn.AvgPool2d defaults: zero padding counted, divisor always 9 -- torchlib/models.py:386-387)
template <typename T>
__global__ void avgpool_fwd_kernel(const T* __restrict__ x, int H, int W, int C, int Ho, int Wo, size_t total, T* __restrict__ y) {
  constexpr int V = Vec<T>::N;
  const int CV = C / V;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int cv = (int)(i % CV);
    size_t t = i / CV;
    const int ow = (int)(t % Wo); t /= Wo;
    const int oh = (int)(t % Ho);
    const size_t b = t / Ho;
    float acc[V];
#pragma unroll
    for (int k = 0; k < V; ++k) acc[k] = 0.f;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int ih = oh * 2 - 1 + r;
      if (ih < 0 || ih >= H) continue;
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const int iw = ow * 2 - 1 + s;
        if (iw < 0 || iw >= W) continue;
        float v[V];
        Vec<T>::load(x + ((b * H + ih) * W + iw) * C + cv * V, v);
#pragma unroll
        for (int k = 0; k < V; ++k) acc[k] += v[k];   // window scanned row-major, as ATen's CPU kernel does
      }
    }
#pragma unroll
    for (int k = 0; k < V; ++k) acc[k] = acc[k] / 9.f;
    Vec<T>::store(y + i * V, acc);
  }
}
template <typename T>
__global__ void avgpool_bwd_kernel(const T* __restrict__ dy, int H, int W, int C, int Ho, int Wo, size_t total, T* __restrict__ dx) {
  constexpr int V = Vec<T>::N;
  const int CV = C / V;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int cv = (int)(i % CV);
    size_t t = i / CV;
    const int w = (int)(t % W); t /= W;
    const int h = (int)(t % H);
    const size_t b = t / H;
    float acc[V];
#pragma unroll
    for (int k = 0; k < V; ++k) acc[k] = 0.f;
    for (int oh = h / 2; oh <= (h + 1) / 2; ++oh) {
      if (oh >= Ho) continue;
      for (int ow = w / 2; ow <= (w + 1) / 2; ++ow) {
        if (ow >= Wo) continue;
        float g[V];
        Vec<T>::load(dy + (((b * Ho + oh) * Wo + ow) * CV + cv) * V, g);
#pragma unroll
        for (int k = 0; k < V; ++k) acc[k] += g[k] / 9.f;
      }
    }
    Vec<T>::store(dx + i * V, acc);
  }
}

template <typename T>
int avgpool_fwd_t(const T* x, int B, int H, int W, int C, T* y, pm_stream_t s) {
  PM_CHECK_ARG(x && y && B > 0 && C % Vec<T>::N == 0);
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  const size_t total = (size_t)B * Ho * Wo * (C / Vec<T>::N);
  avgpool_fwd_kernel<T><<<pm_grid(total, 256, 1, 16), 256, 0, S(s)>>>(x, H, W, C, Ho, Wo, total, y);
  PM_LAUNCH_OK();
}
template <typename T>
int avgpool_bwd_t(const T* dy, int B, int H, int W, int C, T* dx, pm_stream_t s) {
  PM_CHECK_ARG(dy && dx && B > 0 && C % Vec<T>::N == 0);
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  const size_t total = (size_t)B * H * W * (C / Vec<T>::N);
  avgpool_bwd_kernel<T><<<pm_grid(total, 256, 1, 16), 256, 0, S(s)>>>(dy, H, W, C, Ho, Wo, total, dx);
  PM_LAUNCH_OK();
}
template <typename T>
int maxpool_fwd_t(const T* x, int B, int H, int W, int C, T* y, uint8_t* idx, pm_stream_t s) {
  PM_CHECK_ARG(x && y && idx && B > 0);
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  PM_CHECK_ARG(C % Vec<T>::N == 0);
  const size_t total = (size_t)B * Ho * Wo * (C / Vec<T>::N);
  maxpool_fwd_kernel<T><<<pm_grid(total, 256, 1, 16), 256, 0, S(s)>>>(x, H, W, C, Ho, Wo, total, y, idx);
  PM_LAUNCH_OK();
}
template <typename T>
int maxpool_bwd_t(const T* dy, const uint8_t* idx, int B, int H, int W, int C, T* dx, pm_stream_t s) {
  PM_CHECK_ARG(dy && dx && idx && B > 0);
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  PM_CHECK_ARG(C % Vec<T>::N == 0);
  const size_t total = (size_t)B * H * W * (C / Vec<T>::N);
  maxpool_bwd_kernel<T><<<pm_grid(total, 256, 1, 16), 256, 0, S(s)>>>(dy, idx, H, W, C, Ho, Wo, total, dx);
  PM_LAUNCH_OK();
}

}  // namespace

typedef __nv_bfloat16 bf16;

extern "C" {

int pm_bn_stats_f32(const float* x, size_t P, int C, double* stats, pm_stream_t s) { return bn_stats_t<float>(x, P, C, stats, s); }
int pm_bn_stats_bf16(const void* x, size_t P, int C, double* stats, pm_stream_t s) { return bn_stats_t<bf16>((const bf16*)x, P, C, stats, s); }

int pm_bn_finalize(const double* stats, size_t P, int C, float eps, float momentum, float* mean, float* invstd,
                   float* running_mean, float* running_var, pm_stream_t s) {
  PM_CHECK_ARG(stats && mean && invstd && P > 0 && C > 0 && ((running_mean == nullptr) == (running_var == nullptr)));
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, S(s)>>>(stats, (double)P, C, eps, momentum, mean, invstd, running_mean,
                                                        running_var);
  PM_LAUNCH_OK();
}

int pm_bn_apply_f32(const float* x, const float* mean, const float* invstd, const float* gamma, const float* beta,
                    const float* residual, int relu, size_t P, int C, float* y, pm_stream_t s) {
  return bn_apply_t<float>(x, mean, invstd, gamma, beta, residual, relu, P, C, y, s);
}
int pm_bn_apply_bf16(const void* x, const float* mean, const float* invstd, const float* gamma, const float* beta,
                     const void* residual, int relu, size_t P, int C, void* y, pm_stream_t s) {
  return bn_apply_t<bf16>((const bf16*)x, mean, invstd, gamma, beta, (const bf16*)residual, relu, P, C, (bf16*)y, s);
}
int pm_bn_fwd_fused_f32(const float* x, const double* stats, size_t P, int C, float eps, float momentum,
                        const float* gamma, const float* beta, const float* residual, int relu, float* y, float* mean,
                        float* invstd, float* running_mean, float* running_var, pm_stream_t s) {
  return bn_fwd_fused_t<float>(x, stats, P, C, eps, momentum, gamma, beta, residual, relu, y, mean, invstd, running_mean, running_var, s);
}
int pm_bn_fwd_fused_bf16(const void* x, const double* stats, size_t P, int C, float eps, float momentum,
                         const float* gamma, const float* beta, const void* residual, int relu, void* y, float* mean,
                         float* invstd, float* running_mean, float* running_var, pm_stream_t s) {
  return bn_fwd_fused_t<bf16>((const bf16*)x, stats, P, C, eps, momentum, gamma, beta, (const bf16*)residual, relu, (bf16*)y, mean,
                              invstd, running_mean, running_var, s);
}
int pm_bn_bwd_reduce_f32(const float* dy, const float* y_out, const float* x, const float* mean, const float* invstd,
                         size_t P, int C, double* sums, float* g_out, pm_stream_t s) {
  return bn_bwd_reduce_t<float>(dy, y_out, x, mean, invstd, P, C, sums, g_out, s);
}
int pm_bn_bwd_reduce_bf16(const void* dy, const void* y_out, const void* x, const float* mean, const float* invstd,
                          size_t P, int C, double* sums, void* g_out, pm_stream_t s) {
  return bn_bwd_reduce_t<bf16>((const bf16*)dy, (const bf16*)y_out, (const bf16*)x, mean, invstd, P, C, sums, (bf16*)g_out, s);
}
int pm_bn_bwd_apply_f32(const float* dy, const float* y_out, const float* x, const float* mean, const float* invstd,
                        const float* gamma, const double* sums, size_t P, int C, float* dx, float* dgamma,
                        float* dbeta, pm_stream_t s) {
  return bn_bwd_apply_t<float>(dy, y_out, x, mean, invstd, gamma, sums, P, C, dx, dgamma, dbeta, s);
}
int pm_bn_bwd_apply_bf16(const void* dy, const void* y_out, const void* x, const float* mean, const float* invstd,
                         const float* gamma, const double* sums, size_t P, int C, void* dx, float* dgamma,
                         float* dbeta, pm_stream_t s) {
  return bn_bwd_apply_t<bf16>((const bf16*)dy, (const bf16*)y_out, (const bf16*)x, mean, invstd, gamma, sums, P, C, (bf16*)dx,
                              dgamma, dbeta, s);
}

size_t pm_bn_bwd_fused_ws_doubles(int C) { return 2 + 4 * 512 + (size_t)pm_num_sms() * 2 * 2 * (size_t)(C > 0 ? C : 512); }
int pm_bn_bwd_fused_f32(const float* dy, const float* y_out, const float* x, const float* mean, const float* invstd,
                        const float* gamma, size_t P, int C, double* ws, float* g_out, float* dx, float* dgamma, float* dbeta,
                        pm_stream_t s) {
  return bn_bwd_fused_t<float>(dy, y_out, x, mean, invstd, gamma, P, C, ws, g_out, dx, dgamma, dbeta, s);
}
int pm_bn_bwd_fused_bf16(const void* dy, const void* y_out, const void* x, const float* mean, const float* invstd,
                         const float* gamma, size_t P, int C, double* ws, void* g_out, void* dx, float* dgamma, float* dbeta,
                         pm_stream_t s) {
  return bn_bwd_fused_t<bf16>((const bf16*)dy, (const bf16*)y_out, (const bf16*)x, mean, invstd, gamma, P, C, ws, (bf16*)g_out,
                              (bf16*)dx, dgamma, dbeta, s);
}

int pm_bn_bwd_fused_xmask_bf16(const void* dy, const void* x, const float* mean, const float* invstd, const float* gamma,
                               const float* beta, size_t P, int C, double* ws, void* dx, float* dgamma, float* dbeta,
                               pm_stream_t s) {
  PM_CHECK_ARG(beta != nullptr);
  return bn_bwd_fused_t<bf16>((const bf16*)dy, nullptr, (const bf16*)x, mean, invstd, gamma, P, C, ws, nullptr, (bf16*)dx, dgamma,
                              dbeta, s, nullptr, PoolGeo{0, 0, 0, 0}, beta);
}

int pm_bn_relu_maxpool_fwd_bf16(const void* x, const double* stats, int B, int H, int W, int C, float eps, float momentum,
                                const float* gamma, const float* beta, void* y, uint8_t* idx, void* xmax, float* mean, float* invstd,
                                float* running_mean, float* running_var, pm_stream_t s) {
  PM_CHECK_ARG(x && stats && gamma && beta && y && idx && mean && invstd && B > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0);
  PM_CHECK_ARG((running_mean == nullptr) == (running_var == nullptr));
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  const size_t total = (size_t)B * Ho * Wo * (C / 8);
  PM_CHECK_ARG(BT % (C / 8) == 0);
  PM_CUDA(pm_launch(bn_relu_maxpool_fwd_kernel, dim3(pm_grid(total, BT, 1, 16)), dim3(BT), 3 * C * sizeof(float), S(s), (const bf16*)x, stats,
                    (double)B * H * W, eps, momentum, mean, invstd, running_mean, running_var, gamma, beta, H, W, C, Ho, Wo, total,
                    (bf16*)y, idx, (bf16*)xmax));
  PM_LAUNCH_OK();
}

int pm_stem_pool_bn_bwd_bf16(const void* dpool, const uint8_t* pool_idx, const void* xmax, int B, int H, int W, const void* x,
                             const float* mean, const float* invstd, const float* gamma, int C, double* sums, void* dx, float* dgamma,
                             float* dbeta, pm_stream_t s) {
  PM_CHECK_ARG(dpool && pool_idx && x && mean && invstd && gamma && sums && dx && B > 0 && H > 0 && W > 0);
  PM_CHECK_ARG(H % 2 == 0 && W % 2 == 0 && C % 8 == 0 && BT % (C / 8) == 0 && ((dgamma == nullptr) == (dbeta == nullptr)));
  const size_t total = (size_t)B * (H / 2) * (W / 2) * (C / 8);
  const int grid = pm_grid(total, BT, 1, 8);
  const double invP = 1.0 / ((double)B * H * W);
  if (xmax)
    PM_CUDA(pm_launch(stem_pool_reduce_kernel, dim3(pm_grid(total, BT, 2, 8)), dim3(BT), 0, S(s), (const bf16*)dpool, pool_idx,
                      (const bf16*)xmax, mean, invstd, C, total, sums));
  else
    PM_CUDA(pm_launch(stem_pool_bn_bwd_kernel<0>, dim3(grid), dim3(BT), 0, S(s), (const bf16*)dpool, pool_idx, (const bf16*)x, mean, invstd,
                    gamma, sums, invP, H, W, C, total, (bf16*)nullptr, (float*)nullptr, (float*)nullptr));
  PM_CUDA(pm_launch(stem_pool_bn_bwd_kernel<1>, dim3(grid), dim3(BT), 0, S(s), (const bf16*)dpool, pool_idx, (const bf16*)x, mean, invstd,
                    gamma, sums, invP, H, W, C, total, (bf16*)dx, dgamma, dbeta));
  PM_LAUNCH_OK();
}

int pm_bn_bwd_fused_pool_bf16(const void* dpool, const uint8_t* pool_idx, int B, int H, int W, const void* y_out, const void* x,
                              const float* mean, const float* invstd, const float* gamma, int C, double* ws, void* g_scratch,
                              void* dx, float* dgamma, float* dbeta, pm_stream_t s) {
  PM_CHECK_ARG(dpool && pool_idx && B > 0 && H > 0 && W > 0);
  const PoolGeo pg{H, W, (H + 2 - 3) / 2 + 1, (W + 2 - 3) / 2 + 1};
  return bn_bwd_fused_t<bf16>((const bf16*)dpool, (const bf16*)y_out, (const bf16*)x, mean, invstd, gamma, (size_t)B * H * W, C, ws,
                              (bf16*)g_scratch, (bf16*)dx, dgamma, dbeta, s, pool_idx, pg);
}
int pm_bn_bwd_fused_pool_f32(const float* dpool, const uint8_t* pool_idx, int B, int H, int W, const float* y_out, const float* x,
                             const float* mean, const float* invstd, const float* gamma, int C, double* ws, float* g_scratch,
                             float* dx, float* dgamma, float* dbeta, pm_stream_t s) {
  PM_CHECK_ARG(dpool && pool_idx && B > 0 && H > 0 && W > 0);
  const PoolGeo pg{H, W, (H + 2 - 3) / 2 + 1, (W + 2 - 3) / 2 + 1};
  return bn_bwd_fused_t<float>(dpool, y_out, x, mean, invstd, gamma, (size_t)B * H * W, C, ws, g_scratch, dx, dgamma, dbeta, s,
                               pool_idx, pg);
}

int pm_maxpool3s2_fwd_f32(const float* x, int B, int H, int W, int C, float* y, uint8_t* idx, pm_stream_t s) {
  return maxpool_fwd_t<float>(x, B, H, W, C, y, idx, s);
}
int pm_maxpool3s2_bwd_f32(const float* dy, const uint8_t* idx, int B, int H, int W, int C, float* dx, pm_stream_t s) {
  return maxpool_bwd_t<float>(dy, idx, B, H, W, C, dx, s);
}
int pm_avgpool3s2_fwd_f32(const float* x, int B, int H, int W, int C, float* y, pm_stream_t s) { return avgpool_fwd_t<float>(x, B, H, W, C, y, s); }
int pm_avgpool3s2_bwd_f32(const float* dy, int B, int H, int W, int C, float* dx, pm_stream_t s) { return avgpool_bwd_t<float>(dy, B, H, W, C, dx, s); }
int pm_avgpool3s2_fwd_bf16(const void* x, int B, int H, int W, int C, void* y, pm_stream_t s) {
  return avgpool_fwd_t<bf16>((const bf16*)x, B, H, W, C, (bf16*)y, s);
}
int pm_avgpool3s2_bwd_bf16(const void* dy, int B, int H, int W, int C, void* dx, pm_stream_t s) {
  return avgpool_bwd_t<bf16>((const bf16*)dy, B, H, W, C, (bf16*)dx, s);
}
int pm_maxpool3s2_fwd_bf16(const void* x, int B, int H, int W, int C, void* y, uint8_t* idx, pm_stream_t s) {
  return maxpool_fwd_t<bf16>((const bf16*)x, B, H, W, C, (bf16*)y, idx, s);
}
int pm_maxpool3s2_bwd_bf16(const void* dy, const uint8_t* idx, int B, int H, int W, int C, void* dx, pm_stream_t s) {
  return maxpool_bwd_t<bf16>((const bf16*)dy, idx, B, H, W, C, (bf16*)dx, s);
}

int pm_gap_fwd_f32(const float* x, int B, int HW, int C, float* y, pm_stream_t s) {
  PM_CHECK_ARG(x && y);
  const size_t total = (size_t)B * C;
  gap_fwd_kernel<float><<<pm_grid(total, 128), 128, 0, S(s)>>>(x, HW, C, total, y);
  PM_LAUNCH_OK();
}
int pm_gap_fwd_bf16(const void* x, int B, int HW, int C, float* y, pm_stream_t s) {
  PM_CHECK_ARG(x && y);
  const size_t total = (size_t)B * C;
  gap_fwd_kernel<bf16><<<pm_grid(total, 128), 128, 0, S(s)>>>((const bf16*)x, HW, C, total, y);
  PM_LAUNCH_OK();
}
int pm_gap_bwd_f32(const float* dy, int B, int HW, int C, float* dx, pm_stream_t s) {
  PM_CHECK_ARG(dy && dx);
  const size_t total = (size_t)B * HW * C;
  gap_bwd_kernel<float><<<pm_grid(total, 256), 256, 0, S(s)>>>(dy, HW, C, total, dx);
  PM_LAUNCH_OK();
}
int pm_gap_bwd_bf16(const float* dy, int B, int HW, int C, void* dx, pm_stream_t s) {
  PM_CHECK_ARG(dy && dx);
  const size_t total = (size_t)B * HW * C;
  gap_bwd_kernel<bf16><<<pm_grid(total, 256), 256, 0, S(s)>>>(dy, HW, C, total, (bf16*)dx);
  PM_LAUNCH_OK();
}

}  // extern "C"
