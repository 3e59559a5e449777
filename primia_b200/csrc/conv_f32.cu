// Path T, fp32 "parity mode" convolutions: implicit-GEMM on the FP32 pipe (FFMA), NHWC / KRSC.
// These are the kernels the 1e-5 relative parity gate runs on; the throughput mode is conv_tc.cu
// (bf16 operands, tcgen05 tensor cores, fp32 accumulation in TMEM).
//
//   fwd  : y[m, n]  = sum_{r,s,c} x[b, oh*st-pad+r, ow*st-pad+s, c] * w[n, r, s, c]      m = (b, oh, ow)
//   dgrad: dx[m, c] = sum_{r,s,k} dy[b, (h+pad-r)/st, (w+pad-s)/st, k] * w[k, r, s, c]    m = (b, h, w)
//   wgrad: dw[n, r, s, c] = sum_m dy[m, n] * x[m @ (r,s), c]                              split over m
#include "common.cuh"

namespace {

constexpr int BM = 128, BN = 64, BK = 16, NT = 256, APAD = 4, BPAD = 4;

struct ConvP {
  int B, H, W, C, K, R, S, stride, pad, Ho, Wo;
};

// MODE 0 = forward, 1 = dgrad.  GENERIC: per-element k decode (needed when the reduction-channel count
// is not a multiple of BK, i.e. conv1 with C == 3).
template <int MODE, bool GENERIC>
__global__ void __launch_bounds__(NT, 2)
conv_f32_kernel(ConvP p, const float* __restrict__ src, const float* __restrict__ w, float* __restrict__ dst,
                int accumulate) {
  __shared__ __align__(16) float As[2][BK][BM + APAD];
  __shared__ __align__(16) float Bs[2][BK][BN + BPAD];

  const int tid = threadIdx.x;
  // GEMM view
  const int IH = MODE == 0 ? p.H : p.Ho, IW = MODE == 0 ? p.W : p.Wo;    // spatial dims of src
  const int OH = MODE == 0 ? p.Ho : p.H, OW = MODE == 0 ? p.Wo : p.W;    // spatial dims of dst
  const int CR = MODE == 0 ? p.C : p.K;                                   // reduction channels (src channels)
  const int N = MODE == 0 ? p.K : p.C;                                    // dst channels
  const int M = p.B * OH * OW;
  const int Kg = p.R * p.S * CR;
  const int RSC = p.R * p.S * p.C;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;

  // ---- A loader: thread owns row (tid >> 1), 8 consecutive k at (tid & 1) * 8
  const int a_row = tid >> 1, a_k8 = (tid & 1) * 8;
  const int am = m0 + a_row;
  const bool a_row_ok = am < M;
  int ab = 0, aoh = 0, aow = 0;
  if (a_row_ok) {
    ab = am / (OH * OW);
    const int rem = am - ab * OH * OW;
    aoh = rem / OW;
    aow = rem - aoh * OW;
  }
  const float* src_b = src + (size_t)ab * IH * IW * CR;
  // ---- B loader
  const int b_n_f = tid >> 2, b_k4_f = (tid & 3) * 4;   // fwd: n = tid/4, 4 consecutive k
  const int b_k_d = tid >> 4, b_n4_d = (tid & 15) * 4;  // dgrad: k = tid/16, 4 consecutive n

  const int tx = tid & 15, ty = tid >> 4;
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  float ra[8], rb[4];
  const int nchunks = (Kg + BK - 1) / BK;

  auto src_coord = [&](int r, int s, int& ih, int& iw) -> bool {
    if (MODE == 0) {
      ih = aoh * p.stride - p.pad + r;
      iw = aow * p.stride - p.pad + s;
      return ih >= 0 && ih < IH && iw >= 0 && iw < IW;
    } else {
      const int th = aoh + p.pad - r, tw = aow + p.pad - s;
      if (th < 0 || tw < 0) return false;
      if (p.stride != 1) {
        if (th % p.stride || tw % p.stride) return false;
        ih = th / p.stride;
        iw = tw / p.stride;
      } else {
        ih = th;
        iw = tw;
      }
      return ih < IH && iw < IW;
    }
  };

  auto gload = [&](int chunk) {
    const int k0 = chunk * BK;
    if (!GENERIC) {
      const int tap = k0 / CR, c0 = k0 - tap * CR;
      const int r = tap / p.S, s = tap - r * p.S;
      int ih, iw;
      if (a_row_ok && src_coord(r, s, ih, iw)) {
        const float4* g = reinterpret_cast<const float4*>(src_b + ((size_t)ih * IW + iw) * CR + c0 + a_k8);
        const float4 v0 = __ldg(g), v1 = __ldg(g + 1);
        ra[0] = v0.x; ra[1] = v0.y; ra[2] = v0.z; ra[3] = v0.w;
        ra[4] = v1.x; ra[5] = v1.y; ra[6] = v1.z; ra[7] = v1.w;
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) ra[i] = 0.f;
      }
      if (MODE == 0) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(w + (size_t)(n0 + b_n_f) * RSC + tap * p.C + c0 + b_k4_f));
        rb[0] = v.x; rb[1] = v.y; rb[2] = v.z; rb[3] = v.w;
      } else {
        // Bw[(tap,k), c] = w[k][tap][c]
        const float4 v = __ldg(reinterpret_cast<const float4*>(w + (size_t)(c0 + b_k_d) * RSC + tap * p.C + n0 + b_n4_d));
        rb[0] = v.x; rb[1] = v.y; rb[2] = v.z; rb[3] = v.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int k = k0 + a_k8 + i;
        float v = 0.f;
        if (a_row_ok && k < Kg) {
          const int tap = k / CR, c = k - tap * CR;
          const int r = tap / p.S, s = tap - r * p.S;
          int ih, iw;
          if (src_coord(r, s, ih, iw)) v = __ldg(src_b + ((size_t)ih * IW + iw) * CR + c);
        }
        ra[i] = v;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int k = k0 + b_k4_f + i;
        float v = 0.f;
        if (k < Kg) {
          if (MODE == 0) v = __ldg(w + (size_t)(n0 + b_n_f) * RSC + k);
          else {
            const int tap = k / CR, kk = k - tap * CR;
            v = __ldg(w + (size_t)kk * RSC + tap * p.C + n0 + b_n_f);  // generic dgrad: n = b_n_f (unused in practice)
          }
        }
        rb[i] = v;
      }
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 8; ++i) As[buf][a_k8 + i][a_row] = ra[i];
    if (MODE == 0 || GENERIC) {
#pragma unroll
      for (int i = 0; i < 4; ++i) Bs[buf][b_k4_f + i][b_n_f] = rb[i];
    } else {
      *reinterpret_cast<float4*>(&Bs[buf][b_k_d][b_n4_d]) = make_float4(rb[0], rb[1], rb[2], rb[3]);
    }
  };

  gload(0);
  sstore(0);
  __syncthreads();
  int buf = 0;
  for (int chunk = 0; chunk < nchunks; ++chunk) {
    const bool has_next = chunk + 1 < nchunks;
    if (has_next) gload(chunk + 1);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8 + 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[4] = {b0.x, b0.y, b0.z, b0.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (has_next) {
      sstore(buf ^ 1);
      __syncthreads();
      buf ^= 1;
    }
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + ty * 8 + i;
    if (m >= M) continue;
    float4* o = reinterpret_cast<float4*>(dst + (size_t)m * N + n0 + tx * 4);
    float4 v = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    if (accumulate) {
      const float4 old = *o;
      v.x += old.x; v.y += old.y; v.z += old.z; v.w += old.w;
    }
    *o = v;
  }
}

// ---------------------------------------------------------------- wgrad: 64(n) x 64(kg) tile, reduce over rows
constexpr int WT = 64, WK = 16;

template <bool GENERIC>
__global__ void __launch_bounds__(NT, 2)
wgrad_f32_kernel(ConvP p, const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ part,
                 int rows_per_split) {
  __shared__ __align__(16) float Ds[2][WK][WT + 4];  // dy  [m][n]
  __shared__ __align__(16) float Xs[2][WK][WT + 4];  // x@tap [m][kg]
  const int tid = threadIdx.x;
  const int M = p.B * p.Ho * p.Wo;
  const int Kg = p.R * p.S * p.C;
  const int n0 = blockIdx.y * WT, kg0 = blockIdx.x * WT;
  const int split = blockIdx.z;
  const int m_begin = split * rows_per_split;
  const int m_end = min(M, m_begin + rows_per_split);

  const int l_m = tid >> 4, l_c4 = (tid & 15) * 4;  // loader: row within chunk, 4 consecutive columns
  const int tx = tid & 15, ty = tid >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  // non-generic: the 64 kg columns live inside one tap
  int tap = 0, c0 = 0, r = 0, s = 0;
  if (!GENERIC) {
    tap = kg0 / p.C;
    c0 = kg0 - tap * p.C;
    r = tap / p.S;
    s = tap - r * p.S;
  }
  float4 rd, rx;
  auto gload = [&](int mbase) {
    const int m = mbase + l_m;
    rd = make_float4(0.f, 0.f, 0.f, 0.f);
    rx = rd;
    if (m < m_end) {
      rd = __ldg(reinterpret_cast<const float4*>(dy + (size_t)m * p.K + n0 + l_c4));
      const int b = m / (p.Ho * p.Wo);
      const int rem = m - b * p.Ho * p.Wo;
      const int oh = rem / p.Wo, ow = rem - oh * p.Wo;
      if (!GENERIC) {
        const int ih = oh * p.stride - p.pad + r, iw = ow * p.stride - p.pad + s;
        if (ih >= 0 && ih < p.H && iw >= 0 && iw < p.W)
          rx = __ldg(reinterpret_cast<const float4*>(x + (((size_t)b * p.H + ih) * p.W + iw) * p.C + c0 + l_c4));
      } else {
        float t[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int kg = kg0 + l_c4 + i;
          t[i] = 0.f;
          if (kg < Kg) {
            const int tp = kg / p.C, c = kg - tp * p.C;
            const int rr = tp / p.S, ss = tp - rr * p.S;
            const int ih = oh * p.stride - p.pad + rr, iw = ow * p.stride - p.pad + ss;
            if (ih >= 0 && ih < p.H && iw >= 0 && iw < p.W) t[i] = __ldg(x + (((size_t)b * p.H + ih) * p.W + iw) * p.C + c);
          }
        }
        rx = make_float4(t[0], t[1], t[2], t[3]);
      }
    }
  };
  auto sstore = [&](int buf) {
    *reinterpret_cast<float4*>(&Ds[buf][l_m][l_c4]) = rd;
    *reinterpret_cast<float4*>(&Xs[buf][l_m][l_c4]) = rx;
  };
  if (m_begin < m_end) {
    gload(m_begin);
    sstore(0);
    __syncthreads();
    int buf = 0;
    for (int mb = m_begin; mb < m_end; mb += WK) {
      const bool has_next = mb + WK < m_end;
      if (has_next) gload(mb + WK);
#pragma unroll
      for (int k = 0; k < WK; ++k) {
        const float4 a = *reinterpret_cast<const float4*>(&Ds[buf][k][ty * 4]);
        const float4 b = *reinterpret_cast<const float4*>(&Xs[buf][k][tx * 4]);
        const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
      if (has_next) {
        sstore(buf ^ 1);
        __syncthreads();
        buf ^= 1;
      }
    }
  }
  float* out = part + (size_t)split * p.K * Kg;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = n0 + ty * 4 + i;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int kg = kg0 + tx * 4 + j;
      if (kg < Kg) out[(size_t)n * Kg + kg] = acc[i][j];
    }
  }
}

__global__ void split_reduce_kernel(const float* __restrict__ part, int splits, size_t n, float* __restrict__ out) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    float s = 0.f;
    for (int k = 0; k < splits; ++k) s += part[(size_t)k * n + i];
    out[i] = s;
  }
}

int wgrad_splits(const pm_conv_t* p) {
  const long M = (long)p->B * p->Ho * p->Wo;
  const long Kg = (long)p->R * p->S * p->C;
  const long tiles = ((Kg + WT - 1) / WT) * (p->K / WT);
  const long target = 4L * pm_num_sms();
  long splits = (target + tiles - 1) / tiles;
  const long max_splits = (M + 4 * WK - 1) / (4 * WK);  // >= 64 rows per split
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  if (splits > 512) splits = 512;
  return (int)splits;
}

ConvP to_p(const pm_conv_t* p) { return ConvP{p->B, p->H, p->W, p->C, p->K, p->R, p->S, p->stride, p->pad, p->Ho, p->Wo}; }

bool conv_ok(const pm_conv_t* p) {
  return p && p->B > 0 && p->H > 0 && p->W > 0 && p->C > 0 && p->K > 0 && p->R > 0 && p->S > 0 && p->stride > 0 && p->pad >= 0 &&
         p->Ho == (p->H + 2 * p->pad - p->R) / p->stride + 1 && p->Wo == (p->W + 2 * p->pad - p->S) / p->stride + 1;
}

// layout shuffles
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, int C, int HW, float* __restrict__ out) {
  // per batch b (blockIdx.z): [C,HW] -> [HW,C]
  __shared__ float tile[32][33];
  const int b = blockIdx.z, hw0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const float* xb = x + (size_t)b * C * HW;
  float* ob = out + (size_t)b * C * HW;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int c = c0 + r, hw = hw0 + threadIdx.x;
    if (c < C && hw < HW) tile[r][threadIdx.x] = xb[(size_t)c * HW + hw];
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int hw = hw0 + r, c = c0 + threadIdx.x;
    if (c < C && hw < HW) ob[(size_t)hw * C + c] = tile[threadIdx.x][r];
  }
}

__global__ void kcrs_krsc_kernel(const float* __restrict__ w, int C, int RS, size_t total, int to_krsc,
                                 float* __restrict__ out) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    // i indexes the destination
    const size_t k = i / ((size_t)C * RS);
    const int rem = (int)(i - k * C * RS);
    if (to_krsc) {  // dst [k][rs][c]  <- src [k][c][rs]
      const int rs = rem / C, c = rem - rs * C;
      out[i] = w[(k * C + c) * RS + rs];
    } else {        // dst [k][c][rs]  <- src [k][rs][c]
      const int c = rem / RS, rs = rem - c * RS;
      out[i] = w[(k * RS + rs) * C + c];
    }
  }
}

}  // namespace

extern "C" {

int pm_conv_fwd_f32(const pm_conv_t* p, const float* x, const float* w, float* y, pm_stream_t s) {
  PM_CHECK_ARG(conv_ok(p) && x && w && y && p->K % BN == 0);
  const int M = p->B * p->Ho * p->Wo;
  dim3 grid((M + BM - 1) / BM, p->K / BN);
  if (p->C % BK == 0) conv_f32_kernel<0, false><<<grid, NT, 0, S(s)>>>(to_p(p), x, w, y, 0);
  else conv_f32_kernel<0, true><<<grid, NT, 0, S(s)>>>(to_p(p), x, w, y, 0);
  PM_LAUNCH_OK();
}

int pm_conv_dgrad_f32(const pm_conv_t* p, const float* dy, const float* w, float* dx, int accumulate, pm_stream_t s) {
  PM_CHECK_ARG(conv_ok(p) && dy && w && dx && p->C % BN == 0 && p->K % BK == 0);
  const int M = p->B * p->H * p->W;
  dim3 grid((M + BM - 1) / BM, p->C / BN);
  conv_f32_kernel<1, false><<<grid, NT, 0, S(s)>>>(to_p(p), dy, w, dx, accumulate);
  PM_LAUNCH_OK();
}

size_t pm_conv_wgrad_ws_bytes(const pm_conv_t* p) {
  if (!conv_ok(p)) return 0;
  return (size_t)wgrad_splits(p) * p->K * p->R * p->S * p->C * sizeof(float);
}

int pm_conv_wgrad_f32(const pm_conv_t* p, const float* x, const float* dy, float* dw, void* ws, pm_stream_t s) {
  PM_CHECK_ARG(conv_ok(p) && x && dy && dw && ws && p->K % WT == 0);
  const int M = p->B * p->Ho * p->Wo;
  const int Kg = p->R * p->S * p->C;
  const int splits = wgrad_splits(p);
  int rows = (M + splits - 1) / splits;
  rows = (rows + WK - 1) / WK * WK;
  dim3 grid((Kg + WT - 1) / WT, p->K / WT, splits);
  if (p->C % WT == 0) wgrad_f32_kernel<false><<<grid, NT, 0, S(s)>>>(to_p(p), x, dy, (float*)ws, rows);
  else wgrad_f32_kernel<true><<<grid, NT, 0, S(s)>>>(to_p(p), x, dy, (float*)ws, rows);
  const size_t n = (size_t)p->K * Kg;
  split_reduce_kernel<<<pm_grid(n, 256), 256, 0, S(s)>>>((const float*)ws, splits, n, dw);
  PM_LAUNCH_OK();
}

/* DP-SGD, fp32 parity mode: per-sample weight gradients dw[b][K][R*S*C], one deterministic wgrad per image */
int pm_conv_wgrad_persample_f32(const pm_conv_t* p, const float* x, const float* dy, float* dw, void* ws, pm_stream_t s) {
  PM_CHECK_ARG(conv_ok(p) && x && dy && dw && ws);
  pm_conv_t one = *p;
  one.B = 1;
  const size_t xs = (size_t)p->H * p->W * p->C, ys = (size_t)p->Ho * p->Wo * p->K, ws_n = (size_t)p->K * p->R * p->S * p->C;
  for (int b = 0; b < p->B; ++b) {
    const int r = pm_conv_wgrad_f32(&one, x + b * xs, dy + b * ys, dw + b * ws_n, ws, s);
    if (r != PM_OK) return r;
  }
  return PM_OK;
}

int pm_nchw_to_nhwc_f32(const float* x, int B, int C, int H, int W, float* out, pm_stream_t s) {
  PM_CHECK_ARG(x && out && B > 0 && B <= 65535 && C > 0);
  dim3 grid((H * W + 31) / 32, (C + 31) / 32, B), block(32, 8);
  nchw_to_nhwc_kernel<<<grid, block, 0, S(s)>>>(x, C, H * W, out);
  PM_LAUNCH_OK();
}

int pm_kcrs_to_krsc_f32(const float* w, int K, int C, int R, int S_, float* out, pm_stream_t s) {
  PM_CHECK_ARG(w && out);
  const size_t total = (size_t)K * C * R * S_;
  kcrs_krsc_kernel<<<pm_grid(total, 256), 256, 0, S(s)>>>(w, C, R * S_, total, 1, out);
  PM_LAUNCH_OK();
}

int pm_krsc_to_kcrs_f32(const float* w, int K, int C, int R, int S_, float* out, pm_stream_t s) {
  PM_CHECK_ARG(w && out);
  const size_t total = (size_t)K * C * R * S_;
  kcrs_krsc_kernel<<<pm_grid(total, 256), 256, 0, S(s)>>>(w, C, R * S_, total, 0, out);
  PM_LAUNCH_OK();
}

}  // extern "C"
