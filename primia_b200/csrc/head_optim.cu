// Path T: classifier head (Linear + cross-entropy, forward and backward), optimizers on the flat
// parameter buffer, FedAvg scaling and dtype/layout helpers.
#include "common.cuh"

namespace {

constexpr int MAXC = 16;  // max classes handled by the fused head

// one block per sample: logits, log-softmax, per-sample loss weight & nll, dlogits (unnormalised)
__global__ void head_fwd_kernel(const float* __restrict__ feat, const float* __restrict__ W,
                                const float* __restrict__ bias, const int64_t* __restrict__ labels,
                                const float* __restrict__ soft, const float* __restrict__ cw, int F, int ncls,
                                float* __restrict__ logits, float* __restrict__ ws) {
  const int b = blockIdx.x;
  __shared__ float red[MAXC][32];
  __shared__ float lg[MAXC];
  const float* f = feat + (size_t)b * F;
  float part[MAXC];
  for (int j = 0; j < ncls; ++j) part[j] = 0.f;
  for (int i = threadIdx.x; i < F; i += blockDim.x) {
    const float v = f[i];
    for (int j = 0; j < ncls; ++j) part[j] = fmaf(v, W[(size_t)j * F + i], part[j]);
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int j = 0; j < ncls; ++j) {
    const float v = warp_sum(part[j]);
    if (lane == 0) red[j][wid] = v;
  }
  __syncthreads();
  if (threadIdx.x < ncls) {
    float v = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += red[threadIdx.x][w];
    v += bias[threadIdx.x];
    lg[threadIdx.x] = v;
    logits[(size_t)b * ncls + threadIdx.x] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0 && ws) {
    float mx = lg[0];
    for (int j = 1; j < ncls; ++j) mx = fmaxf(mx, lg[j]);
    float se = 0.f;
    for (int j = 0; j < ncls; ++j) se += expf(lg[j] - mx);
    const float lse = mx + logf(se);
    float* dl = ws + (size_t)b * (ncls + 1);  // [ncls] dlogits (unnormalised) + [1] weighted nll
    if (labels) {
      const int y = (int)labels[b];
      const float wgt = cw ? cw[y] : 1.f;
      for (int j = 0; j < ncls; ++j) dl[j] = wgt * (expf(lg[j] - lse) - (j == y ? 1.f : 0.f));
      dl[ncls] = wgt * (lse - lg[y]);
      // denominators are accumulated by head_bwd_kernel (sum of wgt) -- stash wgt in the last slot of ws
      atomicAdd(ws + (size_t)gridDim.x * (ncls + 1), wgt);
    } else {
      const float* t = soft + (size_t)b * ncls;
      float wsum = 1.f, tsum = 0.f, nll = 0.f;
      if (cw) { wsum = 0.f; for (int j = 0; j < ncls; ++j) wsum += cw[j] * t[j]; }
      for (int j = 0; j < ncls; ++j) { tsum += t[j]; nll += -t[j] * (lg[j] - lse); }
      for (int j = 0; j < ncls; ++j) dl[j] = wsum * (expf(lg[j] - lse) * tsum - t[j]);
      dl[ncls] = wsum * nll;
      atomicAdd(ws + (size_t)gridDim.x * (ncls + 1), 1.f);
    }
  }
}

// normalise, loss, dfeat, dW, db.  grid: (ceil(F/128)), block 128: each thread owns one feature column f.
__global__ void head_bwd_kernel(const float* __restrict__ feat, const float* __restrict__ W, int B, int F, int ncls,
                                const float* __restrict__ ws, float* __restrict__ loss, float* __restrict__ dfeat,
                                float* __restrict__ dW, float* __restrict__ db) {
  const float inv = 1.f / ws[(size_t)B * (ncls + 1)];
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f < F) {
    float dw[MAXC], wcol[MAXC];
    for (int j = 0; j < ncls; ++j) { dw[j] = 0.f; wcol[j] = W[(size_t)j * F + f]; }
    // rows in batches of 8: the 8 feature loads are issued together (the loop was one dependent ~600-cycle load per row);
    // the accumulation order over b is unchanged
    for (int b0 = 0; b0 < B; b0 += 8) {
      float xs[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) xs[u] = b0 + u < B ? feat[(size_t)(b0 + u) * F + f] : 0.f;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int b = b0 + u;
        if (b < B) {
          const float* dl = ws + (size_t)b * (ncls + 1);
          float df = 0.f;
          for (int j = 0; j < ncls; ++j) {
            const float g = dl[j] * inv;
            dw[j] = fmaf(g, xs[u], dw[j]);
            df = fmaf(g, wcol[j], df);
          }
          dfeat[(size_t)b * F + f] = df;
        }
      }
    }
    for (int j = 0; j < ncls; ++j) dW[(size_t)j * F + f] = dw[j];
  }
  if (blockIdx.x == 0 && threadIdx.x < ncls) {
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += ws[(size_t)b * (ncls + 1) + threadIdx.x];
    db[threadIdx.x] = s * inv;
  }
  if (blockIdx.x == 0 && threadIdx.x == 32) {
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += ws[(size_t)b * (ncls + 1) + ncls];
    loss[0] = s * inv;
  }
}

// ---- the whole classifier head of one sample in ONE launch (critical path of a step: forward tail -> backward head):
//   AvgPool2d(HW) of the last activation -> feat ; logits = feat W^T + b ; per-sample CE gradient dl (unnormalised, into ws as
//   head_fwd_kernel leaves it) ; dfeat = (dl / denom) W ; gradient of the average pool written straight into d_out [B,HW,F].
// The batch-reduced quantities (dW, db, loss) are NOT needed by the backward chain: head_grads_kernel computes them from
// (ws, feat) afterwards, on the side stream.  denom = sum_b wgt_b is recomputed by every block from the labels (B <= a few
// hundred) so that no grid-wide step is needed; block 0 also stores it in ws[B*(ncls+1)].
template <typename T>
__global__ void __launch_bounds__(256)
head_fused_kernel(const T* __restrict__ x, int HW, const float* __restrict__ W, const float* __restrict__ bias,
                  const int64_t* __restrict__ labels, const float* __restrict__ soft, const float* __restrict__ cw, int B, int F, int ncls,
                  float* __restrict__ feat, float* __restrict__ logits, float* __restrict__ ws, float* __restrict__ dfeat,
                  T* __restrict__ d_out) {
  const int b = blockIdx.x;
  extern __shared__ float sh[];          // [F] feat, then [MAXC][8] warp partials
  float* sfeat = sh;
  float (*red)[8] = reinterpret_cast<float (*)[8]>(sh + F);
  __shared__ float lg[MAXC], dl[MAXC];
  __shared__ float s_inv;
  pm_pdl_sync();
  for (int f = threadIdx.x; f < F; f += blockDim.x) {
    float acc = 0.f;
    for (int p = 0; p < HW; ++p) acc += to_f<T>(x[((size_t)b * HW + p) * F + f]);
    const float v = acc / (float)HW;       // == gap_fwd_kernel
    sfeat[f] = v;
    feat[(size_t)b * F + f] = v;
  }
  __syncthreads();
  float part[MAXC];
  for (int j = 0; j < ncls; ++j) part[j] = 0.f;
  for (int i = threadIdx.x; i < F; i += blockDim.x) {
    const float v = sfeat[i];
    for (int j = 0; j < ncls; ++j) part[j] = fmaf(v, W[(size_t)j * F + i], part[j]);
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int j = 0; j < ncls; ++j) {
    const float v = warp_sum(part[j]);
    if (lane == 0) red[j][wid] = v;
  }
  __syncthreads();
  if (threadIdx.x < ncls) {
    float v = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += red[threadIdx.x][w];
    v += bias[threadIdx.x];
    lg[threadIdx.x] = v;
    logits[(size_t)b * ncls + threadIdx.x] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float mx = lg[0];
    for (int j = 1; j < ncls; ++j) mx = fmaxf(mx, lg[j]);
    float se = 0.f;
    for (int j = 0; j < ncls; ++j) se += expf(lg[j] - mx);
    const float lse = mx + logf(se);
    float* out = ws + (size_t)b * (ncls + 1);
    float denom = 0.f;
    if (labels) {
      const int y = (int)labels[b];
      const float wgt = cw ? cw[y] : 1.f;
      for (int j = 0; j < ncls; ++j) { dl[j] = wgt * (expf(lg[j] - lse) - (j == y ? 1.f : 0.f)); out[j] = dl[j]; }
      out[ncls] = wgt * (lse - lg[y]);
      if (cw) { for (int q = 0; q < B; ++q) denom += cw[(int)labels[q]]; } else denom = (float)B;
    } else {
      const float* t = soft + (size_t)b * ncls;
      float wsum = 1.f, tsum = 0.f, nll = 0.f;
      if (cw) { wsum = 0.f; for (int j = 0; j < ncls; ++j) wsum += cw[j] * t[j]; }
      for (int j = 0; j < ncls; ++j) { tsum += t[j]; nll += -t[j] * (lg[j] - lse); }
      for (int j = 0; j < ncls; ++j) { dl[j] = wsum * (expf(lg[j] - lse) * tsum - t[j]); out[j] = dl[j]; }
      out[ncls] = wsum * nll;
      denom = (float)B;
    }
    s_inv = 1.f / denom;
    if (b == 0) ws[(size_t)B * (ncls + 1)] = denom;
  }
  __syncthreads();
  const float inv = s_inv;
  for (int f = threadIdx.x; f < F; f += blockDim.x) {
    float df = 0.f;
    for (int j = 0; j < ncls; ++j) df = fmaf(dl[j] * inv, W[(size_t)j * F + f], df);   // == head_bwd_kernel
    dfeat[(size_t)b * F + f] = df;
    const T g = from_f<T>(df / (float)HW);                                            // == gap_bwd_kernel
    for (int p = 0; p < HW; ++p) d_out[((size_t)b * HW + p) * F + f] = g;
  }
}

// dW, db, loss from (ws, feat) -- the batch reductions of the head, off the critical path.  grid (ceil(F/64)), block 64.
__global__ void head_grads_kernel(const float* __restrict__ feat, int B, int F, int ncls, const float* __restrict__ ws,
                                  float* __restrict__ loss, float* __restrict__ dW, float* __restrict__ db) {
  pm_pdl_sync();
  const float inv = 1.f / ws[(size_t)B * (ncls + 1)];
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f < F) {
    float dw[MAXC];
    for (int j = 0; j < ncls; ++j) dw[j] = 0.f;
    for (int b0 = 0; b0 < B; b0 += 8) {
      float xs[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) xs[u] = b0 + u < B ? feat[(size_t)(b0 + u) * F + f] : 0.f;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int b = b0 + u;
        if (b < B) {
          const float* dl = ws + (size_t)b * (ncls + 1);
          for (int j = 0; j < ncls; ++j) dw[j] = fmaf(dl[j] * inv, xs[u], dw[j]);
        }
      }
    }
    for (int j = 0; j < ncls; ++j) dW[(size_t)j * F + f] = dw[j];
  }
  if (blockIdx.x == 0 && threadIdx.x < ncls) {
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += ws[(size_t)b * (ncls + 1) + threadIdx.x];
    db[threadIdx.x] = s * inv;
  }
  if (blockIdx.x == 0 && threadIdx.x == 32) {
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += ws[(size_t)b * (ncls + 1) + ncls];
    loss[0] = s * inv;
  }
}

// FIRST: the optimizer was (re-)created since the last step, i.e. m = v = 0 (the reference does this after every aggregation,
// utils.py:1209-1218): the moments are not read -- identical arithmetic with the zeros folded in -- and nobody has to clear them.
template <bool FIRST>
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, size_t n, float lr, float b1, float b2, float eps, float wd,
                            float bc1, float bc2_sqrt) {
  // torch.optim.Adam (single-tensor path): grad += wd*p ; m = b1*m + (1-b1)*grad ; v = b2*v + (1-b2)*grad^2 ;
  // denom = sqrt(v)/sqrt(bias_correction2) + eps ; p -= (lr/bias_correction1) * m/denom
  const float step = lr / bc1;
  pm_pdl_sync();
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float pi = p[i];
    const float gi = fmaf(wd, pi, g[i]);
    const float m0 = FIRST ? 0.f : m[i], v0 = FIRST ? 0.f : v[i];
    const float mi = m0 + (1.f - b1) * (gi - m0);  // lerp_(grad, 1-beta1)
    const float vi = fmaf(b2, v0, (1.f - b2) * gi * gi);
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = pi - step * (mi / denom);
  }
}

__global__ void sgd_kernel(float* __restrict__ p, const float* __restrict__ g, size_t n, float lr, float wd) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float pi = p[i];
    p[i] = pi - lr * fmaf(wd, pi, g[i]);
  }
}

__global__ void scale_kernel(float* __restrict__ x, float scale, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) x[i] *= scale;
}

__global__ void f32_to_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = __float2bfloat16_rn(x[i]);
}

// NCHW fp32 -> NHWC bf16 with channel padding (Cpad >= C, zero filled)
__global__ void nchw_to_nhwc_bf16_kernel(const float* __restrict__ x, int C, int HW, int Cpad, size_t total,
                                         __nv_bfloat16* __restrict__ out) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % Cpad);
    const size_t t = i / Cpad;
    const size_t hw = t % HW, b = t / HW;
    out[i] = c < C ? __float2bfloat16_rn(x[(b * C + c) * HW + hw]) : __float2bfloat16_rn(0.f);
  }
}

// fp32 KRSC master -> bf16 forward operand [K][R][S][Cpad] and dgrad operand [C][R][S][K]
__global__ void krsc_to_bf16_kernel(const float* __restrict__ w, int K, int C, int RS, int Cpad,
                                    __nv_bfloat16* __restrict__ wf, __nv_bfloat16* __restrict__ wd) {
  const size_t nf = (size_t)K * RS * Cpad;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nf; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % Cpad);
    const size_t t = i / Cpad;
    const int rs = (int)(t % RS);
    const size_t k = t / RS;
    const float v = c < C ? w[(k * RS + rs) * C + c] : 0.f;
    wf[i] = __float2bfloat16_rn(v);
    if (wd && c < C) wd[((size_t)c * RS + rs) * K + k] = __float2bfloat16_rn(v);
  }
}

// Batched weight conversion, one launch for the whole model.  blockIdx.y = conv; blockIdx.x enumerates 32x32 (k, c) tiles
// of every filter tap (early exit past the conv's own tile count).  Each tile is read once from the fp32 master
// (coalesced along c), written to the forward operand [K][RS][Cpad] (coalesced along c) and, transposed through shared
// memory, to the dgrad operand [C][RS][K] (coalesced along k).
__global__ void __launch_bounds__(256)
krsc_to_bf16_batched_kernel(const pm_wcvt_t* __restrict__ table) {
  __shared__ float tile[32][33];
  const pm_wcvt_t e = table[blockIdx.y];
  const int tk = (e.K + 31) / 32, tc = (e.Cpad + 31) / 32;
  const int ntiles = e.RS * tk * tc;
  if ((int)blockIdx.x >= ntiles) return;
  const int rs = blockIdx.x / (tk * tc);
  const int rem = blockIdx.x - rs * tk * tc;
  const int k0 = (rem / tc) * 32, c0 = (rem % tc) * 32;
  __nv_bfloat16* wf = (__nv_bfloat16*)e.w_fwd;
  __nv_bfloat16* wd = (__nv_bfloat16*)e.w_dgrad;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    const int k = k0 + i, c = c0 + tx;
    float v = 0.f;
    if (k < e.K && c < e.C) v = e.w[((size_t)k * e.RS + rs) * e.C + c];
    tile[i][tx] = v;
    if (k < e.K && c < e.Cpad) wf[((size_t)k * e.RS + rs) * e.Cpad + c] = __float2bfloat16_rn(v);
  }
  if (wd) {
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
      const int c = c0 + i, k = k0 + tx;
      if (c < e.C && k < e.K) wd[((size_t)c * e.RS + rs) * e.K + k] = __float2bfloat16_rn(tile[tx][i]);
    }
  }
}

// Same conversion with an EXACT grid: blockIdx.x enumerates the tiles of all entries back to back (the rectangular
// (max_tiles, n) grid above launches ~4x more blocks than there are tiles, and retiring empty blocks costs more than the
// conversion itself).  Warp 0 finds the entry with a ballot over the per-entry tile counts.
__global__ void __launch_bounds__(256)
krsc_to_bf16_exact_kernel(const pm_wcvt_t* __restrict__ table, int n) {
  __shared__ float tile[32][33];
  __shared__ int s_entry, s_local;
  pm_pdl_sync();
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    int base = 0, found = 0;
    for (int e0 = 0; e0 < n && !found; e0 += 32) {
      int nt = 0;
      if (e0 + lane < n) {
        const pm_wcvt_t e = table[e0 + lane];
        nt = e.RS * ((e.K + 31) / 32) * ((e.Cpad + 31) / 32);
      }
      int incl = nt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      const unsigned hit = __ballot_sync(0xffffffffu, base + incl > (int)blockIdx.x);
      if (hit) {
        const int l = __ffs(hit) - 1;
        const int excl = __shfl_sync(0xffffffffu, incl - nt, l);
        if (lane == 0) { s_entry = e0 + l; s_local = (int)blockIdx.x - base - excl; }
        found = 1;
      }
      base += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (!found && lane == 0) s_entry = -1;
  }
  __syncthreads();
  if (s_entry < 0) return;
  const pm_wcvt_t e = table[s_entry];
  const int bx = s_local;
  const int tk = (e.K + 31) / 32, tc = (e.Cpad + 31) / 32;
  const int rs = bx / (tk * tc);
  const int rem = bx - rs * tk * tc;
  const int k0 = (rem / tc) * 32, c0 = (rem % tc) * 32;
  __nv_bfloat16* wf = (__nv_bfloat16*)e.w_fwd;
  __nv_bfloat16* wd = (__nv_bfloat16*)e.w_dgrad;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    const int k = k0 + i, c = c0 + tx;
    float v = 0.f;
    if (k < e.K && c < e.C) v = e.w[((size_t)k * e.RS + rs) * e.C + c];
    tile[i][tx] = v;
    if (k < e.K && c < e.Cpad) wf[((size_t)k * e.RS + rs) * e.Cpad + c] = __float2bfloat16_rn(v);
  }
  if (wd) {
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
      const int c = c0 + i, k = k0 + tx;
      if (c < e.C && k < e.K) wd[((size_t)c * e.RS + rs) * e.K + k] = __float2bfloat16_rn(tile[tx][i]);
    }
  }
}

// one block per (image, output row): the R input rows it needs are staged in shared memory (coalesced), then every
// thread assembles 16-byte chunks (8 consecutive k = (r*S+s)*Cin + c) of the im2col rows.  blockDim = 10 pixels x
// (Kpad/8) chunks: a thread always produces the same chunk index, so its 8 smem offsets live in registers.
__global__ void __launch_bounds__(256)
im2col_stem_kernel(const float* __restrict__ x, int Cin, int H, int W, int R, int stride, int pad, int Ho, int Wo, int Kpad,
                   __nv_bfloat16* __restrict__ out) {
  extern __shared__ float srow[];  // [Cin*R][W + 2*pad]
  const int WP = W + 2 * pad;
  const int oh = blockIdx.x, b = blockIdx.y;
  const int ih0 = oh * stride - pad;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int cr = warp; cr < Cin * R; cr += nwarps) {
    const int c = cr / R, r = cr - c * R;
    const int ih = ih0 + r;
    const bool row_ok = ih >= 0 && ih < H;
    const float* src = x + (((size_t)b * Cin + c) * H + (row_ok ? ih : 0)) * W;
    for (int wp = lane; wp < WP; wp += 32) {
      const int iw = wp - pad;
      srow[cr * WP + wp] = (row_ok && iw >= 0 && iw < W) ? src[iw] : 0.f;
    }
  }
  const int chunks = Kpad / 8;
  const int Ktrue = R * R * Cin;
  const int j = threadIdx.x % chunks, p0 = threadIdx.x / chunks, ppb = blockDim.x / chunks;
  int off[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int k = j * 8 + e;
    off[e] = -1;
    if (k < Ktrue) {
      const int c = k % Cin, rs = k / Cin;
      const int s_ = rs % R, r = rs / R;
      off[e] = (c * R + r) * WP + s_;
    }
  }
  __syncthreads();
  __nv_bfloat16* orow = out + ((size_t)b * Ho + oh) * Wo * Kpad;
  if (p0 < ppb) {
    for (int ow = p0; ow < Wo; ow += ppb) {
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = off[e] >= 0 ? srow[off[e] + ow * stride] : 0.f;
      uint4 pk;
      __nv_bfloat162* ph = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
      for (int e = 0; e < 4; ++e) ph[e] = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
      *reinterpret_cast<uint4*>(orow + ((size_t)ow * chunks + j) * 8) = pk;
    }
  }
}

}  // namespace

template <typename T>
static int head_fused_t(const T* x, int HW, const float* W, const float* bias, const int64_t* labels, const float* soft,
                        const float* class_w, int B, int F, int ncls, float* feat, float* logits, float* ws, float* dfeat, T* d_out,
                        pm_stream_t s) {
  PM_CHECK_ARG(x && W && bias && feat && logits && ws && dfeat && d_out && B > 0 && F > 0 && HW > 0 && ncls > 0 && ncls <= MAXC);
  PM_CHECK_ARG((labels != nullptr) != (soft != nullptr));
  const size_t smem = ((size_t)F + MAXC * 8) * sizeof(float);
  PM_CUDA(pm_launch(head_fused_kernel<T>, dim3(B), dim3(256), smem, S(s), x, HW, W, bias, labels, soft, class_w, B, F, ncls, feat, logits, ws,
                    dfeat, d_out));
  PM_LAUNCH_OK();
}

extern "C" {

int pm_linear_ce_f32(const float* feat, const float* W, const float* bias, const int64_t* labels, const float* soft,
                     const float* class_w, int B, int F, int ncls, float* logits, float* loss, float* dfeat,
                     float* dW, float* db, float* ws, pm_stream_t s) {
  PM_CHECK_ARG(feat && W && bias && logits && loss && dfeat && dW && db && ws && B > 0 && F > 0 && ncls > 0 && ncls <= MAXC);
  PM_CHECK_ARG((labels != nullptr) != (soft != nullptr));
  PM_CUDA(cudaMemsetAsync(ws + (size_t)B * (ncls + 1), 0, sizeof(float), S(s)));
  head_fwd_kernel<<<B, 128, 0, S(s)>>>(feat, W, bias, labels, soft, class_w, F, ncls, logits, ws);
  head_bwd_kernel<<<(F + 63) / 64, 64, 0, S(s)>>>(feat, W, B, F, ncls, ws, loss, dfeat, dW, db);  // thread 32 of block 0 writes the loss
  PM_LAUNCH_OK();
}

int pm_head_fused_f32(const float* x, int HW, const float* W, const float* bias, const int64_t* labels, const float* soft,
                      const float* class_w, int B, int F, int ncls, float* feat, float* logits, float* ws, float* dfeat, float* d_out,
                      pm_stream_t s) {
  return head_fused_t<float>(x, HW, W, bias, labels, soft, class_w, B, F, ncls, feat, logits, ws, dfeat, d_out, s);
}
int pm_head_fused_bf16(const void* x, int HW, const float* W, const float* bias, const int64_t* labels, const float* soft,
                       const float* class_w, int B, int F, int ncls, float* feat, float* logits, float* ws, float* dfeat, void* d_out,
                       pm_stream_t s) {
  return head_fused_t<__nv_bfloat16>((const __nv_bfloat16*)x, HW, W, bias, labels, soft, class_w, B, F, ncls, feat, logits, ws, dfeat,
                                     (__nv_bfloat16*)d_out, s);
}
int pm_head_grads_f32(const float* feat, int B, int F, int ncls, const float* ws, float* loss, float* dW, float* db, pm_stream_t s) {
  PM_CHECK_ARG(feat && ws && loss && dW && db && B > 0 && F > 0 && ncls > 0 && ncls <= MAXC);
  head_grads_kernel<<<(F + 63) / 64, 64, 0, S(s)>>>(feat, B, F, ncls, ws, loss, dW, db);
  PM_LAUNCH_OK();
}

int pm_linear_fwd_f32(const float* feat, const float* W, const float* bias, int B, int F, int ncls, float* logits,
                      pm_stream_t s) {
  PM_CHECK_ARG(feat && W && bias && logits && B > 0 && F > 0 && ncls > 0 && ncls <= MAXC);
  head_fwd_kernel<<<B, 128, 0, S(s)>>>(feat, W, bias, nullptr, nullptr, nullptr, F, ncls, logits, nullptr);
  PM_LAUNCH_OK();
}

static int adam_launch(bool first, float* p, const float* g, float* m, float* v, size_t n, float lr, float beta1, float beta2,
                       float eps, float weight_decay, int step, pm_stream_t s) {
  PM_CHECK_ARG(p && g && m && v && step >= 1);
  if (n == 0) return PM_OK;
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  if (first)
    PM_CUDA(pm_launch(adam_kernel<true>, dim3(pm_grid(n, 256, 1, 16)), dim3(256), 0, S(s), p, g, m, v, n, lr, beta1, beta2, eps,
                      weight_decay, (float)bc1, (float)sqrt(bc2)));
  else
    PM_CUDA(pm_launch(adam_kernel<false>, dim3(pm_grid(n, 256, 1, 16)), dim3(256), 0, S(s), p, g, m, v, n, lr, beta1, beta2, eps,
                      weight_decay, (float)bc1, (float)sqrt(bc2)));
  PM_LAUNCH_OK();
}
int pm_adam_step_f32(float* p, const float* g, float* m, float* v, size_t n, float lr, float beta1, float beta2,
                     float eps, float weight_decay, int step, pm_stream_t s) {
  return adam_launch(false, p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, step, s);
}
int pm_adam_first_step_f32(float* p, const float* g, float* m, float* v, size_t n, float lr, float beta1, float beta2,
                           float eps, float weight_decay, pm_stream_t s) {
  return adam_launch(true, p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, 1, s);
}

int pm_sgd_step_f32(float* p, const float* g, size_t n, float lr, float weight_decay, pm_stream_t s) {
  PM_CHECK_ARG(p && g);
  if (n == 0) return PM_OK;
  sgd_kernel<<<pm_grid(n, 256, 1, 16), 256, 0, S(s)>>>(p, g, n, lr, weight_decay);
  PM_LAUNCH_OK();
}

int pm_scale_f32(float* x, float scale, size_t n, pm_stream_t s) {
  PM_CHECK_ARG(x);
  if (n == 0) return PM_OK;
  scale_kernel<<<pm_grid(n, 256, 1, 16), 256, 0, S(s)>>>(x, scale, n);
  PM_LAUNCH_OK();
}

int pm_f32_to_bf16(const float* x, void* out, size_t n, pm_stream_t s) {
  PM_CHECK_ARG(x && out);
  if (n == 0) return PM_OK;
  f32_to_bf16_kernel<<<pm_grid(n, 256, 1, 16), 256, 0, S(s)>>>(x, (__nv_bfloat16*)out, n);
  PM_LAUNCH_OK();
}

int pm_nchw_to_nhwc_f32_bf16(const float* x, int B, int C, int H, int W, int Cpad, void* out, pm_stream_t s) {
  PM_CHECK_ARG(x && out && Cpad >= C);
  const size_t total = (size_t)B * H * W * Cpad;
  nchw_to_nhwc_bf16_kernel<<<pm_grid(total, 256, 1, 16), 256, 0, S(s)>>>(x, C, H * W, Cpad, total, (__nv_bfloat16*)out);
  PM_LAUNCH_OK();
}

int pm_krsc_to_bf16_fwd_dgrad(const float* w, int K, int C, int R, int S_, int Cpad, void* w_fwd, void* w_dgrad,
                              pm_stream_t s) {
  PM_CHECK_ARG(w && w_fwd && Cpad >= C);
  const size_t total = (size_t)K * R * S_ * Cpad;
  krsc_to_bf16_kernel<<<pm_grid(total, 256, 1, 16), 256, 0, S(s)>>>(w, K, C, R * S_, Cpad, (__nv_bfloat16*)w_fwd,
                                                                    (__nv_bfloat16*)w_dgrad);
  PM_LAUNCH_OK();
}

int pm_krsc_to_bf16_batched(const pm_wcvt_t* table, int n, int max_tiles, pm_stream_t s) {
  PM_CHECK_ARG(table && n > 0 && n <= 65535 && max_tiles > 0);
  dim3 grid(max_tiles, n);  // max_tiles = max over entries of RS * ceil(K/32) * ceil(Cpad/32)
  krsc_to_bf16_batched_kernel<<<grid, 256, 0, S(s)>>>(table);
  PM_LAUNCH_OK();
}

int pm_krsc_to_bf16_batched_exact(const pm_wcvt_t* table, int n, int total_tiles, pm_stream_t s) {
  PM_CHECK_ARG(table && n > 0 && total_tiles > 0);
  PM_CUDA(pm_launch(krsc_to_bf16_exact_kernel, dim3(total_tiles), dim3(256), 0, S(s), table, n));
  PM_LAUNCH_OK();
}

int pm_im2col_stem_bf16(const float* x, int B, int Cin, int H, int W, int R, int stride, int pad, int Kpad, void* out,
                        pm_stream_t s) {
  PM_CHECK_ARG(x && out && B > 0 && B <= 65535 && Kpad % 8 == 0 && Kpad >= R * R * Cin);
  const int Ho = (H + 2 * pad - R) / stride + 1, Wo = (W + 2 * pad - R) / stride + 1;
  const size_t smem = (size_t)Cin * R * (W + 2 * pad) * sizeof(float);
  PM_CHECK_ARG(smem <= 48 * 1024);
  dim3 grid(Ho, B);
  const int chunks = Kpad / 8;
  PM_CHECK_ARG(chunks <= 256);
  const int threads = (256 / chunks) * chunks;  // whole pixels per pass
  im2col_stem_kernel<<<grid, threads, smem, S(s)>>>(x, Cin, H, W, R, stride, pad, Ho, Wo, Kpad, (__nv_bfloat16*)out);
  PM_LAUNCH_OK();
}

}  // extern "C"
