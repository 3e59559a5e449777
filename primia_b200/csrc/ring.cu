// Path E: SPDZ additive-secret-shared fixed-precision arithmetic in the ring Z_2^64 (two's-complement
// int64 wraparound == the reference's native torch.int64 behaviour).  All kernels are exact integer
// arithmetic: results are bit-identical to the CPU oracle for identical inputs.
//
// The dominant kernel is ring_gemm2_kernel: C (+)= A1@B1 + A2@B2 over int64 on the integer pipe
// (no int64 MMA exists; 64-bit MAC = IMAD.WIDE.U32 + 2 IMAD).  Integer addition is associative, so
// split-K with 64-bit atomics stays bit-exact.
#include "common.cuh"

typedef unsigned long long u64;

// ------------------------------------------------------------------------------------------------
// elementwise helpers
// ------------------------------------------------------------------------------------------------
__global__ void encode_kernel(const float* __restrict__ x, float scale, int64_t* __restrict__ q, size_t n,
                              int* __restrict__ overflow) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  bool bad = false;
  for (; i < n; i += stride) {
    float v = x[i] * scale;  // fp32 multiply, as torch does for float_tensor * python_scalar (precision.py:121)
    // .long(): truncation toward zero; out-of-range is what the reference's assert rejects
    if (!(fabsf(v) < 9223372036854775808.0f)) bad = true;
    q[i] = (int64_t)v;
  }
  if (bad && overflow) atomicExch(overflow, 1);
}

__global__ void decode_kernel(const int64_t* __restrict__ q, float scale, float* __restrict__ x, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) x[i] = (float)q[i] / scale;  // .float() is round-to-nearest-even like __ll2float_rn
}

// Philox4x32-10 (Salmon et al. 2011) written out; counter = (idx, offset), key = seed.
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t out[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__device__ __forceinline__ int64_t clamp_share_range(u64 r) {
  // reference range is [-2^63, 2^63-2] (random_(min_value, max_value) / randint(low, high) exclude the top value)
  int64_t v = (int64_t)r;
  return v == INT64_MAX ? (int64_t)0 : v;
}

__global__ void epoch_bump_kernel(unsigned long long* e) { *e += 1ull; }

// `epoch` (device uint64, may be NULL): added to the high half of the Philox offset, so that a CAPTURED generation graph draws a
// fresh stream at every replay (the host-side offset is baked into the launch; the epoch is bumped by a node of the graph).
template <bool SHARE>
__global__ void philox_kernel(const int64_t* __restrict__ q, u64 seed, u64 offset, const unsigned long long* __restrict__ epoch,
                              int64_t* __restrict__ s0, int64_t* __restrict__ s1, size_t n) {
  if (epoch) offset += (u64)(*epoch) << 32;
  size_t pairs = (n + 1) / 2;
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < pairs; i += stride) {
    uint32_t o[4];
    philox4x32_10((uint32_t)i, (uint32_t)(i >> 32), (uint32_t)offset, (uint32_t)(offset >> 32), (uint32_t)seed,
                  (uint32_t)(seed >> 32), o);
    int64_t r0 = clamp_share_range(((u64)o[1] << 32) | o[0]);
    int64_t r1 = clamp_share_range(((u64)o[3] << 32) | o[2]);
    size_t e = 2 * i;
    s0[e] = r0;
    if (SHARE) s1[e] = (int64_t)((u64)q[e] - (u64)r0);
    if (e + 1 < n) {
      s0[e + 1] = r1;
      if (SHARE) s1[e + 1] = (int64_t)((u64)q[e + 1] - (u64)r1);
    }
  }
}

__global__ void sub_kernel(const int64_t* __restrict__ x, const int64_t* __restrict__ a, int64_t* __restrict__ d,
                           size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) d[i] = (int64_t)((u64)x[i] - (u64)a[i]);
}

__global__ void add_kernel(const int64_t* __restrict__ x, const int64_t* __restrict__ y, int64_t* __restrict__ d,
                           size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) d[i] = (int64_t)((u64)x[i] + (u64)y[i]);
}

__global__ void axpby_kernel(int64_t alpha, const int64_t* __restrict__ x, int64_t beta,
                             const int64_t* __restrict__ y, int ybcast, size_t n, size_t C,
                             int64_t* __restrict__ out) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    u64 v = (u64)alpha * (u64)x[i];
    if (y) {
      size_t yi = ybcast == 0 ? i : (ybcast == 1 ? i % C : 0);
      v += (u64)beta * (u64)y[yi];
    }
    out[i] = (int64_t)v;
  }
}

// ---- one elementwise pass of a HOISTED Beaver product with a per-channel operand, NCHW (ring/functional.py batch_norm):
//   out[i] = T( sc[c] * (u[i] + peer[i]) + add[i] ) + cs * chan[c] + es * elem[i],   c = (i / HW) % C
// with T = C-style division by `div` (div > 1) or the identity.  NULL operands drop out (sc -> 1).  `peer` is the other party's
// masked share (the opening, spdz.py:162-163; possibly a peer-mapped pointer), everything else is local.  All arithmetic wraps
// mod 2^64 exactly as the share-level torch ops of the reference do.
__global__ void spdz_affine_kernel(const int64_t* __restrict__ u, const int64_t* __restrict__ peer, const int64_t* __restrict__ sc,
                                   const int64_t* __restrict__ add, int64_t div, const int64_t* __restrict__ chan, int cs,
                                   const int64_t* __restrict__ elem, int es, int C, int HW, size_t n, int64_t* __restrict__ out) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    const int c = (int)((i / (size_t)HW) % (size_t)C);
    u64 v = (u64)u[i];
    if (peer) v += (u64)peer[i];
    if (sc) v *= (u64)sc[c];
    if (add) v += (u64)add[i];
    int64_t t = (int64_t)v;
    if (div > 1) t = t / div;
    u64 r = (u64)t;
    if (chan) r += (u64)(int64_t)cs * (u64)chan[c];
    if (elem) r += (u64)(int64_t)es * (u64)elem[i];
    out[i] = (int64_t)r;
  }
}

__global__ void trunc_div_kernel(const int64_t* __restrict__ x, int64_t div, int64_t* __restrict__ out, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = x[i] / div;  // C++ '/' on signed ints truncates toward zero
}

// z [B,M,N] -> out [B,N,M] with truncation (+bias[n])
__global__ void trunc_post_conv_kernel(const int64_t* __restrict__ z, int64_t div, const int64_t* __restrict__ bias,
                                       int M, int N, int64_t* __restrict__ out) {
  __shared__ int64_t tile[32][33];
  int b = blockIdx.z;
  int m0 = blockIdx.x * 32, n0 = blockIdx.y * 32;
  const int64_t* zb = z + (size_t)b * M * N;
  int64_t* ob = out + (size_t)b * M * N;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    int m = m0 + r, n = n0 + threadIdx.x;
    if (m < M && n < N) {
      int64_t v = zb[(size_t)m * N + n];
      if (div != 1) v = v / div;
      if (bias) v = (int64_t)((u64)v + (u64)bias[n]);
      tile[r][threadIdx.x] = v;
    }
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    int n = n0 + r, m = m0 + threadIdx.x;
    if (m < M && n < N) ob[(size_t)n * M + m] = tile[threadIdx.x][r];
  }
}

// generic [B, R, Ccols] -> [B, Ccols, R] transpose of int64 (used for NCHW <-> [P,C] shuffles)
__global__ void transpose_i64_kernel(const int64_t* __restrict__ x, int R, int Cc, int64_t* __restrict__ out) {
  __shared__ int64_t tile[32][33];
  int b = blockIdx.z;
  int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int64_t* xb = x + (size_t)b * R * Cc;
  int64_t* ob = out + (size_t)b * R * Cc;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    int rr = r0 + r, cc = c0 + threadIdx.x;
    if (rr < R && cc < Cc) tile[r][threadIdx.x] = xb[(size_t)rr * Cc + cc];
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    int cc = c0 + r, rr = r0 + threadIdx.x;
    if (rr < R && cc < Cc) ob[(size_t)cc * R + rr] = tile[threadIdx.x][r];
  }
}

// ------------------------------------------------------------------------------------------------
// im2col (reference order k = ch*kh*kw + r*kw + c ; functional.py:129-149), optionally fused with "- a"
// ------------------------------------------------------------------------------------------------
template <bool MASK>
__global__ void im2col_kernel(const int64_t* __restrict__ x, int C, int H, int W, int kh, int kw, int stride,
                              int pad, int dil, int Ho, int Wo, const int64_t* __restrict__ a,
                              int64_t* __restrict__ out, size_t total) {
  const int K = C * kh * kw;
  const int M = Ho * Wo;
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t gstride = (size_t)gridDim.x * blockDim.x;
  for (; i < total; i += gstride) {
    int k = (int)(i % K);
    size_t bm = i / K;
    int m = (int)(bm % M);
    int b = (int)(bm / M);
    int cc = k % kw;
    int r = (k / kw) % kh;
    int ch = k / (kw * kh);
    int oh = m / Wo, ow = m % Wo;
    int ih = oh * stride - pad + r * dil;
    int iw = ow * stride - pad + cc * dil;
    int64_t v = 0;
    if (ih >= 0 && ih < H && iw >= 0 && iw < W) v = x[(((size_t)b * C + ch) * H + ih) * W + iw];
    if (MASK) v = (int64_t)((u64)v - (u64)a[i]);
    out[i] = v;
  }
}

// eps[k,n] = w[n*K+k] - b[k*N+n]
__global__ void mask_wt_kernel(const int64_t* __restrict__ w, int N, int K, const int64_t* __restrict__ b,
                               int64_t* __restrict__ eps) {
  __shared__ int64_t tile[32][33];
  int n0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    int n = n0 + r, k = k0 + threadIdx.x;
    if (n < N && k < K) tile[r][threadIdx.x] = w[(size_t)n * K + k];
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    int k = k0 + r, n = n0 + threadIdx.x;
    if (n < N && k < K) {
      size_t o = (size_t)k * N + n;
      eps[o] = (int64_t)((u64)tile[threadIdx.x][r] - (u64)b[o]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// int64 GEMM with two (A,B) segments:  C[b] (+)= Cinit[b] + A1[b]@B1 + A2[b]@B2
// ------------------------------------------------------------------------------------------------
constexpr int GBM = 64, GBN = 64, GBK = 16, GTHREADS = 256;
constexpr int GPAD = 2;

__global__ void __launch_bounds__(GTHREADS, 2)
ring_gemm2_kernel(const u64* __restrict__ A1, const u64* __restrict__ B1, const u64* __restrict__ A2,
                  const u64* __restrict__ B2, const u64* __restrict__ Cinit, u64* __restrict__ Cout, int M, int K,
                  int N, int splitK) {
  __shared__ __align__(16) u64 As[2][GBK][GBM + GPAD];
  __shared__ __align__(16) u64 Bs[2][GBK][GBN + GPAD];

  const int tid = threadIdx.x;
  const int n0 = blockIdx.x * GBN, m0 = blockIdx.y * GBM;
  const int batch = blockIdx.z / splitK, split = blockIdx.z % splitK;
  const size_t a_off = (size_t)batch * M * K, c_off = (size_t)batch * M * N;

  const int nk = (K + GBK - 1) / GBK;             // chunks per segment
  const int nseg = A2 ? 2 : 1;
  const int total = nk * nseg;
  const int per = (total + splitK - 1) / splitK;
  const int c_begin = split * per;
  const int c_end = min(total, c_begin + per);

  // loader mapping
  const int a_k = tid & 15, a_m = tid >> 4;       // A: 16 k x 16 m per pass, 4 passes
  const int b_n = tid & 63, b_k = tid >> 6;       // B: 64 n x 4 k per pass, 4 passes
  // compute mapping
  const int tx = tid & 15, ty = tid >> 4;

  // 64-bit MAC mod 2^64 in three integer-pipe instructions: lo(a)*lo(b) is accumulated at full width with
  // IMAD.WIDE.U32, the two cross terms lo*hi + hi*lo only matter mod 2^32 and go to a separate 32-bit accumulator that
  // is folded in (<< 32) once at the end; hi*hi vanishes mod 2^64.
  u64 acc[4][4];
  uint32_t acch[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { acc[i][j] = 0; acch[i][j] = 0; }

  u64 ra[4], rb[4];
  auto gload = [&](int chunk) {
    const int seg = chunk / nk, kc = chunk - seg * nk;
    const u64* A = (seg == 0 ? A1 : A2) + a_off;
    const u64* Bm = seg == 0 ? B1 : B2;
    const int k = kc * GBK + a_k;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = m0 + a_m + 16 * i;
      ra[i] = (m < M && k < K) ? A[(size_t)m * K + k] : 0ull;
    }
    const int n = n0 + b_n;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int kk = kc * GBK + b_k + 4 * i;
      rb[i] = (n < N && kk < K) ? Bm[(size_t)kk * N + n] : 0ull;
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 4; ++i) As[buf][a_k][a_m + 16 * i] = ra[i];
#pragma unroll
    for (int i = 0; i < 4; ++i) Bs[buf][b_k + 4 * i][b_n] = rb[i];
  };

  if (c_begin < c_end) {
    gload(c_begin);
    sstore(0);
    __syncthreads();
    int buf = 0;
    for (int chunk = c_begin; chunk < c_end; ++chunk) {
      const bool has_next = chunk + 1 < c_end;
      if (has_next) gload(chunk + 1);
#pragma unroll
      for (int k = 0; k < GBK; ++k) {
        u64 a[4], b[4];
        const ulonglong2 a01 = *reinterpret_cast<const ulonglong2*>(&As[buf][k][ty * 4]);
        const ulonglong2 a23 = *reinterpret_cast<const ulonglong2*>(&As[buf][k][ty * 4 + 2]);
        const ulonglong2 b01 = *reinterpret_cast<const ulonglong2*>(&Bs[buf][k][tx * 4]);
        const ulonglong2 b23 = *reinterpret_cast<const ulonglong2*>(&Bs[buf][k][tx * 4 + 2]);
        a[0] = a01.x; a[1] = a01.y; a[2] = a23.x; a[3] = a23.y;
        b[0] = b01.x; b[1] = b01.y; b[2] = b23.x; b[3] = b23.y;
        uint32_t alo[4], ahi[4], blo[4], bhi[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          alo[i] = (uint32_t)a[i]; ahi[i] = (uint32_t)(a[i] >> 32);
          blo[i] = (uint32_t)b[i]; bhi[i] = (uint32_t)(b[i] >> 32);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i][j]) : "r"(alo[i]), "r"(blo[j]));
            acch[i][j] += alo[i] * bhi[j];
            acch[i][j] += ahi[i] * blo[j];
          }
      }
      if (has_next) {
        sstore(buf ^ 1);
        __syncthreads();
        buf ^= 1;
      }
    }
  }

  const bool atomic = splitK > 1;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      const size_t o = c_off + (size_t)m * N + n;
      u64 v = acc[i][j] + ((u64)acch[i][j] << 32);
      if (split == 0 && Cinit) v += Cinit[o];
      if (atomic) atomicAdd(&Cout[o], v);
      else Cout[o] = v;
    }
  }
}

static int launch_gemm2(const int64_t* A1, const int64_t* B1, const int64_t* A2, const int64_t* B2,
                        const int64_t* Cinit, int64_t* C, int B, int M, int K, int N, cudaStream_t st) {
  const int tm = (M + GBM - 1) / GBM, tn = (N + GBN - 1) / GBN;
  const int nk = (K + GBK - 1) / GBK;
  const int total = nk * (A2 ? 2 : 1);
  const long tiles = (long)tm * tn * B;
  int splitK = 1;
  const long target = 2L * pm_num_sms() * 2;  // 2 CTAs/SM resident, aim for >= 2 waves
  if (tiles < target) {
    splitK = (int)((target + tiles - 1) / tiles);
    int max_split = total / 4;  // keep >= 4 chunks (64 k-steps) per split
    if (max_split < 1) max_split = 1;
    if (splitK > max_split) splitK = max_split;
  }
  if ((long)B * splitK > 65535) return PM_EINVAL;
  if (splitK > 1) {
    cudaError_t e = cudaMemsetAsync(C, 0, (size_t)B * M * N * sizeof(int64_t), st);
    if (e != cudaSuccess) return pm_set_err(__FILE__, __LINE__, cudaGetErrorString(e));
  }
  dim3 grid(tn, tm, B * splitK);
  ring_gemm2_kernel<<<grid, GTHREADS, 0, st>>>((const u64*)A1, (const u64*)B1, (const u64*)A2, (const u64*)B2,
                                               (const u64*)Cinit, (u64*)C, M, K, N, splitK);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return pm_set_err(__FILE__, __LINE__, cudaGetErrorString(e));
  return PM_OK;
}

int pm_set_err(const char* file, int line, const char* msg) {
  snprintf(g_pm_err, sizeof(g_pm_err), "%s:%d: %s", file, line, msg);
  return PM_ECUDA;
}

// elementwise Beaver combine with broadcasting
//   mode 0: all [n] ; mode 1: delta,a [C] vs eps,b,c,z [P,C] ; mode 2: delta,a,c,z [P,C] vs eps,b [C]
__global__ void combine_mul_kernel(int j, const u64* __restrict__ delta, const u64* __restrict__ eps,
                                   const u64* __restrict__ a, const u64* __restrict__ b, const u64* __restrict__ c,
                                   int mode, size_t n, size_t C, u64* __restrict__ z) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    const size_t il = mode == 1 ? i % C : i;
    const size_t ir = mode == 2 ? i % C : i;
    const u64 d = delta[il], e = eps[ir];
    u64 v = d * b[ir] + a[il] * e + c[i];
    if (j == 0) v += d * e;
    z[i] = v;
  }
}

// spdz_mask of BOTH operands of an elementwise product in one pass (spdz.py:22-45): delta_j = x_j - a_j, eps_j = y_j - b_j
__global__ void mask2_kernel(const u64* __restrict__ x, const u64* __restrict__ a, const u64* __restrict__ y, const u64* __restrict__ b,
                             u64* __restrict__ d, u64* __restrict__ e, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) { d[i] = x[i] - a[i]; e[i] = y[i] - b[i]; }
}

// spdz_compute for same-shape operands with the two openings fused (spdz.py:64-122,162-163): delta = d_own + d_peer,
// eps = e_own + e_peer are formed in registers (the peer pointers may be peer-mapped), never stored
__global__ void combine_mul_open_kernel(int j, const u64* __restrict__ d_own, const u64* __restrict__ d_peer,
                                        const u64* __restrict__ e_own, const u64* __restrict__ e_peer, const u64* __restrict__ a,
                                        const u64* __restrict__ b, const u64* __restrict__ c, size_t n, u64* __restrict__ z) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    const u64 d = d_own[i] + d_peer[i], e = e_own[i] + e_peer[i];
    u64 v = d * b[i] + a[i] * e + c[i];
    if (j == 0) v += d * e;
    z[i] = v;
  }
}

// exact C-style (truncating) int64 division by a positive invariant divisor without the ~100-instruction emulated
// 64-bit divide: two double-precision quotient estimates, each followed by an exact remainder in wrapping integer
// arithmetic, then at most two unit corrections.  inv_d = 1.0 / d.
__device__ __forceinline__ int64_t trunc_div_inv(int64_t x, int64_t d, double inv_d) {
  int64_t q = __double2ll_rz((double)x * inv_d);
  int64_t r = (int64_t)((u64)x - (u64)q * (u64)d);           // |r| <= ~2^11 * d: exact despite the wrap
  const int64_t q2 = __double2ll_rz((double)r * inv_d);
  q += q2;
  r -= q2 * d;                                                // now |r| < 2d
  if (x >= 0) {
    if (r < 0) { --q; r += d; }
    if (r < 0) { --q; r += d; }
    if (r >= d) { ++q; r -= d; }
    if (r >= d) { ++q; }
  } else {
    if (r > 0) { ++q; r -= d; }
    if (r > 0) { ++q; r -= d; }
    if (r <= -d) { --q; r += d; }
    if (r <= -d) { --q; }
  }
  return q;
}

// reciprocal(method="newton") (precision.py:507-518) run to completion for [C] vectors when both share holders live on
// the same device: thread = channel, both parties' shares in registers, the openings of the 3*(iters-1) Beaver products
// happen in registers.  Arithmetic per party is exactly spdz_mask / spdz_compute / truncate / __rsub__ / __truediv__.
// a*,b*,c* : [3*(iters-1)][C] triple shares in consumption order (x*x, v*(xx), y*x per iteration); k* : [iters] shares of
// the public constant (C+1) encoded (additive_shared.py:473-487).  blockIdx.y selects the job (one BatchNorm layer each):
// the inverse square roots depend on the model only, so all layers of a forward pass are issued as one launch.
struct NewtonJobs {
  pm_newton_job_t job[PM_NEWTON_MAX_JOBS];
};

__global__ void __launch_bounds__(64)
bn_newton_fused_kernel(const __grid_constant__ NewtonJobs jobs, int iters, int64_t div, int64_t Cc) {
  const pm_newton_job_t& J = jobs.job[blockIdx.y];
  const int C = J.C;
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= C) return;
  const u64 *a0 = (const u64*)J.a0, *b0 = (const u64*)J.b0, *c0 = (const u64*)J.c0;
  const u64 *a1 = (const u64*)J.a1, *b1 = (const u64*)J.b1, *c1 = (const u64*)J.c1;
  const u64 *k0 = (const u64*)J.k0, *k1 = (const u64*)J.k1;
  const double inv_div = 1.0 / (double)div, inv_c = 1.0 / (double)Cc;
  const u64 V0 = (u64)J.v0[ch], V1 = (u64)J.v1[ch];
  u64 X0 = (u64)trunc_div_inv((int64_t)(0 - (V0 - k0[0])), Cc, inv_c);
  u64 X1 = (u64)trunc_div_inv((int64_t)(0 - (V1 - k1[0])), Cc, inv_c);
  // operands of one iteration: 3 products x (a0,b0,c0,a1,b1,c1); loaded one iteration ahead (their addresses do not
  // depend on the data), so the dependent chain below never waits on memory
  u64 T[3][6], Tn[3][6], K0n = 0, K1n = 0;
  auto load = [&](int it, u64 (&t)[3][6], u64& kk0, u64& kk1) {
#pragma unroll
    for (int m = 0; m < 3; ++m) {
      const size_t o = (size_t)(3 * (it - 1) + m) * C + ch;
      t[m][0] = a0[o]; t[m][1] = b0[o]; t[m][2] = c0[o]; t[m][3] = a1[o]; t[m][4] = b1[o]; t[m][5] = c1[o];
    }
    kk0 = k0[it]; kk1 = k1[it];
  };
  auto beaver = [&](u64 p0, u64 p1, u64 q0, u64 q1, const u64 (&t)[6], u64& z0, u64& z1) {
    const u64 d = (p0 - t[0]) + (p1 - t[3]), e = (q0 - t[1]) + (q1 - t[4]);
    z0 = (u64)trunc_div_inv((int64_t)(d * t[1] + t[0] * e + t[2] + d * e), div, inv_div);
    z1 = (u64)trunc_div_inv((int64_t)(d * t[4] + t[3] * e + t[5]), div, inv_div);
  };
  if (iters > 1) load(1, Tn, K0n, K1n);
  for (int it = 1; it < iters; ++it) {
    const u64 K0 = K0n, K1 = K1n;
#pragma unroll
    for (int m = 0; m < 3; ++m)
#pragma unroll
      for (int q = 0; q < 6; ++q) T[m][q] = Tn[m][q];
    if (it + 1 < iters) load(it + 1, Tn, K0n, K1n);
    u64 xx0, xx1, w0, w1, t0, t1;
    beaver(X0, X1, X0, X1, T[0], xx0, xx1);
    beaver(V0, V1, xx0, xx1, T[1], w0, w1);
    const u64 y0 = 0 - (w0 - K0), y1 = 0 - (w1 - K1);
    beaver(y0, y1, X0, X1, T[2], t0, t1);
    X0 = (u64)trunc_div_inv((int64_t)t0, Cc, inv_c);
    X1 = (u64)trunc_div_inv((int64_t)t1, Cc, inv_c);
  }
  J.x0[ch] = (int64_t)X0;
  J.x1[ch] = (int64_t)X1;
}


// The same iteration when the two share holders live on DIFFERENT GPUs: each party runs this kernel on its own device and
// the openings travel over NVLink peer-to-peer.  Per Beaver product a thread (= channel) writes its masked operands
// (delta_j, eps_j) into the PEER's mailbox (remote stores through the peer-mapped pointer), publishes them with a
// release-store of the launch's epoch number into the mailbox flag, and polls its OWN mailbox (local memory) for the
// peer's message of the same round.  Mailbox slot = (job, round, channel): written once per launch, so there is no reuse
// hazard inside a launch; across launches the flag value (epoch, bumped by the host-ordered bump kernel before every
// launch on both devices) tells a fresh message from last image's.  The two kernels must be able to run concurrently
// (they sit on different GPUs and neither launch is stream-ordered after the other).  `err` is set when a message does
// not arrive within ~2^24 polls, i.e. seconds (peer kernel never launched): the thread then gives up instead of hanging the GPU.
struct NewtonMsg { u64 d, e, flag, pad; };
struct NewtonP2PJobs { pm_newton_p2p_job_t job[PM_NEWTON_MAX_JOBS]; };

__global__ void newton_epoch_bump_kernel(unsigned long long* epoch) { *epoch += 1ull; }

__global__ void __launch_bounds__(64)
bn_newton_p2p_kernel(const __grid_constant__ NewtonP2PJobs jobs, int party, int iters, int64_t div, int64_t Cc,
                     NewtonMsg* inbox, NewtonMsg* peer_inbox, const unsigned long long* __restrict__ epoch_p,
                     int slot_stride /* channels per (job, round) */, int* __restrict__ err) {
  const pm_newton_p2p_job_t& J = jobs.job[blockIdx.y];
  const int C = J.C;
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= C) return;
  const u64 epoch = *epoch_p;
  const u64 *ta = (const u64*)J.a, *tb = (const u64*)J.b, *tc = (const u64*)J.c, *kk = (const u64*)J.k;
  const double inv_div = 1.0 / (double)div, inv_c = 1.0 / (double)Cc;
  const int rounds = 3 * (iters - 1);
  const size_t slot0 = (size_t)blockIdx.y * rounds * slot_stride + ch;
  const u64 V = (u64)J.v[ch];
  u64 X = (u64)trunc_div_inv((int64_t)(0 - (V - kk[0])), Cc, inv_c);
  bool dead = false;
  auto beaver = [&](u64 p, u64 q, int round) -> u64 {
    const size_t o = (size_t)round * C + ch;
    const u64 a = ta[o], b = tb[o], c = tc[o];
    const u64 dj = p - a, ej = q - b;                                    // spdz_mask
    NewtonMsg* out = peer_inbox + slot0 + (size_t)round * slot_stride;
    out->d = dj; out->e = ej;
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(&out->flag), "l"(epoch) : "memory");
    const NewtonMsg* in = inbox + slot0 + (size_t)round * slot_stride;
    u64 f = 0;
    if (!dead) {
      unsigned long long spins = 0;
      do {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(f) : "l"(&in->flag) : "memory");
        if (++spins > (1ull << 24)) { dead = true; atomicExch(err, 1); break; }  // ~10 s
      } while (f != epoch);
    }
    // the mailbox is written by the other GPU while this kernel runs: volatile loads, ordered after the acquire above
    const u64 d = dj + *reinterpret_cast<const volatile u64*>(&in->d), e = ej + *reinterpret_cast<const volatile u64*>(&in->e);  // opening (spdz.py:162-163)
    u64 z = d * b + a * e + c;                                           // spdz_compute
    if (party == 0) z += d * e;
    return (u64)trunc_div_inv((int64_t)z, div, inv_div);                 // truncate
  };
  for (int it = 1; it < iters; ++it) {
    const u64 K = kk[it];
    const u64 xx = beaver(X, X, 3 * (it - 1));
    const u64 w = beaver(V, xx, 3 * (it - 1) + 1);
    const u64 y = 0 - (w - K);
    const u64 t = beaver(y, X, 3 * (it - 1) + 2);
    X = (u64)trunc_div_inv((int64_t)t, Cc, inv_c);
  }
  J.x[ch] = (int64_t)X;
}

__global__ void avgpool_kernel(const int64_t* __restrict__ x, int H, int W, int k, int64_t* __restrict__ out,
                               size_t total) {
  const int Ho = H / k, Wo = W / k;
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    int ow = (int)(i % Wo);
    int oh = (int)((i / Wo) % Ho);
    size_t bc = i / ((size_t)Wo * Ho);
    const int64_t* p = x + (bc * H + (size_t)oh * k) * W + (size_t)ow * k;
    u64 sacc = 0;
    for (int r = 0; r < k; ++r)
      for (int c = 0; c < k; ++c) sacc += (u64)p[(size_t)r * W + c];
    out[i] = (int64_t)sacc / (int64_t)(k * k);
  }
}

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

int pm_encode_f32_i64(const float* x, float scale, int64_t* q, size_t n, int* overflow, pm_stream_t s) {
  if (n == 0) return PM_OK;
  PM_CHECK_ARG(x && q);
  encode_kernel<<<pm_grid(n, 256), 256, 0, S(s)>>>(x, scale, q, n, overflow);
  PM_LAUNCH_OK();
}

int pm_decode_i64_f32(const int64_t* q, float scale, float* x, size_t n, pm_stream_t s) {
  if (n == 0) return PM_OK;
  PM_CHECK_ARG(x && q);
  decode_kernel<<<pm_grid(n, 256), 256, 0, S(s)>>>(q, scale, x, n);
  PM_LAUNCH_OK();
}

int pm_share_gen_i64(const int64_t* q, uint64_t seed, uint64_t offset, const uint64_t* epoch, int64_t* s0, int64_t* s1, size_t n,
                     pm_stream_t s) {
  if (n == 0) return PM_OK;
  PM_CHECK_ARG(q && s0 && s1);
  philox_kernel<true><<<pm_grid((n + 1) / 2, 256), 256, 0, S(s)>>>(q, seed, offset, (const unsigned long long*)epoch, s0, s1, n);
  PM_LAUNCH_OK();
}

int pm_random_i64(uint64_t seed, uint64_t offset, const uint64_t* epoch, int64_t* out, size_t n, pm_stream_t s) {
  if (n == 0) return PM_OK;
  PM_CHECK_ARG(out);
  philox_kernel<false><<<pm_grid((n + 1) / 2, 256), 256, 0, S(s)>>>(nullptr, seed, offset, (const unsigned long long*)epoch, out, nullptr, n);
  PM_LAUNCH_OK();
}

int pm_epoch_bump(uint64_t* epoch, pm_stream_t s) {
  PM_CHECK_ARG(epoch);
  epoch_bump_kernel<<<1, 1, 0, S(s)>>>((unsigned long long*)epoch);
  PM_LAUNCH_OK();
}

static int conv_out(int H, int k, int stride, int pad, int dil) { return (H + 2 * pad - dil * (k - 1) - 1) / stride + 1; }

int pm_im2col_i64(const int64_t* x, int B, int C, int H, int W, int kh, int kw, int stride, int pad, int dil,
                  int64_t* im, pm_stream_t s) {
  PM_CHECK_ARG(x && im && B > 0 && C > 0 && stride > 0 && dil > 0 && pad >= 0);
  const int Ho = conv_out(H, kh, stride, pad, dil), Wo = conv_out(W, kw, stride, pad, dil);
  PM_CHECK_ARG(Ho > 0 && Wo > 0);
  const size_t total = (size_t)B * Ho * Wo * C * kh * kw;
  im2col_kernel<false><<<pm_grid(total, 256), 256, 0, S(s)>>>(x, C, H, W, kh, kw, stride, pad, dil, Ho, Wo, nullptr,
                                                               im, total);
  PM_LAUNCH_OK();
}

int pm_spdz_mask_im2col_i64(const int64_t* x, int B, int C, int H, int W, int kh, int kw, int stride, int pad,
                            int dil, const int64_t* a, int64_t* delta, pm_stream_t s) {
  PM_CHECK_ARG(x && a && delta && B > 0 && C > 0 && stride > 0 && dil > 0 && pad >= 0);
  const int Ho = conv_out(H, kh, stride, pad, dil), Wo = conv_out(W, kw, stride, pad, dil);
  PM_CHECK_ARG(Ho > 0 && Wo > 0);
  const size_t total = (size_t)B * Ho * Wo * C * kh * kw;
  im2col_kernel<true><<<pm_grid(total, 256), 256, 0, S(s)>>>(x, C, H, W, kh, kw, stride, pad, dil, Ho, Wo, a, delta,
                                                              total);
  PM_LAUNCH_OK();
}

int pm_spdz_mask_i64(const int64_t* x, const int64_t* a, int64_t* delta, size_t n, pm_stream_t s) {
  if (n == 0) return PM_OK;
  PM_CHECK_ARG(x && a && delta);
  sub_kernel<<<pm_grid(n, 256), 256, 0, S(s)>>>(x, a, delta, n);
  PM_LAUNCH_OK();
}

int pm_spdz_mask2_i64(const int64_t* x, const int64_t* a, const int64_t* y, const int64_t* b, int64_t* delta, int64_t* eps, size_t n,
                      pm_stream_t s) {
  if (n == 0) return PM_OK;
  PM_CHECK_ARG(x && a && y && b && delta && eps);
  mask2_kernel<<<pm_grid(n, 256), 256, 0, S(s)>>>((const u64*)x, (const u64*)a, (const u64*)y, (const u64*)b, (u64*)delta, (u64*)eps, n);
  PM_LAUNCH_OK();
}

int pm_spdz_combine_mul_open_i64(int j, const int64_t* d_own, const int64_t* d_peer, const int64_t* e_own, const int64_t* e_peer,
                                 const int64_t* a, const int64_t* b, const int64_t* c, size_t n, int64_t* z, pm_stream_t s) {
  if (n == 0) return PM_OK;
  PM_CHECK_ARG((j == 0 || j == 1) && d_own && d_peer && e_own && e_peer && a && b && c && z);
  combine_mul_open_kernel<<<pm_grid(n, 256), 256, 0, S(s)>>>(j, (const u64*)d_own, (const u64*)d_peer, (const u64*)e_own,
                                                             (const u64*)e_peer, (const u64*)a, (const u64*)b, (const u64*)c, n, (u64*)z);
  PM_LAUNCH_OK();
}

int pm_spdz_mask_wt_i64(const int64_t* w, int N, int K, const int64_t* b, int64_t* eps, pm_stream_t s) {
  PM_CHECK_ARG(w && b && eps && N > 0 && K > 0);
  dim3 grid((N + 31) / 32, (K + 31) / 32), block(32, 8);
  mask_wt_kernel<<<grid, block, 0, S(s)>>>(w, N, K, b, eps);
  PM_LAUNCH_OK();
}

int pm_open_add_i64(const int64_t* local, const int64_t* peer, int64_t* out, size_t n, pm_stream_t s) {
  if (n == 0) return PM_OK;
  PM_CHECK_ARG(local && peer && out);
  add_kernel<<<pm_grid(n, 256), 256, 0, S(s)>>>(local, peer, out, n);
  PM_LAUNCH_OK();
}

int pm_spdz_combine_matmul_i64(int j, const int64_t* delta, const int64_t* eps, const int64_t* a,
                               const int64_t* b, const int64_t* c, int B, int M, int K, int N, int64_t* ws,
                               int64_t* z, pm_stream_t s) {
  PM_CHECK_ARG(delta && eps && a && b && c && z && B > 0 && M > 0 && K > 0 && N > 0 && (j == 0 || j == 1));
  const int64_t* b_eff = b;
  if (j == 0) {
    PM_CHECK_ARG(ws != nullptr);
    const size_t n = (size_t)K * N;
    add_kernel<<<pm_grid(n, 256), 256, 0, S(s)>>>(b, eps, ws, n);  // delta@b + delta@eps == delta@(b+eps) in Z_2^64
    b_eff = ws;
  }
  return launch_gemm2(delta, b_eff, a, eps, c, z, B, M, K, N, S(s));
}

int pm_matmul_i64(const int64_t* A, const int64_t* Bm, int B, int M, int K, int N, int64_t* C, pm_stream_t s) {
  PM_CHECK_ARG(A && Bm && C && B > 0 && M > 0 && K > 0 && N > 0);
  return launch_gemm2(A, Bm, nullptr, nullptr, nullptr, C, B, M, K, N, S(s));
}

int pm_spdz_combine_mul_i64(int j, const int64_t* delta, const int64_t* eps, const int64_t* a, const int64_t* b,
                            const int64_t* c, int mode, size_t P, size_t C, int64_t* z, pm_stream_t s) {
  PM_CHECK_ARG(delta && eps && a && b && c && z && mode >= 0 && mode <= 2 && (j == 0 || j == 1));
  const size_t n = P * C;
  if (n == 0) return PM_OK;
  combine_mul_kernel<<<pm_grid(n, 256), 256, 0, S(s)>>>(j, (const u64*)delta, (const u64*)eps, (const u64*)a,
                                                         (const u64*)b, (const u64*)c, mode, n, C, (u64*)z);
  PM_LAUNCH_OK();
}

int pm_trunc_div_i64(const int64_t* x, int64_t divisor, int64_t* out, size_t n, pm_stream_t s) {
  if (n == 0) return PM_OK;
  PM_CHECK_ARG(x && out && divisor != 0);
  trunc_div_kernel<<<pm_grid(n, 256), 256, 0, S(s)>>>(x, divisor, out, n);
  PM_LAUNCH_OK();
}

int pm_trunc_post_conv_i64(const int64_t* z, int64_t divisor, const int64_t* bias, int B, int M, int N,
                           int64_t* out, pm_stream_t s) {
  PM_CHECK_ARG(z && out && divisor != 0 && B > 0 && M > 0 && N > 0 && B <= 65535);
  dim3 grid((M + 31) / 32, (N + 31) / 32, B), block(32, 8);
  trunc_post_conv_kernel<<<grid, block, 0, S(s)>>>(z, divisor, bias, M, N, out);
  PM_LAUNCH_OK();
}

int pm_axpby_i64(int64_t alpha, const int64_t* x, int64_t beta, const int64_t* y, int ybcast, size_t P, size_t C,
                 int64_t* out, pm_stream_t s) {
  PM_CHECK_ARG(x && out && ybcast >= 0 && ybcast <= 2);
  const size_t n = P * C;
  if (n == 0) return PM_OK;
  axpby_kernel<<<pm_grid(n, 256), 256, 0, S(s)>>>(alpha, x, beta, y, ybcast, n, C, out);
  PM_LAUNCH_OK();
}

int pm_bn_newton_fused_i64(const pm_newton_job_t* jobs, int n_jobs, int iters, int64_t divisor, int64_t newton_c,
                           pm_stream_t s) {
  PM_CHECK_ARG(jobs && n_jobs > 0 && n_jobs <= PM_NEWTON_MAX_JOBS && iters >= 1 && divisor > 0 && newton_c > 0);
  NewtonJobs J;
  int maxC = 0;
  for (int i = 0; i < n_jobs; ++i) {
    const pm_newton_job_t& j = jobs[i];
    PM_CHECK_ARG(j.v0 && j.v1 && j.k0 && j.k1 && j.x0 && j.x1 && j.C > 0);
    PM_CHECK_ARG(iters == 1 || (j.a0 && j.b0 && j.c0 && j.a1 && j.b1 && j.c1));
    J.job[i] = j;
    if (j.C > maxC) maxC = j.C;
  }
  bn_newton_fused_kernel<<<dim3((maxC + 63) / 64, n_jobs), 64, 0, S(s)>>>(J, iters, divisor, newton_c);
  PM_LAUNCH_OK();
}

size_t pm_bn_newton_p2p_mailbox_bytes(int n_jobs, int iters, int max_channels) {
  return (size_t)n_jobs * 3 * (size_t)(iters > 1 ? iters - 1 : 0) * (size_t)max_channels * sizeof(NewtonMsg);
}

int pm_bn_newton_p2p_i64(int party, const pm_newton_p2p_job_t* jobs, int n_jobs, int iters, int64_t divisor, int64_t newton_c,
                         void* inbox, void* peer_inbox, uint64_t* epoch, int max_channels, int* err, pm_stream_t s) {
  PM_CHECK_ARG(jobs && n_jobs > 0 && n_jobs <= PM_NEWTON_MAX_JOBS && iters >= 1 && divisor > 0 && newton_c > 0);
  PM_CHECK_ARG((party == 0 || party == 1) && inbox && peer_inbox && epoch && err && max_channels > 0);
  NewtonP2PJobs J;
  int maxC = 0;
  for (int i = 0; i < n_jobs; ++i) {
    const pm_newton_p2p_job_t& j = jobs[i];
    PM_CHECK_ARG(j.v && j.k && j.x && j.C > 0 && j.C <= max_channels);
    PM_CHECK_ARG(iters == 1 || (j.a && j.b && j.c));
    J.job[i] = j;
    if (j.C > maxC) maxC = j.C;
  }
  newton_epoch_bump_kernel<<<1, 1, 0, S(s)>>>((unsigned long long*)epoch);
  bn_newton_p2p_kernel<<<dim3((maxC + 63) / 64, n_jobs), 64, 0, S(s)>>>(J, party, iters, divisor, newton_c, (NewtonMsg*)inbox,
                                                                        (NewtonMsg*)peer_inbox, (const unsigned long long*)epoch,
                                                                        max_channels, err);
  PM_LAUNCH_OK();
}

int pm_avgpool_i64(const int64_t* x, int B, int C, int H, int W, int k, int64_t* out, pm_stream_t s) {
  PM_CHECK_ARG(x && out && k > 0 && H % k == 0 && W % k == 0);
  const size_t total = (size_t)B * C * (H / k) * (W / k);
  avgpool_kernel<<<pm_grid(total, 128), 128, 0, S(s)>>>(x, H, W, k, out, total);
  PM_LAUNCH_OK();
}

int pm_spdz_affine_i64(const int64_t* u, const int64_t* peer, const int64_t* sc, const int64_t* add, int64_t div, const int64_t* chan,
                       int cs, const int64_t* elem, int es, int C, int HW, size_t n, int64_t* out, pm_stream_t s) {
  PM_CHECK_ARG(u && out && div >= 1 && C >= 1 && HW >= 1 && n % ((size_t)C * HW) == 0 && (cs == 0 || chan) && (es == 0 || elem) &&
               cs >= -1 && cs <= 1 && es >= -1 && es <= 1);
  if (n == 0) return 0;
  spdz_affine_kernel<<<pm_grid(n, 256), 256, 0, S(s)>>>(u, peer, sc, add, div, cs ? chan : nullptr, cs, es ? elem : nullptr, es, C, HW, n, out);
  PM_LAUNCH_OK();
}

int pm_nchw_to_pc_i64(const int64_t* x, int B, int C, int HW, int64_t* out, pm_stream_t s) {
  // functional.py:52-55: permute(1,0,2,3).reshape(C,-1).t() -> row p = b*HW + hw, col c
  PM_CHECK_ARG(x && out && B > 0 && C > 0 && HW > 0 && B <= 65535);
  dim3 grid((C + 31) / 32, (HW + 31) / 32, B), block(32, 8);
  transpose_i64_kernel<<<grid, block, 0, S(s)>>>(x, C, HW, out);  // per b: [C,HW] -> [HW,C]
  PM_LAUNCH_OK();
}

int pm_pc_to_nchw_i64(const int64_t* x, int B, int C, int HW, int64_t* out, pm_stream_t s) {
  PM_CHECK_ARG(x && out && B > 0 && C > 0 && HW > 0 && B <= 65535);
  dim3 grid((HW + 31) / 32, (C + 31) / 32, B), block(32, 8);
  transpose_i64_kernel<<<grid, block, 0, S(s)>>>(x, HW, C, out);  // per b: [HW,C] -> [C,HW]
  PM_LAUNCH_OK();
}

}  // extern "C"
