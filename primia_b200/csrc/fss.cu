// Function secret sharing for the comparison on shares (ReLU / max-pool of the encrypted forward), sm_100a.
//
// Replaces syft/frameworks/torch/mpc/fss.py: DIF.keygen :341-399, DIF.eval :401-428, the PRG H :553-601 (SHA-512 of the
// 16-byte seed through the external `shaloop` C extension), compress/uncompress :431-479, mask_builder :189-204 and the
// opening mod 2^32 of fss_op :158.  One thread owns one comparison instance and walks the 32 levels of its GGM tree;
// key material is laid out structure-of-arrays ([level][word][instance]) so every load of a warp is one coalesced
// 256-byte row.  The kernels are bound by the integer ALU pipe (32 SHA-512 compressions per evaluated element).
#include "common.cuh"

namespace {

__constant__ uint64_t K512[80] = {
    0x428a2f98d728ae22ULL, 0x7137449123ef65cdULL, 0xb5c0fbcfec4d3b2fULL, 0xe9b5dba58189dbbcULL, 0x3956c25bf348b538ULL,
    0x59f111f1b605d019ULL, 0x923f82a4af194f9bULL, 0xab1c5ed5da6d8118ULL, 0xd807aa98a3030242ULL, 0x12835b0145706fbeULL,
    0x243185be4ee4b28cULL, 0x550c7dc3d5ffb4e2ULL, 0x72be5d74f27b896fULL, 0x80deb1fe3b1696b1ULL, 0x9bdc06a725c71235ULL,
    0xc19bf174cf692694ULL, 0xe49b69c19ef14ad2ULL, 0xefbe4786384f25e3ULL, 0x0fc19dc68b8cd5b5ULL, 0x240ca1cc77ac9c65ULL,
    0x2de92c6f592b0275ULL, 0x4a7484aa6ea6e483ULL, 0x5cb0a9dcbd41fbd4ULL, 0x76f988da831153b5ULL, 0x983e5152ee66dfabULL,
    0xa831c66d2db43210ULL, 0xb00327c898fb213fULL, 0xbf597fc7beef0ee4ULL, 0xc6e00bf33da88fc2ULL, 0xd5a79147930aa725ULL,
    0x06ca6351e003826fULL, 0x142929670a0e6e70ULL, 0x27b70a8546d22ffcULL, 0x2e1b21385c26c926ULL, 0x4d2c6dfc5ac42aedULL,
    0x53380d139d95b3dfULL, 0x650a73548baf63deULL, 0x766a0abb3c77b2a8ULL, 0x81c2c92e47edaee6ULL, 0x92722c851482353bULL,
    0xa2bfe8a14cf10364ULL, 0xa81a664bbc423001ULL, 0xc24b8b70d0f89791ULL, 0xc76c51a30654be30ULL, 0xd192e819d6ef5218ULL,
    0xd69906245565a910ULL, 0xf40e35855771202aULL, 0x106aa07032bbd1b8ULL, 0x19a4c116b8d2d0c8ULL, 0x1e376c085141ab53ULL,
    0x2748774cdf8eeb99ULL, 0x34b0bcb5e19b48a8ULL, 0x391c0cb3c5c95a63ULL, 0x4ed8aa4ae3418acbULL, 0x5b9cca4f7763e373ULL,
    0x682e6ff3d6b2b8a3ULL, 0x748f82ee5defb2fcULL, 0x78a5636f43172f60ULL, 0x84c87814a1f0ab72ULL, 0x8cc702081a6439ecULL,
    0x90befffa23631e28ULL, 0xa4506cebde82bde9ULL, 0xbef9a3f7b2c67915ULL, 0xc67178f2e372532bULL, 0xca273eceea26619cULL,
    0xd186b8c721c0c207ULL, 0xeada7dd6cde0eb1eULL, 0xf57d4f7fee6ed178ULL, 0x06f067aa72176fbaULL, 0x0a637dc5a2c898a6ULL,
    0x113f9804bef90daeULL, 0x1b710b35131c471bULL, 0x28db77f523047d84ULL, 0x32caab7b40c72493ULL, 0x3c9ebe0a15c9bebcULL,
    0x431d67c49c100d4cULL, 0x4cc5d4becb3e42b6ULL, 0x597f299cfc657e2aULL, 0x5fcb6fab3ad6faecULL, 0x6c44198c4a475817ULL};

template <int N>
__device__ __forceinline__ uint64_t rotr64(uint64_t x) {
  const uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
  uint32_t rl, rh;
  if constexpr (N < 32) {
    rl = __funnelshift_r(lo, hi, N);
    rh = __funnelshift_r(hi, lo, N);
  } else {
    rl = __funnelshift_r(hi, lo, N - 32);
    rh = __funnelshift_r(lo, hi, N - 32);
  }
  return ((uint64_t)rh << 32) | rl;
}
__device__ __forceinline__ uint64_t bswap64(uint64_t x) {
  const uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
  return ((uint64_t)__byte_perm(lo, 0, 0x0123) << 32) | __byte_perm(hi, 0, 0x0123);
}
__device__ __forceinline__ uint64_t S0(uint64_t x) { return rotr64<28>(x) ^ rotr64<34>(x) ^ rotr64<39>(x); }
__device__ __forceinline__ uint64_t S1(uint64_t x) { return rotr64<14>(x) ^ rotr64<18>(x) ^ rotr64<41>(x); }
__device__ __forceinline__ uint64_t s0(uint64_t x) { return rotr64<1>(x) ^ rotr64<8>(x) ^ (x >> 7); }
__device__ __forceinline__ uint64_t s1(uint64_t x) { return rotr64<19>(x) ^ rotr64<61>(x) ^ (x >> 6); }

// One SHA-512 round on NH interleaved hashes; (a..h) are passed rotated by the caller so no register moves are needed.
#define PM_SHA_ROUND(a, b, c, d, e, f, g, h, kw)                                       \
  _Pragma("unroll") for (int q = 0; q < NH; ++q) {                                     \
    h[q] += S1(e[q]) + ((e[q] & f[q]) ^ (~e[q] & g[q])) + (kw)[q];                     \
    d[q] += h[q];                                                                      \
    h[q] += S0(a[q]) + ((a[q] & b[q]) ^ (a[q] & c[q]) ^ (b[q] & c[q]));                \
  }

// SHA-512 (FIPS 180-4) of NH independent 16-byte messages; message q = the two little-endian words (m0[q], m1[q]) of a
// seed, digest returned as its eight little-endian words -- the reference's `buffer.view(np.uint64)` of shaloop's output
// (fss.py:581-586).  Single padded block: W2 = 0x80.., W15 = 128 (bit length).
template <int NH>
__device__ __forceinline__ void sha512_seed16(const uint64_t (&m0)[NH], const uint64_t (&m1)[NH], uint64_t (&dig)[NH][8]) {
  uint64_t W[16][NH];
  uint64_t a[NH], b[NH], c[NH], d[NH], e[NH], f[NH], g[NH], h[NH];
#pragma unroll
  for (int q = 0; q < NH; ++q) {
    W[0][q] = bswap64(m0[q]);
    W[1][q] = bswap64(m1[q]);
    W[2][q] = 0x8000000000000000ULL;
#pragma unroll
    for (int i = 3; i < 15; ++i) W[i][q] = 0;
    W[15][q] = 128;
    a[q] = 0x6a09e667f3bcc908ULL; b[q] = 0xbb67ae8584caa73bULL; c[q] = 0x3c6ef372fe94f82bULL; d[q] = 0xa54ff53a5f1d36f1ULL;
    e[q] = 0x510e527fade682d1ULL; f[q] = 0x9b05688c2b3e6c1fULL; g[q] = 0x1f83d9abfb41bd6bULL; h[q] = 0x5be0cd19137e2179ULL;
  }
  uint64_t kw[NH];
#define PM_KW(i, base)                                                      \
  _Pragma("unroll") for (int q = 0; q < NH; ++q) kw[q] = K512[(base) + (i)] + W[(i)][q];
#define PM_SCHED(i)                                                          \
  _Pragma("unroll") for (int q = 0; q < NH; ++q)                             \
      W[(i)][q] += s1(W[((i) + 14) & 15][q]) + W[((i) + 9) & 15][q] + s0(W[((i) + 1) & 15][q]);
#define PM_8ROUNDS(o, base, SCHED)                                           \
  SCHED((o) + 0) PM_KW((o) + 0, base) PM_SHA_ROUND(a, b, c, d, e, f, g, h, kw) \
  SCHED((o) + 1) PM_KW((o) + 1, base) PM_SHA_ROUND(h, a, b, c, d, e, f, g, kw) \
  SCHED((o) + 2) PM_KW((o) + 2, base) PM_SHA_ROUND(g, h, a, b, c, d, e, f, kw) \
  SCHED((o) + 3) PM_KW((o) + 3, base) PM_SHA_ROUND(f, g, h, a, b, c, d, e, kw) \
  SCHED((o) + 4) PM_KW((o) + 4, base) PM_SHA_ROUND(e, f, g, h, a, b, c, d, kw) \
  SCHED((o) + 5) PM_KW((o) + 5, base) PM_SHA_ROUND(d, e, f, g, h, a, b, c, kw) \
  SCHED((o) + 6) PM_KW((o) + 6, base) PM_SHA_ROUND(c, d, e, f, g, h, a, b, kw) \
  SCHED((o) + 7) PM_KW((o) + 7, base) PM_SHA_ROUND(b, c, d, e, f, g, h, a, kw)
#define PM_NOSCHED(i)
  // rounds 0..15: most of the message words are constants, folded by the compiler
  PM_8ROUNDS(0, 0, PM_NOSCHED)
  PM_8ROUNDS(8, 0, PM_NOSCHED)
#pragma unroll 1
  for (int r = 16; r < 80; r += 16) {
    PM_8ROUNDS(0, r, PM_SCHED)
    PM_8ROUNDS(8, r, PM_SCHED)
  }
#undef PM_NOSCHED
#undef PM_8ROUNDS
#undef PM_SCHED
#undef PM_KW
#pragma unroll
  for (int q = 0; q < NH; ++q) {
    dig[q][0] = bswap64(a[q] + 0x6a09e667f3bcc908ULL);
    dig[q][1] = bswap64(b[q] + 0xbb67ae8584caa73bULL);
    dig[q][2] = bswap64(c[q] + 0x3c6ef372fe94f82bULL);
    dig[q][3] = bswap64(d[q] + 0xa54ff53a5f1d36f1ULL);
    dig[q][4] = bswap64(e[q] + 0x510e527fade682d1ULL);
    dig[q][5] = bswap64(f[q] + 0x9b05688c2b3e6c1fULL);
    dig[q][6] = bswap64(g[q] + 0x1f83d9abfb41bd6bULL);
    dig[q][7] = bswap64(h[q] + 0x5be0cd19137e2179ULL);
  }
}

constexpr uint64_t CLR1 = 0xFFFFFFFFFFFFFFFEULL;
constexpr uint64_t M31 = 0x7FFFFFFFULL;  // convert(): the 31 low bits of the last word (fss.py:655-661)

// raw PRG (for the known-answer test against the reference's H)
__global__ void fss_prg_kernel(const uint64_t* __restrict__ seed, size_t n, uint64_t* __restrict__ out) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    uint64_t m0[1] = {seed[i]}, m1[1] = {seed[n + i]}, dig[1][8];
    sha512_seed16<1>(m0, m1, dig);
#pragma unroll
    for (int k = 0; k < 8; ++k) out[k * n + i] = dig[0][k];
  }
}

// DIF.keygen (fss.py:341-399).  alpha [n]; seeds [2 parties][2 words][n]; outputs are the compressed correction words.
__global__ void __launch_bounds__(128)
fss_dif_keygen_kernel(const uint64_t* __restrict__ alpha, const uint64_t* __restrict__ seeds, size_t n, size_t stride,
                      uint8_t* __restrict__ bits, uint64_t* __restrict__ sigma_cw, uint64_t* __restrict__ s_cw,
                      int32_t* __restrict__ leaf) {
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
    const uint32_t al = (uint32_t)alpha[idx];
    uint64_t sa[2] = {seeds[idx], seeds[2 * n + idx]};            // word 0 of party 0 / party 1
    uint64_t sb[2] = {seeds[n + idx], seeds[3 * n + idx]};        // word 1
    uint64_t t[2] = {0, 1};
#pragma unroll 1
    for (int i = 0; i < 32; ++i) {
      const uint32_t ai = (al >> (31 - i)) & 1u;
      uint64_t h[2][8];
      sha512_seed16<2>(sa, sb, h);
      // x = h0 ^ h1 as (sigma[2], tau, s[2], t) per direction r
      uint64_t xs[2][6], hh[2][2][6];
#pragma unroll
      for (int p = 0; p < 2; ++p)
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const uint64_t* w = &h[p][4 * r];
          hh[p][r][0] = w[0] & CLR1; hh[p][r][1] = w[1]; hh[p][r][2] = w[0] & 1;
          hh[p][r][3] = w[2] & CLR1; hh[p][r][4] = w[3]; hh[p][r][5] = w[2] & 1;
        }
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int k = 0; k < 6; ++k) xs[r][k] = hh[0][r][k] ^ hh[1][r][k];
      // SwitchTableDIF (fss.py:628-645) ^ h0 ^ h1, then compress (fss.py:431-455)
      const uint64_t sr0 = ai ? xs[0][3] : xs[1][3], sr1 = ai ? xs[0][4] : xs[1][4];
      const uint64_t gr0 = ai ? xs[0][0] : xs[1][0], gr1 = ai ? xs[0][1] : xs[1][1];
      uint64_t cw[2][6];
      // row 0 (left): leaf part switched on by alpha_i, next part by 1-alpha_i ; row 1 the opposite
      cw[0][0] = (ai ? gr0 : 0) ^ xs[0][0]; cw[0][1] = (ai ? gr1 : 0) ^ xs[0][1]; cw[0][2] = (uint64_t)ai ^ xs[0][2];
      cw[1][0] = (ai ? 0 : gr0) ^ xs[1][0]; cw[1][1] = (ai ? 0 : gr1) ^ xs[1][1]; cw[1][2] = (uint64_t)(1 - ai) ^ xs[1][2];
      cw[0][3] = (ai ? 0 : sr0) ^ xs[0][3]; cw[0][4] = (ai ? 0 : sr1) ^ xs[0][4]; cw[0][5] = (uint64_t)(1 - ai) ^ xs[0][5];
      cw[1][3] = (ai ? sr0 : 0) ^ xs[1][3]; cw[1][4] = (ai ? sr1 : 0) ^ xs[1][4]; cw[1][5] = (uint64_t)ai ^ xs[1][5];
      const uint32_t bt = (uint32_t)(cw[0][2] | (cw[0][5] << 1) | (cw[1][2] << 2) | (cw[1][5] << 3));
      const uint64_t sg0 = ai ? cw[1][0] : cw[0][0], sg1 = ai ? cw[1][1] : cw[0][1];
      const uint64_t sc0 = ai ? cw[0][3] : cw[1][3], sc1 = ai ? cw[0][4] : cw[1][4];
      bits[(size_t)i * stride + idx] = (uint8_t)bt;
      sigma_cw[(size_t)(2 * i) * stride + idx] = sg0;
      sigma_cw[(size_t)(2 * i + 1) * stride + idx] = sg1;
      s_cw[(size_t)(2 * i) * stride + idx] = sc0;
      s_cw[(size_t)(2 * i + 1) * stride + idx] = sc1;
      // uncompress (fss.py:458-479) and advance both parties
      int64_t cs[2];
      uint64_t tau[2];
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        uint64_t dual[2][6];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const uint64_t m = t[p] ? ~0ULL : 0ULL;
          dual[r][0] = hh[p][r][0] ^ (m & sg0); dual[r][1] = hh[p][r][1] ^ (m & sg1);
          dual[r][2] = hh[p][r][2] ^ (m & ((bt >> (2 * r)) & 1));
          dual[r][3] = hh[p][r][3] ^ (m & sc0); dual[r][4] = hh[p][r][4] ^ (m & sc1);
          dual[r][5] = hh[p][r][5] ^ (m & ((bt >> (2 * r + 1)) & 1));
        }
        // keep = dual[alpha_i] (special path), anti = dual[1 - alpha_i]
        sa[p] = ai ? dual[1][3] : dual[0][3];
        sb[p] = ai ? dual[1][4] : dual[0][4];
        t[p] = ai ? dual[1][5] : dual[0][5];
        cs[p] = (int64_t)((ai ? dual[0][1] : dual[1][1]) & M31);
        tau[p] = ai ? dual[0][2] : dual[1][2];
      }
      int64_t lf = 1 - cs[0] + cs[1] - (int64_t)(1 - ai);
      if (tau[1]) lf = -lf;
      leaf[(size_t)i * stride + idx] = (int32_t)lf;
    }
    int64_t lf = 1 - (int64_t)(sb[0] & M31) + (int64_t)(sb[1] & M31);
    if (t[1]) lf = -lf;
    leaf[(size_t)32 * stride + idx] = (int32_t)lf;
  }
}

// DIF.eval (fss.py:401-428): party b's int64 share of [x <= alpha] over the low 32 bits of x.
__global__ void __launch_bounds__(128)
fss_dif_eval_kernel(int b, const int64_t* __restrict__ x, const int64_t* __restrict__ peer /* != NULL: x + peer is opened here */,
                    const uint64_t* __restrict__ s0, const uint8_t* __restrict__ bits,
                    const uint64_t* __restrict__ sigma_cw, const uint64_t* __restrict__ s_cw, const int32_t* __restrict__ leaf,
                    size_t n, size_t stride, int64_t* __restrict__ out) {
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
    const uint32_t xv = (uint32_t)x[idx] + (peer ? (uint32_t)peer[idx] : 0u);   // the low 32 bits are all DIF reads
    uint64_t sa[1] = {s0[idx]}, sb[1] = {s0[stride + idx]};
    uint64_t t = (uint64_t)b;
    int64_t acc = 0;
#pragma unroll 1
    for (int i = 0; i < 32; ++i) {
      // correction words of this level: issued before the hash so the loads overlap it
      // (sigma word 0 only carries tau in its low bit, which travels in `bits`; convert() reads word 1)
      const uint64_t sg1 = sigma_cw[(size_t)(2 * i + 1) * stride + idx];
      const uint64_t sc0 = s_cw[(size_t)(2 * i) * stride + idx], sc1 = s_cw[(size_t)(2 * i + 1) * stride + idx];
      const uint32_t bt = bits[(size_t)i * stride + idx];
      const int64_t lf = leaf[(size_t)i * stride + idx];
      uint64_t h[1][8];
      sha512_seed16<1>(sa, sb, h);
      const uint32_t xb = (xv >> (31 - i)) & 1u;
      const uint64_t w0 = xb ? h[0][4] : h[0][0], w1 = xb ? h[0][5] : h[0][1];
      const uint64_t w2 = xb ? h[0][6] : h[0][2], w3 = xb ? h[0][7] : h[0][3];
      const uint64_t m = t ? ~0ULL : 0ULL;
      const uint64_t sig1 = w1 ^ (m & sg1);
      const uint64_t tau = (w0 & 1) ^ (m & ((bt >> (2 * xb)) & 1));
      sa[0] = (w2 & CLR1) ^ (m & sc0);
      sb[0] = w3 ^ (m & sc1);
      t = (w2 & 1) ^ (m & ((bt >> (2 * xb + 1)) & 1));
      acc += (tau ? lf : 0) + (int64_t)(sig1 & M31);
    }
    const int64_t lf = leaf[(size_t)32 * stride + idx];
    acc += (t ? lf : 0) + (int64_t)(sb[0] & M31);
    out[idx] = b ? -acc : acc;
  }
}

// mask_builder (fss.py:189-204): r_j = x1_j - x2_j + alpha_j ; either operand may be absent (public 0)
__global__ void fss_mask_kernel(const int64_t* __restrict__ x1, const int64_t* __restrict__ x2, const int64_t* __restrict__ alpha,
                                int64_t* __restrict__ r, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const uint64_t a = x1 ? (uint64_t)x1[i] : 0, c = x2 ? (uint64_t)x2[i] : 0;
    r[i] = (int64_t)(a - c + (uint64_t)alpha[i]);
  }
}
// opening of the masked value: sum(shares) % 2**32 (fss.py:158); `peer` may be a peer-mapped pointer
__global__ void fss_open_kernel(const int64_t* __restrict__ local, const int64_t* __restrict__ peer, int64_t* __restrict__ out,
                                size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = (int64_t)(((uint64_t)local[i] + (uint64_t)peer[i]) & 0xFFFFFFFFULL);
}
// randomness of keygen (fss.py:346,354,498-505; primitives.py:245-251) conditioned from raw 64-bit words:
// alpha, mask in [0,2^32); seed word 0 < 2^63; alpha share of party 0 = (alpha - mask) mod 2^32
__global__ void fss_condition_kernel(uint64_t* __restrict__ alpha, uint64_t* __restrict__ mask, uint64_t* __restrict__ seeds,
                                     int64_t* __restrict__ alpha0, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const uint64_t a = alpha[i] & 0xFFFFFFFFULL, m = mask[i] & 0xFFFFFFFFULL;
    alpha[i] = a;
    mask[i] = m;
    alpha0[i] = (int64_t)((a - m) & 0xFFFFFFFFULL);
    seeds[i] &= 0x7FFFFFFFFFFFFFFFULL;
    seeds[2 * n + i] &= 0x7FFFFFFFFFFFFFFFULL;
  }
}

// _pre_pool (nn/functional.py:312-390): x [B,C,H,W] -> [B,C,M=Ho*Wo,k*k], zero padding, window index = r*k + c
__global__ void pre_pool_kernel(const int64_t* __restrict__ x, int H, int W, int k, int stride, int pad, int Ho, int Wo,
                                int64_t* __restrict__ out, size_t total) {
  const int kk = k * k;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int tap = (int)(i % kk);
    const size_t m = i / kk;
    const int ow = (int)(m % Wo), oh = (int)((m / Wo) % Ho);
    const size_t bc = m / ((size_t)Wo * Ho);
    const int ih = oh * stride - pad + tap / k, iw = ow * stride - pad + tap % k;
    out[i] = (ih >= 0 && ih < H && iw >= 0 && iw < W) ? x[(bc * H + ih) * W + iw] : 0;
  }
}
// t[..., start:start+len] of a [rows, L] tensor, made contiguous (the halves of max_half_split, functional.py:489-508)
__global__ void slice_lastdim_kernel(const int64_t* __restrict__ src, int L, int start, int len, int64_t* __restrict__ dst,
                                     size_t total) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = src[(i / len) * L + start + (i % len)];
}

}  // namespace

extern "C" int pm_pre_pool_i64(const int64_t* x, int B, int C, int H, int W, int k, int stride, int pad, int64_t* out,
                               pm_stream_t s) {
  PM_CHECK_ARG(x && out && B > 0 && C > 0 && k > 0 && stride > 0 && pad >= 0);
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  PM_CHECK_ARG(Ho > 0 && Wo > 0);
  const size_t total = (size_t)B * C * Ho * Wo * k * k;
  pre_pool_kernel<<<pm_grid(total, 256), 256, 0, S(s)>>>(x, H, W, k, stride, pad, Ho, Wo, out, total);
  PM_LAUNCH_OK();
}

extern "C" int pm_slice_lastdim_i64(const int64_t* src, size_t rows, int L, int start, int len, int64_t* dst, pm_stream_t s) {
  if (rows == 0 || len == 0) return PM_OK;
  PM_CHECK_ARG(src && dst && start >= 0 && len > 0 && start + len <= L);
  const size_t total = rows * (size_t)len;
  slice_lastdim_kernel<<<pm_grid(total, 256), 256, 0, S(s)>>>(src, L, start, len, dst, total);
  PM_LAUNCH_OK();
}

extern "C" int pm_fss_prg_sha512(const uint64_t* seed, size_t n, uint64_t* out, pm_stream_t s) {
  if (n == 0) return PM_OK;
  PM_CHECK_ARG(seed && out);
  fss_prg_kernel<<<pm_grid(n, 128), 128, 0, S(s)>>>(seed, n, out);
  PM_LAUNCH_OK();
}

extern "C" int pm_fss_dif_keygen(const uint64_t* alpha, const uint64_t* seeds, size_t n, size_t stride, uint8_t* bits,
                                 uint64_t* sigma_cw, uint64_t* s_cw, int32_t* leaf, pm_stream_t s) {
  if (n == 0) return PM_OK;
  PM_CHECK_ARG(alpha && seeds && bits && sigma_cw && s_cw && leaf && stride >= n);
  fss_dif_keygen_kernel<<<pm_grid(n, 128, 1, 64), 128, 0, S(s)>>>(alpha, seeds, n, stride, bits, sigma_cw, s_cw, leaf);
  PM_LAUNCH_OK();
}

extern "C" int pm_fss_dif_eval(int b, const int64_t* x_masked, const uint64_t* s0, const uint8_t* bits,
                               const uint64_t* sigma_cw, const uint64_t* s_cw, const int32_t* leaf, size_t n, size_t stride,
                               int64_t* out, pm_stream_t s) {
  if (n == 0) return PM_OK;
  PM_CHECK_ARG((b == 0 || b == 1) && x_masked && s0 && bits && sigma_cw && s_cw && leaf && out && stride >= n);
  fss_dif_eval_kernel<<<pm_grid(n, 128, 1, 64), 128, 0, S(s)>>>(b, x_masked, nullptr, s0, bits, sigma_cw, s_cw, leaf, n, stride, out);
  PM_LAUNCH_OK();
}

extern "C" int pm_fss_dif_eval_open(int b, const int64_t* r_own, const int64_t* r_peer, const uint64_t* s0, const uint8_t* bits,
                                    const uint64_t* sigma_cw, const uint64_t* s_cw, const int32_t* leaf, size_t n, size_t stride,
                                    int64_t* out, pm_stream_t s) {
  if (n == 0) return PM_OK;
  PM_CHECK_ARG((b == 0 || b == 1) && r_own && r_peer && s0 && bits && sigma_cw && s_cw && leaf && out && stride >= n);
  fss_dif_eval_kernel<<<pm_grid(n, 128, 1, 64), 128, 0, S(s)>>>(b, r_own, r_peer, s0, bits, sigma_cw, s_cw, leaf, n, stride, out);
  PM_LAUNCH_OK();
}

extern "C" int pm_fss_mask_i64(const int64_t* x1, const int64_t* x2, const int64_t* alpha_share, int64_t* r, size_t n,
                               pm_stream_t s) {
  if (n == 0) return PM_OK;
  PM_CHECK_ARG(alpha_share && r);
  fss_mask_kernel<<<pm_grid(n, 256), 256, 0, S(s)>>>(x1, x2, alpha_share, r, n);
  PM_LAUNCH_OK();
}

extern "C" int pm_fss_open_mod32_i64(const int64_t* local, const int64_t* peer, int64_t* out, size_t n, pm_stream_t s) {
  if (n == 0) return PM_OK;
  PM_CHECK_ARG(local && peer && out);
  fss_open_kernel<<<pm_grid(n, 256), 256, 0, S(s)>>>(local, peer, out, n);
  PM_LAUNCH_OK();
}

extern "C" int pm_fss_condition_randomness(uint64_t* alpha, uint64_t* mask, uint64_t* seeds, int64_t* alpha0, size_t n,
                                           pm_stream_t s) {
  if (n == 0) return PM_OK;
  PM_CHECK_ARG(alpha && mask && seeds && alpha0);
  fss_condition_kernel<<<pm_grid(n, 256), 256, 0, S(s)>>>(alpha, mask, seeds, alpha0, n);
  PM_LAUNCH_OK();
}
