/* CPU oracle for the FSS comparison, C restatement -- TEST INFRASTRUCTURE ONLY (see oracle/fss_oracle.py, which this file
 * follows function by function and against which tests/test_oracle_fss.py pins it, together with the fixtures produced by
 * executing the reference's own syft/frameworks/torch/mpc/fss.py).  Only tests/, __graft_entry__.smoke() and bench.py's
 * CPU-baseline legs may load it.  Exists because the full-size (224 x 224, 3.3 M comparisons, 2 x 106 M SHA-512) encrypted
 * forward takes ~10 minutes through hashlib's per-call overhead and seconds here.
 *
 *   H          fss.py:553-601   PRG: SHA-512 (FIPS 180-4) of the 16-byte seed; shaloop.sha512_loop_func in the reference
 *   dif_eval   fss.py:401-428   DIF.eval
 *   dif_keygen fss.py:341-399   DIF.keygen (+ compress / uncompress :431-479, SwitchTableDIF :628-645)
 * Layouts are the numpy oracle's: seeds [2][n] (word, value), bits [32][4][n] (tauL, tL, tauR, tR), sigma_cw / s_cw [32][2][n],
 * leaf [33][n] int32.
 *
 *   gcc -O3 -fopenmp -shared -fPIC -o oracle/_build/libfss_oracle.so oracle/fss_oracle_c.c
 */
#include <stddef.h>
#include <stdint.h>
#include <string.h>

static const uint64_t K512[80] = {
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

#define ROR(x, n) (((x) >> (n)) | ((x) << (64 - (n))))

/* SHA-512 of the 16 bytes (w0 little-endian, w1 little-endian); digest returned as the 8 little-endian uint64 words of its
 * bytes -- what `dig.view(np.uint64)` gives the numpy oracle. */
static void sha512_16(uint64_t w0, uint64_t w1, uint64_t out[8]) {
  uint64_t W[80];
  W[0] = __builtin_bswap64(w0);
  W[1] = __builtin_bswap64(w1);
  W[2] = 0x8000000000000000ULL;
  for (int i = 3; i < 15; ++i) W[i] = 0;
  W[15] = 128; /* message length in bits */
  for (int i = 16; i < 80; ++i) {
    const uint64_t s0 = ROR(W[i - 15], 1) ^ ROR(W[i - 15], 8) ^ (W[i - 15] >> 7);
    const uint64_t s1 = ROR(W[i - 2], 19) ^ ROR(W[i - 2], 61) ^ (W[i - 2] >> 6);
    W[i] = W[i - 16] + s0 + W[i - 7] + s1;
  }
  uint64_t h[8] = {0x6a09e667f3bcc908ULL, 0xbb67ae8584caa73bULL, 0x3c6ef372fe94f82bULL, 0xa54ff53a5f1d36f1ULL,
                   0x510e527fade682d1ULL, 0x9b05688c2b3e6c1fULL, 0x1f83d9abfb41bd6bULL, 0x5be0cd19137e2179ULL};
  uint64_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
  for (int i = 0; i < 80; ++i) {
    const uint64_t S1 = ROR(e, 14) ^ ROR(e, 18) ^ ROR(e, 41);
    const uint64_t ch = (e & f) ^ (~e & g);
    const uint64_t t1 = hh + S1 + ch + K512[i] + W[i];
    const uint64_t S0 = ROR(a, 28) ^ ROR(a, 34) ^ ROR(a, 39);
    const uint64_t mj = (a & b) ^ (a & c) ^ (b & c);
    const uint64_t t2 = S0 + mj;
    hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
  }
  h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
  for (int i = 0; i < 8; ++i) out[i] = __builtin_bswap64(h[i]);
}

#define CLR1 0xFFFFFFFFFFFFFFFEULL
#define MASK31 0x7FFFFFFFULL

/* valuebits of one seed: v[r] = (sigma0, sigma1, tau, s0, s1, t) for direction r (0 left, 1 right) -- fss_oracle.H */
static void H1(uint64_t w0, uint64_t w1, uint64_t v[2][6]) {
  uint64_t d[8];
  sha512_16(w0, w1, d);
  for (int r = 0; r < 2; ++r) {
    const uint64_t* w = d + 4 * r;
    v[r][0] = w[0] & CLR1; v[r][1] = w[1]; v[r][2] = w[0] & 1ULL;
    v[r][3] = w[2] & CLR1; v[r][4] = w[3]; v[r][5] = w[2] & 1ULL;
  }
}

void fss_H(const uint64_t* seed /* [2][n] */, size_t n, uint64_t* out /* [2][6][n] */) {
#pragma omp parallel for schedule(static)
  for (long i = 0; i < (long)n; ++i) {
    uint64_t v[2][6];
    H1(seed[i], seed[n + i], v);
    for (int r = 0; r < 2; ++r)
      for (int k = 0; k < 6; ++k) out[((size_t)r * 6 + k) * n + i] = v[r][k];
  }
}

void fss_dif_eval(int b, const int64_t* x, const uint64_t* s0 /* [2][n] */, const uint8_t* bits /* [32][4][n] */,
                  const uint64_t* sigma_cw /* [32][2][n] */, const uint64_t* s_cw /* [32][2][n] */, const int32_t* leaf /* [33][n] */,
                  size_t n, int64_t* out) {
#pragma omp parallel for schedule(static)
  for (long i = 0; i < (long)n; ++i) {
    uint64_t s[2] = {s0[i], s0[n + i]};
    uint64_t t = (uint64_t)b;
    const uint64_t sign = b ? (uint64_t)-1 : 1ULL; /* arithmetic mod 2^64 == numpy int64 wraparound */
    uint64_t acc = 0;
    const uint32_t xv = (uint32_t)(uint64_t)x[i];
    for (int l = 0; l < 32; ++l) {
      uint64_t h[2][6];
      H1(s[0], s[1], h);
      const int xb = (xv >> (31 - l)) & 1;
      const size_t o2 = ((size_t)l * 2) * n + i, o4 = ((size_t)l * 4) * n + i;
      const uint64_t cw[6] = {sigma_cw[o2], sigma_cw[o2 + n], bits[o4 + (size_t)(2 * xb) * n], s_cw[o2], s_cw[o2 + n],
                              bits[o4 + (size_t)(2 * xb + 1) * n]};
      uint64_t st[6];
      for (int k = 0; k < 6; ++k) st[k] = h[xb][k] ^ (t * cw[k]);
      const uint64_t tau = st[2];
      acc += sign * (tau * (uint64_t)(int64_t)leaf[(size_t)l * n + i] + (st[1] & MASK31));
      s[0] = st[3]; s[1] = st[4]; t = st[5];
    }
    acc += sign * (t * (uint64_t)(int64_t)leaf[(size_t)32 * n + i] + (s[1] & MASK31));
    out[i] = (int64_t)acc;
  }
}

void fss_dif_keygen(const uint64_t* alpha /* [n] */, const uint64_t* seeds /* [2 parties][2 words][n] */, size_t n, uint8_t* bits,
                    uint64_t* sigma_cw, uint64_t* s_cw, int32_t* leaf) {
#pragma omp parallel for schedule(static)
  for (long i = 0; i < (long)n; ++i) {
    uint64_t s[2][2] = {{seeds[i], seeds[n + i]}, {seeds[2 * n + i], seeds[3 * n + i]}};
    uint64_t t[2] = {0, 1};
    const uint32_t av = (uint32_t)alpha[i];
    for (int l = 0; l < 32; ++l) {
      const uint64_t ai = (av >> (31 - l)) & 1;
      uint64_t h[2][2][6], x[2][6], table[2][6], cw[2][6];
      H1(s[0][0], s[0][1], h[0]);
      H1(s[1][0], s[1][1], h[1]);
      for (int r = 0; r < 2; ++r)
        for (int k = 0; k < 6; ++k) x[r][k] = h[0][r][k] ^ h[1][r][k];
      /* SwitchTableDIF (fss.py:628-645): leaf part switched by 1 - alpha_i, next part by alpha_i */
      const uint64_t s_rand[2] = {ai ? x[0][3] : x[1][3], ai ? x[0][4] : x[1][4]};
      const uint64_t sg_rand[2] = {ai ? x[0][0] : x[1][0], ai ? x[0][1] : x[1][1]};
      const uint64_t na = 1 - ai;
      table[0][0] = sg_rand[0] * ai; table[0][1] = sg_rand[1] * ai; table[0][2] = ai;
      table[1][0] = sg_rand[0] * na; table[1][1] = sg_rand[1] * na; table[1][2] = na;
      table[0][3] = s_rand[0] * na; table[0][4] = s_rand[1] * na; table[0][5] = na;
      table[1][3] = s_rand[0] * ai; table[1][4] = s_rand[1] * ai; table[1][5] = ai;
      for (int r = 0; r < 2; ++r)
        for (int k = 0; k < 6; ++k) cw[r][k] = table[r][k] ^ x[r][k];
      /* compress (fss.py:431-455) */
      const size_t o2 = ((size_t)l * 2) * n + i, o4 = ((size_t)l * 4) * n + i;
      const uint8_t bt[4] = {(uint8_t)cw[0][2], (uint8_t)cw[0][5], (uint8_t)cw[1][2], (uint8_t)cw[1][5]};
      for (int k = 0; k < 4; ++k) bits[o4 + (size_t)k * n] = bt[k];
      const uint64_t sg[2] = {ai ? cw[1][0] : cw[0][0], ai ? cw[1][1] : cw[0][1]};
      const uint64_t sc[2] = {ai ? cw[0][3] : cw[1][3], ai ? cw[0][4] : cw[1][4]};
      sigma_cw[o2] = sg[0]; sigma_cw[o2 + n] = sg[1];
      s_cw[o2] = sc[0]; s_cw[o2 + n] = sc[1];
      /* uncompress (:458-479) and one evaluation step for both parties */
      uint64_t sig1[2], tau[2];
      for (int p = 0; p < 2; ++p) {
        uint64_t dual[2][6];
        for (int r = 0; r < 2; ++r) {
          const uint64_t cwi[6] = {sg[0], sg[1], bt[2 * r], sc[0], sc[1], bt[2 * r + 1]};
          for (int k = 0; k < 6; ++k) dual[r][k] = h[p][r][k] ^ (t[p] * cwi[k]);
        }
        const uint64_t* keep = ai ? dual[1] : dual[0]; /* follow the special path */
        const uint64_t* anti = ai ? dual[0] : dual[1]; /* leave it */
        s[p][0] = keep[3]; s[p][1] = keep[4]; t[p] = keep[5];
        sig1[p] = anti[1]; tau[p] = anti[2];
      }
      const int64_t sign = tau[1] ? -1 : 1;
      const int64_t v = sign * (1 - (int64_t)(sig1[0] & MASK31) + (int64_t)(sig1[1] & MASK31) - (1 - (int64_t)ai));
      leaf[(size_t)l * n + i] = (int32_t)v;
    }
    const int64_t sign = t[1] ? -1 : 1;
    leaf[(size_t)32 * n + i] = (int32_t)(sign * (1 - (int64_t)(s[0][1] & MASK31) + (int64_t)(s[1][1] & MASK31)));
  }
}
