/* primia_b200 -- C ABI of the B200-native hot paths of PriMIA.
 *
 * Every entry point takes raw DEVICE pointers (unless the name ends in _host), explicit sizes
 * and a CUDA stream (cudaStream_t passed as void*), never allocates, and returns 0 on success
 * or a PM_E* code (cudaGetLastError is folded into PM_ECUDA).  Callers own all buffers.
 *
 * The reference (gkaissis/PriMIA) has no FFI: its seam is Python operator overloading by dotted
 * name.  Each group below cites the reference function(s) whose arithmetic it
 * replaces (paths relative to the reference checkout).  The Python host side
 * (primia_b200/ring, primia_b200/train, primia_b200/sy) mirrors those names on top of this ABI.
 */
#ifndef PRIMIA_B200_H
#define PRIMIA_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PM_OK 0
#define PM_EINVAL 1  /* bad shape / argument                      */
#define PM_ECUDA 2   /* CUDA runtime error (see pm_last_error)    */
#define PM_ERANGE 3  /* fixed-point encode overflow (precision.py:122-127 assert) */

typedef void* pm_stream_t; /* cudaStream_t */

const char* pm_last_error(void);
int pm_version(void);
int pm_device_sm_count(int* out);

/* ===================================================================== path E : ring Z_2^64 */

/* FixedPrecisionTensor.fix_precision  syft/frameworks/torch/tensors/interpreters/precision.py:117-132
 * q = (int64) trunc( x * scale_f32 ); *overflow (device int, may be NULL) is set to 1 when |x*scale| >= 2^63. */
int pm_encode_f32_i64(const float* x, float scale, int64_t* q, size_t n, int* overflow, pm_stream_t s);
/* FixedPrecisionTensor.float_precision  precision.py:134-144 : x = (float) q / scale_f32 */
int pm_decode_i64_f32(const int64_t* q, float scale, float* x, size_t n, pm_stream_t s);

/* AdditiveSharingTensor.generate_shares (n_workers == 2)  additive_shared.py:336-365
 * s0 ~ Philox4x32-10(seed, offset) over [-2^63, 2^63-2]; s1 = q - s0 (wraps).
 * epoch (device uint64, may be NULL): the stream actually used is offset + (*epoch << 32) -- a generation captured in a CUDA
 * graph then draws fresh randomness at every replay once a pm_epoch_bump node precedes it. */
int pm_share_gen_i64(const int64_t* q, uint64_t seed, uint64_t offset, const uint64_t* epoch, int64_t* s0, int64_t* s1, size_t n,
                     pm_stream_t s);
/* randint(-2^63, 2^63-1) of build_triple  syft/frameworks/torch/mpc/beaver.py:32-34 */
int pm_random_i64(uint64_t seed, uint64_t offset, const uint64_t* epoch, int64_t* out, size_t n, pm_stream_t s);
int pm_epoch_bump(uint64_t* epoch, pm_stream_t s);

/* _pre_conv im2col  syft/frameworks/torch/nn/functional.py:79-166 (groups==1)
 * x [B,C,H,W] -> im [B, M=Ho*Wo, K=C*kh*kw], k = ch*kh*kw + r*kw + c, zero padding. */
int pm_im2col_i64(const int64_t* x, int B, int C, int H, int W, int kh, int kw, int stride, int pad, int dil,
                  int64_t* im, pm_stream_t s);
/* spdz_mask  syft/frameworks/torch/mpc/spdz.py:22-45 : delta_j = x_j - a_j (elementwise, same shape) */
int pm_spdz_mask_i64(const int64_t* x, const int64_t* a, int64_t* delta, size_t n, pm_stream_t s);
/* im2col-free fusion of _pre_conv + spdz_mask for the left operand: delta_j = im2col(x_j) - a_j */
int pm_spdz_mask_im2col_i64(const int64_t* x, int B, int C, int H, int W, int kh, int kw, int stride, int pad,
                            int dil, const int64_t* a, int64_t* delta, pm_stream_t s);
/* Elementwise Beaver product of two same-shape operands (AdditiveSharingTensor.mul -> spdz_mul, spdz.py:125-197; the ReLU and
 * max-pool selects) in two launches per party: spdz_mask of both operands in one pass, then spdz_compute with the two openings
 * fused -- delta = d_own + d_peer and eps = e_own + e_peer are formed in registers (peer pointers may be peer-mapped). */
int pm_spdz_mask2_i64(const int64_t* x, const int64_t* a, const int64_t* y, const int64_t* b, int64_t* delta, int64_t* eps, size_t n,
                      pm_stream_t s);
int pm_spdz_combine_mul_open_i64(int j, const int64_t* d_own, const int64_t* d_peer, const int64_t* e_own, const int64_t* e_peer,
                                 const int64_t* a, const int64_t* b, const int64_t* c, size_t n, int64_t* z, pm_stream_t s);
/* right operand: weight_reshaped = w.reshape(Cout,-1).t() (functional.py:157) fused with spdz_mask:
 * eps_j[k,n] = w_j[n*K + k] - b_j[k*N + n] */
int pm_spdz_mask_wt_i64(const int64_t* w, int N, int K, const int64_t* b, int64_t* eps, pm_stream_t s);
/* opening delta = delta_0 + delta_1  spdz.py:162-163.  `peer` may be a peer-mapped pointer of another GPU
 * (NVLink P2P load) -- replaces the orchestrator hub. */
int pm_open_add_i64(const int64_t* local, const int64_t* peer, int64_t* out, size_t n, pm_stream_t s);

/* spdz_compute, op == "matmul"  spdz.py:64-122 (+ triple_mat_mul :54-59):
 *   z_j = delta@b_j + a_j@eps + c_j (+ delta@eps if j == 0)   all mod 2^64
 * delta,a_j [B,M,K]; eps,b_j [K,N]; c_j,z_j [B,M,N].  z_j must not alias inputs. `ws` is a [K,N] int64
 * scratch (used when j == 0 for b_j + eps: delta@b + delta@eps == delta@(b+eps) exactly in the ring). */
int pm_spdz_combine_matmul_i64(int j, const int64_t* delta, const int64_t* eps, const int64_t* a,
                               const int64_t* b, const int64_t* c, int B, int M, int K, int N, int64_t* ws,
                               int64_t* z, pm_stream_t s);
/* The same contraction on the INT8 tensor cores (exact): every int64 is split into 8 unsigned byte limbs and
 * C = Cinit + A1@B1 (+ A2@B2) mod 2^64 is evaluated as 36 u8 x u8 -> s32 limb-pair GEMMs (tcgen05.mma.kind::i8) whose 14
 * accumulators are recombined with shifts in the epilogue (primia_b200/csrc/ring_i8.cu).  Replaces triple_mat_mul /
 * spdz_compute's three th.matmul calls (spdz.py:54-59,90-122) for N % 32 == 0.  rows = B*M (the batch is folded);
 * A1,A2 [rows,K]; B1,B2 [K,N]; Cinit,C [rows,N]; ws: pm_ring_tc_ws_bytes(...) bytes of scratch for the limb planes. */
size_t pm_ring_tc_ws_bytes(int rows, int K, int N, int nseg);
int pm_ring_tc_supported(int rows, int K, int N);
int pm_ring_gemm2_tc_i64(const int64_t* A1, const int64_t* B1, const int64_t* A2, const int64_t* B2, const int64_t* Cinit, int rows,
                         int K, int N, void* ws, int64_t* C, pm_stream_t s);
/* The same GEMM in pieces.  In spdz_mul (spdz.py:125-197) only delta = x - a depends on the image: the triple's a and b and the
 * opened weight mask eps = w - b are known as soon as the triple exists, so their limb planes are built once in the OFFLINE
 * phase (ring/functional.py prepare_weight_side) and the online phase of a layer is  mask -> open+planarise(delta) -> GEMM ->
 * truncate.  pm_ring_planarize_rows_i64 with peer != NULL fuses the opening (spdz.py:162-163: delta = sum of the parties'
 * masked shares; peer may be a peer-mapped pointer into the other party's GPU): planes of (A + peer) mod 2^64.
 * planes: pm_ring_planes_bytes(n, K) bytes, n = rows (left operand) or N (right operand). */
size_t pm_ring_planes_bytes(int n, int K);
int pm_ring_planarize_rows_i64(const int64_t* A, const int64_t* peer, int rows, int K, void* planes, pm_stream_t s);
int pm_ring_planarize_cols_i64(const int64_t* B, int K, int N, void* planes, pm_stream_t s);
int pm_ring_gemm_planes_i64(const void* pa1, const void* pb1, const void* pa2, const void* pb2, const int64_t* Cinit, int rows, int K,
                            int N, int64_t* C, pm_stream_t s);
/* spdz_compute, op == "mul" with torch broadcasting of a [C]-vector against [P,C]:
 * mode 0: same shape n ; mode 1: left is [C], right is [P,C] ; mode 2: left is [P,C], right is [C]. */
int pm_spdz_combine_mul_i64(int j, const int64_t* delta, const int64_t* eps, const int64_t* a, const int64_t* b,
                            const int64_t* c, int mode, size_t P, size_t C, int64_t* z, pm_stream_t s);
/* build_triple c = a @ b  beaver.py:36-52 ; also the plain ring matmul. C[B,M,N] = A[B,M,K] @ Bm[K,N] */
int pm_matmul_i64(const int64_t* A, const int64_t* Bm, int B, int M, int K, int N, int64_t* C, pm_stream_t s);

/* FixedPrecisionTensor.truncate -> AdditiveSharingTensor._public_div  precision.py:146-154,
 * additive_shared.py:673-678: per-share C-style truncating divide. In place allowed. */
int pm_trunc_div_i64(const int64_t* x, int64_t divisor, int64_t* out, size_t n, pm_stream_t s);
/* truncate fused with _post_conv  functional.py:170-201: z [B,M,N] -> out [B,N,Ho*Wo] (= NCHW), (+bias[N]) */
int pm_trunc_post_conv_i64(const int64_t* z, int64_t divisor, const int64_t* bias, int B, int M, int N,
                           int64_t* out, pm_stream_t s);
/* share-wise linear ops: out = alpha*x + beta*y (+gamma)  (AST add/sub/public mul, additive_shared.py:455-588).
 * y may be NULL. ybcast: 0 same shape, 1 y is a length-C vector broadcast over rows of x [P,C], 2 y is a scalar[1]. */
int pm_axpby_i64(int64_t alpha, const int64_t* x, int64_t beta, const int64_t* y, int ybcast, size_t P, size_t C,
                 int64_t* out, pm_stream_t s);
/* FixedPrecisionTensor.reciprocal(method="newton")  precision.py:507-518 as used by batch_norm (nn/functional.py:62-64):
 * x0 = (C+1 - v)/C ; x <- x*(C+1 - v*x*x)/C, iters-1 more times, every product a Beaver multiplication followed by
 * truncate(divisor = base**pf) and "/ C" a per-share truncating divide.  This entry point runs the whole iteration for
 * both share holders when they are resident on the SAME device (the openings stay in registers), for up to
 * PM_NEWTON_MAX_JOBS vectors (BatchNorm layers) in one launch.  Per job: v*: [C] shares of the variance, a*,b*,c*:
 * [3*(iters-1)][C] triple shares in consumption order, k*: [iters] shares of encode(C+1), x*: [C] outputs.
 * `jobs` is a HOST array (copied into the kernel's parameter space). */
#define PM_NEWTON_MAX_JOBS 32
typedef struct {
  const int64_t *v0, *v1, *a0, *b0, *c0, *a1, *b1, *c1, *k0, *k1;
  int64_t *x0, *x1;
  int C;
} pm_newton_job_t;
int pm_bn_newton_fused_i64(const pm_newton_job_t* jobs, int n_jobs, int iters, int64_t divisor, int64_t newton_c,
                           pm_stream_t s);
/* The same Newton iteration when the two share holders sit on DIFFERENT GPUs: each party calls this
 * once, on its own device and stream, with ITS shares only; the 3*(iters-1) openings per channel travel over NVLink through
 * mailboxes: `inbox` is this party's (local) mailbox, `peer_inbox` the other party's mailbox as a peer-mapped pointer, both
 * pm_bn_newton_p2p_mailbox_bytes(n_jobs, iters, max_channels) bytes, zero-initialised once; `epoch` (device uint64, zero-
 * initialised once, one per party) is bumped by every call so messages of successive images are told apart (CUDA-graph
 * replays included); `err` (device int) is set to 1 if the peer's messages never arrive (the peer kernel was not launched).
 * The two calls must not be stream-ordered after one another: the kernels exchange data while both are running.
 * Per job: v,x [C]; a,b,c [3*(iters-1)][C]; k [iters] -- this party's shares. */
typedef struct {
  const int64_t *v, *a, *b, *c, *k;
  int64_t* x;
  int C;
} pm_newton_p2p_job_t;
size_t pm_bn_newton_p2p_mailbox_bytes(int n_jobs, int iters, int max_channels);
int pm_bn_newton_p2p_i64(int party, const pm_newton_p2p_job_t* jobs, int n_jobs, int iters, int64_t divisor, int64_t newton_c,
                         void* inbox, void* peer_inbox, uint64_t* epoch, int max_channels, int* err, pm_stream_t s);
/* avg pool k x k, stride k on one share: sum / (k*k) with trunc  functional.py:460-525 + additive_shared.py:720-729 */
int pm_avgpool_i64(const int64_t* x, int B, int C, int H, int W, int k, int64_t* out, pm_stream_t s);
/* One elementwise pass of a hoisted Beaver product with a per-channel operand on NCHW shares (nn/functional.py:44-75 batch_norm:
 * x * (flat - mean), normalized * weight + bias).  With the operand that depends only on the model opened in the offline phase,
 * spdz_compute (spdz.py:64-122) collapses to  z_j = s_j[c] * open(masked)[i] + d_j[i]  and the whole layer is three passes per party:
 *   out[i] = T( sc[c] * (u[i] + peer[i]) + add[i] ) + cs * chan[c] + es * elem[i],  c = (i / HW) % C,
 * T = C-style division by div (precision.py:146-160) when div > 1.  NULL operands drop out; cs, es in {-1, 0, +1}. */
int pm_spdz_affine_i64(const int64_t* u, const int64_t* peer, const int64_t* sc, const int64_t* add, int64_t div, const int64_t* chan,
                       int cs, const int64_t* elem, int es, int C, int HW, size_t n, int64_t* out, pm_stream_t s);
/* batch_norm layout shuffles functional.py:52-55,70-73: NCHW [B,C,H,W] <-> [P=B*H*W, C] */
int pm_nchw_to_pc_i64(const int64_t* x, int B, int C, int HW, int64_t* out, pm_stream_t s);
int pm_pc_to_nchw_i64(const int64_t* x, int B, int C, int HW, int64_t* out, pm_stream_t s);

/* ----- function secret sharing: the comparison behind ReLU and max-pool on shares (protocol="fss").
 * Key material of n comparison instances, structure-of-arrays with row stride `stride` (>= n) so that a store can hand out
 * sub-ranges of a larger pool:  s0 [2][stride] (party's root seed, 2 x uint64), bits [32][stride] (uint8: tauL | tL<<1 |
 * tauR<<2 | tR<<3, the compressed correction bits fss.py:431-455), sigma_cw [32][2][stride], s_cw [32][2][stride],
 * leaf [33][stride] int32. */
/* the PRG H  syft/frameworks/torch/mpc/fss.py:553-601 : SHA-512 of each 16-byte seed (seed [2][n]) -> out [8][n], the digest
 * as little-endian uint64 words (the reference calls the external `shaloop.sha512_loop_func`). */
int pm_fss_prg_sha512(const uint64_t* seed, size_t n, uint64_t* out, pm_stream_t s);
/* DIF.keygen  fss.py:341-399.  alpha [n] (values < 2^32), seeds [2 parties][2 words][n] (word 0 < 2^63, randbit :498-505).
 * Both parties receive the same correction words; party b's key is (seeds[b], bits, sigma_cw, s_cw, leaf). */
int pm_fss_dif_keygen(const uint64_t* alpha, const uint64_t* seeds, size_t n, size_t stride, uint8_t* bits,
                      uint64_t* sigma_cw, uint64_t* s_cw, int32_t* leaf, pm_stream_t s);
/* DIF.eval  fss.py:401-428 (reached through evaluate / comp_evaluate :208-275): out[i] = party b's int64 share of
 * [ (x_masked[i] mod 2^32) <= alpha[i] ]. */
int pm_fss_dif_eval(int b, const int64_t* x_masked, const uint64_t* s0, const uint8_t* bits, const uint64_t* sigma_cw,
                    const uint64_t* s_cw, const int32_t* leaf, size_t n, size_t stride, int64_t* out, pm_stream_t s);
/* the same evaluation with the opening of the masked difference fused (fss.py:158): x = (r_own + r_peer) mod 2^32 is formed in
 * the kernel; r_peer may be a peer-mapped pointer into the other party's GPU */
int pm_fss_dif_eval_open(int b, const int64_t* r_own, const int64_t* r_peer, const uint64_t* s0, const uint8_t* bits,
                         const uint64_t* sigma_cw, const uint64_t* s_cw, const int32_t* leaf, size_t n, size_t stride, int64_t* out,
                         pm_stream_t s);
/* mask_builder  fss.py:189-204 : r_j = x1_j - x2_j + alpha_j ; x1 or x2 may be NULL (a public 0 operand) */
int pm_fss_mask_i64(const int64_t* x1, const int64_t* x2, const int64_t* alpha_share, int64_t* r, size_t n, pm_stream_t s);
/* opening of the masked difference  fss.py:158 : out = (local + peer) mod 2^32 ; `peer` may be peer-mapped (NVLink) */
int pm_fss_open_mod32_i64(const int64_t* local, const int64_t* peer, int64_t* out, size_t n, pm_stream_t s);
/* conditions raw 64-bit random words into keygen's randomness (fss.py:346,354,498-505; primitives.py:245-251), in place:
 * alpha, mask -> [0,2^32); seeds [2][2][n] word 0 -> [0,2^63); alpha0 = (alpha - mask) mod 2^32 (party 0's share of alpha,
 * party 1 holds mask). */
int pm_fss_condition_randomness(uint64_t* alpha, uint64_t* mask, uint64_t* seeds, int64_t* alpha0, size_t n, pm_stream_t s);
/* _pre_pool  syft/frameworks/torch/nn/functional.py:312-390 : x [B,C,H,W] -> [B,C,Ho*Wo,k*k] (zero padding) */
int pm_pre_pool_i64(const int64_t* x, int B, int C, int H, int W, int k, int stride, int pad, int64_t* out, pm_stream_t s);
/* t[..., start:start+len] of a [rows, L] tensor made contiguous (max_half_split halves, functional.py:489-508) */
int pm_slice_lastdim_i64(const int64_t* src, size_t rows, int L, int start, int len, int64_t* dst, pm_stream_t s);

/* ===================================================================== path T : float training */
/* Layout: activations NHWC; conv weights KRSC ([Cout][kh][kw][Cin]); fp32 ("_f32", parity mode) or
 * bf16 activations with fp32 accumulation ("_bf16", throughput mode). Replaces the ATen CPU ops the
 * worker executes for torchlib/models.py:466-482,268-284 via syft/workers/message_handler.py:105-118. */

int pm_nchw_to_nhwc_f32(const float* x, int B, int C, int H, int W, float* out, pm_stream_t s);
int pm_nchw_to_nhwc_f32_bf16(const float* x, int B, int C, int H, int W, int Cpad, void* out, pm_stream_t s);
int pm_f32_to_bf16(const float* x, void* out, size_t n, pm_stream_t s);

typedef struct {
  int B, H, W, C;      /* input  NHWC */
  int K, R, S;         /* filters: Cout, kh, kw */
  int stride, pad;
  int Ho, Wo;          /* output spatial */
} pm_conv_t;

/* conv2d forward: y[B,Ho,Wo,K] = x (*) w   (F.conv2d, models.py:219-235,379-381) */
int pm_conv_fwd_f32(const pm_conv_t* p, const float* x, const float* w, float* y, pm_stream_t s);
/* data gradient: dx[B,H,W,C] (= or +=) dy (*)^T w ; accumulate != 0 adds into dx */
int pm_conv_dgrad_f32(const pm_conv_t* p, const float* dy, const float* w, float* dx, int accumulate, pm_stream_t s);
/* weight gradient: dw[K,R,S,C] = sum_pix dy x ; ws: scratch of pm_conv_wgrad_ws_bytes() bytes */
size_t pm_conv_wgrad_ws_bytes(const pm_conv_t* p);
int pm_conv_wgrad_f32(const pm_conv_t* p, const float* x, const float* dy, float* dw, void* ws, pm_stream_t s);

/* bf16 tensor-core (tcgen05 / TMEM) implicit-GEMM versions; x,w,y bf16; optional fused per-channel
 * sum / sum-of-squares of the fp32 accumulators into stats[2*K] (doubles) for BatchNorm. */
int pm_conv_fwd_bf16(const pm_conv_t* p, const void* x, const void* w, void* y, double* stats, pm_stream_t s);
int pm_conv_dgrad_bf16(const pm_conv_t* p, const void* dy, const void* wt, void* dx, int accumulate, pm_stream_t s);
/* NB: the bf16 wgrad ACCUMULATES into dw (fp32 red.add over pixel splits): the caller zeroes dw (the engine clears the whole
 * flat gradient buffer once per step). */
int pm_conv_wgrad_bf16(const pm_conv_t* p, const void* x, const void* dy, float* dw, void* ws, pm_stream_t s);
/* ResNet stem (conv 7x7 / stride 2 / pad 3, 3 -> 64; models.py:379,468) as a direct tcgen05 implicit GEMM over the fp32 NCHW
 * batch -- no im2col matrix (conv_stem.cu).  w192: bf16 [64][192] produced by pm_stem_prep_w_bf16 from the fp32 KRSC master
 * [64][7][7][3].  fwd: y bf16 NHWC [B,Ho,Wo,64], optional fused BatchNorm statistics stats[128] (sum, sum of squares; doubles,
 * accumulated).  wgrad: dw fp32 KRSC [64][7][7][3] += sum_pix dy x (fp32 atomics: the caller zeroes dw). */
int pm_stem_prep_w_bf16(const float* w_krsc, void* w192, pm_stream_t s);
int pm_stem_conv_fwd_bf16(const float* x_nchw, const void* w192, int B, int H, int W, void* y, double* stats, pm_stream_t s);
int pm_stem_conv_wgrad_bf16(const float* x_nchw, const void* dy, int B, int H, int W, float* dw_krsc, pm_stream_t s);

/* diagnostics for the halo-strip 3x3 kernel (conv_halo.cu): enable != 0 makes the next launches record per-CTA wait / busy
 * cycle counters; out_host (may be NULL) receives the [160][8] int64 counters of the last profiled launch, then the event
 * count and up to 256 (code, clock) pairs of CTA 0's trace: 160*8 + 1 + 512 int64 in all. */
int pm_halo_prof(int enable, int64_t* out_host);

/* BatchNorm2d (training) -- F.batch_norm, models.py:261,264,382 ; P = B*H*W rows of C channels.
 * stats: [2*C] doubles (sum, sumsq), zeroed by the caller. */
int pm_bn_stats_f32(const float* x, size_t P, int C, double* stats, pm_stream_t s);
int pm_bn_stats_bf16(const void* x, size_t P, int C, double* stats, pm_stream_t s);
/* mean/invstd from stats; running_mean/var momentum update (unbiased var), eps */
int pm_bn_finalize(const double* stats, size_t P, int C, float eps, float momentum, float* mean, float* invstd,
                   float* running_mean, float* running_var, pm_stream_t s);
/* y = act( (x-mean)*invstd*gamma + beta (+ residual) ) ; relu != 0 applies ReLU */
int pm_bn_apply_f32(const float* x, const float* mean, const float* invstd, const float* gamma, const float* beta,
                    const float* residual, int relu, size_t P, int C, float* y, pm_stream_t s);
int pm_bn_apply_bf16(const void* x, const float* mean, const float* invstd, const float* gamma, const float* beta,
                     const void* residual, int relu, size_t P, int C, void* y, pm_stream_t s);
/* fused finalize + apply (training): mean/invstd are derived from `stats` inside the kernel (one launch instead of two);
 * block 0 also stores mean/invstd (needed by the backward pass) and updates the running statistics. */
int pm_bn_fwd_fused_f32(const float* x, const double* stats, size_t P, int C, float eps, float momentum,
                        const float* gamma, const float* beta, const float* residual, int relu, float* y, float* mean,
                        float* invstd, float* running_mean, float* running_var, pm_stream_t s);
int pm_bn_fwd_fused_bf16(const void* x, const double* stats, size_t P, int C, float eps, float momentum,
                         const float* gamma, const float* beta, const void* residual, int relu, void* y, float* mean,
                         float* invstd, float* running_mean, float* running_var, pm_stream_t s);
/* backward: g = dy * (y_out > 0 if relu_mask) ; sums[0..C) = sum g , sums[C..2C) = sum g*xhat (doubles, zeroed by caller).
 * If g_out != NULL the masked gradient is also written (used for the identity branch). */
int pm_bn_bwd_reduce_f32(const float* dy, const float* y_out, const float* x, const float* mean, const float* invstd,
                         size_t P, int C, double* sums, float* g_out, pm_stream_t s);
int pm_bn_bwd_reduce_bf16(const void* dy, const void* y_out, const void* x, const float* mean, const float* invstd,
                          size_t P, int C, double* sums, void* g_out, pm_stream_t s);
/* dx = gamma*invstd*( g - sum_g/P - xhat*sum_gxhat/P ) ; dgamma = sum_gxhat ; dbeta = sum_g (written into fp32 grads) */
int pm_bn_bwd_apply_f32(const float* dy, const float* y_out, const float* x, const float* mean, const float* invstd,
                        const float* gamma, const double* sums, size_t P, int C, float* dx, float* dgamma,
                        float* dbeta, pm_stream_t s);
int pm_bn_bwd_apply_bf16(const void* dy, const void* y_out, const void* x, const float* mean, const float* invstd,
                         const float* gamma, const double* sums, size_t P, int C, void* dx, float* dgamma,
                         float* dbeta, pm_stream_t s);

/* BatchNorm backward in ONE launch (reduce -> grid barrier -> apply; deterministic, no floating-point atomics).
 * g = dy * (y_out > 0) when y_out != NULL; if g_out != NULL the masked gradient is also written (identity branch).
 * ws: pm_bn_bwd_fused_ws_doubles(C) doubles, ZERO-INITIALISED ONCE by the caller (the first 16 bytes are the barrier state,
 * which the kernel leaves ready for the next launch, CUDA-graph replays included). g_out may alias dy; dx must not alias g_out. */
size_t pm_bn_bwd_fused_ws_doubles(int C);
int pm_bn_bwd_fused_f32(const float* dy, const float* y_out, const float* x, const float* mean, const float* invstd,
                        const float* gamma, size_t P, int C, double* ws, float* g_out, float* dx, float* dgamma, float* dbeta,
                        pm_stream_t s);
int pm_bn_bwd_fused_bf16(const void* dy, const void* y_out, const void* x, const float* mean, const float* invstd,
                         const float* gamma, size_t P, int C, double* ws, void* g_out, void* dx, float* dgamma, float* dbeta,
                         pm_stream_t s);

/* bf16 throughput mode, BN followed by ReLU with no residual: the ReLU decision is recomputed from x and the BN constants
 * ((x - mean) * (invstd * gamma) + beta > 0, the forward's own evaluation) instead of reading the stored activation. */
int pm_bn_bwd_fused_xmask_bf16(const void* dy, const void* x, const float* mean, const float* invstd, const float* gamma,
                               const float* beta, size_t P, int C, double* ws, void* dx, float* dgamma, float* dbeta,
                               pm_stream_t s);
/* Stem forward in one pass (bf16 mode): BatchNorm2d (batch statistics from `stats`, running stats updated, mean/invstd
 * written) + ReLU + MaxPool2d(3,2,1) of x [B,H,W,C] -> y [B,Ho,Wo,C]; the full-resolution activation is not materialised.
 * idx: argmax per output (first strict maximum of the raw conv outputs, sign-adjusted per channel: BN+ReLU is monotonic), or
 * 255 where the maximum is not positive (no gradient passes the ReLU there).  xmax (bf16 [B,Ho,Wo,C], may be NULL): the RAW
 * conv output at each window's argmax, kept so that the backward's batch sums are a reduction over the pooled tensors only. */
int pm_bn_relu_maxpool_fwd_bf16(const void* x, const double* stats, int B, int H, int W, int C, float eps, float momentum,
                                const float* gamma, const float* beta, void* y, uint8_t* idx, void* xmax, float* mean, float* invstd,
                                float* running_mean, float* running_var, pm_stream_t s);

/* Backward of pm_bn_relu_maxpool_fwd_bf16 (H, W even): max-pool backward gathered from (dpool, pool_idx) + BatchNorm backward
 * as a reduce launch and an apply launch over 2x2 input blocks; sums: [2*C] doubles ZEROED by the caller; dx [B,H,W,C]. */
int pm_stem_pool_bn_bwd_bf16(const void* dpool, const uint8_t* pool_idx, const void* xmax, int B, int H, int W, const void* x,
                             const float* mean, const float* invstd, const float* gamma, int C, double* sums, void* dx, float* dgamma,
                             float* dbeta, pm_stream_t s);

/* Stem variant: the BN input gradient is MaxPool2d(3,2,1)'s backward of `dpool` [B,Ho,Wo,C], gathered from the stored
 * argmax inside the reduce pass (saves the separate max-pool-backward launch and one pass over the full-resolution
 * gradient); `g_scratch` [B,H,W,C] receives the gathered+masked gradient for the apply pass (NULL: the apply pass gathers
 * again instead); x / y_out / dx are [B,H,W,C]; y_out may be NULL when `pool_idx` already encodes the ReLU decision (255 marks
 * of pm_bn_relu_maxpool_fwd_bf16). */
int pm_bn_bwd_fused_pool_f32(const float* dpool, const uint8_t* pool_idx, int B, int H, int W, const float* y_out, const float* x,
                             const float* mean, const float* invstd, const float* gamma, int C, double* ws, float* g_scratch,
                             float* dx, float* dgamma, float* dbeta, pm_stream_t s);
int pm_bn_bwd_fused_pool_bf16(const void* dpool, const uint8_t* pool_idx, int B, int H, int W, const void* y_out, const void* x,
                              const float* mean, const float* invstd, const float* gamma, int C, double* ws, void* g_scratch,
                              void* dx, float* dgamma, float* dbeta, pm_stream_t s);

/* MaxPool2d(3,2,1) / AvgPool2d(3,2,1) (models.py:384-389) on NHWC; idx: uint8 argmax per output (first max wins) */
int pm_maxpool3s2_fwd_f32(const float* x, int B, int H, int W, int C, float* y, uint8_t* idx, pm_stream_t s);
int pm_maxpool3s2_bwd_f32(const float* dy, const uint8_t* idx, int B, int H, int W, int C, float* dx, pm_stream_t s);
int pm_maxpool3s2_fwd_bf16(const void* x, int B, int H, int W, int C, void* y, uint8_t* idx, pm_stream_t s);
int pm_maxpool3s2_bwd_bf16(const void* dy, const uint8_t* idx, int B, int H, int W, int C, void* dx, pm_stream_t s);
/* AvgPool2d(3,2,1) (pooling="avg", models.py:386-387; zero padding counted, divisor 9): y [B,Ho,Wo,C]; backward dx [B,H,W,C] */
int pm_avgpool3s2_fwd_f32(const float* x, int B, int H, int W, int C, float* y, pm_stream_t s);
int pm_avgpool3s2_bwd_f32(const float* dy, int B, int H, int W, int C, float* dx, pm_stream_t s);
int pm_avgpool3s2_fwd_bf16(const void* x, int B, int H, int W, int C, void* y, pm_stream_t s);
int pm_avgpool3s2_bwd_bf16(const void* dy, int B, int H, int W, int C, void* dx, pm_stream_t s);
/* AvgPool2d(HW) -> [B,C] (models.py:400-404,477) and its backward (broadcast / HW) */
int pm_gap_fwd_f32(const float* x, int B, int HW, int C, float* y, pm_stream_t s);
int pm_gap_bwd_f32(const float* dy, int B, int HW, int C, float* dx, pm_stream_t s);
int pm_gap_fwd_bf16(const void* x, int B, int HW, int C, float* y, pm_stream_t s);
int pm_gap_bwd_bf16(const float* dy, int B, int HW, int C, void* dx, pm_stream_t s);

/* Linear(F,ncls) + CrossEntropy (hard labels w/ optional class weights: nn.CrossEntropyLoss(weight,"mean");
 * or soft targets: Cross_entropy_one_hot torchlib/utils.py:404-441), forward AND backward:
 * logits[B,ncls]; loss[1]; dfeat[B,F]; dW[ncls,F]; db[ncls].  labels (int64) or soft (float [B,ncls]) -- one is NULL.
 * ws: scratch of B*(ncls+1)+1 floats. */
int pm_linear_ce_f32(const float* feat, const float* W, const float* bias, const int64_t* labels, const float* soft,
                     const float* class_w, int B, int F, int ncls, float* logits, float* loss, float* dfeat,
                     float* dW, float* db, float* ws, pm_stream_t s);
/* The head of a training step split by what the backward chain needs.  pm_head_fused_*: one launch (one block per sample) does
 * AvgPool2d(HW) of the last activation x [B,HW,F] -> feat, logits, the per-sample CE gradient (left unnormalised in ws, as
 * pm_linear_ce_f32 leaves it: ws[b*(ncls+1) + j], nll in slot ncls, denominator in ws[B*(ncls+1)]), dfeat and the gradient of the
 * average pool written straight into d_out [B,HW,F].  pm_head_grads_f32: the batch reductions dW, db, loss from (ws, feat) -- not
 * on the critical path, the engine issues it on its side stream.  ws: B*(ncls+1)+1 floats. */
int pm_head_fused_f32(const float* x, int HW, const float* W, const float* bias, const int64_t* labels, const float* soft,
                      const float* class_w, int B, int F, int ncls, float* feat, float* logits, float* ws, float* dfeat, float* d_out,
                      pm_stream_t s);
int pm_head_fused_bf16(const void* x, int HW, const float* W, const float* bias, const int64_t* labels, const float* soft,
                       const float* class_w, int B, int F, int ncls, float* feat, float* logits, float* ws, float* dfeat, void* d_out,
                       pm_stream_t s);
int pm_head_grads_f32(const float* feat, int B, int F, int ncls, const float* ws, float* loss, float* dW, float* db, pm_stream_t s);
/* inference-only head: logits = feat @ W^T + bias */
int pm_linear_fwd_f32(const float* feat, const float* W, const float* bias, int B, int F, int ncls, float* logits,
                      pm_stream_t s);

/* torch.optim.Adam.step (train.py:280-303; L2 weight decay added to grad) on the flat parameter buffer */
int pm_adam_step_f32(float* p, const float* g, float* m, float* v, size_t n, float lr, float beta1, float beta2,
                     float eps, float weight_decay, int step, pm_stream_t s);
/* the same step for a freshly created optimizer (step index 1, m = v = 0: the reference re-creates the optimizers after every
 * aggregation, utils.py:1209-1218): bit-identical to pm_adam_step_f32 on zeroed moments, but m and v are only WRITTEN, so the
 * caller never has to clear them */
int pm_adam_first_step_f32(float* p, const float* g, float* m, float* v, size_t n, float lr, float beta1, float beta2,
                           float eps, float weight_decay, pm_stream_t s);
/* torch.optim.SGD.step (no momentum; weight decay) */
int pm_sgd_step_f32(float* p, const float* g, size_t n, float lr, float weight_decay, pm_stream_t s);
/* FedAvg tail (torchlib/utils.py:1078-1090): x *= scale (after the NCCL sum) ; and pre-scale for weighted averaging */
int pm_scale_f32(float* x, float scale, size_t n, pm_stream_t s);
/* KCRS (torch) <-> KRSC (ours) weight relayout, and the dgrad operand [C][R][S][K] in bf16 */
int pm_kcrs_to_krsc_f32(const float* w, int K, int C, int R, int S, float* out, pm_stream_t s);
int pm_krsc_to_kcrs_f32(const float* w, int K, int C, int R, int S, float* out, pm_stream_t s);
int pm_krsc_to_bf16_fwd_dgrad(const float* w, int K, int C, int R, int S, int Cpad, void* w_fwd, void* w_dgrad,
                              pm_stream_t s);
/* the same conversion for a whole model in ONE launch: `table` is a DEVICE array of n entries */
typedef struct {
  const float* w;   /* fp32 master  [K][RS][C]            */
  void* w_fwd;      /* bf16         [K][RS][Cpad]         */
  void* w_dgrad;    /* bf16         [C][RS][K]  or NULL   */
  int K, C, RS, Cpad;
} pm_wcvt_t;
/* max_tiles = max over entries of RS * ceil(K/32) * ceil(Cpad/32) (grid.x; smaller entries exit early) */
int pm_krsc_to_bf16_batched(const pm_wcvt_t* table, int n, int max_tiles, pm_stream_t s);
/* same, one block per existing tile: total_tiles = sum over entries of RS * ceil(K/32) * ceil(Cpad/32) */
int pm_krsc_to_bf16_batched_exact(const pm_wcvt_t* table, int n, int total_tiles, pm_stream_t s);
/* stem im2col for the bf16 path: x NCHW fp32 [B,Cin,H,W] -> [B,Ho,Wo,Kpad] bf16 with k = (r*S+s)*Cin + c (zeros for
 * k >= R*S*Cin), so that the 7x7/stride-2 stem becomes a dense 1x1 problem for the TMA-fed tensor-core kernels */
int pm_im2col_stem_bf16(const float* x, int B, int Cin, int H, int W, int R, int stride, int pad, int Kpad, void* out,
                        pm_stream_t s);

/* ----- DP-SGD (train.py:304-334: pytorch-dp PrivacyEngine(noise_multiplier, max_grad_norm).attach(optimizer)).
 * Per-sample weight gradients dw[b][K][R*S*C] of one conv (written, not accumulated); bf16: tcgen05 kernel with the sample as
 * grid.z; f32: the deterministic parity kernel once per image (ws as pm_conv_wgrad_f32). */
/* norm2 (may be NULL): norm2[b] += |dw[b]|^2, accumulated in the epilogue (saves the separate pm_dp_sqnorm_f32 pass) */
int pm_conv_wgrad_persample_bf16(const pm_conv_t* p, const void* x, const void* dy, float* dw, double* norm2, pm_stream_t s);
int pm_conv_wgrad_persample_f32(const pm_conv_t* p, const float* x, const float* dy, float* dw, void* ws, pm_stream_t s);
/* BatchNorm as a frozen per-channel affine map (running statistics; DP needs per-sample gradients, which batch statistics do
 * not have -- the reference refuses BatchNorm models under DP, train.py:306-310).  g = dy * (y_out > 0) (y_out may be NULL),
 * g_out (may be NULL) receives g, dx = g * gamma * invstd. */
int pm_bn_eval_bwd_f32(const float* dy, const float* y_out, const float* gamma, const float* invstd, size_t P, int C, float* g_out,
                       float* dx, pm_stream_t s);
int pm_bn_eval_bwd_bf16(const void* dy, const void* y_out, const float* gamma, const float* invstd, size_t P, int C, void* g_out,
                        void* dx, pm_stream_t s);
/* per-sample dgamma[b][c] = sum_pix g * xhat, dbeta[b][c] = sum_pix g over image b's HW pixels (outputs are cleared here) */
int pm_bn_persample_param_grads_f32(const float* g, const float* x, const float* mean, const float* invstd, int B, int HW, int C,
                                    float* dgamma, float* dbeta, pm_stream_t s);
int pm_bn_persample_param_grads_bf16(const void* g, const void* x, const float* mean, const float* invstd, int B, int HW, int C,
                                     float* dgamma, float* dbeta, pm_stream_t s);
/* norm2[b] += sum_j g[b][j]^2 for a [B][n] per-sample gradient block (norm2: doubles, cleared by the caller once per step) */
int pm_dp_sqnorm_f32(const float* g, int B, size_t n, double* norm2, pm_stream_t s);
/* the Linear layer's share of the per-sample norm without materialising dW_b = dlogits_b (x) feat_b:
 * norm2[b] += |dl_scale * dlogits_b|^2 (|feat_b|^2 + 1) ; dlogits rows have stride ld (pm_linear_ce_f32 leaves the
 * unnormalised per-sample dlogits in its scratch with ld = ncls + 1) */
int pm_dp_fc_sqnorm_f32(const float* dlogits, int ld, float dl_scale, const float* feat, int B, int F, int ncls, double* norm2,
                        pm_stream_t s);
/* factors[b] = min(1, max_grad_norm / (scale * sqrt(norm2[b]) + 1e-6)) ; norms_out (may be NULL) = scale * sqrt(norm2) */
int pm_dp_clip_factors(const double* norm2, int B, double scale, double max_grad_norm, float* factors, float* norms_out, pm_stream_t s);
/* out[j] (= or +=) sum_b factors[b] * g[b][j], b in ascending order (deterministic) */
int pm_dp_weighted_sum_f32(const float* g, const float* factors, int B, size_t n, float* out, int accumulate, pm_stream_t s);
int pm_dp_fc_weighted_f32(const float* dlogits, int ld, float dl_scale, const float* feat, const float* factors, int B, int F, int ncls,
                          float* dW, float* db, pm_stream_t s);
/* g[j] += stddev * N(0,1), Philox4x32-10(seed, offset + *counter_dev) + Box-Muller; counter_dev (device uint64, may be NULL) is
 * incremented afterwards, so a captured CUDA graph draws fresh noise at every replay */
int pm_dp_add_noise_f32(float* g, size_t n, float stddev, uint64_t seed, uint64_t offset, uint64_t* counter_dev, pm_stream_t s);
/* g = (g + a * x) * post  (explicit noise tensor, tests) */
int pm_dp_axpy_scale_f32(float* g, const float* x, float a, float post, size_t n, pm_stream_t s);

/* ----- training-side image front end
 * torchlib/dataloader.py:138-217 create_albu_transform, per image on the CPU in the reference: torchvision RandomAffine (PIL AFFINE,
 * NEAREST, fill 0) -> albumentations Resize(inference_resolution) (cv::resize 8U INTER_LINEAR) -> RandomCrop(train_resolution) ->
 * VerticalFlip / GaussNoise -> ToFloat(255) -> Normalize(mean, std).  One launch turns the raw uint8 images of a batch into the fp32
 * NCHW batch; PIL's 16.16 and OpenCV's 11-bit fixed-point arithmetic are reproduced bit for bit (tests/test_oracle_augment.py pins the
 * restatement to the libraries, tests/test_augment_gpu.py the kernel to the restatement).  The random parameters are drawn on the host.
 *   src      packed uint8 images, HWC, sample b at src + samples[b].src_off
 *   tables   int32; sample b's eight resize tables of R entries each at tables + samples[b].tab_off:
 *            sx0 sx1 ax0 ax1 (columns) sy0 sy1 by0 by1 (rows) -- cv::resize's offsets and cvRound(fraction * 2048) coefficients
 *   mean, rstd   HOST pointers, Cout floats each: Normalize's mean and reciprocal(std)
 *   out      fp32 [B, Cout, T, T]; out_u8 (may be NULL): the uint8 image before ToFloat, same layout (tests) */
typedef struct {
  int64_t src_off;      /* byte offset of the image in src */
  int32_t Hs, Ws, C;    /* source height, width, channels (1 or 3) */
  int32_t tab_off;      /* int32 offset of the resize tables */
  int32_t fix[6];       /* PIL affine_fixed coefficients: FIX(a) FIX(b) FIX(c + a/2 + b/2) FIX(d) FIX(e) FIX(f + d/2 + e/2) */
  int32_t cy, cx;       /* RandomCrop offset inside the resized R x R image */
  int32_t flip;         /* VerticalFlip */
  int32_t area2;        /* source is exactly 2R x 2R: cv::resize uses the 2x2 box mean */
  float noise_sigma;    /* GaussNoise on the uint8 image: 0 = off */
  uint32_t reserved;
  uint64_t noise_seed;
} pm_aug_sample_t;
int pm_augment_batch_u8_f32(const uint8_t* src, const pm_aug_sample_t* samples, const int32_t* tables, int B, int R, int T, int Cout,
                            const float* mean, const float* rstd, float* out, uint8_t* out_u8, pm_stream_t s);
/* clahe = yes (dataloader.py:150-156: a.CLAHE(clip_limit=(1, 1), always_apply) after the crop, before the flip / noise group):
 * cv::CLAHE_Impl (clahe.cpp) restated bit for bit for T divisible by `tiles`.  Three launches: pm_augment_batch_u8_f32 with
 * Cout = 1, out = NULL and flip / noise cleared writes the uint8 crop [B,T,T]; pm_clahe_luts_u8 builds the per-tile tables
 * luts [B, tiles*tiles, 256]; pm_augment_clahe_finish_f32 blends them per pixel and runs the tail of the pipeline.
 * 3-channel models (albumentations: RGB -> LAB, CLAHE on L, LAB -> RGB) are served for GREY sources, the X-ray case: pre[256] =
 * L of the grey pixel (v, v, v), post[256][3] = RGB of (L', 128, 128), both OpenCV's own 8-bit conversion tabulated
 * (primia_b200/train/_lab_tables.py); pre / post NULL = single-channel model. */
int pm_clahe_luts_u8(const uint8_t* img, const uint8_t* pre, int B, int T, int tiles, float clip_limit, uint8_t* luts, pm_stream_t s);
int pm_augment_clahe_finish_f32(const uint8_t* img, const uint8_t* pre, const uint8_t* luts, const uint8_t* post,
                                const pm_aug_sample_t* samples, int B, int T, int tiles, int Cout, const float* mean, const float* rstd,
                                float* out, uint8_t* out_u8, pm_stream_t s);

#ifdef __cplusplus
}
#endif
#endif /* PRIMIA_B200_H */
