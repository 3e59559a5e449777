// Probe: how does tcgen05.mma (kind::f16, bf16 operands, fp32 accumulate in TMEM) round when it accumulates over a long K?
// A [128][K], B [64][K] hold bf16 values (every product is exact in fp32), D = A B^T accumulated over K/16 MMA instructions
// into one TMEM accumulator.  The result is compared with the float64 sum: signed mean relative error (a bias means
// truncation rather than round-to-nearest inside the adder) and rms, for K = 576 ... 27648 and for signed / all-positive data.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_acc umma_acc.cu -lcuda ; ./umma_acc
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
typedef __nv_bfloat16 bf16;
#define ROWS 256
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tWAIT_LOOP:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra WAIT_DONE;\n\tbra WAIT_LOOP;\n\tWAIT_DONE:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t base_off) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(base_off & 7) << 49;
  d |= (uint64_t)2 << 61;
  return d;
}
constexpr uint32_t make_idesc(int M, int N) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }


__global__ void __launch_bounds__(128, 1) probe(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int nchunks, float* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t a_s = smem_u32(smem), b_s = a_s + 128 * 128;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 128 * 128 + 64 * 128);
  const uint32_t full = smem_u32(bars), done = full + 8;
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) { mbar_init(full, 1); mbar_init(done, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(64) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot;
  if (tid == 0) {
    constexpr uint32_t idesc = make_idesc(128, 64);
    for (int c = 0; c < nchunks; ++c) {
      mbar_expect_tx(full, 128 * 128 + 64 * 128);
      tma_load_2d(a_s, &tmA, full, c * 64, 0);
      tma_load_2d(b_s, &tmB, full, c * 64, 0);
      mbar_wait(full, c & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint64_t adesc = make_desc(a_s, 16, 1024, 0), bdesc = make_desc(b_s, 16, 1024, 0);
      for (int k = 0; k < 4; ++k) {
        uint32_t acc = (c | k) != 0;
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(adesc + (uint64_t)(k * 2)), "l"(bdesc + (uint64_t)(k * 2)), "r"(idesc), "r"(acc) : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(done) : "memory");
      mbar_wait(done, c & 1);   // single-buffered: the tile is overwritten by the next chunk
    }
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int cc = 0; cc < 2; ++cc) {
    uint32_t v[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(tmem + ((uint32_t)(warp * 32) << 16) + cc * 32));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int e = 0; e < 32; ++e) out[(size_t)tid * 64 + cc * 32 + e] = __uint_as_float(v[e]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64) : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static bool map2d(EncodeTiledFn enc, CUtensorMap* tm, void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
#include <cmath>
static float gauss(uint64_t& s) {
  float acc = 0.f;
  for (int i = 0; i < 12; ++i) { s = s * 6364136223846793005ull + 1442695040888963407ull; acc += (float)((s >> 40) & 0xFFFFFF) / 16777216.f; }
  return acc - 6.f;
}
int main() {
  cudaDriverEntryPointQueryResult q;
  void* f = nullptr;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)f;
  const int Ks[] = {576, 4608, 27648};
  for (int positive = 0; positive < 2; ++positive)
    for (int K : Ks) {
      bf16 *dA, *dB; float* dO;
      cudaMalloc(&dA, (size_t)128 * K * 2); cudaMalloc(&dB, (size_t)64 * K * 2); cudaMalloc(&dO, 128 * 64 * 4);
      std::vector<bf16> hA((size_t)128 * K), hB((size_t)64 * K);
      uint64_t seed = 42 + K;
      for (auto& v : hA) { float g = gauss(seed); v = __float2bfloat16(positive ? fabsf(g) : g); }
      for (auto& v : hB) { float g = gauss(seed); v = __float2bfloat16(positive ? fabsf(g) : g); }
      cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
      cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
      CUtensorMap tmA, tmB;
      if (!map2d(enc, &tmA, dA, 128, K, 128) || !map2d(enc, &tmB, dB, 64, K, 64)) { printf("tensor map failed\n"); return 1; }
      const int smem = 128 * 128 + 64 * 128 + 1024 + 64;
      cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      probe<<<1, 128, smem>>>(tmA, tmB, K / 64, dO);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 2; }
      std::vector<float> o(128 * 64);
      cudaMemcpy(o.data(), dO, o.size() * 4, cudaMemcpyDeviceToHost);
      double num = 0, den = 0, bias = 0, f32num = 0; int cnt = 0;
      for (int m = 0; m < 128; ++m) for (int n = 0; n < 64; ++n) {
        double ref = 0; float f32 = 0.f;
        for (int k = 0; k < K; ++k) { const float p = __bfloat162float(hA[(size_t)m * K + k]) * __bfloat162float(hB[(size_t)n * K + k]); ref += (double)p; f32 += p; }
        const double d = (double)o[m * 64 + n] - ref;
        num += d * d; den += ref * ref; f32num += ((double)f32 - ref) * ((double)f32 - ref);
        if (fabs(ref) > 1e-3) { bias += d / fabs(ref) * (ref > 0 ? 1 : -1); ++cnt; }
      }
      printf("%s data, K = %5d (%4d MMAs): norm-wise rel err tcgen05 %.3e  (sequential fp32 FMA-free CPU sum %.3e) ; mean signed rel err toward +inf of |ref| %.3e\n",
             positive ? "positive" : "signed  ", K, K / 16, sqrt(num / den), sqrt(f32num / den), bias / cnt);
      cudaFree(dA); cudaFree(dB); cudaFree(dO);
    }
  return 0;
}
