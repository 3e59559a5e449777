// Probe: which fp32 / SWIZZLE_NONE 4-D tiled TMA boxes over an NCHW tensor are legal (cp.async.bulk.tensor.4d)?
//   ./tma_f32_box <box0> <box1> <box2> <c0>   -> loads one box at coords (c0, 2, 0, 0) and checks it against the host tensor
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(const __grid_constant__ CUtensorMap tm, int c0, int nbytes, float* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 65536);
  const uint32_t b = smem_u32(bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(nbytes) : "memory");
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(smem)), "l"(&tm), "r"(b), "r"(c0), "r"(2), "r"(0), "r"(0) : "memory");
    asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(b) : "memory");
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nbytes / 4; i += blockDim.x) out[i] = reinterpret_cast<float*>(smem)[i];
}
int main(int argc, char** argv) {
  const int b0 = atoi(argv[1]), b1 = atoi(argv[2]), b2 = atoi(argv[3]), c0 = atoi(argv[4]);
  const int W = 40, H = 40, B = 2;
  std::vector<float> h((size_t)B * 3 * H * W);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (float)i;
  float *d, *o;
  cudaMalloc(&d, h.size() * 4); cudaMalloc(&o, 65536);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  cudaDriverEntryPointQueryResult q; void* f = nullptr;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q);
  typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  CUtensorMap tm;
  cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, 3, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4, (cuuint64_t)3 * H * W * 4};
  cuuint32_t box[4] = {(cuuint32_t)b0, (cuuint32_t)b1, (cuuint32_t)b2, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = ((Fn)f)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("box {%d,%d,%d,1}: encode failed %d\n", b0, b1, b2, (int)r); return 0; }
  const int nbytes = b0 * b1 * b2 * 4;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 64);
  k<<<1, 128, 65536 + 64>>>(tm, c0, nbytes, o);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("box {%d,%d,%d,1} c0=%d: %s\n", b0, b1, b2, c0, cudaGetErrorString(e)); return 0; }
  std::vector<float> ho(nbytes / 4);
  cudaMemcpy(ho.data(), o, nbytes, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int c = 0; c < b2; ++c) for (int r_ = 0; r_ < b1; ++r_) for (int j = 0; j < b0; ++j) {
    const int iw = c0 + j, ih = 2 + r_;
    const float want = (iw >= 0 && iw < W && ih < H) ? h[((size_t)c * H + ih) * W + iw] : 0.f;
    if (ho[(c * b1 + r_) * b0 + j] != want) ++bad;
  }
  printf("box {%d,%d,%d,1} c0=%d: ok, mismatches %d of %d\n", b0, b1, b2, c0, bad, nbytes / 4);
  return 0;
}
