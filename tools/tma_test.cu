// Stand-alone TMA probe: nvcc -gencode arch=compute_100a,code=sm_100a -o tma_test tma_test.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("ERR %s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(const __grid_constant__ CUtensorMap tm, int bw, int bh, int x, int y, int variant, unsigned* out) {
  extern __shared__ __align__(128) uint8_t sm[];
  __shared__ __align__(8) uint64_t bar;
  const uint32_t b = smem_u32(&bar), dst = smem_u32(sm);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(1));
    if (variant & 1) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    else asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bw * bh) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(&tm)), "r"(b), "r"(x), "r"(y) : "memory");
  }
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(b), "r"(0) : "memory");
  }
  unsigned s = 0;
  for (int i = threadIdx.x; i < bw * bh; i += blockDim.x) s += sm[i];
  atomicAdd(out, s);
}
int main(int argc, char** argv) {
  int W = 1024, H = 1024, bw = atoi(argv[1]), bh = atoi(argv[2]), x = atoi(argv[3]), y = atoi(argv[4]), variant = atoi(argv[5]);
  size_t pitch = 1024;
  uint8_t* d; CK(cudaMalloc(&d, pitch * H)); CK(cudaMemset(d, 1, pitch * H));
  unsigned* out; CK(cudaMalloc(&out, 4)); CK(cudaMemset(out, 0, 4));
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  typedef CUresult (*PFN)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  CUtensorMap tm;
  cuuint64_t dims[2] = {(cuuint64_t)W, (cuuint64_t)H}; cuuint64_t strides[1] = {pitch}; cuuint32_t box[2] = {(cuuint32_t)bw, (cuuint32_t)bh}; cuuint32_t es[2] = {1, 1};
  CUresult r = ((PFN)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                         (variant & 2) ? CU_TENSOR_MAP_L2_PROMOTION_NONE : CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode box %dx%d -> %d; ", bw, bh, (int)r);
  if (r) { printf("\n"); return 1; }
  k<<<1, 32, bw * bh + 128>>>(tm, bw, bh, x, y, variant, out);
  cudaError_t e = cudaDeviceSynchronize();
  unsigned h = 0; cudaMemcpy(&h, out, 4, cudaMemcpyDeviceToHost);
  printf("xy (%d,%d) variant %d: %s sum=%u (box cells %d)\n", x, y, variant, cudaGetErrorString(e), h, bw * bh);
  return 0;
}
