// Dependent-chain latencies on sm_100a for the ops the cost-to-go kernels are made of (one warp, clock64 around 1024 ops).
#include <cstdio>
#include <cuda_runtime.h>
#define N 1024
template <int OP>
__global__ void k(double* out, long long* cyc, double a, double b, int ia) {
  __shared__ double sm[64];
  sm[threadIdx.x] = a + threadIdx.x; sm[threadIdx.x + 32] = b;
  __syncwarp();
  double x = a + threadIdx.x;
  int ix = ia + threadIdx.x;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) {
    if (OP == 0) x = __dadd_rn(x, b);
    if (OP == 1) x = __dmul_rn(x, b);
    if (OP == 2) x = x < b ? __dadd_rn(x, 1.0) : __dadd_rn(x, 2.0);          // add + compare + select
    if (OP == 3) x = __shfl_up_sync(0xFFFFFFFFu, x, 1);
    if (OP == 4) { ix = (ix * 7 + 1) & 31; x = sm[ix]; ix += (int)x; }          // LDS with dependent address
    if (OP == 5) ix += __ballot_sync(0xFFFFFFFFu, ix & 1) & 3;
    if (OP == 6) ix = __reduce_min_sync(0xFFFFFFFFu, ix) + 1;
    if (OP == 7) ix += __any_sync(0xFFFFFFFFu, ix == 12345) ? 1 : 2;
    if (OP == 8) x = fmin(x, __dadd_rn(x, b));                                 // dadd + dmin
    if (OP == 9) { float f = (float)ix; f = f * 1.0001f + 1.0f; ix = (int)f; }
    if (OP == 10) x = __dsqrt_rn(x);
  }
  long long t1 = clock64();
  out[threadIdx.x] = x + ix;
  if (threadIdx.x == 0) cyc[OP] = t1 - t0;
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 256); cudaMallocManaged(&cyc, 16 * 8);
  const char* names[] = {"dadd", "dmul", "dadd+cmp+sel", "shfl f64", "lds dep", "ballot", "redux.min", "vote.any", "dadd+dmin", "i2f+ffma+f2i", "dsqrt"};
#define RUN(OP) k<OP><<<1, 32>>>(out, cyc, 1.0, 1.0000001, 3); cudaDeviceSynchronize(); printf("%-14s %.1f cycles/op\n", names[OP], (double)cyc[OP] / N);
  RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9) RUN(10)
  return 0;
}
