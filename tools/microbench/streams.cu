// dev micro-benchmark: achievable HBM bandwidth of an n-stream SoA copy (what a D3Q19 update looks
// like to the DRAM system) against the plain 2-stream copy MEASURED_PEAKS.json quotes.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
__global__ void k_copy_soa(const double* __restrict__ in, double* __restrict__ out, int64_t n, int nq, int64_t S) {
  const int64_t i = (int64_t)blockIdx.x*blockDim.x + threadIdx.x;
  if (i >= n) return;
  double f[19];
#pragma unroll
  for (int q = 0; q < 19; q++) if (q < nq) f[q] = __ldg(in + q*S + i);
#pragma unroll
  for (int q = 0; q < 19; q++) if (q < nq) out[q*S + i] = f[q];
}
__global__ void k_copy_flat(const double4* __restrict__ in, double4* __restrict__ out, int64_t n4) {
  for (int64_t i = (int64_t)blockIdx.x*blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x*blockDim.x) out[i] = in[i];
}
int main() {
  const int64_t n = 256LL*256*256, S = n + 2*65536;
  double *a, *b; cudaMalloc(&a, 19*S*8); cudaMalloc(&b, 19*S*8);
  cudaMemset(a, 0, 19*S*8); cudaMemset(b, 0, 19*S*8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int nq : {1, 2, 5, 10, 19}) {
    for (int w = 0; w < 3; w++) k_copy_soa<<<(n + 255)/256, 256>>>(a, b, n, nq, S);
    cudaEventRecord(e0);
    for (int r = 0; r < 20; r++) k_copy_soa<<<(n + 255)/256, 256>>>(r & 1 ? b : a, r & 1 ? a : b, n, nq, S);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 20;
    printf("soa copy nq=%2d: %.4f ms  %.0f GB/s\n", nq, ms, 2.0*nq*n*8/ms/1e6);
  }
  const int64_t n4 = 19*n/4;
  for (int w = 0; w < 3; w++) k_copy_flat<<<148*16, 256>>>((double4*)a, (double4*)b, n4);
  cudaEventRecord(e0);
  for (int r = 0; r < 20; r++) k_copy_flat<<<148*16, 256>>>((double4*)(r & 1 ? b : a), (double4*)(r & 1 ? a : b), n4);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 20;
  printf("flat copy (19 n doubles): %.4f ms  %.0f GB/s\n", ms, 2.0*19*n*8/ms/1e6);
  cudaEventRecord(e0);
  for (int r = 0; r < 20; r++) cudaMemcpyAsync(b, a, 19*n*8, cudaMemcpyDeviceToDevice);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  cudaEventElapsedTime(&ms, e0, e1); ms /= 20;
  printf("cudaMemcpy D2D: %.4f ms  %.0f GB/s\n", ms, 2.0*19*n*8/ms/1e6);
  return 0;
}
