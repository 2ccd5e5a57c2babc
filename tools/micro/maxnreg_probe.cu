// Probe: does a __maxnreg__(200) kernel launch with 320 threads and 165 KB dynamic shared memory?
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __maxnreg__(200) probe(double* out, int n) {
  extern __shared__ double sm[];
  double a[90];
#pragma unroll
  for (int i = 0; i < 90; ++i) a[i] = out[(threadIdx.x + i * 7) % n];
  for (int it = 0; it < n; ++it) {
#pragma unroll
    for (int i = 0; i < 90; ++i) a[i] = fma(a[i], a[(i + 1) % 90], 1.0);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 90; ++i) s += a[i];
  sm[threadIdx.x] = s;
  out[threadIdx.x] = sm[threadIdx.x];
}
int main() {
  double* out;
  cudaMalloc(&out, 4096 * 8);
  cudaMemset(out, 0, 4096 * 8);
  cudaFuncAttributes fa;
  cudaFuncGetAttributes(&fa, probe);
  printf("numRegs %d maxThreadsPerBlock %d\n", fa.numRegs, fa.maxThreadsPerBlock);
  for (int smem : {0, 100 * 1024, 165 * 1024}) {
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int threads : {256, 288, 320}) {
      probe<<<1, threads, smem>>>(out, 3);
      cudaError_t e = cudaGetLastError();
      cudaDeviceSynchronize();
      printf("smem %d threads %d: %s\n", smem, threads, cudaGetErrorString(e));
    }
  }
  int regs = 0;
  cudaDeviceGetAttribute(&regs, cudaDevAttrMaxRegistersPerBlock, 0);
  printf("MaxRegistersPerBlock %d\n", regs);
  cudaDeviceGetAttribute(&regs, cudaDevAttrMaxRegistersPerMultiprocessor, 0);
  printf("MaxRegistersPerMultiprocessor %d\n", regs);
  return 0;
}
