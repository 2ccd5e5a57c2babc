// Micro-benchmark: FP64 vector (DFMA) vs FP64 tensor (mma.sync.m8n8k4.f64) issue rates on B200,
// alone and side by side in different warps of the same SM sub-partition.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_pipes fp64_pipes.cu && ./fp64_pipes
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double (&d)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d[0]), "+d"(d[1])
               : "d"(a), "d"(b));
}

// mode bit 0: warps with (warp % 2 == 0 or all if mode==1) run DFMA; bit 1: DMMA
__global__ void k(int mode, int iters, double* out) {
  const int warp = threadIdx.x >> 5;
  double x = threadIdx.x * 1e-3, y = 1.0000001, acc[8] = {0, 1, 2, 3, 4, 5, 6, 7};
  double d[8][2] = {};
  const bool do_fma = (mode == 1) || (mode == 3 && (warp & 4) == 0);
  const bool do_mma = (mode == 2) || (mode == 3 && (warp & 4) != 0);
  if (do_fma) {
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = fma(acc[j], y, x);
    }
  }
  if (do_mma) {
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int j = 0; j < 8; ++j) dmma(d[j], x, y);
    }
  }
  double s = 0;
  for (int j = 0; j < 8; ++j) s += acc[j] + d[j][0] + d[j][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  double* out;
  cudaMalloc(&out, 148 * 1024 * sizeof(double));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int iters = 20000;
  for (int threads : {128, 256, 512}) {
    for (int mode = 1; mode <= 3; ++mode) {
      k<<<148, threads>>>(mode, 100, out);
      cudaDeviceSynchronize();
      cudaEventRecord(e0);
      k<<<148, threads>>>(mode, iters, out);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      const int warps = threads / 32;
      int wf = mode == 1 ? warps : mode == 2 ? 0 : warps / 2 + ((warps % 8) > 4 ? 0 : 0);
      int wm = mode == 2 ? warps : mode == 1 ? 0 : warps - wf;
      if (mode == 3) { wf = 0; wm = 0; for (int w = 0; w < warps; ++w) ((w & 4) == 0 ? wf : wm)++; }
      double fma_flops = 148.0 * wf * 32 * 8.0 * iters * 2;
      double mma_flops = 148.0 * wm * 8.0 * iters * 2 * 256;
      printf("threads %4d mode %d: %.3f ms  vector %.2f TF  tensor %.2f TF  (%s)\n", threads, mode, ms,
             fma_flops / ms / 1e9, mma_flops / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
    }
  }
  return 0;
}
