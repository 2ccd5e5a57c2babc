// Micro-benchmark of the peer-memory exchange the slab decomposition is built on (common.cuh: GramPeers /
// HaloFold): how long does "store a block into the peer's buffer, fence at system scope, publish a sequence
// number with st.release.sys" take until the peer's ld.acquire.sys spin sees it?  One process, GPU 0 against
// every other GPU of the box in turn; a ping-pong of ITERS round trips between two resident single-CTA kernels,
// one-way latency = round trip / 2.  Payloads: none (flag only), 1 152 B (the halo of two sites at N = 12),
// 2 304 B (one N x N Gram block at N = 12), 9 216 B (four of them).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -cudart shared -o p2p_latency.bin p2p_latency.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CU(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      std::fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_));  \
      std::exit(1);                                                                    \
    }                                                                                  \
  } while (0)

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// role 0 sends first and times; role 1 answers.  peer_buf / peer_flag live on the OTHER GPU, my_flag on this one.
__global__ void pingpong(int role, int iters, int payload16, double2* peer_buf, unsigned long long* peer_flag,
                         const unsigned long long* my_flag, const double2* my_buf, double* out_ns, double2* sink) {
  const int tid = threadIdx.x;
  __shared__ int gave_up;
  if (tid == 0) gave_up = 0;
  __syncthreads();
  double2 acc = make_double2(0.0, 0.0);
  unsigned long long t0 = 0;
  for (int it = 1; it <= iters + 16; ++it) {
    if (it == 17 && tid == 0) t0 = globaltimer();  // 16 warm-up round trips
    if (role == 1) {
      if (tid == 0) {
        const long long c0 = clock64();
        while (ld_acquire_sys(my_flag) < static_cast<unsigned long long>(it))
          if (clock64() - c0 > 6000000000ll) {  // ~3 s: the peer kernel never came up -- give up instead of hanging
            gave_up = 1;
            break;
          }
      }
      __syncthreads();
      if (gave_up) return;
      if (tid < payload16) {  // consume what the peer sent (as the coefficient kernels do)
        const double2 v = my_buf[tid];
        acc.x += v.x;
        acc.y += v.y;
      }
    }
    if (tid < payload16) peer_buf[tid] = make_double2(it + tid, acc.x);
    __threadfence_system();
    __syncthreads();
    if (tid == 0) st_release_sys(peer_flag, static_cast<unsigned long long>(it));
    if (role == 0) {
      if (tid == 0) {
        const long long c0 = clock64();
        while (ld_acquire_sys(my_flag) < static_cast<unsigned long long>(it))
          if (clock64() - c0 > 6000000000ll) {  // ~3 s: the peer kernel never came up -- give up instead of hanging
            gave_up = 1;
            break;
          }
      }
      __syncthreads();
      if (gave_up) return;
      if (tid < payload16) {
        const double2 v = my_buf[tid];
        acc.x += v.x;
        acc.y += v.y;
      }
    }
  }
  if (role == 0 && tid == 0) *out_ns = static_cast<double>(globaltimer() - t0) / iters;
  if (tid < payload16) sink[tid] = acc;
}

int main(int argc, char** argv) {
  int ndev = 0;
  CU(cudaGetDeviceCount(&ndev));
  const int iters = argc > 1 ? std::atoi(argv[1]) : 20000;
  if (ndev < 2) {
    std::printf("{\"error\": \"needs 2 GPUs, found %d\"}\n", ndev);
    return 0;
  }
  const int payloads[] = {0, 1152, 2304, 9216};
  for (int peer = 1; peer < ndev; ++peer) {
    int can01 = 0, can10 = 0;
    CU(cudaDeviceCanAccessPeer(&can01, 0, peer));
    CU(cudaDeviceCanAccessPeer(&can10, peer, 0));
    if (!can01 || !can10) {
      std::printf("{\"pair\": [0, %d], \"error\": \"no peer access\"}\n", peer);
      continue;
    }
    double2 *buf[2], *sink[2];
    unsigned long long* flag[2];
    double* out_ns;
    cudaStream_t st[2];
    const int dev[2] = {0, peer};
    for (int r = 0; r < 2; ++r) {
      CU(cudaSetDevice(dev[r]));
      cudaError_t e = cudaDeviceEnablePeerAccess(dev[1 - r], 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CU(e);
      (void)cudaGetLastError();
      CU(cudaMalloc(&buf[r], 16384));
      CU(cudaMalloc(&sink[r], 16384));
      CU(cudaMalloc(&flag[r], 128));
      CU(cudaStreamCreateWithFlags(&st[r], cudaStreamNonBlocking));
    }
    CU(cudaSetDevice(0));
    CU(cudaMalloc(&out_ns, sizeof(double)));
    for (int pl : payloads) {
      for (int r = 0; r < 2; ++r) {
        CU(cudaSetDevice(dev[r]));
        CU(cudaMemset(buf[r], 0, 16384));
        CU(cudaMemset(flag[r], 0, 128));
        CU(cudaDeviceSynchronize());
      }
      const int p16 = pl / 16;
      const int threads = p16 < 32 ? 32 : ((p16 + 31) / 32) * 32;
      // the answering side first, so that it is resident when the first ping arrives
      CU(cudaSetDevice(dev[1]));
      pingpong<<<1, threads, 0, st[1]>>>(1, iters, p16, buf[0], flag[0], flag[1], buf[1], nullptr, sink[1]);
      CU(cudaGetLastError());
      CU(cudaSetDevice(dev[0]));
      pingpong<<<1, threads, 0, st[0]>>>(0, iters, p16, buf[1], flag[1], flag[0], buf[0], out_ns, sink[0]);
      CU(cudaGetLastError());
      CU(cudaStreamSynchronize(st[0]));
      CU(cudaSetDevice(dev[1]));
      CU(cudaStreamSynchronize(st[1]));
      double ns = 0.0;
      CU(cudaSetDevice(0));
      CU(cudaMemcpy(&ns, out_ns, sizeof ns, cudaMemcpyDeviceToHost));
      std::printf("{\"pair\": [0, %d], \"payload_bytes\": %d, \"round_trips\": %d, \"round_trip_us\": %.3f, \"one_way_us\": %.3f}\n",
                  peer, pl, iters, ns * 1e-3, ns * 0.5e-3);
      std::fflush(stdout);
    }
    for (int r = 0; r < 2; ++r) {
      CU(cudaSetDevice(dev[r]));
      CU(cudaFree(buf[r]));
      CU(cudaFree(sink[r]));
      CU(cudaFree(flag[r]));
      CU(cudaStreamDestroy(st[r]));
    }
    CU(cudaSetDevice(0));
    CU(cudaFree(out_ns));
  }
  return 0;
}
