// Single-CTA N x N complex algebra in shared memory and the per-iteration
// "coefficient" kernels of BCG / BCGrQ / SBCGrQ.  These replace the Eigen
// fixed-size calls of inc/block_solvers.hpp (fullPivLu().solve, llt(), N x N
// products, rowwise().norm()) so that the loop never leaves the device.
//
// All matrices are column-major (element (i,j) at i + N*j), N is a run-time
// argument here (the work is latency-, not throughput-bound).
#pragma once
#include "common.cuh"

namespace bcg {

constexpr int kSmallThreads = 256;
constexpr double kEps = 2.220446049250313e-16;

// Device-memory block of coefficient matrices (offsets in units of N*N complex).
enum MatSlot {
  M_ALPHA = 0,
  M_NEGALPHA,
  M_ALPHA_INV0,  // alpha^-1, double-buffered by iteration parity (old = other slot)
  M_ALPHA_INV1,
  M_RHO0,        // rho, double-buffered likewise
  M_RHO1,
  M_RHO_CUR,     // newest rho with its diagonal inverted (operand of the in-kernel back substitution)
  M_DELTA,
  M_R2,          // BCG: r2 ; BCGrQ init: Gram of B
  M_R2_OLD,
  M_SCRATCH,     // host-supplied operand for the stand-alone primitives
  M_FIXED_COUNT
};
// then: A[s], B[s] (field-update operands), alpha_s[s], beta_s[s] for s < max_shifts
struct MatLayout {
  int N, S;
  __host__ __device__ size_t nn() const { return static_cast<size_t>(N) * N; }
  __host__ __device__ size_t fixed(int slot) const { return slot * nn(); }
  __host__ __device__ size_t A(int s) const { return (M_FIXED_COUNT + s) * nn(); }
  __host__ __device__ size_t B(int s) const { return (M_FIXED_COUNT + S + s) * nn(); }
  __host__ __device__ size_t alpha_s(int s) const { return (M_FIXED_COUNT + 2 * S + s) * nn(); }
  __host__ __device__ size_t beta_s(int s) const { return (M_FIXED_COUNT + 3 * S + s) * nn(); }
  __host__ __device__ size_t total() const { return (M_FIXED_COUNT + 4 * S) * nn(); }
};

// ---- CTA-wide primitives; every function ends with __syncthreads() -------------------
__device__ __forceinline__ void sm_copy(cd* dst, const cd* src, int n) {
  for (int e = threadIdx.x; e < n; e += blockDim.x) dst[e] = src[e];
  __syncthreads();
}
__device__ __forceinline__ void sm_identity(cd* dst, int N) {
  for (int e = threadIdx.x; e < N * N; e += blockDim.x) dst[e] = cmake((e % N) == (e / N) ? 1.0 : 0.0, 0.0);
  __syncthreads();
}
// C = A*B ; C must not alias A or B
__device__ __forceinline__ void sm_mm(cd* C, const cd* A, const cd* B, int N) {
  for (int e = threadIdx.x; e < N * N; e += blockDim.x) {
    const int i = e % N, j = e / N;
    cd s = czero();
    for (int k = 0; k < N; ++k) cmac(s, A[i + N * k], B[k + N * j]);
    C[e] = s;
  }
  __syncthreads();
}
// C = A * B^dag
__device__ __forceinline__ void sm_mm_adj(cd* C, const cd* A, const cd* B, int N) {
  for (int e = threadIdx.x; e < N * N; e += blockDim.x) {
    const int i = e % N, j = e / N;
    cd s = czero();
    for (int k = 0; k < N; ++k) cmac(s, A[i + N * k], cconj(B[j + N * k]));
    C[e] = s;
  }
  __syncthreads();
}
__device__ __forceinline__ void sm_adjoint(cd* C, const cd* A, int N) {
  for (int e = threadIdx.x; e < N * N; e += blockDim.x) {
    const int i = e % N, j = e / N;
    C[e] = cconj(A[j + N * i]);
  }
  __syncthreads();
}
// out[i] = || row i of A ||_2   (delta.rowwise().norm())
__device__ __forceinline__ void sm_rownorms(double* out, const cd* A, int N) {
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    double s = 0.0;
    for (int j = 0; j < N; ++j) s += cabs2(A[i + N * j]);
    out[i] = sqrt(s);
  }
  __syncthreads();
}

// Deterministic reduction of the per-CTA partial Grams (fixed order over p),
// lower triangle + diagonal only, upper = conjugate mirror (fields.hpp:103-122).
__device__ __forceinline__ void sm_reduce_gram(cd* G, const cd* __restrict__ gpart, int nparts, int N) {
  for (int e = threadIdx.x; e < N * N; e += blockDim.x) {
    const int i = e % N, j = e / N;
    if (i >= j) {
      double re = 0.0, im = 0.0;
      int p = 0;
      for (; p + 4 <= nparts; p += 4) {  // 4 loads in flight, summed in index order
        const cd a = gpart[static_cast<size_t>(p) * N * N + e];
        const cd b = gpart[static_cast<size_t>(p + 1) * N * N + e];
        const cd c = gpart[static_cast<size_t>(p + 2) * N * N + e];
        const cd d = gpart[static_cast<size_t>(p + 3) * N * N + e];
        re += a.x; im += a.y;
        re += b.x; im += b.y;
        re += c.x; im += c.y;
        re += d.x; im += d.y;
      }
      for (; p < nparts; ++p) {
        const cd a = gpart[static_cast<size_t>(p) * N * N + e];
        re += a.x; im += a.y;
      }
      G[e] = cmake(re, im);
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < N * N; e += blockDim.x) {
    const int i = e % N, j = e / N;
    if (i < j) G[e] = cconj(G[j + N * i]);
  }
  __syncthreads();
}

// Cholesky (lower, unblocked, left-looking as Eigen LLT.h:301-328) of the
// Hermitian G, result returned as R = L^dag (upper, exactly zero below the
// diagonal), as fields.hpp:142.  Returns -1 or the index of the first
// non-positive pivot (uniform across the CTA).  Lw: N*N scratch.
__device__ __forceinline__ int sm_chol_upper(cd* R, const cd* G, cd* Lw, int N, int* s_info) {
  sm_copy(Lw, G, N * N);
  if (threadIdx.x == 0) *s_info = -1;
  __syncthreads();
  for (int k = 0; k < N; ++k) {
    // column k below the diagonal: s_i = L(i,k) - sum_j<k L(i,j) conj(L(k,j)); pivot row i == k
    for (int i = k + threadIdx.x; i < N; i += blockDim.x) {
      if (i == k) {
        double x = Lw[k + N * k].x;
        for (int j = 0; j < k; ++j) x -= cabs2(Lw[k + N * j]);
        if (!(x > 0.0)) *s_info = k;
        Lw[k + N * k] = cmake(sqrt(x), 0.0);
      } else {
        cd s = Lw[i + N * k];
        for (int j = 0; j < k; ++j) cmsub(s, Lw[i + N * j], cconj(Lw[k + N * j]));
        Lw[i + N * k] = s;
      }
    }
    __syncthreads();
    if (*s_info >= 0) break;
    const double x = Lw[k + N * k].x;
    for (int i = k + 1 + threadIdx.x; i < N; i += blockDim.x) Lw[i + N * k] = cscale(Lw[i + N * k], 1.0 / x);
    __syncthreads();
  }
  const int info = *s_info;
  for (int e = threadIdx.x; e < N * N; e += blockDim.x) {
    const int i = e % N, j = e / N;
    R[e] = (j >= i) ? cconj(Lw[j + N * i]) : czero();
  }
  __syncthreads();
  return info;
}

// Full-pivoting LU solve A X = B following Eigen FullPivLU (LU/FullPivLU.h:
// 487-590 compute, 745-790 solve, 317-341 rank threshold eps*N*|maxpivot|):
// pivot = entry of largest modulus in the trailing block, first one in
// column-major order on ties.  lu, X: N*N each; X enters as B.  ws: workspace
// of N*N complex + 4N ints + 64 doubles (carved below).
struct LuWork {
  cd* c;        // N*N
  int* rt;      // N
  int* ct;      // N
  int* p;       // N
  int* q;       // N
  double* red_v;  // 32
  int* red_i;     // 32
  double* misc;   // [0] maxpivot
  int* imisc;     // [0] nonzero [1] rank [2] br [3] bc
};

__device__ __forceinline__ void sm_lu_solve(cd* X, cd* lu, const LuWork& w, int N) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
  if (tid == 0) {
    w.misc[0] = 0.0;
    w.imisc[0] = N;
  }
  __syncthreads();
  for (int k = 0; k < N; ++k) {
    // ---- pivot search over the trailing (N-k)^2 block ----
    const int m = N - k;
    double bv = -1.0;
    int bi = 0x7fffffff;
    for (int e = tid; e < m * m; e += blockDim.x) {
      const int i = k + e % m, j = k + e / m;
      const double a = hypot(lu[i + N * j].x, lu[i + N * j].y);
      const int lin = i + N * j;  // column-major scan order == Eigen's visitor order
      if (a > bv || (a == bv && lin < bi)) {
        bv = a;
        bi = lin;
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, bv, off);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
      if (ov > bv || (ov == bv && oi < bi)) {
        bv = ov;
        bi = oi;
      }
    }
    if (lane == 0) {
      w.red_v[warp] = bv;
      w.red_i[warp] = bi;
    }
    __syncthreads();
    if (tid == 0) {
      double v = w.red_v[0];
      int ix = w.red_i[0];
      for (int ww = 1; ww < nwarp; ++ww)
        if (w.red_v[ww] > v || (w.red_v[ww] == v && w.red_i[ww] < ix)) {
          v = w.red_v[ww];
          ix = w.red_i[ww];
        }
      if (v == 0.0) {
        w.imisc[0] = k;  // nonzero pivots
        w.imisc[2] = -1;
      } else {
        if (v > w.misc[0]) w.misc[0] = v;
        w.imisc[2] = ix % N;
        w.imisc[3] = ix / N;
        w.rt[k] = ix % N;
        w.ct[k] = ix / N;
      }
    }
    __syncthreads();
    if (w.imisc[2] < 0) {
      for (int i = k + tid; i < N; i += blockDim.x) w.rt[i] = w.ct[i] = i;
      __syncthreads();
      break;
    }
    const int br = w.imisc[2], bc = w.imisc[3];
    if (br != k)
      for (int j = tid; j < N; j += blockDim.x) {
        const cd t = lu[k + N * j];
        lu[k + N * j] = lu[br + N * j];
        lu[br + N * j] = t;
      }
    __syncthreads();
    if (bc != k)
      for (int i = tid; i < N; i += blockDim.x) {
        const cd t = lu[i + N * k];
        lu[i + N * k] = lu[i + N * bc];
        lu[i + N * bc] = t;
      }
    __syncthreads();
    const cd piv = lu[k + N * k];
    for (int i = k + 1 + tid; i < N; i += blockDim.x) lu[i + N * k] = cdiv(lu[i + N * k], piv);
    __syncthreads();
    const int m1 = N - k - 1;
    for (int e = tid; e < m1 * m1; e += blockDim.x) {
      const int i = k + 1 + e % m1, j = k + 1 + e / m1;
      cmsub(lu[i + N * j], lu[i + N * k], lu[k + N * j]);
    }
    __syncthreads();
  }
  // ---- permutations and rank ----
  if (tid == 0) {
    for (int i = 0; i < N; ++i) w.p[i] = w.q[i] = i;
    for (int k = N - 1; k >= 0; --k) {
      const int t = w.p[k];
      w.p[k] = w.p[w.rt[k]];
      w.p[w.rt[k]] = t;
    }
    for (int k = 0; k < N; ++k) {
      const int t = w.q[k];
      w.q[k] = w.q[w.ct[k]];
      w.q[w.ct[k]] = t;
    }
    const double thr = w.misc[0] * (kEps * N);
    int rank = 0;
    for (int i = 0; i < w.imisc[0]; ++i) rank += (hypot(lu[i + N * i].x, lu[i + N * i].y) > thr);
    w.imisc[1] = rank;
  }
  __syncthreads();
  const int rank = w.imisc[1];
  if (rank == 0) {
    for (int e = tid; e < N * N; e += blockDim.x) X[e] = czero();
    __syncthreads();
    return;
  }
  // c = P * B : row p[i] of c = row i of B
  for (int e = tid; e < N * N; e += blockDim.x) {
    const int i = e % N, col = e / N;
    w.c[w.p[i] + N * col] = X[i + N * col];
  }
  __syncthreads();
  // unit-lower forward substitution, column oriented
  for (int j = 0; j < N - 1; ++j) {
    for (int e = tid; e < (N - 1 - j) * N; e += blockDim.x) {
      const int i = j + 1 + e % (N - 1 - j), col = e / (N - 1 - j);
      cmsub(w.c[i + N * col], lu[i + N * j], w.c[j + N * col]);
    }
    __syncthreads();
  }
  // upper backward substitution on the leading rank x rank block
  for (int i = rank - 1; i >= 0; --i) {
    const cd d = lu[i + N * i];
    for (int col = tid; col < N; col += blockDim.x) w.c[i + N * col] = cdiv(w.c[i + N * col], d);
    __syncthreads();
    for (int e = tid; e < i * N; e += blockDim.x) {
      const int r = e % i, col = e / i;
      cmsub(w.c[r + N * col], lu[r + N * i], w.c[i + N * col]);
    }
    __syncthreads();
  }
  for (int e = tid; e < N * N; e += blockDim.x) {
    const int i = e % N, col = e / N;
    X[w.q[i] + N * col] = (i < rank) ? w.c[i + N * col] : czero();
  }
  __syncthreads();
}

// shared-memory carve-up for the coefficient kernels: NMAT matrices + LU work
struct SmallSmem {
  cd* mat[12];
  LuWork lw;
  double* vec;  // 2*N doubles
  int* info;
  __device__ __forceinline__ void carve(unsigned char* raw, int N) {
    cd* base = reinterpret_cast<cd*>(raw);
    const int nn = N * N;
    for (int i = 0; i < 12; ++i) mat[i] = base + i * nn;
    lw.c = base + 12 * nn;
    double* d = reinterpret_cast<double*>(base + 13 * nn);
    lw.red_v = d;
    lw.misc = d + 32;
    vec = d + 40;
    int* ip = reinterpret_cast<int*>(d + 40 + 2 * N);
    lw.red_i = ip;
    lw.imisc = ip + 32;
    lw.rt = ip + 40;
    lw.ct = lw.rt + N;
    lw.p = lw.ct + N;
    lw.q = lw.p + N;
    info = lw.q + N;
  }
  static size_t bytes(int N) {
    return sizeof(cd) * 13 * N * N + sizeof(double) * (40 + 2 * N) + sizeof(int) * (40 + 4 * N + 8);
  }
};

// ---- stand-alone helpers used by the primitives (bcg_gram, bcg_thinqr) -------------------
// out = reduced Gram
__global__ void __launch_bounds__(kSmallThreads)
gram_reduce_kernel(cd* __restrict__ out, const cd* __restrict__ gpart, int nparts, int N) {
  extern __shared__ __align__(16) unsigned char raw[];
  cd* G = reinterpret_cast<cd*>(raw);
  sm_reduce_gram(G, gpart, nparts, N);
  for (int e = threadIdx.x; e < N * N; e += blockDim.x) out[e] = G[e];
}

// R = chol(G)^dag  (G given reduced in device memory); status -> ctrl->status
__global__ void __launch_bounds__(kSmallThreads)
chol_kernel(cd* __restrict__ R, const cd* __restrict__ G, int N, Ctrl* __restrict__ ctrl) {
  extern __shared__ __align__(16) unsigned char raw[];
  SmallSmem s;
  s.carve(raw, N);
  sm_copy(s.mat[0], G, N * N);
  const int info = sm_chol_upper(s.mat[1], s.mat[0], s.mat[2], N, s.info);
  for (int e = threadIdx.x; e < N * N; e += blockDim.x) R[e] = s.mat[1][e];
  if (threadIdx.x == 0 && info >= 0 && ctrl) ctrl->status = 3;
}

// ---- SBCGrQ / BCGrQ ----------------------------------------------------------------------
// Setup after the Gram of B: delta = chol(B^dag B)^dag; rho = delta; alpha^-1 = I;
// alpha_s = beta_s = I; b_norm = rowwise norms of delta  (block_solvers.hpp:103-131).
__global__ void __launch_bounds__(kSmallThreads)
rq_init_kernel(cd* __restrict__ mats, MatLayout L, double* __restrict__ b_norm, const cd* __restrict__ gpart,
               int nparts, Ctrl* __restrict__ ctrl) {
  extern __shared__ __align__(16) unsigned char raw[];
  const int N = L.N, nn = N * N;
  SmallSmem s;
  s.carve(raw, N);
  sm_reduce_gram(s.mat[0], gpart, nparts, N);
  const int info = sm_chol_upper(s.mat[1], s.mat[0], s.mat[2], N, s.info);
  sm_identity(s.mat[3], N);
  sm_rownorms(s.vec, s.mat[1], N);
  for (int e = threadIdx.x; e < nn; e += blockDim.x) {
    const cd d = s.mat[1][e], id = s.mat[3][e];
    mats[L.fixed(M_DELTA) + e] = d;
    mats[L.fixed(M_RHO0) + e] = d;
    mats[L.fixed(M_RHO1) + e] = d;
    mats[L.fixed(M_RHO_CUR) + e] = ((e % N) == (e / N)) ? cdiv(cmake(1.0, 0.0), d) : d;
    mats[L.fixed(M_ALPHA_INV0) + e] = id;
    mats[L.fixed(M_ALPHA_INV1) + e] = id;
    for (int sh = 0; sh < L.S; ++sh) {
      mats[L.alpha_s(sh) + e] = id;
      mats[L.beta_s(sh) + e] = id;
    }
  }
  for (int i = threadIdx.x; i < N; i += blockDim.x) b_norm[i] = s.vec[i];
  if (threadIdx.x == 0 && info >= 0) {
    ctrl->status = 3;
    ctrl->done = 1;
  }
}

// A-step (after the stencil): alpha^-1 = herm(P0^dag T) ; alpha = LU-inverse ;
// A_0 = alpha*delta (old delta!) ; -alpha.   (block_solvers.hpp:139-148)
// Also: iteration bookkeeping (iter++, promote stop -> done, retire converged shifts).
__global__ void __launch_bounds__(kSmallThreads)
rq_step_a_kernel(cd* __restrict__ mats, MatLayout L, const cd* __restrict__ gpart, int nparts,
                 Ctrl* __restrict__ ctrl) {
  if (ctrl->done) return;
  if (ctrl->stop) {
    __syncthreads();
    if (threadIdx.x == 0) ctrl->done = 1;
    return;
  }
  extern __shared__ __align__(16) unsigned char raw[];
  const int N = L.N, nn = N * N;
  SmallSmem s;
  s.carve(raw, N);
  const int iter = ctrl->iter + 1;
  const int n_unconv_old = ctrl->n_unconv;
  __syncthreads();
  if (threadIdx.x == 0) {
    ctrl->iter = iter;
    // shifts that passed the test in the previous iteration were still updated in
    // it and drop out from this one on (block_solvers.hpp:161,175-181): every
    // passing shift decrements the count, which always retires the highest index.
    int n = n_unconv_old;
    for (int sh = 1; sh < n_unconv_old; ++sh)
      if (ctrl->conv[sh]) {
        --n;
        ctrl->conv[sh] = 0;
      }
    ctrl->n_unconv = n;
  }
  cd* Ainv = s.mat[0];
  cd* lu = s.mat[1];
  cd* alpha = s.mat[2];
  cd* delta = s.mat[3];
  cd* ad = s.mat[4];
  sm_reduce_gram(Ainv, gpart, nparts, N);
  sm_copy(lu, Ainv, nn);
  sm_identity(alpha, N);
  sm_lu_solve(alpha, lu, s.lw, N);
  sm_copy(delta, mats + L.fixed(M_DELTA), nn);
  sm_mm(ad, alpha, delta, N);
  cd* ainv_g = mats + L.fixed((iter & 1) ? M_ALPHA_INV1 : M_ALPHA_INV0);
  for (int e = threadIdx.x; e < nn; e += blockDim.x) {
    ainv_g[e] = Ainv[e];
    mats[L.fixed(M_ALPHA) + e] = alpha[e];
    mats[L.fixed(M_NEGALPHA) + e] = cmake(-alpha[e].x, -alpha[e].y);
    mats[L.A(0) + shift_mat_index(N, e % N, e / N)] = ad[e];
  }
}

// B-step (after Q -= T alpha and its Gram): one CTA per shift.
//   every CTA : G = herm(Q^dag Q) ; rho = chol(G)^dag            (fields.hpp:142)
//   CTA 0     : delta = rho*delta ; residual ; stop test ; B_0 = rho^dag
//               (block_solvers.hpp:152-158)
//   CTA s>=1  : beta_s, alpha_s, shifted residual, A_s = alpha_s, B_s = beta_s rho^dag
//               (block_solvers.hpp:163-181)
__global__ void __launch_bounds__(kSmallThreads)
rq_step_b_kernel(cd* __restrict__ mats, MatLayout L, const double* __restrict__ b_norm,
                 const cd* __restrict__ gpart, int nparts, Ctrl* __restrict__ ctrl) {
  if (ctrl->done) return;
  const int sh = blockIdx.x;
  if (sh >= ctrl->n_unconv) return;
  extern __shared__ __align__(16) unsigned char raw[];
  const int N = L.N, nn = N * N;
  SmallSmem s;
  s.carve(raw, N);
  const int iter = ctrl->iter;
  cd* G = s.mat[0];
  cd* rho = s.mat[1];
  cd* t0 = s.mat[2];
  sm_reduce_gram(G, gpart, nparts, N);
  const int info = sm_chol_upper(rho, G, t0, N, s.info);
  cd* rho_g = mats + L.fixed((iter & 1) ? M_RHO1 : M_RHO0);
  const cd* rho_old_g = mats + L.fixed((iter & 1) ? M_RHO0 : M_RHO1);
  if (sh == 0) {
    cd* delta = s.mat[3];
    cd* dn = s.mat[4];
    sm_copy(delta, mats + L.fixed(M_DELTA), nn);
    sm_mm(dn, rho, delta, N);
    sm_rownorms(s.vec, dn, N);
    for (int e = threadIdx.x; e < nn; e += blockDim.x) {
      const int i = e % N, j = e / N;
      rho_g[e] = rho[e];
      mats[L.fixed(M_RHO_CUR) + e] = (i == j) ? cdiv(cmake(1.0, 0.0), rho[e]) : rho[e];
      mats[L.fixed(M_DELTA) + e] = dn[e];
      mats[L.B(0) + shift_mat_index(N, i, j)] = cconj(rho[j + N * i]);
    }
    if (threadIdx.x == 0) {
      double r = 0.0;
      bool nan = false;
      for (int i = 0; i < N; ++i) {
        const double v = s.vec[i] / b_norm[i];
        if (v != v) nan = true;
        r = fmax(r, v);
      }
      if (nan) r = nan ? (0.0 / 0.0) : r;
      ctrl->residual = r;
      // while (residual > eps && iter < max_iterations)  -- NaN ends the loop as in the reference
      if (!(r > ctrl->eps) || iter >= ctrl->max_it) ctrl->stop = 1;
      if (info >= 0) {
        ctrl->status = 3;
        ctrl->stop = 1;
      } else if (nan) {
        ctrl->status = 6;
      }
    }
    return;
  }
  // ---- shifted coefficients ----
  cd* alpha = s.mat[3];
  cd* rho_old = s.mat[4];
  cd* ainv_old = s.mat[5];
  cd* beta = s.mat[6];
  cd* t1 = s.mat[7];
  cd* t2 = s.mat[8];
  cd* as = s.mat[9];
  cd* ainv = s.mat[10];
  cd* lu = s.mat[11];
  for (int e = threadIdx.x; e < nn; e += blockDim.x) {
    alpha[e] = mats[L.fixed(M_ALPHA) + e];
    rho_old[e] = rho_old_g[e];
    ainv_old[e] = mats[L.fixed((iter & 1) ? M_ALPHA_INV0 : M_ALPHA_INV1) + e];
    ainv[e] = mats[L.fixed((iter & 1) ? M_ALPHA_INV1 : M_ALPHA_INV0) + e];
    beta[e] = mats[L.beta_s(sh) + e];
    as[e] = mats[L.alpha_s(sh) + e];
  }
  __syncthreads();
  // beta_s_inv = I + (sigma_s - sigma_0) alpha + alpha rho_old alpha_inv_old (I - beta_s) rho_old^dag
  sm_mm(t1, alpha, rho_old, N);
  sm_mm(t2, t1, ainv_old, N);
  for (int e = threadIdx.x; e < nn; e += blockDim.x) {
    const double id = ((e % N) == (e / N)) ? 1.0 : 0.0;
    t1[e] = cmake(id - beta[e].x, -beta[e].y);
  }
  __syncthreads();
  sm_mm(G, t2, t1, N);           // G reused as scratch from here on
  sm_mm_adj(t1, G, rho_old, N);  // ... * rho_old^dag
  const double ds = ctrl->sigma[sh] - ctrl->sigma[0];
  for (int e = threadIdx.x; e < nn; e += blockDim.x) {
    const double id = ((e % N) == (e / N)) ? 1.0 : 0.0;
    lu[e] = cmake((id + ds * alpha[e].x) + t1[e].x, (ds * alpha[e].y) + t1[e].y);
  }
  sm_identity(beta, N);
  sm_lu_solve(beta, lu, s.lw, N);  // beta_s = beta_s_inv^-1
  // alpha_s = beta_s alpha rho_old alpha_inv_old alpha_s  (left to right)
  sm_mm(t1, beta, alpha, N);
  sm_mm(t2, t1, rho_old, N);
  sm_mm(t1, t2, ainv_old, N);
  sm_mm(t2, t1, as, N);  // new alpha_s
  // residual_shift = max_i || row_i(rho alpha_inv alpha_s) || / b_norm_i
  sm_mm(t1, rho, ainv, N);
  sm_mm(G, t1, t2, N);
  sm_rownorms(s.vec, G, N);
  sm_mm_adj(t1, beta, rho, N);  // B_s = beta_s rho^dag
  for (int e = threadIdx.x; e < nn; e += blockDim.x) {
    mats[L.alpha_s(sh) + e] = t2[e];
    mats[L.beta_s(sh) + e] = beta[e];
    mats[L.A(sh) + shift_mat_index(N, e % N, e / N)] = t2[e];
    mats[L.B(sh) + shift_mat_index(N, e % N, e / N)] = t1[e];
  }
  if (threadIdx.x == 0) {
    double r = 0.0;
    for (int i = 0; i < N; ++i) r = fmax(r, s.vec[i] / b_norm[i]);
    ctrl->resid_shift[sh] = r;
    ctrl->conv[sh] = (r < ctrl->eps_shifts) ? 1 : 0;
  }
}

// ---- BCG (block_solvers.hpp:10-45) -----------------------------------------------------------
// init: r2 = R^dag R ; residual_norms_i = sqrt(r2_ii)
__global__ void __launch_bounds__(kSmallThreads)
bcg_init_kernel(cd* __restrict__ mats, MatLayout L, double* __restrict__ b_norm, const cd* __restrict__ gpart,
                int nparts) {
  extern __shared__ __align__(16) unsigned char raw[];
  const int N = L.N, nn = N * N;
  SmallSmem s;
  s.carve(raw, N);
  sm_reduce_gram(s.mat[0], gpart, nparts, N);
  for (int e = threadIdx.x; e < nn; e += blockDim.x) mats[L.fixed(M_R2) + e] = s.mat[0][e];
  for (int i = threadIdx.x; i < N; i += blockDim.x) b_norm[i] = sqrt(s.mat[0][i + N * i].x);
}
// A-step: alpha = LU(P^dag T).solve(r2) ; -alpha ; A_0 = alpha
__global__ void __launch_bounds__(kSmallThreads)
bcg_step_a_kernel(cd* __restrict__ mats, MatLayout L, const cd* __restrict__ gpart, int nparts,
                  Ctrl* __restrict__ ctrl) {
  if (ctrl->done) return;
  if (ctrl->stop) {
    __syncthreads();
    if (threadIdx.x == 0) ctrl->done = 1;
    return;
  }
  extern __shared__ __align__(16) unsigned char raw[];
  const int N = L.N, nn = N * N;
  SmallSmem s;
  s.carve(raw, N);
  const int iter = ctrl->iter + 1;
  __syncthreads();
  if (threadIdx.x == 0) ctrl->iter = iter;
  sm_reduce_gram(s.mat[0], gpart, nparts, N);
  sm_copy(s.mat[1], mats + L.fixed(M_R2), nn);  // X enters as B = r2
  sm_lu_solve(s.mat[1], s.mat[0], s.lw, N);
  for (int e = threadIdx.x; e < nn; e += blockDim.x) {
    const cd a = s.mat[1][e];
    mats[L.fixed(M_ALPHA) + e] = a;
    mats[L.fixed(M_NEGALPHA) + e] = cmake(-a.x, -a.y);
    mats[L.A(0) + shift_mat_index(N, e % N, e / N)] = a;
  }
}
// B-step: r2_old = r2 ; r2 = R^dag R ; beta = LU(r2_old).solve(r2) ; residual ; B_0 = beta
__global__ void __launch_bounds__(kSmallThreads)
bcg_step_b_kernel(cd* __restrict__ mats, MatLayout L, const double* __restrict__ b_norm,
                  const cd* __restrict__ gpart, int nparts, Ctrl* __restrict__ ctrl) {
  if (ctrl->done) return;
  extern __shared__ __align__(16) unsigned char raw[];
  const int N = L.N, nn = N * N;
  SmallSmem s;
  s.carve(raw, N);
  cd* r2 = s.mat[0];
  cd* r2old = s.mat[1];
  cd* beta = s.mat[2];
  sm_reduce_gram(r2, gpart, nparts, N);
  sm_copy(r2old, mats + L.fixed(M_R2), nn);
  sm_copy(beta, r2, nn);
  sm_lu_solve(beta, r2old, s.lw, N);
  for (int e = threadIdx.x; e < nn; e += blockDim.x) {
    mats[L.fixed(M_R2) + e] = r2[e];
    mats[L.B(0) + shift_mat_index(N, e % N, e / N)] = beta[e];
  }
  if (threadIdx.x == 0) {
    double r = 0.0;
    bool nan = false;
    for (int i = 0; i < N; ++i) {
      const double v = sqrt(r2[i + N * i].x) / b_norm[i];
      if (v != v) nan = true;
      r = fmax(r, v);
    }
    if (nan) r = 0.0 / 0.0;
    ctrl->residual = r;
    if (!(r > ctrl->eps) || ctrl->iter >= ctrl->max_it) ctrl->stop = 1;
    if (nan) ctrl->status = 6;
  }
}

}  // namespace bcg
