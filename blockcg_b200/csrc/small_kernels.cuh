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

constexpr int kSmallThreads = 512;
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
  // paired multishift update (shift_pair.cuh): the operands of odd iterations live in a second set of
  // slots, so that the even iteration's launch still finds them
  int pair = 0;   // schedule of the multishift update: 0 plain, 1 alternating, 2 staggered (build_shift_items),
                  // 3 staggered of depth `depth` with streamed operands (build_stag_items)
  int depth = 2;  // deferral depth k of the schedules 1..3
  int ring = 2;   // operand sets / Q fields / n_act slots in use: those of iteration i are number i % ring
                  // (= depth, or depth + 1 when the shifted systems' launch overlaps the next iterations)
  int overlap = 0;
  // operand set t of [A[S], B[S]]: set 0 at the historical place, sets 1 .. kMaxDepth-1 behind alpha_s / beta_s
  __host__ __device__ size_t Aset(int s, int t) const { return (t == 0 ? M_FIXED_COUNT + s : M_FIXED_COUNT + 4 * S + 2 * S * (t - 1) + s) * nn(); }
  __host__ __device__ size_t Bset(int s, int t) const { return (t == 0 ? M_FIXED_COUNT + S + s : M_FIXED_COUNT + 4 * S + 2 * S * (t - 1) + S + s) * nn(); }
  __host__ __device__ int set_of(int iter) const { return pair ? iter % ring : 0; }
  __host__ __device__ size_t A(int s, int iter) const { return Aset(s, set_of(iter)); }
  __host__ __device__ size_t B(int s, int iter) const { return Bset(s, set_of(iter)); }
  __host__ __device__ size_t total() const { return (M_FIXED_COUNT + 4 * S + 2 * S * (kMaxDepth - 1)) * nn(); }
};

// (row, column) of every linear matrix index, filled once per kernel: an integer division by
// the run-time N costs more than a whole complex multiply-add chain of these little kernels.
__shared__ int g_ij[1024];  // i | j << 16
__shared__ int g_il[1024];  // shift_mat_index(N, i, j): position in the column-group-interleaved operand layout
__device__ __forceinline__ void sm_init_ij(int N) {
  for (int e = threadIdx.x; e < N * N; e += blockDim.x) {
    g_ij[e] = (e % N) | ((e / N) << 16);
    g_il[e] = shift_mat_index(N, e % N, e / N);
  }
  __syncthreads();
}
#define BCG_IJ(e, i, j) const int ij_##i = g_ij[e]; const int i = ij_##i & 0xffff, j = ij_##i >> 16

// ---- CTA-wide primitives; every function ends with __syncthreads() -------------------
__device__ __noinline__ void sm_copy(cd* dst, const cd* src, int n) {
  for (int e = threadIdx.x; e < n; e += blockDim.x) dst[e] = src[e];
  __syncthreads();
}
__device__ __noinline__ void sm_identity(cd* dst, int N) {
  for (int e = threadIdx.x; e < N * N; e += blockDim.x) {
    BCG_IJ(e, i, j);
    dst[e] = cmake(i == j ? 1.0 : 0.0, 0.0);
  }
  __syncthreads();
}
// C = A*B ; C must not alias A or B.  (A fully unrolled body for the compile-time sizes 4 / 8 / 12 / 16 was measured
// SLOWER than this loop, 0.3-0.5 us per kernel: these kernels run once per iteration, and straight-line code is
// fetched once where a loop body is fetched once and reused -- profiles/r02_ab_coefficient_kernels.jsonl.)
__device__ __noinline__ void sm_mm(cd* C, const cd* A, const cd* B, int N) {
  for (int e = threadIdx.x; e < N * N; e += blockDim.x) {
    BCG_IJ(e, i, j);
    cd s0 = czero(), s1 = czero();  // two interleaved partial sums halve the dependent chain
    int k = 0;
#pragma unroll 2
    for (; k + 1 < N; k += 2) {
      cmac(s0, A[i + N * k], B[k + N * j]);
      cmac(s1, A[i + N * (k + 1)], B[k + 1 + N * j]);
    }
    if (k < N) cmac(s0, A[i + N * k], B[k + N * j]);
    C[e] = cadd(s0, s1);
  }
  __syncthreads();
}
// C = A * B^dag
__device__ __noinline__ void sm_mm_adj(cd* C, const cd* A, const cd* B, int N) {
  for (int e = threadIdx.x; e < N * N; e += blockDim.x) {
    BCG_IJ(e, i, j);
    cd s0 = czero(), s1 = czero();
    int k = 0;
#pragma unroll 2
    for (; k + 1 < N; k += 2) {
      cmac(s0, A[i + N * k], cconj(B[j + N * k]));
      cmac(s1, A[i + N * (k + 1)], cconj(B[j + N * (k + 1)]));
    }
    if (k < N) cmac(s0, A[i + N * k], cconj(B[j + N * k]));
    C[e] = cadd(s0, s1);
  }
  __syncthreads();
}
__device__ __noinline__ void sm_adjoint(cd* C, const cd* A, int N) {
  for (int e = threadIdx.x; e < N * N; e += blockDim.x) {
    BCG_IJ(e, i, j);
    C[e] = cconj(A[j + N * i]);
  }
  __syncthreads();
}
// out[i] = || row i of A ||_2   (delta.rowwise().norm())
__device__ __noinline__ void sm_rownorms(double* out, const cd* A, int N) {
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    double s = 0.0;
    for (int j = 0; j < N; ++j) s += cabs2(A[i + N * j]);
    out[i] = sqrt(s);
  }
  __syncthreads();
}

// Deterministic reduction of the per-CTA partial Grams, lower triangle + diagonal only, upper =
// conjugate mirror (fields.hpp:103-122).  The loop over partials is latency-bound (one L2 round
// trip per load), so every lower-triangle entry is summed by NSL = min(kRedSlices, blockDim/E)
// threads, thread q taking partials q, q+NSL, ... in index order with 16 loads in flight, and the
// slice sums are then combined in slice order: a fixed summation tree, identical from run to run.
// `scratch`: kRedSlices * N*(N+1)/2 complex (may not alias G).
constexpr int kRedSlices = 8;
__device__ __noinline__ void sm_reduce_gram(cd* G, const cd* __restrict__ gpart, int nparts, int N,
                                               cd* scratch) {
  const int nn = N * N, E = N * (N + 1) / 2;
  if (nparts == 1) {  // the producing kernel has already reduced: copy the lower triangle, mirror it
    for (int e = threadIdx.x; e < nn; e += blockDim.x) {
      BCG_IJ(e, i, j);
      G[e] = (i >= j) ? gpart[e] : cconj(gpart[j + N * i]);
    }
    __syncthreads();
    return;
  }
  int nsl = static_cast<int>(blockDim.x) / E;
  nsl = nsl < 1 ? 1 : (nsl > kRedSlices ? kRedSlices : nsl);
  for (int w = threadIdx.x; w < E * nsl; w += blockDim.x) {
    const int t = w % E, q = w / E;
    int j = 0, rest = t;  // packed lower triangle, column by column: column j holds N - j entries
    while (rest >= N - j) {
      rest -= N - j;
      ++j;
    }
    const int e = (j + rest) + N * j;
    double re = 0.0, im = 0.0;
    int p = q;
    for (; p + 15 * nsl < nparts; p += 16 * nsl) {
      cd v[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) v[u] = gpart[static_cast<size_t>(p + u * nsl) * nn + e];
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        re += v[u].x;
        im += v[u].y;
      }
    }
    for (; p + 3 * nsl < nparts; p += 4 * nsl) {
      cd v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = gpart[static_cast<size_t>(p + u * nsl) * nn + e];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        re += v[u].x;
        im += v[u].y;
      }
    }
    for (; p < nparts; p += nsl) {
      const cd a = gpart[static_cast<size_t>(p) * nn + e];
      re += a.x;
      im += a.y;
    }
    scratch[w] = cmake(re, im);
  }
  __syncthreads();
  for (int t = threadIdx.x; t < E; t += blockDim.x) {
    int j = 0, rest = t;
    while (rest >= N - j) {
      rest -= N - j;
      ++j;
    }
    const int i = j + rest;
    cd sum = scratch[t];
    for (int q = 1; q < nsl; ++q) sum = cadd(sum, scratch[q * E + t]);
    G[i + N * j] = sum;
    if (i != j) G[j + N * i] = cconj(sum);
  }
  __syncthreads();
}

// One entry of column step k.  A warp holds entries of all three kinds (column k, trailing block, untouched),
// so the cases are evaluated without branches and selected: a divergent warp would walk through the three
// dependent chains one after the other, and these kernels are nothing but latency.  Every entry's arithmetic
// is that of its own case, unchanged.
__device__ __forceinline__ cd chol_entry(const cd* src, int N, int e, int i, int j, int k, double x, double r, double r2) {
  const cd v = src[e];
  const cd aik = src[i + N * k], ajk = src[j + N * k];
  cd upd = v;
  cmsub(upd, cscale(aik, r2), cconj(ajk));
  const cd col = (i == k) ? cmake(x * r, 0.0) : cscale(v, r);
  const bool active = i >= j && j >= k;
  return active ? (j == k ? col : upd) : v;
}

// Cholesky of the Hermitian G (real diagonal + lower triangle are read, as Eigen LLT.h:301-328),
// returned as R = L^dag (upper, exactly zero below the diagonal), as fields.hpp:142.  Returns -1 or
// the index of the first non-positive pivot (uniform across the CTA).  Lw: N*N scratch.
// Right-looking, one matrix entry per thread and ONE barrier per column (the factorisation
// ping-pongs between Lw and R): column k is scaled by 1/sqrt(pivot) and the trailing block gets
// its rank-1 update in the same step; the reciprocal square root replaces sqrt + division.
// (Eigen's unblocked LLT is left-looking: same factor, sums associated differently.)
__device__ __noinline__ int sm_chol_upper(cd* R, const cd* G, cd* Lw, int N, int* s_info) {
  const int nn = N * N;
  // one entry per thread (the shape the coefficient kernels are launched with): (i, j) stay in registers
  const bool one = nn <= static_cast<int>(blockDim.x);
  const int e0 = threadIdx.x;
  int i0 = 0, j0 = 0;
  if (one && e0 < nn) {
    BCG_IJ(e0, i, j);
    i0 = i;
    j0 = j;
  }
  sm_copy(Lw, G, nn);
  cd* src = Lw;
  cd* dst = R;
  int info = -1;
  for (int k = 0; k < N; ++k) {
    const double x = src[k + N * k].x;
    if (!(x > 0.0)) {
      info = k;
      break;
    }
    const double r = rsqrt(x), r2 = r * r;
    if (one) {
      if (e0 < nn) dst[e0] = chol_entry(src, N, e0, i0, j0, k, x, r, r2);
    } else {
      for (int e = threadIdx.x; e < nn; e += blockDim.x) {
        BCG_IJ(e, i, j);
        dst[e] = chol_entry(src, N, e, i, j, k, x, r, r2);
      }
    }
    __syncthreads();
    cd* t = src;
    src = dst;
    dst = t;
  }
  if (src != Lw) {
    for (int e = threadIdx.x; e < nn; e += blockDim.x) Lw[e] = src[e];
    __syncthreads();
  }
  for (int e = threadIdx.x; e < nn; e += blockDim.x) {
    BCG_IJ(e, i, j);
    R[e] = (j >= i) ? cconj(Lw[j + N * i]) : czero();
  }
  __syncthreads();
  if (threadIdx.x == 0) *s_info = info;
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

// The factorisation is a chain of ~10 N dependent little steps: it runs in ONE warp (the steps
// are separated by __syncwarp, an order of magnitude cheaper than a CTA barrier) while the
// other warps of the CTA wait at the closing __syncthreads.
// Lanes form a 16 x 2 grid over (row, column) so that no loop needs an integer division; pivots
// are compared by squared modulus (same order as Eigen's abs unless two candidates agree to
// rounding), and divisions by a pivot become one reciprocal per step plus multiplications.
__device__ __noinline__ void warp_lu_solve(cd* X, cd* lu, const LuWork& w, int N) {
  const int lane = threadIdx.x & 31, li = lane & 15, lj = lane >> 4;
  double maxpivot2 = 0.0;
  int nonzero = N;
  for (int k = 0; k < N; ++k) {
    // ---- pivot search over the trailing (N-k)^2 block ----
    double bv = -1.0;
    int bi = 0x7fffffff;
    for (int j = k + lj; j < N; j += 2)
      for (int i = k + li; i < N; i += 16) {
        const double a = cabs2(lu[i + N * j]);
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
    if (bv == 0.0) {  // the whole trailing block vanishes
      nonzero = k;
      for (int i = k + lane; i < N; i += 32) w.rt[i] = w.ct[i] = i;
      __syncwarp();
      break;
    }
    if (bv > maxpivot2) maxpivot2 = bv;
    const int bc = bi / N, br = bi - bc * N;
    if (lane == 0) {
      w.rt[k] = br;
      w.ct[k] = bc;
    }
    if (br != k)
      for (int j = lane; j < N; j += 32) {
        const cd t = lu[k + N * j];
        lu[k + N * j] = lu[br + N * j];
        lu[br + N * j] = t;
      }
    __syncwarp();
    if (bc != k)
      for (int i = lane; i < N; i += 32) {
        const cd t = lu[i + N * k];
        lu[i + N * k] = lu[i + N * bc];
        lu[i + N * bc] = t;
      }
    __syncwarp();
    const cd rpiv = cdiv(cmake(1.0, 0.0), lu[k + N * k]);
    for (int i = k + 1 + lane; i < N; i += 32) lu[i + N * k] = cmul(lu[i + N * k], rpiv);
    __syncwarp();
    for (int j = k + 1 + lj; j < N; j += 2)
      for (int i = k + 1 + li; i < N; i += 16) cmsub(lu[i + N * j], lu[i + N * k], lu[k + N * j]);
    __syncwarp();
  }
  // ---- permutations and rank (FullPivLU.h:317-341: pivots above eps * N * |maxpivot| count) ----
  if (lane == 0) {
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
  }
  const double thr = sqrt(maxpivot2) * (kEps * N);
  int rank = 0;
  for (int i0 = 0; i0 < nonzero; i0 += 32) {
    const int i = i0 + lane;
    const bool big = (i < nonzero) && (cabs2(lu[i + N * i]) > thr * thr);
    rank += __popc(__ballot_sync(0xffffffffu, big));
  }
  __syncwarp();
  if (rank == 0) {
    for (int e = lane; e < N * N; e += 32) X[e] = czero();
    return;
  }
  // c = P * B : row p[i] of c = row i of B ; reciprocals of the pivots into the diagonal of lu
  for (int col = lj; col < N; col += 2)
    for (int i = li; i < N; i += 16) w.c[w.p[i] + N * col] = X[i + N * col];
  for (int i = lane; i < rank; i += 32) lu[i + N * i] = cdiv(cmake(1.0, 0.0), lu[i + N * i]);
  __syncwarp();
  // unit-lower forward substitution, column oriented
  for (int j = 0; j < N - 1; ++j) {
    for (int col = lj; col < N; col += 2)
      for (int i = j + 1 + li; i < N; i += 16) cmsub(w.c[i + N * col], lu[i + N * j], w.c[j + N * col]);
    __syncwarp();
  }
  // upper backward substitution on the leading rank x rank block
  for (int i = rank - 1; i >= 0; --i) {
    const cd rd = lu[i + N * i];
    for (int col = lane; col < N; col += 32) w.c[i + N * col] = cmul(w.c[i + N * col], rd);
    __syncwarp();
    for (int col = lj; col < N; col += 2)
      for (int r = li; r < i; r += 16) cmsub(w.c[r + N * col], lu[r + N * i], w.c[i + N * col]);
    __syncwarp();
  }
  for (int col = lj; col < N; col += 2)
    for (int i = li; i < N; i += 16) X[w.q[i] + N * col] = (i < rank) ? w.c[i + N * col] : czero();
}
__device__ __forceinline__ void sm_lu_solve(cd* X, cd* lu, const LuWork& w, int N) {
  if (threadIdx.x < 32) warp_lu_solve(X, lu, w, N);
  __syncthreads();
}

// Inverse by Gauss-Jordan elimination with row pivoting, one matrix entry per thread and one
// CTA barrier per column.  This replaces the reference's fullPivLu().solve(I)
// (block_solvers.hpp:142,166) on the device: a sequential full-pivot LU is a chain of ~10 N
// dependent steps in one warp (measured 90k cycles at N = 12, i.e. ~45 us in every iteration),
// the elimination below exposes N^2-way parallelism and takes a few microseconds.  Both are
// backward stable on the matrices of this path (alpha^-1 = P^dag T is Hermitian positive
// definite, beta_s^-1 = I + O(shift) perturbation); Eigen's rank truncation of pivots below
// eps * N * |maxpivot| (FullPivLU.h:317-341) is NOT reproduced: a vanishing pivot raises
// *s_info instead (divergence documented in DESIGN.md).
// piv: N ints of scratch; s_info: -1 on success, else the column whose pivot column vanished.
// The elimination ping-pongs between A and the scratch matrix W (one barrier per column);
// every warp finds the pivot row itself (16 candidates per shuffle tree, no hand-off).
// PIVOT = false takes the diagonal entry (Hermitian positive definite input).
// (one entry of an elimination step: gj_entry in common.cuh, shared with the folded A-step of axpy_pipe.cuh)
template <bool PIVOT>
__device__ __noinline__ void sm_inverse(cd* A, cd* W, int N, int* piv, int* s_info) {
  const int tid = threadIdx.x, lane = tid & 31, nn = N * N;
  const bool one = nn <= static_cast<int>(blockDim.x);  // one entry per thread: (i, j) stay in registers
  int i0 = 0, j0 = 0;
  if (one && tid < nn) {
    BCG_IJ(tid, i, j);
    i0 = i;
    j0 = j;
  }
  if (tid == 0) *s_info = -1;
  cd* src = A;
  cd* dst = W;
  for (int k = 0; k < N; ++k) {
    int br = k;
    if (PIVOT) {
      // row of (nearly) largest |A(r,k)|, r >= k: the top 26 bits of |a|^2 (sign 0, exponent, 15
      // mantissa bits) and the row number in the low 6 bits make one 32-bit key per candidate;
      // one hardware max-reduction picks the winner.  Candidates that agree to 5 digits are
      // equally good pivots; among those the highest row wins.
      unsigned key = 0u;
      for (int r = k + lane; r < N; r += 32) {
        const unsigned kr = (static_cast<unsigned>(__double2hiint(cabs2(src[r + N * k]))) & ~63u) | static_cast<unsigned>(r);
        key = kr > key ? kr : key;
      }
      key = __reduce_max_sync(0xffffffffu, key);
      br = static_cast<int>(key & 63u);
      if (br < k) br = k;  // all candidates zero
      if (tid == 0) piv[k] = br;
    }
    const cd pk = src[br + N * k];
    const double pa = cabs2(pk);
    if (tid == 0 && !(pa > 0.0)) *s_info = k;
    const double pn = __drcp_rn(pa);
    const cd rp = cmake(pk.x * pn, -pk.y * pn);  // 1 / pivot
    if (one) {
      if (tid < nn) dst[tid] = gj_entry(src, N, i0, j0, k, br, rp);
    } else {
      for (int e = tid; e < nn; e += blockDim.x) {
        BCG_IJ(e, i, j);
        dst[e] = gj_entry(src, N, i, j, k, br, rp);
      }
    }
    __syncthreads();
    cd* t = src;
    src = dst;
    dst = t;
  }
  if (PIVOT) {
    // undo the row exchanges as column exchanges, last first
    for (int k = N - 1; k >= 0; --k) {
      const int br = piv[k];
      if (br != k) {
        for (int i = tid; i < N; i += blockDim.x) {
          const cd t = src[i + N * k];
          src[i + N * k] = src[i + N * br];
          src[i + N * br] = t;
        }
        __syncthreads();
      }
    }
  }
  if (src != A) {
    for (int e = tid; e < nn; e += blockDim.x) A[e] = src[e];
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

// Slab decomposition: wait until every rank's block with sequence number k has landed in this
// rank's communication buffer, then point (src, nsrc) at the nranks blocks (summed in rank order
// by sm_reduce_gram: every rank computes the bit-identical matrix).  Returns false on time-out.
__device__ __forceinline__ bool sm_wait_peers(const GramWait& w, unsigned long long k, int nn, const cd*& src,
                                              int& nsrc, Ctrl* ctrl) {
  if (w.nranks == 0) return true;
  __shared__ int timed_out;
  if (threadIdx.x == 0) timed_out = 0;
  __syncthreads();
  if (threadIdx.x < w.nranks) {
    const long long t0 = clock64();
    while (ld_acquire_sys(w.seq + threadIdx.x) < k)
      if (clock64() - t0 > kSpinTimeoutClocks) {
        timed_out = 1;
        break;
      }
  }
  __syncthreads();
  if (timed_out) {
    if (threadIdx.x == 0) {
      ctrl->status = 4;
      ctrl->done = 1;
    }
    return false;
  }
  src = w.slots + static_cast<size_t>(k & 1ull) * w.nranks * nn;
  nsrc = w.nranks;
  return true;
}

// ---- stand-alone helpers used by the primitives (bcg_gram, bcg_thinqr) -------------------
// out = reduced Gram
__global__ void __launch_bounds__(kSmallThreads)
gram_reduce_kernel(cd* __restrict__ out, const cd* __restrict__ gpart, int nparts, int N) {
  extern __shared__ __align__(16) unsigned char raw[];
  cd* G = reinterpret_cast<cd*>(raw);
  sm_init_ij(N);
  sm_reduce_gram(G, gpart, nparts, N, G + N * N);
  for (int e = threadIdx.x; e < N * N; e += blockDim.x) out[e] = G[e];
}

// R = chol(G)^dag  (G given reduced in device memory); status -> ctrl->status
__global__ void __launch_bounds__(kSmallThreads)
chol_kernel(cd* __restrict__ R, const cd* __restrict__ G, int N, Ctrl* __restrict__ ctrl) {
  extern __shared__ __align__(16) unsigned char raw[];
  SmallSmem s;
  s.carve(raw, N);
  sm_init_ij(N);
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
  sm_init_ij(N);
  sm_reduce_gram(s.mat[0], gpart, nparts, N, s.mat[4]);
  const int info = sm_chol_upper(s.mat[1], s.mat[0], s.mat[2], N, s.info);
  sm_identity(s.mat[3], N);
  sm_rownorms(s.vec, s.mat[1], N);
  for (int e = threadIdx.x; e < nn; e += blockDim.x) {
    const cd d = s.mat[1][e], id = s.mat[3][e];
    mats[L.fixed(M_DELTA) + e] = d;
    mats[L.fixed(M_RHO0) + e] = d;
    mats[L.fixed(M_RHO1) + e] = d;
    mats[L.fixed(M_RHO_CUR) + e] = ((g_ij[e] & 0xffff) == (g_ij[e] >> 16)) ? cdiv(cmake(1.0, 0.0), d) : d;
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

// A-step (after the stencil), one CTA per shift.
//   every CTA : alpha^-1 = herm(P0^dag T) ; alpha = its inverse        (block_solvers.hpp:139-142)
//   CTA 0     : A_0 = alpha*delta (old delta!) ; -alpha ; iteration bookkeeping (iter++, retire
//               converged shifts)                                     (block_solvers.hpp:145-148)
//   CTA s>=1  : beta_s of this iteration (block_solvers.hpp:163-166).  It depends on alpha, rho_old,
//               alpha_inv_old and the old beta_s only -- not on the rho that the B-step will
//               compute -- so this half of the shifted coefficients (five products and a pivoted
//               inverse) runs here, next to the A-step's own inverse; the B-step does the other
//               half (alpha_s, residual, operands).  Each kernel is as long as its longest CTA, so
//               the chain is split where the two kernels come out about equal.
// The CTAs may not race on the control block: all of them derive the iteration number and the
// active-shift count from the copies (iter_b, n_unconv_b) the previous B-step left behind, and
// only CTA 0 writes iter / n_unconv.
__global__ void __launch_bounds__(kSmallThreads)
rq_step_a_kernel(cd* __restrict__ mats, MatLayout L, const cd* __restrict__ gpart, int nparts,
                 Ctrl* __restrict__ ctrl, const GramWait gw) {
  pdl_wait();
  pdl_trigger();
  // Everything the prologue needs from the control block in ONE round trip to L2 (the loads are independent and
  // issued before the first branch): these kernels are chains of latencies, and a dependent global load is ~1/3 us.
  const int c_done = ctrl->done, c_stop = ctrl->stop;
  const int iter = ctrl->iter_b + 1;
  const int n_unconv_old = ctrl->n_unconv_b;
  unsigned conv_mask = 0u;
#pragma unroll
  for (int q = 1; q < kMaxShifts; ++q) conv_mask |= (ctrl->conv[q] != 0 ? 1u : 0u) << q;
  const unsigned long long seq_base = gw.nranks ? ctrl->seq_base : 0ull;
  if (c_done) return;
  if (c_stop) {
    __syncthreads();
    if (threadIdx.x == 0) ctrl->done = 1;
    return;
  }
  const int sh = blockIdx.x;
  extern __shared__ __align__(16) unsigned char raw[];
  const int N = L.N, nn = N * N;
  // shifts that passed the test in the previous iteration were still updated in it and drop out
  // from this one on (block_solvers.hpp:161,175-181): every passing shift decrements the count,
  // which always retires the highest index.
  int n_unconv = n_unconv_old;
  for (int q = 1; q < n_unconv_old; ++q)
    if ((conv_mask >> q) & 1u) --n_unconv;
  if (sh >= n_unconv && sh > 0) return;
  // the operands of the second half of this kernel, fetched now (one entry per thread) so that their latency
  // hides behind the Gram wait and the first inverse; all of them were written by earlier kernels
  const bool one = nn <= static_cast<int>(blockDim.x);
  const bool own = one && static_cast<int>(threadIdx.x) < nn;
  const cd* rho_old_g = mats + L.fixed((iter & 1) ? M_RHO0 : M_RHO1);
  const cd* ainv_old_g = mats + L.fixed((iter & 1) ? M_ALPHA_INV0 : M_ALPHA_INV1);
  const double sig_s = ctrl->sigma[sh], sig_0 = ctrl->sigma[0];
  cd pre0 = czero(), pre1 = czero(), pre2 = czero();
  if (own) {
    if (sh == 0) {
      pre0 = mats[L.fixed(M_DELTA) + threadIdx.x];
    } else {
      pre0 = rho_old_g[threadIdx.x];
      pre1 = ainv_old_g[threadIdx.x];
      pre2 = mats[L.beta_s(sh) + threadIdx.x];
    }
  }
  SmallSmem s;
  s.carve(raw, N);
  sm_init_ij(N);
  if (sh == 0 && threadIdx.x == 0) {
    ctrl->iter = iter;
    ctrl->n_unconv = n_unconv;
  }
  cd* Ainv = s.mat[0];
  cd* lu = s.mat[1];
  cd* alpha = s.mat[2];
  const cd* gsrc = gpart;
  int nsrc = nparts;
  if (!sm_wait_peers(gw, seq_base + static_cast<unsigned long long>(iter), nn, gsrc, nsrc, ctrl)) return;
  sm_reduce_gram(Ainv, gsrc, nsrc, N, s.mat[4]);
  sm_copy(alpha, Ainv, nn);
  sm_inverse<false>(alpha, lu, N, s.lw.rt, s.info);  // alpha = (P0^dag T)^-1, Hermitian positive definite
  if (sh == 0) {
    cd* delta = s.mat[3];
    cd* ad = s.mat[4];
    if (threadIdx.x == 0 && *s.info >= 0) {  // singular P0^dag T: the operator is not positive definite
      ctrl->status = 3;
      ctrl->stop = 1;
    }
    if (one) {
      if (own) delta[threadIdx.x] = pre0;
      __syncthreads();
    } else {
      sm_copy(delta, mats + L.fixed(M_DELTA), nn);
    }
    sm_mm(ad, alpha, delta, N);
    cd* ainv_g = mats + L.fixed((iter & 1) ? M_ALPHA_INV1 : M_ALPHA_INV0);
    for (int e = threadIdx.x; e < nn; e += blockDim.x) {
      ainv_g[e] = Ainv[e];
      mats[L.fixed(M_ALPHA) + e] = alpha[e];
      mats[L.fixed(M_NEGALPHA) + e] = cmake(-alpha[e].x, -alpha[e].y);
      mats[L.A(0, iter) + g_il[e]] = ad[e];
    }
    return;
  }
  // ---- shifted coefficients that do not need the new rho ----
  cd* rho_old = s.mat[4];
  cd* ainv_old = s.mat[5];
  cd* beta = s.mat[6];
  cd* t1 = s.mat[7];
  cd* t2 = s.mat[8];
  cd* G = s.mat[0];  // Ainv is not needed any more: scratch
  if (one) {
    if (own) {
      rho_old[threadIdx.x] = pre0;
      ainv_old[threadIdx.x] = pre1;
      beta[threadIdx.x] = pre2;
    }
  } else {
    for (int e = threadIdx.x; e < nn; e += blockDim.x) {
      rho_old[e] = rho_old_g[e];
      ainv_old[e] = ainv_old_g[e];
      beta[e] = mats[L.beta_s(sh) + e];
    }
  }
  __syncthreads();
  // beta_s_inv = I + (sigma_s - sigma_0) alpha + alpha rho_old alpha_inv_old (I - beta_s) rho_old^dag
  sm_mm(t1, alpha, rho_old, N);
  sm_mm(t2, t1, ainv_old, N);
  for (int e = threadIdx.x; e < nn; e += blockDim.x) {
    const double id = ((g_ij[e] & 0xffff) == (g_ij[e] >> 16)) ? 1.0 : 0.0;
    t1[e] = cmake(id - beta[e].x, -beta[e].y);
  }
  __syncthreads();
  sm_mm(G, t2, t1, N);
  sm_mm_adj(t1, G, rho_old, N);  // ... * rho_old^dag
  const double ds = sig_s - sig_0;
  for (int e = threadIdx.x; e < nn; e += blockDim.x) {
    const double id = ((g_ij[e] & 0xffff) == (g_ij[e] >> 16)) ? 1.0 : 0.0;
    lu[e] = cmake((id + ds * alpha[e].x) + t1[e].x, (ds * alpha[e].y) + t1[e].y);
  }
  sm_copy(beta, lu, nn);
  sm_inverse<true>(beta, lu, N, s.lw.rt, s.info);  // beta_s = beta_s_inv^-1 (general complex matrix)
  for (int e = threadIdx.x; e < nn; e += blockDim.x) mats[L.beta_s(sh) + e] = beta[e];
}

// B-step (after Q -= T alpha and its Gram): one CTA per shift.
//   every CTA : G = herm(Q^dag Q) ; rho = chol(G)^dag            (fields.hpp:142)
//   CTA 0     : delta = rho*delta ; residual ; stop test ; B_0 = rho^dag
//               (block_solvers.hpp:152-158)
//   CTA s>=1  : alpha_s (with the beta_s the A-step has prepared), shifted residual
//               rho alpha^-1 alpha_s, A_s = alpha_s, B_s = beta_s rho^dag   (block_solvers.hpp:167-181)
__global__ void __launch_bounds__(kSmallThreads)
rq_step_b_kernel(cd* __restrict__ mats, MatLayout L, const double* __restrict__ b_norm,
                 const cd* __restrict__ gpart, int nparts, Ctrl* __restrict__ ctrl, const GramWait gw) {
  pdl_wait();
  pdl_trigger();
  // the control block in one round trip, the operands of the second half fetched before the Gram wait (see the A-step)
  const int c_done = ctrl->done, c_n_unconv = ctrl->n_unconv, iter = ctrl->iter;
  const unsigned long long seq_base = gw.nranks ? ctrl->seq_base : 0ull;
  if (c_done) return;
  const int sh = blockIdx.x;
  if (sh >= c_n_unconv) return;
  extern __shared__ __align__(16) unsigned char raw[];
  const int N = L.N, nn = N * N;
  const bool one = nn <= static_cast<int>(blockDim.x);
  const bool own = one && static_cast<int>(threadIdx.x) < nn;
  const cd* rho_old_g = mats + L.fixed((iter & 1) ? M_RHO0 : M_RHO1);
  const cd* ainv_old_g = mats + L.fixed((iter & 1) ? M_ALPHA_INV0 : M_ALPHA_INV1);
  const cd* ainv_g = mats + L.fixed((iter & 1) ? M_ALPHA_INV1 : M_ALPHA_INV0);
  cd pre[6];
#pragma unroll
  for (int t = 0; t < 6; ++t) pre[t] = czero();
  if (own) {
    if (sh == 0) {
      pre[0] = mats[L.fixed(M_DELTA) + threadIdx.x];
    } else {
      pre[0] = mats[L.fixed(M_ALPHA) + threadIdx.x];
      pre[1] = rho_old_g[threadIdx.x];
      pre[2] = ainv_old_g[threadIdx.x];
      pre[3] = ainv_g[threadIdx.x];
      pre[4] = mats[L.beta_s(sh) + threadIdx.x];  // this iteration's beta_s, from the A-step
      pre[5] = mats[L.alpha_s(sh) + threadIdx.x];
    }
  }
  SmallSmem s;
  s.carve(raw, N);
  sm_init_ij(N);
  cd* G = s.mat[0];
  cd* rho = s.mat[1];
  cd* t0 = s.mat[2];
  const cd* gsrc = gpart;
  int nsrc = nparts;
  if (!sm_wait_peers(gw, seq_base + static_cast<unsigned long long>(iter), nn, gsrc, nsrc, ctrl)) return;
  sm_reduce_gram(G, gsrc, nsrc, N, s.mat[4]);
  const int info = sm_chol_upper(rho, G, t0, N, s.info);
  cd* rho_g = mats + L.fixed((iter & 1) ? M_RHO1 : M_RHO0);
  if (sh == 0) {
    cd* delta = s.mat[3];
    cd* dn = s.mat[4];
    if (one) {
      if (own) delta[threadIdx.x] = pre[0];
      __syncthreads();
    } else {
      sm_copy(delta, mats + L.fixed(M_DELTA), nn);
    }
    sm_mm(dn, rho, delta, N);
    sm_rownorms(s.vec, dn, N);
    for (int e = threadIdx.x; e < nn; e += blockDim.x) {
      BCG_IJ(e, i, j);
      rho_g[e] = rho[e];
      mats[L.fixed(M_RHO_CUR) + e] = (i == j) ? cmake(__drcp_rn(rho[e].x), 0.0) : rho[e];  // diagonal of chol: real > 0
      mats[L.fixed(M_DELTA) + e] = dn[e];
      mats[L.B(0, iter) + g_il[e]] = cconj(rho[j + N * i]);
    }
    if (threadIdx.x == 0) {
      double r = 0.0;
      bool nan = false;
      for (int i = 0; i < N; ++i) {
        const double v = s.vec[i] / b_norm[i];
        if (v != v) nan = true;
        r = fmax(r, v);
      }
      if (nan) r = 0.0 / 0.0;
      ctrl->residual = r;
      // the copies the next A-step reads (its CTAs must not race with CTA 0's own updates)
      ctrl->iter_b = iter;
      ctrl->n_unconv_b = ctrl->n_unconv;
      ctrl->n_act[L.set_of(iter)] = ctrl->n_unconv;
      // while (residual > eps && iter < max_iterations)  -- NaN ends the loop as in the reference
      int stop = 0;
      if (!(r > ctrl->eps) || iter >= ctrl->max_it) stop = 1;
      if (info >= 0) {
        ctrl->status = 3;
        stop = 1;
      } else if (nan) {
        ctrl->status = 6;
      }
      if (stop) ctrl->stop = 1;
      // statistics: systems updated in this iteration, bytes the multishift launch that follows will move
      const int na = ctrl->n_unconv;
      ctrl->hist[na < 0 ? 0 : (na > kMaxShifts ? kMaxShifts : na)] += 1u;
      int passes = 0;
      if (L.pair == 3) {
        int ring[kMaxDepth];
        for (int t = 0; t < kMaxDepth; ++t) ring[t] = ctrl->n_act[t];
        if (L.overlap) {
          // the launch that serves the shifted systems runs beside the next iterations: leave it this
          // iteration's state (the slot is free again: the launch that read it two iterations ago has been
          // joined before this iteration's Q update)
          Ctrl::Snap& sn = ctrl->snap[iter & 1];
          sn.stop = stop;
          sn.n_now = na;
          for (int t = 0; t < kMaxDepth; ++t) sn.n_ring[t] = ring[t];
          sn.iter = iter;
          int p2 = 0;
          build_stag_items(L.depth, L.ring, 1, iter, stop, na, ring, nullptr, &passes);
          build_stag_items(L.depth, L.ring, 2, iter, stop, na, ring, nullptr, &p2);
          passes += p2;
        } else {
          build_stag_items(L.depth, L.ring, 0, iter, stop, na, ring, nullptr, &passes);
        }
      } else {
        build_shift_items(L.pair, iter, stop, na, ctrl->n_act[(iter - 1) & 1], nullptr, &passes);
      }
      ctrl->shift_passes += static_cast<unsigned long long>(passes);
    }
    return;
  }
  // ---- alpha_s, shifted residual and the operands of the field update ----
  cd* alpha = s.mat[3];
  cd* rho_old = s.mat[4];
  cd* ainv_old = s.mat[5];
  cd* beta = s.mat[6];
  cd* t1 = s.mat[7];
  cd* t2 = s.mat[8];
  cd* as = s.mat[9];
  cd* ainv = s.mat[10];
  if (one) {
    if (own) {
      alpha[threadIdx.x] = pre[0];
      rho_old[threadIdx.x] = pre[1];
      ainv_old[threadIdx.x] = pre[2];
      ainv[threadIdx.x] = pre[3];
      beta[threadIdx.x] = pre[4];
      as[threadIdx.x] = pre[5];
    }
  } else {
    for (int e = threadIdx.x; e < nn; e += blockDim.x) {
      alpha[e] = mats[L.fixed(M_ALPHA) + e];
      rho_old[e] = rho_old_g[e];
      ainv_old[e] = ainv_old_g[e];
      ainv[e] = ainv_g[e];
      beta[e] = mats[L.beta_s(sh) + e];  // this iteration's beta_s, from the A-step
      as[e] = mats[L.alpha_s(sh) + e];
    }
  }
  __syncthreads();
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
    mats[L.A(sh, iter) + g_il[e]] = t2[e];
    mats[L.B(sh, iter) + g_il[e]] = t1[e];
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
  sm_init_ij(N);
  sm_reduce_gram(s.mat[0], gpart, nparts, N, s.mat[4]);
  for (int e = threadIdx.x; e < nn; e += blockDim.x) mats[L.fixed(M_R2) + e] = s.mat[0][e];
  for (int i = threadIdx.x; i < N; i += blockDim.x) b_norm[i] = sqrt(s.mat[0][i + N * i].x);
}
// A-step: alpha = LU(P^dag T).solve(r2) ; -alpha ; A_0 = alpha
__global__ void __launch_bounds__(kSmallThreads)
bcg_step_a_kernel(cd* __restrict__ mats, MatLayout L, const cd* __restrict__ gpart, int nparts,
                  Ctrl* __restrict__ ctrl, const GramWait gw) {
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
  sm_init_ij(N);
  const int iter = ctrl->iter + 1;
  __syncthreads();
  if (threadIdx.x == 0) ctrl->iter = iter;
  const cd* gsrc = gpart;
  int nsrc = nparts;
  if (!sm_wait_peers(gw, ctrl->seq_base + static_cast<unsigned long long>(iter), nn, gsrc, nsrc, ctrl)) return;
  sm_reduce_gram(s.mat[0], gsrc, nsrc, N, s.mat[4]);
  sm_copy(s.mat[1], mats + L.fixed(M_R2), nn);  // X enters as B = r2
  sm_lu_solve(s.mat[1], s.mat[0], s.lw, N);
  for (int e = threadIdx.x; e < nn; e += blockDim.x) {
    const cd a = s.mat[1][e];
    mats[L.fixed(M_ALPHA) + e] = a;
    mats[L.fixed(M_NEGALPHA) + e] = cmake(-a.x, -a.y);
    mats[L.A(0) + g_il[e]] = a;
  }
}
// B-step: r2_old = r2 ; r2 = R^dag R ; beta = LU(r2_old).solve(r2) ; residual ; B_0 = beta
__global__ void __launch_bounds__(kSmallThreads)
bcg_step_b_kernel(cd* __restrict__ mats, MatLayout L, const double* __restrict__ b_norm,
                  const cd* __restrict__ gpart, int nparts, Ctrl* __restrict__ ctrl, const GramWait gw) {
  if (ctrl->done) return;
  extern __shared__ __align__(16) unsigned char raw[];
  const int N = L.N, nn = N * N;
  SmallSmem s;
  s.carve(raw, N);
  sm_init_ij(N);
  cd* r2 = s.mat[0];
  cd* r2old = s.mat[1];
  cd* beta = s.mat[2];
  const cd* gsrc = gpart;
  int nsrc = nparts;
  if (!sm_wait_peers(gw, ctrl->seq_base + static_cast<unsigned long long>(ctrl->iter), nn, gsrc, nsrc, ctrl)) return;
  sm_reduce_gram(r2, gsrc, nsrc, N, s.mat[4]);
  sm_copy(r2old, mats + L.fixed(M_R2), nn);
  sm_copy(beta, r2, nn);
  sm_lu_solve(beta, r2old, s.lw, N);
  for (int e = threadIdx.x; e < nn; e += blockDim.x) {
    mats[L.fixed(M_R2) + e] = r2[e];
    mats[L.B(0) + g_il[e]] = beta[e];
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


// ---- unit-test entry points of the device small-matrix routines (bcg_small_*) ----------------
// out = in^-1 by the loop's own Gauss-Jordan inverse (pivot != 0: row pivoting, as used for
// beta_s^-1; 0: the pivot-free variant used for the Hermitian positive definite P0^dag T).
__global__ void __launch_bounds__(kSmallThreads)
small_inverse_kernel(cd* __restrict__ out, const cd* __restrict__ in, int N, int pivot, int* __restrict__ info_out) {
  extern __shared__ __align__(16) unsigned char raw[];
  SmallSmem s;
  s.carve(raw, N);
  sm_init_ij(N);
  sm_copy(s.mat[0], in, N * N);
  if (pivot) sm_inverse<true>(s.mat[0], s.mat[1], N, s.lw.rt, s.info);
  else sm_inverse<false>(s.mat[0], s.mat[1], N, s.lw.rt, s.info);
  for (int e = threadIdx.x; e < N * N; e += blockDim.x) out[e] = s.mat[0][e];
  if (threadIdx.x == 0) *info_out = *s.info;
}
// X = A^-1 B by the Eigen-faithful full-pivoting LU of the BCG loop (FullPivLU.h:487-590,745-790)
__global__ void __launch_bounds__(kSmallThreads)
small_lu_solve_kernel(cd* __restrict__ X, const cd* __restrict__ A, const cd* __restrict__ B, int N) {
  extern __shared__ __align__(16) unsigned char raw[];
  SmallSmem s;
  s.carve(raw, N);
  sm_init_ij(N);
  sm_copy(s.mat[0], A, N * N);
  sm_copy(s.mat[1], B, N * N);
  sm_lu_solve(s.mat[1], s.mat[0], s.lw, N);
  for (int e = threadIdx.x; e < N * N; e += blockDim.x) X[e] = s.mat[1][e];
}

// ---- CG / SCG: the reference's scalar-coefficient solvers for one right-hand side --------------
// (src/standard_solvers.cpp:3-32 and :34-95).  Same device-resident loop as the block solvers --
// stencil with fused p.t, r -= t alpha with fused r.r, one streaming update of all systems -- but
// the coefficients are real scalars kept in the control block: no N x N algebra at all.
// real_dot(a, b) = sum_x Re(a[x]^dag b[x]) (fields.hpp:93-100) = real part of the 1 x 1 Gram.
__device__ __forceinline__ double sc_reduce_real(const cd* __restrict__ gpart, int nparts, double* red) {
  double sum = 0.0;
  for (int p = threadIdx.x; p < nparts; p += blockDim.x) sum += gpart[p].x;  // N = 1: one complex per partial
  red[threadIdx.x] = sum;
  __syncthreads();
  for (int off = blockDim.x / 2; off > 0; off >>= 1) {  // fixed tree: deterministic
    if (static_cast<int>(threadIdx.x) < off) red[threadIdx.x] += red[threadIdx.x + off];
    __syncthreads();
  }
  return red[0];
}
constexpr int kScalarThreads = 256;
// init: r2 = r.r ; alpha = 1, beta = 0, zeta = theta = 1 (:13-14,50-53) ; eps *= sqrt(r2) is kept as r2_0
__global__ void __launch_bounds__(kScalarThreads)
scg_init_kernel(const cd* __restrict__ gpart, int nparts, Ctrl* __restrict__ ctrl) {
  __shared__ double red[kScalarThreads];
  const double r2 = sc_reduce_real(gpart, nparts, red);
  if (threadIdx.x == 0) {
    ctrl->sc_r2 = r2;
    ctrl->sc_r2_0 = r2;
    ctrl->sc_alpha = 1.0;
    ctrl->sc_beta = 0.0;
    for (int s = 0; s < kMaxShifts; ++s) ctrl->sc_zeta[s] = ctrl->sc_theta[s] = 1.0;
    ctrl->residual = 1.0;
    // while (sqrt(r2) > eps * sqrt(r2_0) && iter < max_iterations): may already be false (b = 0)
    if (!(sqrt(r2) > ctrl->eps * sqrt(r2)) || ctrl->max_it <= 0) ctrl->done = 1;
  }
}
// A-step: alpha = r2 / p0.t ; -alpha is the operand of r -= t alpha
__global__ void __launch_bounds__(kScalarThreads)
scg_step_a_kernel(cd* __restrict__ neg_alpha, const cd* __restrict__ gpart, int nparts, Ctrl* __restrict__ ctrl) {
  if (ctrl->done) return;
  if (ctrl->stop) {
    __syncthreads();
    if (threadIdx.x == 0) ctrl->done = 1;
    return;
  }
  __shared__ double red[kScalarThreads];
  const double pt = sc_reduce_real(gpart, nparts, red);
  if (threadIdx.x == 0) {
    ctrl->iter += 1;
    ctrl->sc_alpha_old = ctrl->sc_alpha;
    const double alpha = ctrl->sc_r2 / pt;
    ctrl->sc_alpha = alpha;
    *neg_alpha = cmake(-alpha, 0.0);
  }
}
// B-step: r2, beta, the shifted coefficients (:71-87) and the stopping tests (:57,89-92)
__global__ void __launch_bounds__(kScalarThreads)
scg_step_b_kernel(const cd* __restrict__ gpart, int nparts, Ctrl* __restrict__ ctrl) {
  if (ctrl->done) return;
  __shared__ double red[kScalarThreads];
  const double r2 = sc_reduce_real(gpart, nparts, red);
  if (threadIdx.x != 0) return;
  const double r2_old = ctrl->sc_r2;
  const double beta_old = ctrl->sc_beta, alpha = ctrl->sc_alpha, alpha_old = ctrl->sc_alpha_old;
  const double beta = r2 / r2_old;
  ctrl->sc_r2 = r2;
  ctrl->sc_beta_old = beta_old;
  ctrl->sc_beta = beta;
  const int na = ctrl->n_unconv;
  ctrl->sc_ax[0] = alpha;
  ctrl->sc_bp[0] = beta;
  ctrl->sc_zr[0] = 1.0;
  for (int s = na - 1; s > 0; --s) {
    double inv_theta = 1.0 + (ctrl->sigma[s] - ctrl->sigma[0]) * alpha;
    inv_theta += beta_old * (alpha / alpha_old) * (1.0 - ctrl->sc_theta[s]);
    const double theta = 1.0 / inv_theta;
    ctrl->sc_theta[s] = theta;
    const double zeta = ctrl->sc_zeta[s] * theta;
    ctrl->sc_zeta[s] = zeta;
    ctrl->sc_ax[s] = alpha * theta;
    ctrl->sc_bp[s] = beta * theta * theta;
    ctrl->sc_zr[s] = zeta;
  }
  // the update kernel of this iteration serves `serve` systems: x_0 and p_0 are updated outside the
  // reference's shift loop (:65-69), i.e. always, even once the count has dropped to zero ...
  const int serve = na > 1 ? na : 1;
  ctrl->n_act[0] = serve;
  ctrl->hist[serve > kMaxShifts ? kMaxShifts : serve] += 1u;
  ctrl->shift_passes += static_cast<unsigned long long>(1 + 4 * serve);
  // ... and the highest one is dropped from the next iteration on once its residual is small enough.  (The
  // reference reads zeta[n_unconverged_shifts - 1] unguarded, :89: zeta[-1] once the count is zero.)
  if (na > 0 && sqrt(r2) * ctrl->sc_zeta[na - 1] < ctrl->eps_shifts) ctrl->n_unconv = na - 1;
  const double rel = sqrt(r2) / sqrt(ctrl->sc_r2_0);
  ctrl->residual = rel;
  if (!(sqrt(r2) > ctrl->eps * sqrt(ctrl->sc_r2_0)) || ctrl->iter >= ctrl->max_it) ctrl->stop = 1;
  if (rel != rel) ctrl->status = 6;
}
// x_s += p_s ax_s ; p_s = p_s bp_s + r zr_s for the systems served in this iteration (:65-87): one
// streaming pass, r read once.  n = complex numbers per field (3 V).
struct ScalarPtrs {
  cd* X[kMaxShifts];
  cd* P[kMaxShifts];
};
static __global__ void __launch_bounds__(256)
scg_update_kernel(ScalarPtrs fp, const cd* __restrict__ r, long long n, const Ctrl* __restrict__ ctrl) {
  if (ctrl->done) return;
  const int na = ctrl->n_act[0];
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const cd rv = r[i];
    for (int s = 0; s < na; ++s) {
      const double ax = ctrl->sc_ax[s], bp = ctrl->sc_bp[s], zr = ctrl->sc_zr[s];
      const cd p = fp.P[s][i];
      cd x = fp.X[s][i];
      x.x += p.x * ax;  // this += rhs * scalar (fields.hpp:70-77)
      x.y += p.y * ax;
      fp.X[s][i] = x;
      // tmp = this * lhs ; tmp += rhs * rhs_multiplier (fields.hpp:85-86)
      fp.P[s][i] = cmake(p.x * bp + rv.x * zr, p.y * bp + rv.y * zr);
    }
  }
}

}  // namespace bcg
