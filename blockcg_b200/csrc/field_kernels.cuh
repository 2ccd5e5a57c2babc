// Field-sized kernels of the block-CG hot path, templated on N (= N_rhs).
//
// Layout (the reference's, inc/fields.hpp:18-30): field[x][r][c], complex128,
// c fastest; element (x, r, c) at complex index (x*N + r)*3 + c.  Device
// fields carry a 2-site halo on both ends (pointer = site 0), links likewise
// (element (i,j) of U[x] at x*9 + i + 3*j).
//
//  K1 dirac_kernel      T = (m^2 + sigma) P - D(D(P))  [+ partial Gram P^dag T]
//  K2 gram_kernel       partial Gram A^dag B
//  K3 axpy_gram_kernel  Q += T*M  [+ partial Gram Q^dag Q]
//  K4 shift_update_kernel  Q <- Q rho^-1 ; for every active shift s:
//                          X_s += P_s A_s ;  P_s <- P_s B_s + Q
//  plus the stand-alone primitives (add, rescale_add, trsm, halo wrap ...).
//
// Partial Grams: one N x N block per CTA, reduced in a fixed order by the
// small-matrix kernels (no floating-point atomics anywhere => bitwise
// reproducible run to run).
#pragma once
#include "common.cuh"

namespace bcg {

// ------------------------------------------------------------------------------------
// Gram building block: a warp accumulates one 4x4 block (ti,tj) of A^dag B over
// rows (site,colour) of two shared-memory tiles.  acc[ii][jj] += conj(a_ii) b_jj.
// ------------------------------------------------------------------------------------
template <int N>
struct GramGeom {
  static constexpr int NB = (N + 3) / 4;             // 4-wide blocks per dimension
  static constexpr int NTASK = NB * (NB + 1) / 2;    // lower-triangular blocks
};

__device__ __forceinline__ void gram_task_to_block(int t, int& ti, int& tj) {
  // t = ti*(ti+1)/2 + tj, tj <= ti
  ti = 0;
  while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
  tj = t - ti * (ti + 1) / 2;
}

template <int N>
__device__ __forceinline__ void gram_rows(const cd* __restrict__ sA, const cd* __restrict__ sB, int nrows,
                                          int ti, int tj, int row0, int rowstep, cd (&acc)[4][4]) {
  for (int row = row0; row < nrows; row += rowstep) {
    const int site = row / 3;
    const int base = site * (3 * N) + (row - 3 * site);
    cd a[4], b[4];
#pragma unroll
    for (int ii = 0; ii < 4; ++ii) {
      const int ka = 4 * ti + ii, kb = 4 * tj + ii;
      a[ii] = (N % 4 == 0 || ka < N) ? sA[base + 3 * ka] : czero();
      b[ii] = (N % 4 == 0 || kb < N) ? sB[base + 3 * kb] : czero();
    }
#pragma unroll
    for (int ii = 0; ii < 4; ++ii)
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) cmac_conj(acc[ii][jj], a[ii], b[jj]);
  }
}

// Reduce a warp's 4x4 accumulator over lanes (fixed xor tree) and let lane 0
// store it into an N x N column-major block (entries outside the matrix skipped).
template <int N>
__device__ __forceinline__ void gram_warp_store(cd (&acc)[4][4], int ti, int tj, cd* dstNN, bool accumulate) {
#pragma unroll
  for (int ii = 0; ii < 4; ++ii)
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      double re = acc[ii][jj].x, im = acc[ii][jj].y;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        re += __shfl_xor_sync(0xffffffffu, re, off);
        im += __shfl_xor_sync(0xffffffffu, im, off);
      }
      const int i = 4 * ti + ii, j = 4 * tj + jj;
      if ((threadIdx.x & 31) == 0 && i < N && j < N) {
        cd* d = dstNN + i + N * j;
        if (accumulate) {
          d->x += re;
          d->y += im;
        } else {
          *d = cmake(re, im);
        }
      }
    }
}

// CTA-level Gram bookkeeping shared by K1/K2/K3.  NW warps; warp w owns task
// (w % NTASK) and row slice (w / NTASK) of NSLICE = NW / NTASK (>= 1 required).
template <int N, int NW>
struct GramCta {
  static constexpr int NTASK = GramGeom<N>::NTASK;
  static constexpr int NSLICE = NW / NTASK;
  static_assert(NSLICE >= 1, "not enough warps for the fused Gram at this N");
  static constexpr int NACTIVE = NSLICE * NTASK;  // warps that take part

  int ti, tj, slice;
  bool active;
  cd acc[4][4];

  __device__ __forceinline__ void init() {
    const int w = threadIdx.x >> 5;
    active = w < NACTIVE;
    const int t = w % NTASK;
    slice = w / NTASK;
    gram_task_to_block(t, ti, tj);
#pragma unroll
    for (int ii = 0; ii < 4; ++ii)
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) acc[ii][jj] = czero();
  }
  __device__ __forceinline__ void accumulate(const cd* sA, const cd* sB, int nrows) {
    if (active) gram_rows<N>(sA, sB, nrows, ti, tj, slice * 32 + (threadIdx.x & 31), NSLICE * 32, acc);
  }
  // sG: shared scratch of NSLICE * N*N complex.  Writes the CTA's partial
  // (lower-triangular blocks valid) to gpart[N*N].
  __device__ __forceinline__ void finish(cd* sG, cd* __restrict__ gpart) {
    __syncthreads();
    for (int e = threadIdx.x; e < NSLICE * N * N; e += blockDim.x) sG[e] = czero();
    __syncthreads();
    if (active) gram_warp_store<N>(acc, ti, tj, sG + slice * N * N, false);
    __syncthreads();
    for (int e = threadIdx.x; e < N * N; e += blockDim.x) {
      cd s = sG[e];
#pragma unroll
      for (int sl = 1; sl < NSLICE; ++sl) s = cadd(s, sG[sl * N * N + e]);
      gpart[e] = s;
    }
  }
};

// ------------------------------------------------------------------------------------
// K1: block Dirac apply.  1-D chain operator of the reference
// (inc/dirac_op.hpp:14-21,36-43): D v[x] = 1/2 U[x] v[x+1] - 1/2 U[x-1]^dag v[x-1],
// out = m^2 v - D(D v) (+ sigma v, block_solvers.hpp:136), both sweeps in one
// kernel with the intermediate in shared memory.  Each link is fetched once per
// CTA tile and reused for all N right-hand sides.
//
// Work item = (site, group of R rhs columns): thread keeps R*3 accumulators.
// ------------------------------------------------------------------------------------
template <int N, int R>
__device__ __forceinline__ void load_cols(const cd* __restrict__ s, cd (&v)[R][3]) {
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) v[r][c] = s[r * 3 + c];
}

// acc[r][i] += sum_j U(i,j) v[r][j]
template <int R>
__device__ __forceinline__ void apply_link(const cd* __restrict__ sU, const cd (&v)[R][3], cd (&acc)[R][3]) {
#pragma unroll
  for (int j = 0; j < 3; ++j)
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const cd u = sU[i + 3 * j];
#pragma unroll
      for (int r = 0; r < R; ++r) cmac(acc[r][i], u, v[r][j]);
    }
}
// acc[r][i] -= sum_j conj(U(j,i)) v[r][j]
template <int R>
__device__ __forceinline__ void apply_link_dag_sub(const cd* __restrict__ sU, const cd (&v)[R][3],
                                                   cd (&acc)[R][3]) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const cd u = cconj(sU[j + 3 * i]);
#pragma unroll
      for (int r = 0; r < R; ++r) cmsub(acc[r][i], u, v[r][j]);
    }
}

template <int N, int R, int NT>
struct DiracGeom {
  static_assert(N % R == 0, "R must divide N");
  static constexpr int G = N / R;           // work items per site
  static constexpr int TS1 = NT / G;        // sites covered by the first sweep
  static constexpr int TS = TS1 - 2;        // output sites per tile
  static_assert(TS >= 1, "block too small for this N/R");
  static constexpr int SITE = 3 * N;        // complex per site
  static constexpr int IN_ELEMS = (TS + 4) * SITE;
  static constexpr int TMP_ELEMS = (TS + 2) * SITE;
  static constexpr int OUT_ELEMS = TS * SITE;
  static constexpr int U_ELEMS = (TS + 3) * 9;
  static constexpr int NW = NT / 32;
  static constexpr bool CAN_GRAM = (NW >= GramGeom<N>::NTASK);
  static constexpr int GRAM_ELEMS = CAN_GRAM ? (NW / GramGeom<N>::NTASK) * N * N : 0;
  static constexpr size_t SMEM_BYTES =
      sizeof(cd) * (IN_ELEMS + TMP_ELEMS + OUT_ELEMS + U_ELEMS + GRAM_ELEMS) + 16;
};

template <int N, int R, int NT, bool GRAM>
__global__ void __launch_bounds__(NT)
dirac_kernel(const cd* __restrict__ in, cd* __restrict__ out, const cd* __restrict__ U, long long V,
             double m2, double sigma, cd* __restrict__ gpart, const Ctrl* __restrict__ ctrl) {
  using Geo = DiracGeom<N, R, NT>;
  constexpr int G = Geo::G, TS = Geo::TS, SITE = Geo::SITE;
  if (ctrl != nullptr && (ctrl->done | ctrl->stop)) return;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  cd* sIn = reinterpret_cast<cd*>(smem_raw);
  cd* sTmp = sIn + Geo::IN_ELEMS;
  cd* sOut = sTmp + Geo::TMP_ELEMS;
  cd* sU = sOut + Geo::OUT_ELEMS;
  cd* sG = sU + Geo::U_ELEMS;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sG + Geo::GRAM_ELEMS);

  const int tid = threadIdx.x;
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_fence_init();
  }
  __syncthreads();

  GramCta<N, GRAM ? Geo::NW : GramGeom<N>::NTASK> gram;  // dummy geometry when !GRAM
  if (GRAM) gram.init();

  const long long ntiles = (V + TS - 1) / TS;
  uint32_t phase = 0;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long x0 = tile * TS;
    const int ns = static_cast<int>(min(static_cast<long long>(TS), V - x0));
    // ---- stage P[x0-2 .. x0+ns+2) and U[x0-2 .. x0+ns+1) with two bulk copies ----
    if (tid == 0) {
      const uint32_t bytes_in = static_cast<uint32_t>((ns + 4) * SITE * sizeof(cd));
      const uint32_t bytes_u = static_cast<uint32_t>((ns + 3) * 9 * sizeof(cd));
      mbar_arrive_expect_tx(bar, bytes_in + bytes_u);
      bulk_g2s(sIn, in + (x0 - 2) * SITE, bytes_in, bar);
      bulk_g2s(sU, U + (x0 - 2) * 9, bytes_u, bar);
    }
    mbar_wait(bar, phase);
    phase ^= 1u;

    // ---- sweep 1: tmp[y] for y = x0-1 .. x0+ns   (local ls = 0 .. ns+1) ----
    for (int item = tid; item < (ns + 2) * G; item += NT) {
      const int ls = item / G, g = item - ls * G;
      cd v[R][3], acc[R][3];
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) acc[r][c] = czero();
      load_cols<N, R>(sIn + (ls + 2) * SITE + g * R * 3, v);  // P[y+1]
      apply_link<R>(sU + (ls + 1) * 9, v, acc);               // U[y]
      load_cols<N, R>(sIn + ls * SITE + g * R * 3, v);        // P[y-1]
      apply_link_dag_sub<R>(sU + ls * 9, v, acc);             // U[y-1]^dag
      cd* t = sTmp + ls * SITE + g * R * 3;
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) t[r * 3 + c] = cscale(acc[r][c], 0.5);
    }
    if (tid == 0) bulk_wait_read0();  // previous tile's bulk store has drained sOut
    __syncthreads();

    // ---- sweep 2 + mass/shift term: out[x] for x = x0 .. x0+ns-1 ----
    for (int item = tid; item < ns * G; item += NT) {
      const int ls = item / G, g = item - ls * G;
      cd v[R][3], acc[R][3];
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) acc[r][c] = czero();
      load_cols<N, R>(sTmp + (ls + 2) * SITE + g * R * 3, v);  // tmp[x+1]
      apply_link<R>(sU + (ls + 2) * 9, v, acc);                // U[x]
      load_cols<N, R>(sTmp + ls * SITE + g * R * 3, v);        // tmp[x-1]
      apply_link_dag_sub<R>(sU + (ls + 1) * 9, v, acc);        // U[x-1]^dag
      load_cols<N, R>(sIn + (ls + 2) * SITE + g * R * 3, v);   // P[x]
      cd* o = sOut + ls * SITE + g * R * 3;
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          // lhs = -D(D rhs) + m^2 rhs   (dirac_op.hpp:42), then += sigma rhs
          cd t = cmake(fma(m2, v[r][c].x, -0.5 * acc[r][c].x), fma(m2, v[r][c].y, -0.5 * acc[r][c].y));
          t.x = fma(sigma, v[r][c].x, t.x);
          t.y = fma(sigma, v[r][c].y, t.y);
          o[r * 3 + c] = t;
        }
    }
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      bulk_s2g(out + x0 * SITE, sOut, static_cast<uint32_t>(ns * SITE * sizeof(cd)));
      bulk_commit();
    }
    // ---- fused Gram epilogue: P^dag T over this tile's rows ----
    if (GRAM) {
      gram.accumulate(sIn + 2 * SITE, sOut, 3 * ns);
    }
    __syncthreads();  // tile buffers free for the next round
  }
  if (tid == 0) bulk_wait0();
  if (GRAM) gram.finish(sG, gpart + static_cast<size_t>(blockIdx.x) * N * N);
}

// ------------------------------------------------------------------------------------
// K2: stand-alone partial Gram A^dag B.  blockIdx.y selects a group of NW tasks
// (one 4x4 block per warp) so any N fits; tiles are staged by bulk copies.
// gpart[blockIdx.x][N*N]; blocks of different y write disjoint entries.
// ------------------------------------------------------------------------------------
template <int N, int NT>
struct GramKGeom {
  static constexpr int TS = 32;
  static constexpr int SITE = 3 * N;
  static constexpr size_t SMEM_BYTES = sizeof(cd) * (2 * TS * SITE) + 16;
};

template <int N, int NT>
__global__ void __launch_bounds__(NT)
gram_kernel(const cd* __restrict__ A, const cd* __restrict__ B, long long V, cd* __restrict__ gpart,
            const Ctrl* __restrict__ ctrl) {
  using Geo = GramKGeom<N, NT>;
  constexpr int TS = Geo::TS, SITE = Geo::SITE, NW = NT / 32;
  constexpr int NTASK = GramGeom<N>::NTASK;
  if (ctrl != nullptr && ctrl->done) return;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  cd* sA = reinterpret_cast<cd*>(smem_raw);
  cd* sB = sA + TS * SITE;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + TS * SITE);
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  const bool same = (A == B);
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_fence_init();
  }
  __syncthreads();
  const int task = blockIdx.y * NW + w;
  const bool active = task < NTASK;
  int ti = 0, tj = 0;
  if (active) gram_task_to_block(task, ti, tj);
  cd acc[4][4];
#pragma unroll
  for (int ii = 0; ii < 4; ++ii)
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) acc[ii][jj] = czero();

  const long long ntiles = (V + TS - 1) / TS;
  uint32_t phase = 0;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long x0 = tile * TS;
    const int ns = static_cast<int>(min(static_cast<long long>(TS), V - x0));
    if (tid == 0) {
      const uint32_t bytes = static_cast<uint32_t>(ns * SITE * sizeof(cd));
      mbar_arrive_expect_tx(bar, same ? bytes : 2 * bytes);
      bulk_g2s(sA, A + x0 * SITE, bytes, bar);
      if (!same) bulk_g2s(sB, B + x0 * SITE, bytes, bar);
    }
    mbar_wait(bar, phase);
    phase ^= 1u;
    if (active) gram_rows<N>(sA, same ? sA : sB, 3 * ns, ti, tj, lane, 32, acc);
    __syncthreads();
  }
  if (active) gram_warp_store<N>(acc, ti, tj, gpart + static_cast<size_t>(blockIdx.x) * N * N, false);
}

// ------------------------------------------------------------------------------------
// Row helpers: one thread owns one (site, colour) row of N complex numbers,
// element k at stride 3.  out[j] (+)= sum_k p[k] M(k,j), M column-major in smem.
// ------------------------------------------------------------------------------------
template <int N>
__device__ __forceinline__ void row_load(const cd* __restrict__ f, long long row_base, cd (&r)[N]) {
#pragma unroll
  for (int k = 0; k < N; ++k) r[k] = f[row_base + 3 * k];
}
template <int N>
__device__ __forceinline__ void row_store(cd* __restrict__ f, long long row_base, const cd (&r)[N]) {
#pragma unroll
  for (int k = 0; k < N; ++k) f[row_base + 3 * k] = r[k];
}
// same product with the operand in the interleaved layout of shift_mat_index()
template <int N>
__device__ __forceinline__ void row_mm_acc_il(cd (&out)[N], const cd (&p)[N], const cd* __restrict__ sM) {
  constexpr int NS = shift_nsplit(N), JC = N / NS;
#pragma unroll
  for (int k = 0; k < N; ++k)
#pragma unroll
    for (int j = 0; j < N; ++j) cmac(out[j], p[k], lds_cd(sM + (k * JC + (j % JC)) * NS + j / JC));
}
template <int N>
__device__ __forceinline__ void row_mm_acc(cd (&out)[N], const cd (&p)[N], const cd* __restrict__ sM) {
  // k outer / j inner: N independent accumulator chains between dependent DFMAs
#pragma unroll
  for (int k = 0; k < N; ++k)
#pragma unroll
    for (int j = 0; j < N; ++j) cmac(out[j], p[k], lds_cd(sM + k + N * j));
}
__device__ __forceinline__ long long row_base_of(long long row, int N) {
  const long long site = row / 3;
  return site * (3 * N) + (row - 3 * site);
}

// K3: Q += T*M with the partial Gram Q^dag Q of the updated rows fused in.
// CTA tile = NT rows = NT/3 sites (NT % 3 == 0); updated rows are staged in
// shared memory for the Gram warps.
template <int N, int NT>
struct AxpyGeom {
  static_assert(NT % 3 == 0, "block must hold whole sites");
  static constexpr int TS = NT / 3;
  static constexpr int NW = NT / 32;
  static constexpr bool CAN_GRAM = (NW >= GramGeom<N>::NTASK);
  static constexpr int GRAM_ELEMS = CAN_GRAM ? (NW / GramGeom<N>::NTASK) * N * N : 0;
  static constexpr size_t SMEM_BYTES = sizeof(cd) * (N * N + (CAN_GRAM ? TS * 3 * N : 0) + GRAM_ELEMS);
};

template <int N, int NT, bool GRAM>
__global__ void __launch_bounds__(NT)
axpy_gram_kernel(cd* __restrict__ Q, const cd* __restrict__ T, const cd* __restrict__ M, long long V,
                 cd* __restrict__ gpart, const Ctrl* __restrict__ ctrl) {
  using Geo = AxpyGeom<N, NT>;
  if (ctrl != nullptr && ctrl->done) return;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  cd* sM = reinterpret_cast<cd*>(smem_raw);
  cd* sQ = sM + N * N;
  cd* sG = sQ + (GRAM ? Geo::TS * 3 * N : 0);
  const int tid = threadIdx.x;
  for (int e = tid; e < N * N; e += NT) sM[e] = M[e];
  GramCta<N, GRAM ? Geo::NW : GramGeom<N>::NTASK> gram;
  if (GRAM) gram.init();
  __syncthreads();
  const long long nrows = 3 * V;
  const long long ntiles = (nrows + NT - 1) / NT;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long row = tile * NT + tid;
    const int rows_here = static_cast<int>(min(static_cast<long long>(NT), nrows - tile * NT));
    if (row < nrows) {
      const long long rb = row_base_of(row, N);
      cd t[N], q[N], dq[N];
      row_load<N>(T, rb, t);
      row_load<N>(Q, rb, q);
#pragma unroll
      for (int j = 0; j < N; ++j) dq[j] = czero();
      row_mm_acc<N>(dq, t, sM);
#pragma unroll
      for (int j = 0; j < N; ++j) q[j] = cadd(q[j], dq[j]);  // product first, then one addition (fields.hpp:74)
      row_store<N>(Q, rb, q);
      if (GRAM) {
        const int lb = (tid / 3) * (3 * N) + (tid % 3);
#pragma unroll
        for (int k = 0; k < N; ++k) sQ[lb + 3 * k] = q[k];
      }
    }
    if (GRAM) {
      __syncthreads();
      gram.accumulate(sQ, sQ, rows_here);
      __syncthreads();
    }
  }
  if (GRAM) gram.finish(sG, gpart + static_cast<size_t>(blockIdx.x) * N * N);
}

// dst = dst*L + r*src  (fields.hpp:79-90);  with L == nullptr: dst += src*Madd (fields.hpp:70-77)
template <int N, int NT>
__global__ void __launch_bounds__(NT)
rescale_add_kernel(cd* __restrict__ dst, const cd* __restrict__ L, const cd* __restrict__ src, double r,
                   long long V) {
  __shared__ cd sM[N * N];
  const int tid = threadIdx.x;
  for (int e = tid; e < N * N; e += NT) sM[e] = L[e];
  __syncthreads();
  const long long nrows = 3 * V;
  for (long long row = static_cast<long long>(blockIdx.x) * NT + tid; row < nrows;
       row += static_cast<long long>(gridDim.x) * NT) {
    const long long rb = row_base_of(row, N);
    cd d[N], s[N], o[N];
    row_load<N>(dst, rb, d);
    row_load<N>(src, rb, s);
#pragma unroll
    for (int j = 0; j < N; ++j) o[j] = czero();
    row_mm_acc<N>(o, d, sM);
#pragma unroll
    for (int j = 0; j < N; ++j) {
      o[j].x = fma(s[j].x, r, o[j].x);
      o[j].y = fma(s[j].y, r, o[j].y);
    }
    row_store<N>(dst, rb, o);
  }
}

// dst += s*src, element-wise (block_solvers.hpp:136) ; also plain copies / sets
static __global__ void axpy_scalar_kernel(cd* __restrict__ dst, const cd* __restrict__ src, double s, long long n) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    cd d = dst[i], v = src[i];
    d.x = fma(v.x, s, d.x);
    d.y = fma(v.y, s, d.y);
    dst[i] = d;
  }
}
// dst = a - b (verification: AX -= B)
static __global__ void sub_kernel(cd* dst, const cd* a, const cd* __restrict__ b, long long n) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    dst[i] = csub(a[i], b[i]);
}

// In-register back substitution of one row: q <- q R^-1 in the reference's
// column order (fields.hpp:125-136); R upper triangular, column-major in smem.
// The diagonal of sR must already hold the RECIPROCALS 1/R(i,i) (see load_tri_recip):
// one complex multiply per column instead of a double-precision division per row.
template <int N>
__device__ __forceinline__ void row_backsub(cd (&q)[N], const cd* __restrict__ sR) {
#pragma unroll
  for (int i = 0; i < N; ++i) {
#pragma unroll
    for (int j = 0; j < i; ++j) cmsub(q[i], lds_cd(sR + j + N * i), q[j]);
    q[i] = cmul(q[i], lds_cd(sR + i + N * i));
  }
}
// stage an upper-triangular R into shared memory with its diagonal inverted
template <int N>
__device__ __forceinline__ void load_tri_recip(cd* sR, const cd* __restrict__ Rm) {
  for (int e = threadIdx.x; e < N * N; e += blockDim.x) {
    cd v = Rm[e];
    if (e % N == e / N) v = cdiv(cmake(1.0, 0.0), v);
    sR[e] = v;
  }
}

template <int N, int NT>
__global__ void __launch_bounds__(NT)
trsm_kernel(cd* __restrict__ Q, const cd* __restrict__ Rm, long long V, const Ctrl* __restrict__ ctrl) {
  if (ctrl != nullptr && ctrl->done) return;
  __shared__ cd sR[N * N];
  const int tid = threadIdx.x;
  load_tri_recip<N>(sR, Rm);
  __syncthreads();
  const long long nrows = 3 * V;
  for (long long row = static_cast<long long>(blockIdx.x) * NT + tid; row < nrows;
       row += static_cast<long long>(gridDim.x) * NT) {
    const long long rb = row_base_of(row, N);
    cd q[N];
    row_load<N>(Q, rb, q);
    row_backsub<N>(q, sR);
    row_store<N>(Q, rb, q);
  }
}

// ------------------------------------------------------------------------------------
// K4: the multishift update, one streaming pass over 2 + 4*S_active fields:
//   Q  <- Q rho^-1                                (thinQR back substitution)
//   X_s += P_s A_s ;  P_s <- P_s B_s + Q          for s = 0 .. n_active-1
// with A_0 = alpha*delta_old, B_0 = rho^dag, A_s = alpha_s, B_s = beta_s rho^dag
// (block_solvers.hpp:145,152,158,175,177).  Q is read once and kept in
// registers across all shifts.  n_active is read from the control block, so
// converged shifts drop out without host involvement.
// With do_backsub == 0 the kernel is the BCG update (X += P A; P <- P B + R).
// ------------------------------------------------------------------------------------
struct ShiftPtrs {
  cd* X[kMaxShifts];
  cd* P[kMaxShifts];
};

template <int N, int NT>
__global__ void __launch_bounds__(NT, (N <= 12) ? 2 : 1)
shift_update_kernel(cd* __restrict__ Q, ShiftPtrs fp, const cd* __restrict__ Rm,
                    const cd* __restrict__ Amats, const cd* __restrict__ Bmats, long long V,
                    int do_backsub, int n_active_fixed, const Ctrl* __restrict__ ctrl) {
  if (ctrl != nullptr && ctrl->done) return;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  cd* sR = reinterpret_cast<cd*>(smem_raw);
  const int tid = threadIdx.x;
  const int n_active = (n_active_fixed > 0) ? n_active_fixed : ctrl->n_unconv;
  if (do_backsub)
    for (int e = tid; e < N * N; e += NT) sR[e] = Rm[e];  // diagonal already inverted (M_RHO_RECIP)
  const long long nrows = 3 * V;
  const long long ntiles = (nrows + NT - 1) / NT;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long row = tile * NT + tid;
    const bool live = row < nrows;
    const long long rb = live ? row_base_of(row, N) : 0;
    cd q[N];
    for (int s = 0; s < n_active; ++s) {
      cd* sAb = sR + (1 + (s & 1)) * N * N;  // double-buffered coefficient pair
      cd* sBb = sR + (3 + (s & 1)) * N * N;
      for (int e = tid; e < N * N; e += NT) {
        sAb[e] = Amats[static_cast<size_t>(s) * N * N + e];
        sBb[e] = Bmats[static_cast<size_t>(s) * N * N + e];
      }
      __syncthreads();
      if (live) {
        if (s == 0) {
          row_load<N>(Q, rb, q);
          if (do_backsub) {
            row_backsub<N>(q, sR);
            row_store<N>(Q, rb, q);
          }
        }
        cd p[N];
        row_load<N>(fp.P[s], rb, p);
        {
          cd x[N], dx[N];
          row_load<N>(fp.X[s], rb, x);
#pragma unroll
          for (int j = 0; j < N; ++j) dx[j] = czero();
          row_mm_acc_il<N>(dx, p, sAb);
#pragma unroll
          for (int j = 0; j < N; ++j) x[j] = cadd(x[j], dx[j]);  // product first, one addition into X (fields.hpp:74)
          row_store<N>(fp.X[s], rb, x);
        }
        {
          cd pn[N];
#pragma unroll
          for (int j = 0; j < N; ++j) pn[j] = czero();
          row_mm_acc_il<N>(pn, p, sBb);
#pragma unroll
          for (int j = 0; j < N; ++j) pn[j] = cadd(pn[j], q[j]);  // tmp = P*L ; tmp += Q*1.0 (fields.hpp:85-86)
          row_store<N>(fp.P[s], rb, pn);
        }
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------
// K4 (pipelined): the same multishift update as a warp-specialised bulk-TMA pipeline.
//
//   1 producer warp : per-site cp.async.bulk loads of {P_s tile, X_s tile} plus the
//                     coefficient pair {A_s, B_s} (and {Q tile, R'} once per tile) into a
//                     2-stage shared-memory ring, and the bulk stores of the updated
//                     tiles back to HBM.  Sites land with a padded pitch so that the
//                     compute warps' row accesses are bank-conflict free.
//   NCW compute warps: a site (3 colour rows) is owned by NSPLIT lanes of one warp, lane h
//                     producing JC = N/NSPLIT output columns for all three rows, so every
//                     coefficient word fetched from shared memory feeds 3 complex MACs
//                     (12 DFMA) and every P word feeds JC of them.  Rows are updated in
//                     place in the stage buffer.
// Stage hand-off is by mbarriers only (full[]: TMA -> compute, computed[]: compute ->
// producer).  The one intra-warp hazard (a lane overwriting P columns its partner lanes
// still read) is closed by a __syncwarp() between the last read and the first write.
// Coefficients arrive in the interleaved layout of shift_mat_index(); Rrecip is rho with
// its diagonal already inverted.  Bit 1 of do_backsub skips the arithmetic (diagnostic:
// measures the bare load/store ring).
// ------------------------------------------------------------------------------------
template <int N, int TS>
struct ShiftGeom {
  static constexpr int NSPLIT = shift_nsplit(N);
  static constexpr int JC = N / NSPLIT;
  static constexpr int SPW = 32 / NSPLIT;  // sites per compute warp
  static_assert(TS % SPW == 0, "tile must fill whole warps");
  static constexpr int NCW = TS / SPW;     // compute warps
  static constexpr int NT = (NCW + 1) * 32;
  static constexpr int SITE = 3 * N;       // complex per site in HBM
  // Sites are staged in PAIRS (one bulk copy of 2 sites) with one 16-byte word of padding
  // per pair when the pair would otherwise start on the same banks as its neighbour:
  // the 8 sites a warp touches per instruction then start on 8 distinct 4-bank groups.
  static_assert(TS % 2 == 0, "tile holds whole site pairs");
  static constexpr int PAIR = 2 * SITE + (((2 * SITE) % 2 == 0) ? 1 : 0);
  static constexpr int TILE = (TS / 2) * PAIR;  // complex per staged field tile
  // tensor-map copies need 128-byte aligned destinations: stages are whole multiples of 8 complex
  static constexpr int STAGE_ELEMS = (2 * TILE + 2 * N * N + 7) / 8 * 8;
  static constexpr int NSTAGE = 2;
  static constexpr size_t SMEM_BYTES = sizeof(cd) * NSTAGE * STAGE_ELEMS + 64;
  // registers are allocated per 4 warps: two resident CTAs of <= 8 warp slots each may use 128 per
  // thread, one CTA of <= 8 warps 232, one of 9..12 warps 168
  static constexpr int MAXREG = (2 * SMEM_BYTES <= 220 * 1024 && NT <= 256) ? 128 : (NT <= 256 ? 232 : 168);
};

// tensor maps of the fields one multishift launch touches (kernel parameter, __grid_constant__)
struct ShiftMaps {
  CUtensorMap Q;
  CUtensorMap P[kMaxShifts];
  CUtensorMap X[kMaxShifts];
};

template <int N, int TS>
__global__ void __maxnreg__((ShiftGeom<N, TS>::MAXREG))
shift_pipe_kernel(const __grid_constant__ ShiftMaps maps, const cd* __restrict__ Rrecip,
                  const cd* __restrict__ Amats, const cd* __restrict__ Bmats, long long V, int do_backsub,
                  int n_active_fixed, const Ctrl* __restrict__ ctrl) {
  using Geo = ShiftGeom<N, TS>;
  constexpr int NSPLIT = Geo::NSPLIT, JC = Geo::JC, SPW = Geo::SPW, NCW = Geo::NCW, SITE = Geo::SITE;
  constexpr int PAIR = Geo::PAIR, TILE = Geo::TILE, STAGE = Geo::STAGE_ELEMS;
  if (ctrl != nullptr && ctrl->done) return;
  const int n_active = (n_active_fixed > 0) ? n_active_fixed : ctrl->n_unconv;
  const bool backsub = (do_backsub & 1) != 0, arith = (do_backsub & 2) == 0;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  cd* sbuf = reinterpret_cast<cd*>(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(sbuf + Geo::NSTAGE * STAGE);
  uint64_t* computed = full + Geo::NSTAGE;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < Geo::NSTAGE; ++s) {
      mbar_init(full + s, 1);
      mbar_init(computed + s, NCW * 32);
    }
    mbar_fence_init();
  }
  __syncthreads();

  const long long ntiles = (V + TS - 1) / TS;
  constexpr uint32_t MAT_BYTES = N * N * sizeof(cd);
  constexpr uint32_t TILE_BYTES = TILE * sizeof(cd);

  if (warp == NCW) {
    // ===================== producer: one lane, tensor copies of whole padded tiles =====================
    // A field tile is a box of TS/2 site pairs, one complex wider than a pair, of the
    // [pair][2 * 3N] view of the field: pairs land PAIR apart in shared memory, the surplus
    // element is out of bounds (zero on load, not written on store), rows past the end of the
    // field likewise.
    if (lane != 0) return;
    long long it = 0;
    int rp_a = 0, rp_b = 0, rs_a = 0, rs_b = 0;  // descriptors of the two items in flight
    auto store_item = [&](int st) {
      const cd* buf = sbuf + st * STAGE;
      const int s = st ? rs_b : rs_a, pr = st ? rp_b : rp_a;
      if (s < 0) {
        if (backsub) tma_store_2d(&maps.Q, 0, pr, buf);
      } else {
        tma_store_2d(&maps.P[s], 0, pr, buf);
        tma_store_2d(&maps.X[s], 0, pr, buf + TILE);
      }
      bulk_commit();
    };
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int pair0 = static_cast<int>(tile * (TS / 2));
      for (int s = -1; s < n_active; ++s, ++it) {
        const int st = static_cast<int>(it & 1);
        const uint32_t use = static_cast<uint32_t>(it >> 1);
        if (it >= 2) {
          mbar_wait(computed + st, (use - 1) & 1u);  // item it-2 has been computed in place
          store_item(st);
          bulk_wait_read0();  // the stores have drained the buffer
        }
        if (st) {
          rp_b = pair0;
          rs_b = s;
        } else {
          rp_a = pair0;
          rs_a = s;
        }
        cd* buf = sbuf + st * STAGE;
        if (s < 0) {
          mbar_arrive_expect_tx(full + st, TILE_BYTES + (backsub ? MAT_BYTES : 0u));
          if (backsub) bulk_g2s(buf + 2 * TILE, Rrecip, MAT_BYTES, full + st);
          tma_load_2d(buf, &maps.Q, 0, pair0, full + st);
        } else {
          mbar_arrive_expect_tx(full + st, 2 * TILE_BYTES + 2 * MAT_BYTES);
          bulk_g2s(buf + 2 * TILE, Amats + static_cast<size_t>(s) * N * N, MAT_BYTES, full + st);
          bulk_g2s(buf + 2 * TILE + N * N, Bmats + static_cast<size_t>(s) * N * N, MAT_BYTES, full + st);
          tma_load_2d(buf, &maps.P[s], 0, pair0, full + st);
          tma_load_2d(buf + TILE, &maps.X[s], 0, pair0, full + st);
        }
      }
    }
    // drain the last (up to two) items
    for (long long k = (it >= 2 ? it - 2 : 0); k < it; ++k) {
      const int st = static_cast<int>(k & 1);
      mbar_wait(computed + st, static_cast<uint32_t>(k >> 1) & 1u);
      store_item(st);
    }
    bulk_wait0();
    return;
  }

  // ===================== compute warps =====================
  const int h = lane % NSPLIT;                 // column group of this lane
  const int lsite = warp * SPW + lane / NSPLIT;  // site within the tile
  const int sbase = (lsite >> 1) * PAIR + (lsite & 1) * SITE;
  long long it = 0;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long x0 = tile * TS;
    const int ns = static_cast<int>(min(static_cast<long long>(TS), V - x0));
    const bool live = lsite < ns;
    cd qh[3][JC];
    {  // ---- Q item: lanes h = 0,1,2 back-substitute colour row h, then everybody picks its columns ----
      const int st = static_cast<int>(it & 1);
      mbar_wait(full + st, static_cast<uint32_t>(it >> 1) & 1u);
      cd* buf = sbuf + st * STAGE;
      if (backsub && arith) {
        for (int c = h; c < 3; c += NSPLIT) {
          if (live) {
            cd q[N];
#pragma unroll
            for (int k = 0; k < N; ++k) q[k] = buf[sbase + 3 * k + c];
            row_backsub<N>(q, buf + 2 * TILE);
#pragma unroll
            for (int k = 0; k < N; ++k) buf[sbase + 3 * k + c] = q[k];
          }
        }
      }
      __syncwarp();
      if (live) {
#pragma unroll
        for (int j = 0; j < JC; ++j)
#pragma unroll
          for (int c = 0; c < 3; ++c) qh[c][j] = buf[sbase + 3 * (h * JC + j) + c];
      }
      fence_proxy_async();
      mbar_arrive(computed + st);
      ++it;
    }
    for (int s = 0; s < n_active; ++s, ++it) {
      const int st = static_cast<int>(it & 1);
      mbar_wait(full + st, static_cast<uint32_t>(it >> 1) & 1u);
      cd* sP = sbuf + st * STAGE + sbase;
      cd* sX = sP + TILE;
      const cd* sA = sbuf + st * STAGE + 2 * TILE;
      const cd* sB = sA + N * N;
      cd acc[3][JC];
      if (live && arith) {
        // ---- X_s += P_s A_s ----
#pragma unroll
        for (int j = 0; j < JC; ++j)
#pragma unroll
          for (int c = 0; c < 3; ++c) acc[c][j] = czero();
#pragma unroll
        for (int k = 0; k < N; ++k) {
          const cd p0 = sP[3 * k], p1 = sP[3 * k + 1], p2 = sP[3 * k + 2];
#pragma unroll
          for (int j = 0; j < JC; ++j) {
            const cd m = lds_cd(sA + (k * JC + j) * NSPLIT + h);
            cmac(acc[0][j], p0, m);
            cmac(acc[1][j], p1, m);
            cmac(acc[2][j], p2, m);
          }
        }
        // the product first, ONE addition into X last (as the reference: tmp = rhs * M; this += tmp,
        // fields.hpp:74): late in a solve |X| >> |increment|, and accumulating the N terms straight into
        // X would round N times at the size of X -- measured as a 5x higher true-residual floor
#pragma unroll
        for (int j = 0; j < JC; ++j)
#pragma unroll
          for (int c = 0; c < 3; ++c) sX[3 * (h * JC + j) + c] = cadd(sX[3 * (h * JC + j) + c], acc[c][j]);
        // ---- P_s <- P_s B_s + Q ----
#pragma unroll
        for (int j = 0; j < JC; ++j)
#pragma unroll
          for (int c = 0; c < 3; ++c) acc[c][j] = czero();
#pragma unroll
        for (int k = 0; k < N; ++k) {
          const cd p0 = sP[3 * k], p1 = sP[3 * k + 1], p2 = sP[3 * k + 2];
#pragma unroll
          for (int j = 0; j < JC; ++j) {
            const cd m = lds_cd(sB + (k * JC + j) * NSPLIT + h);
            cmac(acc[0][j], p0, m);
            cmac(acc[1][j], p1, m);
            cmac(acc[2][j], p2, m);
          }
        }
      }
      __syncwarp();  // partner lanes have finished reading the P rows before they are overwritten
      if (live && arith) {
#pragma unroll
        for (int j = 0; j < JC; ++j)
#pragma unroll
          for (int c = 0; c < 3; ++c)
            sP[3 * (h * JC + j) + c] = cadd(acc[c][j], qh[c][j]);  // tmp = P*L ; tmp += Q (fields.hpp:85-86)
      }
      fence_proxy_async();
      mbar_arrive(computed + st);
    }
  }
}

// Counter-based uniform numbers in [-1, 1) for inputs generated in place (the reference fills
// its links and sources with Eigen's setRandom, i.e. -1 + 2*rand()/RAND_MAX: inc/dirac_op.hpp:24-32,
// benchmark.cpp:60-62 -- the same distribution, not the same stream: libc's rand() is sequential).
// Double number `first + i` of the GLOBAL array gets mix(seed, stream, first + i), so a field does
// not depend on how many ranks it is split over.  mix = the SplitMix64 finaliser.
__host__ __device__ inline double counter_uniform(unsigned long long seed, unsigned long long stream,
                                                  unsigned long long index) {
  unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (index + 1) + 0xD1B54A32D192ED03ull * stream;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return static_cast<double>(z >> 11) * 0x1.0p-52 - 1.0;  // exact: a multiple of 2^-52
}
static __global__ void fill_uniform_kernel(double* __restrict__ dst, long long n, unsigned long long first,
                                           unsigned long long seed, unsigned long long stream) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
    dst[i] = counter_uniform(seed, stream, first + static_cast<unsigned long long>(i));
}

// Halo refresh for one rank: slots -H..-1 and V..V+H-1 <- periodic images (any V >= 1).
// `site` = complex numbers per site (3N for fields, 9 or 36 for links); H = 2 for the 1-D
// chain, one x3-slice for the 4-D operator.
static __global__ void halo_wrap_kernel(cd* __restrict__ f, long long V, int site, long long H,
                                        const Ctrl* __restrict__ ctrl) {
  pdl_wait();
  pdl_trigger();
  if (ctrl != nullptr && ctrl->done) return;
  const long long n = 2 * H * site;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long slot = i / site;
    const int e = static_cast<int>(i - slot * site);
    const long long h = (slot < H) ? (slot - H) : (V + slot - H);
    long long src = h % V;
    if (src < 0) src += V;
    f[h * site + e] = f[src * site + e];
  }
}

// Halo exchange of a slab decomposition over peer memory.  Sequence number of the exchange
// that follows iteration i: seq_base + i (i = 0: the refresh before the loop).
static __global__ void halo_push_kernel(const cd* __restrict__ f, long long V, int site, HaloPeers hp,
                                        const Ctrl* __restrict__ ctrl) {
  pdl_wait();
  pdl_trigger();
  if (ctrl->done) return;
  const unsigned long long k = ctrl->seq_base + static_cast<unsigned long long>(ctrl->iter);
  const int n = 2 * site;
  cd* to_left = hp.hi_of_left + (k & 1ull) * n;    // my sites 0,1    -> left neighbour's slots V, V+1
  cd* to_right = hp.lo_of_right + (k & 1ull) * n;  // my sites V-2,V-1 -> right neighbour's slots -2,-1
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    to_left[i] = f[i];
    to_right[i] = f[(V - 2) * site + i];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    st_release_sys(hp.seq_hi_of_left, k);
    st_release_sys(hp.seq_lo_of_right, k);
  }
}
static __global__ void halo_wait_unpack_kernel(cd* __restrict__ f, long long V, int site, HaloPeers hp,
                                               Ctrl* __restrict__ ctrl) {
  pdl_wait();
  pdl_trigger();
  if (ctrl->done) return;
  const unsigned long long k = ctrl->seq_base + static_cast<unsigned long long>(ctrl->iter);
  __shared__ int timed_out;
  if (threadIdx.x == 0) timed_out = 0;
  __syncthreads();
  if (threadIdx.x < 2) {
    const unsigned long long* w = threadIdx.x ? hp.my_seq_hi : hp.my_seq_lo;
    const long long t0 = clock64();
    while (ld_acquire_sys(w) < k)
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
    return;
  }
  const int n = 2 * site;
  const cd* lo = hp.my_lo + (k & 1ull) * n;
  const cd* hi = hp.my_hi + (k & 1ull) * n;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    f[i - n] = __ldcg(reinterpret_cast<const double2*>(lo + i));                  // slots -2, -1
    f[V * site + i] = __ldcg(reinterpret_cast<const double2*>(hi + i));          // slots V, V+1
  }
}

// pack the two boundary slabs (first 2 / last 2 sites) for a neighbour exchange
static __global__ void halo_pack_kernel(const cd* __restrict__ f, long long V, int site, cd* __restrict__ send_lo,
                                 cd* __restrict__ send_hi) {
  const int n = 2 * site;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    send_lo[i] = f[i];                    // sites 0,1   -> left neighbour's slots V, V+1
    send_hi[i] = f[(V - 2) * site + i];   // sites V-2,V-1 -> right neighbour's slots -2,-1
  }
}

}  // namespace bcg
