// 4-D extension of the block Dirac apply.  NOT part of the reference: its
// operator is a 1-D chain (inc/dirac_op.hpp:13-21).  This is the same construction with four
// directions on an L0 x L1 x L2 x L3 periodic lattice, x = x0 + L0 (x1 + L1 (x2 + L2 x3)), four
// random 3x3 links per site (U[x][mu], column-major):
//     D v[x] = 1/2 sum_mu ( U_mu[x] v[x + mu] - U_mu[x - mu]^dag v[x - mu] )      (anti-Hermitian)
//     T      = (m^2 + sigma) P - D (D P)                                          (Hermitian pos. def.)
// Checked against a CPU restatement of the same formula only (tests/): parity is UNPINNED by
// the reference, which has no such operator.
//
// First version: one launch per sweep, the intermediate D P goes through HBM, neighbours are
// gathered straight from global memory (L1 / L2 serve the 8-fold reuse).  Directions 0..2 wrap
// inside the local volume; direction 3 (the slab-decomposed "t") never wraps: fields carry a
// halo of one x3-slice (L0*L1*L2 sites) on either side, filled by a wrap copy or by the
// neighbour exchange.  Thread = (site, group of R right-hand sides).
#pragma once
#include "common.cuh"
#include "field_kernels.cuh"

namespace bcg {

struct Lattice4 {
  int L0, L1, L2, L3;       // local extents (L3 = this rank's slab thickness)
  long long s2, s3;         // strides of directions 2 and 3: L0*L1, L0*L1*L2
};

// SECOND = false:  out = D in                         (sweep 1)
// SECOND = true :  out = (m2 + sigma) p0 - D in       (sweep 2; in = D p0)
template <int N, int R, bool SECOND>
__global__ void __launch_bounds__(128)
dirac4_kernel(const cd* __restrict__ in, const cd* __restrict__ p0, cd* __restrict__ out,
              const cd* __restrict__ U, Lattice4 lat, long long x_begin, long long x_end, double m2, double sigma,
              const Ctrl* __restrict__ ctrl) {
  // sites [x_begin, x_end): the whole local volume, or a range of x3-slices when the halo
  // exchange of the boundary slices is overlapped with the interior
  constexpr int G = N / R, SITE = 3 * N;
  if (ctrl != nullptr && (ctrl->done | ctrl->stop)) return;
  const long long item = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (item >= (x_end - x_begin) * G) return;
  const long long x = x_begin + item / G;
  const int g = static_cast<int>(item % G);
  const int col0 = g * R * 3;
  const int x0 = static_cast<int>(x % lat.L0);
  const long long q1 = x / lat.L0;
  const int x1 = static_cast<int>(q1 % lat.L1);
  const int x2 = static_cast<int>((q1 / lat.L1) % lat.L2);
  // neighbour offsets (in sites); direction 3 reads the halo slices, no wrap
  const long long fwd[4] = {(x0 + 1 == lat.L0) ? -(lat.L0 - 1) : 1,
                            (x1 + 1 == lat.L1) ? -static_cast<long long>(lat.L1 - 1) * lat.L0 : lat.L0,
                            (x2 + 1 == lat.L2) ? -static_cast<long long>(lat.L2 - 1) * lat.s2 : lat.s2, lat.s3};
  const long long bwd[4] = {(x0 == 0) ? (lat.L0 - 1) : -1,
                            (x1 == 0) ? static_cast<long long>(lat.L1 - 1) * lat.L0 : -static_cast<long long>(lat.L0),
                            (x2 == 0) ? static_cast<long long>(lat.L2 - 1) * lat.s2 : -lat.s2, -lat.s3};
  cd acc[R][3];
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) acc[r][c] = czero();
#pragma unroll
  for (int mu = 0; mu < 4; ++mu) {
    cd v[R][3];
    const long long xp = x + fwd[mu], xm = x + bwd[mu];
    load_cols<N, R>(in + xp * SITE + col0, v);
    apply_link<R>(U + (x * 4 + mu) * 9, v, acc);
    load_cols<N, R>(in + xm * SITE + col0, v);
    apply_link_dag_sub<R>(U + (xm * 4 + mu) * 9, v, acc);
  }
  cd* o = out + x * SITE + col0;
  if (!SECOND) {
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) o[r * 3 + c] = cscale(acc[r][c], 0.5);
  } else {
    cd v[R][3];
    load_cols<N, R>(p0 + x * SITE + col0, v);
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        cd t = cmake(fma(m2, v[r][c].x, -0.5 * acc[r][c].x), fma(m2, v[r][c].y, -0.5 * acc[r][c].y));
        t.x = fma(sigma, v[r][c].x, t.x);
        t.y = fma(sigma, v[r][c].y, t.y);
        o[r * 3 + c] = t;
      }
  }
}

}  // namespace bcg
