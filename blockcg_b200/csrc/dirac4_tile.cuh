// 4-D extension of the block Dirac apply, second generation (first one: dirac4d.cuh, which
// stays as the fallback and as the bit-for-bit cross-check).  NOT part of the reference (its operator is the 1-D
// chain of inc/dirac_op.hpp:13-21): parity UNPINNED, checked against the CPU restatement of the same formula only.
//
//     D v[x] = 1/2 sum_mu ( U_mu[x] v[x + mu] - U_mu[x - mu]^dag v[x - mu] ),     T = (m^2 + sigma) P - D (D P)
//
// The first version gathered every neighbour and every link straight from global memory, one thread per (site,
// right-hand side): 96 16-byte loads per thread and sweep, three quarters of them re-reading the same 3x3 link in all
// N threads of a site -- bound by L1 wavefronts (508 us per apply at 24^4, N = 12).
// Here nothing is gathered.  The unit of data movement is one x0-ROW of the lattice (L0 consecutive sites: L0 * 48N
// contiguous bytes of a field, L0 * 144 contiguous bytes of one direction's links, which are kept direction-major
// for that purpose: Ut[mu][site][3][3]).  A CTA owns a tile of b consecutive rows (b = 1 for the shapes the launcher
// picks).  For every one of the eight directions the neighbour sites of the tile are again whole rows (the same rows
// for +-x0, rows with x1, x2 or x3 stepped -- periodically, or into the x3 halo slices -- for the others), so a
// producer lane streams, stage by stage, [neighbour rows of the field | the rows of U_mu that direction needs] into a
// shared-memory ring with bulk copies, and the compute threads -- (site, R right-hand sides), 6R accumulators in
// registers for the whole tile -- consume the stages in the order mu = 0+, 0-, 1+, 1-, ... of the first version, so
// the results are bit-identical to it.  The result tile goes back through shared memory and one bulk store; the
// second sweep takes P itself as a last stage for the mass term and, where the shape has enough warps for it,
// accumulates the Gram P^dag T of the tile on the way (GramCta).
//
// What bounds it (ncu, profiles/r02_ncu_dirac4_tile.txt): nothing is saturated -- FP64 pipe 28 %, L2 30 %, DRAM 29 %
// with one 6-warp CTA per SM and a four-stage ring; the compute warps wait for row deliveries (7 x 48N + 8 x 144
// bytes per site and sweep come from L2: every field row is fetched by the seven tiles that touch it).  Independent
// pipelines hide that latency better than a deeper ring: 24^4, N = 12, per apply (two sweeps) 518 us with one 6-warp
// CTA per SM, 397 us with two, 331 us with four 3-warp CTAs (one row each, two stages) -- the shape the launcher
// therefore picks (ops.cuh: the smallest CTA whose tile holds one row).  Cutting the seven fetches per row needs tiles
// that are blocks in x1, x2 (and a sliding window in x3); with 13.8 KB per row that does not fit beside the ring.
#pragma once
#include "common.cuh"
#include "field_kernels.cuh"
#include "dirac4d.cuh"

namespace bcg {

// NCW_ compute warps + one producer warp: the launcher picks, among the shapes compiled in, the one whose
// tile (whole rows, one thread per site and R right-hand sides) leaves the fewest threads idle at this L0
template <int N, int NCW_, int NSMAX = 2>
struct Dirac4TileGeom {
  static constexpr int R = (N % 3 == 0) ? 3 : (N % 2 == 0) ? 2 : 1;  // right-hand sides per thread
  static constexpr int G = N / R;                                      // threads per site
  static constexpr int NCW = NCW_, NTC = NCW * 32, NT = NTC + 32;
  static constexpr int TSMAX = NTC / G;                                // sites per tile at most
  static constexpr int SITE = 3 * N;
  static constexpr int STAGE_ELEMS = TSMAX * (SITE + 9);              // [field rows | link rows]
  static constexpr int OUT_ELEMS = TSMAX * SITE;
  // the Gram of the tile on the way: only where it costs few registers (N <= 4: one 4x4 block per warp) -- at N = 8 the
  // fused sweep measured slower than the plain sweep + the stand-alone Gram kernel (384 vs 349 us at 24^4)
  static constexpr bool CAN_GRAM = NCW >= GramGeom<N>::NTASK && N <= 4;
  static constexpr int GRAM_ELEMS = CAN_GRAM ? (NCW / GramGeom<N>::NTASK) * N * N : 0;
  static constexpr size_t FIXED_BYTES = sizeof(cd) * (OUT_ELEMS + GRAM_ELEMS) + 128;
  static constexpr int NSFIT = (227 * 1024 - static_cast<int>(FIXED_BYTES)) / static_cast<int>(sizeof(cd) * STAGE_ELEMS);
  static constexpr int NSTAGE = NSFIT >= NSMAX ? NSMAX : NSFIT;  // NSMAX = 2: a ring short enough for two CTAs per SM
  static constexpr bool OK = NSTAGE >= 2 && TSMAX >= 1 && NT <= 1024;
  static constexpr size_t SMEM_BYTES = sizeof(cd) * (NSTAGE * STAGE_ELEMS) + FIXED_BYTES;
  static constexpr int CTAS_SMEM = static_cast<int>((227 * 1024) / (SMEM_BYTES + 1024));
  static constexpr int CTAS_THREADS = 1024 / NT;  // at most 1024 threads per SM: 64 registers each would be too few
  static constexpr int CTAS_PER_SM = CTAS_SMEM < 1 ? 1 : (CTAS_SMEM < CTAS_THREADS ? CTAS_SMEM : (CTAS_THREADS < 1 ? 1 : CTAS_THREADS));
};

// rows of the local lattice: r = x1 + L1 (x2 + L2 x3); site of (r, x0) = r L0 + x0
struct Rows4 {
  int L0, L1, L2, L3;
  long long site_stride_mu;  // sites between the link arrays of consecutive directions in Ut
};

// the row whose sites are the mu-neighbours (dir = +1 / -1) of the sites of row r; mu = 3 never wraps (halo slices)
__device__ __forceinline__ long long row_neighbour(const Rows4& g, long long r, int mu, int dir) {
  if (mu == 1) {
    const int x1 = static_cast<int>(r % g.L1);
    if (dir > 0) return (x1 + 1 == g.L1) ? r - (g.L1 - 1) : r + 1;
    return (x1 == 0) ? r + (g.L1 - 1) : r - 1;
  }
  if (mu == 2) {
    const int x2 = static_cast<int>((r / g.L1) % g.L2);
    const long long s = g.L1;
    if (dir > 0) return (x2 + 1 == g.L2) ? r - (g.L2 - 1) * s : r + s;
    return (x2 == 0) ? r + (g.L2 - 1) * s : r - s;
  }
  return r + dir * static_cast<long long>(g.L1) * g.L2;
}

// Ut[mu][site][9] <- U[site][mu][9], halo slices included (sites -H .. V+H-1)
static __global__ void links4_transpose_kernel(cd* __restrict__ Ut, const cd* __restrict__ U, long long first, long long n_sites,
                                               long long site_stride_mu) {
  const long long n = n_sites * 36;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long s = i / 36;
    const int e = static_cast<int>(i - s * 36), mu = e / 9;
    Ut[mu * site_stride_mu * 9 + (first + s) * 9 + (e - 9 * mu)] = U[(first + s) * 36 + e];
  }
}

// SECOND = false:  out = D in                         (sweep 1)
// SECOND = true :  out = (m2 + sigma) p0 - D in       (sweep 2; in = D p0) [+ partial Gram p0^dag out per CTA]
// rows [row_begin, row_end) of the local lattice; b rows per tile (b L0 G <= NTC).
template <int N, int NCW_, int NSMAX, bool SECOND, bool GRAM>
__global__ void __launch_bounds__((Dirac4TileGeom<N, NCW_, NSMAX>::NT), (Dirac4TileGeom<N, NCW_, NSMAX>::CTAS_PER_SM))
dirac4_tile_kernel(const cd* __restrict__ in, const cd* __restrict__ p0, cd* __restrict__ out, const cd* __restrict__ Ut,
                   Rows4 geo, long long row_begin, long long row_end, int b, double m2, double sigma,
                   cd* __restrict__ gpart, const Ctrl* __restrict__ ctrl) {
  using Geo = Dirac4TileGeom<N, NCW_, NSMAX>;
  constexpr int R = Geo::R, G = Geo::G, SITE = Geo::SITE, NS = Geo::NSTAGE, NTC = Geo::NTC;
  constexpr int NDIR = SECOND ? 8 : 7;  // stages per tile: own rows (both x0 directions), 1+, 1-, 2+, 2-, 3+, 3-, [p0]
  if (ctrl != nullptr && (ctrl->done | ctrl->stop)) return;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  cd* sStage = reinterpret_cast<cd*>(smem_raw);
  cd* sOut = sStage + NS * Geo::STAGE_ELEMS;
  cd* sG = sOut + Geo::OUT_ELEMS;
  uint64_t* full = reinterpret_cast<uint64_t*>(sG + Geo::GRAM_ELEMS);
  uint64_t* empty = full + NS;

  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, Geo::NCW);  // one arrival per compute warp (256 single arrivals would serialise on the barrier word)
    }
    mbar_fence_init();
  }
  __syncthreads();

  const int L0 = geo.L0;
  const long long nrows = row_end - row_begin;
  const long long ntiles = (nrows + b - 1) / b;
  const uint32_t row_field_bytes = static_cast<uint32_t>(L0 * SITE * sizeof(cd));
  const uint32_t row_link_bytes = static_cast<uint32_t>(L0 * 9 * sizeof(cd));
  const long long mu_stride = geo.site_stride_mu * 9;

  if (warp == Geo::NCW) {
    // ===================== producer: one lane issues every bulk copy =====================
    if (tid != NTC) return;
    long long it = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const long long r0 = row_begin + tile * b;
      const int nr = static_cast<int>(min(static_cast<long long>(b), row_end - r0));
      for (int k = 0; k < NDIR; ++k, ++it) {
        const int st = static_cast<int>(it % NS);
        if (it >= NS) mbar_wait(empty + st, static_cast<uint32_t>((it / NS - 1) & 1));
        cd* sF = sStage + static_cast<size_t>(st) * Geo::STAGE_ELEMS;
        cd* sU = sF + Geo::TSMAX * SITE;
        const bool is_p0 = SECOND && k == 7;
        const int mu = (k == 0) ? 0 : (k + 1) / 2, dir = (k & 1) ? +1 : -1;  // k = 1, 2 -> mu 1 (+, -) ; 3, 4 -> mu 2 ; 5, 6 -> mu 3
        mbar_arrive_expect_tx(full + st, static_cast<uint32_t>(nr) * (row_field_bytes + (is_p0 ? 0u : row_link_bytes)));
        for (int j = 0; j < nr; ++j) {
          const long long r = r0 + j;
          if (is_p0) {
            bulk_g2s(sF + static_cast<size_t>(j) * L0 * SITE, p0 + r * L0 * SITE, row_field_bytes, full + st);
            continue;
          }
          const long long rn = (k == 0) ? r : row_neighbour(geo, r, mu, dir);
          // forward hop: the link sits on the site itself; backward hop: on the neighbour (its adjoint is applied)
          const long long rl = (k == 0 || dir > 0) ? r : rn;
          bulk_g2s(sF + static_cast<size_t>(j) * L0 * SITE, in + rn * L0 * SITE, row_field_bytes, full + st);
          bulk_g2s(sU + static_cast<size_t>(j) * L0 * 9, Ut + mu * mu_stride + rl * L0 * 9, row_link_bytes, full + st);
        }
      }
    }
    return;
  }

  // ===================== compute warps: thread = (site of the tile, R right-hand sides) =====================
  const int ls = tid / G, g = tid - ls * G;
  const int jrow = ls / L0, x0 = ls - jrow * L0;
  const int lsp = jrow * L0 + ((x0 + 1 == L0) ? 0 : x0 + 1);   // x0 neighbours inside the row (periodic)
  const int lsm = jrow * L0 + ((x0 == 0) ? L0 - 1 : x0 - 1);
  const int col0 = g * R * 3;
  GramCta<N, GRAM ? Geo::NCW : GramGeom<N>::NTASK> gram;  // dummy geometry when !GRAM
  if (GRAM) gram.init();
  long long it = 0;
  bool store_pending = false;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long r0 = row_begin + tile * b;
    const int nr = static_cast<int>(min(static_cast<long long>(b), row_end - r0));
    const int ns = nr * L0;
    const bool live = ls < ns;
    cd acc[R][3];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) acc[r][c] = czero();
#pragma unroll 1
    for (int k = 0; k < 7; ++k, ++it) {
      const int st = static_cast<int>(it % NS);
      mbar_wait(full + st, static_cast<uint32_t>(it / NS) & 1u);
      const cd* sF = sStage + static_cast<size_t>(st) * Geo::STAGE_ELEMS;
      const cd* sU = sF + Geo::TSMAX * SITE;
      if (live) {
        cd v[R][3];
        if (k == 0) {
          load_cols<N, R>(sF + lsp * SITE + col0, v);
          apply_link<R>(sU + ls * 9, v, acc);
          load_cols<N, R>(sF + lsm * SITE + col0, v);
          apply_link_dag_sub<R>(sU + lsm * 9, v, acc);
        } else {
          load_cols<N, R>(sF + ls * SITE + col0, v);
          if (k & 1)
            apply_link<R>(sU + ls * 9, v, acc);
          else
            apply_link_dag_sub<R>(sU + ls * 9, v, acc);
        }
      }
      __syncwarp();
      if ((tid & 31) == 0) mbar_arrive(empty + st);
    }
    // ---- the result tile: through shared memory, one bulk store ----
    if (tid == 0 && store_pending) bulk_wait_read0();  // the previous tile's store has drained sOut
    named_bar_sync(1, NTC);
    int st_p0 = 0;
    if (SECOND) {
      st_p0 = static_cast<int>(it % NS);
      mbar_wait(full + st_p0, static_cast<uint32_t>(it / NS) & 1u);
      ++it;
    }
    const cd* sP = sStage + static_cast<size_t>(st_p0) * Geo::STAGE_ELEMS;
    if (live) {
      cd* o = sOut + ls * SITE + col0;
      if (!SECOND) {
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
          for (int c = 0; c < 3; ++c) o[r * 3 + c] = cscale(acc[r][c], 0.5);
      } else {
        cd v[R][3];
        load_cols<N, R>(sP + ls * SITE + col0, v);
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
    fence_proxy_async();
    named_bar_sync(1, NTC);
    if (tid == 0) {
      bulk_s2g(out + r0 * L0 * SITE, sOut, static_cast<uint32_t>(ns) * SITE * sizeof(cd));
      bulk_commit();
    }
    store_pending = true;
    if (GRAM) gram.accumulate(sP, sOut, 3 * ns);  // P^dag T over the rows of this tile
    if (SECOND) {
      __syncwarp();
      if ((tid & 31) == 0) mbar_arrive(empty + st_p0);
    }
  }
  if (tid == 0) bulk_wait0();
  if (GRAM) {
    // GramCta::finish synchronises with __syncthreads: here only the compute warps take part
    named_bar_sync(1, NTC);
    for (int e = tid; e < Geo::GRAM_ELEMS; e += NTC) sG[e] = czero();
    named_bar_sync(1, NTC);
    if (gram.active) gram_warp_store<N>(gram.acc, gram.ti, gram.tj, sG + gram.slice * N * N, false);
    named_bar_sync(1, NTC);
    constexpr int NSL = GramCta<N, GRAM ? Geo::NCW : GramGeom<N>::NTASK>::NSLICE;
    for (int e = tid; e < N * N; e += NTC) {
      cd s = sG[e];
#pragma unroll
      for (int sl = 1; sl < NSL; ++sl) s = cadd(s, sG[sl * N * N + e]);
      gpart[static_cast<size_t>(blockIdx.x) * N * N + e] = s;
    }
  }
}

}  // namespace bcg
