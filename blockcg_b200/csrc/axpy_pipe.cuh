// K3 (second generation): Q += T*M with the Gram Q^dag Q of the updated rows fused in, as a
// warp-specialised tensor-map TMA pipeline (reference: block_fermion_field::add, inc/fields.hpp:70-77,
// followed by hermitian_dot of the result, inc/fields.hpp:103-122,142; call sites
// block_solvers.hpp:148-152 / :33-35).
//
// CTA = one SM, persistent, tiles of TS consecutive sites:
//   1 loader lane   : two tensor copies per tile (T tile, Q tile).  The fields are viewed as
//                     [site pair][2 * 3N complex] with a box one complex wider than a pair, so
//                     every pair lands in shared memory one 16-byte word further round the banks
//                     (the surplus element is out of bounds: zero on load, skipped on store).
//   NCW update warps: a site is owned by NSPLIT lanes, lane h producing JC = N/NSPLIT columns for
//                     the three colour rows; Q is updated in place in the stage buffer.
//   4 Gram warps    : a quarter of the lower triangle of Q^dag Q each (GramPart) over the
//                     finished tile; one partial N x N block per CTA at the end.
//   1 storer lane   : one tensor store per tile.
// Four stages; hand-off by mbarriers only.
#pragma once
#include "common.cuh"
#include "dirac_chain.cuh"

namespace bcg {

template <int N, int TS>
struct AxpyPipeGeom {
  static_assert(N % 2 == 0 && N >= 4, "pipeline variant needs an even N >= 4");
  static constexpr int NSPLIT = shift_nsplit(N);
  static constexpr int JC = N / NSPLIT;
  static constexpr int SPW = 32 / NSPLIT;            // sites per update warp
  static_assert(TS % SPW == 0 && TS % 2 == 0, "tile must fill whole warps and whole pairs");
  static constexpr int NCW = TS / SPW;               // update warps
  static constexpr int NGW = 4;
  static constexpr int NWARPS = NCW + NGW + 2;
  static constexpr int NT = NWARPS * 32;
  static constexpr int WARP_LOAD = NCW + NGW, WARP_STORE = NCW + NGW + 1;
  static constexpr int SITE = 3 * N;
  static constexpr int PAIR = 2 * SITE + ((1 - (2 * SITE) % 8 + 8) % 8);  // pair pitch, 1 (mod 8)
  static constexpr int TILE = (TS / 2) * PAIR;       // complex per staged field tile
  static constexpr int NSTAGE = 4;
  static constexpr int STAGE = 2 * TILE;             // T tile, Q tile
  static constexpr size_t BASE_BYTES = sizeof(cd) * (NSTAGE * STAGE + N * N) + 8 * 4 * NSTAGE + 16;
  // folded A-step (AlphaFold): two N x N work matrices + the slices of the rank-order sum of the peer blocks
  static constexpr int FOLD_ELEMS = 2 * N * N + 8 * (N * (N + 1) / 2);
  static constexpr bool FOLD_OK = BASE_BYTES + sizeof(cd) * FOLD_ELEMS + 1024 <= 227 * 1024 && N * N <= NT;
  static constexpr size_t SMEM_BYTES = BASE_BYTES + (FOLD_OK ? sizeof(cd) * FOLD_ELEMS : 0);
  static constexpr int ROWS = 3 * TS;
};

// alpha = (P0^dag T)^-1 in the prologue of the Q update, exactly as the A-step forms it (small_kernels.cuh:
// sm_reduce_gram, sm_inverse<false> -- the same sums in the same order, the same elimination entry by entry), one
// matrix entry per thread.  w: FOLD_ELEMS complex of scratch; the returned pointer (one of the two work matrices)
// holds alpha.  Every thread of the CTA must call it (CTA barriers inside).
template <int N>
__device__ __noinline__ const cd* fold_alpha(cd* w, const cd* __restrict__ gsrc, int nsrc, int step_threads) {
  constexpr int nn = N * N, E = N * (N + 1) / 2;
  const int tid = threadIdx.x, nthr = blockDim.x;
  cd* A = w;
  cd* W = w + nn;
  cd* slices = w + 2 * nn;
  const int i = tid % N, j = tid / N;  // this thread's entry (tid < nn)
  if (nsrc == 1) {  // already reduced by the stencil: lower triangle + conjugate mirror (fields.hpp:103-122)
    if (tid < nn) A[tid] = (i >= j) ? gsrc[tid] : cconj(gsrc[j + N * i]);
    __syncthreads();
  } else {
    // the blocks of the ranks summed in rank order, sliced as the A-step's own threads slice it (sm_reduce_gram)
    int nsl = step_threads / E;
    nsl = nsl < 1 ? 1 : (nsl > 8 ? 8 : nsl);
    for (int wk = tid; wk < E * nsl; wk += nthr) {
      const int t = wk % E, q = wk / E;
      int jj = 0, rest = t;
      while (rest >= N - jj) {
        rest -= N - jj;
        ++jj;
      }
      const int e = (jj + rest) + N * jj;
      double re = 0.0, im = 0.0;
      for (int p = q; p < nsrc; p += nsl) {
        const cd a = gsrc[static_cast<size_t>(p) * nn + e];
        re += a.x;
        im += a.y;
      }
      slices[wk] = cmake(re, im);
    }
    __syncthreads();
    for (int t = tid; t < E; t += nthr) {
      int jj = 0, rest = t;
      while (rest >= N - jj) {
        rest -= N - jj;
        ++jj;
      }
      const int ii = jj + rest;
      cd sum = slices[t];
      for (int q = 1; q < nsl; ++q) sum = cadd(sum, slices[q * E + t]);
      A[ii + N * jj] = sum;
      if (ii != jj) A[jj + N * ii] = cconj(sum);
    }
    __syncthreads();
  }
  cd* src = A;
  cd* dst = W;
#pragma unroll 1
  for (int k = 0; k < N; ++k) {  // Gauss-Jordan, no pivoting (Hermitian positive definite): sm_inverse<false>
    const cd pk = src[k + N * k];
    const double pn = __drcp_rn(cabs2(pk));
    const cd rp = cmake(pk.x * pn, -pk.y * pn);
    if (tid < nn) dst[tid] = gj_entry(src, N, i, j, k, k, rp);
    __syncthreads();
    cd* t = src;
    src = dst;
    dst = t;
  }
  return src;
}

// GRAM: 0 none, 1 four Gram warps with DFMA (GramPart), 2 the same warps on the FP64 tensor instruction (GramDmma)
template <int N, int TS, int GRAM>
__global__ void __launch_bounds__(AxpyPipeGeom<N, TS>::NT, 1)
axpy_pipe_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmQout,
                 const __grid_constant__ CUtensorMap tmT,
                 const cd* __restrict__ M, long long V, cd* __restrict__ gpart, const Ctrl* __restrict__ ctrl,
                 const GramPeers peers, int reverse, const AlphaFold fold) {
  // reverse != 0: tiles are visited from the high end of the field down.  The stencil that ran
  // just before wrote T from site 0 upwards, so the top ~100 MB of T are still in L2; and the
  // multishift update that runs next reads Q from site 0 upwards, i.e. the part written last.
  using Geo = AxpyPipeGeom<N, TS>;
  constexpr int NSPLIT = Geo::NSPLIT, JC = Geo::JC, SPW = Geo::SPW, NCW = Geo::NCW, SITE = Geo::SITE;
  constexpr int PAIR = Geo::PAIR, TILE = Geo::TILE, STAGE = Geo::STAGE, NS = Geo::NSTAGE;
  pdl_wait();
  pdl_trigger();
  // fold.on: M = -alpha is formed here, from the stencil's Gram, and the A-step runs beside this kernel -- so `done`
  // (which that A-step sets after the last iteration) may not be up yet: `stop` says the same one iteration earlier
  if (ctrl != nullptr && (ctrl->done | (fold.on ? ctrl->stop : 0))) return;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  cd* sbuf = reinterpret_cast<cd*>(smem_raw);
  cd* sM = sbuf + NS * STAGE;
  uint64_t* full = reinterpret_cast<uint64_t*>(sM + N * N);
  uint64_t* cdone = full + NS;
  uint64_t* gdone = cdone + NS;
  uint64_t* sdone = gdone + NS;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < NS; ++i) {
      mbar_init(full + i, 1);
      mbar_init(cdone + i, NCW);
      mbar_init(gdone + i, GRAM ? Geo::NGW : NCW);
      mbar_init(sdone + i, 1);
    }
    mbar_fence_init();
  }
  const long long ntiles = (V + TS - 1) / TS;
  const int nmine = static_cast<int>(blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0);
  auto tile_pair0 = [&](int i) {  // first site pair of this CTA's i-th tile
    const long long tidx = static_cast<long long>(blockIdx.x) + static_cast<long long>(i) * gridDim.x;
    return static_cast<int>((reverse ? ntiles - 1 - tidx : tidx) * (TS / 2));
  };
  constexpr uint32_t TILE_BYTES = TILE * sizeof(cd);
  int preloaded = 0;  // tiles the loader lane has requested before the folded A-step

  bool folded = false;
  if constexpr (Geo::FOLD_OK) {
    if (fold.on) {
      folded = true;
      __syncthreads();  // barriers initialised
      if (warp == Geo::WARP_LOAD && lane == 0) {  // the first tiles travel while alpha is being formed
        for (; preloaded < nmine && preloaded < NS; ++preloaded) {
          mbar_arrive_expect_tx(full + preloaded, 2 * TILE_BYTES);
          tma_load_2d(sbuf + preloaded * STAGE, &tmT, 0, tile_pair0(preloaded), full + preloaded);
          tma_load_2d(sbuf + preloaded * STAGE + TILE, &tmQ, 0, tile_pair0(preloaded), full + preloaded);
        }
      }
      cd* w = reinterpret_cast<cd*>(smem_raw + ((Geo::BASE_BYTES + 15) / 16) * 16);
      const cd* gsrc = fold.gsrc;
      int nsrc = fold.nsrc;
      if (fold.gw.nranks > 0) {  // slab decomposition: every rank's block of this iteration has to have landed
        __shared__ int timed_out;
        if (tid == 0) timed_out = 0;
        __syncthreads();
        const unsigned long long kseq = ctrl->seq_base + static_cast<unsigned long long>(ctrl->iter_b + 1);
        if (tid < fold.gw.nranks) {
          const long long t0 = clock64();
          while (ld_acquire_sys(fold.gw.seq + tid) < kseq)
            if (clock64() - t0 > kSpinTimeoutClocks) {  // the A-step beside this kernel reports it
              timed_out = 1;
              break;
            }
        }
        __syncthreads();
        if (timed_out) {
          if (warp == Geo::WARP_LOAD && lane == 0)
            for (int i = 0; i < preloaded; ++i) mbar_wait(full + i, 0u);  // no copy may outlive the CTA
          return;
        }
        gsrc = fold.gw.slots + static_cast<size_t>(kseq & 1ull) * fold.gw.nranks * (N * N);
        nsrc = fold.gw.nranks;
      }
      const cd* alpha = fold_alpha<N>(w, gsrc, nsrc, fold.step_threads);
      for (int e = tid; e < N * N; e += Geo::NT) {
        const int k = e % N, j = e / N;
        sM[(k * JC + (j % JC)) * NSPLIT + j / JC] = cmake(-alpha[e].x, -alpha[e].y);
      }
    }
  }
  if (!folded) {
    // coefficient matrix, column groups interleaved: element (k, j) at (k*JC + j%JC)*NSPLIT + j/JC
    for (int e = tid; e < N * N; e += Geo::NT) {
      const int k = e % N, j = e / N;
      sM[(k * JC + (j % JC)) * NSPLIT + j / JC] = M[e];
    }
  }
  __syncthreads();

  if (warp == Geo::WARP_LOAD) {
    if (lane != 0) return;
    for (int i = preloaded; i < nmine; ++i) {
      const int st = i % NS;
      if (i >= NS) {  // stage free: Gram and store of the tile that used it are done
        const uint32_t par = static_cast<uint32_t>((i / NS - 1) & 1);
        mbar_wait(gdone + st, par);
        mbar_wait(sdone + st, par);
      }
      mbar_arrive_expect_tx(full + st, 2 * TILE_BYTES);
      tma_load_2d(sbuf + st * STAGE, &tmT, 0, tile_pair0(i), full + st);
      tma_load_2d(sbuf + st * STAGE + TILE, &tmQ, 0, tile_pair0(i), full + st);
    }
    return;
  }

  if (warp == Geo::WARP_STORE) {
    if (lane != 0) return;
    for (int i = 0; i < nmine; ++i) {
      const int st = i % NS;
      mbar_wait(cdone + st, static_cast<uint32_t>((i / NS) & 1));
      tma_store_2d(&tmQout, 0, tile_pair0(i), sbuf + st * STAGE + TILE);  // in place, or into the other Q field (staggered schedule)
      bulk_commit();
      bulk_wait_read0();
      mbar_arrive(sdone + st);
    }
    bulk_wait0();
    return;
  }

  if (warp >= NCW) {
    // ===================== Gram warps =====================
    if (!GRAM) return;
    if constexpr (GRAM == 2 && N % 4 == 0) {
      // row quad = site pairs 4 gw .. 4 gw + 3 of the tile at one (site-in-pair sp, colour c); the four Gram
      // warps split the TS / 2 = 16 pairs, each takes all six (sp, c)
      static_assert(TS == 32, "the tensor-instruction Gram splits 16 site pairs over four warps");
      const int gw = warp - NCW, q = lane & 3, mm = lane >> 2;
      const int off = 2 * ((4 * gw + q) * PAIR) + 6 * (mm >> 1) + (mm & 1);
      GramDmma<N> gd;
      gd.init();
      for (int i = 0; i < nmine; ++i) {
        const int st = i % NS;
        mbar_wait(cdone + st, static_cast<uint32_t>((i / NS) & 1));
        const double* dQ = reinterpret_cast<const double*>(sbuf + st * STAGE + TILE) + off;
        const long long site = 2LL * tile_pair0(i) + 2 * (4 * gw + q);  // first site of this lane's pair
#pragma unroll
        for (int sc = 0; sc < 6; ++sc) {
          const int sp = sc / 3, c = sc - 3 * sp;
          gd.quad(dQ + 2 * (sp * SITE + c), dQ + 2 * (sp * SITE + c), site + sp < V);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(gdone + st);
      }
      gd.store(gpart + (static_cast<size_t>(kGramRawOff) + blockIdx.x) * N * N, sbuf, gw);
      gram_group_reduce<N>(gpart, gw, peers, ctrl, 1);
      return;
    }
    if constexpr (GRAM == 1) {
    // row = (colour c, site-in-pair sp, pair m), m fastest across lanes: 8 consecutive pairs start
    // on 8 different 16-byte bank groups
    auto gram_loop = [&](auto& part) {
      part.init();
      for (int i = 0; i < nmine; ++i) {
        const int st = i % NS;
        mbar_wait(cdone + st, static_cast<uint32_t>((i / NS) & 1));
        const cd* tQ = sbuf + st * STAGE + TILE;
        const long long site0 = 2LL * tile_pair0(i);
        // not unrolled: four different Gram loops plus the stencil loop must stay resident in the
        // instruction cache together (stall_no_instruction tripled when they did not)
#pragma unroll 1
        for (int it = 0; it < (Geo::ROWS + 31) / 32; ++it) {
          const int rr = lane + 32 * it;
          const int m = rr % (TS / 2), sp = (rr / (TS / 2)) % 2, c = rr / TS;
          if (rr < Geo::ROWS && site0 + 2 * m + sp < V) {
            const cd* row = tQ + m * PAIR + sp * SITE + c;
            part.row(row, row);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(gdone + st);
      }
      // scratch: the T tile of stage (warp - NCW); every update has finished with it by now
      part.store(gpart + (static_cast<size_t>(kGramRawOff) + blockIdx.x) * N * N, sbuf + (warp - NCW) * STAGE);
      gram_group_reduce<N>(gpart, warp - NCW, peers, ctrl, 1);
    };
    switch (warp - NCW) {
      case 0: { GramPart<N, 0> part; gram_loop(part); break; }
      case 1: { GramPart<N, 1> part; gram_loop(part); break; }
      case 2: { GramPart<N, 2> part; gram_loop(part); break; }
      default: { GramPart<N, 3> part; gram_loop(part); break; }
    }
    }
    return;
  }

  // ===================== update warps =====================
  const int h = lane % NSPLIT;
  const int lsite = warp * SPW + lane / NSPLIT;
  const int sbase = (lsite >> 1) * PAIR + (lsite & 1) * SITE;
  for (int i = 0; i < nmine; ++i) {
    const int st = i % NS;
    mbar_wait(full + st, static_cast<uint32_t>((i / NS) & 1));
    const cd* sT = sbuf + st * STAGE + sbase;
    cd* sQ = sbuf + st * STAGE + TILE + sbase;
    cd acc[3][JC];
#pragma unroll
    for (int j = 0; j < JC; ++j)
#pragma unroll
      for (int c = 0; c < 3; ++c) acc[c][j] = czero();
#pragma unroll
    for (int k = 0; k < N; ++k) {
      const cd p0 = sT[3 * k], p1 = sT[3 * k + 1], p2 = sT[3 * k + 2];
#pragma unroll
      for (int j = 0; j < JC; ++j) {
        const cd m = lds_cd(sM + (k * JC + j) * NSPLIT + h);
        cmac(acc[0][j], p0, m);
        cmac(acc[1][j], p1, m);
        cmac(acc[2][j], p2, m);
      }
    }
#pragma unroll
    for (int j = 0; j < JC; ++j)
#pragma unroll
      for (int c = 0; c < 3; ++c)  // product first, then one addition (reference: tmp = rhs * M; this += tmp)
        sQ[3 * (h * JC + j) + c] = cadd(sQ[3 * (h * JC + j) + c], acc[c][j]);
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
      mbar_arrive(cdone + st);
      if (!GRAM) mbar_arrive(gdone + st);
    }
  }
}

}  // namespace bcg
