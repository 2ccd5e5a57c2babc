// K4, paired variant: the multishift update with the shifted systems served every SECOND
// iteration (reference: the per-shift updates of SBCGrQ, inc/block_solvers.hpp:175-177, and
// thinQR's back-substitution, inc/fields.hpp:125-136).
//
// X_s and P_s of the shifted systems (s >= 1) never feed back into the main recurrence: only
// P_0, Q and T do.  Their two field updates per iteration can therefore wait one iteration and be
// applied two at a time, from the Q of the previous and of the current iteration:
//     iteration 2k+1 ("first") : Q <- Q rho^-1, kept also as Qprev ; shift 0 as usual ;
//                                A_s, B_s and the active count stay behind in their parity slots
//     iteration 2k+2 ("second"): Q <- Q rho^-1 ; shift 0 as usual ; for every shifted system
//                                X += P A' ; P <- P B' + Qprev ; X += P A ; P <- P B + Q
// which reads and writes X_s, P_s once per two iterations: the launch pair moves
// (14 + 4 (S-1) + 1) F bytes instead of 2 (4 S + 2) F (46 F against 76 F at S = 9), with the
// arithmetic, its order and hence every bit of the result unchanged.  A shift that retired
// between the two iterations gets the first update only; if the loop ends on a "first"
// iteration (`stop` is set by its B-step), or once only the unshifted system is left, the launch
// does the plain update.
// The phase is read from the device-side iteration counter, so every launch of a captured
// graph is the same node.
//
// Structure as shift_pipe_kernel (one producer lane issuing tensor copies into a two-stage ring,
// four compute warps, stages updated in place and stored back); the previous Q tile passes
// through the ring as an item of its own and every thread keeps its nine words of it in a
// private shared-memory column.
#pragma once
#include "common.cuh"
#include "field_kernels.cuh"

namespace bcg {

template <int N, int TS>
struct ShiftPairGeom {
  using Base = ShiftGeom<N, TS>;
  static constexpr int NCT = Base::NCW * 32;                                   // compute threads
  static constexpr int STAGE_ELEMS = (2 * Base::TILE + 4 * N * N + 7) / 8 * 8; // P, X tiles + (A', B', A, B)
  static constexpr int SCRATCH = 3 * Base::JC * NCT;                           // Qprev words, [word][thread]
  static constexpr size_t SMEM_BYTES = sizeof(cd) * (Base::NSTAGE * STAGE_ELEMS + SCRATCH) + 64;
  static constexpr int MAXREG =
      (2 * (SMEM_BYTES + 1024) <= 227 * 1024 && Base::NT <= 256) ? 128 : (Base::NT <= 256 ? 232 : 168);
};

struct ShiftPairMaps {
  CUtensorMap Q, Qprev;
  CUtensorMap P[kMaxShifts];
  CUtensorMap X[kMaxShifts];
};

// What the k-th item of a tile is.  mode 0: plain (Q, then every active shift); 1: first of a
// pair (Q kept, shift 0); 2: second of a pair (Q, Qprev, shift 0, shifted systems).
struct PairPlan {
  int mode, n_items, n2;
  __device__ __forceinline__ void item(int k, int& kind, int& s) const {
    if (k == 0) {
      kind = (mode == 1) ? KQ_KEEP : KQ;
      s = -1;
    } else if (mode != 2) {
      kind = KCUR;
      s = k - 1;
    } else if (k == 1) {
      kind = KQPREV;
      s = -1;
    } else {
      s = k - 2;
      kind = (s == 0) ? KCUR : (s < n2 ? KBOTH : KPREV);
    }
  }
};

template <int N, int TS>
__global__ void __maxnreg__((ShiftPairGeom<N, TS>::MAXREG))
shift_pair_kernel(const __grid_constant__ ShiftPairMaps maps, const cd* __restrict__ Rrecip,
                  const cd* __restrict__ Aodd, const cd* __restrict__ Bodd, const cd* __restrict__ Aeven,
                  const cd* __restrict__ Beven, long long V, const Ctrl* __restrict__ ctrl) {
  // Aodd/Bodd: operand slots written in odd iterations, Aeven/Beven: in even ones ([shift][N*N] each)
  using Geo = ShiftGeom<N, TS>;
  using PG = ShiftPairGeom<N, TS>;
  constexpr int NSPLIT = Geo::NSPLIT, JC = Geo::JC, SPW = Geo::SPW, NCW = Geo::NCW, SITE = Geo::SITE;
  constexpr int PAIR = Geo::PAIR, TILE = Geo::TILE, STAGE = PG::STAGE_ELEMS, NN = N * N;
  if (ctrl->done) return;
  const int iter = ctrl->iter;
  const bool odd = (iter & 1) != 0;
  PairPlan plan;
  {
    // the alternating schedule of build_shift_items(), in closed form
    const int n2 = ctrl->n_unconv, n1 = odd ? n2 : ctrl->n_act[1];
    plan.n2 = n2;
    if (odd && !ctrl->stop && n1 > 1) {
      plan.mode = 1;
      plan.n_items = 2;
    } else if (!odd && n1 > 1) {
      plan.mode = 2;
      plan.n_items = 2 + n1;
    } else {
      plan.mode = 0;
      plan.n_items = 1 + n2;
    }
  }
  const cd* Acur = odd ? Aodd : Aeven;
  const cd* Bcur = odd ? Bodd : Beven;
  const cd* Aprev = Aodd;  // only used in even iterations
  const cd* Bprev = Bodd;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  cd* sbuf = reinterpret_cast<cd*>(smem_raw);
  cd* scratch = sbuf + Geo::NSTAGE * STAGE;
  uint64_t* full = reinterpret_cast<uint64_t*>(scratch + PG::SCRATCH);
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
  constexpr uint32_t MAT_BYTES = NN * sizeof(cd);
  constexpr uint32_t TILE_BYTES = TILE * sizeof(cd);

  if (warp == NCW) {
    // ===================== producer =====================
    if (lane != 0) return;
    long long it = 0;
    int d_pair[2] = {0, 0}, d_kind[2] = {KQPREV, KQPREV}, d_s[2] = {-1, -1};  // items in flight
    auto store_item = [&](int st) {
      const cd* buf = sbuf + st * STAGE;
      const int kind = d_kind[st], s = d_s[st], pr = d_pair[st];
      if (kind == KQ || kind == KQ_KEEP) {
        tma_store_2d(&maps.Q, 0, pr, buf);
        if (kind == KQ_KEEP) tma_store_2d(&maps.Qprev, 0, pr, buf);
      } else if (kind != KQPREV) {
        tma_store_2d(&maps.P[s], 0, pr, buf);
        tma_store_2d(&maps.X[s], 0, pr, buf + TILE);
      }
      bulk_commit();
    };
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int pair0 = static_cast<int>(tile * (TS / 2));
      for (int k = 0; k < plan.n_items; ++k, ++it) {
        const int st = static_cast<int>(it & 1);
        const uint32_t use = static_cast<uint32_t>(it >> 1);
        if (it >= 2) {
          mbar_wait(computed + st, (use - 1) & 1u);  // item it-2 has been computed in place
          store_item(st);
          bulk_wait_read0();
        }
        int kind, s;
        plan.item(k, kind, s);
        d_pair[st] = pair0;
        d_kind[st] = kind;
        d_s[st] = s;
        cd* buf = sbuf + st * STAGE;
        cd* mats = buf + 2 * TILE;
        if (kind == KQ || kind == KQ_KEEP) {
          mbar_arrive_expect_tx(full + st, TILE_BYTES + MAT_BYTES);
          bulk_g2s(mats, Rrecip, MAT_BYTES, full + st);
          tma_load_2d(buf, &maps.Q, 0, pair0, full + st);
        } else if (kind == KQPREV) {
          mbar_arrive_expect_tx(full + st, TILE_BYTES);
          tma_load_2d(buf, &maps.Qprev, 0, pair0, full + st);
        } else {
          const size_t off = static_cast<size_t>(s) * NN;
          mbar_arrive_expect_tx(full + st, 2 * TILE_BYTES + (kind == KBOTH ? 4 : 2) * MAT_BYTES);
          const cd* a_first = (kind == KCUR) ? Acur : Aprev;
          const cd* b_first = (kind == KCUR) ? Bcur : Bprev;
          bulk_g2s(mats, a_first + off, MAT_BYTES, full + st);
          bulk_g2s(mats + NN, b_first + off, MAT_BYTES, full + st);
          if (kind == KBOTH) {
            bulk_g2s(mats + 2 * NN, Acur + off, MAT_BYTES, full + st);
            bulk_g2s(mats + 3 * NN, Bcur + off, MAT_BYTES, full + st);
          }
          tma_load_2d(buf, &maps.P[s], 0, pair0, full + st);
          tma_load_2d(buf + TILE, &maps.X[s], 0, pair0, full + st);
        }
      }
    }
    for (long long k = (it >= 2 ? it - 2 : 0); k < it; ++k) {  // drain the last (up to two) items
      const int st = static_cast<int>(k & 1);
      mbar_wait(computed + st, static_cast<uint32_t>(k >> 1) & 1u);
      store_item(st);
    }
    bulk_wait0();
    return;
  }

  // ===================== compute warps =====================
  const int h = lane % NSPLIT;                   // column group of this lane
  const int lsite = warp * SPW + lane / NSPLIT;  // site within the tile
  const int sbase = (lsite >> 1) * PAIR + (lsite & 1) * SITE;
  cd* myq = scratch + tid;                       // word w of this thread's Qprev at myq[w * NCT]
  long long it = 0;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long x0 = tile * TS;
    const int ns = static_cast<int>(min(static_cast<long long>(TS), V - x0));
    const bool live = lsite < ns;
    cd qh[3][JC];
    for (int k = 0; k < plan.n_items; ++k, ++it) {
      int kind, s;
      plan.item(k, kind, s);
      const int st = static_cast<int>(it & 1);
      mbar_wait(full + st, static_cast<uint32_t>(it >> 1) & 1u);
      cd* buf = sbuf + st * STAGE;
      if (kind == KQ || kind == KQ_KEEP) {
        // lanes h = 0,1,2 back-substitute colour row h, then everybody picks its columns
        for (int c = h; c < 3; c += NSPLIT) {
          if (live) {
            cd q[N];
#pragma unroll
            for (int kk = 0; kk < N; ++kk) q[kk] = buf[sbase + 3 * kk + c];
            row_backsub<N>(q, buf + 2 * TILE);
#pragma unroll
            for (int kk = 0; kk < N; ++kk) buf[sbase + 3 * kk + c] = q[kk];
          }
        }
        __syncwarp();
        if (live) {
#pragma unroll
          for (int j = 0; j < JC; ++j)
#pragma unroll
            for (int c = 0; c < 3; ++c) qh[c][j] = buf[sbase + 3 * (h * JC + j) + c];
        }
      } else if (kind == KQPREV) {
        if (live) {
#pragma unroll
          for (int j = 0; j < JC; ++j)
#pragma unroll
            for (int c = 0; c < 3; ++c) myq[(c * JC + j) * PG::NCT] = buf[sbase + 3 * (h * JC + j) + c];
        }
      } else {
        cd* sP = buf + sbase;
        cd* sX = sP + TILE;
        const int nup = (kind == KBOTH) ? 2 : 1;
#pragma unroll 1
        for (int u = 0; u < nup; ++u) {
          const cd* sA = buf + 2 * TILE + 2 * u * NN;
          const cd* sB = sA + NN;
          const bool from_prev = (kind == KPREV) || (kind == KBOTH && u == 0);
          cd acc[3][JC];
          if (live) {
            // ---- X_s += P_s A : the product first, one addition into X (fields.hpp:74) ----
#pragma unroll
            for (int j = 0; j < JC; ++j)
#pragma unroll
              for (int c = 0; c < 3; ++c) acc[c][j] = czero();
#pragma unroll
            for (int kk = 0; kk < N; ++kk) {
              const cd p0 = sP[3 * kk], p1 = sP[3 * kk + 1], p2 = sP[3 * kk + 2];
#pragma unroll
              for (int j = 0; j < JC; ++j) {
                const cd m = lds_cd(sA + (kk * JC + j) * NSPLIT + h);
                cmac(acc[0][j], p0, m);
                cmac(acc[1][j], p1, m);
                cmac(acc[2][j], p2, m);
              }
            }
#pragma unroll
            for (int j = 0; j < JC; ++j)
#pragma unroll
              for (int c = 0; c < 3; ++c) sX[3 * (h * JC + j) + c] = cadd(sX[3 * (h * JC + j) + c], acc[c][j]);
            // ---- P_s <- P_s B + Q ----
#pragma unroll
            for (int j = 0; j < JC; ++j)
#pragma unroll
              for (int c = 0; c < 3; ++c) acc[c][j] = czero();
#pragma unroll
            for (int kk = 0; kk < N; ++kk) {
              const cd p0 = sP[3 * kk], p1 = sP[3 * kk + 1], p2 = sP[3 * kk + 2];
#pragma unroll
              for (int j = 0; j < JC; ++j) {
                const cd m = lds_cd(sB + (kk * JC + j) * NSPLIT + h);
                cmac(acc[0][j], p0, m);
                cmac(acc[1][j], p1, m);
                cmac(acc[2][j], p2, m);
              }
            }
          }
          __syncwarp();  // partner lanes have finished reading the P rows before they are overwritten
          if (live) {
#pragma unroll
            for (int j = 0; j < JC; ++j)
#pragma unroll
              for (int c = 0; c < 3; ++c) {
                const cd q = from_prev ? myq[(c * JC + j) * PG::NCT] : qh[c][j];
                sP[3 * (h * JC + j) + c] = cadd(acc[c][j], q);  // tmp = P*L ; tmp += Q (fields.hpp:85-86)
              }
          }
          __syncwarp();  // the new rows are complete before the second update reads them
        }
      }
      fence_proxy_async();
      mbar_arrive(computed + st);
    }
  }
}

}  // namespace bcg
