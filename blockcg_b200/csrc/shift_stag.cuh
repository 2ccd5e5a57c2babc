// K4, staggered deferral of depth k with streamed operands (schedule 3, build_stag_items in common.cuh).
//
// Same arithmetic as shift_dmma_kernel (the multishift update on the FP64 tensor instruction, reference
// inc/block_solvers.hpp:145,152,158,175-177; inc/fields.hpp:70-90,125-136) -- and therefore the same bits --
// but a shifted system is visited once per k iterations and then receives its k pending updates, so X_s and
// P_s cross HBM once per k iterations instead of once per two.  What keeps a deeper deferral from fitting into
// shift_dmma_kernel is shared memory: k-1 earlier Q tiles per thread and 2k coefficient matrices per stage.
// Here nothing but the P and X tiles lives in the two pipeline stages; the operands of ONE update at a time
// are streamed behind them:
//   coefficient ring : three slots of (A_u, B_u) (or rho^-1 for the Q item), filled by the producer in the
//                      order the updates consume them (full / empty mbarriers per slot);
//   Q-history slot   : the Q tile of the update's own iteration when that is an earlier one (the newest Q stays
//                      in registers, as in shift_dmma_kernel); an update needs it only in its last step
//                      (P <- P B + Q), so one slot is enough: the producer refills it while the tensor
//                      instructions of the next update run.  For the second and later systems of a tile these
//                      reads hit L2.
// 108.5 KB per CTA: two CTAs per SM as before.  Two producer lanes (tiles; operands), four compute warps.
//
// Overlapped variant (BCG_OVERLAP=1): nothing in the main recurrence waits for the shifted systems, so the launch
// is split -- part 1 (Q <- Q rho^-1, system 0, halo) stays on the loop's stream, part 2 (the shifted systems, every
// Q read from the ring, state from a snapshot in the control block) goes to a second stream and runs beside the
// stencil, the Q update and above all the two coefficient kernels of the next iteration, whose latency chains
// otherwise leave the GPU idle (the limiter of a slab decomposition over many GPUs).
#pragma once
#include "common.cuh"
#include "field_kernels.cuh"
#include "shift_dmma.cuh"

namespace bcg {

template <int N, int TS>
struct ShiftStagGeom {
  using D = ShiftDmmaGeom<N, TS, 2>;
  static constexpr int NCW = D::NCW, NT = (D::NCW + 2) * 32, NCT = D::NCT, SITE = D::SITE, PAIR = D::PAIR, TILE = D::TILE;
  static constexpr int KS = D::KS, NTL = D::NTL;
  static constexpr int NSTAGE = 2;
  static constexpr int STAGE_ELEMS = 2 * TILE;               // P, X tiles (or the Q tile)
  static constexpr int NCOEF = 3;
  static constexpr int COEF_ELEMS = 2 * N * N;               // (A_u, B_u)
  static_assert((TILE * sizeof(cd)) % 128 == 0 && (COEF_ELEMS * sizeof(cd)) % 128 == 0, "tensor-copy destinations are 128-byte aligned");
  static constexpr int NBAR = 2 * NSTAGE + 2 * NCOEF + 2;
  static constexpr size_t SMEM_BYTES = sizeof(cd) * (NSTAGE * STAGE_ELEMS + NCOEF * COEF_ELEMS + TILE) + 8 * NBAR + 64;
  static constexpr bool TWO_CTAS = 2 * (SMEM_BYTES + 1024) <= 227 * 1024 && NT <= 256;
  static constexpr int MAXREG = TWO_CTAS ? 152 : (NT <= 256 ? 232 : 168);
};

struct ShiftStagMaps {
  CUtensorMap Q[kMaxDepth];  // the ring of Q fields: Q of iteration i lives in field i % depth
  CUtensorMap P[kMaxShifts];
  CUtensorMap X[kMaxShifts];
};
struct ShiftStagCoefs {      // operand sets: A[t], B[t] = [shift][N*N] written in the iterations with i % depth == t
  const cd* A[kMaxDepth];
  const cd* B[kMaxDepth];
};

// after the overlapped launch (same stream): slot `slot` has been served
static __global__ void bulk_mark_kernel(Ctrl* __restrict__ ctrl, int slot) {
  ctrl->bulk_served[slot] = ctrl->snap[slot].iter;
}

template <int N, int TS>
__global__ void __maxnreg__((ShiftStagGeom<N, TS>::MAXREG))
shift_stag_kernel(const __grid_constant__ ShiftStagMaps maps, const cd* __restrict__ Rrecip, const ShiftStagCoefs coefs,
                  long long V, const Ctrl* __restrict__ ctrl, int depth, int ring, int part, int slot,
                  cd* __restrict__ p0_halo, const HaloFold hf) {
  // depth = k, ring = number of Q fields / operand sets in use, part = 0 whole launch / 1 critical part (Q and
  // system 0) / 2 shifted systems only, from the snapshot in ctrl->snap[slot] (build_stag_items)
  using Geo = ShiftStagGeom<N, TS>;
  constexpr int NS = Geo::NSTAGE, NC = Geo::NCOEF;
  constexpr int NCW = Geo::NCW, SITE = Geo::SITE, PAIR = Geo::PAIR, TILE = Geo::TILE, STAGE = Geo::STAGE_ELEMS;
  constexpr int NN = N * N, KS = Geo::KS, NTL = Geo::NTL;
  pdl_wait();
  pdl_trigger();
  const bool bulk = part == 2;
  int iter;
  if (bulk) {
    // this launch runs beside the kernels of the next iterations: the loop's own control fields have moved on
    // (and `done` may be set although the last iteration's updates are still to come) -- the snapshot tells
    // which iteration to serve, bulk_served whether an earlier launch has served it already (loop over)
    iter = ctrl->snap[slot].iter;
    if (iter == ctrl->bulk_served[slot]) return;
  } else {
    if (ctrl->done) return;
    iter = ctrl->iter;
  }
  __shared__ StagItem s_items[kMaxShiftItems];
  __shared__ int s_n_items;
  if (threadIdx.x == 0) {
    int nr[kMaxDepth];
    if (bulk) {
      for (int t = 0; t < kMaxDepth; ++t) nr[t] = ctrl->snap[slot].n_ring[t];
      s_n_items = build_stag_items(depth, ring, part, iter, ctrl->snap[slot].stop, ctrl->snap[slot].n_now, nr, s_items, nullptr);
    } else {
      for (int t = 0; t < kMaxDepth; ++t) nr[t] = ctrl->n_act[t];
      s_n_items = build_stag_items(depth, ring, part, iter, ctrl->stop, ctrl->n_unconv, nr, s_items, nullptr);
    }
  }

  extern __shared__ __align__(128) unsigned char smem_raw[];
  cd* sbuf = reinterpret_cast<cd*>(smem_raw);
  cd* scoef = sbuf + NS * STAGE;
  cd* sqh = scoef + NC * Geo::COEF_ELEMS;
  uint64_t* full = reinterpret_cast<uint64_t*>(sqh + TILE);
  uint64_t* computed = full + NS;
  uint64_t* cfull = computed + NS;
  uint64_t* cempty = cfull + NC;
  uint64_t* qfull = cempty + NC;
  uint64_t* qempty = qfull + 1;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(full + s, 1);
      mbar_init(computed + s, NCW * 32);
    }
    for (int s = 0; s < NC; ++s) {
      mbar_init(cfull + s, 1);
      mbar_init(cempty + s, NCW * 32);
    }
    mbar_init(qfull, 1);
    mbar_init(qempty, NCW * 32);
    mbar_fence_init();
  }
  __syncthreads();

  const long long ntiles = (V + TS - 1) / TS;
  constexpr uint32_t MAT_BYTES = NN * sizeof(cd);
  constexpr uint32_t TILE_BYTES = TILE * sizeof(cd);
  const int n_items = s_n_items;
  const int cur = iter % ring;  // field / operand set of this iteration

  if (warp == NCW) {
    // ===================== tile producer: P / X / Q tiles in, updated tiles out =====================
    if (lane != 0) return;
    long long it = 0;
    int d_pair[NS], d_s[NS];  // items in flight (s == -1: the Q item)
    for (int i = 0; i < NS; ++i) {
      d_pair[i] = 0;
      d_s[i] = -2;
    }
    auto store_item = [&](int st) {
      const cd* buf = sbuf + st * STAGE;
      const int s = d_s[st], pr = d_pair[st];
      if (s == -1) {
        tma_store_2d(&maps.Q[cur], 0, pr, buf);
      } else if (s >= 0) {
        tma_store_2d(&maps.P[s], 0, pr, buf);
        tma_store_2d(&maps.X[s], 0, pr, buf + TILE);
      }
      bulk_commit();
    };
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int pair0 = static_cast<int>(tile * (TS / 2));
      for (int k = 0; k < n_items; ++k, ++it) {
        const int st = static_cast<int>(it % NS);
        if (it >= NS) {
          mbar_wait(computed + st, static_cast<uint32_t>((it / NS - 1) & 1));  // item it-NS has been computed in place
          store_item(st);
          bulk_wait_read0();
        }
        const StagItem item = s_items[k];
        d_pair[st] = pair0;
        cd* buf = sbuf + st * STAGE;
        if (item.kind == KQ) {
          d_s[st] = -1;
          mbar_arrive_expect_tx(full + st, TILE_BYTES);
          tma_load_2d(buf, &maps.Q[cur], 0, pair0, full + st);
        } else {
          d_s[st] = item.s;
          mbar_arrive_expect_tx(full + st, 2 * TILE_BYTES);
          tma_load_2d(buf, &maps.P[item.s], 0, pair0, full + st);
          tma_load_2d(buf + TILE, &maps.X[item.s], 0, pair0, full + st);
        }
      }
    }
    for (long long k = (it >= NS ? it - NS : 0); k < it; ++k) {  // drain the last (up to NS) items
      const int st = static_cast<int>(k % NS);
      mbar_wait(computed + st, static_cast<uint32_t>(k / NS) & 1u);
      store_item(st);
    }
    bulk_wait0();
    return;
  }
  if (warp == NCW + 1) {
    // ===================== operand producer: coefficient ring and earlier Q tiles, in consumption order ==========
    // (a lane of its own: a full coefficient ring or an occupied Q slot must not hold up the next tile's loads)
    if (lane != 0) return;
    long long cit = 0, qit = 0;
    auto coef_slot = [&]() {  // next slot of the coefficient ring, free again
      const int cs = static_cast<int>(cit % NC);
      if (cit >= NC) mbar_wait(cempty + cs, static_cast<uint32_t>((cit / NC - 1) & 1));
      ++cit;
      return cs;
    };
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int pair0 = static_cast<int>(tile * (TS / 2));
      for (int k = 0; k < n_items; ++k) {
        const StagItem item = s_items[k];
        if (item.kind == KQ) {
          const int cs = coef_slot();
          mbar_arrive_expect_tx(cfull + cs, MAT_BYTES);
          bulk_g2s(scoef + cs * Geo::COEF_ELEMS, Rrecip, MAT_BYTES, cfull + cs);
          continue;
        }
        const size_t off = static_cast<size_t>(item.s) * NN;
        for (int u = 0; u < item.m; ++u) {
          const int d = item.d_first - u;                      // the update of iteration iter - d
          const int set = ((iter - d) % ring + ring) % ring;
          const int cs = coef_slot();
          mbar_arrive_expect_tx(cfull + cs, 2 * MAT_BYTES);
          bulk_g2s(scoef + cs * Geo::COEF_ELEMS, coefs.A[set] + off, MAT_BYTES, cfull + cs);
          bulk_g2s(scoef + cs * Geo::COEF_ELEMS + NN, coefs.B[set] + off, MAT_BYTES, cfull + cs);
          if (d > 0 || bulk) {  // an earlier iteration's Q (overlapped launch: every Q), from its field of the ring
            if (qit >= 1) mbar_wait(qempty, static_cast<uint32_t>((qit - 1) & 1));
            ++qit;
            mbar_arrive_expect_tx(qfull, TILE_BYTES);
            tma_load_2d(sqh, &maps.Q[set], 0, pair0, qfull);
          }
        }
      }
    }
    return;
  }

  // ===================== compute warps (fragment mapping as shift_dmma_kernel) =====================
  const int m = lane >> 2, q = lane & 3;
  const int lsite = warp * Geo::D::SPW + m;
  const int sbase = (lsite >> 1) * PAIR + (lsite & 1) * SITE;
  long long it = 0, cit = 0, qit = 0;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long x0 = tile * TS;
    const int ns = static_cast<int>(min(static_cast<long long>(TS), V - x0));
    const bool live = lsite < ns;
    const long long xs = x0 + lsite;
    cd qf[3][NTL];  // this lane's C-fragment positions of the newest Q: (colour c, column 4 jt + q)
    for (int k = 0; k < n_items; ++k, ++it) {
      const StagItem item = s_items[k];
      const int st = static_cast<int>(it % NS);
      mbar_wait(full + st, static_cast<uint32_t>(it / NS) & 1u);
      cd* buf = sbuf + st * STAGE;
      if (item.kind == KQ) {
        const int cs = static_cast<int>(cit % NC);
        mbar_wait(cfull + cs, static_cast<uint32_t>(cit / NC) & 1u);
        ++cit;
        if (q < 3 && live) {
          cd qr[N];
#pragma unroll
          for (int kk = 0; kk < N; ++kk) qr[kk] = buf[sbase + 3 * kk + q];
          row_backsub<N>(qr, scoef + cs * Geo::COEF_ELEMS);
#pragma unroll
          for (int kk = 0; kk < N; ++kk) buf[sbase + 3 * kk + q] = qr[kk];
        }
        __syncwarp();
        mbar_arrive(cempty + cs);
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
          for (int jt = 0; jt < NTL; ++jt) qf[c][jt] = buf[sbase + 3 * (4 * jt + q) + c];
      } else {
        cd* sP = buf + sbase;
        cd* sX = sP + TILE;
        const double* dP = reinterpret_cast<const double*>(sP) + 6 * (q >> 1) + (q & 1);  // A fragment: k = q
        const bool halo_item = p0_halo != nullptr && item.s == 0;
        const bool push_item = hf.on && item.s == 0;
        const unsigned long long kseq = ctrl->seq_base + static_cast<unsigned long long>(iter);
        cd* to_left = hf.hp.hi_of_left + (kseq & 1ull) * (2 * SITE);
        cd* to_right = hf.hp.lo_of_right + (kseq & 1ull) * (2 * SITE);
#pragma unroll 1
        for (int u = 0; u < item.m; ++u) {
          const bool from_hist = bulk || item.d_first - u > 0;
          const int cs = static_cast<int>(cit % NC);
          mbar_wait(cfull + cs, static_cast<uint32_t>(cit / NC) & 1u);
          ++cit;
          const cd* sA = scoef + cs * Geo::COEF_ELEMS;
          const cd* sB = sA + NN;
          double f[KS][NTL];
          // ---- X_s += P_s A : the product first, one addition into X (fields.hpp:74) ----
          load_coef_frags<N>(sA, lane, f);
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            double acc[NTL][2];
#pragma unroll
            for (int jt = 0; jt < NTL; ++jt) acc[jt][0] = acc[jt][1] = 0.0;
#pragma unroll
            for (int js = 0; js < KS; ++js) {
              const double a = lds_f64(dP + 2 * (6 * js + c));
#pragma unroll
              for (int jt = 0; jt < NTL; ++jt) dmma_m8n8k4(acc[jt][0], acc[jt][1], a, f[js][jt]);
            }
            if (live) {
#pragma unroll
              for (int jt = 0; jt < NTL; ++jt) {
                cd* px = sX + 3 * (4 * jt + q) + c;
                const cd x = *px;
                *px = cmake(x.x + acc[jt][0], x.y + acc[jt][1]);
              }
            }
          }
          // ---- P_s <- P_s B + Q : tmp = P * L ; tmp += Q (fields.hpp:85-86) ----
          load_coef_frags<N>(sB, lane, f);
          mbar_arrive(cempty + cs);  // both matrices are in registers: the slot may be refilled
          if (from_hist) mbar_wait(qfull, static_cast<uint32_t>(qit & 1));
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            double acc[NTL][2];
#pragma unroll
            for (int jt = 0; jt < NTL; ++jt) acc[jt][0] = acc[jt][1] = 0.0;
#pragma unroll
            for (int js = 0; js < KS; ++js) {
              const double a = lds_f64(dP + 2 * (6 * js + c));
#pragma unroll
              for (int jt = 0; jt < NTL; ++jt) dmma_m8n8k4(acc[jt][0], acc[jt][1], a, f[js][jt]);
            }
            // every lane of the warp has read its fragments of colour row c before any lane gets here
            if (live) {
#pragma unroll
              for (int jt = 0; jt < NTL; ++jt) {
                const cd qq = from_hist ? sqh[sbase + 3 * (4 * jt + q) + c] : qf[c][jt];
                const cd pn = cmake(acc[jt][0] + qq.x, acc[jt][1] + qq.y);
                sP[3 * (4 * jt + q) + c] = pn;
                if (halo_item) {
                  if (xs < 2) p0_halo[(V + xs) * SITE + 3 * (4 * jt + q) + c] = pn;
                  if (xs >= V - 2) p0_halo[(xs - V) * SITE + 3 * (4 * jt + q) + c] = pn;
                }
                if (push_item) {
                  if (xs < 2) to_left[xs * SITE + 3 * (4 * jt + q) + c] = pn;
                  if (xs >= V - 2) to_right[(xs - (V - 2)) * SITE + 3 * (4 * jt + q) + c] = pn;
                }
              }
            }
          }
          if (from_hist) {
            mbar_arrive(qempty);  // this thread has read its words of the earlier Q
            ++qit;
          }
          __syncwarp();  // the new rows are complete before the next update reads them
          if (push_item) {
            const bool has_lo = x0 == 0 && warp == 0;
            const bool has_hi = V - 2 >= x0 && V - 2 < x0 + TS && warp == static_cast<int>((V - 2 - x0) / Geo::D::SPW);
            if (has_lo || has_hi) {
              __threadfence_system();
              __syncwarp();
              if (lane == 0) {
                if (has_lo) st_release_sys(hf.hp.seq_hi_of_left, kseq);
                if (has_hi) st_release_sys(hf.hp.seq_lo_of_right, kseq);
              }
            }
          }
        }
      }
      fence_proxy_async();
      mbar_arrive(computed + st);
    }
  }
}

}  // namespace bcg
