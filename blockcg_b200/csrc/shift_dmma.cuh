// K4, third generation: the multishift update on the FP64 tensor instruction.
//
//     Q <- Q rho^-1 ;  for every active system:  X_s += P_s A_s ;  P_s <- P_s B_s + Q
// (reference: inc/block_solvers.hpp:145,152,158,175-177; inc/fields.hpp:70-90,125-136), in the plain
// and in the paired schedule of shift_pair.cuh (shifted systems served every second iteration).
//
// Why: with DFMA the update is bound by shared-memory wavefronts, not by HBM or the FP64 pipe
// (ncu, round 1: L1/shared 78 %, FP64 57 %, DRAM 45 %): per 36 DFMA a lane issues six 16-byte
// LDS -- three P words and three coefficient words -- and half of them re-read the same N x N
// matrix for every site.  `mma.sync.m8n8k4.f64` (SASS DMMA) runs on the same FP64 pipe at the same
// rate (tools/micro/fp64_pipes.cu), so it buys no flops, but it takes its operands as register
// FRAGMENTS: the coefficient matrix is read from shared memory once per warp and item instead of
// once per site, and a P row is read once per product.  ~34 instead of ~90 wavefronts per site and
// system update: the kernel becomes FP64-pipe / HBM bound.
//
// The complex product is evaluated as a real one with no padding at all:
//   [Xre Xim] (rows x 2N) += [Pre Pim] (rows x 2N) . [[Are, Aim], [-Aim, Are]] (2N x 2N)
// K = 2N real = N/2 steps of 4, N-dimension 2N real = N/4 tiles of 8 (N a multiple of 4), M = 8 rows
// = the 8 sites of a warp at one colour.  Real index 2 kk + part interleaves (re, im), so
//   * an A fragment element is one double of the staged P tile (row m = lane / 4, k = lane % 4),
//   * a C fragment pair (c0, c1) is exactly one complex number (row lane / 4, column 4 jt + lane % 4):
//     the epilogue adds it to X, or adds Q to it and stores it as the new P, with 16-byte accesses.
// Same flop count as the complex DFMA form (4 real multiply-adds per complex one).
//
// Arithmetic order differs from the DFMA kernels (the four products of a k-step are summed inside the
// instruction), so solutions differ from theirs in the last bits.  Every (S)BCGrQ update -- plain,
// paired, single-shift -- goes through THIS kernel when it is enabled, so the invariants the tests
// check bit for bit (paired == plain, SBCGrQ shift 0 == BCGrQ) still hold.
//
// Structure as shift_pair_kernel: one producer lane issuing tensor copies into a two-stage ring, four
// compute warps, stages updated in place and stored back, two CTAs per SM.
#pragma once
#include "common.cuh"
#include "field_kernels.cuh"
#include "shift_pair.cuh"

namespace bcg {

template <int N, int TS, int NST = 2>
struct ShiftDmmaGeom {
  static_assert(N % 4 == 0, "the real-expanded product needs 2N a multiple of 8");
  static constexpr int NSPLIT = shift_nsplit(N), JC = N / NSPLIT;  // layout of the coefficient operands (shift_mat_index)
  static constexpr int SPW = 8;                 // sites per compute warp = rows of one DMMA
  static_assert(TS % SPW == 0 && TS % 2 == 0, "tile must fill whole warps and whole site pairs");
  static constexpr int NCW = TS / SPW;          // compute warps
  static constexpr int NT = (NCW + 1) * 32;
  static constexpr int NCT = NCW * 32;
  static constexpr int SITE = 3 * N;
  // pair pitch = 2 (mod 8) sixteen-byte words: with SITE = 4 (mod 8) (N = 4, 12) the four sites a half warp
  // touches start on word residues {0, 4, 2, 6}: fragment loads (LDS.64) and epilogue accesses
  // (LDS/STS.128) are bank-conflict free
  static constexpr int PAIR = 2 * SITE + ((2 - (2 * SITE) % 8 + 8) % 8);
  static constexpr int TILE = (TS / 2) * PAIR;
  static constexpr int KS = N / 2, NTL = N / 4;  // k-steps, n-tiles of one product
  static constexpr int NSTAGE = NST;
  static constexpr int STAGE_ELEMS = (2 * TILE + 4 * N * N + 7) / 8 * 8;  // P, X tiles + (A', B', A, B)
  static constexpr int SCRATCH = 3 * NTL * NCT;                           // Qprev words, [word][thread]
  static constexpr size_t SMEM_BYTES = sizeof(cd) * (NSTAGE * STAGE_ELEMS + SCRATCH) + 64;
  static constexpr bool TWO_CTAS = 2 * (SMEM_BYTES + 1024) <= 227 * 1024 && NT <= 256;
  // registers are granted per thread: two CTAs of 160 threads may use up to 204 each
  static constexpr int MAXREG = TWO_CTAS ? 152 : (NT <= 256 ? 232 : 168);
};

__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}
__device__ __forceinline__ double lds_f64(const double* p) {
  double r;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(r) : "r"(smem_u32(p)));
  return r;
}

// B fragments of one N x N complex coefficient matrix (interleaved layout of shift_mat_index) as the
// 2N x 2N real matrix [[Mre, Mim], [-Mim, Mre]]: fragment (js, jt), lane (k = lane % 4, n = lane / 4)
// holds entry (4 js + k, 8 jt + n) = part (pk ^ pn) of M(2 js + k / 2, 4 jt + n / 2), negated for
// (pk, pn) = (im, re).
template <int N>
__device__ __forceinline__ void load_coef_frags(const cd* __restrict__ sM, int lane, double (&f)[N / 2][N / 4]) {
  constexpr int NS = shift_nsplit(N), JCc = N / NS;
  const int k = lane & 3, n = lane >> 2;
  const int pk = k & 1, pn = n & 1;
  const double* base = reinterpret_cast<const double*>(sM) + (pk ^ pn);
  const int sign = (pk && !pn) ? static_cast<int>(0x80000000u) : 0;  // negation = one integer XOR on the high word
#pragma unroll
  for (int js = 0; js < N / 2; ++js) {
    const int kk = 2 * js + (k >> 1);
#pragma unroll
    for (int jt = 0; jt < N / 4; ++jt) {
      const int j = 4 * jt + (n >> 1);
      const double v = lds_f64(base + 2 * ((kk * JCc + (j % JCc)) * NS + j / JCc));
      f[js][jt] = __hiloint2double(__double2hiint(v) ^ sign, __double2loint(v));
    }
  }
}

// (Measured and dropped, profiles/r02_ab_shift_onepass_experiment.jsonl: both coefficient matrices of an update
// register-resident, every P row read once for both products -- 180 registers, 1.39 ms per launch against 0.86;
// profiles/r02_ab_shift_chain_scheduling_experiment.jsonl: tensor instructions not `volatile`, or all three colour
// rows of a product in flight at once (nine accumulation chains) -- no gain: the accumulation chains are not what
// the kernel waits for.)
template <int N, int TS, int NST>
__global__ void __maxnreg__((ShiftDmmaGeom<N, TS, NST>::MAXREG))
shift_dmma_kernel(const __grid_constant__ ShiftPairMaps maps, const cd* __restrict__ Rrecip,
                  const cd* __restrict__ Aodd, const cd* __restrict__ Bodd, const cd* __restrict__ Aeven,
                  const cd* __restrict__ Beven, long long V, const Ctrl* __restrict__ ctrl, int schedule,
                  cd* __restrict__ p0_halo, const HaloFold hf) {
  // p0_halo != nullptr: field P of system 0 (site 0); the kernel then also writes the periodic images of its
  // first and last two sites into the halo slots (sites V, V+1 and -2, -1), which the stencil reads next
  // Aodd/Bodd: operand slots written in odd iterations, Aeven/Beven: in even ones ([shift][N*N] each;
  // the same slots when the schedule is not paired)
  using Geo = ShiftDmmaGeom<N, TS, NST>;
  constexpr int NS = Geo::NSTAGE;
  constexpr int NCW = Geo::NCW, SITE = Geo::SITE, PAIR = Geo::PAIR, TILE = Geo::TILE, STAGE = Geo::STAGE_ELEMS;
  constexpr int NN = N * N, KS = Geo::KS, NTL = Geo::NTL;
  pdl_wait();
  pdl_trigger();
  if (ctrl->done) return;
  const int iter = ctrl->iter;
  const bool odd = (iter & 1) != 0;
  const cd* Acur = odd ? Aodd : Aeven;
  const cd* Bcur = odd ? Bodd : Beven;
  const cd* Aprev = odd ? Aeven : Aodd;  // operands the previous iteration left in the other parity's slots
  const cd* Bprev = odd ? Beven : Bodd;
  // the items of a tile (same list in every thread: built once, read from shared memory)
  __shared__ ShiftItem s_items[kMaxShiftItems];
  __shared__ int s_n_items;
  if (threadIdx.x == 0)
    s_n_items = build_shift_items(schedule, iter, ctrl->stop, ctrl->n_unconv, ctrl->n_act[(iter - 1) & 1], s_items, nullptr);

  extern __shared__ __align__(128) unsigned char smem_raw[];
  cd* sbuf = reinterpret_cast<cd*>(smem_raw);
  cd* scratch = sbuf + Geo::NSTAGE * STAGE;
  uint64_t* full = reinterpret_cast<uint64_t*>(scratch + Geo::SCRATCH);
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
  const int n_items = s_n_items;

  if (warp == NCW) {
    // ===================== producer (as shift_pair_kernel) =====================
    if (lane != 0) return;
    long long it = 0;
    int d_pair[NS], d_kind[NS], d_s[NS];  // items in flight
    for (int i = 0; i < NS; ++i) {
      d_pair[i] = 0;
      d_kind[i] = KQPREV;
      d_s[i] = -1;
    }
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
      for (int k = 0; k < n_items; ++k, ++it) {
        const int st = static_cast<int>(it % NS);
        const uint32_t use = static_cast<uint32_t>(it / NS);
        if (it >= NS) {
          mbar_wait(computed + st, (use - 1) & 1u);  // item it-NS has been computed in place
          store_item(st);
          bulk_wait_read0();
        }
        const int kind = s_items[k].kind, s = s_items[k].s;
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
    for (long long k = (it >= NS ? it - NS : 0); k < it; ++k) {  // drain the last (up to NS) items
      const int st = static_cast<int>(k % NS);
      mbar_wait(computed + st, static_cast<uint32_t>(k / NS) & 1u);
      store_item(st);
    }
    bulk_wait0();
    return;
  }

  // ===================== compute warps =====================
  const int m = lane >> 2, q = lane & 3;  // fragment coordinates: row (site of the warp) / k or column index
  const int lsite = warp * Geo::SPW + m;  // site within the tile
  const int sbase = (lsite >> 1) * PAIR + (lsite & 1) * SITE;
  // back-substitution of the Q tile keeps the DFMA form: lanes q = 0, 1, 2 take colour row q of the site
  cd* myq = scratch + tid;  // word w of this thread's Qprev at myq[w * NCT]
  long long it = 0;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long x0 = tile * TS;
    const int ns = static_cast<int>(min(static_cast<long long>(TS), V - x0));
    const bool live = lsite < ns;
    cd qf[3][NTL];  // this lane's C-fragment positions of the new Q: (colour c, column 4 jt + q)
    for (int k = 0; k < n_items; ++k, ++it) {
      const int kind = s_items[k].kind;
      const bool halo_item = p0_halo != nullptr && s_items[k].s == 0;  // new P_0: sites 0, 1 and V-2, V-1 also go to the halo slots
      // slab decomposition: ... or straight into the neighbours' buffers (P2P stores over NVLink), published below
      const bool push_item = hf.on && s_items[k].s == 0;
      const unsigned long long kseq = ctrl->seq_base + static_cast<unsigned long long>(iter);
      cd* to_left = hf.hp.hi_of_left + (kseq & 1ull) * (2 * SITE);    // my sites 0, 1     -> left neighbour's slots V, V+1
      cd* to_right = hf.hp.lo_of_right + (kseq & 1ull) * (2 * SITE);  // my sites V-2, V-1 -> right neighbour's slots -2, -1
      const long long xs = x0 + lsite;
      const int st = static_cast<int>(it % NS);
      mbar_wait(full + st, static_cast<uint32_t>(it / NS) & 1u);
      cd* buf = sbuf + st * STAGE;
      if (kind == KQ || kind == KQ_KEEP) {
        if (q < 3 && live) {
          cd qr[N];
#pragma unroll
          for (int kk = 0; kk < N; ++kk) qr[kk] = buf[sbase + 3 * kk + q];
          row_backsub<N>(qr, buf + 2 * TILE);
#pragma unroll
          for (int kk = 0; kk < N; ++kk) buf[sbase + 3 * kk + q] = qr[kk];
        }
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
          for (int jt = 0; jt < NTL; ++jt) qf[c][jt] = buf[sbase + 3 * (4 * jt + q) + c];
      } else if (kind == KQPREV) {
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
          for (int jt = 0; jt < NTL; ++jt) myq[(c * NTL + jt) * Geo::NCT] = buf[sbase + 3 * (4 * jt + q) + c];
      } else {
        cd* sP = buf + sbase;
        cd* sX = sP + TILE;
        const double* dP = reinterpret_cast<const double*>(sP) + 6 * (q >> 1) + (q & 1);  // A fragment: k = q
        const int nup = (kind == KBOTH) ? 2 : 1;
#pragma unroll 1
        for (int u = 0; u < nup; ++u) {
          const cd* sA = buf + 2 * TILE + 2 * u * NN;
          const cd* sB = sA + NN;
          const bool from_prev = (kind == KPREV) || (kind == KBOTH && u == 0);
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
              const double a = lds_f64(dP + 2 * (6 * js + c));  // part (q & 1) of P(row m, column 2 js + q / 2)
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
            // every lane of the warp has read its fragments of colour row c (the instruction is warp-wide)
            // before any lane gets here: the rows can be overwritten in place
            if (live) {
#pragma unroll
              for (int jt = 0; jt < NTL; ++jt) {
                const cd qq = from_prev ? myq[(c * NTL + jt) * Geo::NCT] : qf[c][jt];
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
          __syncwarp();  // the new rows are complete before the second update reads them
          if (push_item) {
            // both boundary sites of a side sit in one warp (V even, tiles and warps hold whole site pairs): the stores
            // of its lanes are fenced at system scope, then one lane publishes the iteration's sequence number
            const bool has_lo = x0 == 0 && warp == 0;
            const bool has_hi = V - 2 >= x0 && V - 2 < x0 + TS && warp == static_cast<int>((V - 2 - x0) / Geo::SPW);
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
