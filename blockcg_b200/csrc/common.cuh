// Shared device-side definitions for the block-CG kernels (sm_100a).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <cstdlib>
#include <utility>

namespace bcg {

// BCG_PDL=1 launches the loop's kernels with the programmatic stream-serialization attribute.  OFF by default:
// measured on B200 (profiles/r02_ab_pdl.jsonl) it changes nothing at 24^4 (1.046 vs 1.036 ms per iteration) and
// costs 5 % at 41 472 sites (0.193 vs 0.184 ms): the persistent kernels fill every SM's shared memory, so the
// next kernel's CTAs cannot become resident before the previous grid has drained anyway.
inline bool pdl_enabled() {
  const char* e = std::getenv("BCG_PDL");  // read per launch: A/B runs in one process
  return e ? std::atoi(e) != 0 : false;
}
// kernel<<<grid, block, smem, st>>>(args...) with the programmatic stream-serialization attribute
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(std::forward<Args>(args))...);
}

typedef double2 cd;  // complex128: x = re, y = im

constexpr int kMaxShifts = 32;
constexpr int kNc = 3;  // colours (N_f in the reference, inc/fields.hpp:18)

// ---- complex helpers (each cmac = 4 DFMA) -----------------------------------------
__device__ __forceinline__ cd cmake(double re, double im) { return make_double2(re, im); }
__device__ __forceinline__ cd czero() { return make_double2(0.0, 0.0); }
__device__ __forceinline__ cd cconj(cd a) { return make_double2(a.x, -a.y); }
__device__ __forceinline__ cd cadd(cd a, cd b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cd csub(cd a, cd b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ cd cmul(cd a, cd b) {
  return make_double2(fma(-a.y, b.y, a.x * b.x), fma(a.y, b.x, a.x * b.y));
}
__device__ __forceinline__ cd cscale(cd a, double s) { return make_double2(a.x * s, a.y * s); }
// acc += a*b
__device__ __forceinline__ void cmac(cd& acc, cd a, cd b) {
  acc.x = fma(a.x, b.x, acc.x);
  acc.x = fma(-a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y);
  acc.y = fma(a.y, b.x, acc.y);
}
// acc -= a*b
__device__ __forceinline__ void cmsub(cd& acc, cd a, cd b) {
  acc.x = fma(-a.x, b.x, acc.x);
  acc.x = fma(a.y, b.y, acc.x);
  acc.y = fma(-a.x, b.y, acc.y);
  acc.y = fma(-a.y, b.x, acc.y);
}
// acc += conj(a)*b
__device__ __forceinline__ void cmac_conj(cd& acc, cd a, cd b) {
  acc.x = fma(a.x, b.x, acc.x);
  acc.x = fma(a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y);
  acc.y = fma(-a.y, b.x, acc.y);
}
__device__ __forceinline__ double cabs2(cd a) { return fma(a.x, a.x, a.y * a.y); }
__device__ __forceinline__ cd cdiv(cd a, cd b) {
  // scaled division (robust against over/underflow of |b|^2)
  double s = fmax(fabs(b.x), fabs(b.y));
  double br = b.x / s, bi = b.y / s;
  double d = (br * br + bi * bi) * s;
  return make_double2((a.x * br + a.y * bi) / d, (a.y * br - a.x * bi) / d);
}

// Coefficient operands of the multishift update are stored "interleaved by column
// group": the pipelined kernel gives every site to NSPLIT lanes, lane h producing the
// JC = N/NSPLIT output columns h*JC .. h*JC+JC-1, and the NSPLIT lanes read adjacent
// 16-byte words: element (k, j) of the N x N matrix sits at
// ((k*JC + j%JC)*NSPLIT + j/JC).
__host__ __device__ constexpr int shift_nsplit(int N) {
  return (N % 4 == 0 && N >= 8) ? 4 : (N % 2 == 0 && N >= 4) ? 2 : 1;
}
__host__ __device__ inline int shift_mat_index(int N, int k, int j) {
  const int ns = shift_nsplit(N), jc = N / ns;
  return (k * jc + (j % jc)) * ns + j / jc;
}

// ---- loop control block living in device memory -------------------------------------
// The iteration loop never round-trips to the host: every kernel of an
// iteration looks at `done` (and the stencil at `stop`) and returns early, so
// the host can enqueue batches of iterations blindly and poll a pinned mirror.
struct Ctrl {
  int iter;       // operator applications so far (the reference's return value)
  int stop;       // set by the B-step of the last iteration; the iteration still completes
  int done;       // loop finished: all later kernels are no-ops
  int n_unconv;   // shifts still being updated (block_solvers.hpp:104,179-181)
  int status;     // bcg_status raised on device (not-PD Gram, NaN)
  int max_it;
  int n_shifts;
  int pad0;
  int iter_b;     // copies of iter / n_unconv left by the B-step for the CTAs of the next A-step
  int n_unconv_b;
  double eps;
  double eps_shifts;
  double residual;
  unsigned long long seq_base;     // sequence number of iteration 0 of this solve in the peer-memory exchange
  int conv[kMaxShifts];            // shift converged in the current iteration
  double resid_shift[kMaxShifts];  // last shifted residual estimate
  double sigma[kMaxShifts];
  int n_act[4];                    // n_unconv of the last iterations ([iter % ring]; deferred multishift updates)
  // ---- overlapped multishift update (schedule 3 with BCG_OVERLAP): the shifted systems are updated by a launch
  // that runs BESIDE the next iterations' kernels, so it reads the state of ITS iteration from a snapshot the
  // B-step leaves in slot [iter & 1] (the loop's own fields above have moved on by then) ----
  struct Snap {
    int iter, stop, n_now;
    int n_ring[4];
  } snap[2];
  int bulk_served[2];              // snap[p].iter of the last launch that served slot p (a later one is a no-op)
  // ---- statistics of the solve (read back by bcg_last_solve_stats; nothing in the loop depends on them) ----
  unsigned hist[kMaxShifts + 1];       // hist[a] = iterations that ran with a systems still being updated
  unsigned long long shift_passes;     // field-sized passes (units of F = 48 N V bytes) the multishift update has moved
  // ---- scalar coefficients of CG / SCG (src/standard_solvers.cpp:3-95), N_rhs = 1 only ----
  double sc_r2, sc_r2_0, sc_alpha, sc_beta, sc_alpha_old, sc_beta_old;
  double sc_zeta[kMaxShifts], sc_theta[kMaxShifts];
  double sc_ax[kMaxShifts], sc_bp[kMaxShifts], sc_zr[kMaxShifts];  // per system: x += p*ax ; p = p*bp + r*zr
};

// What one launch of the multishift update does, as a function of the control block the B-step of
// the same iteration has left behind: the list of ITEMS a tile goes through.  Shared by the kernels
// (shift_pair.cuh, shift_dmma.cuh) and by the B-step's byte accounting.
//   schedule 0  plain      : every active system is updated in every iteration.
//   schedule 1  alternating: the shifted systems (s >= 1) never feed back into the main recurrence, so
//                            their two updates of an odd + even iteration pair are applied together in the
//                            even one, from the Q of both iterations (the odd one's is kept in a second field).
//   schedule 2  staggered  : the same deferral, but odd-numbered systems are served in odd iterations and
//                            even-numbered ones in even iterations, so every launch carries the same load
//                            (FP64 work and HBM traffic balanced launch by launch); Q ping-pongs between two
//                            fields (Q -= T alpha writes into the other one), so the previous Q is kept for free.
// A system that retired between the two iterations gets the earlier update only; the iteration the
// loop ends on (`stop`) brings every system up to date.
enum : int { KQ = 0, KQ_KEEP = 1, KQPREV = 2, KCUR = 3, KPREV = 4, KBOTH = 5 };
struct ShiftItem {
  signed char kind;  // KQ: Q <- Q rho^-1 ; KQ_KEEP: ... and kept as Qprev ; KQPREV: previous Q (read only) ;
  signed char s;     // KCUR / KPREV / KBOTH: system s gets this iteration's / the previous one's / both updates
};
constexpr int kMaxShiftItems = kMaxShifts + 2;
// n_now: systems active in this iteration; n_prev: in the previous one (n_act[(iter - 1) & 1]).
// Returns the number of items; *passes = field-sized passes through HBM (reads + writes).
__host__ __device__ inline int build_shift_items(int schedule, int iter, int stop, int n_now, int n_prev,
                                                 ShiftItem* out, int* passes) {
  int n = 0, systems = 0, qpasses = 2;  // Q in, Q out
  const bool odd = (iter & 1) != 0;
#define BCG_ADD_ITEM(k_, s_)                        \
  do {                                              \
    if (out) {                                      \
      out[n].kind = static_cast<signed char>(k_);   \
      out[n].s = static_cast<signed char>(s_);      \
    }                                               \
    ++n;                                            \
  } while (0)
  if (schedule == 1) {
    const int n1 = odd ? n_now : n_prev;  // systems active in the odd iteration of this pair
    if (odd && !stop && n1 > 1) {         // first of a pair: nothing is deferred when the loop ends here
      BCG_ADD_ITEM(KQ_KEEP, -1);
      BCG_ADD_ITEM(KCUR, 0);
      qpasses = 3;
      systems = 1;
    } else if (!odd && n1 > 1) {          // second of a pair
      BCG_ADD_ITEM(KQ, -1);
      BCG_ADD_ITEM(KQPREV, -1);
      BCG_ADD_ITEM(KCUR, 0);
      for (int s = 1; s < n1; ++s) BCG_ADD_ITEM(s < n_now ? KBOTH : KPREV, s);
      qpasses = 3;
      systems = n1;
    } else {
      BCG_ADD_ITEM(KQ, -1);
      for (int s = 0; s < n_now; ++s) BCG_ADD_ITEM(KCUR, s);
      systems = n_now;
    }
  } else if (schedule == 2) {
    const int np = (iter >= 2) ? n_prev : 0;  // the first iteration has no predecessor
    BCG_ADD_ITEM(KQ, -1);
    bool need_prev = false;
    for (int s = 1; s < np; ++s)
      if ((s & 1) == (iter & 1)) need_prev = true;
    if (need_prev) {
      BCG_ADD_ITEM(KQPREV, -1);
      qpasses = 3;
    }
    BCG_ADD_ITEM(KCUR, 0);
    systems = 1;
    const int top = np > n_now ? np : n_now;
    for (int s = 1; s < top; ++s) {
      const bool mine = (s & 1) == (iter & 1), ap = s < np, ac = s < n_now;
      if (mine && (ap || ac)) {
        BCG_ADD_ITEM(ap && ac ? KBOTH : (ap ? KPREV : KCUR), s);
        ++systems;
      } else if (!mine && stop && ac) {  // the loop ends here: this iteration's update of the other group is not deferred
        BCG_ADD_ITEM(KCUR, s);
        ++systems;
      }
    }
  } else {
    BCG_ADD_ITEM(KQ, -1);
    for (int s = 0; s < n_now; ++s) BCG_ADD_ITEM(KCUR, s);
    systems = n_now;
  }
#undef BCG_ADD_ITEM
  if (passes) *passes = qpasses + 4 * systems;
  return n;
}

// ---- schedule 3: staggered deferral of depth k (2 <= k <= ring <= kMaxDepth), streamed operands ---------------
// System s >= 1 is served in the iterations i with i % k == s % k and then receives the updates of the k
// iterations (i-k, i] in order, each with the coefficients and the Q of its own iteration (Q lives in a ring of
// `ring` >= k fields: Q -= T alpha of iteration i reads field (i-1) % ring and writes field i % ring, so the last
// `ring` are intact).  X_s, P_s are read and written once per k iterations: (7 + (k-1) + 4 (S_act-1)/k) F per
// iteration, balanced launch by launch.  A system that retired keeps only the updates of the iterations in which
// it was active; the iteration the loop ends on flushes everything that is pending.  k = 2 is schedule 2.
// `part`: 0 the whole launch ; 1 the critical part only (Q <- Q rho^-1 and system 0: what the next stencil waits
// for) ; 2 the shifted systems only, every Q read from the ring (the launch that overlaps the next iterations,
// ring = k + 1: the fields and operand sets of the last k iterations then survive one more iteration).
constexpr int kMaxDepth = 4;
struct StagItem {
  signed char kind;     // KQ: Q <- Q rho^-1 ; KCUR: system s gets m updates
  signed char s;
  signed char d_first;  // its first pending update is that of iteration iter - d_first
  signed char m;        // number of pending updates (consecutive iterations from there)
};
// n_ring[j % ring] = systems active in iteration j for the previous k-1 iterations; n_now: in this one.
// *passes = field-sized passes through HBM (Q in / out, every distinct Q read from the ring once, 4 per system served).
__host__ __device__ inline int build_stag_items(int k, int ring, int part, int iter, int stop, int n_now,
                                                const int* n_ring, StagItem* out, int* passes) {
  int n = 0, systems = 0;
  unsigned hist_mask = 0;
  if (part != 2) {
    if (out) {
      out[0].kind = KQ; out[0].s = -1; out[0].d_first = 0; out[0].m = 0;
      out[1].kind = KCUR; out[1].s = 0; out[1].d_first = 0; out[1].m = 1;
    }
    n = 2;
    systems = 1;
  }
  int top = n_now;  // no system beyond the largest active count of the last k iterations has anything pending
  for (int t = 0; t < ring; ++t) top = n_ring[t] > top ? n_ring[t] : top;
  if (top > kMaxShifts) top = kMaxShifts;
  for (int s = 1; s < top && part != 1; ++s) {
    const bool mine = (s % k) == (iter % k);
    if (!mine && !stop) continue;
    int first;
    if (mine) {
      first = iter - k + 1;
    } else {
      int back = (iter - s) % k;  // iterations since this system's group was served last
      if (back < 0) back += k;
      first = iter - back + 1;
    }
    if (first < 1) first = 1;
    int m = 0;
    for (int j = first; j <= iter; ++j) {
      const int na = (j == iter) ? n_now : n_ring[j % ring];
      if (s < na) ++m; else break;  // retirement is permanent: the active iterations are a prefix
    }
    if (m == 0) continue;
    if (out) {
      out[n].kind = KCUR;
      out[n].s = static_cast<signed char>(s);
      out[n].d_first = static_cast<signed char>(iter - first);
      out[n].m = static_cast<signed char>(m);
    }
    for (int u = 0; u < m; ++u)
      if (iter - first - u > 0 || part == 2) hist_mask |= 1u << (iter - first - u);
    ++n;
    ++systems;
  }
  int hist = 0;
  for (int d = 0; d < kMaxDepth; ++d) hist += (hist_mask >> d) & 1u;
  if (passes) *passes = (part != 2 ? 2 : 0) + hist + 4 * systems;
  return n;
}

// ---- peer-memory exchange between the ranks of a slab decomposition (NVLink P2P) -------------
// Every rank owns a communication buffer that all peers have mapped (CUDA IPC).  A producer
// writes its contribution straight into the consumers' buffers with ordinary stores, fences at
// system scope and then publishes a sequence number; the consumer spins on the sequence word in
// its OWN memory.  Slots are double-buffered by sequence parity; every exchange is a barrier
// between the ranks, so a producer is never more than one sequence number ahead.
constexpr int kMaxRanks = 8;
struct GramPeers {                      // one of the two Gram channels (0: P^dag T, 1: Q^dag Q)
  cd* slot[kMaxRanks];                  // rank r's block area: [2 parities][nranks][N*N]
  unsigned long long* seq[kMaxRanks];   // rank r's sequence words: [nranks]
  int nranks;                           // 0 = no exchange (one rank, or the NCCL path of the primitives)
  int rank;
  int iter_from_b;                      // channel 1 only: take the iteration number from ctrl->iter_b + 1 (the A-step, which
                                        // writes ctrl->iter, runs BESIDE the producing kernel: AlphaFold)
};
struct HaloPeers {
  cd* lo_of_right;                      // right neighbour's "from the left" slots: [2 parities][2 sites]
  cd* hi_of_left;                       // left neighbour's "from the right" slots
  unsigned long long* seq_lo_of_right;  // sequence word next to each
  unsigned long long* seq_hi_of_left;
  const cd* my_lo;                      // this rank's own slots and sequence words
  const cd* my_hi;
  const unsigned long long* my_seq_lo;
  const unsigned long long* my_seq_hi;
};
// Slab decomposition with the halo exchange folded into the kernels that produce / consume it: the update
// kernel stores the first and last two sites of the new P_0 straight into the neighbours' buffers and
// publishes the sequence number; the stencil of the next iteration waits for its neighbours' and copies
// them into the halo slots of the field in its prologue -- no halo kernel in the loop.
struct HaloFold {
  HaloPeers hp;
  int on;
};
struct GramWait {                       // consumer side of a Gram channel
  const cd* slots;                      // this rank's block area: [2 parities][nranks][N*N]
  const unsigned long long* seq;        // this rank's sequence words: [nranks]
  int nranks;                           // 0 = plain mode (blocks handed over in stream order)
};
// The A-step folded into the Q update (axpy_pipe.cuh): every CTA of `Q' = Q - T alpha` forms alpha = (P0^dag T)^-1
// itself in its prologue, from the Gram block(s) the stencil left behind, while its first tiles are in flight; the
// A-step kernel (which still produces alpha for the B-step, A_0 and the beta_s) leaves the critical path and runs on
// a second stream beside the Q update.  Same functions of the same data in the same order: the same bits.
struct AlphaFold {
  const cd* gsrc;      // reduced Gram block (nsrc == 1) or, with gw.nranks > 0, taken from the peer slots
  int nsrc;
  GramWait gw;
  int step_threads;    // threads of the A-step kernel: its rank-order sum of the peer blocks is sliced by that number
  int on;
};
// One entry of elimination step k of the Gauss-Jordan inverse (pivot row br, rp = 1 / pivot), branch-free: pivot
// row, pivot column and the rest are three dependent chains that a divergent warp would run in turn.
__device__ __forceinline__ cd gj_entry(const cd* src, int N, int i, int j, int k, int br, cd rp) {
  const int si = (i == k) ? br : (i == br) ? k : i;  // source row after the exchange k <-> br
  const cd ask = src[si + N * k], abj = src[br + N * j], asj = src[si + N * j];
  cd rest = asj;
  cmsub(rest, cmul(ask, rp), abj);
  const cd row = cmul(abj, rp);
  const cd col = cmul(cmake(-ask.x, -ask.y), rp);
  return (i == k) ? ((j == k) ? rp : row) : ((j == k) ? col : rest);
}
constexpr long long kSpinTimeoutClocks = 20000000000LL;  // ~10 s: a peer that never arrives is an error, not a hang
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// ---- programmatic dependent launch (griddepcontrol) --------------------------------------------
// Every kernel of the iteration loop starts with pdl_wait(): when it was launched with the programmatic
// stream-serialization attribute its CTAs may become resident -- and run their prologue (barrier
// initialisation, descriptor fetches) -- while the previous kernel of the stream is still draining; the wait
// returns once that kernel has completed and its memory is visible.  Launched normally it is a no-op.
// pdl_trigger() lets the NEXT kernel's CTAs be scheduled as soon as this grid's CTAs have all started.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---- mbarrier + 1-D bulk TMA (cp.async.bulk) ------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// shared -> global, tracked by a bulk async-group
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst),
               "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// make generic-proxy smem writes visible to the async proxy (before a bulk store)
// named barrier `id` (1..15) among `nthreads` threads of the CTA (a multiple of 32; whole warps take part)
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- tensor-map TMA (cp.async.bulk.tensor, SASS UTMALDG / UTMASTG) ----------------------------
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, int c0, int c1, int c2,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
          smem_u32(smem_dst)),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, int c0, int c1, int c2, const void* smem_src) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(map),
               "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(smem_src))
               : "memory");
}

// 2-D tensor-map TMA
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          smem_u32(smem_dst)),
      "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, int c0, int c1, const void* smem_src) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(map), "r"(c0),
               "r"(c1), "r"(smem_u32(smem_src))
               : "memory");
}

// Shared-memory operand load that the high-level optimiser may not hoist out of
// the row loops (LICM of a whole N x N coefficient matrix into registers is a
// guaranteed spill); ptxas still schedules it freely.
__device__ __forceinline__ cd lds_cd(const cd* p) {
  cd r;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "r"(smem_u32(p)));
  return r;
}

// streaming global access (fields are touched once per kernel)
__device__ __forceinline__ cd ldg_stream(const cd* p) {
  cd r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  return r;
}

}  // namespace bcg
