// K1 (second generation): block Dirac apply as a warp-specialised parity-chain pipeline.
//
// Operator (reference inc/dirac_op.hpp:14-21,36-43; shift term block_solvers.hpp:136):
//   D v[x] = 1/2 U[x] v[x+1] - 1/2 U[x-1]^dag v[x-1],   T = (m^2 + sigma) P - D(D P)
// T[x] only couples to P[x-2], P[x], P[x+2]: the even and the odd sub-lattice are closed
// under the operator.  A stencil thread therefore walks a *parity chain* x, x+2, x+4, ...
// for a group of R right-hand sides and keeps (D P)[x-1] of the in-between site in
// registers from one step to the next:
//     tp     = 1/2 (U[x+1] P[x+2] - U[x]^dag P[x])            = (D P)[x+1]
//     T[x]   = (m^2+sigma) P[x] - 1/2 (U[x] tp - U[x-1]^dag tm)        tm = (D P)[x-1]
//     tm <- tp ; x <- x + 2
// so D P is computed exactly once per site and never stored anywhere, and every P / T
// element crosses shared memory once (the first-generation kernel moved 2.5x as much
// through shared memory, which -- not HBM -- was what bound it).
//
// CTA = one SM, persistent:
//   NSW stencil warps : lane = (column group g, parity p, sub-chain k').  The CTA owns K
//                       sub-chains (contiguous site ranges); every tile advances each of
//                       them by W sites.
//   4 Gram warps      : a quarter of the lower triangle of P^dag T each (GramPart), over the
//                       rows of the finished tile, accumulated in registers over the whole
//                       kernel; at the end the CTA's block is reduced across CTAs (group of 8,
//                       then grid; always the last arriver adds, in index order => deterministic)
//                       and, in a slab decomposition, stored into every peer's buffer over NVLink.
//   1 loader lane     : three tensor-map TMA copies per tile (SASS UTMALDG): the P windows and
//                       link windows of all K sub-chains as boxes of a [window][tile][chain]
//                       view, up to three tiles ahead.
//   1 storer lane     : one tensor store (UTMASTG) of the T windows per tile.
// Hand-off is by mbarriers only (inb: TMA -> stencil, ofull: stencil -> Gram + storer,
// gdone: Gram -> loader + stencil, sdone: storer -> stencil); no __syncthreads in the loop.
#pragma once
#include "common.cuh"
#include "field_kernels.cuh"

namespace bcg {

// ---- balanced four-way split of the lower triangle of an N x N Gram block (N even, H = N/2) ----
//   part 0: rows [0,H)      x cols [0,H)  lower triangle      part 1: rows [H,H+H1)  x cols [0,H)
//   part 3: rows [H,N)      x cols [H,N)  lower triangle      part 2: rows [H+H1,N)  x cols [0,H)
// H(H+1)/2, H1*H, (H-H1)*H, H(H+1)/2 entries: 21/18/18/21 at N = 12 -- one warp per part, one
// warp per SM sub-partition, so the FP64 work of the Gram is spread evenly over the four
// schedulers.  Only entries with row >= col are produced (fields.hpp:103-122).
template <int N, int PART>
struct GramPart {
  static_assert(N % 2 == 0, "GramPart needs an even N");
  static constexpr int H = N / 2, H1 = (H + 1) / 2;
  static constexpr int R0 = (PART == 0) ? 0 : (PART == 1) ? H : (PART == 2) ? H + H1 : H;
  static constexpr int NR = (PART == 0) ? H : (PART == 1) ? H1 : (PART == 2) ? H - H1 : H;
  static constexpr int C0 = (PART == 3) ? H : 0;
  static constexpr int NC = H;
  static constexpr bool LOWER = (PART == 0 || PART == 3);
  cd acc[NR > 0 ? NR : 1][NC];
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int i = 0; i < NR; ++i)
#pragma unroll
      for (int j = 0; j < NC; ++j) acc[i][j] = czero();
  }
  // one row (site, colour): pa / pb point at column 0 of the row in A / B, columns 3 apart
  __device__ __forceinline__ void row(const cd* __restrict__ pa, const cd* __restrict__ pb) {
    cd a[NR > 0 ? NR : 1], b[NC];
#pragma unroll
    for (int i = 0; i < NR; ++i) a[i] = pa[3 * (R0 + i)];
#pragma unroll
    for (int j = 0; j < NC; ++j) b[j] = pb[3 * (C0 + j)];
#pragma unroll
    for (int i = 0; i < NR; ++i)
#pragma unroll
      for (int j = 0; j < NC; ++j)
        if (!LOWER || j <= i) cmac_conj(acc[i][j], a[i], b[j]);
  }
  // Cross-lane sum and store of the finished accumulators.  Code that runs once per launch is
  // paid for in instruction-cache misses (~40 cycles per cold instruction), so nothing here is
  // unrolled beyond the register-to-shared spill: every lane parks its NENT accumulators in
  // `scratch` ([entry][lane], NENT * 32 complex, private to this warp), then lane t adds up the
  // 32 values of entry t in lane order (fixed => deterministic) and writes it into the N x N
  // column-major block.
  static constexpr int NENT = LOWER ? NC * (NC + 1) / 2 : NR * NC;
  __device__ __forceinline__ void store(cd* __restrict__ dstNN, cd* __restrict__ scratch) {
    const int lane = threadIdx.x & 31;
    int idx = 0;
#pragma unroll
    for (int i = 0; i < NR; ++i)
#pragma unroll
      for (int j = 0; j < NC; ++j) {
        if (LOWER && j > i) continue;
        scratch[idx * 32 + lane] = acc[i][j];
        ++idx;
      }
    __syncwarp();
#pragma unroll 1
    for (int t = lane; t < NENT; t += 32) {
      int i = 0, j = t;  // entry t in the enumeration order above
      if (LOWER) {
        while (j > i) {
          j -= i + 1;
          ++i;
        }
      } else {
        i = t / NC;
        j = t - i * NC;
      }
      double re = 0.0, im = 0.0;
#pragma unroll 4
      for (int l = 0; l < 32; ++l) {
        const cd v = scratch[t * 32 + l];
        re += v.x;
        im += v.y;
      }
      dstNN[(R0 + i) + N * (C0 + j)] = cmake(re, im);
    }
  }
};

// ---- the same lower triangle on the FP64 tensor instruction (mma.sync.m8n8k4.f64, SASS DMMA) -------
// G = A^dag B as a real product over rows:  S[(i,pi)][(j,pj)] = sum_rows A[row][i].pi * B[row][j].pj
// (2N x 2N real, parts interleaved: real index 2 i + part), then  Re G_ij = S(i0,j0) + S(i1,j1),
// Im G_ij = S(i0,j1) - S(i1,j0).  One instruction takes K = 4 rows and an 8 x 8 tile of S (4 x 4 complex
// entries); only tiles on or below the diagonal are computed (N/4 (N/4 + 1) / 2 of them: 6 at N = 12, i.e.
// 384 multiply-adds per row against 312 for the exact triangle) -- a few more flops on the same FP64 pipe,
// but 1.5 DMMA + 1.5 LDS.64 per row instead of ~10 DFMA + 0.4 LDS.128 per row and lane: the Gram warps stop
// competing with the stencil warps for issue slots, and their six accumulation chains are independent.
// Lane (q = lane % 4, mm = lane / 4): A fragment = A[row q][real column 8 mt + mm], B fragment likewise.
template <int N>
struct GramDmma {
  static_assert(N % 4 == 0, "GramDmma needs N a multiple of 4");
  static constexpr int NB = N / 4;
  static constexpr int NTILE = NB * (NB + 1) / 2;
  double acc[NTILE][2];
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int t = 0; t < NTILE; ++t) acc[t][0] = acc[t][1] = 0.0;
  }
  // pa / pb: this lane's double of tile 0 in row q of the quad (column block t is 24 doubles = 4 complex
  // columns x 3 colours further on); rows outside the field contribute zero
  __device__ __forceinline__ void quad(const double* __restrict__ pa, const double* __restrict__ pb, bool valid) {
    double a[NB], b[NB];
#pragma unroll
    for (int t = 0; t < NB; ++t) {
      double va, vb;
      asm volatile("ld.shared.f64 %0, [%1];" : "=d"(va) : "r"(smem_u32(pa + 24 * t)));
      asm volatile("ld.shared.f64 %0, [%1];" : "=d"(vb) : "r"(smem_u32(pb + 24 * t)));
      a[t] = valid ? va : 0.0;
      b[t] = valid ? vb : 0.0;
    }
    int idx = 0;
#pragma unroll
    for (int mt = 0; mt < NB; ++mt)
#pragma unroll
      for (int nt = 0; nt <= mt; ++nt) {
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                     : "+d"(acc[idx][0]), "+d"(acc[idx][1])
                     : "d"(a[mt]), "d"(b[nt]));
        ++idx;
      }
  }
  // Combine the (re, im) parts across lane pairs, park this warp's tiles in `scratch` ([4 warps][N*N] complex,
  // shared by the four Gram warps), then the 128 threads add the four blocks in warp order (fixed =>
  // deterministic) and write the lower block triangle of the N x N column-major block.
  __device__ __forceinline__ void store(cd* __restrict__ dstNN, cd* __restrict__ scratch, int gw) {
    const int lane = threadIdx.x & 31;
    int idx = 0;
#pragma unroll
    for (int mt = 0; mt < NB; ++mt)
#pragma unroll
      for (int nt = 0; nt <= mt; ++nt) {
        const double o0 = __shfl_xor_sync(0xffffffffu, acc[idx][0], 4);
        const double o1 = __shfl_xor_sync(0xffffffffu, acc[idx][1], 4);
        if (((lane >> 2) & 1) == 0) {
          const int i = 4 * mt + (lane >> 3), j = 4 * nt + (lane & 3);
          scratch[gw * N * N + i + N * j] = cmake(acc[idx][0] + o1, acc[idx][1] - o0);
        }
        ++idx;
      }
    asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll 1
    for (int e = gw * 32 + lane; e < N * N; e += 128) {
      const int i = e % N, j = e / N;
      if (i / 4 < j / 4) continue;  // block above the diagonal: never produced, never consumed
      cd sum = scratch[e];
#pragma unroll
      for (int w = 1; w < 4; ++w) sum = cadd(sum, scratch[w * N * N + e]);
      dstNN[e] = sum;
    }
  }
};

// ---- second level of the Gram reduction, inside the producing kernel ---------------------------
// The per-CTA partial blocks are not handed to the coefficient kernel one by one (a single CTA
// pulling ~150 blocks out of L2 is a 5-10 us latency chain in every iteration): CTAs form groups
// of kGramGroup consecutive blockIdx, and the group member that finishes LAST adds the group's
// blocks, in blockIdx order, into one block.  Which member is last varies from run to run, what
// it computes does not => still bit-reproducible, no floating-point atomics.
// Buffer layout (complex units of N*N): [0] final block | [kGramSumOff, +ngroups) group sums |
// [kGramRawOff, +grid) raw partial blocks | [kGramCntOff] arrival counters (unsigned, zero at rest:
// whoever completes a count resets it).
constexpr int kGramGroup = 8;
constexpr int kGramSumOff = 8;
constexpr int kGramRawOff = 1024;
constexpr int kGramCntOff = 3072;
// Third level, same pattern: the CTA that completes the LAST group adds the group sums in group
// order into the final block gbuf[0] -- and, in a slab decomposition, stores that block straight
// into every peer's communication buffer over NVLink and publishes the sequence number of the
// iteration: the all-"reduce" of the Gram matrix is the tail of the kernel that produced it
// (each rank then adds the nranks blocks in rank order inside its coefficient kernel, so all
// ranks get bit-identical matrices and stay in lock-step without a broadcast).
// Called by the 4 Gram warps (128 threads, named barrier 1) after each has stored its part of the
// raw block.  Counters: [0, 1023) groups, [1023] number of finished groups.
template <int N>
__device__ __noinline__ void gram_group_reduce(cd* __restrict__ gbuf, int gram_warp, const GramPeers& peers,
                                               const Ctrl* __restrict__ ctrl, int channel) {
  constexpr int nn = N * N;
  __shared__ unsigned s_old;
  const int t = gram_warp * 32 + (threadIdx.x & 31);  // 0 .. 127
  const int group = blockIdx.x / kGramGroup;
  const int ngroups = (static_cast<int>(gridDim.x) + kGramGroup - 1) / kGramGroup;
  const int first = group * kGramGroup;
  const int members = min(kGramGroup, static_cast<int>(gridDim.x) - first);
  unsigned* cnt = reinterpret_cast<unsigned*>(gbuf + static_cast<size_t>(kGramCntOff) * nn);
  __threadfence();
  asm volatile("bar.sync 1, 128;" ::: "memory");
  if (t == 0) s_old = atomicAdd(cnt + group, 1u);
  asm volatile("bar.sync 1, 128;" ::: "memory");
  if (s_old != static_cast<unsigned>(members - 1)) return;
  __threadfence();
  // ---- this CTA completed its group: add the group's raw blocks in blockIdx order ----
  const cd* raw = gbuf + (static_cast<size_t>(kGramRawOff) + first) * nn;
  cd* gsum = gbuf + static_cast<size_t>(kGramSumOff) * nn;
#pragma unroll 1
  for (int e = t; e < nn; e += 128) {
    const int i = e % N, j = e / N;
    if (i < j) continue;  // only the lower triangle is produced and consumed
    cd v[kGramGroup];
#pragma unroll
    for (int m = 0; m < kGramGroup; ++m)  // all loads in flight, then a fixed-order sum
      v[m] = (m < members) ? __ldcg(reinterpret_cast<const double2*>(raw + static_cast<size_t>(m) * nn + e)) : czero();
    cd sum = v[0];
#pragma unroll
    for (int m = 1; m < kGramGroup; ++m) sum = cadd(sum, v[m]);  // absent members contribute +0
    gsum[static_cast<size_t>(group) * nn + e] = sum;
  }
  __threadfence();
  asm volatile("bar.sync 1, 128;" ::: "memory");
  if (t == 0) {
    cnt[group] = 0u;
    s_old = atomicAdd(cnt + 1023, 1u);
  }
  asm volatile("bar.sync 1, 128;" ::: "memory");
  if (s_old != static_cast<unsigned>(ngroups - 1)) return;
  __threadfence();
  // ---- ... and the last group: add the group sums in group order, publish ----
  // iteration number of this Gram: the A-step increments ctrl->iter between the two channels; with the A-step running
  // beside the Q update (AlphaFold) channel 1 derives it from the previous B-step's copy instead
  const int it_now = (ctrl == nullptr) ? 0 : (channel == 0 ? ctrl->iter + 1 : (peers.iter_from_b ? ctrl->iter_b + 1 : ctrl->iter));
  const unsigned long long k = (ctrl != nullptr) ? ctrl->seq_base + static_cast<unsigned long long>(it_now) : 0ull;
  const int np = (ctrl != nullptr) ? peers.nranks : 0;
  constexpr int CH = 20;  // group sums fetched per round (148 SMs / 8 = 19 groups: one round)
#pragma unroll 1
  for (int e = t; e < nn; e += 128) {
    const int i = e % N, j = e / N;
    if (i < j) continue;
    double re = 0.0, im = 0.0;
#pragma unroll 1
    for (int g0 = 0; g0 < ngroups; g0 += CH) {
      cd v[CH];
#pragma unroll
      for (int g = 0; g < CH; ++g)
        v[g] = (g0 + g < ngroups) ? __ldcg(reinterpret_cast<const double2*>(gsum + static_cast<size_t>(g0 + g) * nn + e))
                                  : czero();
#pragma unroll
      for (int g = 0; g < CH; ++g) {
        re += v[g].x;
        im += v[g].y;
      }
    }
    const cd sum = cmake(re, im);
    gbuf[e] = sum;
#pragma unroll 1
    for (int r = 0; r < np; ++r)
      peers.slot[r][(static_cast<size_t>(k & 1ull) * np + peers.rank) * nn + e] = sum;
  }
  if (t == 0) cnt[1023] = 0u;
  if (np > 0) {
    __threadfence_system();
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (t < np) st_release_sys(peers.seq[t] + peers.rank, k);
  }
}

template <int N, int G, int K, int W>
struct ChainGeom {
  static_assert(N % G == 0, "G must divide N");
  static_assert(G == 1 || G == 2 || G == 4 || G == 8 || G == 16, "G must be a power of two <= 16");
  static_assert(W % 2 == 0 && W >= 2, "window must hold whole parity pairs");
  static constexpr int R = N / G;                 // rhs columns per stencil thread
  static constexpr int SITE = 3 * N;              // complex per site
  static constexpr int KW = 16 / G;               // sub-chains per stencil warp
  static_assert(K % KW == 0, "K must fill whole warps");
  static constexpr int NSW = K / KW;              // stencil warps
  static constexpr int NGW = 4;                   // Gram warps: one GramPart each, one per SM sub-partition
  static constexpr int NWARPS = NSW + NGW + 2;
  static constexpr int NT = NWARPS * 32;
  static constexpr int WARP_LOAD = NSW + NGW, WARP_STORE = NSW + NGW + 1;
  // window pitches in complex (16-byte) units: field windows start on residues 0,1,2,.. mod 8
  // (conflict-free Gram rows), link windows on residues 0,6,4,2 (conflict-free link broadcast)
  static constexpr int PP = W * SITE + ((1 - (W * SITE) % 8 + 8) % 8);
  static constexpr int PU = (W + 2) * 9 + ((6 - ((W + 2) * 9) % 8 + 8) % 8);
  static constexpr int SP = 5, SU = 4, SO = 2;    // ring depths: P windows, link windows, T windows
  static constexpr int NBAR_IN = 4, NBAR_G = 4;
  static constexpr int P_ELEMS = SP * K * PP, U_ELEMS = SU * K * PU, O_ELEMS = SO * K * PP;
  static constexpr size_t SMEM_BYTES =
      sizeof(cd) * (P_ELEMS + U_ELEMS + O_ELEMS) + 8 * (NBAR_IN + SO + NBAR_G + SO) + 16;
  static constexpr int ROWS = 3 * K * W;          // Gram rows per tile
  // lanes of a quarter warp must hit 8 distinct 16-byte bank groups when they read their P
  // columns: when consecutive sites are 4 (mod 8) units apart the parity bit goes inside
  // the quarter warp, otherwise a sub-chain bit does (windows are 1 (mod 8) apart).
  static constexpr bool PARITY_INNER = (G != 4) || (SITE % 8 == 4);
};

// GMODE 0: no Gram.  1: four dedicated Gram warps working one tile behind the stencil warps (DFMA, GramPart).
// 3: the same four warps on the FP64 tensor instruction (GramDmma).
// 2 (experiment, not built by default): no dedicated warps -- after each tile the four stencil
// warps wait for one another and compute one GramPart each over the tile they have just
// finished, accumulators in their own registers (255 per thread with 6 warps per CTA).  Measured
// 3x SLOWER than mode 1 (334 vs 106 us at 24^4, N=12): every warp then runs its own ~18 KB copy
// of stencil + Gram loop, which no longer fits the per-sub-partition instruction cache.  The
// lesson cuts the other way too: hot loops are kept small and rolled wherever that is free.
template <int N, int G, int K, int W, int GMODE>
__global__ void __launch_bounds__((GMODE == 2 ? (ChainGeom<N, G, K, W>::NSW + 2) * 32 : ChainGeom<N, G, K, W>::NT), 1)
dirac_chain_kernel(const __grid_constant__ CUtensorMap tmP, const __grid_constant__ CUtensorMap tmO,
                   const __grid_constant__ CUtensorMap tmU, const cd* __restrict__ in,
                   const cd* __restrict__ U, long long V, long long L, double m2, double sigma,
                   cd* __restrict__ gpart, const Ctrl* __restrict__ ctrl, const GramPeers peers, const HaloFold hf) {
  using Geo = ChainGeom<N, G, K, W>;
  constexpr int WARP_LOAD = (GMODE == 2) ? Geo::NSW : Geo::WARP_LOAD;
  constexpr int WARP_STORE = WARP_LOAD + 1;
  constexpr int R = Geo::R, SITE = Geo::SITE, PP = Geo::PP, PU = Geo::PU;
  constexpr int SP = Geo::SP, SU = Geo::SU, SO = Geo::SO;
  pdl_wait();
  pdl_trigger();
  if (ctrl != nullptr && (ctrl->done | ctrl->stop)) return;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  cd* sP = reinterpret_cast<cd*>(smem_raw);
  cd* sU = sP + Geo::P_ELEMS;
  cd* sO = sU + Geo::U_ELEMS;
  uint64_t* inb = reinterpret_cast<uint64_t*>(sO + Geo::O_ELEMS);
  uint64_t* ofull = inb + Geo::NBAR_IN;
  uint64_t* gdone = ofull + SO;
  uint64_t* sdone = gdone + Geo::NBAR_G;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < Geo::NBAR_IN; ++i) mbar_init(inb + i, 1);
    for (int i = 0; i < SO; ++i) mbar_init(ofull + i, Geo::NSW);
    for (int i = 0; i < Geo::NBAR_G; ++i) mbar_init(gdone + i, (GMODE == 1 || GMODE == 3) ? Geo::NGW : Geo::NSW);
    for (int i = 0; i < SO; ++i) mbar_init(sdone + i, 1);
    mbar_fence_init();
  }
  __syncthreads();

  const int T = static_cast<int>(L / W);                      // tiles
  const long long chain0 = static_cast<long long>(blockIdx.x) * K;  // first sub-chain of this CTA
  // ---- folded halo exchange (slab decomposition): wait for the neighbours' boundary sites of this P ----
  const cd* lo_src = nullptr;  // sites -2, -1 as the left neighbour stored them (read by the chain start of chain 0)
  if (hf.on && ctrl != nullptr) {
    const unsigned long long kseq = ctrl->seq_base + static_cast<unsigned long long>(ctrl->iter);
    const int hn = 2 * SITE;
    // sites this CTA reads: its chains' starts (generic loads) and windows 0 .. T of chains chain0 .. chain0 + K (TMA)
    const bool need_lo = chain0 == 0;
    const bool need_hi = chain0 * L <= V + 1 && (chain0 + K + 1) * L + W > V;
    if (need_lo || need_hi) {
      __shared__ int halo_timed_out;
      if (tid == 0) halo_timed_out = 0;
      __syncthreads();
      if ((tid == 0 && need_lo) || (tid == 1 && need_hi)) {
        const unsigned long long* wseq = tid ? hf.hp.my_seq_hi : hf.hp.my_seq_lo;
        const long long t0 = clock64();
        while (ld_acquire_sys(wseq) < kseq)
          if (clock64() - t0 > kSpinTimeoutClocks) {
            halo_timed_out = 1;
            break;
          }
      }
      __syncthreads();
      if (halo_timed_out) {
        if (tid == 0) {
          Ctrl* cw = const_cast<Ctrl*>(ctrl);
          cw->status = 4;
          cw->done = 1;
        }
        return;
      }
      if (need_hi) {  // slots V, V+1 are read through the tensor map: they must be in the field itself
        cd* fw = const_cast<cd*>(in);
        const cd* hi = hf.hp.my_hi + (kseq & 1ull) * hn;
        for (int i = tid; i < hn; i += blockDim.x) fw[V * SITE + i] = __ldcg(reinterpret_cast<const double2*>(hi + i));
        __threadfence();
        asm volatile("fence.proxy.async;" ::: "memory");  // generic-proxy stores -> visible to the tensor copies below
      }
      __syncthreads();
    }
    lo_src = hf.hp.my_lo + (kseq & 1ull) * hn;
  }
  // sub-chain q covers out sites [q*L, min((q+1)*L, V))
  auto chain_start = [&](int k) { return (chain0 + k) * L; };
  auto chain_end = [&](int k) {
    const long long e = (chain0 + k + 1) * L;
    return e < V ? e : V;
  };

  constexpr uint32_t P_BOX_BYTES = K * PP * sizeof(cd), U_BOX_BYTES = K * PU * sizeof(cd);
  const int q0 = static_cast<int>(chain0);

  if (warp == WARP_LOAD) {
    // ===================== loader: three tensor copies per tile, one elected lane =====================
    // input barrier of tile t covers P-load t+1 (sites o0+W .. o0+2W) and link-load t
    // (links o0-1 .. o0+W+1) of all K sub-chains; tile 0 also brings P-load 0.
    // Window j of sub-chain q is box row (j, q) of the 3-D view (see make_chain_maps):
    // the view is flat in (window, chain), so window T of chain q is window 0 of chain q+1.
    if (lane != 0) return;
    for (int t = 0; t < T; ++t) {
      if (t >= 4) mbar_wait(gdone + (t & 3), static_cast<uint32_t>(((t - 4) >> 2) & 1));
      uint64_t* bar = inb + (t & 3);
      mbar_arrive_expect_tx(bar, (t == 0 ? 2u : 1u) * P_BOX_BYTES + U_BOX_BYTES);
      if (t == 0) tma_load_3d(sP, &tmP, 0, 0, q0, bar);
      const int j = t + 1;
      tma_load_3d(sP + (j % SP) * (K * PP), &tmP, 0, j == T ? 0 : j, j == T ? q0 + 1 : q0, bar);
      tma_load_3d(sU + (t % SU) * (K * PU), &tmU, 0, t, q0, bar);
    }
    return;
  }

  if (warp == WARP_STORE) {
    // ===================== storer: one tensor store per tile =====================
    // the pad element of every window row lies outside dimension 0 of the view: not written
    if (lane != 0) return;
    for (int t = 0; t < T; ++t) {
      mbar_wait(ofull + (t % SO), static_cast<uint32_t>((t / SO) & 1));
      tma_store_3d(&tmO, 0, t, q0, sO + (t % SO) * (K * PP));
      bulk_commit();
      bulk_wait_read0();  // shared memory of this tile has been read
      mbar_arrive(sdone + (t % SO));
    }
    bulk_wait0();
    return;
  }

  if (warp >= Geo::NSW) {
    // ===================== Gram warps: block (ti,tj) of P^dag T =====================
    if constexpr (GMODE == 3) {
      // row quad = sub-chains 4 gw .. 4 gw + 3 at one (site s of the window, colour c): lane q = lane % 4 reads
      // sub-chain 4 gw + q; the four Gram warps split the K = 16 sub-chains, each takes all W * 3 (s, c)
      static_assert(K == 16, "the tensor-instruction Gram splits 16 sub-chains over four warps");
      const int gw = warp - Geo::NSW, q = lane & 3, mm = lane >> 2;
      const int kch = 4 * gw + q;
      const long long rem = chain_end(kch) - chain_start(kch);  // sites of this lane's sub-chain (<= 0: empty)
      const int off = 2 * (kch * PP) + 6 * (mm >> 1) + (mm & 1);
      GramDmma<N> gd;
      gd.init();
      for (int t = 0; t < T; ++t) {
        mbar_wait(ofull + (t % SO), static_cast<uint32_t>((t / SO) & 1));
        const double* dP = reinterpret_cast<const double*>(sP + (t % SP) * (K * PP)) + off;
        const double* dO = reinterpret_cast<const double*>(sO + (t % SO) * (K * PP)) + off;
#pragma unroll
        for (int sc = 0; sc < 3 * W; ++sc) {
          const int s = sc / 3, c = sc - 3 * s;
          gd.quad(dP + 2 * (s * SITE + c), dO + 2 * (s * SITE + c), t * W + s < rem);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(gdone + (t & 3));
      }
      gd.store(gpart + (static_cast<size_t>(kGramRawOff) + blockIdx.x) * N * N, sP + ((T + 1) % SP) * (K * PP), gw);
      gram_group_reduce<N>(gpart, gw, peers, ctrl, 0);
      return;
    }
    if (GMODE != 1) return;
    if constexpr (GMODE == 1) {
    // row = (colour c, site s of the window, sub-chain k), k fastest across lanes: the 8 lanes of a
    // quarter warp read 8 consecutive windows, which start on 8 different 16-byte bank groups
    auto gram_loop = [&](auto& part) {
      part.init();
      for (int t = 0; t < T; ++t) {
        mbar_wait(ofull + (t % SO), static_cast<uint32_t>((t / SO) & 1));
        const cd* tP = sP + (t % SP) * (K * PP);  // P-load t holds P at this tile's out sites
        const cd* tO = sO + (t % SO) * (K * PP);
        // not unrolled: four different Gram loops plus the stencil loop must stay resident in the
        // instruction cache together (stall_no_instruction tripled when they did not)
#pragma unroll 1
        for (int it = 0; it < (Geo::ROWS + 31) / 32; ++it) {
          const int rr = lane + 32 * it;
          const int k = rr % K, s = (rr / K) % W, c = rr / (K * W);
          const long long rem = chain_end(k) - chain_start(k);  // sites of sub-chain k (<= 0: empty)
          if (rr < Geo::ROWS && t * W + s < rem) {
            const int base = k * PP + s * SITE + c;
            part.row(tP + base, tO + base);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(gdone + (t & 3));
      }
      // scratch: a P ring slot the last tile does not use (slot (T-1) % SP is still being read)
      part.store(gpart + (static_cast<size_t>(kGramRawOff) + blockIdx.x) * N * N,
                 sP + ((T + (warp - Geo::NSW)) % SP) * (K * PP));
      gram_group_reduce<N>(gpart, warp - Geo::NSW, peers, ctrl, 0);
    };
    switch (warp - Geo::NSW) {
      case 0: { GramPart<N, 0> part; gram_loop(part); break; }
      case 1: { GramPart<N, 1> part; gram_loop(part); break; }
      case 2: { GramPart<N, 2> part; gram_loop(part); break; }
      default: { GramPart<N, 3> part; gram_loop(part); break; }
    }
    }
    return;
  }

  // ===================== stencil warps =====================
  struct NoPart {
    __device__ __forceinline__ void init() {}
    __device__ __forceinline__ void row(const cd*, const cd*) {}
    __device__ __forceinline__ void store(cd*, cd*) {}
  };
  auto stencil = [&](auto& part) {
    part.init();
    int g, p, kq;
    {
      const int l8 = lane & 7, q8 = lane >> 3;
      if (G == 4) {
        g = l8 & 3;
        if (Geo::PARITY_INNER) {
          p = l8 >> 2;
          kq = q8;
        } else {
          p = q8 & 1;
          kq = (q8 >> 1) * 2 + (l8 >> 2);
        }
      } else {
        g = lane % G;
        p = (lane / G) & 1;
        kq = lane / (2 * G);
      }
    }
    const int k = warp * Geo::KW + kq;
    const long long cs = chain_start(k);
    const long long rem_ll = chain_end(k) - cs;
    const int rem = rem_ll < 0 ? 0 : static_cast<int>(rem_ll);  // valid out sites of this sub-chain
    const int col0 = g * R * 3;  // first complex of this thread's columns inside a site
    const double ms = m2;

    cd tm[R][3];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) tm[r][c] = czero();
    if (p < rem) {
      // chain start: (D P)[x-1] for the first site of this chain, straight from global memory
      const long long x = cs + p;
      cd v[R][3], acc[R][3];
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) acc[r][c] = czero();
      load_cols<N, R>(in + x * SITE + col0, v);
      apply_link<R>(U + (x - 1) * 9, v, acc);
      if (lo_src != nullptr && x < 2) {  // sites -2, -1: straight from the communication buffer (written by the peer)
        const cd* src = lo_src + x * SITE + col0;
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
          for (int c = 0; c < 3; ++c) v[r][c] = __ldcg(reinterpret_cast<const double2*>(src + r * 3 + c));
      } else {
        load_cols<N, R>(in + (x - 2) * SITE + col0, v);
      }
      apply_link_dag_sub<R>(U + (x - 2) * 9, v, acc);
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) tm[r][c] = cscale(acc[r][c], 0.5);
    }

    for (int t = 0; t < T; ++t) {
      mbar_wait(inb + (t & 3), static_cast<uint32_t>((t >> 2) & 1));
      if (t >= SO) {  // T window slot free again: Gram and store of tile t-SO are through with it
        if (GMODE == 1 || GMODE == 3) mbar_wait(gdone + ((t - SO) & 3), static_cast<uint32_t>(((t - SO) >> 2) & 1));
        mbar_wait(sdone + (t % SO), static_cast<uint32_t>(((t - SO) / SO) & 1));
      }
      const cd* tP0 = sP + (t % SP) * (K * PP) + k * PP;         // P at out sites o0 .. o0+W
      const cd* tP1 = sP + ((t + 1) % SP) * (K * PP) + k * PP;   // P at o0+W .. o0+2W
      const cd* tU = sU + (t % SU) * (K * PU) + k * PU;          // links o0-1 .. o0+W+1
      cd* tO = sO + (t % SO) * (K * PP) + k * PP;
#pragma unroll
      for (int i = 0; i < W / 2; ++i) {
        const int ls = p + 2 * i;  // out site o0 + ls
        cd v[R][3], tp[R][3], acc[R][3];
        // tp = 1/2 (U[x+1] P[x+2] - U[x]^dag P[x])
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
          for (int c = 0; c < 3; ++c) acc[r][c] = czero();
        load_cols<N, R>((ls + 2 < W ? tP0 + (ls + 2) * SITE : tP1 + (ls + 2 - W) * SITE) + col0, v);
        apply_link<R>(tU + (ls + 2) * 9, v, acc);
        load_cols<N, R>(tP0 + ls * SITE + col0, v);
        apply_link_dag_sub<R>(tU + (ls + 1) * 9, v, acc);
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
          for (int c = 0; c < 3; ++c) tp[r][c] = cscale(acc[r][c], 0.5);
        // T[x] = (m^2 + sigma) P[x] - 1/2 (U[x] tp - U[x-1]^dag tm)
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
          for (int c = 0; c < 3; ++c) acc[r][c] = czero();
        apply_link<R>(tU + (ls + 1) * 9, tp, acc);
        apply_link_dag_sub<R>(tU + ls * 9, tm, acc);
        {  // sites past the end of the field land in the allocation slack (never read back)
          cd* o = tO + ls * SITE + col0;
#pragma unroll
          for (int r = 0; r < R; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              // lhs = -D(D rhs) + m^2 rhs (dirac_op.hpp:42), then += sigma rhs (block_solvers.hpp:136)
              cd tt = cmake(fma(ms, v[r][c].x, -0.5 * acc[r][c].x), fma(ms, v[r][c].y, -0.5 * acc[r][c].y));
              tt.x = fma(sigma, v[r][c].x, tt.x);
              tt.y = fma(sigma, v[r][c].y, tt.y);
              o[r * 3 + c] = tt;
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
          for (int c = 0; c < 3; ++c) tm[r][c] = tp[r][c];
      }
      fence_proxy_async();  // T window visible to the bulk store
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(ofull + (t % SO));
        if (GMODE == 0) mbar_arrive(gdone + (t & 3));  // no Gram: the stencil releases the input slots itself
      }
      if (GMODE == 2) {
        // every stencil warp has written its T windows of tile t: this warp's quarter of the Gram
        mbar_wait(ofull + (t % SO), static_cast<uint32_t>((t / SO) & 1));
        const cd* gP = sP + (t % SP) * (K * PP);
        const cd* gO = sO + (t % SO) * (K * PP);
#pragma unroll 1
        for (int it = 0; it < (Geo::ROWS + 31) / 32; ++it) {
          const int rr = lane + 32 * it;
          const int gk = rr % K, gs = (rr / K) % W, gc = rr / (K * W);
          const long long grem = chain_end(gk) - chain_start(gk);
          if (rr < Geo::ROWS && t * W + gs < grem) {
            const int base = gk * PP + gs * SITE + gc;
            part.row(gP + base, gO + base);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(gdone + (t & 3));
      }
    }
    if (GMODE == 2) {
      part.store(gpart + (static_cast<size_t>(kGramRawOff) + blockIdx.x) * N * N, sP + ((T + warp) % SP) * (K * PP));
      gram_group_reduce<N>(gpart, warp, peers, ctrl, 0);
    }
  };
  if constexpr (GMODE == 2) {
    static_assert(GMODE != 2 || Geo::NSW == 4, "one GramPart per stencil warp");
    switch (warp) {
      case 0: { GramPart<N, 0> part; stencil(part); break; }
      case 1: { GramPart<N, 1> part; stencil(part); break; }
      case 2: { GramPart<N, 2> part; stencil(part); break; }
      default: { GramPart<N, 3> part; stencil(part); break; }
    }
  } else {
    NoPart part;
    stencil(part);
  }
}

}  // namespace bcg
