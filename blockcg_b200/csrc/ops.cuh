// Host-side launchers, one table per compiled N (see inst.cu / registry in capi.cu).
#pragma once
#include <cstdlib>
#include <cstring>
#include <mutex>
#include "common.cuh"
#include "field_kernels.cuh"
#include "dirac_chain.cuh"
#include "axpy_pipe.cuh"
#include "dirac4d.cuh"
#include "dirac4_tile.cuh"
#include "shift_pair.cuh"
#include "shift_dmma.cuh"
#include "shift_stag.cuh"

namespace bcg {

struct OpsTable {
  int N;
  int dirac_rhs_per_thread;
  int dirac_tile_sites;
  int fused_gram;  // 1 if the Gram epilogue is fused into K1/K3 at this N
  // Each launcher returns the number of partial Gram blocks it produced (0 if none),
  // or a negative cudaError_t.
  // peers: peer-memory targets of the fused Gram exchange (nullptr = none)
  int (*dirac)(cudaStream_t st, const cd* in, cd* out, const cd* U, long long V, double m2, double sigma,
               cd* gpart, const Ctrl* ctrl, int sms, int* launches, const GramPeers* peers, const HaloFold* hf);
  // first-generation tile kernel (intermediate staged in shared memory), kept for comparison
  int (*dirac_v1)(cudaStream_t st, const cd* in, cd* out, const cd* U, long long V, double m2, double sigma,
                  cd* gpart, const Ctrl* ctrl, int sms, int* launches);
  // one sweep of the 4-D operator (dirac4d.cuh): second == 0: out = D in ; else out = (m2 + sigma) p0 - D in
  // (sites [x_begin, x_end) only)
  int (*dirac4_sweep)(cudaStream_t st, const cd* in, const cd* p0, cd* out, const cd* U, const Lattice4* lat,
                      long long x_begin, long long x_end, double m2, double sigma, int second, const Ctrl* ctrl,
                      int* launches);
  // the same sweep, second generation (dirac4_tile.cuh): rows [row_begin, row_end) of the local lattice (row = L0
  // consecutive sites), Ut = direction-major links.  gpart != nullptr (second sweep only): every CTA also leaves its
  // partial Gram p0^dag out -- returns their number; otherwise 0; < 0: error (-cudaErrorNotSupported: the lattice
  // row does not fit a tile, use dirac4_sweep).  nullptr where the kernel does not fit the SM at this N.
  int (*dirac4_tile)(cudaStream_t st, const cd* in, const cd* p0, cd* out, const cd* Ut, const Lattice4* lat,
                     long long site_stride_mu, long long row_begin, long long row_end, double m2, double sigma, int second,
                     cd* gpart, const Ctrl* ctrl, int sms, int* launches);
  int (*gram)(cudaStream_t st, const cd* A, const cd* B, long long V, cd* gpart, const Ctrl* ctrl, int sms,
              int* launches);
  // Q += T*M (+ Gram of the result); Qout != nullptr: the result goes to Qout, Q is left untouched
  // (only where out_of_place_axpy is set)
  int (*axpy_gram)(cudaStream_t st, cd* Q, const cd* T, const cd* M, long long V, cd* gpart,
                   const Ctrl* ctrl, int sms, int* launches, const GramPeers* peers, cd* Qout);
  int out_of_place_axpy;
  int fused_exchange;  // 1 if dirac / axpy_gram push the final Gram block to the peers themselves
  int (*axpy_gram_v1)(cudaStream_t st, cd* Q, const cd* T, const cd* M, long long V, cd* gpart,
                      const Ctrl* ctrl, int sms, int* launches);
  int (*rescale_add)(cudaStream_t st, cd* dst, const cd* L, const cd* src, double r, long long V, int sms,
                     int* launches);
  int (*trsm)(cudaStream_t st, cd* Q, const cd* R, long long V, const Ctrl* ctrl, int sms, int* launches);
  int (*shift_update)(cudaStream_t st, cd* Q, const ShiftPtrs* fp, const cd* R, const cd* A, const cd* B,
                      long long V, int do_backsub, int n_active_fixed, const Ctrl* ctrl, int sms,
                      int* launches);
  // the same update with the first-generation register-direct kernel (kept for comparison)
  int (*shift_update_direct)(cudaStream_t st, cd* Q, const ShiftPtrs* fp, const cd* R, const cd* A, const cd* B,
                             long long V, int do_backsub, int n_active_fixed, const Ctrl* ctrl, int sms,
                             int* launches);
  // paired multishift update (shift_pair.cuh): shifted systems served every second iteration; nullptr
  // where the pipelined kernel is not built.  A_odd / B_odd: operand slots of odd iterations.
  int (*shift_update_pair)(cudaStream_t st, cd* Q, cd* Qprev, const ShiftPtrs* fp, const cd* Rm, const cd* A_odd,
                           const cd* B_odd, const cd* A_even, const cd* B_even, long long V, const Ctrl* ctrl,
                           int sms, int* launches);
  // the same update (plain or paired schedule) on the FP64 tensor instruction (shift_dmma.cuh); nullptr
  // where it is not built (N not a multiple of 4).  schedule: 0 plain, 1 alternating, 2 staggered
  // (build_shift_items; 2 wants Q = this iteration's Q field, Qprev = the previous iteration's).
  int (*shift_update_dmma)(cudaStream_t st, cd* Q, cd* Qprev, const ShiftPtrs* fp, const cd* Rm, const cd* A_odd,
                           const cd* B_odd, const cd* A_even, const cd* B_even, long long V, const Ctrl* ctrl,
                           int sms, int* launches, int schedule, cd* p0_halo, const HaloFold* hf);
  // the same update with the shifted systems served once per `depth` iterations (schedule 3, shift_stag.cuh):
  // Qring[t] = the Q field of the iterations with i % depth == t, coefs = the operand sets likewise
  // ring = fields / operand sets in use (>= depth), part = 0 whole launch / 1 Q and system 0 / 2 shifted systems
  // from the snapshot `slot`, max_ctas > 0 caps the grid (build_stag_items, shift_stag_kernel)
  int (*shift_update_stag)(cudaStream_t st, cd* const* Qring, int depth, int ring, int part, int slot, int max_ctas,
                           const ShiftPtrs* fp, const cd* Rm, const ShiftStagCoefs* coefs, long long V, const Ctrl* ctrl,
                           int sms, int* launches, cd* p0_halo, const HaloFold* hf);
  int (*max_partials)(int sms);
  // sites that must be allocated after site 0 of every field / of the links (>= V + 2): the
  // tensor-map views of the chain stencil are rectangular and reach past the end of the field
  long long (*field_capacity)(long long V, int sms);
  void (*prepare)(int sms);  // occupancy queries / smem opt-in; call once outside stream capture
  // axpy_gram with M = -alpha formed in the kernel's prologue from the stencil's Gram (AlphaFold, axpy_pipe.cuh): the
  // A-step kernel need not have run.  nullptr where the pipelined kernel is not built or the scratch does not fit.
  int (*axpy_gram_fold)(cudaStream_t st, cd* Q, const cd* T, long long V, cd* gpart, const Ctrl* ctrl, int sms,
                        int* launches, const GramPeers* peers, cd* Qout, const AlphaFold* fold);
};

const OpsTable* get_ops(int N);  // nullptr if N is not compiled in

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time libcuda dependency)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}
// 3-D view (window, tile, sub-chain) of a site-major array, all sizes in complex numbers:
// window = win complex (dimension 0, contiguous), consecutive windows win_stride apart, sub-chain
// rows row_stride = T * win_stride apart (the view is flat in (tile, chain)); box = K rows of `box`
// complex each.  box > win pads every row in shared memory: the surplus elements are outside
// dimension 0, i.e. zero-filled on load and not written on store.
inline int make_chain_map(CUtensorMap* m, const cd* base, int win, int win_stride, int box, long long T,
                          long long rows, int K) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return -static_cast<int>(cudaErrorNotSupported);
  const cuuint64_t dims[3] = {static_cast<cuuint64_t>(2 * win), static_cast<cuuint64_t>(T), static_cast<cuuint64_t>(rows)};
  const cuuint64_t strides[2] = {static_cast<cuuint64_t>(win_stride) * sizeof(cd),
                                 static_cast<cuuint64_t>(win_stride) * sizeof(cd) * static_cast<cuuint64_t>(T)};
  const cuuint32_t bx[3] = {static_cast<cuuint32_t>(2 * box), 1u, static_cast<cuuint32_t>(K)};
  const cuuint32_t es[3] = {1u, 1u, 1u};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<cd*>(base), dims, strides, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -static_cast<int>(cudaErrorInvalidValue);
}

// 2-D view [site pair][2 * site complex] of a field; box = `rows` pairs of `box` complex each
// (box > 2 * site pads every pair in shared memory, see make_chain_map)
inline int make_pair_map(CUtensorMap* m, const cd* base, int site, int box, long long npairs, int rows) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return -static_cast<int>(cudaErrorNotSupported);
  const cuuint64_t dims[2] = {static_cast<cuuint64_t>(4 * site), static_cast<cuuint64_t>(npairs)};
  const cuuint64_t strides[1] = {static_cast<cuuint64_t>(2 * site) * sizeof(cd)};
  const cuuint32_t bx[2] = {static_cast<cuuint32_t>(2 * box), static_cast<cuuint32_t>(rows)};
  const cuuint32_t es[2] = {1u, 1u};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<cd*>(base), dims, strides, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -static_cast<int>(cudaErrorInvalidValue);
}

// traversal order of the pipelined Q += T*M (see axpy_pipe_kernel); BCG_AXPY_REVERSE=0 for A/B runs
inline int axpy_reverse() {
  static const int v = [] {
    const char* e = std::getenv("BCG_AXPY_REVERSE");
    return e ? (e[0] != '0') : 1;
  }();
  return v;
}

// diagnostic: BCG_FORCE_V1 = bit mask of kernels to run in their first-generation form
// (1 stencil, 2 Q += T*M, 4 multishift update) -- for A/B comparisons of accuracy and speed
inline int force_v1() {
  static const int v = [] {
    const char* e = std::getenv("BCG_FORCE_V1");
    return e ? std::atoi(e) : 0;
  }();
  return v;
}

#ifndef BCG_CHAIN_GMODE
#define BCG_CHAIN_GMODE 1
#endif

// fused Gram epilogues of the stencil and of Q += T*M on the FP64 tensor instruction (GramDmma, dirac_chain.cuh);
// BCG_GRAM_DMMA=0 selects the DFMA form (GramPart).  Read per launch: A/B runs in one process.
inline int gram_dmma() {
  const char* e = std::getenv("BCG_GRAM_DMMA");
  return e ? std::atoi(e) : 1;
}

#ifdef BCG_N  // ---- per-N implementation, included only by inst.cu ----------------------------

constexpr int kNT = 192;  // 6 warps: one 4x4 Gram block per warp at N = 12, whole sites per CTA

template <int N>
struct Tune {
  // rhs columns per stencil work item (must divide N)
  static constexpr int R = (N % 3 == 0) ? 3 : (N % 2 == 0) ? 2 : 1;
  // parity-chain stencil (dirac_chain.cuh): column groups per site, 0 = not used at this N
  // (N = 16 would need a finer split of the Gram: 36 accumulators per warp do not fit in registers)
  // N = 16: eight column groups of two (the Gram runs on the tensor instruction there: GramDmma needs no
  // per-entry accumulators, which is what kept the DFMA form from fitting in registers)
  static constexpr int CHAIN_G = (N % 4 == 0 && N <= 12) ? 4 : (N == 16 ? 8 : 0);
  static constexpr int CHAIN_K = 16;  // sub-chains per CTA
  static constexpr int CHAIN_W = 2;   // sites per sub-chain and tile
};

template <typename K>
static int occupancy_blocks(K kernel, int threads, size_t smem, int sms) {
  int nb = 0;
  if (smem > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, threads, smem);
  return (nb < 1 ? 1 : nb) * sms;
}

template <int N>
struct Ops {
  static constexpr int R = Tune<N>::R;
  using DG = DiracGeom<N, R, kNT>;
  using AG = AxpyGeom<N, kNT>;
  using GG = GramKGeom<N, kNT>;
  static constexpr bool FUSED = DG::CAN_GRAM && AG::CAN_GRAM;
  static constexpr int GRAM_Y = (GramGeom<N>::NTASK + (kNT / 32) - 1) / (kNT / 32);
  static constexpr size_t SHIFT_SMEM = sizeof(cd) * 5 * N * N;
  // sites per pipeline tile: 32 with two resident CTAs per SM measured 2.3 % faster than 64 with one
  static constexpr int SHIFT_TS = 32;
  using SG = ShiftGeom<N, SHIFT_TS>;
  // a tensor-map box row (one padded site pair, in doubles) may not exceed 256 elements
  static constexpr bool PIPE_OK = SG::SMEM_BYTES <= 227 * 1024 && 2 * SG::PAIR <= 256;
  using SPG = ShiftPairGeom<N, SHIFT_TS>;
  static constexpr bool PAIR_OK = PIPE_OK && SPG::SMEM_BYTES <= 227 * 1024;
  static constexpr bool DMMA_N = (N % 4 == 0);
  using SDG = ShiftDmmaGeom<DMMA_N ? N : 4, SHIFT_TS>;
  static constexpr bool DMMA_OK = DMMA_N && SDG::SMEM_BYTES <= 227 * 1024 && 2 * SDG::PAIR <= 256;
  // experimental shapes of the same kernel (BCG_DMMA_CFG = 1: 32 sites x 3 stages, one CTA per SM;
  // 2: 64 sites x 2 stages, eight compute warps in one CTA per SM)
  using SDG1 = ShiftDmmaGeom<DMMA_N ? N : 4, 32, 3>;
  using SDG2 = ShiftDmmaGeom<DMMA_N ? N : 4, 64, 2>;
  static constexpr bool DMMA_CFG1 = DMMA_OK && N <= 12 && SDG1::SMEM_BYTES <= 227 * 1024;
  static constexpr bool DMMA_CFG2 = DMMA_OK && N <= 12 && SDG2::SMEM_BYTES <= 227 * 1024;
  static constexpr bool APIPE = (N % 2 == 0 && N >= 4 && N <= 12) || N == 16;  // pipelined Q += T*M (axpy_pipe.cuh)
  static constexpr bool DFMA_GRAM = N <= 12;  // GramPart (DFMA accumulators per entry) fits in registers
  static constexpr int APIPE_TS = 32;
  using APG = AxpyPipeGeom<APIPE ? N : 4, APIPE_TS>;
  static constexpr bool CHAIN = Tune<N>::CHAIN_G > 0;
  static constexpr int CG_ = CHAIN ? Tune<N>::CHAIN_G : 1, CK = Tune<N>::CHAIN_K, CW = Tune<N>::CHAIN_W;
  using CGm = ChainGeom<N, CG_, CHAIN ? CK : 16 / CG_, CW>;
  static constexpr int CGMODE = BCG_CHAIN_GMODE;  // how the fused Gram is scheduled (dirac_chain.cuh)
  static constexpr int CNT_G = (CGMODE == 2) ? (CGm::NSW + 2) * 32 : CGm::NT;

  // resident-CTA capacities (CTAs per SM x SMs), filled once by prepare() --
  // outside any stream capture -- and used to size the persistent grids.
  struct Caps {
    int dirac_g = 0, dirac = 0, gram = 0, axpy_g = 0, axpy = 0, rescale = 0, trsm = 0, shift = 0, pipe = 0, pair = 0, dmma = 0, dmma1 = 0, dmma2 = 0, stag = 0;
  };
  // Both the shared-memory opt-ins (cudaFuncSetAttribute) and the occupancy figures are per
  // DEVICE: one slot per device ordinal, filled under a lock the first time a context of that
  // device asks (contexts on several GPUs in one process, possibly from several host threads).
  static constexpr int kMaxDevices = 64;
  static Caps& caps() {
    static Caps c[kMaxDevices];
    int dev = 0;
    cudaGetDevice(&dev);
    return c[(dev >= 0 && dev < kMaxDevices) ? dev : 0];
  }
  static void prepare(int sms) {
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    Caps& c = caps();
    if (c.dirac) return;
    if constexpr (FUSED) {
      c.dirac_g = occupancy_blocks(dirac_kernel<N, R, kNT, true>, kNT, DG::SMEM_BYTES, sms);
      c.axpy_g = occupancy_blocks(axpy_gram_kernel<N, kNT, true>, kNT, AG::SMEM_BYTES, sms);
    }
    c.dirac = occupancy_blocks(dirac_kernel<N, R, kNT, false>, kNT, DG::SMEM_BYTES, sms);
    c.gram = occupancy_blocks(gram_kernel<N, kNT>, kNT, GG::SMEM_BYTES, sms);
    c.axpy = occupancy_blocks(axpy_gram_kernel<N, kNT, false>, kNT, sizeof(cd) * N * N, sms);
    c.rescale = occupancy_blocks(rescale_add_kernel<N, kNT>, kNT, 0, sms);
    c.trsm = occupancy_blocks(trsm_kernel<N, kNT>, kNT, 0, sms);
    c.shift = occupancy_blocks(shift_update_kernel<N, kNT>, kNT, SHIFT_SMEM, sms);
    if constexpr (PIPE_OK) c.pipe = occupancy_blocks(shift_pipe_kernel<N, SHIFT_TS>, SG::NT, SG::SMEM_BYTES, sms);
    if constexpr (PAIR_OK) c.pair = occupancy_blocks(shift_pair_kernel<N, SHIFT_TS>, SG::NT, SPG::SMEM_BYTES, sms);
    if constexpr (DMMA_OK) c.dmma = occupancy_blocks(shift_dmma_kernel<N, SHIFT_TS, 2>, SDG::NT, SDG::SMEM_BYTES, sms);
    if constexpr (DMMA_CFG1) c.dmma1 = occupancy_blocks(shift_dmma_kernel<N, 32, 3>, SDG1::NT, SDG1::SMEM_BYTES, sms);
    if constexpr (STAG_OK) c.stag = occupancy_blocks(shift_stag_kernel<N, SHIFT_TS>, SSG::NT, SSG::SMEM_BYTES, sms);
    if constexpr (DMMA_CFG2) c.dmma2 = occupancy_blocks(shift_dmma_kernel<N, 64, 2>, SDG2::NT, SDG2::SMEM_BYTES, sms);
    if constexpr (APIPE) {
      if constexpr (DFMA_GRAM)
        cudaFuncSetAttribute(axpy_pipe_kernel<N, APIPE_TS, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)APG::SMEM_BYTES);
      cudaFuncSetAttribute(axpy_pipe_kernel<N, APIPE_TS, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)APG::SMEM_BYTES);
      if constexpr (N % 4 == 0) {
        cudaFuncSetAttribute(axpy_pipe_kernel<N, APIPE_TS, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)APG::SMEM_BYTES);
        // the A-step kernel runs BESIDE this one (AlphaFold): ask for the largest shared-memory carve-out, so that an SM
        // configured for this kernel's CTA still has room for a coefficient CTA (and vice versa)
        cudaFuncSetAttribute(axpy_pipe_kernel<N, APIPE_TS, 2>, cudaFuncAttributePreferredSharedMemoryCarveout,
                             (int)cudaSharedmemCarveoutMaxShared);
      }
    }
    dirac4_tile_prepare<2>();
    dirac4_tile_prepare<3>();
    dirac4_tile_prepare<4>();
    dirac4_tile_prepare<6>();
    dirac4_tile_prepare<8>();
    dirac4_tile_prepare<12>();
    dirac4_tile_prepare<16>();
    dirac4_tile_prepare<24>();
    if constexpr (CHAIN) {
      if constexpr (DFMA_GRAM)
        cudaFuncSetAttribute(dirac_chain_kernel<N, CG_, CK, CW, CGMODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)CGm::SMEM_BYTES);
      cudaFuncSetAttribute(dirac_chain_kernel<N, CG_, CK, CW, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)CGm::SMEM_BYTES);
      if constexpr (N % 4 == 0 && CK == 16)
        cudaFuncSetAttribute(dirac_chain_kernel<N, CG_, CK, CW, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)CGm::SMEM_BYTES);
    }
  }

  // one persistent CTA per SM; the CTA's K sub-chains are contiguous ranges of L sites (T tiles of W)
  struct ChainPlan {
    int grid;
    long long L, T, nchains;
  };
  static ChainPlan chain_plan(long long V, int sms) {
    ChainPlan p;
    const long long per_tile = static_cast<long long>(CK) * CW;
    p.grid = clamp_grid((V + per_tile - 1) / per_tile, sms);
    p.nchains = static_cast<long long>(p.grid) * CK;
    p.L = (V + p.nchains - 1) / p.nchains;
    p.L = (p.L + CW - 1) / CW * CW;
    p.T = p.L / CW;
    return p;
  }
  static long long field_capacity(long long V, int sms) {
    if constexpr (CHAIN) {
      const ChainPlan p = chain_plan(V, sms);
      const long long cap = (p.nchains + 1) * p.L + CW + 2;
      return cap > V + 2 ? cap : V + 2;
    }
    return V + 2;
  }

  static int err() {
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : -static_cast<int>(e);
  }
  static int clamp_grid(long long ntiles, int cap) {
    int grid = static_cast<int>(ntiles < cap ? ntiles : cap);
    return grid < 1 ? 1 : grid;
  }

  static int dirac4_sweep(cudaStream_t st, const cd* in, const cd* p0, cd* out, const cd* U, const Lattice4* lat,
                          long long x_begin, long long x_end, double m2, double sigma, int second, const Ctrl* ctrl,
                          int* launches) {
    if (x_end <= x_begin) return 0;
    // right-hand sides per thread: 1 keeps a warp's 16-byte loads within 12 lines per instruction
    // (3 would spread them over 36); BCG_DIRAC4_R=3 selects the wider variant for comparison
    static const bool wide = [] { const char* e = std::getenv("BCG_DIRAC4_R"); return e && std::atoi(e) == R && R > 1; }();
    if (wide) {
      const unsigned grid = static_cast<unsigned>(((x_end - x_begin) * (N / R) + 127) / 128);
      if (second)
        dirac4_kernel<N, R, true><<<grid, 128, 0, st>>>(in, p0, out, U, *lat, x_begin, x_end, m2, sigma, ctrl);
      else
        dirac4_kernel<N, R, false><<<grid, 128, 0, st>>>(in, p0, out, U, *lat, x_begin, x_end, m2, sigma, ctrl);
    } else {
      const unsigned grid = static_cast<unsigned>(((x_end - x_begin) * N + 127) / 128);
      if (second)
        dirac4_kernel<N, 1, true><<<grid, 128, 0, st>>>(in, p0, out, U, *lat, x_begin, x_end, m2, sigma, ctrl);
      else
        dirac4_kernel<N, 1, false><<<grid, 128, 0, st>>>(in, p0, out, U, *lat, x_begin, x_end, m2, sigma, ctrl);
    }
    if (launches) ++*launches;
    return err();
  }

  // Shapes of the tiled 4-D sweep compiled in: NCW compute warps, a two-stage ring, as many CTAs per SM as fit.
  // Measured (24^4, N = 12, profiles/r02_4d_tile_shapes.jsonl): many small CTAs beat one large one with a deep ring
  // -- 3 warps x 4 CTAs 331 us, 6 warps x 2 CTAs 397 us, 6 warps x 1 CTA with four stages 518 us -- the sweep waits
  // for row deliveries, and independent pipelines hide that better than depth.  So: the smallest shape whose tile
  // holds one lattice row.
  using D4 = Dirac4TileGeom<N, 3, 2>;  // what the shapes share: R, G, SITE
  template <int NCW>
  static int dirac4_tile_launch(cudaStream_t st, const cd* in, const cd* p0, cd* out, const cd* Ut, const Rows4& geo,
                                long long row_begin, long long row_end, double m2, double sigma, int second, cd* gpart,
                                const Ctrl* ctrl, int sms) {
    using Geo = Dirac4TileGeom<N, NCW, 2>;
    if constexpr (Geo::OK) {
      if (gpart != nullptr && !Geo::CAN_GRAM) return -static_cast<int>(cudaErrorNotSupported);
      const int b = Geo::NTC / (Geo::G * geo.L0);  // rows per tile
      const long long ntiles = (row_end - row_begin + b - 1) / b;
      const int grid = clamp_grid(ntiles, Geo::CTAS_PER_SM * sms);   // persistent CTAs, as many as are resident
      if (!second)
        dirac4_tile_kernel<N, NCW, 2, false, false><<<grid, Geo::NT, Geo::SMEM_BYTES, st>>>(in, p0, out, Ut, geo, row_begin, row_end,
                                                                                            b, m2, sigma, nullptr, ctrl);
      else if (gpart == nullptr)
        dirac4_tile_kernel<N, NCW, 2, true, false><<<grid, Geo::NT, Geo::SMEM_BYTES, st>>>(in, p0, out, Ut, geo, row_begin, row_end,
                                                                                           b, m2, sigma, nullptr, ctrl);
      else if constexpr (Geo::CAN_GRAM)
        dirac4_tile_kernel<N, NCW, 2, true, true><<<grid, Geo::NT, Geo::SMEM_BYTES, st>>>(in, p0, out, Ut, geo, row_begin, row_end,
                                                                                          b, m2, sigma, gpart, ctrl);
      return grid;
    }
    return -static_cast<int>(cudaErrorNotSupported);
  }
  template <int NCW>
  static void dirac4_tile_prepare() {
    using Geo = Dirac4TileGeom<N, NCW, 2>;
    if constexpr (Geo::OK) {
      cudaFuncSetAttribute(dirac4_tile_kernel<N, NCW, 2, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Geo::SMEM_BYTES);
      cudaFuncSetAttribute(dirac4_tile_kernel<N, NCW, 2, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Geo::SMEM_BYTES);
      if constexpr (Geo::CAN_GRAM)
        cudaFuncSetAttribute(dirac4_tile_kernel<N, NCW, 2, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Geo::SMEM_BYTES);
    }
  }
  template <int NCW>
  static bool dirac4_fits(int L0) {  // a row of L0 sites fits the tile of this shape
    using Geo = Dirac4TileGeom<N, NCW, 2>;
    return Geo::OK && static_cast<long long>(L0) * Geo::G <= Geo::NTC;
  }
  static int dirac4_tile(cudaStream_t st, const cd* in, const cd* p0, cd* out, const cd* Ut, const Lattice4* lat,
                         long long site_stride_mu, long long row_begin, long long row_end, double m2, double sigma, int second,
                         cd* gpart, const Ctrl* ctrl, int sms, int* launches) {
    if (row_end <= row_begin) return 0;
    if (gpart != nullptr && !second) return -static_cast<int>(cudaErrorNotSupported);
    const int L0 = lat->L0;
    // BCG_DIRAC4_WARPS = 2 / 3 / 4 / 6 / 8 / 12 / 16 / 24 forces a shape (if the row fits it)
    static const int forced = [] { const char* e = std::getenv("BCG_DIRAC4_WARPS"); return e ? std::atoi(e) : 0; }();
    const int cand[8] = {2, 3, 4, 6, 8, 12, 16, 24};
    const bool fits[8] = {dirac4_fits<2>(L0), dirac4_fits<3>(L0), dirac4_fits<4>(L0), dirac4_fits<6>(L0),
                          dirac4_fits<8>(L0), dirac4_fits<12>(L0), dirac4_fits<16>(L0), dirac4_fits<24>(L0)};
    int best = -1;
    for (int i = 7; i >= 0; --i)
      if (fits[i]) best = i;
    for (int i = 0; i < 8; ++i)
      if (forced == cand[i] && fits[i]) best = i;
    if (best < 0) return -static_cast<int>(cudaErrorNotSupported);
    prepare(sms);
    const Rows4 geo = {lat->L0, lat->L1, lat->L2, lat->L3, site_stride_mu};
    int grid = 0;
#define BCG_D4_CASE(W) case W: grid = dirac4_tile_launch<W>(st, in, p0, out, Ut, geo, row_begin, row_end, m2, sigma, second, gpart, ctrl, sms); break;
    switch (cand[best]) {
      BCG_D4_CASE(2) BCG_D4_CASE(3) BCG_D4_CASE(4) BCG_D4_CASE(6) BCG_D4_CASE(8) BCG_D4_CASE(12) BCG_D4_CASE(16)
      default: grid = dirac4_tile_launch<24>(st, in, p0, out, Ut, geo, row_begin, row_end, m2, sigma, second, gpart, ctrl, sms); break;
    }
#undef BCG_D4_CASE
    if (grid < 0) return grid;  // this shape cannot fuse the Gram at this N: the caller runs the plain sweep + the Gram kernel
    if (launches) ++*launches;
    const int e = err();
    return e ? e : (gpart != nullptr ? grid : 0);
  }

  static int gram(cudaStream_t st, const cd* A, const cd* B, long long V, cd* gpart, const Ctrl* ctrl, int sms,
                  int* launches) {
    prepare(sms);
    const int grid = clamp_grid((V + GG::TS - 1) / GG::TS, caps().gram);
    gram_kernel<N, kNT><<<dim3(grid, GRAM_Y), kNT, GG::SMEM_BYTES, st>>>(A, B, V, gpart, ctrl);
    if (launches) ++*launches;
    int e = err();
    return e ? e : grid;
  }

  static int dirac(cudaStream_t st, const cd* in, cd* out, const cd* U, long long V, double m2, double sigma,
                   cd* gpart, const Ctrl* ctrl, int sms, int* launches, const GramPeers* peers, const HaloFold* hfp) {
    GramPeers pe;
    if (peers) pe = *peers; else std::memset(&pe, 0, sizeof pe);
    HaloFold hf;
    if (hfp) hf = *hfp; else std::memset(&hf, 0, sizeof hf);
    if (force_v1() & 1) return dirac_v1(st, in, out, U, V, m2, sigma, gpart, ctrl, sms, launches);
    if constexpr (CHAIN) {
      prepare(sms);
      const ChainPlan pl = chain_plan(V, sms);
      alignas(64) CUtensorMap tmP, tmO, tmU;
      int e = make_chain_map(&tmP, in, CW * 3 * N, CW * 3 * N, CGm::PP, pl.T, pl.nchains + 1, CK);
      if (!e) e = make_chain_map(&tmO, out, CW * 3 * N, CW * 3 * N, CGm::PP, pl.T, pl.nchains + 1, CK);
      if (!e) e = make_chain_map(&tmU, U - 9, (CW + 2) * 9, CW * 9, CGm::PU, pl.T, pl.nchains + 1, CK);
      if (e) return e;
      if (gpart != nullptr) {
        bool done = false;
        if constexpr (N % 4 == 0 && CK == 16) {
          if (gram_dmma() || !DFMA_GRAM) {
            launch_pdl(dirac_chain_kernel<N, CG_, CK, CW, 3>, pl.grid, CGm::NT, CGm::SMEM_BYTES, st, tmP, tmO, tmU, in, U, V,
                       pl.L, m2, sigma, gpart, ctrl, pe, hf);
            done = true;
          }
        }
        if constexpr (DFMA_GRAM) {
          if (!done)
            launch_pdl(dirac_chain_kernel<N, CG_, CK, CW, CGMODE>, pl.grid, CNT_G, CGm::SMEM_BYTES, st, tmP, tmO, tmU, in, U,
                       V, pl.L, m2, sigma, gpart, ctrl, pe, hf);
        }
      } else
        launch_pdl(dirac_chain_kernel<N, CG_, CK, CW, 0>, pl.grid, CGm::NT, CGm::SMEM_BYTES, st, tmP, tmO, tmU, in, U, V, pl.L,
                   m2, sigma, static_cast<cd*>(nullptr), ctrl, pe, hf);
      if (launches) ++*launches;
      e = err();
      return e ? e : (gpart != nullptr ? 1 : 0);  // the kernel leaves the fully reduced block in gpart[0]
    }
    return dirac_v1(st, in, out, U, V, m2, sigma, gpart, ctrl, sms, launches);
  }

  static int dirac_v1(cudaStream_t st, const cd* in, cd* out, const cd* U, long long V, double m2, double sigma,
                      cd* gpart, const Ctrl* ctrl, int sms, int* launches) {
    prepare(sms);
    const long long ntiles = (V + DG::TS - 1) / DG::TS;
    if (gpart != nullptr) {
      if constexpr (FUSED) {
        const int grid = clamp_grid(ntiles, caps().dirac_g);
        dirac_kernel<N, R, kNT, true><<<grid, kNT, DG::SMEM_BYTES, st>>>(in, out, U, V, m2, sigma, gpart, ctrl);
        if (launches) ++*launches;
        int e = err();
        return e ? e : grid;
      }
    }
    const int grid = clamp_grid(ntiles, caps().dirac);
    dirac_kernel<N, R, kNT, false><<<grid, kNT, DG::SMEM_BYTES, st>>>(in, out, U, V, m2, sigma, nullptr, ctrl);
    if (launches) ++*launches;
    int e = err();
    if (e) return e;
    if (gpart != nullptr) return gram(st, in, out, V, gpart, ctrl, sms, launches);
    return 0;
  }

  static constexpr bool AFOLD = APIPE && APG::FOLD_OK;
  static int axpy_gram_fold(cudaStream_t st, cd* Q, const cd* T, long long V, cd* gpart, const Ctrl* ctrl, int sms,
                            int* launches, const GramPeers* peers, cd* Qout, const AlphaFold* fold) {
    if (!AFOLD || fold == nullptr || !fold->on || gpart == nullptr || ctrl == nullptr || (force_v1() & 2))
      return -static_cast<int>(cudaErrorNotSupported);
    return axpy_gram_impl(st, Q, T, nullptr, V, gpart, ctrl, sms, launches, peers, Qout, fold);
  }
  static int axpy_gram(cudaStream_t st, cd* Q, const cd* T, const cd* M, long long V, cd* gpart,
                       const Ctrl* ctrl, int sms, int* launches, const GramPeers* peers, cd* Qout) {
    return axpy_gram_impl(st, Q, T, M, V, gpart, ctrl, sms, launches, peers, Qout, nullptr);
  }
  static int axpy_gram_impl(cudaStream_t st, cd* Q, const cd* T, const cd* M, long long V, cd* gpart,
                            const Ctrl* ctrl, int sms, int* launches, const GramPeers* peers, cd* Qout,
                            const AlphaFold* fold) {
    GramPeers pe;
    if (peers) pe = *peers; else std::memset(&pe, 0, sizeof pe);
    AlphaFold af;
    std::memset(&af, 0, sizeof af);
    if (fold) {
      af = *fold;
      pe.iter_from_b = 1;  // the A-step (which writes ctrl->iter) runs beside this kernel
    }
    if constexpr (!APIPE) {
      if (Qout != nullptr && Qout != Q) return -static_cast<int>(cudaErrorNotSupported);
    }
    if ((force_v1() & 2) && (Qout == nullptr || Qout == Q)) return axpy_gram_v1(st, Q, T, M, V, gpart, ctrl, sms, launches);
    if constexpr (APIPE) {
      prepare(sms);
      alignas(64) CUtensorMap tmQ, tmQo, tmT;
      int e = make_pair_map(&tmQ, Q, 3 * N, APG::PAIR, (V + 1) / 2, APIPE_TS / 2);
      if (!e) e = make_pair_map(&tmQo, Qout ? Qout : Q, 3 * N, APG::PAIR, (V + 1) / 2, APIPE_TS / 2);
      if (!e) e = make_pair_map(&tmT, T, 3 * N, APG::PAIR, (V + 1) / 2, APIPE_TS / 2);
      if (e) return e;
      const int grid = clamp_grid((V + APIPE_TS - 1) / APIPE_TS, sms);
      if (gpart != nullptr) {
        bool done = false;
        if constexpr (N % 4 == 0) {
          if (gram_dmma() || !DFMA_GRAM) {
            launch_pdl(axpy_pipe_kernel<N, APIPE_TS, 2>, grid, APG::NT, APG::SMEM_BYTES, st, tmQ, tmQo, tmT, M, V, gpart, ctrl, pe, axpy_reverse(), af);
            done = true;
          }
        }
        if constexpr (DFMA_GRAM) {
          if (!done)
            launch_pdl(axpy_pipe_kernel<N, APIPE_TS, 1>, grid, APG::NT, APG::SMEM_BYTES, st, tmQ, tmQo, tmT, M, V, gpart, ctrl, pe, axpy_reverse(), af);
        }
      } else {
        launch_pdl(axpy_pipe_kernel<N, APIPE_TS, 0>, grid, APG::NT, APG::SMEM_BYTES, st, tmQ, tmQo, tmT, M, V, static_cast<cd*>(nullptr), ctrl, pe, axpy_reverse(), af);
      }
      if (launches) ++*launches;
      e = err();
      return e ? e : (gpart != nullptr ? 1 : 0);
    }
    return axpy_gram_v1(st, Q, T, M, V, gpart, ctrl, sms, launches);
  }

  static int axpy_gram_v1(cudaStream_t st, cd* Q, const cd* T, const cd* M, long long V, cd* gpart,
                          const Ctrl* ctrl, int sms, int* launches) {
    prepare(sms);
    const long long ntiles = (3 * V + kNT - 1) / kNT;
    if (gpart != nullptr) {
      if constexpr (FUSED) {
        const int grid = clamp_grid(ntiles, caps().axpy_g);
        axpy_gram_kernel<N, kNT, true><<<grid, kNT, AG::SMEM_BYTES, st>>>(Q, T, M, V, gpart, ctrl);
        if (launches) ++*launches;
        int e = err();
        return e ? e : grid;
      }
    }
    const int grid = clamp_grid(ntiles, caps().axpy);
    axpy_gram_kernel<N, kNT, false><<<grid, kNT, sizeof(cd) * N * N, st>>>(Q, T, M, V, nullptr, ctrl);
    if (launches) ++*launches;
    int e = err();
    if (e) return e;
    if (gpart != nullptr) return gram(st, Q, Q, V, gpart, ctrl, sms, launches);
    return 0;
  }

  static int rescale_add(cudaStream_t st, cd* dst, const cd* L, const cd* src, double r, long long V, int sms,
                         int* launches) {
    prepare(sms);
    rescale_add_kernel<N, kNT><<<clamp_grid((3 * V + kNT - 1) / kNT, caps().rescale), kNT, 0, st>>>(dst, L, src, r, V);
    if (launches) ++*launches;
    return err();
  }

  static int trsm(cudaStream_t st, cd* Q, const cd* Rm, long long V, const Ctrl* ctrl, int sms, int* launches) {
    prepare(sms);
    trsm_kernel<N, kNT><<<clamp_grid((3 * V + kNT - 1) / kNT, caps().trsm), kNT, 0, st>>>(Q, Rm, V, ctrl);
    if (launches) ++*launches;
    return err();
  }

  static int shift_update(cudaStream_t st, cd* Q, const ShiftPtrs* fp, const cd* Rm, const cd* A, const cd* B,
                          long long V, int do_backsub, int n_active_fixed, const Ctrl* ctrl, int sms,
                          int* launches) {
    prepare(sms);
    if (force_v1() & 4) return shift_update_direct(st, Q, fp, Rm, A, B, V, do_backsub, n_active_fixed, ctrl, sms, launches);
    if constexpr (PIPE_OK) {
      const int grid = clamp_grid((V + SHIFT_TS - 1) / SHIFT_TS, caps().pipe);
      // tensor maps of every field the launch may touch (n_active is only known on the device)
      alignas(64) ShiftMaps maps;
      const long long npairs = (V + 1) / 2;
      int e = make_pair_map(&maps.Q, Q, 3 * N, SG::PAIR, npairs, SHIFT_TS / 2);
      for (int s = 0; s < kMaxShifts && !e; ++s) {
        if (fp->P[s] == nullptr || fp->X[s] == nullptr) {
          maps.P[s] = maps.Q;  // never used: the device loop stops at n_active
          maps.X[s] = maps.Q;
          continue;
        }
        e = make_pair_map(&maps.P[s], fp->P[s], 3 * N, SG::PAIR, npairs, SHIFT_TS / 2);
        if (!e) e = make_pair_map(&maps.X[s], fp->X[s], 3 * N, SG::PAIR, npairs, SHIFT_TS / 2);
      }
      if (e) return e;
      shift_pipe_kernel<N, SHIFT_TS><<<grid, SG::NT, SG::SMEM_BYTES, st>>>(maps, Rm, A, B, V, do_backsub,
                                                                           n_active_fixed, ctrl);
      if (launches) ++*launches;
      return err();
    }
    return shift_update_direct(st, Q, fp, Rm, A, B, V, do_backsub, n_active_fixed, ctrl, sms, launches);
  }

  static int shift_update_pair(cudaStream_t st, cd* Q, cd* Qprev, const ShiftPtrs* fp, const cd* Rm, const cd* A_odd,
                               const cd* B_odd, const cd* A_even, const cd* B_even, long long V, const Ctrl* ctrl,
                               int sms, int* launches) {
    if constexpr (PAIR_OK) {
      prepare(sms);
      const int grid = clamp_grid((V + SHIFT_TS - 1) / SHIFT_TS, caps().pair);
      alignas(64) ShiftPairMaps maps;
      const long long npairs = (V + 1) / 2;
      int e = make_pair_map(&maps.Q, Q, 3 * N, SG::PAIR, npairs, SHIFT_TS / 2);
      if (!e) e = make_pair_map(&maps.Qprev, Qprev, 3 * N, SG::PAIR, npairs, SHIFT_TS / 2);
      for (int s = 0; s < kMaxShifts && !e; ++s) {
        if (fp->P[s] == nullptr || fp->X[s] == nullptr) {
          maps.P[s] = maps.Q;  // never used: the device loop stops at the active count
          maps.X[s] = maps.Q;
          continue;
        }
        e = make_pair_map(&maps.P[s], fp->P[s], 3 * N, SG::PAIR, npairs, SHIFT_TS / 2);
        if (!e) e = make_pair_map(&maps.X[s], fp->X[s], 3 * N, SG::PAIR, npairs, SHIFT_TS / 2);
      }
      if (e) return e;
      shift_pair_kernel<N, SHIFT_TS><<<grid, SG::NT, SPG::SMEM_BYTES, st>>>(maps, Rm, A_odd, B_odd, A_even, B_even, V,
                                                                            ctrl);
      if (launches) ++*launches;
      return err();
    }
    return -static_cast<int>(cudaErrorNotSupported);
  }

  template <int TSX, int NSTX>
  static int launch_dmma(cudaStream_t st, int cap, cd* Q, cd* Qprev, const ShiftPtrs* fp, const cd* Rm, const cd* A_odd,
                         const cd* B_odd, const cd* A_even, const cd* B_even, long long V, const Ctrl* ctrl,
                         int* launches, int paired, cd* p0_halo, const HaloFold* hfp) {
    using G = ShiftDmmaGeom<N, TSX, NSTX>;
    HaloFold hf;
    if (hfp) hf = *hfp; else std::memset(&hf, 0, sizeof hf);
    const int grid = clamp_grid((V + TSX - 1) / TSX, cap);
    alignas(64) ShiftPairMaps maps;
    const long long npairs = (V + 1) / 2;
    int e = make_pair_map(&maps.Q, Q, 3 * N, G::PAIR, npairs, TSX / 2);
    if (!e) e = make_pair_map(&maps.Qprev, Qprev ? Qprev : Q, 3 * N, G::PAIR, npairs, TSX / 2);
    for (int s = 0; s < kMaxShifts && !e; ++s) {
      if (fp->P[s] == nullptr || fp->X[s] == nullptr) {
        maps.P[s] = maps.Q;  // never used: the device loop stops at the active count
        maps.X[s] = maps.Q;
        continue;
      }
      e = make_pair_map(&maps.P[s], fp->P[s], 3 * N, G::PAIR, npairs, TSX / 2);
      if (!e) e = make_pair_map(&maps.X[s], fp->X[s], 3 * N, G::PAIR, npairs, TSX / 2);
    }
    if (e) return e;
    launch_pdl(shift_dmma_kernel<N, TSX, NSTX>, grid, G::NT, G::SMEM_BYTES, st, maps, Rm, A_odd, B_odd, A_even, B_even, V, ctrl,
               paired, p0_halo, hf);
    if (launches) ++*launches;
    return err();
  }
  static int dmma_cfg() {
    const char* e = std::getenv("BCG_DMMA_CFG");  // read per launch: A/B runs in one process
    return e ? std::atoi(e) : 0;
  }
  static int shift_update_dmma(cudaStream_t st, cd* Q, cd* Qprev, const ShiftPtrs* fp, const cd* Rm, const cd* A_odd,
                               const cd* B_odd, const cd* A_even, const cd* B_even, long long V, const Ctrl* ctrl,
                               int sms, int* launches, int paired, cd* p0_halo, const HaloFold* hfp) {
    if constexpr (DMMA_OK) {
      prepare(sms);
      const int cfg = dmma_cfg();
      if constexpr (DMMA_CFG1)
        if (cfg == 1)
          return launch_dmma<32, 3>(st, caps().dmma1, Q, Qprev, fp, Rm, A_odd, B_odd, A_even, B_even, V, ctrl, launches, paired, p0_halo, hfp);
      if constexpr (DMMA_CFG2)
        if (cfg == 2)
          return launch_dmma<64, 2>(st, caps().dmma2, Q, Qprev, fp, Rm, A_odd, B_odd, A_even, B_even, V, ctrl, launches, paired, p0_halo, hfp);
      return launch_dmma<SHIFT_TS, 2>(st, caps().dmma, Q, Qprev, fp, Rm, A_odd, B_odd, A_even, B_even, V, ctrl, launches, paired, p0_halo, hfp);
    }
    return -static_cast<int>(cudaErrorNotSupported);
  }

  using SSG = ShiftStagGeom<DMMA_N ? N : 4, SHIFT_TS>;
  static constexpr bool STAG_OK = DMMA_OK && SSG::SMEM_BYTES <= 227 * 1024;
  static int shift_update_stag(cudaStream_t st, cd* const* Qring, int depth, int ring, int part, int slot, int max_ctas,
                               const ShiftPtrs* fp, const cd* Rm, const ShiftStagCoefs* coefs, long long V, const Ctrl* ctrl,
                               int sms, int* launches, cd* p0_halo, const HaloFold* hfp) {
    if constexpr (STAG_OK) {
      prepare(sms);
      HaloFold hf;
      if (hfp) hf = *hfp; else std::memset(&hf, 0, sizeof hf);
      int grid = clamp_grid((V + SHIFT_TS - 1) / SHIFT_TS, caps().stag);
      if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
      alignas(64) ShiftStagMaps maps;
      const long long npairs = (V + 1) / 2;
      int e = 0;
      for (int t = 0; t < kMaxDepth && !e; ++t)
        e = make_pair_map(&maps.Q[t], Qring[t < ring ? t : 0], 3 * N, SSG::PAIR, npairs, SHIFT_TS / 2);
      for (int s = 0; s < kMaxShifts && !e; ++s) {
        if (fp->P[s] == nullptr || fp->X[s] == nullptr) {
          maps.P[s] = maps.Q[0];  // never used: the device loop stops at the active count
          maps.X[s] = maps.Q[0];
          continue;
        }
        e = make_pair_map(&maps.P[s], fp->P[s], 3 * N, SSG::PAIR, npairs, SHIFT_TS / 2);
        if (!e) e = make_pair_map(&maps.X[s], fp->X[s], 3 * N, SSG::PAIR, npairs, SHIFT_TS / 2);
      }
      if (e) return e;
      launch_pdl(shift_stag_kernel<N, SHIFT_TS>, grid, SSG::NT, SSG::SMEM_BYTES, st, maps, Rm, *coefs, V, ctrl, depth, ring, part,
                 slot, p0_halo, hf);
      if (launches) ++*launches;
      if (part == 2) {
        bulk_mark_kernel<<<1, 1, 0, st>>>(const_cast<Ctrl*>(ctrl), slot);
        if (launches) ++*launches;
      }
      return err();
    }
    return -static_cast<int>(cudaErrorNotSupported);
  }

  static int shift_update_direct(cudaStream_t st, cd* Q, const ShiftPtrs* fp, const cd* Rm, const cd* A,
                                 const cd* B, long long V, int do_backsub, int n_active_fixed, const Ctrl* ctrl,
                                 int sms, int* launches) {
    prepare(sms);
    shift_update_kernel<N, kNT><<<clamp_grid((3 * V + kNT - 1) / kNT, caps().shift), kNT, SHIFT_SMEM, st>>>(
        Q, *fp, Rm, A, B, V, do_backsub, n_active_fixed, ctrl);
    if (launches) ++*launches;
    return err();
  }

  // see gram_group_reduce() for the layout of the partial-Gram buffer
  static int max_partials(int sms) { return (32 * sms > kGramCntOff + 32) ? 32 * sms : kGramCntOff + 32; }
};

template <int N>
const OpsTable* make_ops() {
  static const OpsTable t = {N,
                             Tune<N>::R,
                             DiracGeom<N, Tune<N>::R, kNT>::TS,
                             Ops<N>::FUSED ? 1 : 0,
                             &Ops<N>::dirac,
                             &Ops<N>::dirac_v1,
                             &Ops<N>::dirac4_sweep,
                             &Ops<N>::dirac4_tile,
                             &Ops<N>::gram,
                             &Ops<N>::axpy_gram,
                             Ops<N>::APIPE ? 1 : 0,
                             (Ops<N>::CHAIN && Ops<N>::APIPE) ? 1 : 0,
                             &Ops<N>::axpy_gram_v1,
                             &Ops<N>::rescale_add,
                             &Ops<N>::trsm,
                             &Ops<N>::shift_update,
                             &Ops<N>::shift_update_direct,
                             Ops<N>::PAIR_OK ? &Ops<N>::shift_update_pair : nullptr,
                             Ops<N>::DMMA_OK ? &Ops<N>::shift_update_dmma : nullptr,
                             Ops<N>::STAG_OK ? &Ops<N>::shift_update_stag : nullptr,
                             &Ops<N>::max_partials,
                             &Ops<N>::field_capacity,
                             &Ops<N>::prepare,
                             Ops<N>::AFOLD ? &Ops<N>::axpy_gram_fold : nullptr};
  return &t;
}
#endif  // BCG_N

}  // namespace bcg
