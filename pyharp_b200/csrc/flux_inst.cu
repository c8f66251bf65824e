// flux_inst.cu -- one instantiation of K1 (the DISORT flux solve kernel) for nn = B200_NN
// streams per hemisphere.  One thread = one (wave, column) solve; grid-stride over solves.
#ifndef B200_NN
#error "compile with -DB200_NN=<nstr/2>"
#endif
#ifndef B200_LOCKSTEP
#define B200_LOCKSTEP (B200_NN >= 6)
#endif
// parking the stack state (Ra, da) in shared memory instead of letting ptxas spill it: measured -5 % at nstr = 16
// (5.0 vs 5.3 M solves/s), so it is off; -DB200_SMEM_STACK=1 turns it on for experiments
#ifndef B200_SMEM_STACK
#define B200_SMEM_STACK 0
#endif
#include <cstdlib>

#include "flux_inst.cuh"
// team form of the adding phase of the layer-split solve (four lanes per solve): nstr = 16
// (and nstr = 8, where the split form serves batches far below one resident wave: pyharp's own 500 x 1 / G x 1 calls)
#ifndef B200_HAVE_TEAM
#define B200_HAVE_TEAM (B200_NN == 8 || B200_NN == 4)
#endif
#if B200_HAVE_TEAM
#include "adding_team.cuh"
#endif
#ifndef B200_TEAM_MINB
#define B200_TEAM_MINB 2  // resident CTAs per SM asked of ptxas for the team adding kernel
#endif

#ifndef B200_NN
#error "compile with -DB200_NN=<nstr/2>"
#endif

namespace b200 {
namespace {

__constant__ QuadTables c_qt;
constexpr int kNsave = B200_NN * B200_NN + 2 * B200_NN + 1;

// resident CTAs per SM asked of ptxas: the small stream counts are latency / bandwidth bound and want warps
// (128 registers per thread), the larger ones want every register they can get
#ifndef B200_MIN_BLOCKS
#define B200_MIN_BLOCKS 0
#endif
#if B200_MIN_BLOCKS == 0
#undef B200_MIN_BLOCKS
#if B200_NN <= 2
#define B200_MIN_BLOCKS 4
#elif B200_NN == 3
#define B200_MIN_BLOCKS 3  // 168 registers: no spills and 12 warps per SM (measured 165 M solves/s against 117 M at 128)
#else
#define B200_MIN_BLOCKS 1
#endif
#endif
#if defined(B200_MAXT)
#define B200_MAX_THREADS B200_MAXT
#elif B200_NN <= 3
#define B200_MAX_THREADS 128
#else
#define B200_MAX_THREADS 256
#endif
// NOMOM: instantiation for launches without phase-function moments (nprop <= 2; amars_sw, amars_lw, lbl shapes)
#if B200_NN <= 4
#define B200_HAVE_NOMOM 1
#else
#define B200_HAVE_NOMOM 0  // large stream counts always come with moments; one (large) kernel is enough
#endif
// BAND: fused band integration -- w_n F_n accumulated per column instead of per-solve flux rows (SolveParams::bflx_t,
// emit_flux in disort_device.cuh); BAND = 2 is the form for batches of fewer than 32 columns (nstr <= 8 only)
#if B200_NN <= 4
#define B200_HAVE_WARPRED 1
#else
#define B200_HAVE_WARPRED 0
#endif
template <bool NOMOM, int BAND>
__global__ void __launch_bounds__(B200_MAX_THREADS, B200_MIN_BLOCKS) disort_flux_kernel(SolveParams p) {
  const long nthreads = (long)gridDim.x * blockDim.x;
  const long tid = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long nsolve = p.nlocal;
  // position in the warp's ring of back-sweep records (stage = pos % stages, barrier parity = pos / stages)
  unsigned ring_pos = 0;
  if constexpr (TmaRing<kNsave>::kUsed) {
    using Ring = TmaRing<kNsave>;
    extern __shared__ __align__(128) double b200_smem[];
    if ((threadIdx.x & 31) == 0) {
      unsigned long long* bars = reinterpret_cast<unsigned long long*>(
          b200_smem + (size_t)(threadIdx.x >> 5) * Ring::kWarpDoubles + Ring::kStages * Ring::kRecDoubles);
      for (int k = 0; k < Ring::kStages; ++k) b200_mbar_init(bars + k, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
  }
  // CTA-uniform trip count: the solver contains CTA barriers (lockstep walk); threads beyond the end of the
  // batch shadow the last solve and write nothing
  BackPend<B200_NN> pend;  // carried back sweep of the previous solve (nstr <= 8)
  int it = 0;
  for (long base = (long)blockIdx.x * blockDim.x; base < nsolve; base += nthreads, ++it) {
    const long s = base + threadIdx.x;
    const bool active = s < nsolve;
    solve_flux_one<B200_NN, NOMOM, BAND>(p, c_qt, active ? s : nsolve - 1, tid, active, ring_pos, pend, it & 1);
  }
  if constexpr (kCarryBack<B200_NN>) {
    __syncwarp();
    drain_back_sweep<B200_NN, BAND>(p, pend, c_qt, tid, ring_pos);
  }
}

// ---------------------------------------------------------------------------
// Warp-per-solve kernel (see the block comment in disort_solver_body.inc): lanes take the layers of one solve up to the
// layer operators, lane 0 runs the adding recursion and the back sweep.  Scratch per warp: nlyr hand-over records +
// nlyr back-sweep records (global memory, L1/L2-resident: written and read by the same warp within microseconds).
// ---------------------------------------------------------------------------
#if B200_NN <= 4
#define B200_HAVE_WARP 1
#else
#define B200_HAVE_WARP 0
#endif
#if B200_HAVE_WARP
template <bool NOMOM, int BAND>
__global__ void __launch_bounds__(128, 1) disort_flux_warp_kernel(SolveParams p, double* scratch) {
  constexpr int NN = B200_NN;
  constexpr int NOPS = kOpsDoubles<NN>;
  const int lane = (int)(threadIdx.x & 31);
  const long warp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((long)gridDim.x * blockDim.x) >> 5;
  const int nlyr = p.nlyr;
  double* opsbuf = scratch + warp * ((long)nlyr * (NOPS + kNsave));
  double* recbuf = opsbuf + (long)nlyr * NOPS;
  for (long s = warp; s < p.nlocal; s += nwarps) {
    SolveCommon<NN> cm;
    solve_common_u<NN>(p, c_qt, s, cm);
    const double* pr = p.prop + s * nlyr * p.nprop;
    int ncut = nlyr, st = cm.st;
    double carry_t = 0.0, carry_a = 0.0;
    bool cut = false;
    for (int base = 0; base < nlyr && !cut; base += 32) {
      const int lc = base + lane;
      double dtp = 0.0, dabs = 0.0;
      if (lc < nlyr) layer_depths_u<NN, NOMOM>(p, pr, lc, dtp, dabs);
      double it = dtp, ia = dabs;  // inclusive prefix sums over the 32 layers of this round
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const double a = __shfl_up_sync(0xffffffffu, it, o), b = __shfl_up_sync(0xffffffffu, ia, o);
        if (lane >= o) { it += a; ia += b; }
      }
      const double taup = carry_t + (it - dtp), abefore = carry_a + (ia - dabs);
      // c_disort_set: the medium is cut at the first layer above which the absorption optical depth reaches 10
      const unsigned m = __ballot_sync(0xffffffffu, !cm.planck && nlyr > 1 && lc < nlyr && abefore >= 10.0);
      if (m) { ncut = base + __ffs(m) - 1; cut = true; }
      if (lc < min(ncut, nlyr)) st = status_merge_u(st, layer_ops_u<NN, NOMOM>(p, c_qt, cm, pr, lc, taup, opsbuf + (long)lc * NOPS));
      carry_t += __shfl_sync(0xffffffffu, it, 31);
      carry_a += __shfl_sync(0xffffffffu, ia, 31);
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) st = status_merge_u(st, __shfl_xor_sync(0xffffffffu, st, o));
    const bool lyrcut = cut || (!cm.planck && nlyr > 1 && carry_a >= 10.0);
    __syncwarp();  // the hand-over records of all lanes are visible to lane 0
    if (lane == 0) adding_sweep_u<NN, BAND>(p, c_qt, cm, s, pr, ncut, lyrcut, st, opsbuf, recbuf);
    __syncwarp();  // ... and consumed before the next solve overwrites them
  }
}

long warp_scratch(int nlyr) { return (long)nlyr * (kOpsDoubles<B200_NN> + kNsave); }

cudaError_t warp_occupancy(int block, int* blocks_per_sm) {
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, disort_flux_warp_kernel<false, 0>, block, 0);
}

cudaError_t launch_warp(const SolveParams& p, unsigned grid, int block, cudaStream_t stream, double* scratch) {
  const bool band = p.bflx_t != nullptr;
#if B200_HAVE_NOMOM
  if (p.nprop <= 2) {
    if (band) disort_flux_warp_kernel<true, 1><<<grid, block, 0, stream>>>(p, scratch);
    else disort_flux_warp_kernel<true, 0><<<grid, block, 0, stream>>>(p, scratch);
  } else
#endif
  {
    if (band) disort_flux_warp_kernel<false, 1><<<grid, block, 0, stream>>>(p, scratch);
    else disort_flux_warp_kernel<false, 0><<<grid, block, 0, stream>>>(p, scratch);
  }
  return cudaGetLastError();
}
#endif

// ---------------------------------------------------------------------------
// Layer-split form (nstr >= 12).  At these stream counts the per-solve state of the fused kernel (~400 doubles at
// nstr = 16) does not fit the register file: every thread carries a 3.5 KB spill frame, the frames of all resident
// threads (~130 MB) overflow the L2 and the spills travel to HBM and back (ncu: 323 KB of local-memory traffic per
// solve, long_scoreboard 3.6 stalls per issue, FP64 pipe 27 % busy).  The layers of a column are independent up to
// the adding recursion, so the solve is cut where its live state is smallest:
//   split_prefix_kernel   thread per solve: delta-M scaled optical depth at every layer top, absorption cut-off level
//   split_layer_kernel    thread per (solve, layer): eigenproblem, particular solutions, R / T, layer sources
//                         (disort_layer_ops.inc) -> hand-over record [warp block][layer][element][lane], 90 doubles at
//                         nstr = 16; no stack operator, no adding-step matrices among its live values
//   split_adding_kernel   thread per solve: adding recursion, bottom boundary, back sweep (disort_adding_step.inc);
//                         none of the eigen-work among its live values
// Both phases execute the statements of the fused kernel (the fragments are shared), so results agree with it to
// rounding (the optical depth at a layer top comes from the same running sum).  The hand-over costs 2 x 720 B of HBM
// traffic per layer and solve -- less than the spill traffic it removes -- and the layer phase has nlyr times the
// parallelism of the batch, so small batches fill the GPU as well.
// ---------------------------------------------------------------------------
#if B200_NN >= 6 || B200_NN == 4 || B200_NN == 2
#define B200_HAVE_SPLIT 1
#else
#define B200_HAVE_SPLIT 0
#endif
// thread form of the adding phase (the team form replaces it where it exists; kept at nstr >= 12 for comparison)
#define B200_HAVE_SPLIT_THREAD (B200_NN >= 6 || B200_NN == 2)
#if B200_HAVE_SPLIT
// Legendre polynomials at the beam angle: 1 = read from the prefix pass's buffer where they are used, 0 = recomputed per item
#ifndef B200_SPLIT_Y0BUF
#define B200_SPLIT_Y0BUF 1
#endif
// threads per CTA of the layer kernel (the CTA walks its items in lockstep)
#ifndef B200_SPLIT_BLOCK_A
#define B200_SPLIT_BLOCK_A 128
#endif
// L2 prefetch of the next item's inputs in the layer kernel: measured 2 % SLOWER (profiles/r02/sweeps.md, e44 -- the other
// resident CTA already covers the item-start loads), so it is off
#ifndef B200_SPLIT_PREFETCH
#define B200_SPLIT_PREFETCH 0
#endif
__device__ __forceinline__ void b200_prefetch_l2(const void* ptr) { asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr)); }
#ifndef B200_SPLIT_MINB_A
#define B200_SPLIT_MINB_A 1
#endif
#ifndef B200_SPLIT_MINB_B
#define B200_SPLIT_MINB_B 1
#endif
constexpr int kSplitNoStatus = 0x7fffffff;

template <bool NOMOM>
__global__ void __launch_bounds__(128) split_prefix_kernel(SolveParams p, long c0, long cn, long cs, double* taup, double* y0b,
                                                           int* cutv, int* stl) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cn) return;
  const long s = c0 + i;
  const int nlyr = p.nlyr;
  {  // Legendre polynomials at the beam angle (c_legendre_poly, m = 0): once per solve, not once per layer
    SolveCommon<B200_NN> cm;
    solve_common_any<B200_NN>(p, c_qt, s, cm, false, true);
#pragma unroll
    for (int l = 0; l < 2 * B200_NN; ++l) y0b[(long)l * cs + i] = cm.y0[l];
  }
  const bool planck = (p.flags & F_PLANCK) != 0;
  const double* pr = p.prop + s * nlyr * p.nprop;
  double tp = 0.0, ab = 0.0;
  int ncut = nlyr;
  bool cut = false;
  for (int lc = 0; lc < nlyr; ++lc) {
    // c_disort_set: the medium is cut at the first layer above which the absorption optical depth reaches 10
    if (!cut && !planck && nlyr > 1 && ab >= 10.0) { ncut = lc; cut = true; }
    taup[(long)lc * cs + i] = tp;
    double dtp, dabs;
    layer_depths_any<B200_NN, NOMOM>(p, pr, lc, dtp, dabs);
    tp += dtp;
    ab += dabs;
  }
  const bool lyrcut = cut || (!planck && nlyr > 1 && ab >= 10.0);
  cutv[i] = ncut | (lyrcut ? 0x10000 : 0);
  stl[i] = kSplitNoStatus;
}

// (gw = solves per group of the hand-over buffer: kWsLanes for the thread form of the adding phase, kGrp for the team form)
template <bool NOMOM>
__global__ void __launch_bounds__(B200_SPLIT_BLOCK_A, B200_SPLIT_MINB_A) split_layer_kernel(SolveParams p, long c0, long cn, long cs,
                                                                              const double* taup, const double* y0b,
                                                                              const int* cutv, double* ops, int* stl, int gw) {
  constexpr int NN = B200_NN, NOPS = kOpsDoubles<NN>;
  const int nlyr = p.nlyr;
  const long items = cn * nlyr, nthreads = (long)gridDim.x * blockDim.x;
  // item t = (layer t / cn, solve t % cn): one division per thread, then both indices advance with the grid stride
  long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  int lc = (int)(t / cn);
  long i = t - (long)lc * cn;
  const int dl = (int)(nthreads / cn);
  const long di = nthreads - (long)dl * cn;
  for (long base = (long)blockIdx.x * blockDim.x; base < items; base += nthreads, t += nthreads) {
    B200_BLOCK_SYNC();  // the CTA's warps start every item together (instruction-cache sharing, as in the fused kernel)
    const bool in = t < items;
    const int lcc = in ? lc : nlyr - 1;  // threads beyond the end shadow the last item (the Jacobi sweeps vote per warp)
    const long ic = in ? i : cn - 1;
    lc += dl;
    i += di;
    if (i >= cn) { i -= cn; ++lc; }
#if B200_SPLIT_PREFETCH
    // the inputs of this thread's NEXT item are pulled into the L2 behind this item's arithmetic: every item starts with a
    // dozen dependent-free global loads (cut-off level, optical depth at the layer top, boundary fields, Legendre
    // polynomials, the layer's prop record), and in lockstep all warps of the CTA wait for them at the same time
    if (t + nthreads < items) {
      const long sn = c0 + i;
      b200_prefetch_l2(cutv + i);
      b200_prefetch_l2(taup + (long)lc * cs + i);
      if (p.umu0) b200_prefetch_l2(p.umu0 + (p.bc_bcast ? 0 : sn));
      if (p.fbeam) b200_prefetch_l2(p.fbeam + (p.bc_bcast ? 0 : sn));
      if (p.albedo) b200_prefetch_l2(p.albedo + (p.bc_bcast ? 0 : sn));
      const double* qn = p.prop + sn * nlyr * p.nprop + (long)(nlyr - 1 - lc) * p.nprop;
      b200_prefetch_l2(qn);
      if (p.nprop > 16) b200_prefetch_l2(qn + 16);
      if (B200_SPLIT_Y0BUF && p.fbeam) {
#pragma unroll
        for (int l = 0; l < 2 * NN; ++l) b200_prefetch_l2(y0b + (long)l * cs + i);
      }
    }
#endif
    const bool active = in && lcc < (cutv[ic] & 0xffff);  // layers below the cut-off are not examined
    if (!__any_sync(0xffffffffu, active)) continue;
    const long s = c0 + ic;
    SolveCommon<NN> cm;
    solve_common_any<NN>(p, c_qt, s, cm, false, !B200_SPLIT_Y0BUF);
    const double* pr = p.prop + s * nlyr * p.nprop;
    double* o = ops + (ic / gw) * ((long)nlyr * NOPS * gw) + (long)lcc * NOPS * gw + (ic % gw);
    const int st = layer_ops_any<NN, NOMOM>(p, c_qt, cm, pr, lcc, taup[(long)lcc * cs + ic], o, gw,
                                            B200_SPLIT_Y0BUF ? y0b + ic : nullptr, B200_SPLIT_Y0BUF ? cs : 1);
    if (active && st != ST_OK) atomicMin(stl + ic, st);  // (bad input < eigen failure < singular: status_merge's order)
  }
}

#if B200_HAVE_SPLIT_THREAD
template <int BAND>
__global__ void __launch_bounds__(128, B200_SPLIT_MINB_B) split_adding_kernel(SolveParams p, long c0, long cn, const int* cutv,
                                                                               const int* stl, const double* ops, double* rec) {
  constexpr int NN = B200_NN, NOPS = kOpsDoubles<NN>;
  const long ii = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = ii < cn;
  if (BAND != kBandWarpRed && !active) return;  // (warp-reduced band integration: every lane of the warp stays)
  const long i = active ? ii : cn - 1;
  const long s = c0 + i;
  const int nlyr = p.nlyr;
  SolveCommon<NN> cm;
  solve_common_any<NN>(p, c_qt, s, cm, true, false);
  const double* pr = p.prop + s * nlyr * p.nprop;
  const int cv = cutv[i], sl = stl[i];
  const int st = status_merge_u(cm.st, sl == kSplitNoStatus ? (int)ST_OK : sl);
  const long blk = i / kWsLanes, lane = i % kWsLanes;
  adding_sweep_any<NN, BAND>(p, c_qt, cm, s, pr, cv & 0xffff, (cv >> 16) != 0, st, ops + blk * ((long)nlyr * NOPS * kWsLanes) + lane,
                             rec + blk * ((long)nlyr * kNsave * kWsLanes) + lane, kWsLanes, lane, active);
}
#endif
#if B200_HAVE_TEAM
template <int BAND>
__global__ void __launch_bounds__(128, B200_TEAM_MINB) team_adding_kernel(SolveParams p, long c0, long cn, const int* cutv, const int* stl,
                                                             const double* ops, double* rec) {
  extern __shared__ __align__(128) double b200_smem[];
  team_adding_sweep<B200_NN, BAND>(p, c_qt, c0, cn, cutv, stl, ops, rec, b200_smem);
}
#endif

long split_round(long chunk) { return (chunk + kWsLanes - 1) / kWsLanes * kWsLanes; }

size_t split_scratch(int nlyr, long chunk) {
  const long cs = split_round(chunk);
  return (size_t)cs * (nlyr * (kOpsDoubles<B200_NN> + kNsave + 1) + 2 * B200_NN) * sizeof(double) + (size_t)cs * 2 * sizeof(int);
}

cudaError_t launch_split(const SolveParams& p, long chunk, int sm_count, cudaStream_t stream, void* scratch) {
  const long cs = split_round(chunk);
  const int nlyr = p.nlyr;
  double* taup = static_cast<double*>(scratch);
  double* y0b = taup + (size_t)cs * nlyr;
  double* ops = y0b + (size_t)cs * 2 * B200_NN;
  double* rec = ops + (size_t)cs * nlyr * kOpsDoubles<B200_NN>;
  int* cutv = reinterpret_cast<int*>(rec + (size_t)cs * nlyr * kNsave);
  int* stl = cutv + cs;
  // (experiment build -DB200_SMEM_Y=1: the Jacobi eigenvector accumulator of every thread in shared memory)
  constexpr size_t kLayerSmem = B200_SMEM_Y ? (size_t)B200_NN * B200_NN * B200_SPLIT_BLOCK_A * sizeof(double) : 0;
  static int occ_a = 0;
  if (occ_a == 0) {
    cudaError_t e = cudaSuccess;
    if (kLayerSmem > 48 * 1024) e = cudaFuncSetAttribute(split_layer_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLayerSmem);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_a, split_layer_kernel<false>, B200_SPLIT_BLOCK_A, kLayerSmem);
    if (e != cudaSuccess) return e;
    if (occ_a < 1) occ_a = 1;
  }
  const bool band = p.bflx_t != nullptr;
  // adding phase: team form (four lanes per solve) where it is instantiated; B200DISORT_TEAM=0 selects the thread form
  static const bool team = B200_HAVE_TEAM && !(B200_HAVE_SPLIT_THREAD && std::getenv("B200DISORT_TEAM") && std::atoi(std::getenv("B200DISORT_TEAM")) == 0);
  const bool nomom = B200_HAVE_NOMOM && p.nprop <= 2;
#if B200_HAVE_TEAM
  constexpr size_t kTeamSmem = (size_t)4 * TeamLayout<B200_NN>::kWarpDoubles * sizeof(double);
  // (function attributes are per device: a C++ host that drives every GPU of the box from one process reaches this with
  // each of them current)
  static bool attr_set[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
  if (team && !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(team_adding_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTeamSmem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(team_adding_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTeamSmem);
    if (e != cudaSuccess) return e;
    attr_set[dev] = true;
  }
#endif
  const int gw = team ? 32 / 4 : kWsLanes;
  for (long c0 = 0; c0 < p.nlocal; c0 += cs) {
    const long cn = (p.nlocal - c0 < cs) ? p.nlocal - c0 : cs;
    const unsigned gs = (unsigned)((cn + 127) / 128);
    const long ga = (cn * nlyr + B200_SPLIT_BLOCK_A - 1) / B200_SPLIT_BLOCK_A, cap = (long)sm_count * occ_a;
    const unsigned gl = (unsigned)(ga < cap ? ga : cap);
#if B200_HAVE_NOMOM
    if (nomom) {
      split_prefix_kernel<true><<<gs, 128, 0, stream>>>(p, c0, cn, cs, taup, y0b, cutv, stl);
      split_layer_kernel<true><<<gl, B200_SPLIT_BLOCK_A, kLayerSmem, stream>>>(p, c0, cn, cs, taup, y0b, cutv, ops, stl, gw);
    } else
#endif
    {
      split_prefix_kernel<false><<<gs, 128, 0, stream>>>(p, c0, cn, cs, taup, y0b, cutv, stl);
      split_layer_kernel<false><<<gl, B200_SPLIT_BLOCK_A, kLayerSmem, stream>>>(p, c0, cn, cs, taup, y0b, cutv, ops, stl, gw);
    }
    (void)nomom;
#if B200_HAVE_TEAM
    if (team) {
      const unsigned gt = (unsigned)((cn + 31) / 32);  // 4 warps x 8 solves per CTA
      if (band) team_adding_kernel<1><<<gt, 128, kTeamSmem, stream>>>(p, c0, cn, cutv, stl, ops, rec);
      else team_adding_kernel<0><<<gt, 128, kTeamSmem, stream>>>(p, c0, cn, cutv, stl, ops, rec);
      continue;
    }
#endif
#if B200_HAVE_SPLIT_THREAD
#if B200_HAVE_WARPRED
    if (band && p.warp_red > 0) {  // fewer than 32 columns: warp-reduced band integration (emit_flux, BAND = 2)
      split_adding_kernel<2><<<gs, 128, 0, stream>>>(p, c0, cn, cutv, stl, ops, rec);
      continue;
    }
#endif
    if (band) split_adding_kernel<1><<<gs, 128, 0, stream>>>(p, c0, cn, cutv, stl, ops, rec);
    else split_adding_kernel<0><<<gs, 128, 0, stream>>>(p, c0, cn, cutv, stl, ops, rec);
#endif
  }
  return cudaGetLastError();
}
#endif

cudaError_t upload_tables(const QuadTables* qt) { return cudaMemcpyToSymbol(c_qt, qt, sizeof(QuadTables)); }

// dynamic shared memory per CTA: the parked stack state (Ra packed symmetric + da) of every thread
size_t smem_bytes(int block) {
  if (TmaRing<kNsave>::kUsed) return (size_t)(block / 32) * TmaRing<kNsave>::kWarpDoubles * sizeof(double);
#if B200_SMEM_Y
  if (B200_NN >= 6) return (size_t)(B200_NN * B200_NN) * block * sizeof(double);
#endif
#if B200_SMEM_STACK
  return (size_t)(B200_NN * (B200_NN + 1) / 2 + B200_NN) * block * sizeof(double);
#else
  (void)block;
  return 0;
#endif
}

cudaError_t occupancy(int block, int* blocks_per_sm) {
  const size_t sm = smem_bytes(block);
  if (sm > 48 * 1024) {
#define B200_SET_SMEM(NOMOM, BAND)                                                                                  \
  {                                                                                                                 \
    cudaError_t e = cudaFuncSetAttribute(disort_flux_kernel<NOMOM, BAND>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                         (int)sm);                                                                  \
    if (e != cudaSuccess) return e;                                                                                 \
  }
    B200_SET_SMEM(false, 0) B200_SET_SMEM(false, 1)
#if B200_HAVE_WARPRED
    B200_SET_SMEM(false, 2)
#endif
#if B200_HAVE_NOMOM
    B200_SET_SMEM(true, 0) B200_SET_SMEM(true, 1)
#if B200_HAVE_WARPRED
    B200_SET_SMEM(true, 2)
#endif
#endif
#undef B200_SET_SMEM
  }
  // (all instantiations run under the same __launch_bounds__, hence the same residency)
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, disort_flux_kernel<false, 0>, block, sm);
}

cudaError_t launch(const SolveParams& p, unsigned grid, int block, cudaStream_t stream) {
  // band mode: plain L2 reductions, or warp-reduced ones when the lanes of a warp share columns (p.warp_red > 0,
  // set by the host for ncol in {1, 2, 4, 8, 16}; only the nstr <= 8 kernels have that form)
  const int band = !p.bflx_t ? 0 : (B200_HAVE_WARPRED && p.warp_red > 0) ? 2 : 1;
  const size_t sm = smem_bytes(block);
#define B200_LAUNCH(NOMOM, BAND) disort_flux_kernel<NOMOM, BAND><<<grid, block, sm, stream>>>(p)
#if B200_HAVE_NOMOM
  if (p.nprop <= 2) {
    if (band == 0) B200_LAUNCH(true, 0);
    else if (band == 1) B200_LAUNCH(true, 1);
#if B200_HAVE_WARPRED
    else B200_LAUNCH(true, 2);
#endif
  } else
#endif
  {
    if (band == 0) B200_LAUNCH(false, 0);
    else if (band == 1) B200_LAUNCH(false, 1);
#if B200_HAVE_WARPRED
    else B200_LAUNCH(false, 2);
#endif
  }
#undef B200_LAUNCH
  return cudaGetLastError();
}

#if B200_HAVE_WARP && B200_HAVE_SPLIT
const FluxInst kInst = {B200_NN, B200_NN * B200_NN + 2 * B200_NN + 1, B200_MAX_THREADS, kCarryBack<B200_NN> ? 2 : 1, upload_tables, occupancy, launch,
                        warp_scratch, warp_occupancy, launch_warp, split_scratch, launch_split};
#elif B200_HAVE_WARP
const FluxInst kInst = {B200_NN, B200_NN * B200_NN + 2 * B200_NN + 1, B200_MAX_THREADS, kCarryBack<B200_NN> ? 2 : 1, upload_tables, occupancy, launch,
                        warp_scratch, warp_occupancy, launch_warp, nullptr, nullptr};
#elif B200_HAVE_SPLIT
const FluxInst kInst = {B200_NN, B200_NN * B200_NN + 2 * B200_NN + 1, B200_MAX_THREADS, kCarryBack<B200_NN> ? 2 : 1, upload_tables, occupancy, launch,
                        nullptr, nullptr, nullptr, split_scratch, launch_split};
#else
const FluxInst kInst = {B200_NN, B200_NN * B200_NN + 2 * B200_NN + 1, B200_MAX_THREADS, kCarryBack<B200_NN> ? 2 : 1, upload_tables, occupancy, launch,
                        nullptr, nullptr, nullptr, nullptr, nullptr};
#endif

}  // namespace

#define B200_CAT2(a, b) a##b
#define B200_CAT(a, b) B200_CAT2(a, b)
const FluxInst* B200_CAT(flux_inst_, B200_NN)() { return &kInst; }

}  // namespace b200
