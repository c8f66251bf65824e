// disort_device.cuh -- per-solve device math of the B200-native DISORT flux path.
//
// One CUDA thread owns one independent (wave, column) solve and walks the column
// top-down.  This is NOT the cdisort algorithm re-typed: the same discrete-ordinate
// boundary-value problem (c_disort: c_sol_eigen/c_asymmetric_matrix, c_upbeam,
// c_upisot, c_set_matrix, c_solve0/c_sgbco/c_sgbsl, c_fluxes -- SURVEY.md 8 a6-a15,
// called from pyharp at src/rtsolver/rt_solver_disort.cpp_:147) is solved with a
// formulation chosen for registers and FP64 FMA throughput:
//
//  * reduced nn x nn (nn = nstr/2) eigenproblem made SYMMETRIC by the stream scaling
//    u^ = sqrt(mu_i w_i) u  ->  Cholesky + cyclic Jacobi, no Hessenberg/QR, no branches
//    on data except convergence;
//  * beam and thermal particular solutions obtained from that eigen-decomposition
//    (no extra nstr x nstr LU as in c_upbeam / c_upisot);
//  * the (9nn-2) x (nstr*nlyr) banded LU of c_set_matrix/c_sgbco is replaced by the
//    mathematically equivalent interaction-principle (adding) sweep on nn x nn layer
//    reflection/transmission operators: O(nn^2) live state per solve, one forward
//    (top-down) sweep that stores (nn^2+2nn+1) doubles per level in a coalesced HBM
//    workspace and one backward sweep that emits the fluxes (c_fluxes).
//
// All of it is FP64.  The file is also compiled by g++ for the test-only host
// simulator (tests/hostsim), hence the B200_HD macros; the product library only
// ever instantiates it inside __global__ kernels.
#pragma once

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define B200_HD __host__ __device__ __forceinline__
#else
#define B200_HD inline
#endif

#ifndef B200_WS_STREAM
#define B200_WS_STREAM 1
#endif
#if defined(__CUDA_ARCH__)
// Reciprocal and reciprocal square root without the range check, branch and slow-path call of the CUDA math
// library versions: MUFU seed (2^-22) plus the same refinement, straight-line.  The library forms end every use
// in a conditional CALL, which splits the instruction stream into short basic blocks -- independent pivot /
// rotation chains could not be interleaved across them (ncu: `wait` was the top stall).  Valid for normal,
// positive-finite arguments only (rsqrt) / normal arguments (rcp); every call site guards its argument
// (pivots and eigenvalues are checked against kTiny, the other arguments are O(1) by construction).
#ifndef B200_FAST_RECIP
#define B200_FAST_RECIP 1
#endif
#if B200_FAST_RECIP
__device__ __forceinline__ double b200_fast_rcp(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  return fma(y, e, y);
}
__device__ __forceinline__ double b200_fast_rsqrt(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double e = fma(-x, y * y, 1.0);  // y (1 + e/2 + 3 e^2/8): third order, 2^-22 -> below 2^-53
  return fma(fma(e, 0.375, 0.5), y * e, y);
}
#define B200_RCP(x) b200_fast_rcp(x)
#define B200_RSQRT(x) b200_fast_rsqrt(x)
#else
#define B200_RCP(x) __drcp_rn(x)
#define B200_RSQRT(x) rsqrt(x)
#endif
// 1 - exp(-y) for y >= 0 with full RELATIVE accuracy (the -expm1(-y) of the math library, straight-line: range
// reduction -y = n ln 2 + r, Taylor polynomial of expm1(r) on |r| <= ln 2 / 2, 1 - 2^n (1 + expm1 r) in one FMA).
// The layer transmission functions k coth(k dtau'/2), k tanh(k dtau'/2) are formed from it; "1.0 - exp(-y)" has an
// absolute error of 1e-16, i.e. a relative one of 1e-16 / y, and the thermal source of a layer divides by dtau':
// across layers of dtau' ~ 1e-9 (upper atmosphere of a weak correlated-k g-point) that cost 7 digits of the flux.
__device__ __forceinline__ double b200_one_minus_exp_neg(double y) {
  const double x = fmax(-y, -700.0);  // exp(-700) = 1e-304: the result is 1 to double precision beyond
  const double n = rint(x * 1.4426950408889634);
  double r = fma(n, -6.93147180559945286e-01, x);
  r = fma(n, -2.31904681384629956e-17, r);
  double p = 1.0 / 87178291200.0;  // 1 / 14!
  p = fma(p, r, 1.0 / 6227020800.0);
  p = fma(p, r, 1.0 / 479001600.0);
  p = fma(p, r, 1.0 / 39916800.0);
  p = fma(p, r, 1.0 / 3628800.0);
  p = fma(p, r, 1.0 / 362880.0);
  p = fma(p, r, 1.0 / 40320.0);
  p = fma(p, r, 1.0 / 5040.0);
  p = fma(p, r, 1.0 / 720.0);
  p = fma(p, r, 1.0 / 120.0);
  p = fma(p, r, 1.0 / 24.0);
  p = fma(p, r, 1.0 / 6.0);
  p = fma(p, r, 0.5);
  const double q = fma(p * r, r, r);  // expm1(r)
  const double s = __hiloint2double(((int)n + 1023) << 20, 0);  // 2^n, n in [-1010, 0]
  return fma(-s, q, 1.0 - s);
}
#define B200_EM(y) b200_one_minus_exp_neg(y)
#define B200_PREFETCH(ptr) asm volatile("prefetch.global.L1 [%0];" ::"l"(ptr))
#define B200_HIWORD(x) __double2hiint(x)
// Streaming (evict-first) accesses for the back-sweep records, which are written once and read once.  At nstr >= 12 the kernel spills ~3 KB per thread, and the
// spill working set of all resident threads (~120 MB) only just fits the L2: keeping the streaming traffic from
// displacing it is worth +11 % on the GCM workload.  At nstr <= 8 there are no spills to protect and the hint costs
// 1 %, so plain accesses are used there (kStreamWs).  The same hint on the layer inputs and the flux rows changes
// nothing (profiles/r01, sweeps e15/e16).
#define B200_ST_STREAM(ptr, v) __stcs((ptr), (v))
#define B200_LD_STREAM(ptr) __ldcs(ptr)
// CTA lockstep (see jacobi_spd in disort_solver_body.inc); every thread of the CTA must reach these.
// Measured on B200: +32 % at nstr = 16 (instruction-fetch bound: 376 KB of unrolled code, I-cache hit rate 51 %),
// -3..6 % at nstr = 8, so flux_inst.cu enables it from nn = 6 up.
#if defined(B200_LOCKSTEP) && B200_LOCKSTEP
#define B200_BLOCK_SYNC() __syncthreads()
#define B200_BLOCK_ANY(pred) (__syncthreads_or(pred) != 0)
// Jacobi convergence vote in the lockstep build: per warp (B200_JACOBI_WARP_VOTE, the sweep body is small enough to
// stay in the instruction cache; the CTA re-aligns at the next layer) or per CTA
#ifndef B200_JACOBI_WARP_VOTE
#define B200_JACOBI_WARP_VOTE 1
#endif
#if B200_JACOBI_WARP_VOTE
#define B200_JACOBI_ANY(pred) (__any_sync(0xffffffffu, (pred)) != 0)
#else
#define B200_JACOBI_ANY(pred) B200_BLOCK_ANY(pred)
#endif
#else
#define B200_BLOCK_SYNC() ((void)0)
#define B200_BLOCK_ANY(pred) (pred)
#define B200_JACOBI_ANY(pred) (pred)
#endif
#else
#define B200_BLOCK_SYNC() ((void)0)
#define B200_BLOCK_ANY(pred) (pred)
#define B200_JACOBI_ANY(pred) (pred)
#define B200_PREFETCH(ptr) ((void)0)
#define B200_RCP(x) (1.0 / (x))
#define B200_RSQRT(x) (1.0 / sqrt(x))
#define B200_EM(y) (-expm1(-(y)))
#define B200_HIWORD(x) b200::host_hiword(x)
#define B200_ST_STREAM(ptr, v) (*(ptr) = (v))
#define B200_LD_STREAM(ptr) (*(ptr))
#include <string.h>
#include <algorithm>
namespace b200 {
inline int host_hiword(double x) {
  long long b;
  memcpy(&b, &x, 8);
  return (int)(b >> 32);
}
using std::max;
}  // namespace b200
#endif

namespace b200 {

constexpr int kMaxNN = 16;  // nstr <= 32
constexpr double kPi = 3.14159265358979323846;
constexpr double kTiny = 1.0e-290;  // pivots / eigenvalues at or below this are treated as zero (singular)

// status words (include/b200disort.h)
enum : int { ST_OK = 0, ST_BAD_INPUT = 1, ST_EIGEN_FAIL = 2, ST_SINGULAR = 3, ST_UNSUPPORTED = 4, ST_BEAM_ANGLE = 5 };

enum : unsigned { F_USRTAU = 1u, F_USRANG = 2u, F_LAMBER = 4u, F_PLANCK = 8u, F_ONLYFL = 16u, F_INTCOR = 32u, F_OLDINTCOR = 64u };

// Quadrature-dependent tables for azimuth mode 0 (built on the host, constant memory on
// the device).  Streams i = 0..nn-1 ascending in mu (double-Gauss on (0,1)).
struct QuadTables {
  double mu[kMaxNN];    // quadrature cosines
  double rmu[kMaxNN];   // 1/mu
  double wt[kMaxNN];    // weights
  double t[kMaxNN];     // sqrt(mu*w)  (stream scaling)
  double ehat[kMaxNN];  // sqrt(w/mu)
  double yhat[2 * kMaxNN][kMaxNN];  // sqrt(w_i/mu_i) * P_l(mu_i), l = 0..nstr-1
};

struct SolveParams {
  int nwave, ncol, nlyr, nprop, nstr;
  unsigned flags;
  const double* prop;  // (nwave,ncol,nlyr,nprop), layer 0 = surface
  const double *umu0, *fbeam, *albedo, *btemp, *ttemp, *temis, *fisot;  // (nwave,ncol) or null
  const double* pkag;        // (nband,ncol,nlyr+1) Planck emission per level, DISORT (top-down) order
  const int* wave2band;      // (nwave)
  const double* wave_lower;  // (nwave) or null
  const double* wave_upper;
  double* flx;     // (nwave,ncol,nlyr+1,2), level 0 = surface
  int32_t* status; // (nwave,ncol)
  long s0;         // batch index of this launch's first solve (chunked host path); prop/bc/flx/status are chunk-local
  long nlocal;     // solves in this launch
  // BAND instantiation (fused band integration, examples/amars_lw.cpp:84-88): instead of the per-spectral-point
  // flux rows the kernel accumulates w_n F_n into bflx_t[(row * 2 + c) * ncol + j] (row = pyharp level index, level
  // 0 = surface; coalesced across the lanes of a warp, which are neighbouring columns) with FP64 reductions in the
  // L2; flx is not written.  The host zeroes bflx_t before the launch.
  double* bflx_t;
  const double* weights;  // (nwave) band weights (BAND only)
  int bc_bcast;    // != 0: the boundary fields are (ncol) arrays shared by every spectral point
  int warp_red;    // BAND = 2: columns per warp (= ncol, a power of two <= 16), see emit_flux
  int nrep;        // BAND = 2: replicas of the band-flux buffer; CTA b accumulates into replica b % nrep, the host's
  long rep_stride; //   finishing kernel adds them (doubles per replica).  With one column every warp of the launch
                   //   reduces into the SAME 2 (nlyr + 1) addresses: unreplicated, the L2 serialises ~3000 reductions
                   //   per address and the fused form was 2x slower than writing the fluxes (10^5 x 1 line-by-line)
  double* ws;      // back-sweep workspace: [half][warp][level][elem][lane] (kWsLanes lanes per group)
  long ws_half;    // doubles per half (carried back sweep: two halves; otherwise unused)
  long ws_threads; // threads the workspace was sized for (a multiple of kWsLanes)
};

// ---------------------------------------------------------------------------
// c_planck_func1 (PLKAVG) on the device: identical series / cut-offs / legacy
// constants as the oracle; parity needs the same series, not a
// better Planck integral.
// ---------------------------------------------------------------------------
B200_HD double plkf(double x) { return x * x * x / (exp(x) - 1.0); }

B200_HD double plkavg(double wnumlo, double wnumhi, double t) {
  const double a1 = 1.0 / 3.0, a2 = -1.0 / 8.0, a3 = 1.0 / 60.0, a4 = -1.0 / 5040.0,
               a5 = 1.0 / 272160.0, a6 = -1.0 / 13305600.0;
  const double c2 = 1.438786, sigma = 5.67032e-8, vcut = 1.5;
  const double vcp[7] = {10.25, 5.7, 3.9, 2.9, 2.3, 1.9, 0.0};
  const double vmax = 709.782712893384;  // log(DBL_MAX)
  const double epsil = 2.220446049250313e-16;
  const double sigdpi = sigma / kPi;
  const double conc = 15.0 / (kPi * kPi * kPi * kPi);
  if (t < 1.0e-4) return 0.0;
  double v[2] = {c2 * wnumlo / t, c2 * wnumhi / t};
  if (v[0] > epsil && v[1] < vmax && (wnumhi - wnumlo) / wnumhi < 1.0e-2) {
    double hh = v[1] - v[0], oldval = 0.0, val = 0.0;
    double val0 = plkf(v[0]) + plkf(v[1]);
    for (int n = 1; n <= 10; ++n) {
      double del = hh / (2 * n);
      val = val0;
      for (int k = 1; k <= 2 * n - 1; ++k) val += 2 * (1 + k % 2) * plkf(v[0] + k * del);
      val = del / 3.0 * val;
      if (fabs((val - oldval) / val) <= 1.0e-6) break;
      oldval = val;
    }
    return sigdpi * t * t * t * t * conc * val;
  }
  double p[2] = {0, 0}, d[2] = {0, 0};
  int smallv = 0;
  for (int i = 0; i < 2; ++i) {
    if (v[i] < vcut) {
      ++smallv;
      double vsq = v[i] * v[i];
      p[i] = conc * vsq * v[i] * (a1 + v[i] * (a2 + v[i] * (a3 + vsq * (a4 + vsq * (a5 + vsq * a6)))));
    } else {
      int mmax = 0;
      do { ++mmax; } while (v[i] < vcp[mmax - 1]);
      double ex = exp(-v[i]), exm = 1.0, acc = 0.0;
      for (int m = 1; m <= mmax; ++m) {
        double mv = m * v[i];
        exm = ex * exm;
        acc += exm * (6.0 + mv * (6.0 + mv * (3.0 + mv))) / ((double)m * m * m * m);
      }
      d[i] = conc * acc;
    }
  }
  double ans;
  if (smallv == 2) ans = p[1] - p[0];
  else if (smallv == 1) ans = 1.0 - p[0] - d[1];
  else ans = d[0] - d[1];
  return sigdpi * t * t * t * t * ans;
}

// packed lower-triangular index (i >= j)
B200_HD constexpr int sym(int i, int j) { return i >= j ? i * (i + 1) / 2 + j : j * (j + 1) / 2 + i; }

// thread-private array in shared memory, element e of this thread at base[e * stride]
struct SmemArray {
  double* base;
  int stride;
  B200_HD double& operator[](int e) const { return base[(long)e * stride]; }
};

// reference to `a` when B, else to `b` (storage-sharing switch in the solver body)
template <bool B, typename T>
B200_HD T& pick_ref(T& a, T& b) {
  if constexpr (B) return a;
  else return b;
}

template <int NN>
constexpr bool kStreamWs = B200_WS_STREAM && NN >= 6;
template <bool STREAM>
B200_HD void st_ws(double* ptr, double v) {
  if constexpr (STREAM) B200_ST_STREAM(ptr, v);
  else *ptr = v;
}
template <bool STREAM>
B200_HD double ld_ws(const double* ptr) {
  if constexpr (STREAM) return B200_LD_STREAM(ptr);
  else return *ptr;
}

constexpr int kMaxSweeps = 16;
// largest nn whose scratch matrices share storage explicitly (disort_solver_body.inc, kShare)
#ifndef B200_SHARE_MAX_NN
#define B200_SHARE_MAX_NN 4
#endif
// largest nn whose phase-function moments are loaded a layer ahead into registers (disort_solver_body.inc, PIPE_PM)
#ifndef B200_PIPE_PM_MAX_NN
#define B200_PIPE_PM_MAX_NN 4
#endif
// smallest nn at which Z = L Y is not formed (disort_layer_ops.inc, kNoZT).  Measured at nstr = 16 (profiles/r02/sweeps.md,
// e37): 4 % SLOWER in the layer kernel although it executes 3 % fewer flops -- L and Y stay live through the particular
// solutions and ptxas spills more (3.5 KB of spill stores per layer against 3.0 KB) -- so it is off
#ifndef B200_NOZT_MIN_NN
#define B200_NOZT_MIN_NN 99
#endif
#ifndef B200_BS_DEPTH
#define B200_BS_DEPTH 3
#endif
#ifndef B200_OPT_SPD2
#define B200_OPT_SPD2 1
#endif

// Back-sweep workspace layout [warp][level][element][lane] (see solve_flux_one): lanes per group.  The host
// simulator runs one solve at a time (one lane).
#if defined(__CUDA_ARCH__) || defined(__CUDACC__)
constexpr int kWsLanes = 32;
#else
constexpr int kWsLanes = 1;
#endif
// levels between a record's L2 prefetch and its register load: enough levels that their arithmetic covers an HBM
// round trip -- a level is ~4 NSAVE instructions, so small records need a longer run-ahead (B200_WS_AHEAD_INSTR
// instructions' worth)
#ifndef B200_WS_AHEAD_INSTR
#define B200_WS_AHEAD_INSTR 400
#endif
template <int NSAVE>
constexpr int kWsPrefetchAhead = (B200_WS_AHEAD_INSTR + 4 * NSAVE - 1) / (4 * NSAVE) < 3 ? 3 : (B200_WS_AHEAD_INSTR + 4 * NSAVE - 1) / (4 * NSAVE);

// Prefetch the (warp, level) record -- NSAVE * kWsLanes doubles, contiguous -- one 128-byte line per lane and
// instruction.  wsl is the lane's own base pointer (it already includes the lane offset).
template <int NSAVE>
B200_HD void ws_prefetch_record(const double* wsl, int lev, long tid) {
#if defined(__CUDA_ARCH__)
  if (lev < 0) return;
  const int lane = (int)(tid % kWsLanes);
  const char* rec = reinterpret_cast<const char*>(wsl - lane + (long)lev * NSAVE * kWsLanes);
  constexpr int kLines = (NSAVE * kWsLanes * 8 + 127) / 128;
#pragma unroll
  for (int l0 = 0; l0 < kLines; l0 += kWsLanes)
    if (l0 + lane < kLines) asm volatile("prefetch.global.L2 [%0];" ::"l"(rec + (long)(l0 + lane) * 128));
#else
  (void)wsl; (void)lev; (void)tid;
#endif
}

// ---------------------------------------------------------------------------
// TMA-fed back sweep (nstr <= 8).  With the [warp][level][element][lane] layout the record of one (warp, level) is a
// contiguous, 256-byte aligned run of NSAVE * 256 bytes, i.e. exactly what one 1-D bulk copy moves.  Each warp owns
// a ring of kStages such records in shared memory; lane 0 issues `cp.async.bulk` for the records several levels
// ahead, completion is counted by one mbarrier per stage, and the 32 lanes then read their elements with
// conflict-free LDS.  Why not plain loads into registers: ptxas puts every LDG of the register ring on ONE
// scoreboard, and a scoreboard can only be waited to zero -- so using the oldest buffer waited for the youngest
// loads as well, one full HBM latency per level (ncu: 21 % of the amars_lw samples on 2.6 % of its instructions).
// The mbarrier ring has per-stage completion, and one instruction per warp replaces 25 loads per thread.
// ---------------------------------------------------------------------------
#ifndef B200_TMA_BACK
#define B200_TMA_BACK 1
#endif
#ifndef B200_EXP_HALFREC
#define B200_EXP_HALFREC 0
#endif
template <int NSAVE>
struct TmaRing {
  // stages: what fits 150-200 KB of shared memory per SM at the kernel's residency (8 warps/SM at nn = 4,
  // 16 warps/SM below), at most 8
#ifdef B200_TMA_STAGES
  static constexpr int kStages = B200_TMA_STAGES;
#else
  static constexpr int kStages = NSAVE >= 16 ? 3 : NSAVE >= 9 ? 5 : 8;
#endif
  static constexpr int kRecDoubles = NSAVE * kWsLanes;
  static constexpr int kRecBytes = kRecDoubles * 8;
  // profiling experiment only (-DB200_EXP_HALFREC=1, never in the product build): elements e >= kLiveElems of a record are
  // neither stored nor copied (the back sweep re-uses element e % kLiveElems) -- same instructions, half the workspace
  // traffic, wrong fluxes: what cutting the back-sweep records in half could buy at most (profiles/r02/sweeps.md, e42)
  static constexpr int kLiveElems = B200_EXP_HALFREC ? (NSAVE + 1) / 2 : NSAVE;
  static constexpr int kCopyBytes = kLiveElems * kWsLanes * 8;
  static constexpr int kWarpDoubles = kStages * kRecDoubles + 16;  // + 128 bytes: the stage barriers
  static constexpr bool kUsed = B200_TMA_BACK && NSAVE <= 32;
};

#if defined(__CUDACC__)
__device__ __forceinline__ unsigned b200_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void b200_mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b200_smem_u32(bar)), "r"(count) : "memory");
}
// arm the barrier with the byte count of the copy and start the copy (issued by one lane)
__device__ __forceinline__ void b200_tma_load(double* dst, const double* src, int bytes, unsigned long long* bar) {
  const unsigned b = b200_smem_u32(bar);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   b200_smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(b)
               : "memory");
}
__device__ __forceinline__ void b200_mbar_wait(unsigned long long* bar, unsigned parity) {
  const unsigned b = b200_smem_u32(bar);
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "B200_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra B200_DONE;\n"
      "bra B200_WAIT;\n"
      "B200_DONE:\n"
      "}\n" ::"r"(b),
      "r"(parity)
      : "memory");
}
#endif

// ---------------------------------------------------------------------------
// Carried back sweep (nstr <= 8): the back sweep of a finished forward sweep rides on the same thread's NEXT
// solve -- one level is emitted at the end of every layer of the next forward sweep, its record copied into the
// warp's ring by TMA at least one layer (~2000 instructions) earlier, so no warp ever waits for HBM between
// two forward sweeps.  Forward sweep alone: 3.85 ms per amars_lw launch; with a stand-alone back sweep: 4.50 ms;
// carried: 4.38 ms (profiles/r01, sweeps e11, e12) -- what remains is the ~100 instructions a level costs at the
// per-warp issue rate of this kernel, not memory latency.  The two solves use alternating halves of the workspace.  Everything that
// synchronises (mbarrier waits, __syncwarp, the refill by lane 0) is warp-uniform: `lev` is the same in all lanes
// (lanes whose medium is cut off higher up skip levels they do not own), and the layer loop of a warp runs until
// all of its lanes are cut off.
// ---------------------------------------------------------------------------
#ifndef B200_CARRY_BACK
#define B200_CARRY_BACK 1
#endif
// Measured (profiles/r01, sweep e12): +3 % at nstr = 8 (amars_lw), none on amars_sw, -2.4 % at nstr = 4 (LBL, where a
// layer of the forward sweep is too short to amortise the per-level bookkeeping) -- so it is used from nn = 4 up.
template <int NN>
constexpr bool kCarryBack = B200_CARRY_BACK && NN >= 4 && TmaRing<NN * NN + 2 * NN + 1>::kUsed;

// One output level of a solve: row = pyharp level index (0 = surface).
//   BAND = 0  the per-solve flux row;
//   BAND = 1  fused band integration: w_n F_n accumulated into the band-flux buffer with FP64 reductions in the L2
//             (RED.E.ADD.F64; the summation order over spectral points is not fixed -- a few ulp between runs);
//   BAND = 2  the same when fewer than 32 columns share the batch (line-by-line: 10^5 spectral points x 1 column), so
//             that the lanes of a warp are spectral points of the SAME column: the warp first adds its lanes' terms
//             per column with shuffles (p.warp_red = ncol, a power of two <= 16; lane l holds column l % ncol) and
//             only `ncol` lanes touch the L2 -- 32 same-address reductions per instruction would serialise there.
//             Must be reached by all 32 lanes with a warp-uniform row; `pred` masks lanes without a value.
constexpr int kBandOff = 0, kBandAtomic = 1, kBandWarpRed = 2;
template <int BAND>
B200_HD void emit_flux(double* out, int row, long stride, double wg, double fup, double fdn, bool pred, int warp_red) {
  if constexpr (BAND == kBandOff) {
    if (pred) {
      out[2 * row] = fup;
      out[2 * row + 1] = fdn;
    }
    (void)stride; (void)wg; (void)warp_red;
  } else {
#if defined(__CUDA_ARCH__)
    if constexpr (BAND == kBandWarpRed) {
      double a = pred ? wg * fup : 0.0, b = pred ? wg * fdn : 0.0;
      for (int o = 16; o >= warp_red; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
      }
      if ((int)(threadIdx.x & 31) < warp_red) {
        atomicAdd(out + (long)(2 * row) * stride, a);
        atomicAdd(out + (long)(2 * row + 1) * stride, b);
      }
    } else {
      if (pred) {
        atomicAdd(out + (long)(2 * row) * stride, wg * fup);
        atomicAdd(out + (long)(2 * row + 1) * stride, wg * fdn);
      }
      (void)warp_red;
    }
#else
    if (pred) {
      out[(long)(2 * row) * stride] += wg * fup;
      out[(long)(2 * row + 1) * stride] += wg * fdn;
    }
    (void)warp_red;
#endif
  }
}

template <int NN>
struct BackPend {
  double up[NN];           // upward radiation at level lev + 1
  double wg = 0.0;         // BAND: band weight of the pending solve
  double* out = nullptr;   // flux rows of the pending solve (BAND: its column in the band-flux buffer)
  const double* wsl = nullptr;  // lane base (level 0) of its workspace half
  int ncut = 0;            // this lane emits levels lev < ncut
  int lev = -1;            // warp-uniform: next level to emit, counting down; < 0: nothing pending
  int nlyr = 0;
  bool bad = false, active = false;
};

#if defined(__CUDA_ARCH__)
#define B200_WARP_ANY(pred) (__any_sync(0xffffffffu, (pred)) != 0)
#define B200_WARP_SYNC() __syncwarp()
#else
#define B200_WARP_ANY(pred) (pred)
#define B200_WARP_SYNC() ((void)0)
#endif

// Validity of pivots / eigenvalues, accumulated on the integer pipe: a double is a usable pivot (finite, > kTiny)
// exactly when its high word lies in [hi(kTiny), hi(DBL_MAX)]; anything else -- zero, negative, tiny, Inf, NaN --
// falls outside after the unsigned shift.  Two integer instructions per pivot instead of a DSETP + selects, and
// one test per layer.
struct PivotCheck {
  unsigned worst = 0u;
  B200_HD void add(double d) { worst = max(worst, (unsigned)B200_HIWORD(d) - 0x03B8F2B0u); }
  B200_HD bool bad() const { return worst > 0x7FEFFFFFu - 0x03B8F2B0u; }
};

// Jacobi rotation that annihilates g_pq (small-angle root): t = tan, s = sin, c = cos of the angle, from the
// half-angle identities with two reciprocal square roots -- no sqrt, no division (measured cost of the classical
// form t = b / (d + sign(d) sqrt(d^2 + b^2)), c = 1/sqrt(1 + t^2), tau = s / (1 + c): 49 DFMA-equivalents, this
// one: 32).  cos 2a = |d| h >= 0, so x = (1 + cos 2a) / 2 is in [1/2, 1] and nothing cancels.  A pair below the
// relative threshold gets the identity (d := 1, b := 0), not a branch.
B200_HD void jacobi_rotation(double gpq, double gpp, double gqq, double eps2, double& t, double& s, double& c) {
  const bool rot = gpq * gpq > eps2 * fabs(gpp * gqq);
  const double d = rot ? gqq - gpp : 1.0, b2 = rot ? 2.0 * gpq : 0.0;
  const double sd = copysign(B200_RSQRT(fma(d, d, b2 * b2)), d);
  const double x = fma(0.5 * d, sd, 0.5);
  const double rc = B200_RSQRT(x);
  c = x * rc;
  s = (0.5 * b2) * sd * rc;
  t = s * rc;
}

// Brent-Luk tournament: data source of position i after one round (n even; pairs are (2k, 2k+1))
B200_HD constexpr int bl_src(int i, int n) {
  return (i == 0) ? 0 : (i == 2) ? 1 : (i % 2 == 0) ? i - 2 : (i == n - 1) ? n - 2 : i + 2;
}

// Eigenvector accumulator Y of the Jacobi iteration: a thread-private array (registers / local memory), or -- nn >= 6,
// -DB200_SMEM_Y=1 -- the thread's slice of the CTA's shared memory ([element][thread], conflict-free).  At nstr = 16
// Y is 64 doubles = 128 registers of a 255-register thread that also holds G, L and the stack operator; ptxas spills
// it and the hot loop streams it through L1/L2 (ncu: long_scoreboard 3.2 stalls per issue).  In shared memory its
// columns are addressed at run time, so the Brent-Luk rounds permute an index vector instead of moving Y.
#ifndef B200_SMEM_Y
#define B200_SMEM_Y 0
#endif
template <int NN, bool SMEM>
struct YRef;
template <int NN>
struct YRef<NN, false> {
  static constexpr bool kSmem = false;
  double (&a)[NN][NN];
  B200_HD explicit YRef(double (&arr)[NN][NN]) : a(arr) {}
  B200_HD double& operator()(int i, int j) const { return a[i][j]; }
};
#if defined(__CUDACC__)
template <int NN>
struct YRef<NN, true> {
  static constexpr bool kSmem = true;
  double* base;
  int stride;
  __device__ __forceinline__ explicit YRef(double (&)[NN][NN]) {
    extern __shared__ __align__(128) double b200_smem[];
    base = b200_smem + threadIdx.x;
    stride = (int)blockDim.x;
  }
  __device__ __forceinline__ double& operator()(int i, int j) const { return base[(i * NN + j) * stride]; }
};
#endif

// Warp-per-solve form (disort_solver_body.inc, last section): per-layer hand-over record
//   Rl | Tl (packed symmetric) | s+ | s- | expbea at the layer top | at the bottom
template <int NN>
constexpr int kOpsDoubles = NN * (NN + 1) + 2 * NN + 2;
// ... and what every phase of a solve needs of its boundary conditions
template <int NN>
struct SolveCommon {
  double umu0, rmu0, fbeam, albedo, fisot, tplank, bplank;
  double y0[2 * NN];
  const double* pk;  // Planck emission per level, DISORT order (null without thermal sources)
  int n, j, st;
  bool beam, planck;
};

// fully unrolled instantiation (small nn: registers)
#define B200_UNROLL _Pragma("unroll")
#define B200_FN(name) name##_u
#include "disort_solver_body.inc"
#undef B200_UNROLL
#undef B200_FN
// rolled instantiation (large nn: local memory)
#define B200_UNROLL _Pragma("unroll 1")
#define B200_FN(name) name##_r
#include "disort_solver_body.inc"
#undef B200_UNROLL
#undef B200_FN

// NOMOM = the launch carries no phase-function moments (nprop <= 2): isotropic scattering, compile-time lmax = 1
template <int NN, bool NOMOM = false, int BAND = 0, typename QT>
B200_HD void solve_flux_one(const SolveParams& p, const QT& qt, long s, long tid, bool active, unsigned& ring_pos,
                            BackPend<NN>& pend, int wshalf) {
#ifndef B200_UNROLL_MAX_NN
#define B200_UNROLL_MAX_NN 4
#endif
  if constexpr (NN <= B200_UNROLL_MAX_NN) solve_flux_one_u<NN, NOMOM, BAND>(p, qt, s, tid, active, ring_pos, pend, wshalf);
  else solve_flux_one_r<NN, NOMOM, BAND>(p, qt, s, tid, active, ring_pos, pend, wshalf);
}

// layer-wise (warp-per-solve) order on one thread: host simulator entry
template <int NN, bool NOMOM = false, int BAND = 0, typename QT>
B200_HD void solve_flux_layerwise(const SolveParams& p, const QT& qt, long s, double* opsbuf, double* recbuf) {
  solve_flux_layerwise_u<NN, NOMOM, BAND>(p, qt, s, opsbuf, recbuf);
}

// Layer-split form (flux_inst.cu, nstr >= 12): the phases of a solve as separate kernels -- unrolled or rolled bodies
// as for solve_flux_one, with their own limit (the split kernels are smaller than the fused one: unrolling them at a
// stream count whose fused kernel stays rolled costs less compile time)
#ifndef B200_SPLIT_UNROLL_MAX_NN
#define B200_SPLIT_UNROLL_MAX_NN B200_UNROLL_MAX_NN
#endif
template <int NN, typename QT>
B200_HD void solve_common_any(const SolveParams& p, const QT& qt, long s, SolveCommon<NN>& c, bool bc_planck, bool legendre) {
  if constexpr (NN <= B200_SPLIT_UNROLL_MAX_NN) solve_common_u<NN>(p, qt, s, c, bc_planck, legendre);
  else solve_common_r<NN>(p, qt, s, c, bc_planck, legendre);
}
template <int NN, bool NOMOM>
B200_HD void layer_depths_any(const SolveParams& p, const double* pr, int lc, double& dtp, double& dabs) {
  layer_depths_u<NN, NOMOM>(p, pr, lc, dtp, dabs);
}
template <int NN, bool NOMOM, typename QT>
B200_HD int layer_ops_any(const SolveParams& p, const QT& qt, const SolveCommon<NN>& cm, const double* pr, int lc, double taup,
                          double* ops, long ostr, const double* y0p, long y0s) {
  if constexpr (NN <= B200_SPLIT_UNROLL_MAX_NN) return layer_ops_u<NN, NOMOM>(p, qt, cm, pr, lc, taup, ops, ostr, y0p, y0s);
  else return layer_ops_r<NN, NOMOM>(p, qt, cm, pr, lc, taup, ops, ostr, y0p, y0s);
}
template <int NN, int BAND, typename QT>
B200_HD void adding_sweep_any(const SolveParams& p, const QT& qt, const SolveCommon<NN>& cm, long s, const double* pr, int ncut,
                              bool lyrcut, int st, const double* opsbuf, double* recbuf, long bstr, long blane, bool active) {
  if constexpr (NN <= B200_SPLIT_UNROLL_MAX_NN) adding_sweep_u<NN, BAND>(p, qt, cm, s, pr, ncut, lyrcut, st, opsbuf, recbuf, bstr, blane, active);
  else adding_sweep_r<NN, BAND>(p, qt, cm, s, pr, ncut, lyrcut, st, opsbuf, recbuf, bstr, blane, active);
}

// finishes a pending (carried) back sweep after a thread's last solve
template <int NN, int BAND = 0, typename QT>
B200_HD void drain_back_sweep(const SolveParams& p, BackPend<NN>& pend, const QT& qt, long tid, unsigned& ring_pos) {
  if constexpr (NN <= B200_UNROLL_MAX_NN) {
    while (pend.lev >= 0) back_level_u<NN, BAND>(p, pend, qt, tid, ring_pos);
  } else {
    while (pend.lev >= 0) back_level_r<NN, BAND>(p, pend, qt, tid, ring_pos);
  }
}

}  // namespace b200
