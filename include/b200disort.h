/*
 * b200disort.h -- C ABI of the B200-native DISORT solve path.
 *
 * This is the drop-in boundary beneath pyharp's solver plugin.  Each entry point
 * names the reference interface it replaces (paths relative to the pyharp tree):
 *
 *   b200disort_create      <- disort::Disort ctor / DisortImpl::reset, which
 *                             allocate one disort_state per (wave, column)
 *                             (tests/test_disort.cpp:29, examples/amars_sw.cpp:216,
 *                             examples/amars_lw.cpp:60, src/radiation/radiation_band.cpp:58-68)
 *   b200disort_solve       <- disort::DisortImpl::forward(prop, bc*, temf?) and the
 *   b200disort_solve_host     c_disort call it loops over (tests/test_disort.cpp:49,
 *                             examples/amars_sw.cpp:280, examples/amars_lw.cpp:80,
 *                             src/radiation/radiation_band.cpp:124-127,
 *                             src/rtsolver/rt_solver_disort.cpp_:147,266)
 *   b200disort_get_rad     <- disort::DisortImpl::get_rad (tests/test_disort.cpp:52)
 *   b200disort_reduce_band <- the band integration bflx = sum_g w_g F_g
 *                             (examples/amars_lw.cpp:84-88,
 *                             src/rtsolver/rt_solver_disort.cpp_:183-184)
 *   b200disort_destroy     <- Disort module destruction
 *
 * Plain pointers and sizes only; no torch / C++ types.  All arithmetic is FP64.
 * There is no CPU fallback: every solve runs as sm_100a CUDA kernels.
 *
 * Tensor contract (identical to pydisort's):
 *   prop  (nwave, ncol, nlyr, nprop) row-major; prop[...,0] = layer optical depth,
 *         prop[...,1] = single-scattering albedo (0 if nprop < 2), prop[...,2+k] =
 *         phase-function Legendre moment k+1 (moment 0 is forced to 1; missing
 *         moments are 0).  LAYER INDEX 0 IS THE SURFACE LAYER (pyharp is bottom-up).
 *   bc    each field (nwave, ncol) or NULL.  NULL => 0, except umu0 => 1.
 *   temf  (ncol, nlyr+1) level temperatures, index 0 = surface level; needed iff PLANCK.
 *   flx   (nwave, ncol, nlev, 2): [...,0] = upward flux (flup), [...,1] = total downward
 *         flux (rfldir + rfldn); nlev = nlyr+1 (index 0 = surface) or, with USRTAU,
 *         ntau user levels in the same bottom-up (reversed) order.
 *   uu    (nwave, ncol, nphi, ntau, numu) intensities (cdisort's uu order, numu fastest),
 *         user levels/angles in the order given (DISORT order, top first).
 *   status (nwave, ncol) int32 per-solve status word (B200DISORT_ST_*).
 */
#ifndef B200DISORT_H
#define B200DISORT_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* flags: the tokens DisortOptions::flags("...") parses onto ds.flag.* */
#define B200DISORT_USRTAU 1u
#define B200DISORT_USRANG 2u
#define B200DISORT_LAMBER 4u
#define B200DISORT_PLANCK 8u
#define B200DISORT_ONLYFL 16u
#define B200DISORT_INTENSITY_CORRECTION 32u
#define B200DISORT_OLD_INTENSITY_CORRECTION 64u
#define B200DISORT_QUIET 128u

/* return codes of the entry points */
#define B200DISORT_OK 0
#define B200DISORT_ERR_INVALID 1     /* bad argument / configuration */
#define B200DISORT_ERR_CUDA 2        /* CUDA runtime failure (no GPU, OOM, launch) */
#define B200DISORT_ERR_UNSUPPORTED 3 /* feature of cdisort not on pyharp's path */
#define B200DISORT_ERR_SOLVE 4       /* at least one solve has non-zero status */

/* per-solve status words (mirror c_check_inputs / c_errmsg conditions) */
#define B200DISORT_ST_OK 0
#define B200DISORT_ST_BAD_INPUT 1   /* tau<0, ssalb outside [0,1], |pmom|>1, T<0, ... */
#define B200DISORT_ST_EIGEN_FAIL 2  /* eigen iteration did not converge */
#define B200DISORT_ST_SINGULAR 3    /* singular particular-solution / boundary system */
#define B200DISORT_ST_UNSUPPORTED 4
#define B200DISORT_ST_BEAM_ANGLE 5  /* umu0 coincides with a quadrature angle */

typedef struct b200disort_handle b200disort_handle;

typedef struct {
  int nwave, ncol, nlyr; /* batch shape: nwave x ncol independent solves */
  int nstr, nmom;        /* streams (even, 2..32; the flux kernel is instantiated for 2,4,6,8,12,16,24,32, other
                            counts run on the general kernel) and phase moments (>= nstr) */
  int ntau, numu, nphi;  /* user grid sizes (used with USRTAU / USRANG) */
  unsigned flags;        /* B200DISORT_* bitmask */
  double accur;          /* azimuth-series convergence criterion */
  const double* user_tau; /* host [ntau], top-down optical depths */
  const double* user_mu;  /* host [numu] */
  const double* user_phi; /* host [nphi], degrees */
  const double* wave_lower; /* host [nwave] or NULL: wvnmlo per spectral point */
  const double* wave_upper; /* host [nwave] or NULL: wvnmhi */
  int device;   /* CUDA device ordinal */
  void* stream; /* cudaStream_t the handle launches on (NULL = default stream) */
} b200disort_config;

typedef struct {
  const double *umu0, *fbeam, *albedo, *btemp, *ttemp, *temis, *fisot, *phi0;
} b200disort_bc;

/* timing / launch statistics of a handle (b200disort_get_stats synchronises, reports, resets) */
typedef struct {
  int kernel_launches;     /* kernels launched by the last solve call */
  int solves_timed;        /* solve calls since the previous get_stats */
  float solve_kernel_ms;   /* summed device time of the dominant solve kernel over those calls
                              (CUDA events on the handle's stream around each launch) */
  float total_device_ms;   /* last call: first kernel start to solve kernel end */
  uint64_t workspace_bytes;
} b200disort_stats;

int b200disort_create(const b200disort_config* cfg, b200disort_handle** out);

/* Device-pointer solve: every pointer is a device pointer on cfg->device.
 * Asynchronous on the handle's stream except that a non-zero per-solve status is
 * only reported by b200disort_check_status / the _host variant. */
int b200disort_solve(b200disort_handle* h, const double* prop, int nprop,
                     const b200disort_bc* bc, const double* temf, double* flx,
                     double* uu, int32_t* status);

/* Host-pointer solve: same contract with HOST buffers; stages through pinned
 * memory, copies H2D, solves, copies D2H and synchronises.  Returns
 * B200DISORT_ERR_SOLVE if any status word is non-zero. */
int b200disort_solve_host(b200disort_handle* h, const double* prop, int nprop,
                          const b200disort_bc* bc, const double* temf, double* flx,
                          double* uu, int32_t* status);

/* Flux-only batches of at most `max_solves` solves run on the warp-per-solve form of the flux kernel (lanes take the
 * layers of a solve: per-call latency 1.6-1.9x lower for pyharp's own 500 x 1 / G x 1 calls); larger ones on the
 * thread-per-solve form.  -1 restores the default (one resident wave of warps), 0 disables the warp form. */
int b200disort_set_warp_solve_max(b200disort_handle* h, long max_solves);

/* Form of the flux solve at nstr = 4, 8 and nstr >= 12.  1 = layer-split (the layer phase -- eigenproblem, particular
 * solutions, reflection / transmission operators: cdisort's c_sol_eigen / c_upbeam / c_upisot -- as one thread per
 * (solve, layer), then the adding recursion and the back sweep -- c_set_matrix / c_solve0 / c_fluxes -- as a second kernel,
 * run by teams of four lanes per solve at nstr = 8 and 16), 0 = one fused kernel (or the warp-per-solve form for the
 * smallest batches), -1 = the default: split from nstr = 16 up, and at nstr = 8 for batches between the warp-per-solve
 * form's range and 8192 solves.  No effect at other stream counts. */
int b200disort_set_layer_split(b200disort_handle* h, int mode);

/* Synchronise and count non-zero status words of a device status array. */
int b200disort_check_status(b200disort_handle* h, const int32_t* status, int* n_bad);

/* Copy the intensities of the last solve into a device buffer
 * (nwave, ncol, nphi, ntau, numu). */
int b200disort_get_rad(b200disort_handle* h, double* uu);

/* bflx(ncol, nlev, 2) = sum_n weights[n] * flx(n, ncol, nlev, 2); device pointers.
 * This is the local (per-GPU) part of the band integration; ranks all-reduce bflx. */
int b200disort_reduce_band(b200disort_handle* h, const double* flx,
                           const double* weights, double* bflx);

int b200disort_get_stats(b200disort_handle* h, b200disort_stats* out);

/* Change the cudaStream_t the handle launches on (NULL = default stream).  Waits for the work already queued on
 * the previous stream, so buffers of earlier calls may be reused by the caller afterwards. */
int b200disort_set_stream(b200disort_handle* h, void* stream);

/* Measured FP64 FMA throughput of the device in TFLOP/s (a ~20 ms DFMA-saturating
 * kernel): the roofline denominator bench.py reports against. */
int b200disort_fp64_peak(int device, double* tflops);

void b200disort_destroy(b200disort_handle* h);
const char* b200disort_strerror(int code);
const char* b200disort_version(void);

#ifdef __cplusplus
}
#endif
#endif /* B200DISORT_H */
