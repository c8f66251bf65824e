/*
 * b200band.h -- C ABI of the band-level forward: one call that runs pyharp's whole per-band pipeline on the GPU,
 *
 *   conc / pres / temp  ->  attenuator optics (RFM ck / line-by-line tables, S8Fuller / H2SO4Simple aerosols)
 *                       ->  property mixing (+ dz)  ->  layer2level + Planck pre-pass
 *                       ->  DISORT flux solve of every (spectral point, column)
 *                       ->  band integration sum_n w_n F_n fused into the solve kernel's flux epilogue
 *                       ->  (optional) heating rate
 *
 * i.e. what RadiationBandImpl::forward does in the reference (src/radiation/radiation_band.cpp:75-128: attenuator
 * loop :86-104, normalisation :107-116, rtsolver.forward(prop, &bc, layer2level(temp)) :123-128, and the band-weight
 * accumulation it leaves commented out :130-131) and what examples/amars_lw.cpp:62-88 and
 * examples/amars_sw.cpp:262-301 do by hand.  Only (ncol, nlyr, .) inputs and the (ncol, nlyr+1, 2) band flux cross
 * the boundary: neither the (nwave, ncol, nlyr, nprop) optical properties nor the (nwave, ncol, nlyr+1, 2)
 * per-spectral-point fluxes are ever copied, and unless the caller asks for them the fluxes are never written.
 *
 * Exported by libb200disort.so.  Plain pointers and sizes; FP64; no CPU fallback.
 *
 *   b200_band_create        <- RadiationBandImpl::reset (src/radiation/radiation_band.cpp:24-73): attenuators from
 *                              their tables, solver from DisortOptions; read_weights_rfm (src/utils/read_weights.cpp)
 *   b200_band_forward       <- RadiationBandImpl::forward on CUDA tensors (device pointers, asynchronous on the stream)
 *   b200_band_forward_host  <- the same on CPU tensors: columns flow through three (stream, buffer) slots so the H2D
 *                              copy of chunk c+1, the kernels of chunk c and the D2H copy of chunk c-1 overlap
 *   b200_band_allreduce     <- the cross-rank sum of the band flux when spectral points are sharded over GPUs
 *                              (the op of examples/amars_lw.cpp:88 across ranks): ncclAllReduce(sum, f64)
 */
#ifndef B200BAND_H
#define B200BAND_H
#include <stdint.h>

#include "b200disort.h"

#ifdef __cplusplus
extern "C" {
#endif

#define B200_ATT_RFM 1     /* harp::RFM: ln(sigma) table on (wave, ln p, T - Tref(p)), nprop = 1 */
#define B200_ATT_AEROSOL 2 /* harp::S8Fuller / harp::H2SO4Simple: (wavelength, k_ext, ssa) table, nprop = 2 */

/* One attenuator of the band; every pointer is a HOST array that b200_band_create copies to the device. */
typedef struct {
  int type;       /* B200_ATT_* */
  int species_id; /* column of conc this attenuator scales with (AttenuatorOptions::species_ids()[0]) */
  /* B200_ATT_RFM (RFMImpl members, src/opacity/rfm.hpp:20-40) */
  const double* kaxis; /* [dims[0] + dims[1] + dims[2]]: wavenumber | ln p | dT axes */
  const double* kdata; /* RFM: (dims[0], dims[1], dims[2]) ln(m^2/kmol); AEROSOL: (rows, 2) = (k_ext m^2/mol, ssa) */
  int64_t dims[3];     /* RFM table shape; dims[0] must equal nwave.  AEROSOL: dims[0] = rows */
  const double* klnp;  /* [nref] reference atmosphere ln p (krefatm[IPR]) */
  const double* ktemp; /* [nref] reference atmosphere T    (krefatm[ITM]) */
  int nref;
  /* B200_ATT_AEROSOL */
  const double* kwave; /* [rows] table wavelengths (um) */
} b200_attenuator;

#define B200_BAND_MAX_ATT 8

typedef struct {
  b200disort_config disort; /* batch shape, streams, flags (ONLYFL required), wave_lower/upper, device, stream */
  int natt;
  const b200_attenuator* att; /* mixed in this order (RadiationBand iterates its std::map in key order) */
  int nspecies;               /* last dimension of conc */
  const double* wave;         /* host [nwave]: spectral coordinate of the aerosol attenuators, or NULL */
  int wave_is_wavenumber;     /* 0: wavelength (um); 1: wavenumber (1/cm), converted 1e4 / wave (s8_fuller.cpp:82-84) */
  const double* weights;      /* host [nwave]: band weights w_n (read_weights_rfm / d(nu), examples/amars_sw.cpp:185) */
  int l2l_order, l2l_blower, l2l_bupper; /* Layer2LevelOptions (B200_L2L_* of b200opacity.h) for temp -> temf */
  /* 1: interpolate with harp::interpn's own arithmetic (IEEE divisions per blend: bit-identical table values);
   * 0: blend weights hoisted out of the spectral loop (<= 1e-13 relative on the optical properties, ~6x cheaper) */
  int exact_interp;
  /* 0: the examples' mixing (examples/amars_sw.cpp:262-273): prop = sum_a att_a; prop *= dz; ssa = prop[1] / prop[0]
   *    (a layer without extinction gets ssa = 0);
   * > 0: RadiationBand's guarded form (radiation_band.cpp:107-116): ssa = sum(ssa ext) / (sum(ext) + mix_eps),
   *    tau = sum(ext) dz; the reference uses 1e-10 */
  double mix_eps;
  /* columns per chunk of b200_band_forward_host; 0 = one resident wave of the solve kernel (the default) */
  int host_chunk_cols;
} b200_band_config;

typedef struct {
  /* inputs */
  const double* conc; /* (ncol, nlyr, nspecies) mol/m^3, layer 0 = surface */
  const double* pres; /* (ncol, nlyr) Pa; needed by RFM attenuators */
  const double* temp; /* (ncol, nlyr) K; needed by RFM attenuators and with PLANCK */
  const double* dz;   /* (nlyr) layer thickness m, or NULL = 1 (examples/amars_lw.cpp:79 "assuming dz = 1") */
  b200disort_bc bc;   /* each field (nwave, ncol) -- or (ncol), shared by all spectral points, if bc_per_column */
  int bc_per_column;
  const double* rho;  /* (ncol, nlyr) kg/m^3, needed iff dTdt */
  double cp;          /* J/(kg K), needed iff dTdt */
  /* outputs */
  double* bflx;    /* (ncol, nlyr+1, 2) band-integrated (up, down) flux, level 0 = surface; required */
  double* dTdt;    /* (ncol, nlyr) K/s heating rate (examples/amars_sw.cpp:290-301), or NULL; needs dz */
  int32_t* status; /* (nwave, ncol) per-solve status words, or NULL */
  double* flx;     /* device entry only: (nwave, ncol, nlyr+1, 2) per-spectral-point fluxes, or NULL.  When given, the
                      solve kernel writes them and the band sum is a second (deterministic) pass over them */
  double* prop;    /* device entry only: (nwave, ncol, nlyr, nprop) mixed optical properties, or NULL
                      (RadiationBand exports them as shared["radiation/<band>/optics"], radiation_band.cpp:119-120) */
} b200_band_io;

typedef struct b200_band b200_band;

int b200_band_create(const b200_band_config* cfg, b200_band** out);
/* number of optical properties per layer the band's attenuators produce (1: extinction only; 2: + ssa) */
int b200_band_nprop(const b200_band* b);
/* Device pointers; asynchronous on the configured stream.  A failed solve / negative interface temperature is
 * reported by b200_band_check (or by the _host entry). */
int b200_band_forward(b200_band* b, const b200_band_io* io);
/* The optics stage alone (device pointers, asynchronous): prop (nwave, ncol, nlyr, nprop) from io->conc / pres /
 * temp / dz -- the tensor RadiationBand exports as shared["radiation/<band>/optics"].  Only those fields and
 * io->prop are read. */
int b200_band_optical_props(b200_band* b, const b200_band_io* io);
/* Host pointers (pinned memory is DMA'd directly, pageable memory staged by the driver); synchronises. */
int b200_band_forward_host(b200_band* b, const b200_band_io* io);
/* Synchronise; B200DISORT_ERR_SOLVE if any solve of the calls since the last check had a non-zero status. */
int b200_band_check(b200_band* b, int* n_bad);
/* In-place sum of bflx (count doubles, device) over the ranks of an NCCL communicator (ncclComm_t passed as
 * void*), on the band's stream.  NCCL is resolved at run time from the libnccl already loaded in the process
 * (torch's, or the system one): the library itself does not link it. */
int b200_band_allreduce(b200_band* b, double* bflx, long count, void* nccl_comm);
/* The solver handle inside the band (statistics: b200disort_get_stats). */
b200disort_handle* b200_band_solver(b200_band* b);
void b200_band_destroy(b200_band* b);

/* The same collective for callers of the solver-level API (b200disort_reduce_band's output). */
int b200disort_allreduce_band(b200disort_handle* h, double* bflx, void* nccl_comm);

/* Solve + fused band integration without the opacity stage: bflx (ncol, nlyr+1, 2) = sum_n weights[n] F_n, the
 * (nwave, ncol, nlyr+1, 2) fluxes are never written.  Device pointers; weights (nwave).  ONLYFL without USRTAU. */
int b200disort_solve_band(b200disort_handle* h, const double* prop, int nprop, const b200disort_bc* bc,
                          const double* temf, const double* weights, double* bflx, int32_t* status);

#ifdef __cplusplus
}
#endif
#endif /* B200BAND_H */
