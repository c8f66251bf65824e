/*
 * b200opacity.h -- C ABI of the GPU opacity / pre-processing kernels that feed the DISORT solve
 * without a host round-trip.  Device pointers, FP64,
 * asynchronous on the given cudaStream_t (NULL = default stream).  Exported by libb200disort.so.
 *
 *   b200_interpn         <- harp::call_interpn_cuda<N>, declared at src/math/interpolation.hpp:12-14 and
 *                           never defined; semantics of harp::interpn<N> / harp::locate
 *                           (src/math/interpn.h:34-76, src/math/locate.h:14-42): bisection incl. the
 *                           exact-endpoint rule, clamp-to-edge, recursion order, same arithmetic
 *                           (no FMA contraction) => bit-identical to the reference CPU path
 *   b200_aerosol_optics  <- S8FullerImpl::forward / H2SO4SimpleImpl::forward on a CUDA tensor
 *                           (src/opacity/s8_fuller.cpp:72-117, h2so4_simple.cpp:72-117; the .is_cuda()
 *                           branch there is an empty stub)
 *   b200_rfm_optics      <- RFMImpl::forward + get_reftemp (src/opacity/rfm.cpp:122-197)
 *   b200_layer2level     <- harp::layer2level (src/utils/layer2level.cpp:7-79)
 *   b200_spherical_flux_correction <- harp::spherical_flux_correction (src/utils/spherical_flux_correction.cpp:3-18)
 *   b200_heating_rate <- the heating-rate post-processing of examples/amars_sw.cpp:290-301
 *   b200_band_optics     <- the optical-property mixing the examples do before the solve
 *                           (examples/amars_sw.cpp:262-273; RadiationBandImpl::forward
 *                           src/radiation/radiation_band.cpp:83-116 in the 4-D layout)
 */
#ifndef B200OPACITY_H
#define B200OPACITY_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* out[p, 0..nval) = multilinear interpolation of data at coord[p * coord_stride + 0..ndim).
 * data is (dims[0], ..., dims[ndim-1], nval) row-major; axis holds the ndim axes back to back.
 * dims is a HOST array.  1 <= ndim <= 4, 1 <= nval <= 8. */
int b200_interpn(const double* coord, long npoint, long coord_stride, const double* data, const double* axis,
                 const int64_t* dims, int ndim, int nval, double* out, void* stream);

/* out (nwave, ncol, nlyr, 2): [..,0] = kext(wave) * conc[col,lyr,species_id]  (1/m),
 *                             [..,1] = ssa(wave) * out[..,0].
 * kwave[rows] table wavelengths (um), kdata (rows, 2) = (kext m^2/mol, ssa); wave[nwave] is the wavelength (um)
 * or, with is_wavenumber != 0, the wavenumber (1/cm) converted as 1e4 / wave like s8_fuller.cpp:82-84. */
int b200_aerosol_optics(const double* kwave, const double* kdata, int rows, const double* wave, int nwave,
                        int is_wavenumber, const double* conc, int ncol, int nlyr, int nspecies, int species_id,
                        double* out, void* stream);

/* out (nwave, ncol, nlyr, 1) = 1e-3 * exp(interp3(kdata; kaxis[w], ln p, T - Tref(ln p))) * conc.
 * kdata (dims[0], dims[1], dims[2]) = ln sigma on (wave, ln p, dT); kaxis = the three axes back to back;
 * klnp / ktemp [nref] reference atmosphere.  get_reftemp interpolates with axis length nlyr
 * (src/opacity/rfm.cpp:178: dims = {nlyr}); nlyr <= nref is required. */
int b200_rfm_optics(const double* kaxis, const double* kdata, const int64_t* dims, const double* klnp,
                    const double* ktemp, int nref, const double* pres, const double* temp, const double* conc,
                    int ncol, int nlyr, int nspecies, int species_id, double* out, void* stream);

#define B200_L2L_2ND_ORDER 0
#define B200_L2L_4TH_ORDER 1
#define B200_L2L_EXTRAPOLATE 0
#define B200_L2L_CONSTANT 1
/* out (nrow, nlyr+1) from var (nrow, nlyr); *negative (device int, may be NULL) is set to 1 if any
 * interface value is negative (the reference's check_positivity). */
int b200_layer2level(const double* var, long nrow, int nlyr, int order, int blower, int bupper, double* out,
                     int* negative, void* stream);

/* prop (nwave, ncol, nlyr, nprop_out): sum of natt attenuator outputs att[a] (nwave, ncol, nlyr, nprops[a])
 * (extinction 1/m, ssa-weighted extinction, ...), then [..,0] *= dz[lyr] and [..,1] /= summed extinction.
 * att / nprops are HOST arrays of device pointers / ints. */
int b200_band_optics(const double* const* att, const int* nprops, int natt, const double* dz, int nwave, int ncol,
                     int nlyr, int nprop_out, double* prop, void* stream);

/* harp::spherical_flux_correction (src/utils/spherical_flux_correction.cpp:3-18), in place on
 * flx (nrow, nx1): x1f[nx1] interface coordinates, area / vol (nrow, nx1).  Walks down from the top level;
 * like the reference (whose `flx_iu` is a view of the tensor it then overwrites) the divergence at level i
 * uses the already corrected flux of level i+1. */
int b200_spherical_flux_correction(double* flx, const double* x1f, const double* area, const double* vol, long nrow,
                                   int nx1, void* stream);

/* Heating rate per layer from band-integrated fluxes, the post-processing of examples/amars_sw.cpp:290-301:
 * bflx (ncol, nlyr+1, 2) = (up, down) with level 0 at the surface (the output of b200disort_reduce_band, i.e. the
 * example's integrate_result, examples/amars_sw.cpp:173-196), rho / dz (ncol, nlyr) layer density [kg/m^3] and
 * thickness [m], cp [J/(kg K)]; dTdt (ncol, nlyr) [K/s] = -(1 / (rho cp)) (net[k+1] - net[k]) / dz, net = up - down. */
int b200_heating_rate(const double* bflx, const double* rho, const double* dz, double cp, long ncol, int nlyr,
                      double* dTdt, void* stream);

#ifdef __cplusplus
}
#endif
#endif
