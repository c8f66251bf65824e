// opacity.cu -- GPU opacity interpolation and pre-processing kernels (include/b200opacity.h).
// HBM-bound gathers / streaming kernels: one thread per output element, outputs written in the
// solver's (nwave, ncol, nlyr, nprop) layout so they feed b200disort_solve directly.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>

#include "../../include/b200opacity.h"
#include "opacity_device.cuh"

namespace b200 {

template <int NDIM, int NVAL>
__global__ void interpn_kernel(const double* __restrict__ coord, long npoint, long cstride, const double* __restrict__ data,
                               const double* __restrict__ axis, Dims d, double* __restrict__ out) {
  for (long p = (long)blockIdx.x * blockDim.x + threadIdx.x; p < npoint; p += (long)gridDim.x * blockDim.x) {
    double c[NDIM], v[NVAL];
#pragma unroll
    for (int k = 0; k < NDIM; ++k) c[k] = coord[p * cstride + k];
    interpn_dev<NDIM, NVAL>(v, c, data, axis, d, 0);
#pragma unroll
    for (int j = 0; j < NVAL; ++j) out[p * NVAL + j] = v[j];
  }
}

// S8Fuller / H2SO4Simple forward: the table lookup depends on the spectral point only, so each CTA row
// (blockIdx.y = spectral point) interpolates once per thread from L1/L2-resident tables and streams
// conc -> out with coalesced 16-byte stores.
__global__ void aerosol_kernel(const double* __restrict__ kwave, const double* __restrict__ kdata, int rows,
                               const double* __restrict__ wave, int is_wavenumber, const double* __restrict__ conc,
                               long ncl, int nspecies, int sid, double* __restrict__ out) {
  const int w = blockIdx.y;
  Dims d{};
  d.len[0] = rows; d.stride[0] = 2; d.axis_off[0] = 0;
  double c = is_wavenumber ? __ddiv_rn(1.0e4, wave[w]) : wave[w], k[2];
  interpn_dev<1, 2>(k, &c, kdata, kwave, d, 0);
  double2* o = reinterpret_cast<double2*>(out) + (long)w * ncl;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < ncl; i += (long)gridDim.x * blockDim.x) {
    const double ext = __dmul_rn(k[0], conc[i * nspecies + sid]);
    o[i] = make_double2(ext, __dmul_rn(k[1], ext));
  }
}

// RFM forward: reference temperature on the (quirky, length-nlyr) reference grid, then the 3-D table.
__global__ void rfm_kernel(const double* __restrict__ kaxis, const double* __restrict__ kdata, Dims d,
                           const double* __restrict__ klnp, const double* __restrict__ ktemp, int nref_used,
                           const double* __restrict__ pres, const double* __restrict__ temp,
                           const double* __restrict__ conc, long ncl, int nspecies, int sid, double* __restrict__ out) {
  const int w = blockIdx.y;
  Dims r{};
  r.len[0] = nref_used; r.stride[0] = 1; r.axis_off[0] = 0;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < ncl; i += (long)gridDim.x * blockDim.x) {
    double lnp = log(pres[i]), tref;
    interpn_dev<1, 1>(&tref, &lnp, ktemp, klnp, r, 0);
    double c[3] = {kaxis[w], lnp, __dsub_rn(temp[i], tref)}, v;
    interpn_dev<3, 1>(&v, c, kdata, kaxis, d, 0);
    out[(long)w * ncl + i] = __dmul_rn(__dmul_rn(1.0e-3, exp(v)), conc[i * nspecies + sid]);
  }
}

__global__ void layer2level_kernel(const double* __restrict__ var, long nrow, int nlyr, int order, int blower, int bupper,
                                   double* __restrict__ out, int* __restrict__ negative) {
  const long total = nrow * (nlyr + 1);
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long row = i / (nlyr + 1);
    const int lev = (int)(i - row * (nlyr + 1));
    const double* v = var + row * nlyr;
    double r;
    if (nlyr == 1) r = v[0];
    else if (lev == 0) r = (blower == B200_L2L_EXTRAPOLATE) ? __ddiv_rn(__dsub_rn(__dmul_rn(3.0, v[0]), v[1]), 2.0) : v[0];
    else if (lev == nlyr)
      r = (bupper == B200_L2L_EXTRAPOLATE) ? __ddiv_rn(__dsub_rn(__dmul_rn(3.0, v[nlyr - 1]), v[nlyr - 2]), 2.0) : v[nlyr - 1];
    else if (order == B200_L2L_2ND_ORDER || lev == 1 || lev == nlyr - 1) r = __ddiv_rn(__dadd_rn(v[lev - 1], v[lev]), 2.0);
    else {
      // Center4Interp: matmul(window, (-1/12, 7/12, 7/12, -1/12)), accumulated left to right
      const double cm0 = -1.0 / 12.0, cm1 = 7.0 / 12.0;
      r = __dmul_rn(v[lev - 2], cm0);
      r = __dadd_rn(r, __dmul_rn(v[lev - 1], cm1));
      r = __dadd_rn(r, __dmul_rn(v[lev], cm1));
      r = __dadd_rn(r, __dmul_rn(v[lev + 1], cm0));
    }
    out[i] = r;
    if (negative && r < 0.0) *negative = 1;
  }
}

// one thread per row (column x band ...); the recurrence runs down the levels
__global__ void spherical_flux_kernel(double* __restrict__ flx, const double* __restrict__ x1f, const double* __restrict__ area,
                                      const double* __restrict__ vol, long nrow, int nx1) {
  for (long r = (long)blockIdx.x * blockDim.x + threadIdx.x; r < nrow; r += (long)gridDim.x * blockDim.x) {
    double* f = flx + r * nx1;
    const double* a = area + r * nx1;
    const double* v = vol + r * nx1;
    double fiu = f[nx1 - 1];
    for (int i = nx1 - 2; i >= 0; --i) {
      const double dx = __dsub_rn(x1f[i + 1], x1f[i]);
      const double volh = __dmul_rn(__ddiv_rn(__dsub_rn(fiu, f[i]), dx), v[i]);
      const double nf = __ddiv_rn(__dsub_rn(__dmul_rn(f[i + 1], a[i + 1]), volh), a[i]);
      f[i] = nf;
      fiu = nf;  // the reference's flx_iu aliases flx[..., i] after the assignment
    }
  }
}

constexpr int kMaxAtt = 8;
struct AttList {
  const double* p[kMaxAtt];
  int nprop[kMaxAtt];
  int n;
};

__global__ void band_optics_kernel(AttList a, const double* __restrict__ dz, long npoint, int nlyr, int nprop_out,
                                   double* __restrict__ prop) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < npoint; i += (long)gridDim.x * blockDim.x) {
    const int lyr = (int)(i % nlyr);
    for (int k = 0; k < nprop_out; ++k) {
      double s = 0.0;
      for (int t = 0; t < a.n; ++t)
        if (k < a.nprop[t]) s = __dadd_rn(s, a.p[t][i * a.nprop[t] + k]);
      prop[i * nprop_out + k] = s;
    }
    const double ext = prop[i * nprop_out];
    prop[i * nprop_out] = __dmul_rn(ext, dz[lyr]);
    // a layer without any extinction (conc = 0 aloft) has no scatterers either: ssa = 0, not 0/0
    if (nprop_out > 1) prop[i * nprop_out + 1] = (ext == 0.0) ? 0.0 : __ddiv_rn(prop[i * nprop_out + 1], ext);
  }
}

static unsigned grid_for(long n, int block) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return (unsigned)std::max<long>(1, std::min<long>((n + block - 1) / block, (long)sms * 16));
}

template <int NDIM>
static int launch_interpn(int nval, unsigned grid, cudaStream_t st, const double* coord, long npoint, long cstride,
                          const double* data, const double* axis, const Dims& d, double* out) {
  switch (nval) {
#define B200_CASE(V) case V: interpn_kernel<NDIM, V><<<grid, 256, 0, st>>>(coord, npoint, cstride, data, axis, d, out); break;
    B200_CASE(1) B200_CASE(2) B200_CASE(3) B200_CASE(4) B200_CASE(5) B200_CASE(6) B200_CASE(7) B200_CASE(8)
#undef B200_CASE
    default: return 1;
  }
  return cudaGetLastError() == cudaSuccess ? 0 : 2;
}

}  // namespace b200

using namespace b200;

extern "C" int b200_interpn(const double* coord, long npoint, long coord_stride, const double* data, const double* axis,
                            const int64_t* dims, int ndim, int nval, double* out, void* stream) {
  if (!coord || !data || !axis || !dims || !out || npoint < 0 || ndim < 1 || ndim > 4 || nval < 1 || nval > 8) return 1;
  for (int k = 0; k < ndim; ++k) if (dims[k] < 1) return 1;
  if (npoint == 0) return 0;
  const Dims d = make_dims(dims, ndim, nval);
  const unsigned grid = grid_for(npoint, 256);
  cudaStream_t st = (cudaStream_t)stream;
  switch (ndim) {
    case 1: return launch_interpn<1>(nval, grid, st, coord, npoint, coord_stride, data, axis, d, out);
    case 2: return launch_interpn<2>(nval, grid, st, coord, npoint, coord_stride, data, axis, d, out);
    case 3: return launch_interpn<3>(nval, grid, st, coord, npoint, coord_stride, data, axis, d, out);
    default: return launch_interpn<4>(nval, grid, st, coord, npoint, coord_stride, data, axis, d, out);
  }
}

extern "C" int b200_aerosol_optics(const double* kwave, const double* kdata, int rows, const double* wave, int nwave,
                                   int is_wavenumber, const double* conc, int ncol, int nlyr, int nspecies, int species_id,
                                   double* out, void* stream) {
  if (!kwave || !kdata || !wave || !conc || !out || rows < 1 || nwave < 1 || ncol < 1 || nlyr < 1) return 1;
  if (species_id < 0 || species_id >= nspecies) return 1;
  const long ncl = (long)ncol * nlyr;
  dim3 grid(std::min<unsigned>(grid_for(ncl, 256), 64u), (unsigned)nwave);
  aerosol_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(kwave, kdata, rows, wave, is_wavenumber, conc, ncl, nspecies, species_id, out);
  return cudaGetLastError() == cudaSuccess ? 0 : 2;
}

extern "C" int b200_rfm_optics(const double* kaxis, const double* kdata, const int64_t* dims, const double* klnp,
                               const double* ktemp, int nref, const double* pres, const double* temp, const double* conc,
                               int ncol, int nlyr, int nspecies, int species_id, double* out, void* stream) {
  if (!kaxis || !kdata || !dims || !klnp || !ktemp || !pres || !temp || !conc || !out) return 1;
  if (species_id < 0 || species_id >= nspecies || nlyr > nref || nlyr < 1 || ncol < 1) return 1;
  const Dims d = make_dims(dims, 3, 1);
  const long ncl = (long)ncol * nlyr;
  dim3 grid(std::min<unsigned>(grid_for(ncl, 256), 64u), (unsigned)dims[0]);
  rfm_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(kaxis, kdata, d, klnp, ktemp, nlyr, pres, temp, conc, ncl, nspecies,
                                                     species_id, out);
  return cudaGetLastError() == cudaSuccess ? 0 : 2;
}

extern "C" int b200_layer2level(const double* var, long nrow, int nlyr, int order, int blower, int bupper, double* out,
                                int* negative, void* stream) {
  if (!var || !out || nrow < 1 || nlyr < 1 || order < 0 || order > 1 || blower < 0 || blower > 1 || bupper < 0 || bupper > 1)
    return 1;
  layer2level_kernel<<<grid_for(nrow * (nlyr + 1), 256), 256, 0, (cudaStream_t)stream>>>(var, nrow, nlyr, order, blower,
                                                                                         bupper, out, negative);
  return cudaGetLastError() == cudaSuccess ? 0 : 2;
}

extern "C" int b200_band_optics(const double* const* att, const int* nprops, int natt, const double* dz, int nwave, int ncol,
                                int nlyr, int nprop_out, double* prop, void* stream) {
  if (!att || !nprops || natt < 1 || natt > kMaxAtt || !dz || !prop || nprop_out < 1) return 1;
  AttList a{};
  a.n = natt;
  for (int t = 0; t < natt; ++t) { a.p[t] = att[t]; a.nprop[t] = nprops[t]; if (!att[t] || nprops[t] < 1) return 1; }
  const long npoint = (long)nwave * ncol * nlyr;
  band_optics_kernel<<<grid_for(npoint, 256), 256, 0, (cudaStream_t)stream>>>(a, dz, npoint, nlyr, nprop_out, prop);
  return cudaGetLastError() == cudaSuccess ? 0 : 2;
}

// heating rate of layer k from the band-integrated fluxes (examples/amars_sw.cpp:290-301): net = up - down at the
// two interfaces, dT/dt = -(1 / (rho cp)) (net[k+1] - net[k]) / dz; same evaluation order as the example
__global__ void heating_rate_kernel(const double* __restrict__ bflx, const double* __restrict__ rho,
                                    const double* __restrict__ dz, double cp, long ncol, int nlyr, double* __restrict__ out) {
  const long n = ncol * nlyr;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const long j = i / nlyr;
    const int k = (int)(i - j * nlyr);
    const double* f = bflx + (j * (nlyr + 1) + k) * 2;
    const double df = f[0] - f[1], df1 = f[2] - f[3];
    out[i] = -(1.0 / (rho[i] * cp)) * (df1 - df) / dz[i];
  }
}

extern "C" int b200_heating_rate(const double* bflx, const double* rho, const double* dz, double cp, long ncol, int nlyr,
                                 double* dTdt, void* stream) {
  if (!bflx || !rho || !dz || !dTdt || ncol < 1 || nlyr < 1 || !(cp > 0.0)) return 1;
  heating_rate_kernel<<<grid_for(ncol * nlyr, 256), 256, 0, (cudaStream_t)stream>>>(bflx, rho, dz, cp, ncol, nlyr, dTdt);
  return cudaGetLastError() == cudaSuccess ? 0 : 2;
}

extern "C" int b200_spherical_flux_correction(double* flx, const double* x1f, const double* area, const double* vol, long nrow,
                                              int nx1, void* stream) {
  if (!flx || !x1f || !area || !vol || nrow < 1 || nx1 < 2) return 1;
  spherical_flux_kernel<<<grid_for(nrow, 128), 128, 0, (cudaStream_t)stream>>>(flx, x1f, area, vol, nrow, nx1);
  return cudaGetLastError() == cudaSuccess ? 0 : 2;
}
