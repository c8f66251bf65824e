// band.cu -- band-level forward (include/b200band.h): attenuator optics -> mixing -> layer2level + Planck pre-pass
// -> DISORT flux solve with the band integration fused into its flux epilogue (-> heating rate), as one call.
// Mirrors RadiationBandImpl::forward (src/radiation/radiation_band.cpp:75-128) and the hand-written pipelines of
// examples/amars_lw.cpp:62-88 / examples/amars_sw.cpp:262-301; sm_100a only, no CPU fallback.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "../../include/b200band.h"
#include "../../include/b200opacity.h"
#include "b200disort_internal.hpp"
#include "opacity_device.cuh"

namespace b200 {

// ---------------------------------------------------------------------------
// Fused band optics: prop(nwave, ncol, nlyr, nprop) straight from conc / pres / temp.
// One thread owns one (column, layer) point and a tile of WT spectral points, so everything that depends on the
// point only -- ln p, the reference temperature, the bracketing rows of the (ln p, dT) axes and their blend weights,
// the concentrations -- is computed once per tile instead of once per spectral point (the stand-alone rfm_kernel
// redoes it for every spectral point: ~600 FP64 instructions per value, this form ~45).  HBM traffic is the
// algorithmic minimum: the (ncol, nlyr) inputs are read once per tile, prop is written once.
// ---------------------------------------------------------------------------
struct DevAtt {
  int type, sid;
  const double* kaxis;  // RFM: [n0 + n1 + n2]
  const double* kdata;  // RFM: (n0, n1, n2)
  const double* kdata_t;  // RFM: the same table as (n1, n2, n0), spectral point fastest (hoisted-weight path)
  int n0, n1, n2;
  const double* klnp;   // RFM reference atmosphere [nref]
  const double* ktemp;
  int on_grid;          // RFM: every spectral point coincides with its own table row (wi = (w, .), ww = (1, 0))
  int sm_off;           // RFM: offset (doubles) of this attenuator's axes in the CTA's shared memory:
                        //      [ln p ref (nlyr) | T ref (nlyr) | ln p axis (n1) | dT axis (n2)]
  int same_axes;        // RFM: axes and reference atmosphere identical to the previous RFM attenuator's (same table
                        //      file): its reference temperature, bracketing rows and blend weights are reused
  const int* wi;        // RFM (nwave, 2): table rows a spectral point blends (it sits on the table's own grid)
  const double* ww;     // RFM (nwave, 2): weight of row wi[.][0] (harp::interpn's b / den) and of wi[.][1] (a / den)
  const double* kw;     // AEROSOL (nwave, 2): (k_ext, ssa) at every spectral point
};

struct BandOpticsParams {
  DevAtt att[B200_BAND_MAX_ATT];
  int natt, nwave, nlyr, nspecies, nprop, exact;
  int sm_doubles;  // shared memory of the axes (all RFM attenuators)
  long npoint;  // ncol * nlyr
  const double *conc, *pres, *temp, *dz;
  double eps;
  double* prop;
};

// harp::interpn's choice of the two table rows around c (src/math/interpn.h:40-52: clamp to the edge) and the blend
// weights of its last line with the division hoisted: value = w1 v[i1] + w2 v[i2]
__device__ __forceinline__ void bracket_dev(const double* ax, int len, double c, int& i1, int& i2, double& w1, double& w2) {
  i1 = locate_dev(ax, c, len);
  if (i1 == -1) { i1 = 0; i2 = 0; }
  else if (i1 == len - 1) i2 = len - 1;
  else i2 = i1 + 1;
  const double x1 = ax[i1], x2 = ax[i2];
  if (x2 != x1) {
    const double r = 1.0 / (x2 - x1);
    w2 = (c - x1) * r;
    w1 = (x2 - c) * r;
  } else {
    w1 = 0.5;
    w2 = 0.5;
  }
}

// 32-byte read-only load (four consecutive doubles, 32-byte aligned)
__device__ __forceinline__ void ld_nc_v4(const double* p, double (&v)[4]) {
  asm volatile("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
}

template <int WT, int MINB = 2>
__global__ void __launch_bounds__(256, MINB) band_optics_fused_kernel(const BandOpticsParams P) {
  const int w0 = blockIdx.x * WT;
  const int nw = min(WT, P.nwave - w0);
  // the interpolation axes live in shared memory: the bisections of harp::locate are chains of dependent loads at
  // addresses that differ from lane to lane -- from global memory each of them costs the L1 a wavefront per lane
  extern __shared__ double b200_axes[];
  for (int a = 0; a < P.natt; ++a) {
    const DevAtt& A = P.att[a];
    if (A.type != B200_ATT_RFM || A.same_axes) continue;
    double* s = b200_axes + A.sm_off;
    const int n = P.nlyr;
    for (int k = threadIdx.x; k < 2 * n + A.n1 + A.n2; k += blockDim.x)
      s[k] = k < n ? A.klnp[k] : k < 2 * n ? A.ktemp[k - n] : A.kaxis[A.n0 + (k - 2 * n)];
  }
  __syncthreads();
  for (long i = (long)blockIdx.y * blockDim.x + threadIdx.x; i < P.npoint; i += (long)gridDim.y * blockDim.x) {
    double ext[WT], sse[WT];
    // state shared by consecutive RFM attenuators on the same axes
    double lnp = 0.0, dT = 0.0, wp1 = 0.0, wp2 = 0.0, wt1 = 0.0, wt2 = 0.0;
    int ip1 = 0, ip2 = 0, it1 = 0, it2 = 0;
#pragma unroll
    for (int k = 0; k < WT; ++k) { ext[k] = 0.0; sse[k] = 0.0; }
    for (int a = 0; a < P.natt; ++a) {
      const DevAtt& A = P.att[a];
      const double cv = P.conc[i * P.nspecies + A.sid];
      if (A.type == B200_ATT_AEROSOL) {
        // S8FullerImpl::forward (src/opacity/s8_fuller.cpp:108-114): ext = k_ext conc, then ssa ext
#pragma unroll
        for (int k = 0; k < WT; ++k)
          if (k < nw) {
            const double e = __dmul_rn(A.kw[2 * (w0 + k)], cv);
            ext[k] = __dadd_rn(ext[k], e);
            sse[k] = __dadd_rn(sse[k], __dmul_rn(A.kw[2 * (w0 + k) + 1], e));
          }
        continue;
      }
      // RFMImpl::forward (src/opacity/rfm.cpp:122-169); get_reftemp interpolates on a reference grid of length
      // nlyr (rfm.cpp:178: dims = {nlyr}), reproduced as is
      const double* sax = b200_axes + A.sm_off;
      if (!A.same_axes) {
        lnp = log(P.pres[i]);
        Dims r{};
        r.len[0] = P.nlyr; r.stride[0] = 1; r.axis_off[0] = 0;
        double tref;
        interpn_dev<1, 1>(&tref, &lnp, sax + P.nlyr, sax, r, 0);
        dT = __dsub_rn(P.temp[i], tref);
        if (!P.exact) {
          bracket_dev(sax + 2 * P.nlyr, A.n1, lnp, ip1, ip2, wp1, wp2);
          bracket_dev(sax + 2 * P.nlyr + A.n1, A.n2, dT, it1, it2, wt1, wt2);
        }
      }
      if (P.exact) {
        Dims d{};
        d.len[0] = A.n0; d.len[1] = A.n1; d.len[2] = A.n2;
        d.stride[0] = (long)A.n1 * A.n2; d.stride[1] = A.n2; d.stride[2] = 1;
        d.axis_off[0] = 0; d.axis_off[1] = A.n0; d.axis_off[2] = A.n0 + A.n1;
#pragma unroll
        for (int k = 0; k < WT; ++k)
          if (k < nw) {
            double c[3] = {A.kaxis[w0 + k], lnp, dT}, v;
            interpn_dev<3, 1>(&v, c, A.kdata, A.kaxis, d, 0);
            ext[k] = __dadd_rn(ext[k], __dmul_rn(__dmul_rn(1.0e-3, exp(v)), cv));
          }
      } else {
        // transposed table (n1, n2, n0): the WT values a corner contributes are contiguous (whole 32-byte sectors
        // per load instead of one double of each), and neighbouring layers -- the lanes of a warp -- read
        // neighbouring rows
        const double* c11 = A.kdata_t + ((long)ip1 * A.n2 + it1) * A.n0 + w0;
        const double* c12 = A.kdata_t + ((long)ip1 * A.n2 + it2) * A.n0 + w0;
        const double* c21 = A.kdata_t + ((long)ip2 * A.n2 + it1) * A.n0 + w0;
        const double* c22 = A.kdata_t + ((long)ip2 * A.n2 + it2) * A.n0 + w0;
        const double f11 = wp1 * wt1, f12 = wp1 * wt2, f21 = wp2 * wt1, f22 = wp2 * wt2;
        const double cs = 1.0e-3 * cv;
        if constexpr (WT % 4 == 0) {
          if (A.on_grid && (A.n0 & 3) == 0 && nw == WT) {
            // four spectral points per 256-bit load (LDG.E.256): the lanes of a warp read different table rows, so
            // every load instruction costs the L1 one wavefront per lane whatever its width -- scalar loads made
            // this kernel L1-bound (ncu: l1tex throughput 88 %, FP64 pipe 10 %)
#pragma unroll
            for (int k = 0; k < WT; k += 4) {
              double a[4], b[4], c[4], d[4];
              ld_nc_v4(c11 + k, a);
              ld_nc_v4(c12 + k, b);
              ld_nc_v4(c21 + k, c);
              ld_nc_v4(c22 + k, d);
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const double v = fma(f11, a[q], fma(f12, b[q], fma(f21, c[q], f22 * d[q])));
                ext[k + q] = fma(exp(v), cs, ext[k + q]);
              }
            }
            continue;
          }
        }
#pragma unroll
        for (int k = 0; k < WT; ++k)
          if (k < nw) {
            double v;
            if (A.on_grid) {  // the spectral point is row w of the table (RFMImpl::forward, rfm.cpp:145)
              v = fma(f11, c11[k], fma(f12, c12[k], fma(f21, c21[k], f22 * c22[k])));
            } else {
              const int r1 = A.wi[2 * (w0 + k)] - w0 - k, r2 = A.wi[2 * (w0 + k) + 1] - w0 - k;
              const double g1 = A.ww[2 * (w0 + k)], g2 = A.ww[2 * (w0 + k) + 1];
              v = g1 * fma(f11, c11[k + r1], fma(f12, c12[k + r1], fma(f21, c21[k + r1], f22 * c22[k + r1])));
              if (g2 != 0.0) v += g2 * fma(f11, c11[k + r2], fma(f12, c12[k + r2], fma(f21, c21[k + r2], f22 * c22[k + r2])));
            }
            ext[k] = fma(exp(v), cs, ext[k]);
          }
      }
    }
    // mixing: examples/amars_sw.cpp:262-273 (mix_eps = 0) or radiation_band.cpp:107-116 (mix_eps > 0)
    const double dzl = P.dz ? P.dz[(int)(i % P.nlyr)] : 1.0;
#pragma unroll
    for (int k = 0; k < WT; ++k)
      if (k < nw) {
        const long o = (long)(w0 + k) * P.npoint + i;
        const double tau = __dmul_rn(ext[k], dzl);
        if (P.nprop == 1) {
          P.prop[o] = tau;
        } else {
          double ssa;
          if (P.eps > 0.0) ssa = __ddiv_rn(sse[k], __dadd_rn(ext[k], P.eps));
          else ssa = (tau == 0.0) ? 0.0 : __ddiv_rn(__dmul_rn(sse[k], dzl), tau);
          reinterpret_cast<double2*>(P.prop)[o] = make_double2(tau, fmin(ssa, 1.0));
        }
      }
  }
}


// ---------------------------------------------------------------------------
// Tile form of the fused optics kernel (the default when every RFM attenuator sits on its table's own spectral grid):
// a warp owns 32 (column, layer) points x all spectral points.
//   phase 1  lane = point: ln p, reference temperature, bracketing rows and blend weights of every attenuator
//            (harp::locate on axes in shared memory) -> a per-warp record in shared memory;
//   phase 2  lane = spectral point: for each of the 32 points the record is read as a broadcast and the four table
//            corners are ONE coalesced 256-byte load each (table transposed to (ln p, dT, wave)) -- the thread-per-point
//            form paid an L1 wavefront per lane for every gather (l1tex 68-88 % busy, FP64 pipe 10-19 %);
//   phase 3  the 32 x 32 (wave, point) tile goes through shared memory so that prop (wave, point) is written with
//            coalesced 256-byte stores.
// Same arithmetic, statement by statement, as the hoisted-weight path of band_optics_fused_kernel.
// ---------------------------------------------------------------------------
constexpr int kTileWarps = 4;
#ifndef B200_TILE_ILP
#define B200_TILE_ILP 4
#endif
constexpr int kPtIlp = B200_TILE_ILP;  // points of a tile evaluated side by side (independent exp chains)
struct TileRec {      // one (point, attenuator)
  int o11, o12, o21, o22;          // element offsets of the four (ln p, dT) corners in the transposed table
  double f11, f12, f21, f22, cs;   // their weights; 1e-3 conc (RFM) or conc (aerosol)
};

__global__ void __launch_bounds__(kTileWarps * 32) band_optics_tile_kernel(const BandOpticsParams P, long ntile) {
  extern __shared__ double b200_axes[];
  for (int a = 0; a < P.natt; ++a) {
    const DevAtt& A = P.att[a];
    if (A.type != B200_ATT_RFM || A.same_axes) continue;
    double* s = b200_axes + A.sm_off;
    const int n = P.nlyr;
    for (int k = threadIdx.x; k < 2 * n + A.n1 + A.n2; k += blockDim.x)
      s[k] = k < n ? A.klnp[k] : k < 2 * n ? A.ktemp[k - n] : A.kaxis[A.n0 + (k - 2 * n)];
  }
  __syncthreads();
  const int lane = (int)(threadIdx.x & 31), wib = (int)(threadIdx.x >> 5);
  // per-warp shared memory behind the axes: records [natt][32], dz [32], tile(s) [nprop][32][33]
  const int rec_doubles = (int)(sizeof(TileRec) / sizeof(double));
  const int per_warp = P.natt * 32 * rec_doubles + 32 + P.nprop * 32 * 33;
  double* wsm = b200_axes + ((P.sm_doubles + 3) & ~3) + (long)wib * per_warp;
  TileRec* recs = reinterpret_cast<TileRec*>(wsm);
  double* dzs = wsm + P.natt * 32 * rec_doubles;
  double* tile = dzs + 32;
  for (long t = (long)blockIdx.x * kTileWarps + wib; t < ntile; t += (long)gridDim.x * kTileWarps) {
    const long i0 = t * 32;
    const long i = i0 + lane;
    const bool valid = i < P.npoint;
    // ---- phase 1: this lane's point
    {
      double lnp = 0.0, dT = 0.0, wp1 = 0.0, wp2 = 0.0, wt1 = 0.0, wt2 = 0.0;
      int ip1 = 0, ip2 = 0, it1 = 0, it2 = 0;
      const long ii = valid ? i : P.npoint - 1;
      for (int a = 0; a < P.natt; ++a) {
        const DevAtt& A = P.att[a];
        TileRec r;
        const double cv = P.conc[ii * P.nspecies + A.sid];
        if (A.type == B200_ATT_AEROSOL) {
          r.o11 = r.o12 = r.o21 = r.o22 = 0;
          r.f11 = r.f12 = r.f21 = r.f22 = 0.0;
          r.cs = cv;
        } else {
          const double* sax = b200_axes + A.sm_off;
          if (!A.same_axes) {
            lnp = log(P.pres[ii]);
            Dims rd{};
            rd.len[0] = P.nlyr; rd.stride[0] = 1; rd.axis_off[0] = 0;
            double tref;
            interpn_dev<1, 1>(&tref, &lnp, sax + P.nlyr, sax, rd, 0);
            dT = __dsub_rn(P.temp[ii], tref);
            bracket_dev(sax + 2 * P.nlyr, A.n1, lnp, ip1, ip2, wp1, wp2);
            bracket_dev(sax + 2 * P.nlyr + A.n1, A.n2, dT, it1, it2, wt1, wt2);
          }
          r.o11 = (ip1 * A.n2 + it1) * A.n0; r.o12 = (ip1 * A.n2 + it2) * A.n0;
          r.o21 = (ip2 * A.n2 + it1) * A.n0; r.o22 = (ip2 * A.n2 + it2) * A.n0;
          r.f11 = wp1 * wt1; r.f12 = wp1 * wt2; r.f21 = wp2 * wt1; r.f22 = wp2 * wt2;
          r.cs = 1.0e-3 * cv;
        }
        recs[a * 32 + lane] = r;
      }
      dzs[lane] = P.dz ? P.dz[(int)(ii % P.nlyr)] : 1.0;
    }
    __syncwarp();
    const int npt = (int)min((long)32, P.npoint - i0);
    for (int wb = 0; wb < P.nwave; wb += 32) {
      const int wv = wb + lane;
      const bool wok = wv < P.nwave;
      const int wc = wok ? wv : P.nwave - 1;
      // ---- phase 2: this lane's spectral point, every point of the tile -- kPtIlp points at a time, so that as many
      //      independent exp() chains per attenuator are in flight (the kernel is bound by their latency)
      for (int p0 = 0; p0 < npt; p0 += kPtIlp) {
        double ext[kPtIlp], sse[kPtIlp];
#pragma unroll
        for (int u = 0; u < kPtIlp; ++u) { ext[u] = 0.0; sse[u] = 0.0; }
        for (int a = 0; a < P.natt; ++a) {
          const DevAtt& A = P.att[a];
          if (A.type == B200_ATT_AEROSOL) {
            const double k0 = A.kw[2 * wc], k1 = A.kw[2 * wc + 1];
#pragma unroll
            for (int u = 0; u < kPtIlp; ++u) {
              const double e = __dmul_rn(k0, recs[a * 32 + min(p0 + u, 31)].cs);
              ext[u] = __dadd_rn(ext[u], e);
              sse[u] = __dadd_rn(sse[u], __dmul_rn(k1, e));
            }
          } else {
            const double* T = A.kdata_t + wc;
            double v[kPtIlp];
#pragma unroll
            for (int u = 0; u < kPtIlp; ++u) {
              const TileRec& r = recs[a * 32 + min(p0 + u, 31)];
              v[u] = fma(r.f11, T[r.o11], fma(r.f12, T[r.o12], fma(r.f21, T[r.o21], r.f22 * T[r.o22])));
            }
#pragma unroll
            for (int u = 0; u < kPtIlp; ++u) ext[u] = fma(exp(v[u]), recs[a * 32 + min(p0 + u, 31)].cs, ext[u]);
          }
        }
#pragma unroll
        for (int u = 0; u < kPtIlp; ++u) {
          const int pt = p0 + u;
          if (pt < npt) {
            const double dzl = dzs[pt];
            const double tau = __dmul_rn(ext[u], dzl);
            tile[lane * 33 + pt] = tau;
            if (P.nprop > 1) {
              double ssa;
              if (P.eps > 0.0) ssa = __ddiv_rn(sse[u], __dadd_rn(ext[u], P.eps));
              else ssa = (tau == 0.0) ? 0.0 : __ddiv_rn(__dmul_rn(sse[u], dzl), tau);
              tile[32 * 33 + lane * 33 + pt] = fmin(ssa, 1.0);
            }
          }
        }
      }
      __syncwarp();
      // ---- phase 3: rows of the tile (one spectral point each) -> prop, lanes = points
      const int nwv = min(32, P.nwave - wb);
      if (valid) {
        if (P.nprop == 1) {
          for (int r = 0; r < nwv; ++r) P.prop[(long)(wb + r) * P.npoint + i] = tile[r * 33 + lane];
        } else {
          double2* o = reinterpret_cast<double2*>(P.prop);
          for (int r = 0; r < nwv; ++r) o[(long)(wb + r) * P.npoint + i] = make_double2(tile[r * 33 + lane], tile[32 * 33 + r * 33 + lane]);
        }
      }
      __syncwarp();
    }
  }
}

// (k_ext, ssa) of an aerosol table at every spectral point: harp::interpn<1> with nval = 2, once per band
__global__ void aerosol_wave_kernel(const double* __restrict__ kwave, const double* __restrict__ kdata, int rows,
                                    const double* __restrict__ wave, int nwave, int is_wavenumber, double* __restrict__ kw) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= nwave) return;
  Dims d{};
  d.len[0] = rows; d.stride[0] = 2; d.axis_off[0] = 0;
  double c = is_wavenumber ? __ddiv_rn(1.0e4, wave[w]) : wave[w], k[2];
  interpn_dev<1, 2>(k, &c, kdata, kwave, d, 0);
  kw[2 * w] = k[0];
  kw[2 * w + 1] = k[1];
}

// acc_t (rows, ncol) -> bflx (ncol, rows): the band-flux buffer the solve kernel accumulates into is column-fastest
// (coalesced reductions), pyharp's tensor is (ncol, nlyr+1, 2)
// (nrep replicas of acc, added in replica order: the warp-reduced form of the solve kernel spreads its reductions)
__global__ void band_finish_kernel(const double* __restrict__ acc, int rows, int ncol, int nrep, double* __restrict__ bflx) {
  __shared__ double tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int dr = threadIdx.y; dr < 32; dr += blockDim.y) {
    const int r = r0 + dr, c = c0 + threadIdx.x;
    if (r < rows && c < ncol) {
      double v = 0.0;
      for (int k = 0; k < nrep; ++k) v += acc[((long)k * rows + r) * ncol + c];
      tile[dr][threadIdx.x] = v;
    }
  }
  __syncthreads();
  for (int dc = threadIdx.y; dc < 32; dc += blockDim.y) {
    const int c = c0 + dc, r = r0 + threadIdx.x;
    if (r < rows && c < ncol) bflx[(long)c * rows + r] = tile[threadIdx.x][dc];
  }
}

// the same for the replicated buffer of the warp-reduced form (fewer than 32 columns, nrep = 64 replicas): one WARP per
// output element, lanes over the replicas, shuffle tree.  The tile kernel above walked the replicas serially in a handful
// of threads -- 64 dependent L2 round trips, ~50 us of a 135 us line-by-line call (profiles/r02/sweeps.md, e47)
__global__ void band_finish_rep_kernel(const double* __restrict__ acc, int rows, int ncol, int nrep, double* __restrict__ bflx) {
  const int lane = threadIdx.x & 31;
  const long wid = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (wid >= (long)rows * ncol) return;
  const int r = (int)(wid / ncol), c = (int)(wid - (long)r * ncol);
  double v = 0.0;
  for (int k = lane; k < nrep; k += 32) v += acc[((long)k * rows + r) * ncol + c];
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (lane == 0) bflx[(long)c * rows + r] = v;
}

static void launch_band_finish(const double* acc, int rows, int ncol, int nrep, double* bflx, cudaStream_t st) {
  if (nrep > 1) {
    const long warps = (long)rows * ncol;
    band_finish_rep_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, st>>>(acc, rows, ncol, nrep, bflx);
  } else {
    band_finish_kernel<<<dim3((unsigned)((ncol + 31) / 32), (unsigned)((rows + 31) / 32)), dim3(32, 8), 0, st>>>(acc, rows, ncol, nrep, bflx);
  }
}

// heating rate from the band flux, examples/amars_sw.cpp:290-301 (same evaluation order); dz per layer
__global__ void band_heating_kernel(const double* __restrict__ bflx, const double* __restrict__ rho,
                                    const double* __restrict__ dz, double cp, long ncol, int nlyr, double* __restrict__ out) {
  const long n = ncol * nlyr;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const long j = i / nlyr;
    const int k = (int)(i - j * nlyr);
    const double* f = bflx + (j * (nlyr + 1) + k) * 2;
    const double df = f[0] - f[1], df1 = f[2] - f[3];
    out[i] = -(1.0 / (rho[i] * cp)) * (df1 - df) / dz[k];
  }
}

// harp::locate on the host (src/math/locate.h:14-42), for the per-band spectral rows of an RFM table
static int locate_host(const double* xx0, double x, int n) {
  const double* xx = xx0 - 1;
  int jl = 0, ju = n + 1;
  const bool ascnd = (xx[n] >= xx[1]);
  while (ju - jl > 1) {
    const int jm = (ju + jl) >> 1;
    if ((x >= xx[jm]) == ascnd) jl = jm;
    else ju = jm;
  }
  int j;
  if (x == xx[1]) j = 1;
  else if (x == xx[n]) j = n;
  else j = jl;
  return j - 1;
}

}  // namespace b200

using namespace b200;

struct b200_band {
  b200disort_handle* h = nullptr;
  int nwave = 0, ncol = 0, nlyr = 0, nlev = 0, nspecies = 0, nprop = 1, natt = 0;
  int l2l_order = 1, l2l_blower = 0, l2l_bupper = 1, exact = 0, host_chunk_cols = 0;
  double eps = 0.0;
  bool planck = false, need_pres = false, tile_smem_set = false;
  DevAtt att[B200_BAND_MAX_ATT];
  std::vector<void*> owned;  // device tables
  double* d_weights = nullptr;
  double* d_dz = nullptr;
  static constexpr int kSlots = 3;
  struct Slot {
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
    double* d_buf = nullptr;
    size_t d_bytes = 0;
    double* d_ws = nullptr;
    size_t ws_bytes = 0;
  } slots[kSlots];
  // scratch of the device-pointer entry (runs on the handle's stream)
  double* d_main = nullptr;
  size_t main_bytes = 0;
  double* d_ws_main = nullptr;
  size_t ws_main_bytes = 0;
  cudaEvent_t ev_pre = nullptr;
};

namespace {

// device buffers of one chunk of columns
struct Chunk {
  int cc = 0;
  const double *conc = nullptr, *pres = nullptr, *temp = nullptr, *rho = nullptr;
  double cp = 0.0;
  b200disort_bc bc{};
  int bc_bcast = 0;
  double *prop = nullptr, *temf = nullptr, *pkag = nullptr, *acc = nullptr;
  int32_t* status = nullptr;
  double *bflx = nullptr, *dTdt = nullptr, *flx = nullptr;
};

template <typename T>
int upload(b200_band* b, const T* src, size_t n, const T** dst) {
  T* d = nullptr;
  CU_TRY(cudaMalloc(&d, std::max<size_t>(1, n) * sizeof(T)));
  b->owned.push_back(d);
  CU_TRY(cudaMemcpy(d, src, n * sizeof(T), cudaMemcpyHostToDevice));
  *dst = d;
  return B200DISORT_OK;
}

int n_bc_fields(const b200disort_bc& bc, const double** f) {
  f[0] = bc.umu0; f[1] = bc.fbeam; f[2] = bc.albedo; f[3] = bc.btemp; f[4] = bc.ttemp; f[5] = bc.temis; f[6] = bc.fisot;
  int n = 0;
  for (int i = 0; i < 7; ++i) n += f[i] ? 1 : 0;
  return n;
}

// sub-buffers of a slot start on 32-byte boundaries (double2 stores of prop, 256-bit loads)
constexpr size_t r4(size_t n) { return (n + 3) & ~(size_t)3; }

unsigned blocks_for(long n, int block, long cap) { return (unsigned)std::max<long>(1, std::min<long>((n + block - 1) / block, cap)); }

// the fused optics kernel for `cc` columns on `st`
int launch_optics(b200_band* b, cudaStream_t st, int cc, const double* conc, const double* pres, const double* temp,
                const double* d_dz, double* prop) {
  b200disort_handle* h = b->h;
  const int nlyr = b->nlyr, nwave = b->nwave;
  BandOpticsParams P{};
  for (int a = 0; a < b->natt; ++a) P.att[a] = b->att[a];
  P.natt = b->natt; P.nwave = nwave; P.nlyr = nlyr; P.nspecies = b->nspecies; P.nprop = b->nprop; P.exact = b->exact;
  P.npoint = (long)cc * nlyr;
  P.conc = conc; P.pres = pres; P.temp = temp; P.dz = d_dz;
  P.eps = b->eps;
  P.prop = prop;
  const unsigned gy = blocks_for(P.npoint, 256, 65535);
  // shared memory of the interpolation axes: offsets per RFM attenuator (those on the previous one's axes share)
  int smd = 0;
  for (int a = 0; a < b->natt; ++a) {
    if (P.att[a].type != B200_ATT_RFM) continue;
    if (P.att[a].same_axes) { P.att[a].sm_off = P.att[a - 1].sm_off; continue; }
    P.att[a].sm_off = smd;
    smd += 2 * nlyr + P.att[a].n1 + P.att[a].n2;
  }
  P.sm_doubles = smd;
  const size_t sm = (size_t)smd * sizeof(double);
  // tile form: hoisted weights, every RFM attenuator on its table's own spectral grid
  {
    static const int env_tile = std::getenv("B200_BAND_TILE") ? std::atoi(std::getenv("B200_BAND_TILE")) : 1;  // profiling knob
    bool ok = env_tile != 0 && !b->exact && P.npoint >= 32;
    for (int a = 0; a < b->natt; ++a)
      if (P.att[a].type == B200_ATT_RFM && (!P.att[a].on_grid || (long)P.att[a].n0 * P.att[a].n1 * P.att[a].n2 >= (1L << 31))) ok = false;
    const int per_warp = b->natt * 32 * (int)(sizeof(TileRec) / sizeof(double)) + 32 + b->nprop * 32 * 33;
    const size_t smt = ((size_t)((smd + 3) & ~3) + (size_t)kTileWarps * per_warp) * sizeof(double);
    if (ok && smt <= 200 * 1024) {
      if (smt > 48 * 1024 && !b->tile_smem_set) {
        CU_TRY(cudaFuncSetAttribute(band_optics_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        b->tile_smem_set = true;
      }
      const long ntile = (P.npoint + 31) / 32;
      const long grid = std::min<long>((ntile + kTileWarps - 1) / kTileWarps, (long)h->sm_count * 16);
      band_optics_tile_kernel<<<(unsigned)grid, kTileWarps * 32, smt, st>>>(P, ntile);
      CU_TRY(cudaGetLastError());
      h->stats.kernel_launches += 1;
      return B200DISORT_OK;
    }
  }
  // enough points to fill the machine with tiles of 8 spectral points per thread?  otherwise one per thread
  static const int env_wt = std::getenv("B200_BAND_WT") ? std::atoi(std::getenv("B200_BAND_WT")) : 0;  // profiling knob: 10 * WT + MINB
  const int sel = env_wt ? env_wt : (nwave >= 16 ? 162 : 82);
  const int wt = sel / 10;
  const bool big = P.npoint * ((nwave + wt - 1) / wt) >= (long)h->sm_count * 2048;
#define B200_OPT(WT_, MB_) band_optics_fused_kernel<WT_, MB_><<<dim3((unsigned)((nwave + WT_ - 1) / WT_), gy), 256, sm, st>>>(P)
  if (!big) B200_OPT(1, 2);
  else if (sel == 82) B200_OPT(8, 2);
  else if (sel == 83) B200_OPT(8, 3);
  else if (sel == 84) B200_OPT(8, 4);
  else if (sel == 163) B200_OPT(16, 3);
  else if (sel == 164) B200_OPT(16, 4);
  else if (sel == 322) B200_OPT(32, 2);
  else if (sel == 42) B200_OPT(4, 2);
  else if (sel == 44) B200_OPT(4, 4);
  else if (sel == 46) B200_OPT(4, 6);
  else B200_OPT(16, 2);
#undef B200_OPT
  CU_TRY(cudaGetLastError());
  h->stats.kernel_launches += 1;
  return B200DISORT_OK;
}

// warp-reduced band integration (emit_flux, BAND = 2): fewer than 32 columns, a power of two, nstr <= 8
constexpr int kBandReplicas = 64;
int band_warp_red(const b200disort_handle* h, int cc) {
  return (h->inst && h->inst->nn <= 4 && cc < 32 && (cc & (cc - 1)) == 0) ? cc : 0;
}
size_t band_acc_doubles(const b200disort_handle* h, int rows, int cc) {
  return (size_t)(band_warp_red(h, cc) ? kBandReplicas : 1) * rows * cc;
}

// every kernel of the band pipeline for one chunk of columns, on `st`
int run_chunk(b200_band* b, cudaStream_t st, const Chunk& c, const double* d_dz, double** ws, size_t* ws_bytes) {
  b200disort_handle* h = b->h;
  const int cc = c.cc, nlyr = b->nlyr, nwave = b->nwave;
  // ---- optical properties
  {
    int rc = launch_optics(b, st, cc, c.conc, c.pres, c.temp, d_dz, c.prop);
    if (rc != B200DISORT_OK) return rc;
  }
  // ---- interface temperatures and Planck emission per (band, column, level)
  if (b->planck) {
    if (b200_layer2level(c.temp, cc, nlyr, b->l2l_order, b->l2l_blower, b->l2l_bupper, c.temf, h->d_flag, st) != 0)
      return B200DISORT_ERR_CUDA;
    int rc = launch_planck_into(h, c.temf, cc, c.pkag, st);
    if (rc != B200DISORT_OK) return rc;
    h->stats.kernel_launches += 1;
  }
  // ---- solve
  SolveParams p{};
  p.nwave = nwave; p.ncol = cc; p.nlyr = nlyr; p.nprop = b->nprop; p.nstr = h->cfg.nstr;
  p.flags = h->cfg.flags;
  p.prop = c.prop;
  p.umu0 = c.bc.umu0; p.fbeam = c.bc.fbeam; p.albedo = c.bc.albedo; p.btemp = c.bc.btemp;
  p.ttemp = c.bc.ttemp; p.temis = c.bc.temis; p.fisot = c.bc.fisot;
  p.bc_bcast = c.bc_bcast;
  p.pkag = c.pkag;
  p.wave2band = h->d_w2b; p.wave_lower = h->d_wlo; p.wave_upper = h->d_whi;
  p.status = c.status;
  p.s0 = 0;
  p.nlocal = (long)nwave * cc;
  const int rows = b->nlev * 2;
  if (c.flx) {  // the caller wants the per-spectral-point fluxes: write them, integrate in a second pass
    p.flx = c.flx;
    int rc = launch_flux(h, p, st, ws, ws_bytes, true);
    if (rc != B200DISORT_OK) return rc;
    rc = launch_reduce_band(h, nwave, (long)cc * rows, c.flx, b->d_weights, c.bflx, st);
    if (rc != B200DISORT_OK) return rc;
    h->stats.kernel_launches += 1;
  } else {
    p.bflx_t = c.acc;
    p.weights = b->d_weights;
    p.warp_red = band_warp_red(h, cc);
    p.nrep = p.warp_red ? kBandReplicas : 1;
    p.rep_stride = (long)rows * cc;
    CU_TRY(cudaMemsetAsync(c.acc, 0, (size_t)p.nrep * rows * cc * sizeof(double), st));
    int rc = launch_flux(h, p, st, ws, ws_bytes, true);
    if (rc != B200DISORT_OK) return rc;
    launch_band_finish(c.acc, rows, cc, p.nrep, c.bflx, st);
    CU_TRY(cudaGetLastError());
    h->stats.kernel_launches += 1;
  }
  int rc = launch_count_status(h, c.status, p.nlocal, st);
  if (rc != B200DISORT_OK) return rc;
  h->stats.kernel_launches += 1;
  if (c.dTdt) {
    band_heating_kernel<<<blocks_for((long)cc * nlyr, 256, (long)h->sm_count * 8), 256, 0, st>>>(c.bflx, c.rho, d_dz, c.cp, cc, nlyr, c.dTdt);
    CU_TRY(cudaGetLastError());
  }
  return B200DISORT_OK;
}

}  // namespace

extern "C" int b200_band_create(const b200_band_config* cfg, b200_band** out) {
  if (!cfg || !out) return B200DISORT_ERR_INVALID;
  *out = nullptr;
  const b200disort_config& dc = cfg->disort;
  if (cfg->natt < 1 || cfg->natt > B200_BAND_MAX_ATT || !cfg->att || cfg->nspecies < 1 || !cfg->weights)
    return B200DISORT_ERR_INVALID;
  if (!(dc.flags & B200DISORT_ONLYFL) || (dc.flags & B200DISORT_USRTAU)) return B200DISORT_ERR_UNSUPPORTED;
  if (cfg->l2l_order < 0 || cfg->l2l_order > 1 || cfg->l2l_blower < 0 || cfg->l2l_blower > 1 || cfg->l2l_bupper < 0 ||
      cfg->l2l_bupper > 1 || !(cfg->mix_eps >= 0.0))
    return B200DISORT_ERR_INVALID;
  for (int a = 0; a < cfg->natt; ++a) {
    const b200_attenuator& A = cfg->att[a];
    if (A.species_id < 0 || A.species_id >= cfg->nspecies || !A.kdata) return B200DISORT_ERR_INVALID;
    if (A.type == B200_ATT_RFM) {
      if (!A.kaxis || !A.klnp || !A.ktemp || A.dims[0] != dc.nwave || A.dims[1] < 1 || A.dims[2] < 1) return B200DISORT_ERR_INVALID;
      if (dc.nlyr > A.nref) return B200DISORT_ERR_INVALID;  // get_reftemp reads nlyr entries of the reference grid
    } else if (A.type == B200_ATT_AEROSOL) {
      if (!A.kwave || A.dims[0] < 1 || !cfg->wave) return B200DISORT_ERR_INVALID;
    } else {
      return B200DISORT_ERR_INVALID;
    }
  }
  b200_band* b = new (std::nothrow) b200_band();
  if (!b) return B200DISORT_ERR_INVALID;
  auto fail = [&](int rc) {
    b200_band_destroy(b);
    return rc;
  };
  int rc = b200disort_create(&dc, &b->h);
  if (rc != B200DISORT_OK) return fail(rc);
  if (!b->h->inst) return fail(B200DISORT_ERR_UNSUPPORTED);  // stream counts without a flux-kernel instantiation
  b->nwave = dc.nwave; b->ncol = dc.ncol; b->nlyr = dc.nlyr; b->nlev = dc.nlyr + 1;
  b->nspecies = cfg->nspecies; b->natt = cfg->natt;
  b->l2l_order = cfg->l2l_order; b->l2l_blower = cfg->l2l_blower; b->l2l_bupper = cfg->l2l_bupper;
  b->exact = cfg->exact_interp; b->eps = cfg->mix_eps; b->host_chunk_cols = cfg->host_chunk_cols;
  b->planck = dc.flags & B200DISORT_PLANCK;
  const double* d_wave = nullptr;
  if (cfg->wave) {
    rc = upload(b, cfg->wave, (size_t)dc.nwave, &d_wave);
    if (rc != B200DISORT_OK) return fail(rc);
  }
  for (int a = 0; a < cfg->natt; ++a) {
    const b200_attenuator& A = cfg->att[a];
    DevAtt& D = b->att[a];
    std::memset(&D, 0, sizeof(D));
    D.type = A.type;
    D.sid = A.species_id;
    if (A.type == B200_ATT_RFM) {
      b->need_pres = true;
      D.n0 = (int)A.dims[0]; D.n1 = (int)A.dims[1]; D.n2 = (int)A.dims[2];
      rc = upload(b, A.kaxis, (size_t)(D.n0 + D.n1 + D.n2), &D.kaxis);
      if (rc == B200DISORT_OK) rc = upload(b, A.kdata, (size_t)D.n0 * D.n1 * D.n2, &D.kdata);
      if (rc == B200DISORT_OK) rc = upload(b, A.klnp, (size_t)A.nref, &D.klnp);
      if (rc == B200DISORT_OK) rc = upload(b, A.ktemp, (size_t)A.nref, &D.ktemp);
      // the spectral coordinate RFMImpl::forward interpolates at is the table's own axis (rfm.cpp:145): rows and
      // weights per spectral point, once
      std::vector<int> wi(2 * (size_t)D.n0);
      std::vector<double> ww(2 * (size_t)D.n0);
      for (int w = 0; w < D.n0; ++w) {
        const double cw = A.kaxis[w];
        int i1 = locate_host(A.kaxis, cw, D.n0), i2;
        if (i1 == -1) { i1 = 0; i2 = 0; }
        else if (i1 == D.n0 - 1) i2 = D.n0 - 1;
        else i2 = i1 + 1;
        const double x1 = A.kaxis[i1], x2 = A.kaxis[i2];
        wi[2 * w] = i1; wi[2 * w + 1] = i2;
        if (x2 != x1) { ww[2 * w] = (x2 - cw) / (x2 - x1); ww[2 * w + 1] = (cw - x1) / (x2 - x1); }
        else { ww[2 * w] = 1.0; ww[2 * w + 1] = 0.0; }  // (v1 + v2) / 2 of the same row
      }
      // same axes and reference atmosphere as the previous attenuator (two species of one table file)?
      D.same_axes = 0;
      if (a > 0 && cfg->att[a - 1].type == B200_ATT_RFM) {
        const b200_attenuator& Pv = cfg->att[a - 1];
        bool same = Pv.dims[0] == A.dims[0] && Pv.dims[1] == A.dims[1] && Pv.dims[2] == A.dims[2] && Pv.nref == A.nref;
        if (same) same = std::memcmp(Pv.kaxis, A.kaxis, sizeof(double) * (size_t)(D.n0 + D.n1 + D.n2)) == 0 &&
                         std::memcmp(Pv.klnp, A.klnp, sizeof(double) * (size_t)dc.nlyr) == 0 &&
                         std::memcmp(Pv.ktemp, A.ktemp, sizeof(double) * (size_t)dc.nlyr) == 0;
        D.same_axes = same ? 1 : 0;
      }
      D.on_grid = 1;
      for (int w = 0; w < D.n0; ++w)
        if (wi[2 * w] != w || ww[2 * w] != 1.0 || ww[2 * w + 1] != 0.0) D.on_grid = 0;
      {
        std::vector<double> kt((size_t)D.n0 * D.n1 * D.n2);
        for (int w = 0; w < D.n0; ++w)
          for (long q = 0; q < (long)D.n1 * D.n2; ++q) kt[(size_t)q * D.n0 + w] = A.kdata[(size_t)w * D.n1 * D.n2 + q];
        if (rc == B200DISORT_OK) rc = upload(b, kt.data(), kt.size(), &D.kdata_t);
      }
      if (rc == B200DISORT_OK) rc = upload(b, wi.data(), wi.size(), &D.wi);
      if (rc == B200DISORT_OK) rc = upload(b, ww.data(), ww.size(), &D.ww);
      if (rc != B200DISORT_OK) return fail(rc);
    } else {
      b->nprop = 2;
      const int rows = (int)A.dims[0];
      const double *d_kwave = nullptr, *d_kdata = nullptr;
      rc = upload(b, A.kwave, (size_t)rows, &d_kwave);
      if (rc == B200DISORT_OK) rc = upload(b, A.kdata, (size_t)rows * 2, &d_kdata);
      if (rc != B200DISORT_OK) return fail(rc);
      double* kw = nullptr;
      if (cudaMalloc(&kw, sizeof(double) * 2 * dc.nwave) != cudaSuccess) return fail(B200DISORT_ERR_CUDA);
      b->owned.push_back(kw);
      aerosol_wave_kernel<<<(dc.nwave + 127) / 128, 128, 0, b->h->stream>>>(d_kwave, d_kdata, rows, d_wave, dc.nwave,
                                                                            cfg->wave_is_wavenumber, kw);
      if (cudaGetLastError() != cudaSuccess) return fail(B200DISORT_ERR_CUDA);
      D.kw = kw;
    }
  }
  {
    const double* dw = nullptr;
    rc = upload(b, cfg->weights, (size_t)dc.nwave, &dw);
    if (rc != B200DISORT_OK) return fail(rc);
    b->d_weights = const_cast<double*>(dw);
    if (cudaMalloc(&b->d_dz, sizeof(double) * dc.nlyr) != cudaSuccess) return fail(B200DISORT_ERR_CUDA);
  }
  if (cudaStreamSynchronize(b->h->stream) != cudaSuccess) return fail(B200DISORT_ERR_CUDA);
  *out = b;
  return B200DISORT_OK;
}

extern "C" void b200_band_destroy(b200_band* b) {
  if (!b) return;
  if (b->h) {
    cudaSetDevice(b->h->cfg.device);
    cudaStreamSynchronize(b->h->stream);
  }
  for (auto& sl : b->slots) {
    if (sl.stream) { cudaStreamSynchronize(sl.stream); cudaStreamDestroy(sl.stream); }
    if (sl.done) cudaEventDestroy(sl.done);
    cudaFree(sl.d_buf);
    cudaFree(sl.d_ws);
  }
  for (void* p : b->owned) cudaFree(p);
  cudaFree(b->d_dz);
  cudaFree(b->d_main);
  cudaFree(b->d_ws_main);
  if (b->ev_pre) cudaEventDestroy(b->ev_pre);
  if (b->h) b200disort_destroy(b->h);
  delete b;
}

extern "C" int b200_band_nprop(const b200_band* b) { return b ? b->nprop : 0; }
extern "C" b200disort_handle* b200_band_solver(b200_band* b) { return b ? b->h : nullptr; }

static int check_io(const b200_band* b, const b200_band_io* io, bool host) {
  if (!b || !io || !io->conc || !io->bflx) return B200DISORT_ERR_INVALID;
  if (b->need_pres && (!io->pres || !io->temp)) return B200DISORT_ERR_INVALID;
  if (b->planck && !io->temp) return B200DISORT_ERR_INVALID;
  if (io->dTdt && (!io->rho || !io->dz || !(io->cp > 0.0))) return B200DISORT_ERR_INVALID;
  if (host && (io->flx || io->prop)) return B200DISORT_ERR_UNSUPPORTED;
  return B200DISORT_OK;
}

extern "C" int b200_band_forward(b200_band* b, const b200_band_io* io) {
  int rc = check_io(b, io, false);
  if (rc != B200DISORT_OK) return rc;
  b200disort_handle* h = b->h;
  CU_TRY(cudaSetDevice(h->cfg.device));
  h->stats.kernel_launches = 0;
  cudaStream_t st = h->stream;
  CU_TRY(cudaEventRecord(h->ev0, st));
  const int ncol = b->ncol, nlyr = b->nlyr, nwave = b->nwave;
  const size_t ns = (size_t)nwave * ncol;
  // scratch: [prop?] [temf] [pkag] [acc] [status?]
  const size_t n_prop = io->prop ? 0 : r4(ns * nlyr * b->nprop);
  const size_t n_temf = b->planck ? r4((size_t)ncol * (nlyr + 1)) : 0, n_pkag = b->planck ? r4((size_t)h->nband * ncol * (nlyr + 1)) : 0;
  const size_t n_acc = io->flx ? 0 : r4(band_acc_doubles(h, b->nlev * 2, ncol)), n_st = io->status ? 0 : r4((ns * sizeof(int32_t) + 7) / 8);
  rc = ensure_dev(&b->d_main, &b->main_bytes, (n_prop + n_temf + n_pkag + n_acc + n_st + 1) * sizeof(double));
  if (rc != B200DISORT_OK) return rc;
  const double* d_dz = nullptr;
  if (io->dz) {
    CU_TRY(cudaMemcpyAsync(b->d_dz, io->dz, sizeof(double) * nlyr, cudaMemcpyDeviceToDevice, st));
    d_dz = b->d_dz;
  }
  Chunk c;
  c.cc = ncol;
  c.conc = io->conc; c.pres = io->pres; c.temp = io->temp; c.rho = io->rho; c.cp = io->cp;
  c.bc = io->bc; c.bc_bcast = io->bc_per_column;
  double* d = b->d_main;
  c.prop = io->prop ? io->prop : d; d += n_prop;
  c.temf = d; d += n_temf;
  c.pkag = d; d += n_pkag;
  c.acc = d; d += n_acc;
  c.status = io->status ? io->status : reinterpret_cast<int32_t*>(d);
  c.bflx = io->bflx; c.dTdt = io->dTdt; c.flx = io->flx;
  return run_chunk(b, st, c, d_dz, &b->d_ws_main, &b->ws_main_bytes);
}

extern "C" int b200_band_optical_props(b200_band* b, const b200_band_io* io) {
  if (!b || !io || !io->conc || !io->prop) return B200DISORT_ERR_INVALID;
  if (b->need_pres && (!io->pres || !io->temp)) return B200DISORT_ERR_INVALID;
  b200disort_handle* h = b->h;
  CU_TRY(cudaSetDevice(h->cfg.device));
  const double* d_dz = nullptr;
  if (io->dz) {
    CU_TRY(cudaMemcpyAsync(b->d_dz, io->dz, sizeof(double) * b->nlyr, cudaMemcpyDeviceToDevice, h->stream));
    d_dz = b->d_dz;
  }
  return launch_optics(b, h->stream, b->ncol, io->conc, io->pres, io->temp, d_dz, io->prop);
}

extern "C" int b200_band_check(b200_band* b, int* n_bad) {
  if (!b) return B200DISORT_ERR_INVALID;
  b200disort_handle* h = b->h;
  CU_TRY(cudaSetDevice(h->cfg.device));
  for (auto& sl : b->slots)
    if (sl.stream) CU_TRY(cudaStreamSynchronize(sl.stream));
  int flags[2] = {0, 0};
  CU_TRY(cudaMemcpyAsync(flags, h->d_flag, sizeof(int) * 2, cudaMemcpyDeviceToHost, h->stream));
  CU_TRY(cudaStreamSynchronize(h->stream));
  if (n_bad) *n_bad = flags[1];
  if (flags[0] || flags[1]) {
    CU_TRY(cudaMemsetAsync(h->d_flag, 0, sizeof(int) * 2, h->stream));
    return B200DISORT_ERR_SOLVE;
  }
  return B200DISORT_OK;
}

constexpr long kChunkWaves = 4;

static int band_forward_host_impl(b200_band* b, const b200_band_io* io);

extern "C" int b200_band_forward_host(b200_band* b, const b200_band_io* io) {
  const int rc = band_forward_host_impl(b, io);
  if (b && rc != B200DISORT_OK && rc != B200DISORT_ERR_SOLVE && rc != B200DISORT_ERR_INVALID && rc != B200DISORT_ERR_UNSUPPORTED) {
    // an error return must not leave copies in flight that still read or write the caller's buffers
    for (auto& sl : b->slots)
      if (sl.stream) cudaStreamSynchronize(sl.stream);
    cudaStreamSynchronize(b->h->stream);
  }
  return rc;
}

static int band_forward_host_impl(b200_band* b, const b200_band_io* io) {
  int rc = check_io(b, io, true);
  if (rc != B200DISORT_OK) return rc;
  b200disort_handle* h = b->h;
  CU_TRY(cudaSetDevice(h->cfg.device));
  h->stats.kernel_launches = 0;
  const int ncol = b->ncol, nlyr = b->nlyr, nwave = b->nwave, nlev = b->nlev, nsp = b->nspecies;
  // chunks of columns, up to kChunkWaves resident waves of the flux kernel each: the kernels of successive chunks run back
  // to back while the copy engines move their neighbours, and with several solves per thread the back sweep of one
  // rides on the forward sweep of the next (carried back sweep, disort_device.cuh).  The first chunks are smaller
  // (1, 2, then 4 waves): nothing overlaps the first chunk's copy-in.
  if (h->occ == 0) {
    int occ = 0;
    CU_TRY(h->inst->occupancy(128, &occ));
    h->occ = occ < 1 ? 1 : occ;
  }
  const long wave = (long)h->sm_count * h->occ * 128;
  static const int env_cc = std::getenv("B200_BAND_CHUNK_COLS") ? std::atoi(std::getenv("B200_BAND_CHUNK_COLS")) : 0;  // profiling knob
  long cc = b->host_chunk_cols > 0 ? b->host_chunk_cols : env_cc > 0 ? env_cc : (kChunkWaves * wave + nwave - 1) / nwave;
  if (cc >= 32 && b->host_chunk_cols <= 0) cc = (cc + 31) / 32 * 32;  // whole warps per spectral point
  cc = std::max<long>(1, std::min<long>(cc, ncol));
  // chunk schedule: start column of every chunk (+ the end)
  std::vector<long> cstart;
  {
    const bool ramp = b->host_chunk_cols <= 0 && env_cc <= 0 && cc >= 4 * 32;
    long c0 = 0, k = 0;
    while (c0 < ncol) {
      cstart.push_back(c0);
      long n = cc;
      if (ramp && k == 0) n = std::max<long>(32, cc / 4 / 32 * 32);
      else if (ramp && k == 1) n = std::max<long>(32, cc / 2 / 32 * 32);
      c0 += n;
      ++k;
    }
    cstart.push_back(ncol);
  }
  const long nchunk = (long)cstart.size() - 1;

  const double* f[7];
  const int n_bc = n_bc_fields(io->bc, f);
  const size_t bc_len = io->bc_per_column ? (size_t)cc : (size_t)nwave * cc;
  const size_t ns = (size_t)nwave * cc;
  // slot layout (doubles): inputs | scratch | outputs
  const size_t n_conc = r4((size_t)cc * nlyr * nsp), n_pt = r4((size_t)cc * nlyr);
  const size_t n_in = n_conc + (io->pres ? n_pt : 0) + (io->temp ? n_pt : 0) + (io->rho ? n_pt : 0) + n_bc * r4(bc_len);
  const size_t n_prop = r4(ns * nlyr * b->nprop), n_temf = b->planck ? r4((size_t)cc * (nlyr + 1)) : 0;
  size_t acc_max = 0;  // (the last chunk may be a narrower, warp-reduced one with replicas)
  for (long k = 0; k + 1 < (long)cstart.size(); ++k)
    acc_max = std::max(acc_max, band_acc_doubles(h, nlev * 2, (int)(cstart[k + 1] - cstart[k])));
  const size_t n_pkag = b->planck ? r4((size_t)h->nband * cc * (nlyr + 1)) : 0, n_acc = r4(acc_max);
  const size_t n_bflx = r4((size_t)cc * nlev * 2), n_dT = io->dTdt ? n_pt : 0, n_st = r4((ns * sizeof(int32_t) + 7) / 8);
  const size_t slot_doubles = n_in + n_prop + n_temf + n_pkag + n_acc + n_bflx + n_dT + n_st;

  if (!b->ev_pre) CU_TRY(cudaEventCreateWithFlags(&b->ev_pre, cudaEventDisableTiming));
  const int nslot = (int)std::min<long>(b200_band::kSlots, nchunk);
  for (int k = 0; k < nslot; ++k) {
    b200_band::Slot& sl = b->slots[k];
    if (!sl.stream) CU_TRY(cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking));
    if (!sl.done) CU_TRY(cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming));
    rc = ensure_dev(&sl.d_buf, &sl.d_bytes, slot_doubles * sizeof(double));
    if (rc != B200DISORT_OK) return rc;
  }
  const double* d_dz = nullptr;
  if (io->dz) {
    CU_TRY(cudaMemcpyAsync(b->d_dz, io->dz, sizeof(double) * nlyr, cudaMemcpyHostToDevice, h->stream));
    d_dz = b->d_dz;
  }
  CU_TRY(cudaEventRecord(h->ev0, h->stream));
  CU_TRY(cudaEventRecord(b->ev_pre, h->stream));

  for (long ck = 0; ck < nchunk; ++ck) {
    b200_band::Slot& sl = b->slots[ck % nslot];
    cudaStream_t st = sl.stream;
    if (ck >= nslot) CU_TRY(cudaStreamWaitEvent(st, sl.done, 0));  // (same stream: already ordered; kept for clarity)
    else CU_TRY(cudaStreamWaitEvent(st, b->ev_pre, 0));
    const long c0 = cstart[ck];
    const int n = (int)(cstart[ck + 1] - c0);
    Chunk c;
    c.cc = n;
    c.cp = io->cp;
    c.bc_bcast = io->bc_per_column;
    double* d = sl.d_buf;
    auto h2d = [&](const double* src, size_t count) -> const double* {
      double* dst = d;
      d += r4(count);
      if (cudaMemcpyAsync(dst, src, count * sizeof(double), cudaMemcpyHostToDevice, st) != cudaSuccess) return nullptr;
      return dst;
    };
    c.conc = h2d(io->conc + (size_t)c0 * nlyr * nsp, (size_t)n * nlyr * nsp);
    if (!c.conc) return B200DISORT_ERR_CUDA;
    if (io->pres && !(c.pres = h2d(io->pres + (size_t)c0 * nlyr, (size_t)n * nlyr))) return B200DISORT_ERR_CUDA;
    if (io->temp && !(c.temp = h2d(io->temp + (size_t)c0 * nlyr, (size_t)n * nlyr))) return B200DISORT_ERR_CUDA;
    if (io->rho && !(c.rho = h2d(io->rho + (size_t)c0 * nlyr, (size_t)n * nlyr))) return B200DISORT_ERR_CUDA;
    {
      const double** dst[7] = {&c.bc.umu0, &c.bc.fbeam, &c.bc.albedo, &c.bc.btemp, &c.bc.ttemp, &c.bc.temis, &c.bc.fisot};
      for (int i = 0; i < 7; ++i) {
        if (!f[i]) continue;
        if (io->bc_per_column) {
          if (!(*dst[i] = h2d(f[i] + c0, (size_t)n))) return B200DISORT_ERR_CUDA;
        } else {  // (nwave, ncol) -> (nwave, n): one row per spectral point
          CU_TRY(cudaMemcpy2DAsync(d, (size_t)n * sizeof(double), f[i] + c0, (size_t)ncol * sizeof(double),
                                   (size_t)n * sizeof(double), (size_t)nwave, cudaMemcpyHostToDevice, st));
          *dst[i] = d;
          d += r4((size_t)nwave * n);
        }
      }
    }
    d = sl.d_buf + n_in;
    c.prop = d; d += n_prop;
    c.temf = d; d += n_temf;
    c.pkag = d; d += n_pkag;
    c.acc = d; d += n_acc;
    c.bflx = d; d += n_bflx;
    c.dTdt = io->dTdt ? d : nullptr; d += n_dT;
    c.status = reinterpret_cast<int32_t*>(d);
    rc = run_chunk(b, st, c, d_dz, &sl.d_ws, &sl.ws_bytes);
    if (rc != B200DISORT_OK) return rc;
    CU_TRY(cudaMemcpyAsync(io->bflx + (size_t)c0 * nlev * 2, c.bflx, (size_t)n * nlev * 2 * sizeof(double),
                           cudaMemcpyDeviceToHost, st));
    if (io->dTdt)
      CU_TRY(cudaMemcpyAsync(io->dTdt + (size_t)c0 * nlyr, c.dTdt, (size_t)n * nlyr * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (io->status)
      CU_TRY(cudaMemcpy2DAsync(io->status + c0, (size_t)ncol * sizeof(int32_t), c.status, (size_t)n * sizeof(int32_t),
                               (size_t)n * sizeof(int32_t), (size_t)nwave, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaEventRecord(sl.done, st));
  }
  return b200_band_check(b, nullptr);
}

// ---------------------------------------------------------------------------
// NCCL, resolved at run time: the communicator comes from the caller, so the calls must go to the libnccl that
// created it -- the one already loaded in the process (torch's bundled copy or the system library).
// ---------------------------------------------------------------------------
namespace {
typedef int (*nccl_allreduce_fn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
nccl_allreduce_fn resolve_nccl_allreduce() {
  static nccl_allreduce_fn fn = nullptr;
  static bool tried = false;
  if (tried) return fn;
  tried = true;
  void* sym = dlsym(RTLD_DEFAULT, "ncclAllReduce");
  if (!sym) {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
      void* lib = dlopen(nm, RTLD_NOW | RTLD_NOLOAD);
      if (!lib) continue;
      if ((sym = dlsym(lib, "ncclAllReduce"))) break;
    }
  }
  if (!sym) {
    if (void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL)) sym = dlsym(lib, "ncclAllReduce");
  }
  fn = reinterpret_cast<nccl_allreduce_fn>(sym);
  return fn;
}

int allreduce_sum_f64(double* buf, long count, void* comm, cudaStream_t st) {
  if (!buf || count < 1 || !comm) return B200DISORT_ERR_INVALID;
  nccl_allreduce_fn fn = resolve_nccl_allreduce();
  if (!fn) {
    std::fprintf(stderr, "b200disort: libnccl is not loaded in this process (ncclAllReduce not found)\n");
    return B200DISORT_ERR_UNSUPPORTED;
  }
  const int kNcclFloat64 = 8, kNcclSum = 0;  // nccl.h: ncclDataType_t / ncclRedOp_t
  const int rc = fn(buf, buf, (size_t)count, kNcclFloat64, kNcclSum, comm, st);
  if (rc != 0) {
    std::fprintf(stderr, "b200disort: ncclAllReduce failed with ncclResult_t %d\n", rc);
    return B200DISORT_ERR_CUDA;
  }
  return B200DISORT_OK;
}
}  // namespace

extern "C" int b200_band_allreduce(b200_band* b, double* bflx, long count, void* nccl_comm) {
  if (!b) return B200DISORT_ERR_INVALID;
  CU_TRY(cudaSetDevice(b->h->cfg.device));
  return allreduce_sum_f64(bflx, count, nccl_comm, b->h->stream);
}

extern "C" int b200disort_allreduce_band(b200disort_handle* h, double* bflx, void* nccl_comm) {
  if (!h) return B200DISORT_ERR_INVALID;
  CU_TRY(cudaSetDevice(h->cfg.device));
  const b200disort_config& c = h->cfg;
  const int nlev = (c.flags & B200DISORT_USRTAU) ? c.ntau : c.nlyr + 1;
  return allreduce_sum_f64(bflx, (long)c.ncol * nlev * 2, nccl_comm, h->stream);
}

extern "C" int b200disort_solve_band(b200disort_handle* h, const double* prop, int nprop, const b200disort_bc* bc,
                                     const double* temf, const double* weights, double* bflx, int32_t* status) {
  if (!h || !prop || !weights || !bflx || !status || nprop < 1) return B200DISORT_ERR_INVALID;
  const b200disort_config& c = h->cfg;
  const bool planck = c.flags & B200DISORT_PLANCK;
  if (!(c.flags & B200DISORT_ONLYFL) || (c.flags & B200DISORT_USRTAU) || !h->inst) return B200DISORT_ERR_UNSUPPORTED;
  if (planck && !temf) return B200DISORT_ERR_INVALID;
  CU_TRY(cudaSetDevice(c.device));
  h->stats.kernel_launches = 0;
  cudaStream_t st = h->stream;
  CU_TRY(cudaEventRecord(h->ev0, st));
  const int rows = (c.nlyr + 1) * 2;
  const size_t n_pkag = planck ? (size_t)h->nband * c.ncol * (c.nlyr + 1) : 0;
  const size_t n_acc = band_acc_doubles(h, rows, c.ncol);
  int rc = ensure_dev(&h->d_acc, &h->acc_bytes, (n_acc + n_pkag) * sizeof(double));
  if (rc != B200DISORT_OK) return rc;
  SolveParams p{};
  p.nwave = c.nwave; p.ncol = c.ncol; p.nlyr = c.nlyr; p.nprop = nprop; p.nstr = c.nstr;
  p.flags = c.flags;
  p.prop = prop;
  if (bc) {
    p.umu0 = bc->umu0; p.fbeam = bc->fbeam; p.albedo = bc->albedo; p.btemp = bc->btemp;
    p.ttemp = bc->ttemp; p.temis = bc->temis; p.fisot = bc->fisot;
  }
  p.status = status;
  p.wave_lower = h->d_wlo; p.wave_upper = h->d_whi; p.wave2band = h->d_w2b;
  p.s0 = 0;
  p.nlocal = (long)c.nwave * c.ncol;
  if (planck) {
    double* pkag = h->d_acc + n_acc;
    rc = launch_planck_into(h, temf, c.ncol, pkag, st);
    if (rc != B200DISORT_OK) return rc;
    p.pkag = pkag;
  }
  CU_TRY(cudaMemsetAsync(h->d_acc, 0, n_acc * sizeof(double), st));
  p.bflx_t = h->d_acc;
  p.weights = weights;
  p.warp_red = band_warp_red(h, c.ncol);
  p.nrep = p.warp_red ? kBandReplicas : 1;
  p.rep_stride = (long)rows * c.ncol;
  rc = launch_flux(h, p, st, &h->d_ws, &h->ws_bytes, true);
  if (rc != B200DISORT_OK) return rc;
  launch_band_finish(h->d_acc, rows, c.ncol, p.nrep, bflx, st);
  CU_TRY(cudaGetLastError());
  h->stats.kernel_launches += 1;
  return B200DISORT_OK;
}
