// opacity_device.cuh -- device forms of harp::locate / harp::interpn<N> shared by opacity.cu (the stand-alone
// interpolation / attenuator kernels) and band.cu (the fused band-optics kernel).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200 {

// harp::locate (src/math/locate.h:14-42): bisection on a monotonic table, exact-endpoint rule.
__device__ __forceinline__ int locate_dev(const double* xx0, double x, int n) {
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

struct Dims {
  long len[4];
  long stride[4];  // elements of `data` per step along each dimension (includes nval)
  long axis_off[4];
};

// harp::interpn<N> (src/math/interpn.h:34-76) with the recursion unrolled over NDIM; the blend uses
// explicit round-to-nearest mul/add so nvcc cannot contract it into FMAs (bit parity with the CPU path).
template <int NDIM, int NVAL>
__device__ void interpn_dev(double* val, const double* coor, const double* data, const double* axis, const Dims& d,
                            int dim0) {
  const long len = d.len[dim0];
  const double* ax = axis + d.axis_off[dim0];
  const double c = coor[0];
  int i1 = locate_dev(ax, c, (int)len), i2;
  if (i1 == -1) { i1 = 0; i2 = 0; }
  else if (i1 == len - 1) { i2 = (int)len - 1; }
  else i2 = i1 + 1;
  double v1[NVAL], v2[NVAL];
  const double x1 = ax[i1], x2 = ax[i2];
  if constexpr (NDIM == 1) {
#pragma unroll
    for (int j = 0; j < NVAL; ++j) { v1[j] = data[i1 * d.stride[dim0] + j]; v2[j] = data[i2 * d.stride[dim0] + j]; }
  } else {
    interpn_dev<NDIM - 1, NVAL>(v1, coor + 1, data + i1 * d.stride[dim0], axis, d, dim0 + 1);
    interpn_dev<NDIM - 1, NVAL>(v2, coor + 1, data + i2 * d.stride[dim0], axis, d, dim0 + 1);
  }
  if (x2 != x1) {
    const double a = __dsub_rn(c, x1), b = __dsub_rn(x2, c), den = __dsub_rn(x2, x1);
#pragma unroll
    for (int j = 0; j < NVAL; ++j) val[j] = __ddiv_rn(__dadd_rn(__dmul_rn(a, v2[j]), __dmul_rn(b, v1[j])), den);
  } else {
#pragma unroll
    for (int j = 0; j < NVAL; ++j) val[j] = __ddiv_rn(__dadd_rn(v1[j], v2[j]), 2.0);
  }
}


inline Dims make_dims(const int64_t* dims, int ndim, int nval) {
  Dims d{};
  long off = 0, stride = nval;
  for (int k = ndim - 1; k >= 0; --k) { d.stride[k] = stride; stride *= dims[k]; }
  for (int k = 0; k < ndim; ++k) { d.len[k] = dims[k]; d.axis_off[k] = off; off += dims[k]; }
  return d;
}

}  // namespace b200
