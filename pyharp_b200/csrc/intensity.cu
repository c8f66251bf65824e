// intensity.cu -- K3 launcher: one CTA per solve (grid-stride over the batch), 64 threads that
// split layers / band rows / (level, angle) outputs between them (see intensity_device.cuh).
#include <cuda_runtime.h>

#include "intensity_device.cuh"
#include "intensity_launch.hpp"

namespace b200 {

__global__ void __launch_bounds__(kIntensityThreads) disort_intensity_kernel(IntensityParams P) {
  double* w = P.ws + (long)blockIdx.x * P.ws_per_cta;
  for (long s = blockIdx.x; s < P.nlocal; s += gridDim.x) solve_intensity_one(P, s, w, threadIdx.x, blockDim.x);
}

long intensity_ws_doubles(int nlyr, int nstr, int ntau, int numu) {
  return intensity_ws_map(nlyr, nstr, ntau, numu, kIntensityThreads).total;
}

cudaError_t launch_intensity(const IntensityParams& P, unsigned grid, cudaStream_t stream) {
  disort_intensity_kernel<<<grid, kIntensityThreads, 0, stream>>>(P);
  return cudaGetLastError();
}

}  // namespace b200
