// intensity_launch.hpp -- host-visible interface of intensity.cu
#pragma once
#include <cuda_runtime.h>

namespace b200 {
struct IntensityParams;
constexpr int kIntensityThreads = 64;
long intensity_ws_doubles(int nlyr, int nstr, int ntau, int numu);
cudaError_t launch_intensity(const IntensityParams& P, unsigned grid, cudaStream_t stream);
}  // namespace b200
