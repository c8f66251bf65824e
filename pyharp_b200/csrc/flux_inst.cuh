// flux_inst.cuh -- per-stream-count instantiations of the flux kernel live in their own
// translation units (flux_inst.cu compiled with -DB200_NN=<nn>) so they build in parallel.
// Each TU owns its constant-memory quadrature table.
#pragma once
#include <cuda_runtime.h>

#include "disort_device.cuh"

namespace b200 {

struct FluxInst {
  int nn;
  int nsave;  // doubles stored per level and solve for the back sweep
  int max_threads;  // __launch_bounds__ of the kernel
  int ws_halves;    // 2 when the back sweep is carried into the thread's next solve (alternating workspace halves)
  cudaError_t (*upload_tables)(const QuadTables*);
  cudaError_t (*occupancy)(int block, int* blocks_per_sm);
  cudaError_t (*launch)(const SolveParams&, unsigned grid, int block, cudaStream_t);
  // warp-per-solve form for batches far below one resident wave (null where it is not instantiated): doubles of
  // scratch per warp, resident CTAs per SM, launch
  long (*warp_scratch)(int nlyr);
  cudaError_t (*warp_occupancy)(int block, int* blocks_per_sm);
  cudaError_t (*launch_warp)(const SolveParams&, unsigned grid, int block, cudaStream_t, double* scratch);
  // layer-split form (null where it is not instantiated; nstr >= 12): the layer phase (eigenproblem ... R / T, sources)
  // as one thread per (solve, layer), the adding recursion + back sweep as one thread per solve, chunk by chunk through a
  // scratch buffer of split_scratch(nlyr, chunk) bytes
  size_t (*split_scratch)(int nlyr, long chunk);
  cudaError_t (*launch_split)(const SolveParams&, long chunk, int sm_count, cudaStream_t, void* scratch);
};

#define B200_DECL_INST(NN) const FluxInst* flux_inst_##NN();
B200_DECL_INST(1)
B200_DECL_INST(2)
B200_DECL_INST(3)
B200_DECL_INST(4)
B200_DECL_INST(6)
B200_DECL_INST(8)
B200_DECL_INST(12)
B200_DECL_INST(16)
#undef B200_DECL_INST

}  // namespace b200
