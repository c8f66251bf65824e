// FP64 vector-pipe microbenchmark for B200 (sm_100a).
// Measures: DFMA throughput vs ILP/occupancy (the roofline denominator for the
// DISORT solve path), and the cost of div / sqrt /
// exp / rcp in DFMA-equivalents.  Standalone: nvcc -O3 -gencode
// arch=compute_100a,code=sm_100a fp64_peak.cu -o fp64_peak
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

template <int ILP>
__global__ void dfma_kernel(double* out, int iters, double a, double b) {
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-9 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i];
  if (s == 12345.678) out[0] = s;
}

template <int OP>
__global__ void op_kernel(double* out, int iters, double a, double b) {
  double x[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) x[i] = 1.0 + threadIdx.x * 1e-6 + i * 0.1;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (OP == 0) x[i] = a / x[i] + b;              // div
      else if (OP == 1) x[i] = sqrt(x[i]) + b;       // sqrt
      else if (OP == 2) x[i] = exp(-x[i]) + b;       // exp
      else if (OP == 3) x[i] = rsqrt(x[i]) + b;      // rsqrt
      else if (OP == 4) x[i] = __drcp_rn(x[i]) + b;  // rcp
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) s += x[i];
  if (s == 12345.678) out[0] = s;
}

template <typename F>
float time_it(F launch) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  launch(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int nsm = p.multiProcessorCount;
  printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d,\n", p.name, nsm, p.clockRate);
  double* out; CK(cudaMalloc(&out, 8));
  const int iters = 4096;
  printf(" \"dfma\": [\n");
  int tpbs[] = {64, 128, 256, 512, 1024};
  bool first = true;
  double best_tf = 0;
  for (int tpb : tpbs) {
    for (int ilp : {1, 2, 4, 8}) {
      int blocks = nsm * (1024 / tpb >= 1 ? 1 : 1);
      // one block per SM of tpb threads, and a second config filling 2048 threads/SM
      for (int fill = 0; fill < 2; ++fill) {
        int nb = fill ? nsm * (2048 / tpb) : nsm;
        float ms;
        if (ilp == 1) ms = time_it([&] { dfma_kernel<1><<<nb, tpb>>>(out, iters, 1.0000001, 1e-9); });
        else if (ilp == 2) ms = time_it([&] { dfma_kernel<2><<<nb, tpb>>>(out, iters, 1.0000001, 1e-9); });
        else if (ilp == 4) ms = time_it([&] { dfma_kernel<4><<<nb, tpb>>>(out, iters, 1.0000001, 1e-9); });
        else ms = time_it([&] { dfma_kernel<8><<<nb, tpb>>>(out, iters, 1.0000001, 1e-9); });
        double flops = 2.0 * (double)nb * tpb * ilp * 8.0 * iters;
        double tf = flops / (ms * 1e-3) / 1e12;
        if (tf > best_tf) best_tf = tf;
        printf("%s  {\"tpb\": %d, \"warps_per_sm\": %d, \"ilp\": %d, \"ms\": %.4f, \"tflops\": %.3f}",
               first ? "" : ",\n", tpb, nb / nsm * tpb / 32, ilp, ms, tf);
        first = false;
        (void)blocks;
      }
    }
  }
  printf("\n ],\n \"dfma_peak_tflops\": %.3f,\n \"ops\": {", best_tf);
  const char* names[] = {"div", "sqrt", "exp", "rsqrt", "rcp"};
  for (int op = 0; op < 5; ++op) {
    int nb = nsm * 8, tpb = 256;
    float ms;
    if (op == 0) ms = time_it([&] { op_kernel<0><<<nb, tpb>>>(out, 1024, 1.5, 0.25); });
    else if (op == 1) ms = time_it([&] { op_kernel<1><<<nb, tpb>>>(out, 1024, 1.5, 0.25); });
    else if (op == 2) ms = time_it([&] { op_kernel<2><<<nb, tpb>>>(out, 1024, 1.5, 0.25); });
    else if (op == 3) ms = time_it([&] { op_kernel<3><<<nb, tpb>>>(out, 1024, 1.5, 0.25); });
    else ms = time_it([&] { op_kernel<4><<<nb, tpb>>>(out, 1024, 1.5, 0.25); });
    double nops = (double)nb * tpb * 4.0 * 1024;
    double gops = nops / (ms * 1e-3) / 1e9;
    // DFMA-equivalents: (peak DFMA/s) / (ops/s)
    double equiv = (best_tf * 1e12 / 2.0) / (gops * 1e9);
    printf("%s\"%s\": {\"gops\": %.2f, \"dfma_equiv\": %.1f}", op ? ", " : "", names[op], gops, equiv);
  }
  printf("}\n}\n");
  return 0;
}
