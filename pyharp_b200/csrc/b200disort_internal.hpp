// b200disort_internal.hpp -- pieces of the C-ABI implementation shared between b200disort.cu (solver entry
// points) and band.cu (band-level forward).  Not installed; the public surface is include/*.h.
#pragma once
#include <cuda_runtime.h>

#include <cstdio>
#include <vector>

#include "../../include/b200disort.h"
#include "disort_device.cuh"
#include "flux_inst.cuh"

struct b200disort_handle {
  b200disort_config cfg;
  const b200::FluxInst* inst = nullptr;
  cudaStream_t stream = nullptr;
  int sm_count = 0;
  int occ = 0;  // resident CTAs per SM of the flux kernel
  int occ_warp = 0;  // ... of its warp-per-solve form
  int layer_split = -1;  // nstr >= 12: 1 = layer-split form, 0 = fused kernel, -1 = default (b200disort_set_layer_split)
  long warp_solve_max = -1;  // largest batch sent to the warp-per-solve form (-1: one wave of warps)
  std::vector<double> user_tau, user_mu, user_phi, wave_lower, wave_upper;
  // device-side constants
  double *d_wlo = nullptr, *d_whi = nullptr, *d_blo = nullptr, *d_bhi = nullptr;
  int* d_w2b = nullptr;
  int nband = 0;
  double* d_pkag = nullptr;
  size_t pkag_elems = 0;
  double* d_ws = nullptr;
  size_t ws_bytes = 0;
  int* d_flag = nullptr;  // [0] planck bad-T flag, [1] bad status counter
  // general (intensity / user-level) path
  double *d_utau = nullptr, *d_umu = nullptr, *d_phi = nullptr, *d_quad = nullptr;  // d_quad = [cmu | cwt]
  double* d_uu = nullptr;   // intensities of the last solve (nwave,ncol,nphi,ntau,numu)
  size_t uu_bytes = 0;
  double* d_iws = nullptr;  // per-CTA workspaces of the intensity kernel
  size_t iws_bytes = 0;
  bool have_uu = false;
  double* d_part = nullptr;  // partial sums of reduce_band
  size_t part_bytes = 0;
  double* d_acc = nullptr;   // b200disort_solve_band: band-flux accumulation buffer [| Planck table]
  size_t acc_bytes = 0;
  // host-pointer entry point: the batch is cut into chunks of solves that flow through kSlots
  // independent (stream, device buffers, pinned staging) slots, so the H2D copy of chunk c+1, the
  // solve of chunk c and the D2H copy of chunk c-1 run concurrently (copy engines + SMs).
  static constexpr int kSlots = 3;
  struct Slot {
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
    double* d_buf = nullptr;   // [inputs | flx | status]
    size_t d_bytes = 0;
    double* h_buf = nullptr;   // pinned staging, same layout; used only for pageable user memory
    size_t h_bytes = 0;
    double* d_ws = nullptr;
    size_t ws_bytes = 0;
    // pending copy-out of a finished chunk (pageable destination)
    long pend_s0 = -1, pend_n = 0;
  } slots[kSlots];
  double* d_temf = nullptr;
  size_t temf_bytes = 0;
  cudaEvent_t ev_pre = nullptr;
  // ring of event pairs around the solve kernel: get_stats sums every solve since the last call
  static constexpr int kRing = 1024;
  cudaEvent_t ev0 = nullptr, evA[kRing] = {nullptr}, evB[kRing] = {nullptr};
  int ring_used = 0;
  double ring_carry_ms = 0.0;  // durations folded in when the ring wrapped
  int solves_since_stats = 0;
  b200disort_stats stats{};
};

#define CU_TRY(expr)                                                                         \
  do {                                                                                       \
    cudaError_t e__ = (expr);                                                                \
    if (e__ != cudaSuccess) {                                                                \
      std::fprintf(stderr, "b200disort: CUDA error '%s' at %s:%d\n", cudaGetErrorString(e__), \
                   __FILE__, __LINE__);                                                      \
      return B200DISORT_ERR_CUDA;                                                            \
    }                                                                                        \
  } while (0)


namespace b200 {
// grow-only device buffer
int ensure_dev(double** ptr, size_t* have, size_t need);
// Launch the flux kernel for p.nlocal solves on `stream` with the given back-sweep workspace; timed = record the
// launch in the handle's event ring (b200disort_get_stats).  p.bflx_t != NULL selects the fused band integration.
int launch_flux(b200disort_handle* h, SolveParams& p, cudaStream_t stream, double** ws, size_t* ws_bytes, bool timed);
// Planck pre-pass: temf (device, (ncol, nlyr+1), level 0 = surface) -> pkag (nband, ncol, nlyr+1), DISORT order
int launch_planck_into(b200disort_handle* h, const double* temf, int ncol, double* pkag, cudaStream_t stream);
// adds the number of non-zero status words to h->d_flag[1] (asynchronous)
int launch_count_status(b200disort_handle* h, const int32_t* status, long n, cudaStream_t stream);
// bflx(inner) = sum_n w[n] flx(n, inner) on `stream` (deterministic two-stage form)
int launch_reduce_band(b200disort_handle* h, int nwave, long inner, const double* flx, const double* weights, double* bflx,
                       cudaStream_t stream);
bool is_pinned(const void* p);
}  // namespace b200
