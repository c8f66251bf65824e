// b200disort.cu -- kernels and C ABI of the B200-native DISORT solve path
// (include/b200disort.h).  sm_100a only; no CPU fallback: every entry point that
// computes needs a CUDA device and fails with B200DISORT_ERR_CUDA otherwise.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <new>
#include <utility>
#include <vector>

#include "b200disort_internal.hpp"
#include "host_tables.hpp"
#include "intensity_device.cuh"
#include "intensity_launch.hpp"

namespace b200 {

static const FluxInst* inst_of(int nstr) {
  switch (nstr) {
    case 2: return flux_inst_1();
    case 4: return flux_inst_2();
    case 6: return flux_inst_3();
    case 8: return flux_inst_4();
    case 12: return flux_inst_6();
    case 16: return flux_inst_8();
    case 24: return flux_inst_12();
    case 32: return flux_inst_16();
    default: return nullptr;
  }
}

// ---------------------------------------------------------------------------
// K2: Planck pre-pass.  pkag[band][col][lev] = PLKAVG(wlo_band, whi_band, T(col, lev)),
// DISORT (top-down) level order; temf is pyharp bottom-up.  Hoisted out of the solve so
// g-points sharing a band (amars_lw: all of them) do not recompute it.
// ---------------------------------------------------------------------------
__global__ void planck_prepass_kernel(int nband, int ncol, int nlyr, const double* __restrict__ blo,
                                      const double* __restrict__ bhi, const double* __restrict__ temf,
                                      double* __restrict__ pkag, int* __restrict__ bad) {
  const long total = (long)nband * ncol * (nlyr + 1);
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int lev = (int)(i % (nlyr + 1));
    const long bj = i / (nlyr + 1);
    const int j = (int)(bj % ncol);
    const int b = (int)(bj / ncol);
    const double t = temf[(long)j * (nlyr + 1) + (nlyr - lev)];
    if (!(t >= 0.0)) {
      *bad = 1;
      pkag[i] = nan("");
    } else {
      pkag[i] = plkavg(blo[b], bhi[b], t);
    }
  }
}

// ---------------------------------------------------------------------------
// K5: weighted band reduction bflx(ncol,nlev,2) = sum_n w_n flx(n,ncol,nlev,2)
// ---------------------------------------------------------------------------
// Deterministic two-stage form: blockIdx.y owns a contiguous slice of spectral points and writes
// a partial sum (or the result when gridDim.y == 1); reduce_parts_kernel adds the partials in
// slice order.  Threads run along `inner` (coalesced); the slices give enough CTAs when nwave is
// large and ncol small (line-by-line: 1e5 spectral points x 1 column).
__global__ void reduce_band_kernel(int nwave, long inner, const double* __restrict__ flx,
                                   const double* __restrict__ w, double* __restrict__ out) {
  const int per = (nwave + gridDim.y - 1) / gridDim.y;
  const int n0 = blockIdx.y * per, n1 = min(nwave, n0 + per);
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < inner; i += (long)gridDim.x * blockDim.x) {
    double acc = 0.0;
    for (int n = n0; n < n1; ++n) acc = fma(w[n], flx[(long)n * inner + i], acc);
    out[(long)blockIdx.y * inner + i] = acc;
  }
}

__global__ void reduce_parts_kernel(int parts, long inner, const double* __restrict__ part,
                                    double* __restrict__ out) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < inner; i += (long)gridDim.x * blockDim.x) {
    double acc = 0.0;
    for (int k = 0; k < parts; ++k) acc += part[(long)k * inner + i];
    out[i] = acc;
  }
}

__global__ void count_status_kernel(const int32_t* __restrict__ st, long n, int* __restrict__ nbad) {
  int local = 0;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    if (st[i] != 0) ++local;
  if (local) atomicAdd(nbad, local);
}

// DFMA-saturating kernel: 8 independent chains per thread
__global__ void dfma_peak_kernel(double* out, int iters, double a, double b) {
  double x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-9 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = fma(x[i], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i];
  if (s == 12345.678) out[0] = s;
}

}  // namespace b200

using namespace b200;

static int nlev_out(const b200disort_config& c) {
  return (c.flags & B200DISORT_USRTAU) ? c.ntau : c.nlyr + 1;
}

extern "C" const char* b200disort_version(void) { return "b200disort 0.1 (sm_100a, fp64)"; }

extern "C" const char* b200disort_strerror(int code) {
  switch (code) {
    case B200DISORT_OK: return "ok";
    case B200DISORT_ERR_INVALID: return "invalid argument or configuration";
    case B200DISORT_ERR_CUDA: return "CUDA failure (no device, out of memory, or launch error); there is no CPU fallback";
    case B200DISORT_ERR_UNSUPPORTED: return "configuration not supported by the B200 solve path";
    case B200DISORT_ERR_SOLVE: return "one or more solves reported a non-zero status";
    default: return "unknown error";
  }
}

extern "C" void b200disort_destroy(b200disort_handle* h);
// everything of b200disort_create that can fail after the handle exists; the caller destroys the handle on error
static int init_handle(b200disort_handle* h, const b200disort_config* cfg) {
  h->stream = (cudaStream_t)cfg->stream;
  if (cfg->user_tau && cfg->ntau > 0) h->user_tau.assign(cfg->user_tau, cfg->user_tau + cfg->ntau);
  if (cfg->user_mu && cfg->numu > 0) h->user_mu.assign(cfg->user_mu, cfg->user_mu + cfg->numu);
  if (cfg->user_phi && cfg->nphi > 0) h->user_phi.assign(cfg->user_phi, cfg->user_phi + cfg->nphi);
  if (cfg->wave_lower) h->wave_lower.assign(cfg->wave_lower, cfg->wave_lower + cfg->nwave);
  if (cfg->wave_upper) h->wave_upper.assign(cfg->wave_upper, cfg->wave_upper + cfg->nwave);
  h->cfg.user_tau = h->user_tau.data();
  h->cfg.user_mu = h->user_mu.data();
  h->cfg.user_phi = h->user_phi.data();
  h->cfg.wave_lower = h->wave_lower.empty() ? nullptr : h->wave_lower.data();
  h->cfg.wave_upper = h->wave_upper.empty() ? nullptr : h->wave_upper.data();

  cudaDeviceProp prop;
  CU_TRY(cudaGetDeviceProperties(&prop, cfg->device));
  h->sm_count = prop.multiProcessorCount;

  // constant tables for this stream count
  QuadTables qt;
  build_quad_tables(cfg->nstr, &qt);
  if (h->inst) CU_TRY(h->inst->upload_tables(&qt));

  // spectral bands: unique (wvnmlo, wvnmhi) pairs
  if (!h->wave_lower.empty()) {
    std::map<std::pair<double, double>, int> bands;
    std::vector<int> w2b(cfg->nwave);
    std::vector<double> blo, bhi;
    for (int n = 0; n < cfg->nwave; ++n) {
      auto key = std::make_pair(h->wave_lower[n], h->wave_upper[n]);
      auto it = bands.find(key);
      if (it == bands.end()) {
        it = bands.emplace(key, (int)blo.size()).first;
        blo.push_back(key.first);
        bhi.push_back(key.second);
      }
      w2b[n] = it->second;
      if (!(key.first >= 0.0) || !(key.second >= key.first)) return B200DISORT_ERR_INVALID;
    }
    h->nband = (int)blo.size();
    CU_TRY(cudaMalloc(&h->d_wlo, sizeof(double) * cfg->nwave));
    CU_TRY(cudaMalloc(&h->d_whi, sizeof(double) * cfg->nwave));
    CU_TRY(cudaMalloc(&h->d_blo, sizeof(double) * h->nband));
    CU_TRY(cudaMalloc(&h->d_bhi, sizeof(double) * h->nband));
    CU_TRY(cudaMalloc(&h->d_w2b, sizeof(int) * cfg->nwave));
    CU_TRY(cudaMemcpy(h->d_wlo, h->wave_lower.data(), sizeof(double) * cfg->nwave, cudaMemcpyHostToDevice));
    CU_TRY(cudaMemcpy(h->d_whi, h->wave_upper.data(), sizeof(double) * cfg->nwave, cudaMemcpyHostToDevice));
    CU_TRY(cudaMemcpy(h->d_blo, blo.data(), sizeof(double) * h->nband, cudaMemcpyHostToDevice));
    CU_TRY(cudaMemcpy(h->d_bhi, bhi.data(), sizeof(double) * h->nband, cudaMemcpyHostToDevice));
    CU_TRY(cudaMemcpy(h->d_w2b, w2b.data(), sizeof(int) * cfg->nwave, cudaMemcpyHostToDevice));
  }
  {
    auto up = [&](const std::vector<double>& v, double** d) -> int {
      if (v.empty()) return B200DISORT_OK;
      CU_TRY(cudaMalloc(d, sizeof(double) * v.size()));
      CU_TRY(cudaMemcpy(*d, v.data(), sizeof(double) * v.size(), cudaMemcpyHostToDevice));
      return B200DISORT_OK;
    };
    std::vector<double> quad(cfg->nstr);
    for (int i = 0; i < cfg->nstr / 2; ++i) { quad[i] = qt.mu[i]; quad[cfg->nstr / 2 + i] = qt.wt[i]; }
    int rc = up(h->user_tau, &h->d_utau);
    if (rc == B200DISORT_OK) rc = up(h->user_mu, &h->d_umu);
    if (rc == B200DISORT_OK) rc = up(h->user_phi, &h->d_phi);
    if (rc == B200DISORT_OK) rc = up(quad, &h->d_quad);
    if (rc != B200DISORT_OK) return rc;
  }
  CU_TRY(cudaMalloc(&h->d_flag, sizeof(int) * 2));
  CU_TRY(cudaMemset(h->d_flag, 0, sizeof(int) * 2));
  CU_TRY(cudaEventCreate(&h->ev0));
  for (int i = 0; i < b200disort_handle::kRing; ++i) {
    CU_TRY(cudaEventCreate(&h->evA[i]));
    CU_TRY(cudaEventCreate(&h->evB[i]));
  }
  return B200DISORT_OK;
}


extern "C" int b200disort_create(const b200disort_config* cfg, b200disort_handle** out) {
  if (!cfg || !out) return B200DISORT_ERR_INVALID;
  *out = nullptr;
  if (cfg->nwave < 1 || cfg->ncol < 1 || cfg->nlyr < 1) return B200DISORT_ERR_INVALID;
  if (cfg->nstr < 2 || (cfg->nstr & 1) || cfg->nmom < cfg->nstr) return B200DISORT_ERR_INVALID;
  if (!(cfg->flags & B200DISORT_LAMBER)) return B200DISORT_ERR_UNSUPPORTED;  // BRDF not on pyharp's path
  // K1 (the flux kernel) is instantiated for nstr in {2,4,6,8,12,16,24,32}; any other even nstr <= 32 is served by
  // the general kernel K3, whose sizes are run-time (one CTA per solve: correct, not tuned for throughput)
  const FluxInst* inst = inst_of(cfg->nstr);
  if (std::getenv("B200DISORT_FORCE_GENERAL")) inst = nullptr;  // profiling knob: every solve on the general kernel
  if (!inst && cfg->nstr > 2 * kMaxNN) return B200DISORT_ERR_UNSUPPORTED;
  if ((cfg->flags & B200DISORT_USRTAU) && (cfg->ntau < 1 || !cfg->user_tau)) return B200DISORT_ERR_INVALID;
  const bool onlyfl = cfg->flags & B200DISORT_ONLYFL;
  if (!onlyfl && (cfg->flags & B200DISORT_USRANG) && (cfg->numu < 1 || !cfg->user_mu)) return B200DISORT_ERR_INVALID;
  if (!onlyfl && (cfg->nphi < 1 || !cfg->user_phi)) return B200DISORT_ERR_INVALID;
  if ((cfg->flags & B200DISORT_PLANCK) && (!cfg->wave_lower || !cfg->wave_upper)) return B200DISORT_ERR_INVALID;

  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    std::fprintf(stderr, "b200disort: no CUDA device; this library has no CPU fallback\n");
    return B200DISORT_ERR_CUDA;
  }
  if (cfg->device < 0 || cfg->device >= ndev) return B200DISORT_ERR_INVALID;
  CU_TRY(cudaSetDevice(cfg->device));

  b200disort_handle* h = new (std::nothrow) b200disort_handle();
  if (!h) return B200DISORT_ERR_INVALID;
  h->cfg = *cfg;
  h->inst = inst;
  const int rc = init_handle(h, cfg);
  if (rc != B200DISORT_OK) {  // release whatever was allocated before the failure
    b200disort_destroy(h);
    return rc;
  }
  *out = h;
  return B200DISORT_OK;
}

extern "C" void b200disort_destroy(b200disort_handle* h) {
  if (!h) return;
  cudaSetDevice(h->cfg.device);
  cudaStreamSynchronize(h->stream);
  cudaFree(h->d_wlo); cudaFree(h->d_whi); cudaFree(h->d_blo); cudaFree(h->d_bhi);
  cudaFree(h->d_w2b); cudaFree(h->d_pkag); cudaFree(h->d_ws); cudaFree(h->d_flag);
  cudaFree(h->d_part); cudaFree(h->d_temf); cudaFree(h->d_acc);
  cudaFree(h->d_utau); cudaFree(h->d_umu); cudaFree(h->d_phi); cudaFree(h->d_quad); cudaFree(h->d_uu); cudaFree(h->d_iws);
  for (auto& sl : h->slots) {
    if (sl.stream) { cudaStreamSynchronize(sl.stream); cudaStreamDestroy(sl.stream); }
    if (sl.done) cudaEventDestroy(sl.done);
    cudaFree(sl.d_buf); cudaFree(sl.d_ws);
    if (sl.h_buf) cudaFreeHost(sl.h_buf);
  }
  if (h->ev_pre) cudaEventDestroy(h->ev_pre);
  if (h->ev0) cudaEventDestroy(h->ev0);
  for (int i = 0; i < b200disort_handle::kRing; ++i) {
    if (h->evA[i]) cudaEventDestroy(h->evA[i]);
    if (h->evB[i]) cudaEventDestroy(h->evB[i]);
  }
  delete h;
}

constexpr long kHostChunkDiv = 1;  // host-path chunk = resident wave / kHostChunkDiv

// grow-only device buffer
int b200::ensure_dev(double** ptr, size_t* have, size_t need) {
  if (need <= *have) return B200DISORT_OK;
  if (*ptr) CU_TRY(cudaFree(*ptr));
  *ptr = nullptr;
  *have = 0;
  CU_TRY(cudaMalloc(ptr, need));
  *have = need;
  return B200DISORT_OK;
}

// Event ring behind b200disort_get_stats: one (begin, end) pair per timed solve-kernel launch on the handle's stream.
static int ring_begin(b200disort_handle* h, cudaStream_t stream) {
  if (h->ring_used == b200disort_handle::kRing) {  // fold the ring (blocks until those solves finished)
    CU_TRY(cudaEventSynchronize(h->evB[h->ring_used - 1]));
    for (int i = 0; i < h->ring_used; ++i) {
      float ms = 0.f;
      CU_TRY(cudaEventElapsedTime(&ms, h->evA[i], h->evB[i]));
      h->ring_carry_ms += ms;
    }
    h->ring_used = 0;
  }
  CU_TRY(cudaEventRecord(h->evA[h->ring_used], stream));
  return B200DISORT_OK;
}

static int ring_end(b200disort_handle* h, cudaStream_t stream) {
  CU_TRY(cudaEventRecord(h->evB[h->ring_used], stream));
  h->ring_used += 1;
  h->solves_since_stats += 1;
  return B200DISORT_OK;
}

// Layer-split form by default (measured, profiles/r02/sweeps.md e36, e45): from nstr = 16 up (nstr = 12 loses 14 %, nstr = 16 /
// 24 / 32 gain 1.3-1.9x / 1.3x / 2x), and at nstr = 8 for batches above the warp-per-solve form's range of at most this many
// solves (1184 < n <= 8192: 0.14-0.21 ms per call against 0.23 ms for the fused kernel)
// (nstr = 4: the split form is instantiated and tested, but never faster than the fused kernel -- a call of 10^3..10^5 solves is
// bound by its ~0.1 ms of launch / host overhead, sweeps e46 -- so it is off by default there)
static long small_max(int nn) { return nn == 4 ? 8192 : 0; }

// Launch the flux kernel for p.nlocal solves on `stream` with the given back-sweep workspace.
// timed = record the launch in the handle's event ring (handle stream only).
int b200::launch_flux(b200disort_handle* h, SolveParams& p, cudaStream_t stream, double** ws, size_t* ws_bytes,
                      bool timed) {
  const FluxInst* inst = h->inst;
  // Layer-split form (flux_inst.cu): layer phase as one thread per (solve, layer), adding recursion + back sweep as a second
  // kernel (teams of four lanes per solve at nstr = 8 / 16), chunk by chunk through a scratch buffer.  It is the default
  // from nstr = 16 up (the fused kernel spills there) and, at nstr = 8, for batches far below one resident wave of the fused
  // kernel -- pyharp's own calls are 500 x 1 and G x 1 solves, and the layer phase has nlyr times the parallelism of the
  // batch.  B200DISORT_SPLIT / the handle's layer_split: 0 = never, 1 = always, -1 = that default;
  // B200DISORT_SPLIT_SMALL_MAX = largest nstr = 8 batch that takes it; B200DISORT_SPLIT_CHUNK = solves per chunk
  // (default: two resident waves of the adding kernel).
  if (inst->launch_split) {
    static const int env_split = std::getenv("B200DISORT_SPLIT") ? std::atoi(std::getenv("B200DISORT_SPLIT")) : -1;
    static const long env_chunk = std::getenv("B200DISORT_SPLIT_CHUNK") ? std::atol(std::getenv("B200DISORT_SPLIT_CHUNK")) : 0;
    static const long env_small = std::getenv("B200DISORT_SPLIT_SMALL_MAX") ? std::atol(std::getenv("B200DISORT_SPLIT_SMALL_MAX")) : -1;
    const int mode = h->layer_split >= 0 ? h->layer_split : env_split;
    // (nstr = 8: up to one resident wave of warps the warp-per-solve form below is faster still -- 0.12 against 0.13 ms per
    // call -- and from ~10^4 solves on the fused kernel is; in between the split form is 1.1-1.5x faster than either)
    static const long env_warp = std::getenv("B200DISORT_WARP_SOLVE_MAX") ? std::atol(std::getenv("B200DISORT_WARP_SOLVE_MAX")) : -1;
    const long warp_max = h->warp_solve_max >= 0 ? h->warp_solve_max : env_warp >= 0 ? env_warp : (long)h->sm_count * 8;
    const bool use = mode == 1 || (mode < 0 && (inst->nn >= 8 || (p.nlocal > warp_max &&
                                                                  p.nlocal <= (env_small >= 0 ? env_small : small_max(inst->nn)))));
    if (use) {
      long chunk = env_chunk > 0 ? env_chunk : (long)h->sm_count * 256 * 2;
      chunk = std::min<long>(chunk, p.nlocal);
      // (the scratch is nlyr x ~(nn^2 + nn) x 16 bytes per solve of the chunk: 82 KB at nstr = 16, 60 layers.  Many-layer
      // columns halve the chunk down to half a resident wave rather than take more than 8 GB)
      if (env_chunk <= 0)
        while (inst->split_scratch(p.nlyr, chunk) > ((size_t)8 << 30) && chunk > (long)h->sm_count * 128) chunk = (chunk + 1) / 2;
      const size_t need = inst->split_scratch(p.nlyr, chunk);
      int rc = ensure_dev(ws, ws_bytes, need);
      if (rc != B200DISORT_OK) return rc;
      if (timed) {
        rc = ring_begin(h, stream);
        if (rc != B200DISORT_OK) return rc;
      }
      CU_TRY(inst->launch_split(p, chunk, h->sm_count, stream, *ws));
      if (timed) {
        rc = ring_end(h, stream);
        if (rc != B200DISORT_OK) return rc;
      }
      h->stats.kernel_launches += 3 * (int)((p.nlocal + chunk - 1) / chunk);
      h->stats.workspace_bytes = *ws_bytes;
      return B200DISORT_OK;
    }
  }
  // Batches far below one resident wave (pyharp's own calls: 500 x 1, G x 1 solves): the warp-per-solve form, whose
  // lanes take the layers of a solve -- a thread-per-solve launch would occupy a handful of SMs for as long as one
  // thread needs to walk the column.  One call of it lasts ~0.55x as long (profiles/r02/sweeps.md, e32), so it is used
  // up to ONE resident wave of warps (148 SMs x 8 warps = 1184 solves); B200DISORT_WARP_SOLVE_MAX / the handle's
  // warp_solve_max override the threshold (0 = never).
  if (inst->launch_warp) {
    static const long env_max = std::getenv("B200DISORT_WARP_SOLVE_MAX") ? std::atol(std::getenv("B200DISORT_WARP_SOLVE_MAX")) : -1;
    if (h->occ_warp == 0) {
      int occ = 0;
      CU_TRY(inst->warp_occupancy(128, &occ));
      h->occ_warp = occ < 1 ? 1 : occ;
    }
    const long wave_warps = (long)h->sm_count * h->occ_warp * 4;
    const long max_solves = h->warp_solve_max >= 0 ? h->warp_solve_max : env_max >= 0 ? env_max : wave_warps;
    if (p.nlocal <= max_solves) {
      const long grid = std::min<long>((p.nlocal + 3) / 4, (long)h->sm_count * h->occ_warp);
      const size_t need = (size_t)grid * 4 * inst->warp_scratch(p.nlyr) * sizeof(double);
      int rc = ensure_dev(ws, ws_bytes, need);
      if (rc != B200DISORT_OK) return rc;
      if (timed) {
        rc = ring_begin(h, stream);
        if (rc != B200DISORT_OK) return rc;
      }
      CU_TRY(inst->launch_warp(p, (unsigned)grid, 128, stream, *ws));
      if (timed) {
        rc = ring_end(h, stream);
        if (rc != B200DISORT_OK) return rc;
      }
      h->stats.kernel_launches += 1;
      h->stats.workspace_bytes = *ws_bytes;
      return B200DISORT_OK;
    }
  }
  // tuning knobs (profiling only): B200DISORT_BLOCK = threads per CTA (<= 128), B200DISORT_CTAS_PER_SM caps residency
  static const int env_block = std::getenv("B200DISORT_BLOCK") ? std::atoi(std::getenv("B200DISORT_BLOCK")) : 0;
  static const int env_ctas = std::getenv("B200DISORT_CTAS_PER_SM") ? std::atoi(std::getenv("B200DISORT_CTAS_PER_SM")) : 0;
  const int block = (env_block >= 32 && env_block <= inst->max_threads && env_block % 32 == 0) ? env_block : 128;  // whole warps: the workspace is laid out per warp
  if (h->occ == 0) {
    int occ = 0;
    CU_TRY(inst->occupancy(block, &occ));
    h->occ = occ < 1 ? 1 : occ;
    if (env_ctas > 0) h->occ = std::min(h->occ, env_ctas);
  }
  long grid = std::min<long>((p.nlocal + block - 1) / block, (long)h->sm_count * h->occ);
  const long nthreads = grid * block;
  const size_t half = (size_t)p.nlyr * inst->nsave * nthreads;
  const size_t need = half * inst->ws_halves * sizeof(double);
  int rc = ensure_dev(ws, ws_bytes, need);
  if (rc != B200DISORT_OK) return rc;
  p.ws = *ws;
  p.ws_threads = nthreads;
  p.ws_half = (long)half;
  if (timed) {
    rc = ring_begin(h, stream);
    if (rc != B200DISORT_OK) return rc;
  }
  CU_TRY(inst->launch(p, (unsigned)grid, block, stream));
  if (timed) {
    rc = ring_end(h, stream);
    if (rc != B200DISORT_OK) return rc;
  }
  h->stats.kernel_launches += 1;
  h->stats.workspace_bytes = *ws_bytes;
  return B200DISORT_OK;
}

int b200::launch_planck_into(b200disort_handle* h, const double* temf, int ncol, double* pkag, cudaStream_t stream) {
  const size_t elems = (size_t)h->nband * ncol * (h->cfg.nlyr + 1);
  const int block = 256;
  const long grid = std::min<long>((long)((elems + block - 1) / block), (long)h->sm_count * 8);
  planck_prepass_kernel<<<(unsigned)grid, block, 0, stream>>>(h->nband, ncol, h->cfg.nlyr, h->d_blo, h->d_bhi, temf, pkag,
                                                             h->d_flag);
  CU_TRY(cudaGetLastError());
  h->stats.kernel_launches += 1;
  return B200DISORT_OK;
}

// Planck pre-pass on `stream`: temf (device) -> h->d_pkag
static int launch_planck(b200disort_handle* h, const double* temf, cudaStream_t stream) {
  const b200disort_config& c = h->cfg;
  const size_t elems = (size_t)h->nband * c.ncol * (c.nlyr + 1);
  size_t have = h->pkag_elems * sizeof(double);
  int rc = ensure_dev(&h->d_pkag, &have, elems * sizeof(double));
  h->pkag_elems = have / sizeof(double);
  if (rc != B200DISORT_OK) return rc;
  return launch_planck_into(h, temf, c.ncol, h->d_pkag, stream);
}

int b200::launch_count_status(b200disort_handle* h, const int32_t* status, long n, cudaStream_t stream) {
  count_status_kernel<<<(unsigned)std::min<long>((n + 255) / 256, 1024), 256, 0, stream>>>(status, n, h->d_flag + 1);
  CU_TRY(cudaGetLastError());
  return B200DISORT_OK;
}

static int ntau_eff(const b200disort_config& c) { return (c.flags & B200DISORT_USRTAU) ? c.ntau : c.nlyr + 1; }
static int numu_eff(const b200disort_config& c) {
  return ((c.flags & B200DISORT_USRANG) && !(c.flags & B200DISORT_ONLYFL)) ? c.numu : c.nstr;
}
static size_t uu_elems(const b200disort_config& c) {
  if (c.flags & B200DISORT_ONLYFL) return 0;
  return (size_t)c.nwave * c.ncol * c.nphi * ntau_eff(c) * numu_eff(c);
}

// General path (user levels / angles, azimuth series, intensity corrections): K3, one CTA per solve.
static int solve_general(b200disort_handle* h, const double* prop, int nprop, const b200disort_bc* bc,
                         const double* temf, double* flx, double* uu, int32_t* status) {
  const b200disort_config& c = h->cfg;
  const bool onlyfl = c.flags & B200DISORT_ONLYFL;
  IntensityParams P{};
  P.nwave = c.nwave; P.ncol = c.ncol; P.nlyr = c.nlyr; P.nprop = nprop; P.nstr = c.nstr; P.nmom = c.nmom;
  P.ntau = ntau_eff(c); P.numu = numu_eff(c); P.nphi = onlyfl ? 0 : c.nphi;
  P.flags = c.flags; P.accur = c.accur;
  P.prop = prop;
  if (bc) {
    P.umu0 = bc->umu0; P.fbeam = bc->fbeam; P.albedo = bc->albedo; P.btemp = bc->btemp;
    P.ttemp = bc->ttemp; P.temis = bc->temis; P.fisot = bc->fisot; P.phi0 = bc->phi0;
  }
  P.temf = temf;
  P.wave_lower = h->d_wlo; P.wave_upper = h->d_whi;
  P.utau = h->d_utau; P.umu = h->d_umu; P.phi = h->d_phi;
  P.cmu = h->d_quad; P.cwt = h->d_quad + c.nstr / 2;
  P.flx = flx; P.status = status;
  const size_t nuu = uu_elems(c);
  if (nuu) {
    int rc = ensure_dev(&h->d_uu, &h->uu_bytes, nuu * sizeof(double));
    if (rc != B200DISORT_OK) return rc;
    P.uu = h->d_uu;
  }
  const long nsolve = (long)c.nwave * c.ncol;
  const long grid = std::min<long>(nsolve, (long)h->sm_count * 4);
  P.ws_per_cta = intensity_ws_doubles(c.nlyr, c.nstr, P.ntau, P.numu);
  int rc = ensure_dev(&h->d_iws, &h->iws_bytes, (size_t)grid * P.ws_per_cta * sizeof(double));
  if (rc != B200DISORT_OK) return rc;
  P.ws = h->d_iws;
  P.s0 = 0; P.nlocal = nsolve;
  rc = ring_begin(h, h->stream);  // the general kernel is timed like the flux kernel (b200disort_get_stats)
  if (rc != B200DISORT_OK) return rc;
  CU_TRY(launch_intensity(P, (unsigned)grid, h->stream));
  rc = ring_end(h, h->stream);
  if (rc != B200DISORT_OK) return rc;
  h->stats.kernel_launches += 1;
  h->stats.workspace_bytes = h->iws_bytes;
  h->have_uu = nuu != 0;
  if (nuu && uu) CU_TRY(cudaMemcpyAsync(uu, h->d_uu, nuu * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  return B200DISORT_OK;
}

extern "C" int b200disort_solve(b200disort_handle* h, const double* prop, int nprop,
                                const b200disort_bc* bc, const double* temf, double* flx,
                                double* uu, int32_t* status) {
  if (!h || !prop || !flx || !status || nprop < 1) return B200DISORT_ERR_INVALID;
  const b200disort_config& c = h->cfg;
  const bool planck = c.flags & B200DISORT_PLANCK;
  const bool onlyfl = c.flags & B200DISORT_ONLYFL;
  if (planck && !temf) return B200DISORT_ERR_INVALID;
  CU_TRY(cudaSetDevice(c.device));
  h->stats.kernel_launches = 0;
  CU_TRY(cudaEventRecord(h->ev0, h->stream));
  if (!onlyfl || (c.flags & B200DISORT_USRTAU) || !h->inst) return solve_general(h, prop, nprop, bc, temf, flx, uu, status);

  SolveParams p{};
  p.nwave = c.nwave; p.ncol = c.ncol; p.nlyr = c.nlyr; p.nprop = nprop; p.nstr = c.nstr;
  p.flags = c.flags;
  p.prop = prop;
  if (bc) {
    p.umu0 = bc->umu0; p.fbeam = bc->fbeam; p.albedo = bc->albedo; p.btemp = bc->btemp;
    p.ttemp = bc->ttemp; p.temis = bc->temis; p.fisot = bc->fisot;
  }
  p.flx = flx;
  p.status = status;
  p.wave_lower = h->d_wlo;
  p.wave_upper = h->d_whi;
  p.wave2band = h->d_w2b;

  p.s0 = 0;
  p.nlocal = (long)c.nwave * c.ncol;
  if (planck) {
    int rc = launch_planck(h, temf, h->stream);
    if (rc != B200DISORT_OK) return rc;
    p.pkag = h->d_pkag;
  }
  return launch_flux(h, p, h->stream, &h->d_ws, &h->ws_bytes, true);
}

extern "C" int b200disort_get_stats(b200disort_handle* h, b200disort_stats* out) {
  if (!h || !out) return B200DISORT_ERR_INVALID;
  CU_TRY(cudaSetDevice(h->cfg.device));
  double sum = h->ring_carry_ms;
  float tot = 0.f;
  if (h->ring_used > 0) {
    CU_TRY(cudaEventSynchronize(h->evB[h->ring_used - 1]));
    for (int i = 0; i < h->ring_used; ++i) {
      float ms = 0.f;
      CU_TRY(cudaEventElapsedTime(&ms, h->evA[i], h->evB[i]));
      sum += ms;
    }
    CU_TRY(cudaEventElapsedTime(&tot, h->ev0, h->evB[h->ring_used - 1]));
  }
  h->stats.solve_kernel_ms = (float)sum;
  h->stats.solves_timed = h->solves_since_stats;
  h->stats.total_device_ms = tot;
  *out = h->stats;
  h->ring_used = 0;
  h->ring_carry_ms = 0.0;
  h->solves_since_stats = 0;
  return B200DISORT_OK;
}

extern "C" int b200disort_set_stream(b200disort_handle* h, void* stream) {
  if (!h) return B200DISORT_ERR_INVALID;
  if (h->stream == (cudaStream_t)stream) return B200DISORT_OK;
  CU_TRY(cudaSetDevice(h->cfg.device));
  CU_TRY(cudaStreamSynchronize(h->stream));
  h->stream = (cudaStream_t)stream;
  h->cfg.stream = stream;
  return B200DISORT_OK;
}

extern "C" int b200disort_set_layer_split(b200disort_handle* h, int mode) {
  if (!h) return B200DISORT_ERR_INVALID;
  h->layer_split = mode;
  return B200DISORT_OK;
}

extern "C" int b200disort_set_warp_solve_max(b200disort_handle* h, long max_solves) {
  if (!h) return B200DISORT_ERR_INVALID;
  h->warp_solve_max = max_solves;
  return B200DISORT_OK;
}

extern "C" int b200disort_check_status(b200disort_handle* h, const int32_t* status, int* n_bad) {
  if (!h || !status || !n_bad) return B200DISORT_ERR_INVALID;
  CU_TRY(cudaSetDevice(h->cfg.device));
  const long n = (long)h->cfg.nwave * h->cfg.ncol;
  CU_TRY(cudaMemsetAsync(h->d_flag + 1, 0, sizeof(int), h->stream));
  int rc = launch_count_status(h, status, n, h->stream);
  if (rc != B200DISORT_OK) return rc;
  int flags[2] = {0, 0};
  CU_TRY(cudaMemcpyAsync(flags, h->d_flag, sizeof(int) * 2, cudaMemcpyDeviceToHost, h->stream));
  CU_TRY(cudaStreamSynchronize(h->stream));
  *n_bad = flags[1];
  if (flags[0]) {  // a negative / NaN level temperature reached the Planck pre-pass
    CU_TRY(cudaMemsetAsync(h->d_flag, 0, sizeof(int), h->stream));
    return B200DISORT_ERR_SOLVE;
  }
  return flags[1] ? B200DISORT_ERR_SOLVE : B200DISORT_OK;
}

bool b200::is_pinned(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost;
}

static int solve_host_impl(b200disort_handle* h, const double* prop, int nprop, const b200disort_bc* bc, const double* temf,
                           double* flx, double* uu, int32_t* status);

extern "C" int b200disort_solve_host(b200disort_handle* h, const double* prop, int nprop,
                                     const b200disort_bc* bc, const double* temf, double* flx,
                                     double* uu, int32_t* status) {
  const int rc = solve_host_impl(h, prop, nprop, bc, temf, flx, uu, status);
  if (h && rc != B200DISORT_OK && rc != B200DISORT_ERR_SOLVE && rc != B200DISORT_ERR_INVALID) {
    // an error return must not leave copies in flight that still read or write the caller's (and the pinned) buffers
    for (auto& sl : h->slots)
      if (sl.stream) cudaStreamSynchronize(sl.stream);
    cudaStreamSynchronize(h->stream);
  }
  return rc;
}

static int solve_host_impl(b200disort_handle* h, const double* prop, int nprop, const b200disort_bc* bc, const double* temf,
                           double* flx, double* uu, int32_t* status) {
  if (!h || !prop || !flx || nprop < 1) return B200DISORT_ERR_INVALID;
  const b200disort_config& c = h->cfg;
  const bool planck = c.flags & B200DISORT_PLANCK;
  const bool onlyfl = c.flags & B200DISORT_ONLYFL;
  if (planck && !temf) return B200DISORT_ERR_INVALID;
  CU_TRY(cudaSetDevice(c.device));
  h->stats.kernel_launches = 0;
  CU_TRY(cudaEventRecord(h->ev0, h->stream));  // (b200disort_get_stats measures from here)
  using Slot = b200disort_handle::Slot;
  if (!onlyfl || (c.flags & B200DISORT_USRTAU) || !h->inst) {
    // cold path (test_disort-sized batches): stage everything, one launch, copy back
    const size_t ns = (size_t)c.nwave * c.ncol;
    const double* src[9] = {bc ? bc->umu0 : nullptr, bc ? bc->fbeam : nullptr, bc ? bc->albedo : nullptr,
                            bc ? bc->btemp : nullptr, bc ? bc->ttemp : nullptr, bc ? bc->temis : nullptr,
                            bc ? bc->fisot : nullptr, bc ? bc->phi0 : nullptr, nullptr};
    const size_t n_prop = ns * c.nlyr * nprop, n_temf = temf ? (size_t)c.ncol * (c.nlyr + 1) : 0;
    const size_t n_flx = ns * ntau_eff(c) * 2, n_st = (ns * sizeof(int32_t) + 7) / 8;
    Slot& sl = h->slots[0];
    int rc = ensure_dev(&sl.d_buf, &sl.d_bytes, (n_prop + n_temf + 8 * ns + n_flx + n_st) * sizeof(double));
    if (rc != B200DISORT_OK) return rc;
    double* d = sl.d_buf;
    CU_TRY(cudaMemcpyAsync(d, prop, n_prop * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    const double* d_prop = d; d += n_prop;
    const double* d_temf = nullptr;
    if (temf) {
      CU_TRY(cudaMemcpyAsync(d, temf, n_temf * sizeof(double), cudaMemcpyHostToDevice, h->stream));
      d_temf = d; d += n_temf;
    }
    const double* dbc[8] = {nullptr};
    for (int i = 0; i < 8; ++i)
      if (src[i]) {
        CU_TRY(cudaMemcpyAsync(d, src[i], ns * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        dbc[i] = d; d += ns;
      }
    double* d_flx = d; d += n_flx;
    int32_t* d_st = reinterpret_cast<int32_t*>(d);
    b200disort_bc dev_bc{dbc[0], dbc[1], dbc[2], dbc[3], dbc[4], dbc[5], dbc[6], dbc[7]};
    rc = solve_general(h, d_prop, nprop, &dev_bc, d_temf, d_flx, nullptr, d_st);
    if (rc != B200DISORT_OK) return rc;
    std::vector<int32_t> st_host(ns);
    CU_TRY(cudaMemcpyAsync(flx, d_flx, n_flx * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(cudaMemcpyAsync(st_host.data(), d_st, ns * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    if (uu && uu_elems(c)) CU_TRY(cudaMemcpyAsync(uu, h->d_uu, uu_elems(c) * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(cudaStreamSynchronize(h->stream));
    long nbad = 0;
    for (size_t i = 0; i < ns; ++i) nbad += st_host[i] != 0;
    if (status) std::memcpy(status, st_host.data(), ns * sizeof(int32_t));
    return nbad ? B200DISORT_ERR_SOLVE : B200DISORT_OK;
  }
  const long nsolve = (long)c.nwave * c.ncol;
  const int nlev = nlev_out(c);
  const double* bcs[7] = {nullptr};
  if (bc) {
    bcs[0] = bc->umu0; bcs[1] = bc->fbeam; bcs[2] = bc->albedo; bcs[3] = bc->btemp;
    bcs[4] = bc->ttemp; bcs[5] = bc->temis; bcs[6] = bc->fisot;
  }
  int n_bc = 0;
  for (int i = 0; i < 7; ++i) n_bc += bcs[i] ? 1 : 0;

  // chunk = one resident wave of the flux kernel (every CTA slot of the GPU busy once), so the
  // kernels of successive chunks run back to back while the copy engines move the neighbours
  if (h->occ == 0) {
    int occ = 0;
    CU_TRY(h->inst->occupancy(128, &occ));
    h->occ = occ < 1 ? 1 : occ;
  }
  const long wave = (long)h->sm_count * h->occ * 128;  // (launch_flux may be told another CTA size by the tuning knobs)
  // The host path is bound by PCIe, not by the kernels (amars_lw: 0.9 ms of copies against 0.3 ms of solve per
  // wave).  Smaller chunks would shrink the first H2D and the last D2H of a call, which nothing overlaps;
  // measured (profiles/r01, sweep e13): +2 % at half a wave on amars_lw, losses elsewhere and below a quarter
  // wave, so the chunk stays one wave.  B200DISORT_HOST_CHUNK_DIV overrides the divisor (profiling knob).
  static const int env_div = std::getenv("B200DISORT_HOST_CHUNK_DIV") ? std::atoi(std::getenv("B200DISORT_HOST_CHUNK_DIV")) : 0;
  const long div = env_div > 0 ? env_div : kHostChunkDiv;
  const long part = std::max<long>(128 * 16, ((wave / div + 127) / 128) * 128);
  const long chunk = std::min(nsolve, part);
  const long nchunk = (nsolve + chunk - 1) / chunk;
  const size_t in_per = (size_t)c.nlyr * nprop + n_bc;  // doubles per solve
  const size_t out_per = (size_t)nlev * 2;
  const size_t st_doubles = ((size_t)chunk * sizeof(int32_t) + 7) / 8;
  const size_t slot_doubles = (size_t)chunk * (in_per + out_per) + st_doubles;

  bool pinned_in = is_pinned(prop), pinned_out = is_pinned(flx) && (!status || is_pinned(status));
  for (int i = 0; i < 7; ++i)
    if (bcs[i] && !is_pinned(bcs[i])) pinned_in = false;
  const bool stage = !(pinned_in && pinned_out);

  if (!h->ev_pre) CU_TRY(cudaEventCreateWithFlags(&h->ev_pre, cudaEventDisableTiming));
  const int nslot = (int)std::min<long>(b200disort_handle::kSlots, nchunk);
  for (int k = 0; k < nslot; ++k) {
    Slot& sl = h->slots[k];
    if (!sl.stream) CU_TRY(cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking));
    if (!sl.done) CU_TRY(cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming));
    int rc = ensure_dev(&sl.d_buf, &sl.d_bytes, slot_doubles * sizeof(double));
    if (rc != B200DISORT_OK) return rc;
    if (stage && slot_doubles * sizeof(double) > sl.h_bytes) {
      if (sl.h_buf) CU_TRY(cudaFreeHost(sl.h_buf));
      sl.h_buf = nullptr;
      sl.h_bytes = 0;
      CU_TRY(cudaMallocHost(&sl.h_buf, slot_doubles * sizeof(double)));
      sl.h_bytes = slot_doubles * sizeof(double);
    }
    sl.pend_s0 = -1;
  }

  // per-column Planck table once, on the handle's stream; the slots wait for it
  if (planck) {
    const size_t nb = (size_t)c.ncol * (c.nlyr + 1) * sizeof(double);
    int rc = ensure_dev(&h->d_temf, &h->temf_bytes, nb);
    if (rc != B200DISORT_OK) return rc;
    CU_TRY(cudaMemcpyAsync(h->d_temf, temf, nb, cudaMemcpyHostToDevice, h->stream));
    rc = launch_planck(h, h->d_temf, h->stream);
    if (rc != B200DISORT_OK) return rc;
  }
  CU_TRY(cudaEventRecord(h->ev_pre, h->stream));

  int32_t* st_out = status;
  std::vector<int32_t> st_tmp;
  if (!st_out) {
    st_tmp.resize((size_t)nsolve);
    st_out = st_tmp.data();
  }
  auto drain = [&](Slot& sl) -> int {  // finish a slot's previous chunk (copy-out when staging)
    if (sl.pend_s0 < 0) return B200DISORT_OK;
    CU_TRY(cudaEventSynchronize(sl.done));
    if (stage) {
      const double* ho = sl.h_buf + (size_t)chunk * in_per;
      std::memcpy(flx + (size_t)sl.pend_s0 * out_per, ho, (size_t)sl.pend_n * out_per * sizeof(double));
      std::memcpy(st_out + sl.pend_s0, ho + (size_t)chunk * out_per, (size_t)sl.pend_n * sizeof(int32_t));
    }
    sl.pend_s0 = -1;
    return B200DISORT_OK;
  };

  for (long ck = 0; ck < nchunk; ++ck) {
    Slot& sl = h->slots[ck % nslot];
    int rc = drain(sl);
    if (rc != B200DISORT_OK) return rc;
    const long s0 = ck * chunk, n = std::min(chunk, nsolve - s0);
    double* d_in = sl.d_buf;
    double* d_flx = sl.d_buf + (size_t)chunk * in_per;
    int32_t* d_st = reinterpret_cast<int32_t*>(d_flx + (size_t)chunk * out_per);
    // inputs: [prop chunk | bc_0 chunk | bc_1 chunk ...]
    const size_t np = (size_t)n * c.nlyr * nprop;
    if (stage) {
      std::memcpy(sl.h_buf, prop + (size_t)s0 * c.nlyr * nprop, np * sizeof(double));
      size_t off = (size_t)chunk * c.nlyr * nprop;
      for (int i = 0; i < 7; ++i)
        if (bcs[i]) {
          std::memcpy(sl.h_buf + off, bcs[i] + s0, (size_t)n * sizeof(double));
          off += chunk;
        }
      CU_TRY(cudaMemcpyAsync(d_in, sl.h_buf, (size_t)chunk * in_per * sizeof(double), cudaMemcpyHostToDevice, sl.stream));
    } else {
      CU_TRY(cudaMemcpyAsync(d_in, prop + (size_t)s0 * c.nlyr * nprop, np * sizeof(double), cudaMemcpyHostToDevice, sl.stream));
      size_t off = (size_t)chunk * c.nlyr * nprop;
      for (int i = 0; i < 7; ++i)
        if (bcs[i]) {
          CU_TRY(cudaMemcpyAsync(d_in + off, bcs[i] + s0, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, sl.stream));
          off += chunk;
        }
    }
    SolveParams p{};
    p.nwave = c.nwave; p.ncol = c.ncol; p.nlyr = c.nlyr; p.nprop = nprop; p.nstr = c.nstr;
    p.flags = c.flags;
    p.prop = d_in;
    {
      const double** dst[7] = {&p.umu0, &p.fbeam, &p.albedo, &p.btemp, &p.ttemp, &p.temis, &p.fisot};
      size_t off = (size_t)chunk * c.nlyr * nprop;
      for (int i = 0; i < 7; ++i)
        if (bcs[i]) {
          *dst[i] = d_in + off;
          off += chunk;
        }
    }
    p.flx = d_flx;
    p.status = d_st;
    p.wave_lower = h->d_wlo; p.wave_upper = h->d_whi; p.wave2band = h->d_w2b;
    p.pkag = h->d_pkag;
    p.s0 = s0;
    p.nlocal = n;
    if (ck < nslot) CU_TRY(cudaStreamWaitEvent(sl.stream, h->ev_pre, 0));
    rc = launch_flux(h, p, sl.stream, &sl.d_ws, &sl.ws_bytes, true);  // timed: b200disort_get_stats sees the host route too
    if (rc != B200DISORT_OK) return rc;
    if (stage) {
      CU_TRY(cudaMemcpyAsync(sl.h_buf + (size_t)chunk * in_per, d_flx,
                             ((size_t)chunk * out_per + st_doubles) * sizeof(double), cudaMemcpyDeviceToHost, sl.stream));
    } else {
      CU_TRY(cudaMemcpyAsync(flx + (size_t)s0 * out_per, d_flx, (size_t)n * out_per * sizeof(double),
                             cudaMemcpyDeviceToHost, sl.stream));
      if (status)
        CU_TRY(cudaMemcpyAsync(status + s0, d_st, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, sl.stream));
      else
        CU_TRY(cudaMemcpyAsync(st_out + s0, d_st, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, sl.stream));
    }
    CU_TRY(cudaEventRecord(sl.done, sl.stream));
    sl.pend_s0 = s0;
    sl.pend_n = n;
  }
  for (int k = 0; k < nslot; ++k) {
    int rc = drain(h->slots[k]);
    if (rc != B200DISORT_OK) return rc;
  }
  CU_TRY(cudaStreamSynchronize(h->stream));
  // a negative / NaN level temperature seen by the Planck pre-pass
  int flag = 0;
  CU_TRY(cudaMemcpy(&flag, h->d_flag, sizeof(int), cudaMemcpyDeviceToHost));
  if (flag) CU_TRY(cudaMemset(h->d_flag, 0, sizeof(int)));
  long nbad = 0;
  for (long i = 0; i < nsolve; ++i) nbad += st_out[i] != 0;
  return (flag || nbad) ? B200DISORT_ERR_SOLVE : B200DISORT_OK;
}

extern "C" int b200disort_get_rad(b200disort_handle* h, double* uu) {
  if (!h || !uu) return B200DISORT_ERR_INVALID;
  if (!h->have_uu) return B200DISORT_ERR_INVALID;  // ONLYFL, or no solve yet
  CU_TRY(cudaSetDevice(h->cfg.device));
  CU_TRY(cudaMemcpyAsync(uu, h->d_uu, uu_elems(h->cfg) * sizeof(double), cudaMemcpyDefault, h->stream));
  CU_TRY(cudaStreamSynchronize(h->stream));
  return B200DISORT_OK;
}

int b200::launch_reduce_band(b200disort_handle* h, int nwave, long inner, const double* flx, const double* weights,
                             double* bflx, cudaStream_t stream) {
  const int block = 256;
  const long grid = std::min<long>((inner + block - 1) / block, (long)h->sm_count * 8);
  const long want = (long)h->sm_count * 8;  // CTAs that fill the machine
  int parts = (int)std::min<long>(std::max<long>(1, want / std::max<long>(1, (inner + block - 1) / block)),
                                  std::max(1, nwave / 16));
  if (parts <= 1) {
    reduce_band_kernel<<<dim3((unsigned)grid, 1), block, 0, stream>>>(nwave, inner, flx, weights, bflx);
    CU_TRY(cudaGetLastError());
    return B200DISORT_OK;
  }
  const size_t need = (size_t)parts * inner * sizeof(double);
  if (need > h->part_bytes) {
    if (h->d_part) CU_TRY(cudaFree(h->d_part));
    h->d_part = nullptr;
    h->part_bytes = 0;
    CU_TRY(cudaMalloc(&h->d_part, need));
    h->part_bytes = need;
  }
  reduce_band_kernel<<<dim3((unsigned)grid, (unsigned)parts), block, 0, stream>>>(nwave, inner, flx, weights, h->d_part);
  CU_TRY(cudaGetLastError());
  reduce_parts_kernel<<<(unsigned)grid, block, 0, stream>>>(parts, inner, h->d_part, bflx);
  CU_TRY(cudaGetLastError());
  return B200DISORT_OK;
}

extern "C" int b200disort_reduce_band(b200disort_handle* h, const double* flx, const double* weights,
                                      double* bflx) {
  if (!h || !flx || !weights || !bflx) return B200DISORT_ERR_INVALID;
  const b200disort_config& c = h->cfg;
  CU_TRY(cudaSetDevice(c.device));
  return launch_reduce_band(h, c.nwave, (long)c.ncol * nlev_out(c) * 2, flx, weights, bflx, h->stream);
}

extern "C" int b200disort_fp64_peak(int device, double* tflops) {
  if (!tflops) return B200DISORT_ERR_INVALID;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return B200DISORT_ERR_CUDA;
  CU_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU_TRY(cudaGetDeviceProperties(&prop, device));
  double* d = nullptr;
  CU_TRY(cudaMalloc(&d, 8));
  cudaEvent_t e0, e1;
  CU_TRY(cudaEventCreate(&e0));
  CU_TRY(cudaEventCreate(&e1));
  const int nb = prop.multiProcessorCount * 4, tpb = 512, iters = 8192;
  dfma_peak_kernel<<<nb, tpb>>>(d, 256, 1.0000001, 1e-9);  // warm-up
  CU_TRY(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    CU_TRY(cudaEventRecord(e0));
    dfma_peak_kernel<<<nb, tpb>>>(d, iters, 1.0000001, 1e-9);
    CU_TRY(cudaEventRecord(e1));
    CU_TRY(cudaEventSynchronize(e1));
    float ms;
    CU_TRY(cudaEventElapsedTime(&ms, e0, e1));
    best = std::min(best, ms);
  }
  *tflops = 2.0 * (double)nb * tpb * 64.0 * iters / (best * 1e-3) / 1e12;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  return B200DISORT_OK;
}
