// adding_team.cuh -- the adding recursion + back sweep of the layer-split form (c_set_matrix + c_solve0 + c_fluxes of
// cdisort in this project's formulation, see disort_adding_step.inc) with a TEAM of four lanes per solve.
//
// Why: as one thread per solve this phase needs ~270 live doubles at nstr = 16 (stack operator Ra, the layer's R and T, the
// system matrix and its nine right-hand sides), spills 2.4 KB per thread, and with every resident thread streaming its
// hand-over and back-sweep records through the L2 the spills go to HBM: ncu showed the thread form HBM-bound (4.6 TB/s,
// 253 KB per solve against 121 KB of records, FP64 pipe 12 % busy, long_scoreboard 13.9 stalls per issue) and as slow as
// the whole layer phase although it executes a third of its flops.  Here the nn x nn operators are distributed by rows
// (lane r of a team owns rows R r .. R r + R - 1, R = nn / 4), which leaves ~100 doubles per lane -- no spills:
//   * the layer's hand-over record of the warp's 8 solves (Rl | Tl packed | s+ | s- | expbea: 90 x 8 doubles at nstr = 16,
//     contiguous in the [group of 8 solves][layer][element][solve] layout the layer kernel writes) is staged in shared
//     memory by ONE bulk copy (cp.async.bulk + mbarrier, two buffers: the copy of layer lc + 2 is issued as soon as
//     layer lc has been read), so no lane ever waits on a global load in the forward sweep;
//   * operands every lane of the team needs (the stack operator, the pivot rows of the elimination, M1 | v1) are
//     exchanged through per-team slabs in shared memory and read back with broadcast loads (row stride nn + 2 doubles
//     and a team stride of an odd number of 16-byte chunks keep the row stores of the 8 lanes of a quarter warp and
//     the broadcast reads of the 8 teams of a warp on distinct banks);
//   * A = I - R Ra is solved against [T | R da + s+] by Gauss-Jordan elimination in the natural order -- the owner of row
//     k normalises it and publishes it, every lane updates its own rows.  A is a non-singular M-matrix in the physical
//     case and its symmetric part is positive definite in general (R and Ra are contractions in the scaled variables), so
//     no pivoting is needed; pivots at or below 1e-290 are reported as ST_SINGULAR through PivotCheck;
//   * the back sweep u+(lev) = M1 u+(lev+1) + v1 keeps the rows of a level's record one level ahead in registers and
//     all-gathers u+ with shuffles.
// Included by flux_inst.cu for nn % 4 == 0, nn >= 8.
#pragma once

#include "disort_device.cuh"

namespace b200 {

// stage buffers for the hand-over records per warp (the record is copied to registers at the top of a layer, so one buffer
// already gives the copy of the next layer a whole layer of arithmetic; two: e43)
#ifndef B200_TEAM_STAGES
#define B200_TEAM_STAGES 2
#endif
constexpr int kTeam = 4;            // lanes per solve
constexpr int kGrp = 32 / kTeam;    // solves per warp = solves per group of the hand-over / record buffers

template <int NN>
struct TeamLayout {
  static constexpr int R = NN / kTeam;
  static constexpr int NOPS = kOpsDoubles<NN>;
  static constexpr int NSAVE = NN * NN + 2 * NN + 1;
  static constexpr int RS = NN + 2;                       // slab row stride (doubles): rows start on 16-byte boundaries
  static constexpr int kPivDoubles = 2 * NN + 2;          // pivot row: A part (NN) | right-hand sides (NN + 1) | pad
  static constexpr int kRaOff = 0;                        // Ra | da            NN rows x RS
  static constexpr int kXOff = NN * RS;                   // M1 | v1            NN rows x RS
  // the pivot rows of the elimination (two buffers of -- R == 2 -- two rows) share the storage of M1 | v1, which is written
  // after the elimination (one warp barrier in between)
  static constexpr int kPivOff = kXOff;
  static constexpr int kTeamRaw = 2 * NN * RS;
  static_assert((R == 2 ? 4 : 2) * kPivDoubles <= NN * RS, "pivot rows fit the M1 | v1 slab");
  // team stride: an odd number of 16-byte chunks (mod 8 chunks = 128 bytes)
  static constexpr int kTeamDoubles = ((kTeamRaw / 2) % 2 == 1) ? kTeamRaw : kTeamRaw + 2;
  static constexpr int kStageDoubles = NOPS * kGrp;       // one layer of the warp's group
  static constexpr int kStageBytes = kStageDoubles * 8;
  static constexpr int kStages = B200_TEAM_STAGES;
  static constexpr int kWarpDoubles = kStages * kStageDoubles + 2 + kGrp * kTeamDoubles;  // stages | 2 barriers | slabs
  static_assert(kStageBytes % 16 == 0, "bulk copies move multiples of 16 bytes");
  static_assert(kWarpDoubles % 2 == 0 && kTeamDoubles % 2 == 0 && RS % 2 == 0, "16-byte alignment of the slabs");
};

__device__ __forceinline__ double2 b200_ld2(const double* p) { return *reinterpret_cast<const double2*>(p); }
__device__ __forceinline__ void b200_st2(double* p, double a, double b) { *reinterpret_cast<double2*>(p) = make_double2(a, b); }

template <int NN, int BAND, typename QT>
__device__ __forceinline__ void team_adding_sweep(const SolveParams& p, const QT& qt, long c0, long cn, const int* cutv,
                                                  const int* stl, const double* ops, double* rec, double* smem) {
  using L = TeamLayout<NN>;
  constexpr int R = L::R, NS = NN * (NN + 1) / 2, NOPS = L::NOPS, NSAVE = L::NSAVE, RS = L::RS;
  const int lane = (int)(threadIdx.x & 31), wid = (int)(threadIdx.x >> 5);
  const int r = lane & (kTeam - 1), tj = lane >> 2;
  const long grp = (long)blockIdx.x * (blockDim.x >> 5) + wid;  // this warp's group of 8 solves (chunk-local)
  if (grp * kGrp >= cn) return;                                 // (warp-uniform)
  const long i = grp * kGrp + tj;
  const bool valid = i < cn;      // teams beyond the end of the chunk shadow its last solve and emit nothing
  const long s = c0 + (valid ? i : cn - 1);
  const int nlyr = p.nlyr, nprop = p.nprop;

  double* wbase = smem + (size_t)wid * L::kWarpDoubles;
  double* stage0 = wbase;
  constexpr int S = L::kStages;
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(wbase + S * L::kStageDoubles);
  double* slab = wbase + S * L::kStageDoubles + 2 + (size_t)tj * L::kTeamDoubles;
  double* RaS = slab + L::kRaOff;
  double* Xs = slab + L::kXOff;
  double* Pv = slab + L::kPivOff;
  if (lane == 0) {
    b200_mbar_init(bars, 1);
    b200_mbar_init(bars + 1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();

  SolveCommon<NN> cm;
  solve_common_any<NN>(p, qt, s, cm, true, false);
  const double* pr = p.prop + s * nlyr * nprop;
  const int cv = valid ? cutv[i] : cutv[cn - 1], sl = valid ? stl[i] : stl[cn - 1];
  const int ncut = cv & 0xffff;
  const bool lyrcut = (cv >> 16) != 0;
  int st = status_merge_u(cm.st, sl == 0x7fffffff ? (int)ST_OK : sl);
  int ncw = ncut;  // the warp walks to its deepest cut-off
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) ncw = max(ncw, __shfl_xor_sync(0xffffffffu, ncw, o));

  const double* ops_g = ops + grp * ((long)nlyr * NOPS * kGrp);
  double* rec_g = rec + grp * ((long)nlyr * NSAVE * kGrp) + tj;
  if (lane == 0) {
    if (ncw > 0) b200_tma_load(stage0, ops_g, L::kStageBytes, bars);
    if (S > 1 && ncw > 1) b200_tma_load(stage0 + L::kStageDoubles, ops_g + (long)NOPS * kGrp, L::kStageBytes, bars + 1);
  }

  // rows of this lane, their packed-symmetric offsets and stream scalings
  int off[R][NN];
  double tq[R];
#pragma unroll
  for (int a = 0; a < R; ++a) {
    const int ia = R * r + a;
    tq[a] = qt.t[ia];
#pragma unroll
    for (int k = 0; k < NN; ++k) off[a][k] = (ia >= k ? ia * (ia + 1) / 2 + k : k * (k + 1) / 2 + ia) * kGrp;
  }
  // stack state: rows of Ra in registers and (all rows, with da as column NN) in the team's slab
  double ra[R][NN], dav[R];
#pragma unroll
  for (int a = 0; a < R; ++a) {
    dav[a] = tq[a] * (cm.fisot + cm.tplank);
#pragma unroll
    for (int k = 0; k < NN; ++k) ra[a][k] = 0.0;
#pragma unroll
    for (int k = 0; k < NN; k += 2) b200_st2(RaS + (R * r + a) * RS + k, 0.0, 0.0);
    RaS[(R * r + a) * RS + NN] = dav[a];
  }
  __syncwarp();

  PivotCheck piv;
  double ebot_last = cm.beam ? 1.0 : 0.0;
  const double fb = cm.umu0 * cm.fbeam;
  for (int lc = 0; lc < ncw; ++lc) {
    const bool act = lc < ncut;
    const int buf = lc % S;
    b200_mbar_wait(bars + buf, (unsigned)(lc / S) & 1u);
    const double* sg = stage0 + buf * L::kStageDoubles + tj;
    double rl[R][NN], tl[R][NN], spv[R], smv[R];
#pragma unroll
    for (int a = 0; a < R; ++a) {
#pragma unroll
      for (int k = 0; k < NN; ++k) {
        rl[a][k] = sg[off[a][k]];
        tl[a][k] = sg[NS * kGrp + off[a][k]];
      }
      spv[a] = sg[(2 * NS + R * r + a) * kGrp];
      smv[a] = sg[(2 * NS + NN + R * r + a) * kGrp];
    }
    const double etop = sg[(2 * NS + 2 * NN) * kGrp];
    if (act) ebot_last = sg[(2 * NS + 2 * NN + 1) * kGrp];
    __syncwarp();  // every lane has taken its values out of the stage: refill it with layer lc + S
    if (lane == 0 && lc + S < ncw)
      b200_tma_load(stage0 + buf * L::kStageDoubles, ops_g + (long)(lc + S) * NOPS * kGrp, L::kStageBytes, bars + buf);

    // ---- A = I - R Ra, right-hand sides [T | R da + s+]; and TR = T Ra, td = T da for the stack update below
    double A[R][NN], B[R][NN + 1], TR[R][NN], td[R];
#pragma unroll
    for (int a = 0; a < R; ++a) {
#pragma unroll
      for (int c = 0; c < NN; ++c) {
        A[a][c] = (R * r + a == c) ? 1.0 : 0.0;
        B[a][c] = tl[a][c];
        TR[a][c] = 0.0;
      }
      B[a][NN] = spv[a];
      td[a] = 0.0;
    }
#pragma unroll
    for (int k = 0; k < NN; ++k) {
      double row[NN + 2];
#pragma unroll
      for (int c = 0; c < NN + 2; c += 2) {
        const double2 v = b200_ld2(RaS + k * RS + c);
        row[c] = v.x;
        row[c + 1] = v.y;
      }
#pragma unroll
      for (int a = 0; a < R; ++a) {
#pragma unroll
        for (int c = 0; c < NN; ++c) {
          A[a][c] = fma(-rl[a][k], row[c], A[a][c]);
          TR[a][c] = fma(tl[a][k], row[c], TR[a][c]);
        }
        B[a][NN] = fma(rl[a][k], row[NN], B[a][NN]);
        td[a] = fma(tl[a][k], row[NN], td[a]);
      }
    }

    // ---- Gauss-Jordan elimination in the natural order: [A | B] -> [I | A^-1 B]
    if constexpr (R == 2) {
      // by 2 x 2 pivot blocks: both rows of a block belong to ONE lane, which applies the inverse of its diagonal block
      // (one reciprocal of the determinant: four dependent reciprocal chains per layer instead of eight, and half the
      // publish -> synchronise -> read round trips) and publishes both rows
#pragma unroll
      for (int q = 0; q < kTeam; ++q) {
        constexpr int PD = L::kPivDoubles;
        double* pv = Pv + (q & 1) * 2 * PD;
        const int k0 = 2 * q;
        if (r == q) {
          const double a00 = A[0][k0], a01 = A[0][k0 + 1], a10 = A[1][k0], a11 = A[1][k0 + 1];
          const double det = fma(a00, a11, -a01 * a10);
          piv.add(a00);
          piv.add(det);
          const double rd = B200_RCP(det);
          const double i00 = a11 * rd, i01 = -a01 * rd, i10 = -a10 * rd, i11 = a00 * rd;
#pragma unroll
          for (int c = 0; c < NN; ++c)
            if (c > k0 + 1) {
              const double x = A[0][c], y = A[1][c];
              A[0][c] = fma(i00, x, i01 * y);
              A[1][c] = fma(i10, x, i11 * y);
            }
#pragma unroll
          for (int c = 0; c <= NN; ++c) {
            const double x = B[0][c], y = B[1][c];
            B[0][c] = fma(i00, x, i01 * y);
            B[1][c] = fma(i10, x, i11 * y);
          }
#pragma unroll
          for (int c = 0; c < 2 * NN + 2; c += 2)
            if (c + 1 > k0 + 1) {
              b200_st2(pv + c, c < NN ? A[0][c] : B[0][c - NN], c + 1 < NN ? A[0][c + 1] : (c + 1 - NN <= NN ? B[0][c + 1 - NN] : 0.0));
              b200_st2(pv + PD + c, c < NN ? A[1][c] : B[1][c - NN], c + 1 < NN ? A[1][c + 1] : (c + 1 - NN <= NN ? B[1][c + 1 - NN] : 0.0));
            }
        }
        __syncwarp();
        double p0[2 * NN + 2], p1[2 * NN + 2];
#pragma unroll
        for (int c = 0; c < 2 * NN + 2; c += 2)
          if (c + 1 > k0 + 1) {
            const double2 u = b200_ld2(pv + c), v = b200_ld2(pv + PD + c);
            p0[c] = u.x; p0[c + 1] = u.y;
            p1[c] = v.x; p1[c + 1] = v.y;
          }
#pragma unroll
        for (int a = 0; a < R; ++a) {
          const double f0 = (r == q) ? 0.0 : A[a][k0], f1 = (r == q) ? 0.0 : A[a][k0 + 1];
#pragma unroll
          for (int c = 0; c < NN; ++c)
            if (c > k0 + 1) A[a][c] = fma(-f1, p1[c], fma(-f0, p0[c], A[a][c]));
#pragma unroll
          for (int c = 0; c <= NN; ++c) B[a][c] = fma(-f1, p1[NN + c], fma(-f0, p0[NN + c], B[a][c]));
        }
      }
    } else {
#pragma unroll
    for (int k = 0; k < NN; ++k) {
      const int owner = k / R, slot = k % R;
      double* pv = Pv + (k & 1) * L::kPivDoubles;
      if (r == owner) {
        const double pk = A[slot][k];
        piv.add(pk);
        const double rp = B200_RCP(pk);
#pragma unroll
        for (int c = 0; c < NN; ++c)
          if (c > k) {
            A[slot][c] *= rp;
            pv[c] = A[slot][c];
          }
#pragma unroll
        for (int c = 0; c <= NN; ++c) {
          B[slot][c] *= rp;
          pv[NN + c] = B[slot][c];
        }
      }
      __syncwarp();
      double prow[2 * NN + 2];
#pragma unroll
      for (int c = 0; c < 2 * NN + 2; c += 2) {
        if (c + 1 > k) {  // (elements <= k of the A part are not used)
          const double2 v = b200_ld2(pv + c);
          prow[c] = v.x;
          prow[c + 1] = v.y;
        }
      }
#pragma unroll
      for (int a = 0; a < R; ++a) {
        const double f = (r == owner && a == slot) ? 0.0 : A[a][k];
#pragma unroll
        for (int c = 0; c < NN; ++c)
          if (c > k) A[a][c] = fma(-f, prow[c], A[a][c]);
#pragma unroll
        for (int c = 0; c <= NN; ++c) B[a][c] = fma(-f, prow[NN + c], B[a][c]);
      }
    }
    }
    // B = [M1 | v1]

    // ---- this level's back-sweep record: M1, v1, c^T Ra, c^T da (c = 2 pi t)
    {
      double cdp = 0.0;
#pragma unroll
      for (int a = 0; a < R; ++a) cdp = fma(tq[a], dav[a], cdp);
      cdp += __shfl_xor_sync(0xffffffffu, cdp, 1);
      cdp += __shfl_xor_sync(0xffffffffu, cdp, 2);
      if (act) {
        double* w = rec_g + (long)lc * NSAVE * kGrp;
#pragma unroll
        for (int a = 0; a < R; ++a) {
          const int ia = R * r + a;
          double cr = 0.0;
#pragma unroll
          for (int c = 0; c < NN; ++c) {
            __stcs(&w[(long)(ia * NN + c) * kGrp], B[a][c]);
            cr = fma(qt.t[c], ra[a][c], cr);  // (Ra is symmetric: column ia = row ia)
          }
          __stcs(&w[(long)(NN * NN + ia) * kGrp], B[a][NN]);
          __stcs(&w[(long)(NN * NN + NN + ia) * kGrp], 2.0 * kPi * cr);
        }
        if (r == 0) __stcs(&w[(long)(NN * NN + 2 * NN) * kGrp], 2.0 * kPi * cdp + fb * etop);
      }
    }

    __syncwarp();  // (the last pivot rows have been read: their storage takes M1 | v1)
    // ---- new stack: Ra' = R + T Ra M1 = R + TR M1,  da' = s- + T (da + Ra v1) = s- + T da + TR v1  (TR = T Ra and T da were
    //      formed with the rows of Ra | da read for A above: one broadcast of [M1 | v1] instead of two exchanges)
#pragma unroll
    for (int a = 0; a < R; ++a) {
#pragma unroll
      for (int c = 0; c < NN; c += 2) b200_st2(Xs + (R * r + a) * RS + c, B[a][c], B[a][c + 1]);
      Xs[(R * r + a) * RS + NN] = B[a][NN];
    }
    __syncwarp();
    double rn[R][NN], dn[R];
#pragma unroll
    for (int a = 0; a < R; ++a) {
#pragma unroll
      for (int c = 0; c < NN; ++c) rn[a][c] = rl[a][c];
      dn[a] = smv[a] + td[a];
    }
#pragma unroll
    for (int k = 0; k < NN; ++k) {
      double row[NN + 2];
#pragma unroll
      for (int c = 0; c < NN + 2; c += 2) {
        const double2 v = b200_ld2(Xs + k * RS + c);
        row[c] = v.x;
        row[c + 1] = v.y;
      }
#pragma unroll
      for (int a = 0; a < R; ++a) {
#pragma unroll
        for (int c = 0; c < NN; ++c) rn[a][c] = fma(TR[a][k], row[c], rn[a][c]);
        dn[a] = fma(TR[a][k], row[NN], dn[a]);
      }
    }
    if (act) {  // (a team whose medium is cut off higher up keeps its stack)
#pragma unroll
      for (int a = 0; a < R; ++a) {
        dav[a] = dn[a];
#pragma unroll
        for (int c = 0; c < NN; ++c) ra[a][c] = rn[a][c];
#pragma unroll
        for (int c = 0; c < NN; c += 2) b200_st2(RaS + (R * r + a) * RS + c, rn[a][c], rn[a][c + 1]);
        RaS[(R * r + a) * RS + NN] = dn[a];
      }
    }
    __syncwarp();
  }
  // pivots of all four lanes
  {
    unsigned w = piv.worst;
    w = max(w, __shfl_xor_sync(0xffffffffu, w, 1));
    w = max(w, __shfl_xor_sync(0xffffffffu, w, 2));
    piv.worst = w;
  }
  if (st == ST_OK && piv.bad()) st = ST_SINGULAR;
  const double etop = ebot_last;  // expbea at the lowest level reached

  // ---- bottom boundary (Lambertian surface), as in solve_flux_one
  double* out = (BAND != kBandOff) ? p.bflx_t + cm.j : p.flx + s * (nlyr + 1) * 2;
  // (batches of fewer than 32 columns: the host keeps p.nrep replicas of the band-flux buffer so that the reductions of the
  // many spectral points of one column do not serialise on the same L2 addresses; CTA b adds into replica b % nrep)
  if (BAND != kBandOff && p.nrep > 1) out += (long)(blockIdx.x % (unsigned)p.nrep) * p.rep_stride;
  const double wg = (BAND != kBandOff) ? p.weights[cm.n] : 1.0;
  const long ostr = p.ncol;
  constexpr int EB = (BAND != kBandOff) ? kBandAtomic : kBandOff;
  const bool bad = (st != ST_OK);
  const double nanv = nan("");
  const bool emit = valid && r == 0;
  double up[NN];  // u+ at the level below the one being evaluated (all streams, every lane)
  {
    double tda = 0.0, tRt = 0.0;
#pragma unroll
    for (int a = 0; a < R; ++a) {
      tda = fma(tq[a], dav[a], tda);
      double v = 0.0;
#pragma unroll
      for (int k = 0; k < NN; ++k) v = fma(ra[a][k], qt.t[k], v);
      tRt = fma(tq[a], v, tRt);
    }
    tda += __shfl_xor_sync(0xffffffffu, tda, 1);
    tRt += __shfl_xor_sync(0xffffffffu, tRt, 1);
    tda += __shfl_xor_sync(0xffffffffu, tda, 2);
    tRt += __shfl_xor_sync(0xffffffffu, tRt, 2);
    double xi = 0.0;
    if (!lyrcut) {
      const double g = cm.albedo * cm.umu0 * cm.fbeam / kPi * etop + (1.0 - cm.albedo) * cm.bplank;
      xi = (2.0 * cm.albedo * tda + g) / (1.0 - 2.0 * cm.albedo * tRt);
    }
    double fup = 0.0, fdn = 0.0;
#pragma unroll
    for (int k = 0; k < NN; ++k) up[k] = qt.t[k] * xi;
#pragma unroll
    for (int a = 0; a < R; ++a) {
      fup = fma(tq[a], tq[a] * xi, fup);
      double v = dav[a];
#pragma unroll
      for (int k = 0; k < NN; ++k) v = fma(ra[a][k], up[k], v);
      fdn = fma(tq[a], v, fdn);
    }
    fup += __shfl_xor_sync(0xffffffffu, fup, 1);
    fdn += __shfl_xor_sync(0xffffffffu, fdn, 1);
    fup += __shfl_xor_sync(0xffffffffu, fup, 2);
    fdn += __shfl_xor_sync(0xffffffffu, fdn, 2);
    if (emit) {
      int nsame = 0;  // levels below the cut-off separated from it by zero-thickness layers only (see solve_flux_one)
      if (lyrcut)
        while (ncut + nsame < nlyr && pr[(long)(nlyr - 1 - (ncut + nsame)) * nprop] == 0.0) ++nsame;
      if constexpr (BAND == kBandOff)
        for (int lev = nlyr; lev > ncut + nsame; --lev) {
          out[(nlyr - lev) * 2 + 0] = 0.0;
          out[(nlyr - lev) * 2 + 1] = 0.0;
        }
      for (int lev = ncut + nsame; lev >= ncut; --lev)
        emit_flux<EB>(out, nlyr - lev, ostr, wg, bad ? nanv : 2.0 * kPi * fup, bad ? nanv : 2.0 * kPi * fdn + fb * etop, true, 0);
    }
  }

  // ---- back sweep (c_fluxes): the rows of level lev - 1 are loaded while level lev is evaluated
  constexpr int NREC = R * (NN + 2) + 1;
  auto load_rec = [&](double(&q)[NREC], int lev) {
    const double* w = rec_g + (long)lev * NSAVE * kGrp;
#pragma unroll
    for (int a = 0; a < R; ++a) {
      const int ia = R * r + a;
#pragma unroll
      for (int c = 0; c < NN; ++c) q[a * (NN + 2) + c] = __ldcs(&w[(long)(ia * NN + c) * kGrp]);
      q[a * (NN + 2) + NN] = __ldcs(&w[(long)(NN * NN + ia) * kGrp]);
      q[a * (NN + 2) + NN + 1] = __ldcs(&w[(long)(NN * NN + NN + ia) * kGrp]);
    }
    q[NREC - 1] = __ldcs(&w[(long)(NN * NN + 2 * NN) * kGrp]);
  };
  double cur[NREC], nxt[NREC];
  if (ncut > 0) load_rec(cur, ncut - 1);
  constexpr int kRecLines = (NSAVE * kGrp * 8 + 127) / 128;  // 128-byte lines of the warp's record of one level
  for (int lev = ncw - 1; lev >= 0; --lev) {  // (warp-uniform trip count: the all-gather below is a warp operation)
    const bool own = lev < ncut;
    if (lev >= 4) {  // the record four levels up is pulled into the L2 behind this level's arithmetic
      const char* pl = reinterpret_cast<const char*>(rec_g - tj + (long)(lev - 4) * NSAVE * kGrp);
#pragma unroll
      for (int l0 = 0; l0 < kRecLines; l0 += 32)
        if (l0 + lane < kRecLines) asm volatile("prefetch.global.L2 [%0];" ::"l"(pl + (long)(l0 + lane) * 128));
    }
    if (own && lev > 0) load_rec(nxt, lev - 1);
    double nu[R];
    double fup = 0.0, fdn = 0.0;
#pragma unroll
    for (int a = 0; a < R; ++a) {
      double v = cur[a * (NN + 2) + NN];
#pragma unroll
      for (int c = 0; c < NN; ++c) v = fma(cur[a * (NN + 2) + c], up[c], v);
      nu[a] = v;
      fup = fma(tq[a], v, fup);
      fdn = fma(cur[a * (NN + 2) + NN + 1], v, fdn);
    }
    fup += __shfl_xor_sync(0xffffffffu, fup, 1);
    fdn += __shfl_xor_sync(0xffffffffu, fdn, 1);
    fup += __shfl_xor_sync(0xffffffffu, fup, 2);
    fdn += __shfl_xor_sync(0xffffffffu, fdn, 2);
    fdn += cur[NREC - 1];
    // all-gather of u+(lev): stream R q + a lives in lane q of the team
#pragma unroll
    for (int q = 0; q < kTeam; ++q)
#pragma unroll
      for (int a = 0; a < R; ++a) {
        const double v = __shfl_sync(0xffffffffu, nu[a], (lane & ~(kTeam - 1)) | q);
        if (own) up[R * q + a] = v;
      }
    if (own && emit) emit_flux<EB>(out, nlyr - lev, ostr, wg, bad ? nanv : 2.0 * kPi * fup, bad ? nanv : fdn, true, 0);
    if (own) {
#pragma unroll
      for (int e = 0; e < NREC; ++e) cur[e] = nxt[e];
    }
  }
  if (emit) p.status[s] = st;
}

}  // namespace b200
