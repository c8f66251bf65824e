#!/usr/bin/env python
"""bench.py -- DISORT solves/s on B200 (BASELINE.json metric), one process per GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload amars_lw|amars_sw|lbl|gcm] [--shard col|wave]
  python bench.py --impl reference ...      # the CPU path (oracle restatement of cdisort) on host cores

Default workload = the configuration the metric is quoted on: amars_lw shape (nstr = 8, 80 layers, 32 correlated-k
g-points x 16384 columns per GPU, thermal emission), driven the way pyharp drives it -- as a radiation BAND
(src/radiation/radiation_band.cpp:75-128, examples/amars_lw.cpp:62-88): a "step" is one pass of the per-band hot
path over one batch of synthetic atmospheres,

    conc / pres / temp -> RFM correlated-k optics -> layer2level + Planck pre-pass -> DISORT flux solve of every
    (g-point, column) -> band integration sum_g w_g F_g (fused into the solve kernel's flux epilogue)

through ONE C-ABI call (include/b200band.h).  Columns are sharded over ranks (weak scaling: fixed columns per GPU);
the band integration is then rank-local and no collective is on the data path.

  value      solves/s with the band's inputs resident in HBM (b200_band_forward; CUDA events, max over ranks)
  e2e        the same call with HOST tensors (b200_band_forward_host): H2D of conc / pres / temp / boundary fields and
             D2H of the (ncol, nlyr+1, 2) band flux inside the timed region
  kernel     the flux-solve kernel alone (events around each launch, b200disort_get_stats)
  roofline   that kernel against the measured FP64 (DFMA) peak: `frac` = executed FP64 flops (ncu count per solve of the
             committed capture of exactly these sources, profiles/r02/traffic.json) x this run's rate / peak;
             `algorithmic_equiv` = the reference algorithm's flop count W at this rate, for comparison
  other_workloads   BASELINE.json configs 2, 4, 5 (amars_sw, line-by-line, full-size GCM batch) on the same box
  opacity    the optics kernels against the HBM copy peak
  cpu_baseline  the oracle on the host cores on a bounded sample

`--workload X` (X != amars_lw) measures the solver-level step of that shape instead (prop resident in HBM ->
Planck pre-pass -> solve + fused band integration; e2e through Disort.forward with host tensors);
`--workload lbl --shard wave` is the spectral-point sharding of BASELINE config 4: 10^5 wavenumbers x 1 column
split over the ranks (strong scaling), one NCCL all-reduce of the band flux per step (b200disort_allreduce_band).
"""
import argparse
import hashlib
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "DISORT solves/sec (nstr=8, 80 layers)"
DEFAULT_COLS = {"amars_lw": 16384, "amars_sw": 1024, "lbl": 8, "gcm": 65536}
DEFAULT_NWAVE = {"amars_lw": 32, "amars_sw": 500, "lbl": 100000, "gcm": 32}
DEFAULT_NLYR = {"amars_lw": 80, "amars_sw": 40, "lbl": 40, "gcm": 60}
GCM_SEED_COLS = 2048  # the full-size GCM batch tiles a seeded block of this many columns (18 GB of inputs)
TRAFFIC = os.path.join(ROOT, "profiles", "r02", "traffic.json")
HASHED = ["pyharp_b200/csrc/disort_solver_body.inc", "pyharp_b200/csrc/disort_layer_ops.inc",
          "pyharp_b200/csrc/disort_adding_step.inc", "pyharp_b200/csrc/disort_device.cuh",
          "pyharp_b200/csrc/flux_inst.cu", "pyharp_b200/csrc/adding_team.cuh", "pyharp_b200/workloads.py"]


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="amars_lw", choices=sorted(DEFAULT_COLS))
    ap.add_argument("--shard", default="col", choices=["col", "wave"])
    ap.add_argument("--ncol", type=int, default=0, help="columns per GPU")
    ap.add_argument("--nwave", type=int, default=0)
    ap.add_argument("--nlyr", type=int, default=0)
    ap.add_argument("--cpu-sample", type=int, default=0, help="solves in the cpu_baseline sample")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-others", action="store_true", help="skip other_workloads / opacity")
    ap.add_argument("--prop-level", action="store_true", help="amars_lw through the solver-level entry (round-1 step)")
    ap.add_argument("--no-numa", action="store_true", help="do not bind the process to the GPU's NUMA node")
    return ap.parse_args()


def shape_of(args, name=None):
    name = name or args.workload
    return (args.nwave or DEFAULT_NWAVE[name], args.ncol or DEFAULT_COLS[name], args.nlyr or DEFAULT_NLYR[name])


def kernel_hash():
    h = hashlib.sha256()
    for f in HASHED:
        h.update(open(os.path.join(ROOT, f), "rb").read())
    return h.hexdigest()[:16]


def traffic_entry(name):
    """ncu counters per solve of the solve kernel for this workload -- only if the committed capture was taken
    with exactly the current kernel sources and workload generator."""
    try:
        tr = json.load(open(TRAFFIC))
    except Exception:  # noqa: BLE001
        return None, "no capture"
    if tr.get("hash") != kernel_hash():
        return None, "stale"
    return tr.get(name), None


# ---------------------------------------------------------------------------------------------
# CPU leg: the oracle (a restatement of cdisort; the real cdisort is un-vendored)
# ---------------------------------------------------------------------------------------------
def cpu_sample(args, wl_name, nsolve_target):
    from pyharp_b200 import workloads as W
    nwave, _, nlyr = shape_of(args, wl_name)
    nw = min(nwave, 32)
    ncol = max(1, nsolve_target // nw)
    return W.make(wl_name, nwave=nw, ncol=ncol, nlyr=nlyr, variant=7)


def run_oracle(wl, nthreads):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    t0 = time.perf_counter()
    _, _, st = O.forward(wl.prop, wl.bc, wl.temf, nstr=wl.nstr, flags=wl.flags, wave_lower=wl.wave_lower,
                         wave_upper=wl.wave_upper, nthreads=nthreads)
    dt = time.perf_counter() - t0
    assert not st.any()
    return dt


def cpu_baseline(args):
    cores = os.cpu_count() or 1
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    O.lib()
    probe = cpu_sample(args, args.workload, 64 * cores)
    dt = run_oracle(probe, cores)
    rate = probe.nsolve / dt
    n = args.cpu_sample or int(min(max(rate * 10.0, 64 * cores), 4e6))
    wl = cpu_sample(args, args.workload, n)
    dt = run_oracle(wl, cores)
    return {"value": wl.nsolve / dt, "unit": "solves/s", "cores": cores, "kind": "port",
            "sample": f"{wl.nsolve} solves of the same workload shape ({wl.shape[0]} spectral points x {wl.shape[1]} "
                      f"columns x {wl.shape[2]} layers) in {dt:.2f} s, OpenMP over solves; oracle/ restatement of "
                      f"cdisort's DISORT solve (cdisort itself is un-vendored in the reference); the opacity stage of "
                      f"the GPU step is not in the CPU sample"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    O.lib()
    probe = cpu_sample(args, args.workload, 64 * cores)
    rate = probe.nsolve / run_oracle(probe, cores)
    # bound the whole run to about two minutes of CPU wall time
    per_step = max(64 * cores, int(rate * min(10.0, 120.0 / max(1, args.steps + args.warmup))))
    wl = cpu_sample(args, args.workload, per_step)
    for _ in range(args.warmup):
        run_oracle(wl, cores)
    t = 0.0
    for _ in range(args.steps):
        t += run_oracle(wl, cores)
    v = wl.nsolve * args.steps / t
    nwave, ncol, nlyr = shape_of(args)
    full = cpu_sample(args, args.workload, 1)  # only for the config text
    sample = (f"each step = {wl.nsolve} solves ({wl.shape[0]} x {wl.shape[1]} columns x {wl.shape[2]} layers) of the "
              f"workload, OpenMP over solves on {cores} threads")
    out = {"metric": METRIC, "value": v, "unit": "solves/s", "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True,
           "scaling": "strong" if args.shard == "wave" else "weak",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
           "config": {"workload": f"{args.workload}-shape synthetic: {nwave} spectral points x {ncol} columns per GPU "
                                  f"x {nlyr} layers, nstr={full.nstr} (CPU arm: bounded sample per step)",
                      "nstr": full.nstr, "nlyr": nlyr},
           "cpu_baseline": {"value": v, "unit": "solves/s", "cores": cores, "kind": "port", "sample": sample},
           "e2e": {"value": v, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


# ---------------------------------------------------------------------------------------------
# clocks sampler (NVML), running during the timed region
# ---------------------------------------------------------------------------------------------
class Clocks:
    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._th = None
        try:
            import pynvml as N
            N.nvmlInit()
            self.N = N
            self.h = N.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = N.nvmlDeviceGetMaxClockInfo(self.h, N.NVML_CLOCK_SM)
        except Exception as e:  # noqa: BLE001
            self.N = None
            self.err = repr(e)

    def _loop(self):
        N = self.N
        names = {getattr(N, k): k.replace("nvmlClocksThrottleReason", "").replace("nvmlClocksEventReason", "")
                 for k in dir(N) if k.startswith("nvmlClocksThrottleReason") and k not in
                 ("nvmlClocksThrottleReasonAll", "nvmlClocksThrottleReasonNone")}
        while not self._stop.is_set():
            try:
                self.samples.append(N.nvmlDeviceGetClockInfo(self.h, N.NVML_CLOCK_SM))
                r = N.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if bit and (r & bit):
                        self.reasons.add(nm)
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(0.02)

    def start(self):
        if self.N:
            self.samples, self.reasons = [], set()
            self._stop.clear()
            self._th = threading.Thread(target=self._loop, daemon=True)
            self._th.start()

    def stop(self):
        if self._th:
            self._stop.set()
            self._th.join()
            self._th = None
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": getattr(self, "err", "no samples")}
        idle = {"GpuIdle", "ApplicationsClocksSetting"}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(r for r in self.reasons if r not in idle), "samples": len(self.samples)}


def bind_numa(local):
    """Pin this process to the CPUs NVML reports as local to its GPU, BEFORE any pinned host memory is allocated, so
    that the staging memory of every rank sits on the NUMA node of its GPU's PCIe root (round 1: all ranks on node 0
    saturated at ~117 GB/s box-wide)."""
    try:
        import pynvml as N
        N.nvmlInit()
        h = N.nvmlDeviceGetHandleByIndex(local)
        ncpu = os.cpu_count() or 1
        words = N.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1 and 64 * i + b < ncpu}
        allowed = os.sched_getaffinity(0)
        cpus = sorted(cpus & allowed)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return {"cpus": f"{cpus[0]}-{cpus[-1]}" if cpus == list(range(cpus[0], cpus[-1] + 1)) else len(cpus),
                    "n": len(cpus)}
    except Exception as e:  # noqa: BLE001
        return {"error": repr(e)}
    return {"n": 0}


# ---------------------------------------------------------------------------------------------
class Ctx:
    pass


def timed_region(ctx, fn, steps):
    """`steps` calls of fn between barrier + synchronize; CUDA events on the current stream; max over ranks."""
    import torch
    import torch.distributed as dist
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if ctx.world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    if ctx.world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], dtype=torch.float64, device=ctx.dev)
    if ctx.world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return ms, float(t.item())


def wall_region(ctx, fn, steps):
    """Host-timed region for calls that synchronise themselves (host-pointer entries); max over ranks."""
    import torch
    import torch.distributed as dist
    if ctx.world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=ctx.dev)
    if ctx.world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def roofline_of(name, W_, Q_, nsolve, kernel_ms, fp64_peak, hbm_peak, hbm_src):
    """Roofline record of the solve kernel: executed FP64 flops (ncu, committed capture of these sources) against the
    measured DFMA peak; the reference-algorithm count W beside it; DRAM traffic against the HBM copy peak."""
    tr, why = traffic_entry(name)
    sec = kernel_ms * 1e-3
    r = {"bound": "fp64", "unit": "TFLOP/s", "peak": fp64_peak,
         "peak_source": "DFMA-saturating kernel measured in this run (b200disort_fp64_peak)",
         "algorithmic_equiv": {"flop_per_solve_W": W_, "tflops": W_ * nsolve / sec / 1e12,
                               "frac": W_ * nsolve / sec / 1e12 / fp64_peak,
                               "note": "reference-algorithm (banded LU + QR) flop count of SURVEY.md 8d at this rate; the "
                                       "adding-method kernel executes about a third of it, so this can exceed 1"},
         "hbm": {"algorithmic_bytes_per_solve_Q": Q_, "algorithmic_gbs": Q_ * nsolve / sec / 1e9, "peak": hbm_peak,
                 "peak_source": hbm_src}}
    if tr is None:
        r.update({"achieved": None, "frac": None, "traffic": None, "executed": None, "stale": why == "stale",
                  "note": f"no ncu capture of the current kernel sources ({why}); see profiles/r02"})
        return r
    ex = tr["executed_flop_per_solve"] * nsolve / sec / 1e12
    traffic = tr["bytes_per_launch"] / tr["solves_per_launch"] * nsolve
    r.update({"achieved": ex, "frac": ex / fp64_peak, "traffic": traffic, "stale": False,
              "executed": {"flop_per_solve": tr["executed_flop_per_solve"], "fp64_pipe_active_pct_ncu": tr["fp64_pipe_active_pct"],
                           "source": tr.get("capture")},
              "note": "achieved = FP64 flops the kernel executes per solve (ncu DFMA x 2 + DMUL + DADD of the committed "
                      "capture, hash-checked against the current sources) x solves per launch / measured launch duration"})
    r["hbm"].update({"traffic_bytes_per_solve": traffic / nsolve, "dram_gbs": traffic / sec / 1e9,
                     "dram_frac_of_hbm_peak": traffic / sec / 1e9 / hbm_peak,
                     "traffic_over_algorithmic": traffic / nsolve / Q_})
    return r


def peaks():
    p = {}
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    return (p["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in p else (6650.0, "fallback (B200_PROFILING.md)")


# ---------------------------------------------------------------------------------------------
# solver-level step of a named workload: prop in HBM -> Planck pre-pass -> solve + fused band integration
# ---------------------------------------------------------------------------------------------
def solver_step_bench(ctx, name, nwave, ncol, nlyr, steps, warmup, variant=0, tile=1, e2e_steps=0):
    import torch
    from pyharp_b200 import Disort, workloads as W
    dev = ctx.dev
    wl = W.make(name, nwave=nwave, ncol=ncol // tile, nlyr=nlyr, variant=variant)
    to_dev = lambda a: None if a is None else torch.from_numpy(a).to(dev)
    rep = lambda t, dim: t if (t is None or tile == 1) else t.repeat(*[tile if i == dim else 1 for i in range(t.dim())])
    d_prop = rep(to_dev(wl.prop), 1)
    d_bc = {k: rep(to_dev(v), 1) for k, v in wl.bc.items()}
    d_temf = rep(to_dev(wl.temf), 0)
    op = wl.options().ncol(ncol)
    ds = Disort(op)
    weights = torch.full((nwave,), 1.0 / nwave, dtype=torch.float64, device=dev)
    step = lambda: ds.forward_band(d_prop, d_bc, d_temf, weights, check=False)
    for _ in range(max(warmup, 3)):
        bflx = step()
    torch.cuda.synchronize()
    assert int(ds.status().abs().sum()) == 0, f"non-zero solve status in the {name} workload"
    assert bool(torch.isfinite(bflx).all())
    ds.stats()
    ctx.clocks.start()
    ms, ms_max = timed_region(ctx, step, steps)
    clk = ctx.clocks.stop()
    st = ds.stats()
    kernel_ms = st["solve_kernel_ms"] / max(1, st["solves_timed"])
    nsolve = nwave * ncol
    out = {"nsolve": nsolve, "ms": ms, "ms_max": ms_max, "kernel_ms": kernel_ms, "clocks": clk, "wl": wl, "ds": ds,
           # (the handle counts its launches per call: one solve kernel, or -- layer-split form, nstr >= 16 -- three per chunk)
           "launches_per_step": max(2 + (1 if wl.thermal else 0), int(st.get("kernel_launches", 0)) + 1),
           "kernel_name": "disort_flux_kernel" if wl.nstr < 16 else "split_prefix_kernel + split_layer_kernel + team_adding_kernel (per chunk)",
           "workspace_bytes": st["workspace_bytes"], "nprop": wl.shape[3],
           "prop_bytes": d_prop.numel() * 8}
    if e2e_steps:
        h_prop = torch.from_numpy(wl.prop).pin_memory()
        h_bc = {k: torch.from_numpy(v).pin_memory() for k, v in wl.bc.items()}
        h_temf = None if wl.temf is None else torch.from_numpy(wl.temf).pin_memory()
        ds1 = ds if tile == 1 else Disort(wl.options())
        for _ in range(2):
            flx = ds1.forward(h_prop, h_bc, h_temf)
        dt = wall_region(ctx, lambda: ds1.forward(h_prop, h_bc, h_temf), e2e_steps)
        out["e2e"] = {"value": wl.nsolve * ctx.world * e2e_steps / dt, "unit": "solves/s",
                      "h2d_bytes_per_step": int(h_prop.nbytes + sum(v.nbytes for v in h_bc.values()) + (0 if h_temf is None else h_temf.nbytes)),
                      "d2h_bytes_per_step": int(flx.nbytes + wl.nsolve * 4), "steps": e2e_steps,
                      "api": "pyharp_b200.Disort.forward(host tensors) -> b200disort_solve_host (per-spectral-point fluxes out)"}
    return out


def nccl_selfcheck(ctx):
    """Spectral-point sharding on a small line-by-line shaped problem: every rank solves its slice of the spectral
    points with the fused band integration, the C-ABI all-reduce (b200disort_allreduce_band on a raw ncclComm_t)
    finishes the band flux, and the result must equal the single-GPU band flux each rank computes by itself."""
    import torch
    from pyharp_b200 import Disort, shard as S, workloads as W
    nw = 64 * ctx.world
    wl = W.make("lbl", nwave=nw, ncol=1, nlyr=12, variant=3)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(ctx.dev)
    w = torch.linspace(0.5, 1.5, nw, dtype=torch.float64, device=ctx.dev)
    whole = Disort(wl.options()).forward_band(t(wl.prop), {k: t(v) for k, v in wl.bc.items()}, None, w)
    sh = S.Shard("wave", *S._split(nw, ctx.world, ctx.rank), ctx.world, ctx.rank)
    prop, bc, _, ww, _, _ = S.take(sh, wl.prop, wl.bc, None, w)
    ds = Disort(wl.options().nwave(sh.size))
    part = ds.forward_band(t(prop), {k: t(v) for k, v in bc.items()}, None, ww)
    ds.allreduce_band(part, ctx.comm)
    torch.cuda.synchronize()
    err = float((part - whole).abs().max() / whole.abs().max())
    assert err < 1e-12, f"NCCL band-flux all-reduce mismatch: {err}"
    return {"ok": True, "max_rel_err": err, "ranks": ctx.world, "nccl_version": ctx.comm.version,
            "what": "spectral-point sharding + b200disort_allreduce_band == single-GPU band flux"}


class stdout_to_stderr:
    """file descriptor 1 -> 2 for the duration of the block (C libraries write to the descriptor, not to sys.stdout)"""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)
        return False


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference)")
    numa = None if args.no_numa else bind_numa(local)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # rank 0's stdout carries ONE JSON line: NCCL's version banner (printed at the first communicator when the box
        # sets NCCL_DEBUG=VERSION) goes to stderr
        with stdout_to_stderr():
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()

    from pyharp_b200 import fp64_peak_tflops, workloads as W
    ctx = Ctx()
    ctx.rank, ctx.world, ctx.local, ctx.dev = rank, world, local, dev
    ctx.clocks = Clocks(local)
    ctx.comm = None
    nccl_check = None
    if world > 1:
        from pyharp_b200.nccl import Communicator
        ctx.comm = Communicator(local)
        nccl_check = nccl_selfcheck(ctx)
    hbm_peak, hbm_src = peaks()
    steps, warmup = args.steps, max(args.warmup, 3)
    nwave, ncol, nlyr = shape_of(args)
    out = None

    if args.workload == "lbl" and args.shard == "wave":
        out = bench_wave_sharded(ctx, args, nwave, nlyr, steps, warmup)
    elif args.workload == "amars_lw" and not args.prop_level:
        out = bench_band(ctx, args, nwave, ncol, nlyr, steps, warmup, hbm_peak, hbm_src)
    else:
        tile = max(1, ncol // GCM_SEED_COLS) if args.workload == "gcm" else 1
        r = solver_step_bench(ctx, args.workload, nwave, ncol, nlyr, steps, warmup, variant=rank, tile=tile,
                              e2e_steps=0 if args.no_e2e else max(3, min(steps, 10)))
        if rank == 0:
            wl = r["wl"]
            fp64_peak = fp64_peak_tflops(local)
            value = r["nsolve"] * world * steps / (r["ms_max"] * 1e-3)
            out = {"metric": METRIC, "value": value, "unit": "solves/s", "n_gpus": world, "steps": steps, "warmup": warmup,
                   "ms_per_step": r["ms_max"] / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                   "dtype": "f64", "data": "synthetic",
                   "config": {"workload": f"{args.workload}-shape synthetic: {nwave} spectral points x {ncol} columns per GPU x "
                                          f"{nlyr} layers, nstr={wl.nstr}, nprop={r['nprop']}, flags={wl.flag_tokens}"
                                          + (f"; a seeded block of {ncol // tile} columns tiled {tile}x" if tile > 1 else ""),
                              "step": "prop resident in HBM -> Planck pre-pass -> flux solve with fused band integration "
                                      "(b200disort_solve_band)",
                              "solves_per_step_per_gpu": r["nsolve"], "nstr": wl.nstr, "nlyr": nlyr, "nwave": nwave,
                              "ncol_per_gpu": ncol, "sharding": f"columns x{world}, no data-path collective",
                              "l2": "inputs (%.0f MB) and back-sweep workspace exceed the 126 MB L2; no explicit flush" % (r["prop_bytes"] / 1e6)},
                   "clocks": r["clocks"], "gpu_launches": (r["launches_per_step"] + 1) * steps,
                   "kernel": {"name": r["kernel_name"], "avg_ms": r["kernel_ms"], "share_of_step": r["kernel_ms"] / (r["ms"] / steps),
                              "workspace_bytes": r["workspace_bytes"]},
                   "roofline": roofline_of(args.workload, wl.flops_per_solve(), wl.bytes_per_solve(), r["nsolve"], r["kernel_ms"],
                                           fp64_peak, hbm_peak, hbm_src),
                   "e2e": r.get("e2e")}
    if rank == 0 and out is not None:
        if numa is not None:
            out["numa_binding"] = numa
        if nccl_check is not None:
            out["nccl_band_allreduce_check"] = nccl_check
        if not args.no_cpu and world == 1:  # the CPU leg is reported at N = 1 only
            out["cpu_baseline"] = cpu_baseline(args)
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        if ctx.comm is not None:
            ctx.comm.destroy()
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------
# default: amars_lw as a radiation band, one C-ABI call per step
# ---------------------------------------------------------------------------------------------
def bench_band(ctx, args, nwave, ncol, nlyr, steps, warmup, hbm_peak, hbm_src):
    import torch
    from pyharp_b200 import fp64_peak_tflops, workloads as W
    dev, world, rank = ctx.dev, ctx.world, ctx.rank
    bw = W.make_band("amars_lw", nwave=nwave, ncol=ncol, nlyr=nlyr, variant=rank)
    band = W.band_of(bw)
    t = lambda a: torch.from_numpy(a).to(dev)
    d = dict(conc=t(bw.conc), dz=t(bw.dz), bc={k: t(v) for k, v in bw.bc.items()}, op={"pres": t(bw.pres), "temp": t(bw.temp)})
    step = lambda: band.forward_band(d["conc"], d["dz"], d["bc"], d["op"], check=False)
    for _ in range(warmup):
        bflx = band.forward_band(d["conc"], d["dz"], d["bc"], d["op"])  # (check=True: raises on a failed solve)
    assert bool(torch.isfinite(bflx).all())
    band.stats()
    ctx.clocks.start()
    ms, ms_max = timed_region(ctx, step, steps)
    clk = ctx.clocks.stop()
    st = band.stats()
    kernel_ms = st["solve_kernel_ms"] / max(1, st["solves_timed"])
    launches_per_step = st["kernel_launches"]
    value = bw.nsolve * world * steps / (ms_max * 1e-3)

    # ---- optics stage alone (HBM-bound streaming kernel): algorithmic bytes = inputs once + prop once
    import ctypes as C
    prop = band.optics(d["conc"], d["dz"], d["op"])
    io, keep = band._io(d["conc"], d["dz"], None, d["op"], dev)  # raw C-ABI calls: no per-call tensor plumbing
    io.prop = prop.data_ptr()
    oms, _ = timed_region(ctx, lambda: band._lib.b200_band_optical_props(band._h, C.byref(io)), 10)
    oms /= 10
    del keep
    obytes = bw.conc.nbytes + bw.pres.nbytes + bw.temp.nbytes + prop.numel() * 8
    optics = {"kernel": "band_optics_fused_kernel (2 RFM attenuators -> mixed prop)", "avg_ms": oms, "algorithmic_bytes": obytes,
              "gbs": obytes / (oms * 1e-3) / 1e9, "frac_of_hbm_peak": obytes / (oms * 1e-3) / 1e9 / hbm_peak,
              "values_per_s": prop.numel() / (oms * 1e-3),
              "note": "compute/L1-bound, not HBM-bound: per value 2 x (4 table gathers, trilinear blend, exp); "
                      "profiles/r02 has the ncu capture"}
    del prop

    # ---- end to end: host tensors through the same call
    e2e = None
    e2e_raw = None
    if not args.no_e2e:
        pin = lambda a: torch.from_numpy(a).pin_memory()
        h = dict(conc=pin(bw.conc), dz=pin(bw.dz), bc={k: pin(v) for k, v in bw.bc.items()}, op={"pres": pin(bw.pres), "temp": pin(bw.temp)})
        for _ in range(3):
            out_h = band.forward_band(h["conc"], h["dz"], h["bc"], h["op"])
        assert float((out_h - bflx.cpu()).abs().max() / bflx.abs().max()) < 1e-12
        esteps = max(3, min(steps, 20))
        dt = wall_region(ctx, lambda: band.forward_band(h["conc"], h["dz"], h["bc"], h["op"]), esteps)
        e2e = {"value": bw.nsolve * world * esteps / dt, "unit": "solves/s", "h2d_bytes_per_step": int(bw.h2d_bytes()),
               "d2h_bytes_per_step": int(out_h.nbytes), "steps": esteps, "ms_per_step": 1e3 * dt / esteps,
               "api": "pyharp_b200.RadiationBand.forward_band(host tensors) -> b200_band_forward_host: conc / pres / temp / "
                      "per-column boundary fields in, band flux (ncol, nlyr+1, 2) out"}
        del h, out_h
    del d
    torch.cuda.empty_cache()

    out = None
    if rank == 0:
        fp64_peak = fp64_peak_tflops(ctx.local)
        wlp = W.make("amars_lw", nwave=nwave, ncol=1, nlyr=nlyr)  # for the W / Q figures of the shape
        W_, Q_ = wlp.flops_per_solve(), wlp.bytes_per_solve()
        out = {"metric": METRIC, "value": value, "unit": "solves/s", "n_gpus": world, "steps": steps, "warmup": warmup,
               "ms_per_step": ms_max / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "f64", "data": "synthetic",
               "config": {"workload": f"amars_lw-shape synthetic band: {nwave} correlated-k g-points x {ncol} columns per GPU x "
                                      f"{nlyr} layers, nstr={bw.nstr}, 2 RFM attenuators (CO2, H2O) on a synthetic ck table, "
                                      f"flags={bw.flag_tokens}",
                          "step": "conc/pres/temp resident in HBM -> fused optics -> layer2level + Planck pre-pass -> flux solve "
                                  "with fused band integration -> (ncol, nlyr+1, 2): one b200_band_forward call",
                          "solves_per_step_per_gpu": bw.nsolve, "nstr": bw.nstr, "nlyr": nlyr, "nwave": nwave,
                          "ncol_per_gpu": ncol, "sharding": f"columns x{world}, no data-path collective",
                          "l2": "optical properties (%.0f MB per step) and back-sweep workspace exceed the 126 MB L2; no "
                                "explicit flush" % (bw.nsolve * nlyr * 8 / 1e6),
                          "flops_per_solve_W": W_, "bytes_per_solve_Q": Q_},
               "clocks": clk, "gpu_launches": launches_per_step * steps,
               "kernel": {"name": "disort_flux_kernel<NOMOM, BAND>", "avg_ms": kernel_ms, "share_of_step": kernel_ms / (ms / steps),
                          "solves_per_s": bw.nsolve / (kernel_ms * 1e-3), "workspace_bytes": st["workspace_bytes"]},
               "roofline": roofline_of("amars_lw", W_, Q_, bw.nsolve, kernel_ms, fp64_peak, hbm_peak, hbm_src),
               "e2e": e2e, "opacity": optics}
    # ---- the round-1 end-to-end route (per-spectral-point fluxes back to the host), for comparison
    if not args.no_e2e and not args.no_others:
        r = solver_step_bench(ctx, "amars_lw", nwave, ncol, nlyr, max(3, steps // 4), 3, variant=rank, e2e_steps=3)
        if rank == 0:
            out["e2e_disort_forward"] = r["e2e"]
            out["solver_level_step"] = {"value": r["nsolve"] * world * max(3, steps // 4) / (r["ms_max"] * 1e-3), "unit": "solves/s",
                                        "kernel_ms": r["kernel_ms"],
                                        "what": "prop resident in HBM -> Planck pre-pass -> solve + fused band integration (round-1 step)"}
        del r
        torch.cuda.empty_cache()
    # ---- the other BASELINE.json configurations on this box (N = 1 only)
    if not args.no_others and world == 1:
        others = []
        fp64_peak = fp64_peak_tflops(ctx.local)
        for name, osteps in (("amars_sw", 10), ("lbl", 10), ("gcm", 4)):
            nw, nc, nl = DEFAULT_NWAVE[name], DEFAULT_COLS[name], DEFAULT_NLYR[name]
            tile = nc // GCM_SEED_COLS if name == "gcm" else 1
            r = solver_step_bench(ctx, name, nw, nc, nl, osteps, 3, tile=tile)
            wl = r["wl"]
            roof = roofline_of(name, wl.flops_per_solve(), wl.bytes_per_solve(), r["nsolve"], r["kernel_ms"], fp64_peak, hbm_peak, hbm_src)
            others.append({"workload": f"{name}: {nw} spectral points x {nc} columns x {nl} layers, nstr={wl.nstr}, nprop={r['nprop']}"
                                       + (f" (a seeded block of {nc // tile} columns tiled {tile}x: {r['prop_bytes'] / 1e9:.1f} GB of inputs)" if tile > 1 else ""),
                           "value": r["nsolve"] * osteps / (r["ms_max"] * 1e-3), "unit": "solves/s", "steps": osteps,
                           "ms_per_step": r["ms_max"] / osteps, "kernel_ms": r["kernel_ms"],
                           "executed_frac_of_fp64_peak": roof["frac"], "algorithmic_equiv_frac": roof["algorithmic_equiv"]["frac"],
                           "dram_frac_of_hbm_peak": roof["hbm"].get("dram_frac_of_hbm_peak"), "stale": roof.get("stale"),
                           "clocks": r["clocks"]})
            del r, wl
            torch.cuda.empty_cache()
        # BASELINE config 2 once more, as a radiation BAND (aerosol optics on the GPU -> mixing -> solve -> band sum in one call)
        if not args.no_e2e:
            others[0]["band_level"] = sw_band_bench(ctx, 10)
        out["other_workloads"] = others
    return out


# ---------------------------------------------------------------------------------------------
# BASELINE config 2 at the band level: what examples/amars_sw.cpp does by hand (S8 + H2SO4 attenuators, mixing, disort,
# integrate_result) as ONE call, inputs resident in HBM and end to end with host tensors
# ---------------------------------------------------------------------------------------------
def sw_band_bench(ctx, steps):
    import torch
    from pyharp_b200 import sw_band
    nwave, ncol, nlyr = DEFAULT_NWAVE["amars_sw"], DEFAULT_COLS["amars_sw"], DEFAULT_NLYR["amars_sw"]
    band, inp = sw_band.make(nwave, ncol, nlyr)
    t = lambda a: torch.from_numpy(a).to(ctx.dev)
    d = dict(conc=t(inp["conc"]), dz=t(inp["dz"]), bc={k: t(v) for k, v in inp["bc"].items()})
    for _ in range(3):
        bflx = band.forward_band(d["conc"], d["dz"], d["bc"], {})
    assert bool(torch.isfinite(bflx).all())
    ms, ms_max = timed_region(ctx, lambda: band.forward_band(d["conc"], d["dz"], d["bc"], {}, check=False), steps)
    pin = lambda a: torch.from_numpy(a).pin_memory()
    h = dict(conc=pin(inp["conc"]), dz=pin(inp["dz"]), bc={k: pin(v) for k, v in inp["bc"].items()})
    for _ in range(2):
        out_h = band.forward_band(h["conc"], h["dz"], h["bc"], {})
    assert float((out_h - bflx.cpu()).abs().max() / bflx.abs().max()) < 1e-12
    dt = wall_region(ctx, lambda: band.forward_band(h["conc"], h["dz"], h["bc"], {}), steps)
    nsolve = nwave * ncol
    return {"workload": f"amars_sw as a band: {nwave} wavenumbers x {ncol} columns x {nlyr} layers, S8 + H2SO4 aerosol attenuators, nstr=8, beam",
            "note": "a synthetic haze on the reference's own S8 / H2SO4 tables (pyharp_b200/sw_band.py): strongly absorbing wavenumbers reach "
                    "the absorption cut-off part of the way down, so a solve is cheaper here than in the solver-level amars_sw line above -- "
                    "compare e2e with value, not with that line",
            "value": nsolve * steps / (ms_max * 1e-3), "unit": "solves/s", "ms_per_step": ms_max / steps,
            "e2e": {"value": nsolve * steps / dt, "unit": "solves/s", "ms_per_step": 1e3 * dt / steps,
                    "h2d_bytes_per_step": int(inp["conc"].nbytes + inp["dz"].nbytes + sum(v.nbytes for v in inp["bc"].values())),
                    "d2h_bytes_per_step": int(out_h.nbytes),
                    "api": "pyharp_b200.RadiationBand.forward_band(host tensors) -> b200_band_forward_host: conc and (nwave, ncol) "
                           "boundary fields in, band flux (ncol, nlyr+1, 2) out"}}


# ---------------------------------------------------------------------------------------------
# BASELINE config 4: 10^5 wavenumbers x 1 column, spectral points sharded over the ranks (strong scaling)
# ---------------------------------------------------------------------------------------------
def bench_wave_sharded(ctx, args, nwave, nlyr, steps, warmup):
    import torch
    from pyharp_b200 import Disort, shard as S, workloads as W
    dev, world, rank = ctx.dev, ctx.world, ctx.rank
    wl = W.make("lbl", nwave=nwave, ncol=1, nlyr=nlyr)
    sh = S.Shard("wave", *S._split(nwave, world, rank), world, rank)  # spectral points, also at world = 1
    w_all = np.full(nwave, 0.1)  # d(nu) of the 0.1 cm^-1 grid
    prop, bc, _, w, _, _ = S.take(sh, wl.prop, wl.bc, None, w_all)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    d_prop, d_bc, d_w = t(prop), {k: t(v) for k, v in bc.items()}, t(w)
    ds = Disort(wl.options().nwave(sh.size))

    def step():
        b = ds.forward_band(d_prop, d_bc, None, d_w, check=False)
        if ctx.comm is not None:
            ds.allreduce_band(b, ctx.comm)
        return b

    for _ in range(warmup):
        bflx = step()
    torch.cuda.synchronize()
    assert int(ds.status().abs().sum()) == 0 and bool(torch.isfinite(bflx).all())
    ds.stats()
    ctx.clocks.start()
    ms, ms_max = timed_region(ctx, step, steps)
    clk = ctx.clocks.stop()
    st = ds.stats()
    kernel_ms = st["solve_kernel_ms"] / max(1, st["solves_timed"])
    ar_ms = None
    if ctx.comm is not None:
        buf = bflx.clone()
        _, ar = timed_region(ctx, lambda: ds.allreduce_band(buf, ctx.comm), 50)
        ar_ms = ar / 50
    # per-call latency with a synchronise after every call (what a host loop that consumes the result sees)
    def sync_step():
        step()
        torch.cuda.synchronize()
    lat = wall_region(ctx, sync_step, steps) / steps
    if rank != 0:
        return None
    return {"metric": "DISORT solves/sec (line-by-line, nstr=4, 40 layers, spectral points sharded)", "value": nwave * steps / (ms_max * 1e-3),
            "unit": "solves/s", "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms_max / steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"lbl-shape synthetic: {nwave} wavenumbers x 1 column x {nlyr} layers, nstr={wl.nstr}, nprop=2, "
                                   f"spectral points sharded over {world} GPU(s): {sh.size} per GPU",
                       "step": "prop slice resident in HBM -> flux solve with fused band integration (warp-reduced L2 "
                               "reductions) -> one ncclAllReduce(sum, f64) of the (1, nlyr+1, 2) band flux",
                       "sharding": f"spectral points x{world}", "solves_per_step_per_gpu": sh.size, "nstr": wl.nstr, "nlyr": nlyr},
            "clocks": clk, "gpu_launches": 2 * steps,
            "kernel": {"name": "disort_flux_kernel<NOMOM, BAND=warp-reduced>", "avg_ms": kernel_ms, "share_of_step": kernel_ms / (ms / steps)},
            "latency_ms_per_call_synchronised": 1e3 * lat, "allreduce_ms": ar_ms,
            "resident_waves": sh.size / (148 * 4 * 128.0),
            "e2e": None}


if __name__ == "__main__":
    main()
