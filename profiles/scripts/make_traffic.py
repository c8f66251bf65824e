#!/usr/bin/env python
"""Writes profiles/r02/traffic.json from the ncu --set full captures of the solve kernel (one .ncu-rep per workload):
DRAM bytes and executed FP64 flops per launch / per solve, FP64-pipe activity -- stamped with the hash of the kernel
sources and the workload generator (bench.py: kernel_hash), so that bench.py only uses counters that belong to the
code it is timing.   python profiles/scripts/make_traffic.py <tag> name=report.ncu-rep:solves ..."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


KERNELS = ("disort_flux_kernel", "split_prefix_kernel", "split_layer_kernel", "team_adding_kernel", "split_adding_kernel")


def counters(rep):
    """the launches of every solve kernel in the capture: one launch of the fused kernel, or the launches of all chunks
    of one step of the layer-split form (their counters are summed)"""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h, u = rows[0], rows[1]
    found = {}
    for v in rows[2:]:
        m = {n: (v[i], u[i]) for i, n in enumerate(h)}
        for k in KERNELS:
            if k in m["Kernel Name"][0]:
                found.setdefault(k, []).append(m)
    if not found:
        raise SystemExit(f"{rep}: no solve kernel launch")
    return found


def num(m, k):
    return float(m[k][0].replace(",", ""))


def main():
    tag = sys.argv[1]
    out = {"comment": "per-launch counters of the solve kernel (nstr = 16: of the three kernels of one chunk of the layer-split form, summed) from the ncu --set full captures of round 2 (profiles/r02/ncu_full_*_"
                      + tag + "_summary.txt): DRAM bytes = dram__bytes_read.sum + dram__bytes_write.sum; executed flop = 2 DFMA + DMUL "
                      "+ DADD thread instructions; FP64 pipe = sm__pipe_fp64_cycles_active.  `hash` = sha256 of the kernel sources and "
                      "the workload generator (bench.py: HASHED); bench.py ignores the file when it does not match",
           "hash": bench.kernel_hash(), "hashed_files": bench.HASHED}
    unit = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    for arg in sys.argv[2:]:
        name, rest = arg.split("=")
        rep, nsolve = rest.rsplit(":", 1)
        nsolve = float(nsolve)
        flop = byt = dur = pipe = 0.0
        parts = {}
        for k, ms in counters(rep).items():
            f = b = t = pk = 0.0
            for m in ms:
                if "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum" in m:  # (--metrics capture: plain counts)
                    f += sum(w * num(m, f"smsp__sass_thread_inst_executed_op_{op}_pred_on.sum") for op, w in (("dfma", 2), ("dmul", 1), ("dadd", 1)))
                else:
                    cyc = num(m, "sm__cycles_elapsed.avg")
                    f += sum(w * num(m, f"smsp__sass_thread_inst_executed_op_{op}_pred_on.sum.per_cycle_elapsed") * cyc
                             for op, w in (("dfma", 2), ("dmul", 1), ("dadd", 1)))
                b += (num(m, "dram__bytes_read.sum") * unit[m["dram__bytes_read.sum"][1]]
                      + num(m, "dram__bytes_write.sum") * unit[m["dram__bytes_write.sum"][1]])
                ti = num(m, "gpu__time_duration.sum") * {"ns": 1e-6, "us": 1e-3, "ms": 1.0}[m["gpu__time_duration.sum"][1]]
                t += ti
                pk += num(m, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active") * ti
            flop, byt, dur, pipe = flop + f, byt + b, dur + t, pipe + pk
            parts[k] = {"launches": len(ms), "ms_under_ncu": round(t, 3), "executed_flop_per_solve": round(f / nsolve, 1),
                        "bytes_per_solve": round(b / nsolve, 1), "fp64_pipe_active_pct": round(pk / t, 1)}
        out[name] = {"bytes_per_launch": byt, "solves_per_launch": nsolve, "executed_flop_per_solve": round(flop / nsolve, 1),
                     "fp64_pipe_active_pct": round(pipe / dur, 1),  # (time-weighted over the kernels of the solve)
                     "capture": f"profiles/r02/ncu_full_{name}_{tag}_summary.txt"}
        if len(parts) > 1:
            out[name]["kernels"] = parts
    json.dump(out, open(os.path.join(ROOT, "profiles", "r02", "traffic.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
