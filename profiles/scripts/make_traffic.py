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


def counters(rep, kernel_regex="disort_flux_kernel"):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h, u = rows[0], rows[1]
    for v in rows[2:]:
        m = {n: (v[i], u[i]) for i, n in enumerate(h)}
        if kernel_regex in m["Kernel Name"][0]:
            return m
    raise SystemExit(f"{rep}: no {kernel_regex} launch")


def num(m, k):
    return float(m[k][0].replace(",", ""))


def main():
    tag = sys.argv[1]
    out = {"comment": "per-launch counters of disort_flux_kernel from the ncu --set full captures of round 2 (profiles/r02/ncu_full_*_"
                      + tag + "_summary.txt): DRAM bytes = dram__bytes_read.sum + dram__bytes_write.sum; executed flop = 2 DFMA + DMUL "
                      "+ DADD thread instructions; FP64 pipe = sm__pipe_fp64_cycles_active.  `hash` = sha256 of the kernel sources and "
                      "the workload generator (bench.py: HASHED); bench.py ignores the file when it does not match",
           "hash": bench.kernel_hash(), "hashed_files": bench.HASHED}
    unit = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    for arg in sys.argv[2:]:
        name, rest = arg.split("=")
        rep, nsolve = rest.rsplit(":", 1)
        nsolve = float(nsolve)
        m = counters(rep)
        cyc = num(m, "sm__cycles_elapsed.avg")
        flop = sum(w * num(m, f"smsp__sass_thread_inst_executed_op_{op}_pred_on.sum.per_cycle_elapsed") * cyc
                   for op, w in (("dfma", 2), ("dmul", 1), ("dadd", 1)))
        rd = num(m, "dram__bytes_read.sum") * unit[m["dram__bytes_read.sum"][1]]
        wr = num(m, "dram__bytes_write.sum") * unit[m["dram__bytes_write.sum"][1]]
        out[name] = {"bytes_per_launch": rd + wr, "solves_per_launch": nsolve, "executed_flop_per_solve": round(flop / nsolve, 1),
                     "fp64_pipe_active_pct": round(num(m, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"), 1),
                     "capture": f"profiles/r02/ncu_full_{name}_{tag}_summary.txt"}
    json.dump(out, open(os.path.join(ROOT, "profiles", "r02", "traffic.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
