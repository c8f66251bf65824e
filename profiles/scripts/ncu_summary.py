"""Reads an .ncu-rep (ncu -i ... --page raw --csv) and prints the handful of counters the roofline discussion
uses.  Usage: python profiles/scripts/ncu_summary.py gpurun_out/prof.ncu-rep <solves per launch>"""
import csv
import subprocess
import sys

rep, nsolve = sys.argv[1], float(sys.argv[2])
kfilter = sys.argv[3] if len(sys.argv) > 3 else "disort_flux_kernel"  # substring of the kernel name to report
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, u = rows[0], rows[1]
v = next((r for r in rows[2:] if kfilter in r[h.index("Kernel Name")]), rows[2])
m = {n: (v[i], u[i]) for i, n in enumerate(h)}
f = lambda k: float(m[k][0].replace(",", ""))
dur_ms = f("gpu__time_duration.sum") * (1e-6 if m["gpu__time_duration.sum"][1] == "ns" else 1e-3 if m["gpu__time_duration.sum"][1] == "us" else 1.0)
cyc = f("sm__cycles_elapsed.avg")
unit = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
rd = f("dram__bytes_read.sum") * unit[m["dram__bytes_read.sum"][1]]
wr = f("dram__bytes_write.sum") * unit[m["dram__bytes_write.sum"][1]]
dfma = f("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed") * cyc
dmul = f("smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed") * cyc
dadd = f("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed") * cyc
print(f"kernel               {m['Kernel Name'][0]}")
print(f"duration             {dur_ms:.3f} ms for {nsolve:.0f} solves")
print(f"registers/thread     {m['launch__registers_per_thread'][0]}   block {m['launch__block_size'][0]}  grid {m['launch__grid_size'][0]}")
print(f"warps active         {f('sm__warps_active.avg.pct_of_peak_sustained_active'):.1f} % of max")
print(f"FP64 pipe active     {f('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active'):.1f} % (sm__pipe_fp64_cycles_active)")
print(f"executed FP64/solve  DFMA {dfma / nsolve:.0f}  DMUL {dmul / nsolve:.0f}  DADD {dadd / nsolve:.0f}  => {(2 * dfma + dmul + dadd) / nsolve:.0f} flop/solve")
print(f"executed FP64 rate   {(2 * dfma + dmul + dadd) / (dur_ms * 1e-3) / 1e12:.2f} TFLOP/s")
print(f"DRAM traffic         read {rd / 1e9:.2f} GB + write {wr / 1e9:.2f} GB = {(rd + wr) / 1e9:.2f} GB/launch ({(rd + wr) / (dur_ms * 1e-3) / 1e12:.2f} TB/s), {(rd + wr) / nsolve:.0f} B/solve")
print(f"I-cache hit rate     {f('sm__icc_request_hit_rate.pct'):.1f} %")
print(f"local ld/st instr    {f('sass__inst_executed_local_loads'):.3g} / {f('sass__inst_executed_local_stores'):.3g}  (L1 hit {f('l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct'):.0f} %)")
st = {k.split("issue_stalled_")[1].split("_per_issue")[0]: f(k) for k in m if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")}
top = sorted(st.items(), key=lambda kv: -kv[1])[:6]
print("stalls per issue     " + ", ".join(f"{k} {x:.2f}" for k, x in top))
