"""Summarise an ncu --set full report into a small JSON (the metrics DESIGN.md / bench.py quote).
usage: python tools/ncu_summary.py report.ncu-rep [kernel-name-substring] > profiles/xxx.json"""
import csv, io, json, subprocess, sys

rep = sys.argv[1]
want_kernel = sys.argv[2] if len(sys.argv) > 2 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEEP = ("Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "launch__registers_per_thread", "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg",
        "sm__cycles_elapsed.avg.per_second", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct")
out = []
for r in rows[2:]:
    d = dict(zip(hdr, r))
    if want_kernel and want_kernel not in d.get("Kernel Name", ""):
        continue
    if d.get("sm__cycles_elapsed.avg", "") in ("", "-nan", "nan"):
        continue  # instance that exited at once (device-side variant selection)
    u = dict(zip(hdr, units))
    e = {k: (d[k] + (" " + u[k] if u.get(k) else "")).strip() for k in KEEP if k in d}
    for k in hdr:
        if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and float(d[k] or 0) >= 0.05:
            e[k.replace("smsp__average_warps_issue_stalled_", "stall_").replace("_per_issue_active.ratio", "")] = round(float(d[k]), 3)
    out.append(e)
print(json.dumps(out, indent=1))
