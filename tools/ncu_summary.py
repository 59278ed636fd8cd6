"""Summarise an `ncu --set full` report (read here, on the CPU box) into a markdown block + a small JSON for bench.py:

    python tools/ncu_summary.py gpurun_out/r02_prof_persistent.ncu-rep profiles/r02_ncu_persistent  "<title>"
"""
import csv, io, json, subprocess, sys
rep, out, title = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg",
        "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_wait", "smsp__pcsamp_warps_issue_stalled_barrier",
        "smsp__pcsamp_warps_issue_stalled_membar", "smsp__pcsamp_warps_issue_stalled_short_scoreboard", "smsp__pcsamp_warps_issue_stalled_selected",
        "smsp__pcsamp_warps_issue_stalled_not_selected", "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle"]
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
md = ["# %s" % title, "# source: %s (`ncu --set full --clock-control none --import-source on`, after the same command exited 0 without ncu)" % rep, ""]
js = {}
for r in rows[2:]:
    d = dict(zip(hdr, r))
    md.append("## %s" % d.get("Kernel Name", "?")[:160])
    for k in WANT:
        if k in d:
            u = units[hdr.index(k)]
            md.append("- %s = %s %s" % (k, d[k], u))
            try:
                v = float(d[k].replace(",", ""))
                js[k] = v * SCALE.get(u, 1.0) if k.startswith("dram__bytes") else v
            except ValueError:
                pass
    md.append("")
open(out + "_summary.md", "w").write("\n".join(md))
json.dump(js, open(out + ".json", "w"), indent=1)
print("\n".join(md))
