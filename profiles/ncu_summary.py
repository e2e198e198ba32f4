"""Summarises one kernel of an .ncu-rep (ncu --set full) as JSON: the metrics
DESIGN.md and bench.py quote.  Usage: python profiles/ncu_summary.py REP OUT.json TREES [note]"""
import csv, json, subprocess, sys

rep, out_path, trees = sys.argv[1], sys.argv[2], int(sys.argv[3])
note = sys.argv[4] if len(sys.argv) > 4 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
f = lambda k: float(d[k][1].replace(",", ""))
keep = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "sm__cycles_elapsed.avg.per_second"]
summary = {"kernel": d["Kernel Name"][1], "report": rep, "note": note,
           "metrics": {k: {"unit": d[k][0], "value": d[k][1]} for k in keep if k in d}}
stalls = {h.split("stalled_")[1]: float(v.replace(",", "")) for h, (u, v) in d.items()
          if "smsp__pcsamp_warps_issue_stalled" in h and "not_issued" not in h}
total = sum(stalls.values())
summary["stall_share_pct"] = {k: round(100 * v / total, 2) for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])
                              if v / total > 0.005}
dram = (f("dram__bytes_read.sum") * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[d["dram__bytes_read.sum"][0]] +
        f("dram__bytes_write.sum") * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[d["dram__bytes_write.sum"][0]])
summary["trees_in_capture"] = trees
summary["git_sha"] = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
summary["dram_bytes_per_tree"] = dram / trees
summary["dram_bytes_per_launch_at_bench_size"] = dram / trees * 1024
summary["algorithmic_bytes_per_tree"] = (10 * 100 - 14) * 32 * 4 * 100000
json.dump(summary, open(out_path, "w"), indent=1)
m = summary["metrics"]
for k in keep:
    if k in m:
        print(f"{k:85s} {m[k]['value']:>16s} {m[k]['unit']}")
print(summary["stall_share_pct"])
print("dram bytes/tree", summary["dram_bytes_per_tree"])
