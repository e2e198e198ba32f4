"""One kernel of an `ncu --set full` capture as a dictionary (duration, DRAM bytes, registers, occupancy, pipe
utilisation, stall shares), and a command line that collects several captures into one JSON file:

    python tools/ncu_kernel_summary.py OUT.json "what was measured" label=REPORT.ncu-rep[:launch] [label=REPORT.ncu-rep[:launch] ...]
"""
import csv
import json
import subprocess
import sys

UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
TIME = {"us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1}


def kernel(report, index=0):
    raw = subprocess.run(["ncu", "-i", report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    d = {h: (u, v) for h, u, v in zip(rows[0], rows[1], rows[2 + index])}  # (row 2 + i = the i-th captured launch)
    f = lambda k: float(d[k][1].replace(",", ""))  # noqa: E731
    dram = f("dram__bytes_read.sum") * UNIT[d["dram__bytes_read.sum"][0]] + \
        f("dram__bytes_write.sum") * UNIT[d["dram__bytes_write.sum"][0]]
    seconds = f("gpu__time_duration.sum") * TIME[d["gpu__time_duration.sum"][0]]
    stalls = {h.split("stalled_")[1]: float(v.replace(",", "")) for h, (u, v) in d.items()
              if "smsp__pcsamp_warps_issue_stalled" in h and "not_issued" not in h}
    total = sum(stalls.values())
    return {"kernel": d["Kernel Name"][1].replace("<unnamed>::", ""), "duration_ms_under_ncu": seconds * 1e3,
            "dram_bytes": dram, "dram_GBps_under_ncu": dram / seconds / 1e9,
            "registers": int(f("launch__registers_per_thread")), "grid": int(f("launch__grid_size")),
            "block": int(f("launch__block_size")),
            "warps_active_pct": f("sm__warps_active.avg.pct_of_peak_sustained_active"),
            "issue_active_pct": f("smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "fp64_pipe_pct": f("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
            "l1tex_throughput_pct": f("l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
            "warp_instructions": f("smsp__inst_executed.sum"),
            "stall_share_pct": {k: round(100 * v / total, 1) for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])
                                if v / total > 0.01}}


def main():
    out_path, what = sys.argv[1], sys.argv[2]
    summary = {"what": what, "git_sha": subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True,
                                                       text=True).stdout.strip(), "kernels": {}}
    for argument in sys.argv[3:]:
        label, report = argument.split("=", 1)
        index = 0
        if ":" in report:  # REPORT.ncu-rep:i = the i-th launch of a capture that holds several
            report, index = report.rsplit(":", 1)
        summary["kernels"][label] = kernel(report, int(index))
        k = summary["kernels"][label]
        print(label, f"{k['duration_ms_under_ncu']:.3f} ms", f"{k['dram_bytes'] / 1e9:.2f} GB", f"{k['dram_GBps_under_ncu']:.0f} GB/s")
    json.dump(summary, open(out_path, "w"), indent=1)


if __name__ == "__main__":
    main()
