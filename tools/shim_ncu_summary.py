"""Merges three `ncu --set full` captures of the BEAGLE-compatible library's kernels (one launch each, taken
under tools/beagle_shim_bench.py) into profiles/r02_beagle_shim_ncu_summary.json under "kernels_<order>".

    python tools/shim_ncu_summary.py ORDER POST.ncu-rep PRE.ncu-rep DERIVATIVES.ncu-rep
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATH = os.path.join(ROOT, "profiles", "r02_beagle_shim_ncu_summary.json")
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
TIME = {"us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1}


def kernel(report):
    raw = subprocess.run(["ncu", "-i", report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    d = {h: (u, v) for h, u, v in zip(rows[0], rows[1], rows[2])}
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
    order, post, pre, derivatives = sys.argv[1:5]
    summary = json.load(open(PATH)) if os.path.exists(PATH) else {}
    summary["kernels_" + order] = {"update_partials": kernel(post), "update_pre_partials": kernel(pre),
                                   "edge_derivatives": kernel(derivatives)}
    summary["git_sha_" + order] = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True,
                                                 text=True).stdout.strip()
    json.dump(summary, open(PATH, "w"), indent=1)
    for key, value in summary["kernels_" + order].items():
        print(order, key, f"{value['duration_ms_under_ncu']:.3f} ms", f"{value['dram_bytes'] / 1e9:.2f} GB")


if __name__ == "__main__":
    main()
