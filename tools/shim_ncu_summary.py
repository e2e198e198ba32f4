"""Merges three `ncu --set full` captures of the BEAGLE-compatible library's kernels (one launch each, taken
under tools/beagle_shim_bench.py) into profiles/r02_beagle_shim_ncu_summary.json under "kernels_<order>".

    python tools/shim_ncu_summary.py ORDER POST.ncu-rep PRE.ncu-rep DERIVATIVES.ncu-rep
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATH = os.path.join(ROOT, "profiles", "r02_beagle_shim_ncu_summary.json")
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ncu_kernel_summary import kernel  # noqa: E402


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
