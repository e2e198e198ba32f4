"""Partial-update HBM GB/s of the BEAGLE-compatible device library (libhmsbeagle_b200.so:
libsbn_b200/csrc/beagle_shim.cu) -- SURVEY.md 8d's second metric, on the materialised
op-at-a-time schedule where it is directly meaningful: one FatBeagle::BranchGradientInternals
call sequence (fat_beagle.cpp:119-175) on a random tree at BASELINE configs[3] size (100 taxa x
100k patterns x 4 categories, GTR), device time by CUDA events per BEAGLE call, bytes by
SURVEY.md 8d's table (U = 32 C P per partial; compact tips and matrices ignored).

    python tools/beagle_shim_bench.py [--taxa 100 --patterns 100000 --categories 4 --repeats 5]
"""
import argparse
import ctypes
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from libsbn_b200 import beagle as harness  # noqa: E402  (the BEAGLE call sequence of fat_beagle.cpp over ctypes)


def measure(n=100, P=100000, C=4, repeats=5, device=None):
    """The op lists in libsbn's own order (depth-first: the reference's call sequence), and as
    `node_id_order` the same measurement with the internal nodes numbered by random joins and the lists in
    id order (consecutive ops rarely parent and child: no operand stays in registers)."""
    if device is not None:
        os.environ["SBNB_BEAGLE_DEVICE"] = str(device)
    out = _measure(n, P, C, repeats, "libsbn")
    other = _measure(n, P, C, repeats, "node_id")
    out["node_id_order"] = {key: other[key] for key in ("update_partials", "update_pre_partials", "edge_derivatives",
                                                        "device_ms_per_logl_plus_gradient")}
    return out


def measure_dram(n, P, C, order):
    """DRAM bytes of the three kernels of ONE call sequence, by ncu on a probe process (two metrics, one pass;
    whatever the probe process prints is discarded).  ({call: bytes}, source) or (None, None)."""
    import subprocess
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        log = os.path.join(tmp, "dram.csv")
        command = ["ncu", "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum", "--clock-control", "none",
                   "--print-units", "base", "--kernel-name-base", "demangled", "-k",
                   "regex:UpdatePartialsPipelinedKernel|EdgeDerivativesKernel", "-c", "3", "--csv", "--log-file", log,
                   sys.executable, os.path.abspath(__file__), "--_probe", order, "--taxa", str(n), "--patterns", str(P),
                   "--categories", str(C)]
        try:
            subprocess.run(command, capture_output=True, text=True, timeout=600, cwd=ROOT)
            totals = {}
            for row in open(log):
                cells = [c.strip('"') for c in row.strip().split('","')]
                if len(cells) > 3 and cells[-3].startswith("dram__bytes_"):
                    name = next(c for c in cells if "Kernel" in c and "(" in c)
                    pre_order = "PipelinedKernel<1," in name or "PipelinedKernel<(bool)1," in name
                    key = "edge_derivatives" if "EdgeDerivatives" in name else (
                        "update_pre_partials" if pre_order else "update_partials")
                    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[cells[-2]]
                    totals[key] = totals.get(key, 0.0) + float(cells[-1].replace(",", "")) * scale
            if len(totals) == 3 and all(v > 0 for v in totals.values()):
                return totals, ("ncu dram__bytes_read.sum + dram__bytes_write.sum of the three launches of one call "
                                "sequence, measured in this run")
        except (OSError, subprocess.SubprocessError, ValueError, KeyError, StopIteration):
            pass
    return None, None


def _measure(n, P, C, repeats, order, probe=True):
    rng = np.random.default_rng(20261017)
    states = rng.integers(0, 4, size=(n, P)).astype(np.int32)
    states[rng.random(states.shape) < 0.01] = 4
    post, pre = harness.random_tree_operations(n, rng, True, order)
    lengths = np.maximum(rng.exponential(0.1, size=2 * n - 1), 1e-6)
    evec, ivec, evals, freqs, q = harness.gtr_eigensystem()
    rates = np.array([0.03, 0.25, 0.8, 2.92])[:C] if C == 4 else np.ones(C)
    rates = rates / rates.mean()
    beagle = harness.Beagle(None, n, P, C, True)
    beagle.lib.sbnbBeagleLastKernelMs.restype = ctypes.c_double
    beagle.set_tips(states, np.ones(P), True)
    beagle.set_model(evec, ivec, evals, freqs, rates, np.full(C, 1.0 / C))
    U = 32.0 * C * P
    N = 2 * n - 1

    def timed(call):
        call()
        return beagle.lib.sbnbBeagleLastKernelMs(beagle.handle)

    derivative_args = [beagle._i(np.arange(N - 1)), beagle._i(np.arange(N - 1) + N), beagle._i(np.full(N - 1, N - 1)),
                       beagle._i([0])]
    sums = np.zeros(N - 1)

    def derivatives():
        beagle.ok(beagle.lib.beagleCalculateEdgeDerivatives(beagle.handle, *[a[1] for a in derivative_args], N - 1, None,
                                                            sums.ctypes.data_as(harness._D), None))

    keep_a, post_ptr = beagle._i(post)
    keep_b, pre_ptr = beagle._i(pre)
    if repeats < 0:  # (the probe process of measure_dram: one call sequence, three kernels)
        beagle.log_likelihood_and_gradient(post, pre, lengths, q, rates, freqs, True, per_site=False)
        beagle.close()
        return None
    results = {"post_ms": [], "pre_ms": [], "derivatives_ms": [], "call_sequence_ms": []}
    for _ in range(repeats + 1):
        t0 = time.perf_counter()
        logl, sums, _, _ = beagle.log_likelihood_and_gradient(post, pre, lengths, q, rates, freqs, True, per_site=False)
        results["call_sequence_ms"].append((time.perf_counter() - t0) * 1e3)
        results["post_ms"].append(timed(lambda: beagle.ok(beagle.lib.beagleUpdatePartials(beagle.handle, post_ptr, len(post), 0))))
        results["derivatives_ms"].append(timed(derivatives))
        results["pre_ms"].append(timed(lambda: beagle.ok(beagle.lib.beagleUpdatePrePartials(beagle.handle, pre_ptr, len(pre), -1))))
    med = {k: float(np.median(v[1:])) for k, v in results.items()}
    # post-order writes n-1, reads n-2 partials; pre-order writes 2n-2, reads (2n-4) + (n-2)
    post_bytes, pre_bytes = (2 * n - 3) * U, (5 * n - 8) * U
    out = {
        "workload": f"BEAGLE-compatible device library, one tree: {n} taxa x {P} patterns x {C} categories, GTR, rescaling on, "
                    "op lists in libsbn's (depth-first) order",
        "partial_bytes_U": U,
        "update_partials": {"ms": med["post_ms"], "algorithmic_GB": post_bytes / 1e9,
                            "GBps": post_bytes / med["post_ms"] / 1e6},
        "update_pre_partials": {"ms": med["pre_ms"], "algorithmic_GB": pre_bytes / 1e9,
                                "GBps": pre_bytes / med["pre_ms"] / 1e6},
        "edge_derivatives": {"ms": med["derivatives_ms"], "algorithmic_GB": (3 * n - 4) * U / 1e9,
                             "GBps": (3 * n - 4) * U / med["derivatives_ms"] / 1e6},
        "device_ms_per_logl_plus_gradient": med["post_ms"] + med["pre_ms"] + med["derivatives_ms"],
        "branch_gradient_call_sequence_ms": med["call_sequence_ms"],
        "note": "call_sequence = FatBeagle's calls with FatBeagle's arguments (sums of derivatives only), including the "
                "host <-> device copies of the BEAGLE API (a U-sized root pre-order partial built and uploaded as "
                "fat_beagle.cpp:316-325 does) and a stream synchronisation per call",
        "log_likelihood": logl,
    }
    # Two fractions per call.  `algorithmic_frac_of_peak`: SURVEY.md 8d's bytes (every partial through HBM)
    # over the time -- it exceeds 1 where the 126 MB L2 serves partials written a few ops earlier.
    # `hbm_frac`: the DRAM bytes of the same launch, measured by ncu IN THIS RUN on a probe process that makes
    # one call sequence (else the committed capture profiles/r02_beagle_shim_ncu_summary.json, and
    # `dram_source` says so), over the time = real HBM utilisation.
    peaks = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm = json.load(open(peaks)).get("hbm_gbs") if os.path.exists(peaks) else None
    dram, source = measure_dram(n, P, C, order) if probe else (None, None)
    if dram is None:
        summary_path = os.path.join(ROOT, "profiles", "r02_beagle_shim_ncu_summary.json")
        committed = json.load(open(summary_path)).get("kernels_" + order, {}) if os.path.exists(summary_path) else {}
        if (n, P, C) == (100, 100000, 4) and committed:
            dram = {key: committed[key]["dram_bytes"] for key in committed}
            source = "profiles/r02_beagle_shim_ncu_summary.json (committed ncu capture; not measured in this run)"
    if hbm:
        out["hbm_peak_GBps_measured"] = hbm
        for key in ("update_partials", "update_pre_partials", "edge_derivatives"):
            out[key]["algorithmic_frac_of_peak"] = out[key]["GBps"] / hbm
            if dram and key in dram:
                out[key]["dram_GB_measured"] = dram[key] / 1e9
                out[key]["hbm_GBps"] = dram[key] / out[key]["ms"] / 1e6
                out[key]["hbm_frac"] = out[key]["hbm_GBps"] / hbm
        if dram:
            out["dram_source"] = source
    beagle.close()
    return out


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--taxa", type=int, default=100)
    parser.add_argument("--patterns", type=int, default=100000)
    parser.add_argument("--categories", type=int, default=4)
    parser.add_argument("--repeats", type=int, default=5)
    parser.add_argument("--_probe", dest="probe", default=None, help=argparse.SUPPRESS)
    args = parser.parse_args()
    if args.probe:  # the process measure_dram() profiles: one call sequence in the given op order
        _measure(args.taxa, args.patterns, args.categories, -1, args.probe, probe=False)
        return
    print(json.dumps(measure(args.taxa, args.patterns, args.categories, args.repeats)))


if __name__ == "__main__":
    main()
