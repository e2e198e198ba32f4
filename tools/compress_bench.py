"""Site-pattern compression (sbnb_compress_site_patterns) on synthetic alignments:
device time (CUDA events around the kernel sequence), algorithmic bytes
(read taxa x sites characters, write taxa x patterns symbols + 8 bytes per weight)
over that time against the measured HBM peak, the end-to-end call (host buffers,
H2D/D2H inside), and the UNMODIFIED reference's SitePattern::Compress
(oracle/_ref/site_pattern_dump on a FASTA of a bounded sample, parsing included)
on the box's host cores.  One JSON line per workload."""
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from libsbn_b200.site_pattern import SitePattern  # noqa: E402


def alignment(taxa, sites, distinct, seed):
    rng = np.random.default_rng(seed)
    alphabet = np.frombuffer(b"ACGT-", dtype=np.uint8)
    if distinct is None:  # iid columns (BASELINE configs[3], [4]): 1 % gaps
        states = rng.integers(0, 4, size=(taxa, sites), dtype=np.uint8)
        states[rng.integers(0, 100, size=(taxa, sites), dtype=np.uint8) == 0] = 4
        return np.ascontiguousarray(alphabet[states])
    pool = alphabet[rng.integers(0, 5, size=(taxa, distinct))]
    return np.ascontiguousarray(pool[:, rng.integers(0, distinct, size=sites)])


def main():
    # (development: `--quick` skips the CPU arm and the small workload)
    quick = "--quick" in sys.argv
    only_repeats = "--only-repeats" in sys.argv  # (profiling: just the 50k-pattern workload)
    peaks = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = json.load(open(peaks))["hbm_gbs"] if os.path.exists(peaks) else 6650.0
    for name, taxa, sites, distinct in [("100 taxa x 100k sites, iid columns", 100, 100000, None),
                                        ("1000 taxa x 1M sites, iid columns", 1000, 1000000, None),
                                        ("1000 taxa x 1M sites, 50k distinct columns", 1000, 1000000, 50000)]:
        if (quick and sites < 1000000) or (only_repeats and distinct is None):
            continue
        sequences = alignment(taxa, sites, distinct, seed=5)
        SitePattern(sequences)  # warm-up (context, allocations)
        device_ms, wall = [], []
        for _ in range(3):
            start = time.perf_counter()
            pattern = SitePattern(sequences)
            wall.append(time.perf_counter() - start)
            device_ms.append(pattern.device_ms)
        ms = min(device_ms)
        algorithmic = taxa * sites + taxa * pattern.pattern_count + 8 * pattern.pattern_count
        line = {"workload": name, "patterns": pattern.pattern_count, "device_ms": ms,
                "sites_per_s_device": sites / (ms * 1e-3), "algorithmic_bytes": algorithmic,
                "algorithmic_GBps": algorithmic / (ms * 1e-3) / 1e9,
                "frac_of_hbm_peak": algorithmic / (ms * 1e-3) / 1e9 / peak, "hbm_peak_GBps": peak,
                "e2e_s_host_buffers": min(wall), "sites_per_s_e2e": sites / min(wall)}
        dump = os.path.join(ROOT, "oracle", "_ref", "site_pattern_dump")
        if os.path.exists(dump) and not quick and not only_repeats:
            sample = min(sites, 100000)
            with tempfile.NamedTemporaryFile("w", suffix=".fasta", delete=False) as handle:
                for t, row in enumerate(sequences):
                    handle.write(f">t{t:04d}\n{bytes(row[:sample]).decode()}\n")
            start = time.perf_counter()
            subprocess.run([dump, handle.name], stdout=subprocess.DEVNULL, check=True)
            seconds = time.perf_counter() - start
            os.unlink(handle.name)
            line["cpu_baseline"] = {"kind": "reference", "cores": 1,
                                    "sample": f"first {sample} sites of the same alignment, FASTA parsing and "
                                              "JSON printing included",
                                    "value": sample / seconds, "unit": "sites/s"}
        print(json.dumps(line))


if __name__ == "__main__":
    main()
