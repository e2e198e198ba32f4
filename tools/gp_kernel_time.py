"""Device time of the GP interpreter on the DS1 subsplit DAG (tests/golden/gp_ds1_dag.npz:
27 taxa, 934 patterns, 612 PLVs, 181 GPCSPs): per op program, the CUDA-event time of the
one kernel launch (sbnb_gp_last_kernel_ms) and the wall time of the C-ABI call with host
buffers (program upload, launch, status read-back).

    python tools/gp_kernel_time.py [--repeats 20]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from libsbn_b200.gp_engine import GPEngine  # noqa: E402


def measure(repeats=20, device=0):
    fx = dict(np.load(os.path.join(ROOT, "tests", "golden", "gp_ds1_dag.npz")))
    engine = GPEngine(fx["tip_states"], fx["pattern_weights"], int(fx["site_count"]), int(fx["plv_count"]),
                      int(fx["gpcsp_count"]), rescaling_threshold=float(fx["rescaling_threshold"]),
                      sbn_prior=fx["sbn_prior"],
                      unconditional_node_probabilities=fx["unconditional_node_probabilities"],
                      inverted_sbn_prior=fx["inverted_sbn_prior"], device=device)
    engine.set_branch_lengths(fx["initial_branch_lengths"])
    out = {"workload": "GP interpreter, DS1 DAG (934 patterns, 612 PLVs, 181 GPCSPs)", "repeats": repeats}
    for name in ("populate_plvs", "compute_likelihoods", "marginal_likelihood", "branch_length_optimization",
                 "optimize_sbn_parameters"):
        program = fx["program_" + name]
        engine.process_operations(fx["program_populate_plvs"])  # a defined state; warm-up
        kernel, wall = [], []
        for _ in range(repeats):
            if name == "branch_length_optimization":
                engine.set_branch_lengths(fx["initial_branch_lengths"])
                engine.process_operations(fx["program_populate_plvs"])
            t0 = time.perf_counter()
            engine.process_operations(program)
            wall.append((time.perf_counter() - t0) * 1e3)
            kernel.append(engine.last_kernel_ms)
        out[name] = {"words": int(program.size), "kernel_ms": float(np.median(kernel)),
                     "call_ms": float(np.median(wall))}
    out["log_marginal_likelihood"] = float(engine.get_log_marginal_likelihood())
    return out


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--repeats", type=int, default=20)
    args = parser.parse_args()
    print(json.dumps(measure(args.repeats)))


if __name__ == "__main__":
    main()
