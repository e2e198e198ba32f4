"""Generalized-pruning benchmark (BASELINE.json configs[2]):
the reference's own `libsbn.gp_instance` Python API on the DS1 subsplit DAG, run
once with the UNMODIFIED reference module (oracle/_ref, Eigen on one host core --
GPEngine is single-threaded by design) and once with the same module built over
libsbn_b200.so (integration/_build: GPEngine replaced by integration/gp_engine.*,
PLVs in HBM, one kernel launch per op program).

The workload: DS1.fasta (27 taxa, 934 site patterns) x the 100 topologies of
DS1.100_topologies.nwk, each rooted as (A,B,C) -> (A,(B,C):1):0.  Timed:
  sweep         estimate_sbn_parameters = PopulatePLVs + ComputeLikelihoods +
                UpdateSBNProbabilities (gp_instance.cpp:177-183; the PLV sweep the
                Python API exposes), mean of --repeats
  optimize      estimate_branch_lengths(tol=1e-4, max_iter=--iterations)
Both arms print the log marginal likelihood they end at, so the numbers can be
compared for parity as well as speed.

    python tools/gp_bench.py            # both arms, one JSON line each + a summary
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RUN_DIR = os.path.join(ROOT, "oracle", "_ref")  # holds data/ as staged by `make -C oracle ref`
ARMS = {"reference": os.path.join(ROOT, "oracle", "_ref"), "ours": os.path.join(ROOT, "integration", "_build")}


def _split_top_level(text):
    parts, depth, start = [], 0, 0
    for i, ch in enumerate(text):
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        elif ch == "," and depth == 0:
            parts.append(text[start:i])
            start = i + 1
    parts.append(text[start:])
    return parts


def root_newick(line):
    """(A,B,C); -> (A,(B,C):1):0;   (documented rooting of SURVEY.md 8d config 3)"""
    body = line.strip().rstrip(";")
    close = body.rindex(")")
    children = _split_top_level(body[1:close])
    if len(children) != 3:
        raise ValueError("expected a trifurcating root")
    return f"({children[0]},({children[1]},{children[2]}):1):0;"


_CHILD = r"""
import json, sys, time
sys.path.insert(0, sys.argv[1])
import libsbn
newick, repeats, iterations = sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
inst = libsbn.gp_instance("_ignore/gp_bench_mmap.data")
inst.read_fasta_file("data/DS1.fasta")
inst.read_newick_file(newick)
t0 = time.perf_counter()
inst.make_engine()
make_engine_s = time.perf_counter() - t0
inst.estimate_sbn_parameters()                            # warm-up
t0 = time.perf_counter()
for _ in range(repeats):
    inst.estimate_sbn_parameters()
populate_s = (time.perf_counter() - t0) / repeats
inst.estimate_branch_lengths(1e-4, 1, True)               # warm-up (one iteration)
t0 = time.perf_counter()
inst.estimate_branch_lengths(1e-4, iterations, True)
optimize_s = time.perf_counter() - t0
import numpy as np
import tempfile, os, csv
path = os.path.join(tempfile.mkdtemp(), "bl.csv")
inst.branch_lengths_to_csv(path)
lengths = [float(row[1]) for row in csv.reader(open(path))]
print(json.dumps({"make_engine_s": make_engine_s, "plv_sweep_s": populate_s,
                  "estimate_branch_lengths_s": optimize_s, "iterations_cap": iterations,
                  "gpcsp_count": len(lengths), "branch_length_sum": float(np.sum(lengths)),
                  "branch_length_head": lengths[:4]}))
"""


def run_arm(arm, newick, repeats, iterations):
    module_dir = ARMS[arm]
    if not any(f.startswith("libsbn") and f.endswith(".so") for f in os.listdir(module_dir)):
        return {"arm": arm, "unavailable": f"no libsbn module under {module_dir}"}
    os.makedirs(os.path.join(RUN_DIR, "_ignore"), exist_ok=True)
    done = subprocess.run([sys.executable, "-c", _CHILD, module_dir, newick, str(repeats), str(iterations)],
                          cwd=RUN_DIR, capture_output=True, text=True, timeout=3000)
    if done.returncode != 0:
        return {"arm": arm, "failed": done.stderr[-1500:]}
    result = json.loads(done.stdout.strip().split("\n")[-1])
    result["arm"] = arm
    return result


def measure(repeats=5, iterations=5, arms=("reference", "ours")):
    source = os.path.join(RUN_DIR, "data", "DS1.100_topologies.nwk")
    if not os.path.exists(source):
        raise RuntimeError(f"{source} missing (staged by make -C oracle ref)")
    with tempfile.TemporaryDirectory() as scratch:
        newick = os.path.join(scratch, "ds1_100_rooted.nwk")
        with open(source) as handle, open(newick, "w") as out:
            for line in handle:
                if line.strip():
                    out.write(root_newick(line) + "\n")
        results = {arm: run_arm(arm, newick, repeats, iterations) for arm in arms}
    if all("plv_sweep_s" in results.get(arm, {}) for arm in ("reference", "ours")):
        ref, ours = results["reference"], results["ours"]
        results["summary"] = {
            "workload": "GP on the DS1 subsplit DAG (27 taxa, 934 patterns, 100 rooted topologies)",
            "plv_sweep_speedup": ref["plv_sweep_s"] / ours["plv_sweep_s"],
            "estimate_branch_lengths_speedup": ref["estimate_branch_lengths_s"] / ours["estimate_branch_lengths_s"],
            "branch_length_sum_rel_diff": abs(ref["branch_length_sum"] - ours["branch_length_sum"]) /
            abs(ref["branch_length_sum"]),
            "cpu_cores_used_by_reference": 1}
    return results


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--repeats", type=int, default=5)
    parser.add_argument("--iterations", type=int, default=5)
    parser.add_argument("--arms", default="reference,ours")
    args = parser.parse_args()
    results = measure(args.repeats, args.iterations, tuple(args.arms.split(",")))
    for value in results.values():
        print(json.dumps(value))


if __name__ == "__main__":
    main()
