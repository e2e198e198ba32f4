"""Generates tests/golden/gp_*.npz by running the UNMODIFIED reference GP code here.

Run in the build container only (needs /root/reference and oracle/_ref/gp_dump,
built by `make -C oracle ref` from oracle/gp_dump.cpp + the reference's own
GPInstance / GPDAG / GPEngine objects):

    python tests/golden/make_gp_fixtures.py

Every fixture is one scenario of the reference's src/gp_doctest.cpp: the
arguments GPInstance::MakeEngine passes to GPEngine, the GPOperation programs
GPDAG schedules (flattened with the record layout of include/sbn_b200_gp.h),
and the reference engine's state after each stage.  The external goldens the
doctest hard-codes are stored next to them.  Nothing here runs on the GPU box.
"""
import json
import os
import re
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
GP_DUMP = os.path.join(ROOT, "oracle", "_ref", "gp_dump")
DATA = "/root/reference/data"


def dump(scenario, fasta, newick, **options):
    args = [GP_DUMP, scenario, fasta if os.path.isabs(fasta) else os.path.join(DATA, fasta),
            newick if os.path.isabs(newick) else os.path.join(DATA, newick)]
    for key, value in options.items():
        if isinstance(value, (list, tuple)):
            value = ",".join(repr(x) for x in value)
        args.append(f"{key}={value}")
    text = subprocess.run(args, check=True, capture_output=True, text=True).stdout
    record = json.loads(text)
    out = {}
    for key, value in record.items():
        if key == "scenario":
            continue
        if key == "tip_states":
            out[key] = np.array(value, dtype=np.uint8).reshape(int(record["taxon_count"]), -1)
        elif key.startswith("program_") or key.endswith("rescaling_counts") or key.startswith("quartet_tips") \
                or key in ("quartet_counts", "gradient_op", "hybrid_requests"):
            out[key] = np.array(value, dtype=np.int32)
        elif key.endswith("_plvs"):
            out[key] = np.array(value, dtype=np.float64).reshape(int(record["plv_count"]), -1, 4)
        elif key.endswith("log_likelihood_matrix"):
            out[key] = np.array(value, dtype=np.float64).reshape(int(record["gpcsp_count"]), -1)
        elif key.endswith("_count") or key in ("estimate_iterations", "estimate_max_iter", "quartet_central_gpcsp"):
            out[key] = np.int64(value)
        else:
            out[key] = np.array(value, dtype=np.float64)
    return out


def save(name, record, **extra):
    record = dict(record, **extra)
    path = os.path.join(HERE, f"gp_{name}.npz")
    np.savez_compressed(path, **record)
    print(f"{path}: {os.path.getsize(path)} bytes, {len(record)} arrays")


def rooted_ds1_newick(path):
    """BASELINE.json configs[2] / SURVEY.md 8d(3): each unrooted DS1 topology
    (A,B,C); is rooted as (A,(B,C):1):0; -- a documented choice, the reference
    ships no rooted DS1 topology set."""
    lines = []
    for line in open(os.path.join(DATA, "DS1.100_topologies.nwk")):
        line = line.strip()
        if not line:
            continue
        body = re.sub(r":[0-9.eE+-]+$", "", line.rstrip(";"))  # drop the root's own ":0"
        assert body[0] == "(" and body[-1] == ")"
        depth, parts, start = 0, [], 1
        for i, ch in enumerate(body[1:-1], start=1):
            if ch == "(":
                depth += 1
            elif ch == ")":
                depth -= 1
            elif ch == "," and depth == 0:
                parts.append(body[start:i])
                start = i + 1
        parts.append(body[start:-1])
        assert len(parts) == 3, len(parts)
        lines.append(f"({parts[0]},({parts[1]},{parts[2]}):1):0;")
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")


def doctest_doubles(anchor):
    """The numbers streamed into an Eigen vector by the `<<` statement that
    starts at `anchor` in src/gp_doctest.cpp (read at generation time instead
    of being retyped)."""
    text = open("/root/reference/src/gp_doctest.cpp").read()
    start = text.index(anchor)
    start = text.index("<<", start) + 2
    end = text.index(";", start)
    return [float(x) for x in re.findall(r"-?\d+\.?\d*(?:[eE][-+]?\d+)?", text[start:end])]


def main():
    # gp_doctest.cpp:33-47, 89-101 -- hello, classical likelihood -84.77961943
    save("hello", dump("hello", "hello.fasta", "hello_rooted.nwk", branch_lengths=[0, 0.22, 0.113, 0.15, 0.1],
                       plvs=1), golden_log_likelihood=-84.77961943)
    # gp_doctest.cpp:218-241 -- gradient on a single nucleotide
    save("hello_gradient", dump("gradient", "hello_single_nucleotide.fasta", "hello_rooted.nwk",
                                branch_lengths=[0, 0.22, 0.113, 0.15, 0.1], plvs=1),
         golden_log_likelihood_and_derivative=np.array([-4.806671945, -0.6109379521]))
    # gp_doctest.cpp:197-216 -- composite marginals after EstimateBranchLengths(1e-4, 100)
    save("hello_two_trees", dump("estimate", "hello.fasta", "hello_rooted_two_trees.nwk", plvs=1))
    save("five_taxon", dump("estimate", "five_taxon.fasta", "five_taxon_rooted.nwk", plvs=1))
    save("ds1_reduced_5", dump("estimate", "ds1-reduced-5.fasta", "ds1-reduced-5.nwk"))
    save("seven_taxon_all_trees", dump("estimate", "7-taxon-slice-of-ds1.fasta",
                                       "simplest-hybrid-marginal-all-trees.nwk"))
    # gp_doctest.cpp:296-309 -- five taxa, EstimateBranchLengths(1e-6, 10)
    save("five_taxon_tight", dump("estimate", "five_taxon.fasta", "five_taxon_rooted.nwk", tol=1e-6, max_iter=10))
    # gp_doctest.cpp:243-253 -- fluA with two rescaling thresholds; equal to 1e-10
    save("flua_threshold_1e-40", dump("flua", "fluA.fa", "fluA.tree", constant=0.01))
    save("flua_threshold_1e-4", dump("flua", "fluA.fa", "fluA.tree", constant=0.01, threshold=1e-4))
    # The reference's 1e-4 never triggers a rescale on fluA (every PLV keeps an entry
    # above it), so two more thresholds exercise RescalePLVIfNeeded (gp_engine.cpp:288-320),
    # PrepForMarginalization and the rescaled IncrementWithWeightedEvolvedPLV for real.
    save("flua_threshold_0.5", dump("flua", "fluA.fa", "fluA.tree", constant=0.01, threshold=0.5))
    save("five_taxon_threshold_0.5", dump("estimate", "five_taxon.fasta", "five_taxon_rooted.nwk", threshold=0.5,
                                          plvs=1))
    # gp_doctest.cpp:522-587 -- quartet hybrid marginals
    save("simplest_hybrid", dump("quartet", "7-taxon-slice-of-ds1.fasta", "simplest-hybrid-marginal.nwk",
                                 branch_lengths=doctest_doubles("branch_lengths << 0.058"),
                                 quartet=[12, 0, 11]))
    save("second_simplest_hybrid", dump("quartet", "7-taxon-slice-of-ds1.fasta",
                                        "second-simplest-hybrid-marginal.nwk",
                                        branch_lengths=doctest_doubles("branch_lengths << 0.09"),
                                        quartet=[12, 1, 11]))
    # BASELINE.json configs[2]: the DS1 subsplit DAG (rooted as documented above)
    with tempfile.TemporaryDirectory() as tmp:
        rooted = os.path.join(tmp, "ds1_rooted.nwk")
        rooted_ds1_newick(rooted)
        save("ds1_dag", dump("estimate", "DS1.fasta", rooted, tol=1e-4, max_iter=2))


if __name__ == "__main__":
    main()
