"""Times the full GTR phylo_gradients call (logL + branch + site-model + substitution-model
gradients) on the configs[3] workload with the analytic and with the finite-difference
substitution gradient.   python tools/full_gradient_bench.py [--trees 296] [--steps 2]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import libsbn_b200 as sbn  # noqa: E402


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--taxa", type=int, default=100)
    parser.add_argument("--patterns", type=int, default=100000)
    parser.add_argument("--trees", type=int, default=296)
    parser.add_argument("--steps", type=int, default=2)
    args = parser.parse_args()
    states, weights, parent_ids, lengths, params = bench.workload(args)
    engine = sbn.Engine(sbn.PhyloModelSpecification("GTR", "weibull+4", "none"), states, weights, 0)
    batch = sbn.TreeBatch(parent_ids, lengths)
    out = {"trees": args.trees, "library": os.environ.get("SBNB_LIBRARY", "default")}
    results = {}
    for mode in ("analytic", "fd"):
        engine.set_substitution_gradient(mode)
        results[mode] = engine.gradients(batch, params, rescaling=True)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            engine.gradients(batch, params, rescaling=True)
        seconds = (time.perf_counter() - t0) / args.steps
        out[mode + "_ms_per_tree"] = seconds * 1e3 / args.trees
        out[mode + "_trees_per_s"] = args.trees / seconds
    a = np.array([g.gradient["substitution_model"] for g in results["analytic"]])
    d = np.array([g.gradient["substitution_model"] for g in results["fd"]])
    out["max_rel_difference"] = float(np.max(np.abs(a - d) / np.max(np.abs(d), axis=1, keepdims=True)))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
