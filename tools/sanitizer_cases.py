"""Small invocations of every tree-walk instantiation, for compute-sanitizer.

  compute-sanitizer --tool memcheck  python tools/sanitizer_cases.py
  compute-sanitizer --tool racecheck python tools/sanitizer_cases.py

Cases: DS1 (27 taxa x 934 patterns) GTR+weibull4 and JC69 from the committed
fixtures, with and without rescaling, logL and gradients; a 1000-taxon ladder
(deepest traversal, one stack slot) and a 300-taxon random tree (several stack
slots) on a few hundred patterns.  Results are compared with the oracle so a
sanitized run that silently computes garbage is not counted as clean.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import libsbn_b200 as sbn  # noqa: E402
from libsbn_b200 import trees  # noqa: E402
from oracle import phylo  # noqa: E402
from conftest import load_fixture  # noqa: E402


def check(name, substitution, site, states, weights, parent_ids, lengths, params, rescaling):
    engine = sbn.Engine(sbn.PhyloModelSpecification(substitution, site, "none"), states, weights, 0)
    batch = sbn.TreeBatch(parent_ids, lengths)
    logl = engine.log_likelihoods(batch, params, rescaling)
    got = engine.gradients(batch, params, rescaling)
    want = phylo.gradients(substitution, site, states, weights, parent_ids, lengths, params, rescaling=rescaling)
    g = np.array([x.gradient["branch_lengths"] for x in got])
    e_logl = np.max(np.abs(logl - want["log_likelihood"]) / np.abs(want["log_likelihood"]))
    e_grad = np.max(np.abs(g - want["branch"])) / np.max(np.abs(want["branch"]))
    print(f"{name:40s} rescaling={int(rescaling)} logL rel.err {e_logl:.1e} gradient rel.err {e_grad:.1e}",
          flush=True)
    assert e_logl < 1e-10 and e_grad < 1e-8, name


def main():
    for name in ["ds1_gtr_weibull4", "ds1_jc69"]:
        fx = load_fixture(name)
        for rescaling in (False, True):
            check(name, fx["substitution"], fx["site"], fx["patterns"], fx["weights"], fx["parent_ids"][:3],
                  fx["branch_lengths"][:3], fx["params"][:3], rescaling)
    row = np.array([[0.05, 0.1, 0.15, 0.20, 0.25, 0.25, 0.1, 0.2, 0.3, 0.4, 0.5]])
    for label, taxa, patterns in [("ladder", 1000, 300), ("random", 300, 700)]:
        states, weights = trees.random_alignment(taxa, patterns, seed=11, gap_fraction=0.02)
        rng = np.random.default_rng(12)
        parent_ids = (trees.ladder_topology(taxa) if label == "ladder"
                      else trees.random_unrooted_topology(taxa, rng))[None, :]
        lengths = np.maximum(rng.exponential(0.05, size=(1, parent_ids.shape[1] + 1)), 1e-6)
        lengths[:, -1] = 0.0
        check(f"{label} {taxa} taxa x {patterns} patterns GTR+weibull4", "GTR", "weibull+4", states, weights,
              parent_ids, lengths, row, True)
    print("sanitizer cases: all results match the oracle")


if __name__ == "__main__":
    main()
