"""Small invocations of the GP interpreter (one CTA, several CTAs through the mailbox, rate categories,
a Brent sweep) and of the BEAGLE-compatible device library, for compute-sanitizer:

  compute-sanitizer --tool memcheck  python tools/sanitizer_cases_gp_beagle.py
  compute-sanitizer --tool racecheck python tools/sanitizer_cases_gp_beagle.py

Results are compared with the numpy / CPU oracles inside the sanitized run.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import libsbn_b200 as sbn  # noqa: E402
from libsbn_b200 import beagle  # noqa: E402
from oracle import gp  # noqa: E402
from conftest import load_fixture  # noqa: E402


def gp_case(name, patterns, site):
    fx = load_fixture("gp_five_taxon")
    rng = np.random.default_rng(patterns)
    tips = rng.integers(0, 4, size=(int(fx["taxon_count"]), patterns)).astype(np.uint8)
    tips[rng.random(tips.shape) < 0.03] = 4
    weights = rng.integers(1, 5, size=patterns).astype(np.float64)
    args = (tips, weights, int(weights.sum()), fx["plv_count"], fx["gpcsp_count"])
    kwargs = dict(rescaling_threshold=1e-3, sbn_prior=fx["sbn_prior"],
                  unconditional_node_probabilities=fx["unconditional_node_probabilities"],
                  inverted_sbn_prior=fx["inverted_sbn_prior"])
    ours, want = sbn.GPEngine(*args, device=0, **kwargs), gp.GPEngineOracle(*args, **kwargs)
    lengths = rng.uniform(0.01, 0.3, size=int(fx["gpcsp_count"]))
    for engine in (ours, want):
        if site != "constant":
            engine.set_site_model(site, [0.6])
        engine.set_branch_lengths(lengths)
        engine.process_operations(fx["program_populate_plvs"])
        engine.process_operations(fx["program_compute_likelihoods"])
        engine.process_operations(fx["program_branch_length_optimization"])
        engine.process_operations(fx["program_populate_plvs"])
        engine.process_operations(fx["program_marginal_likelihood"])
    a, b = ours.get_log_marginal_likelihood(), want.get_log_marginal_likelihood()
    print(f"GP {name:45s} log marginal rel.err {abs(a - b) / abs(b):.1e}", flush=True)
    assert abs(a - b) < 1e-6 * abs(b), name


def beagle_case(categories, use_tip_states, rescaling):
    rng = np.random.default_rng(categories)
    n, P = 13, 1003
    states = rng.integers(0, 4, size=(n, P)).astype(np.int32)
    states[rng.random(states.shape) < 0.05] = 4
    weights = rng.integers(1, 6, size=P).astype(np.float64)
    post, pre = beagle.random_tree_operations(n, rng, rescaling)
    lengths = rng.exponential(0.1, size=2 * n - 1)
    evec, ivec, evals, freqs, q = beagle.gtr_eigensystem()
    rates = np.sort(rng.gamma(2.0, 0.5, size=categories))
    rates /= rates.mean()
    results = []
    for path in (None, os.path.join(ROOT, "oracle", "_build", "libbeagle_oracle.so")):
        instance = beagle.Beagle(path, n, P, categories, use_tip_states)
        instance.set_tips(states, weights, use_tip_states)
        instance.set_model(evec, ivec, evals, freqs, rates, np.full(categories, 1.0 / categories))
        results.append(instance.log_likelihood_and_gradient(post, pre, lengths, q, rates, freqs, rescaling))
        instance.close()
    e_logl = abs(results[0][0] - results[1][0]) / abs(results[1][0])
    e_grad = np.max(np.abs(results[0][1] - results[1][1])) / np.max(np.abs(results[1][1]))
    print(f"BEAGLE library C={categories} tip_states={int(use_tip_states)} rescaling={int(rescaling)}: logL rel.err "
          f"{e_logl:.1e} derivative rel.err {e_grad:.1e}", flush=True)
    assert e_logl < 1e-12 and e_grad < 1e-10


if __name__ == "__main__":
    gp_case("one CTA (100 patterns)", 100, "constant")
    gp_case("8 CTAs through the mailbox (1000 patterns)", 1000, "constant")
    gp_case("weibull+4, 16 CTAs (500 patterns x 4 categories)", 500, "weibull+4")
    gp_case("strided: 148 CTAs x 256 threads (40000 patterns)", 40000, "constant")
    beagle_case(4, True, True)
    beagle_case(3, False, False)
    beagle_case(1, True, True)
