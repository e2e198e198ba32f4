"""Parity tests proper: the CUDA path, called through the C ABI (via the host
mirror in libsbn_b200.engine), against the CPU oracle on identical inputs, the
committed reference fixtures, and the external goldens.

Tolerances are BASELINE.json's: log-likelihoods 1e-10 relative, branch
gradients 1e-8 relative (to max|g| of the tree, because gradients contain exact
zeros), everything in fp64.
"""
import os

import numpy as np
import pytest

from conftest import load_fixture
import libsbn_b200 as sbn
from libsbn_b200 import _capi, trees

pytestmark = pytest.mark.gpu

LOGL_RTOL = 1e-10
GRAD_RTOL = 1e-8

UNROOTED = ["hello_jc69", "ds1_jc69", "ds1_jc69_weibull4", "ds1_gtr_weibull4", "ds1_100_topologies_gtr_weibull4",
            "ds1_100_topologies_jc69", "ds1_tree0_gtr_equal"]
ROOTED = ["flua_jc69_strict", "flua_jc69_varied_rates", "flua_gtr_strict", "flua_jc69_weibull4_strict"]


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))


def grad_rel(a, b):
    a, b = np.atleast_2d(a), np.atleast_2d(b)
    return np.max(np.max(np.abs(a - b), axis=1) / np.maximum(np.max(np.abs(b), axis=1), 1e-300))


def fd_noise(log_likelihoods):
    """Noise floor of central differences with delta 1e-6 on logL values that
    carry ~1e-12 relative rounding (see tests/test_oracle.py)."""
    return np.abs(log_likelihoods).max() * 1e-12 / 1e-6


def engine_of(fx, device=0):
    spec = sbn.PhyloModelSpecification(fx["substitution"], fx["site"], fx["clock"])
    return sbn.Engine(spec, fx["patterns"], fx["weights"], device)


def batch_of(fx):
    if fx["rooted"]:
        return sbn.TreeBatch(fx["parent_ids"], fx["branch_lengths"], fx["rates"], fx["node_heights"],
                             fx["node_bounds"], fx["height_ratios"], 1)
    return sbn.TreeBatch(fx["parent_ids"], fx["branch_lengths"])


def stack(gradients, key):
    return np.array([g.gradient[key] for g in gradients])


@pytest.mark.parametrize("name", UNROOTED)
@pytest.mark.parametrize("rescaling", [False, True], ids=["plain", "rescaled"])
def test_unrooted_log_likelihoods(oracle, name, rescaling):
    fx = load_fixture(name)
    engine = engine_of(fx)
    got = engine.log_likelihoods(batch_of(fx), fx["params"], rescaling)
    want = oracle.log_likelihoods(fx["substitution"], fx["site"], fx["patterns"], fx["weights"],
                                  fx["parent_ids"], fx["branch_lengths"], fx["params"], rescaling=rescaling)
    assert rel(got, want) < LOGL_RTOL
    # and what the unmodified reference returned in the build container
    assert rel(got, fx["log_likelihoods"]) < LOGL_RTOL
    if "golden_log_likelihoods" in fx:
        assert np.max(np.abs(got - fx["golden_log_likelihoods"])) < fx["golden_tol"]


@pytest.mark.parametrize("name", UNROOTED)
@pytest.mark.parametrize("rescaling", [False, True], ids=["plain", "rescaled"])
def test_unrooted_gradients(oracle, name, rescaling):
    fx = load_fixture(name)
    engine = engine_of(fx)
    got = engine.gradients(batch_of(fx), fx["params"], rescaling)
    want = oracle.gradients(fx["substitution"], fx["site"], fx["patterns"], fx["weights"],
                            fx["parent_ids"], fx["branch_lengths"], fx["params"], rescaling=rescaling,
                            reference_quirks=False)
    logl = np.array([g.log_likelihood for g in got])
    assert rel(logl, want["log_likelihood"]) < LOGL_RTOL
    assert grad_rel(stack(got, "branch_lengths"), want["branch"]) < GRAD_RTOL
    tag = "_rescaled" if rescaling else ""
    assert grad_rel(stack(got, "branch_lengths"), fx["grad_branch_lengths" + tag]) < GRAD_RTOL
    if engine.category_count > 1:
        assert rel(stack(got, "site_model")[:, 0], want["site_model"]) < GRAD_RTOL
        # The reference evaluates this block on a model its finite-difference loop
        # left perturbed by 1e-6 (fat_beagle.cpp:433-436); we do not reproduce that.
        assert rel(stack(got, "site_model")[:, 0], fx["grad_site_model" + tag][:, 0]) < 1e-4
    if fx["substitution"] == "GTR":
        noise = fd_noise(logl)
        assert np.max(np.abs(stack(got, "substitution_model") - want["substitution_model"])) < noise
        assert np.max(np.abs(stack(got, "substitution_model") - fx["grad_substitution_model" + tag])) < noise
    else:
        assert "substitution_model" not in got[0].gradient


def test_ds1_physher_gradient_goldens(oracle):
    """unrooted_sbn_instance.hpp:234-257 / 284-335."""
    fx = load_fixture("ds1_jc69")
    got = engine_of(fx).gradients(batch_of(fx), fx["params"], False)
    assert np.max(np.abs(np.sort(got[-1].gradient["branch_lengths"]) - fx["golden_last_gradient_sorted"])) < 1e-4
    fx = load_fixture("ds1_jc69_weibull4")
    got = engine_of(fx).gradients(batch_of(fx), fx["params"], True)
    assert np.max(np.abs(stack(got, "branch_lengths")[:, 0] - fx["golden_first_branch_gradient"])) < 1e-4


@pytest.mark.parametrize("name", ROOTED)
def test_rooted(oracle, name):
    fx = load_fixture(name)
    engine = engine_of(fx)
    batch = batch_of(fx)
    kwargs = dict(rooted=True, rates=fx["rates"], node_heights=fx["node_heights"], node_bounds=fx["node_bounds"])
    args = (fx["substitution"], fx["site"], fx["patterns"], fx["weights"], fx["parent_ids"],
            fx["branch_lengths"], fx["params"])
    got = engine.log_likelihoods(batch, fx["params"], False, rooted=True)
    assert rel(got, oracle.log_likelihoods(*args, **kwargs)) < LOGL_RTOL
    assert rel(got, fx["log_likelihoods"]) < LOGL_RTOL
    plain = engine.unrooted_log_likelihoods(batch, fx["params"], False)
    assert rel(plain, oracle.log_likelihoods(*args)) < LOGL_RTOL
    grads = engine.gradients(batch, fx["params"], False, rooted=True)
    want = oracle.gradients(*args, height_ratios=fx["height_ratios"], **kwargs)
    logl = np.array([g.log_likelihood for g in grads])
    assert rel(logl, want["log_likelihood"]) < LOGL_RTOL
    assert grad_rel(stack(grads, "ratios_root_height"), want["ratios_root_height"]) < GRAD_RTOL
    assert grad_rel(stack(grads, "ratios_root_height"), fx["grad_ratios_root_height"]) < GRAD_RTOL
    assert grad_rel(stack(grads, "clock_model"), want["clock_model"]) < GRAD_RTOL
    assert "branch_lengths" not in grads[0].gradient
    if "golden_ratio_gradient" in fx:
        assert np.max(np.abs(grads[0].gradient["ratios_root_height"] - fx["golden_ratio_gradient"])) < 1e-4
        assert abs(got[0] - (fx["golden_log_likelihood"] + fx["golden_jacobian"])) < 1e-4
    if "golden_site_gradient" in fx:
        assert abs(grads[0].gradient["site_model"][0] - fx["golden_site_gradient"]) < 1e-3
    if "golden_substitution_gradient" in fx:
        assert np.max(np.abs(grads[0].gradient["substitution_model"] - fx["golden_substitution_gradient"])) < 1e-3


def test_relaxed_clock_gradient_layout(oracle):
    """rooted_sbn_instance.hpp:308-323: one rate per branch."""
    fx = load_fixture("flua_jc69_varied_rates")
    engine = engine_of(fx)
    edges = fx["rates"].shape[1]
    batch = sbn.TreeBatch(fx["parent_ids"], fx["branch_lengths"], fx["rates"], fx["node_heights"],
                          fx["node_bounds"], fx["height_ratios"], rate_count=edges)
    grads = engine.gradients(batch, fx["params"], False, rooted=True)
    want = oracle.gradients(fx["substitution"], fx["site"], fx["patterns"], fx["weights"], fx["parent_ids"],
                            fx["branch_lengths"], fx["params"], rooted=True, rates=fx["rates"],
                            node_heights=fx["node_heights"], node_bounds=fx["node_bounds"],
                            height_ratios=fx["height_ratios"], rate_count=edges)
    assert grads[0].gradient["clock_model"].shape == (edges,)
    assert grad_rel(stack(grads, "clock_model"), want["clock_model"]) < GRAD_RTOL


SYNTHETIC = [
    # taxa, patterns, model, site, trees, rooted
    (3, 1, "JC69", "constant", 2, False),
    (4, 77, "GTR", "constant", 3, False),
    (5, 1000, "JC69", "weibull+4", 3, True),
    (16, 513, "GTR", "weibull+3", 4, False),   # 3 categories pad to 4 lanes
    (16, 130, "GTR", "weibull+2", 2, False),
    (16, 130, "JC69", "weibull+8", 2, True),
    (11, 65, "GTR", "weibull+16", 2, False),
    (50, 300, "HKY", "weibull+4", 3, False),
    (100, 257, "GTR", "weibull+4", 5, False),
    (9, 700, "GTR", "weibull+5", 3, False),    # 5 and 6 categories pad to 8 lanes
    (12, 90, "HKY", "weibull+6", 2, True),
    (30, 5000, "GTR", "weibull+4", 2, False),  # many tiles per (tree, chunk) item
    (7, 3000, "JC69", "constant", 6, True),
    # discrete Gamma (not in the reference): the oracle gets scipy's category rates
    (10, 300, "GTR", "gamma+4", 3, False),
    (8, 200, "JC69", "gamma+5", 2, True),
    (12, 150, "HKY", "gamma", 2, False),
]


@pytest.mark.parametrize("taxa,patterns,substitution,site,tree_count,rooted", SYNTHETIC)
@pytest.mark.parametrize("rescaling", [False, True], ids=["plain", "rescaled"])
def test_synthetic_against_oracle(oracle, taxa, patterns, substitution, site, tree_count, rooted, rescaling):
    seed = taxa * 1000 + patterns
    rng = np.random.default_rng(seed)
    states, _ = trees.random_alignment(taxa, patterns, seed, gap_fraction=0.03)
    weights = rng.integers(1, 4, size=patterns).astype(np.float64)
    parent_ids, lengths = trees.random_tree_batch(taxa, tree_count, seed, rooted=rooted)
    engine = sbn.Engine(sbn.PhyloModelSpecification(substitution, site, "none"), states, weights)
    params = np.zeros((tree_count, engine.param_count))
    oracle_params = np.zeros((tree_count, 11))
    oracle_sub = substitution
    for t in range(tree_count):  # a different model per tree
        freqs = rng.dirichlet(np.ones(4) * 4)
        if substitution == "GTR":
            rates = rng.dirichlet(np.ones(6) * 2)
            params[t, :10] = np.concatenate([rates, freqs])
            oracle_params[t, :10] = params[t, :10]
        elif substitution == "HKY":
            kappa = rng.uniform(0.5, 4)
            params[t, :5] = np.concatenate([freqs, [kappa]])
            raw = np.array([1, kappa, 1, 1, kappa, 1.0])
            oracle_params[t, :10] = np.concatenate([raw / raw.sum(), freqs])
            oracle_sub = "GTR"
        if site != "constant":
            params[t, -1] = rng.uniform(0.3, 2)
            oracle_params[t, 10 if oracle_sub == "GTR" else 0] = params[t, -1]
    if oracle_sub != "GTR":
        oracle_params = oracle_params[:, :1]
    batch = sbn.TreeBatch(parent_ids, lengths)
    got = engine.log_likelihoods(batch, params, rescaling)
    want = oracle.log_likelihoods(oracle_sub, site, states, weights, parent_ids, lengths, oracle_params,
                                  rescaling=rescaling)
    assert rel(got, want) < LOGL_RTOL
    grads = engine.gradients(batch, params, rescaling)
    want = oracle.gradients(oracle_sub, site, states, weights, parent_ids, lengths, oracle_params,
                            rescaling=rescaling)
    assert rel([g.log_likelihood for g in grads], want["log_likelihood"]) < LOGL_RTOL
    assert grad_rel(stack(grads, "branch_lengths"), want["branch"]) < GRAD_RTOL
    if engine.category_count > 1:
        assert grad_rel(stack(grads, "site_model").T, want["site_model"][None, :]) < GRAD_RTOL


@pytest.mark.parametrize("shape", ["ladder", "random"])
def test_thousand_taxa_need_rescaling(oracle, shape):
    """At 1000 taxa per-site likelihoods underflow fp64; with rescaling on the
    result matches the (BEAGLE-style rescaled) oracle."""
    taxa, patterns = 1000, 96
    states, weights = trees.random_alignment(taxa, patterns, 11, gap_fraction=0.01)
    rng = np.random.default_rng(12)
    if shape == "ladder":
        parent_ids = trees.ladder_topology(taxa)[None, :]
    else:
        parent_ids = trees.random_unrooted_topology(taxa, rng)[None, :]
    lengths = np.maximum(rng.exponential(0.1, size=(1, 2 * taxa - 2)), 1e-6)
    engine = sbn.Engine(sbn.PhyloModelSpecification("HKY", "weibull+4", "none"), states, weights)
    params = np.array([[0.1, 0.2, 0.3, 0.4, 2.0, 0.5]])
    raw = np.array([1, 2, 1, 1, 2, 1.0])
    oracle_params = np.array([list(raw / raw.sum()) + [0.1, 0.2, 0.3, 0.4, 0.5]])
    batch = sbn.TreeBatch(parent_ids, lengths)
    plain = engine.log_likelihoods(batch, params, False)
    assert not np.isfinite(plain[0])  # underflow without rescaling, as in the reference
    grads = engine.gradients(batch, params, True)
    want = oracle.gradients("GTR", "weibull+4", states, weights, parent_ids, lengths, oracle_params,
                            rescaling=True)
    assert np.isfinite(grads[0].log_likelihood)
    assert rel(grads[0].log_likelihood, want["log_likelihood"][0]) < LOGL_RTOL
    assert grad_rel(grads[0].gradient["branch_lengths"], want["branch"][0]) < GRAD_RTOL


def test_rescaling_is_exact():
    """Power-of-two rescaling does not round: rescaled == plain to the last bits."""
    fx = load_fixture("ds1_gtr_weibull4")
    engine = engine_of(fx)
    plain = engine.gradients(batch_of(fx), fx["params"], False)
    scaled = engine.gradients(batch_of(fx), fx["params"], True)
    assert rel([g.log_likelihood for g in scaled], [g.log_likelihood for g in plain]) < 1e-14
    assert grad_rel(stack(scaled, "branch_lengths"), stack(plain, "branch_lengths")) < 1e-13


def test_staged_api_and_determinism():
    fx = load_fixture("ds1_gtr_weibull4")
    engine = engine_of(fx)
    batch = batch_of(fx)
    staged = engine.stage(batch, fx["params"])
    staged.run(_capi.MODE_LOG_LIKELIHOOD, False)
    logl = staged.fetch()
    assert np.array_equal(logl, engine.log_likelihoods(batch, fx["params"], False))
    staged.run(_capi.MODE_BRANCH_GRADIENT, True)
    logl2, grad, rgrad = staged.fetch(gradients=True)
    staged.run(_capi.MODE_BRANCH_GRADIENT, True)
    logl3, grad3, rgrad3 = staged.fetch(gradients=True)
    assert np.array_equal(logl2, logl3) and np.array_equal(grad, grad3) and np.array_equal(rgrad, rgrad3)
    assert staged.algorithmic_bytes(_capi.MODE_BRANCH_GRADIENT) == \
        (10 * 27 - 14) * 32.0 * 4 * 934 * batch.tree_count
    launches = engine.launch_count
    staged.run(_capi.MODE_LOG_LIKELIHOOD, False)
    assert engine.launch_count - launches == 3  # matrices, tree walk, reduction
    staged.close()


def test_pattern_sharding_sums_to_the_whole():
    """Site-pattern sharding: per-shard sums of logL and edge derivatives add up."""
    fx = load_fixture("ds1_jc69_weibull4")
    batch = batch_of(fx)
    whole = engine_of(fx)
    staged = whole.stage(batch, fx["params"])
    staged.run(_capi.MODE_BRANCH_GRADIENT, True)
    logl, grad, rgrad = staged.fetch(gradients=True)
    parts = []
    for begin, end in [(0, 300), (300, 934)]:
        shard = engine_of(fx)
        shard.set_pattern_range(begin, end)
        s = shard.stage(batch, fx["params"])
        s.run(_capi.MODE_BRANCH_GRADIENT, True)
        parts.append(s.fetch(gradients=True))
    assert rel(parts[0][0] + parts[1][0], logl) < 1e-13
    assert grad_rel(parts[0][1] + parts[1][1], grad) < 1e-12
    assert grad_rel(parts[0][2] + parts[1][2], rgrad) < 1e-12


def test_tree_sharding_is_a_slice():
    fx = load_fixture("ds1_100_topologies_jc69")
    engine = engine_of(fx)
    batch = batch_of(fx)
    whole = engine.log_likelihoods(batch, fx["params"], False)
    halves = [engine.log_likelihoods(batch.slice(0, 37), fx["params"][:37], False),
              engine.log_likelihoods(batch.slice(37, 100), fx["params"][37:], False)]
    assert np.array_equal(np.concatenate(halves), whole)


def test_error_behaviour():
    fx = load_fixture("ds1_gtr_weibull4")
    engine = engine_of(fx)
    bad = fx["params"].copy()
    bad[:, 6:10] = 0.3  # frequencies sum to 1.2
    with pytest.raises(RuntimeError, match="do not sum to 1"):
        engine.log_likelihoods(batch_of(fx), bad, False)
    with pytest.raises(RuntimeError):
        engine.log_likelihoods(sbn.TreeBatch(fx["parent_ids"][:, :-2], fx["branch_lengths"][:, :-2]),
                               fx["params"], False)
    # the engine is still usable after a failed call
    good = engine.log_likelihoods(batch_of(fx), fx["params"], False)
    assert rel(good, fx["log_likelihoods"]) < LOGL_RTOL
    empty = engine.log_likelihoods(batch_of(fx).slice(0, 0), fx["params"][:0], False)
    assert empty.shape == (0,)


def test_unrooted_gradient_of_a_bifurcating_tree_slides_the_root(oracle):
    """The reference's unrooted Gradient applies Tree::SlideRootPosition after
    Detrifurcate (fat_beagle.cpp:470-472, tree.cpp:72-78): a no-op for a detrifurcated
    tree, but for bifurcating input the root's second child gets length 0 and its
    first child the sum -- the engine must do the same (and the oracle does)."""
    fx = load_fixture("flua_jc69_weibull4_strict")
    spec = sbn.PhyloModelSpecification("JC69", "weibull+4", "none")
    engine = sbn.Engine(spec, fx["patterns"], fx["weights"])
    batch = sbn.TreeBatch(fx["parent_ids"], fx["branch_lengths"])
    params = fx["params"][:, :1]
    got = engine.gradients(batch, params, True)
    want = oracle.gradients("JC69", "weibull+4", fx["patterns"], fx["weights"], fx["parent_ids"],
                            fx["branch_lengths"], params, rescaling=True)
    assert rel([g.log_likelihood for g in got], want["log_likelihood"]) < LOGL_RTOL
    assert grad_rel(stack(got, "branch_lengths"), want["branch"]) < GRAD_RTOL
    assert grad_rel(stack(got, "site_model").T, want["site_model"][None, :]) < GRAD_RTOL


@pytest.mark.parametrize("substitution,site", [("GTR", "weibull+4"), ("GTR", "constant"), ("HKY", "weibull+4"),
                                               ("HKY", "gamma+3"), ("GTR", "weibull+8")])
@pytest.mark.parametrize("rescaling", [False, True], ids=["plain", "rescaled"])
def test_analytic_substitution_gradient_matches_finite_differences(substitution, site, rescaling):
    """the exact d logL / d (substitution parameters), accumulated inside the
    gradient sweep, against the reference's own recipe (16 central-difference log-likelihood
    sweeps, fat_beagle.cpp:400-465) run by the same engine: equal at the finite-difference
    noise floor, everything else bit for bit."""
    fx = load_fixture("ds1_gtr_weibull4")
    spec = sbn.PhyloModelSpecification(substitution, site, "none")
    engine = sbn.Engine(spec, fx["patterns"], fx["weights"])
    batch = sbn.TreeBatch(fx["parent_ids"], fx["branch_lengths"])
    rng = np.random.default_rng(5)
    rows = []
    for _ in range(batch.tree_count):  # a different model per tree
        rates, freqs = rng.dirichlet(np.full(6, 5.0)), rng.dirichlet(np.full(4, 5.0))
        sub = list(rates) + list(freqs) if substitution == "GTR" else list(freqs) + [rng.uniform(0.5, 4.0)]
        rows.append(sub + ([] if site == "constant" else [rng.uniform(0.3, 1.5)]))
    params = np.array(rows)
    engine.set_substitution_gradient("fd")
    by_differences = engine.gradients(batch, params, rescaling)
    engine.set_substitution_gradient("analytic")
    analytic = engine.gradients(batch, params, rescaling)
    logl = np.array([g.log_likelihood for g in analytic])
    assert np.array_equal(logl, [g.log_likelihood for g in by_differences])
    assert np.array_equal(stack(analytic, "branch_lengths"), stack(by_differences, "branch_lengths"))
    a, d = stack(analytic, "substitution_model"), stack(by_differences, "substitution_model")
    assert a.shape == (batch.tree_count, 8 if substitution == "GTR" else 4)
    assert np.max(np.abs(a - d)) < 2 * fd_noise(logl), (a[0], d[0])
    assert np.max(np.abs(a - d) / np.max(np.abs(d), axis=1, keepdims=True)) < 1e-4


def test_repeated_calls_on_unchanged_topologies_replay_a_graph(oracle):
    """Variational inference re-evaluates the same trees with new branch lengths and
    parameters (vip/burrito.py:84-117): the traversal programs stay on the device and
    from the third call on the launch sequence is replayed as a CUDA graph.  Every call
    must still see the new lengths / parameters."""
    fx = load_fixture("ds1_gtr_weibull4")
    engine = engine_of(fx)
    rng = np.random.default_rng(3)
    launches = []
    for call in range(6):
        lengths = fx["branch_lengths"] * rng.uniform(0.5, 1.5, size=fx["branch_lengths"].shape)
        lengths[:, -1] = 0.0
        params = fx["params"].copy()
        params[:, -1] = rng.uniform(0.3, 1.2)  # the Weibull shape
        batch = sbn.TreeBatch(fx["parent_ids"], lengths)
        rescaling = bool(call % 2 == 0) if call >= 4 else True
        before = engine.launch_count
        logl = engine.log_likelihoods(batch, params, rescaling)
        got = engine.gradients(batch, params, rescaling)
        launches.append(engine.launch_count - before)
        want = oracle.gradients(fx["substitution"], fx["site"], fx["patterns"], fx["weights"], fx["parent_ids"],
                                lengths, params, rescaling=rescaling)
        assert rel(logl, want["log_likelihood"]) < LOGL_RTOL
        assert rel([g.log_likelihood for g in got], want["log_likelihood"]) < LOGL_RTOL
        assert grad_rel(stack(got, "branch_lengths"), want["branch"]) < GRAD_RTOL
        assert np.max(np.abs(stack(got, "substitution_model") - want["substitution_model"])) < fd_noise(logl)
    assert len(set(launches)) == 1  # kernels counted alike, launched eagerly or replayed
