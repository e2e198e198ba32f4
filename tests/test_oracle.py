"""Pins the CPU oracle (oracle/phylo_oracle.cpp + oracle/beagle_cpu.cpp).

Two anchors per scenario:
  * the outputs of the UNMODIFIED reference run in the build container
    (tests/golden/*.npz, made by tests/golden/make_fixtures.py), and
  * the external goldens (pybeagle / physher / phylotorch) hard-coded in the
    reference's own doctests, at the reference's own tolerances -- and far
    tighter where the golden carries the digits.
"""
import numpy as np
import pytest

from conftest import load_fixture

UNROOTED = ["hello_jc69", "ds1_jc69", "ds1_jc69_weibull4", "ds1_gtr_weibull4", "ds1_100_topologies_gtr_weibull4",
            "ds1_100_topologies_jc69", "ds1_tree0_gtr_equal"]
ROOTED = ["flua_jc69_strict", "flua_jc69_varied_rates", "flua_gtr_strict", "flua_jc69_weibull4_strict"]


def rel(a, b):
    return np.max(np.abs(np.asarray(a) - np.asarray(b)) / np.maximum(np.abs(np.asarray(b)), 1e-300))


def grad_rel(a, b):
    """Relative to max|g| per tree (gradients have entries that are exactly 0)."""
    a, b = np.atleast_2d(a), np.atleast_2d(b)
    return np.max(np.max(np.abs(a - b), axis=1) / np.maximum(np.max(np.abs(b), axis=1), 1e-300))


def fd_noise(log_likelihoods):
    """Noise floor of the reference's central differences (delta = 1e-6,
    fat_beagle.cpp:454): two evaluations of logL that each carry ~1e-12 relative
    summation-order rounding, divided by 2 delta.  The reference's own test
    tolerance for these entries is 1e-3 (rooted_sbn_instance.hpp:349-352)."""
    return np.abs(log_likelihoods).max() * 1e-12 / 1e-6


def model_params(fx):
    """The substitution + site columns of the reference's parameter matrix."""
    return fx["params"]


@pytest.mark.parametrize("name", UNROOTED)
@pytest.mark.parametrize("rescaling", [False, True])
def test_unrooted_log_likelihoods_match_reference(oracle, name, rescaling):
    fx = load_fixture(name)
    got = oracle.log_likelihoods(fx["substitution"], fx["site"], fx["patterns"], fx["weights"],
                                 fx["parent_ids"], fx["branch_lengths"], model_params(fx),
                                 rescaling=rescaling)
    want = fx["log_likelihoods_rescaled" if rescaling else "log_likelihoods"]
    # Same arithmetic, different site-pattern order: summation rounding only.
    assert rel(got, want) < 1e-12


@pytest.mark.parametrize("name", UNROOTED)
def test_tip_partials_equal_tip_states(oracle, name):
    """unrooted_sbn_instance.hpp:219-232 runs every case with both tip encodings."""
    fx = load_fixture(name)
    args = (fx["substitution"], fx["site"], fx["patterns"], fx["weights"], fx["parent_ids"][:3],
            fx["branch_lengths"][:3], model_params(fx)[:3])
    states = oracle.log_likelihoods(*args, use_tip_states=True)
    partials = oracle.log_likelihoods(*args, use_tip_states=False)
    assert rel(states, partials) < 1e-13


@pytest.mark.parametrize("name", UNROOTED)
@pytest.mark.parametrize("rescaling", [False, True])
def test_unrooted_gradients_match_reference(oracle, name, rescaling):
    fx = load_fixture(name)
    tag = "_rescaled" if rescaling else ""
    got = oracle.gradients(fx["substitution"], fx["site"], fx["patterns"], fx["weights"],
                           fx["parent_ids"], fx["branch_lengths"], model_params(fx),
                           rescaling=rescaling, reference_quirks=True)
    assert rel(got["log_likelihood"], fx["grad_log_likelihood" + tag]) < 1e-12
    assert grad_rel(got["branch"], fx["grad_branch_lengths" + tag]) < 1e-10
    if "grad_site_model" + tag in fx:
        assert rel(got["site_model"], fx["grad_site_model" + tag][:, 0]) < 1e-8
    if "grad_substitution_model" + tag in fx:
        assert np.max(np.abs(got["substitution_model"] - fx["grad_substitution_model" + tag])) < \
            fd_noise(fx["grad_log_likelihood" + tag])


def test_site_model_quirk_is_small(oracle):
    """The reference evaluates the site-model sweep on a model left perturbed by
    its finite-difference loop (fat_beagle.cpp:433-436); quantify the effect."""
    fx = load_fixture("ds1_gtr_weibull4")
    args = (fx["substitution"], fx["site"], fx["patterns"], fx["weights"], fx["parent_ids"][:2],
            fx["branch_lengths"][:2], model_params(fx)[:2])
    quirk = oracle.gradients(*args, reference_quirks=True)["site_model"]
    clean = oracle.gradients(*args, reference_quirks=False)["site_model"]
    assert rel(quirk, clean) < 1e-4
    assert rel(quirk, clean) > 0  # it IS different


@pytest.mark.parametrize("name", ROOTED)
def test_rooted_match_reference(oracle, name):
    fx = load_fixture(name)
    kwargs = dict(rooted=True, rates=fx["rates"], node_heights=fx["node_heights"],
                  node_bounds=fx["node_bounds"])
    args = (fx["substitution"], fx["site"], fx["patterns"], fx["weights"], fx["parent_ids"],
            fx["branch_lengths"], model_params(fx))
    got = oracle.log_likelihoods(*args, **kwargs)
    assert rel(got, fx["log_likelihoods"]) < 1e-12
    grads = oracle.gradients(*args, height_ratios=fx["height_ratios"], reference_quirks=True, **kwargs)
    assert rel(grads["log_likelihood"], fx["grad_log_likelihood"]) < 1e-12
    assert grad_rel(grads["ratios_root_height"], fx["grad_ratios_root_height"]) < 1e-9
    assert grad_rel(grads["clock_model"], fx["grad_clock_model"]) < 1e-9
    if "grad_site_model" in fx:
        assert rel(grads["site_model"], fx["grad_site_model"][:, 0]) < 1e-8
    if "grad_substitution_model" in fx:
        assert np.max(np.abs(grads["substitution_model"] - fx["grad_substitution_model"])) < \
            fd_noise(fx["grad_log_likelihood"])


# ---- external goldens (numbers from other programs, in the reference's tests) ----

def test_hello_golden(oracle):
    fx = load_fixture("hello_jc69")
    got = oracle.log_likelihoods("JC69", "constant", fx["patterns"], fx["weights"], fx["parent_ids"],
                                 fx["branch_lengths"])
    assert abs(got[0] - fx["golden_log_likelihoods"][0]) < 1e-6


def test_ds1_pybeagle_goldens(oracle):
    """unrooted_sbn_instance.hpp:225-231 at the reference's 1.1e-4, and at 1e-9
    absolute because the pybeagle numbers are full precision."""
    fx = load_fixture("ds1_jc69")
    got = oracle.log_likelihoods("JC69", "constant", fx["patterns"], fx["weights"], fx["parent_ids"],
                                 fx["branch_lengths"])
    assert np.max(np.abs(got - fx["golden_log_likelihoods"])) < 1e-9
    grads = oracle.gradients("JC69", "constant", fx["patterns"], fx["weights"], fx["parent_ids"],
                             fx["branch_lengths"])
    last = np.sort(grads["branch"][-1])
    assert np.max(np.abs(last - fx["golden_last_gradient_sorted"])) < 1e-4


def test_ds1_physher_weibull_goldens(oracle):
    fx = load_fixture("ds1_jc69_weibull4")
    got = oracle.gradients("JC69", "weibull+4", fx["patterns"], fx["weights"], fx["parent_ids"],
                           fx["branch_lengths"], fx["params"])
    assert np.max(np.abs(got["log_likelihood"] - fx["golden_log_likelihoods"])) < 1e-9
    assert np.max(np.abs(got["branch"][:, 0] - fx["golden_first_branch_gradient"])) < 1e-4


def test_flua_physher_goldens(oracle):
    fx = load_fixture("flua_jc69_strict")
    kwargs = dict(rooted=True, rates=fx["rates"], node_heights=fx["node_heights"],
                  node_bounds=fx["node_bounds"])
    args = ("JC69", "constant", fx["patterns"], fx["weights"], fx["parent_ids"], fx["branch_lengths"],
            fx["params"])
    got = oracle.log_likelihoods(*args, **kwargs)
    assert abs(got[0] - (fx["golden_log_likelihood"] + fx["golden_jacobian"])) < 1e-4
    grads = oracle.gradients(*args, height_ratios=fx["height_ratios"], **kwargs)
    assert abs(grads["log_likelihood"][0] - fx["golden_log_likelihood"]) < 1e-4
    assert np.max(np.abs(grads["ratios_root_height"][0] - fx["golden_ratio_gradient"])) < 1e-4


def test_flua_phylotorch_gtr_goldens(oracle):
    fx = load_fixture("flua_gtr_strict")
    kwargs = dict(rooted=True, rates=fx["rates"], node_heights=fx["node_heights"],
                  node_bounds=fx["node_bounds"])
    grads = oracle.gradients("GTR", "constant", fx["patterns"], fx["weights"], fx["parent_ids"],
                             fx["branch_lengths"], fx["params"], height_ratios=fx["height_ratios"],
                             **kwargs)
    assert abs(grads["log_likelihood"][0] - fx["golden_log_likelihood"]) < 1e-3
    assert np.max(np.abs(grads["substitution_model"][0] - fx["golden_substitution_gradient"])) < 1e-3


def test_flua_physher_weibull_goldens(oracle):
    fx = load_fixture("flua_jc69_weibull4_strict")
    kwargs = dict(rooted=True, rates=fx["rates"], node_heights=fx["node_heights"],
                  node_bounds=fx["node_bounds"])
    grads = oracle.gradients("JC69", "weibull+4", fx["patterns"], fx["weights"], fx["parent_ids"],
                             fx["branch_lengths"], fx["params"], height_ratios=fx["height_ratios"],
                             **kwargs)
    assert abs(grads["log_likelihood"][0] - fx["golden_log_likelihood"]) < 1e-4
    assert abs(grads["site_model"][0] - fx["golden_site_gradient"]) < 1e-3


def test_jc69_equals_gtr_with_equal_parameters(oracle):
    """test/test_libsbn.py:95-118."""
    fx = load_fixture("ds1_tree0_gtr_equal")
    args = (fx["patterns"], fx["weights"], fx["parent_ids"], fx["branch_lengths"])
    jc = oracle.log_likelihoods("JC69", "constant", *args)
    gtr = oracle.log_likelihoods("GTR", "constant", *args, fx["params"])
    assert rel(jc, gtr) < 1e-12
