"""Pins the numpy GP oracle (oracle/gp.py) to the UNMODIFIED reference GP code:
every scenario of tests/gp_cases.py against the fixtures gp_dump produced."""
import numpy as np
import pytest

import gp_cases
from conftest import load_fixture
from oracle import gp


def factory(*args, **kwargs):
    return gp.GPEngineOracle(*args, **kwargs)


def plv(engine, index):
    return engine.plvs[index]


def counts(engine):
    return engine.rescaling_counts


@pytest.mark.parametrize("name", gp_cases.POPULATE_ONLY)
def test_populate_and_likelihoods(name):
    gp_cases.run_populate(factory, name, plv, counts)


@pytest.mark.parametrize("name", [n for n in gp_cases.ESTIMATE if n != "ds1_dag"])
def test_estimate_branch_lengths_and_sbn_parameters(name):
    gp_cases.run_estimate(factory, name)


def test_ds1_dag_populate():
    """BASELINE.json configs[2], PLV sweep + marginal likelihood (the Brent part
    of this DAG is covered on the GPU; in numpy it takes minutes)."""
    fx = load_fixture("gp_ds1_dag")
    engine = gp_cases.make_engine(factory, fx)
    engine.process_operations(fx["program_populate_plvs"])
    engine.process_operations(fx["program_compute_likelihoods"])
    gp_cases.check_state(engine, fx, "populated_", None, lambda: engine.rescaling_counts)


@pytest.mark.parametrize("name", gp_cases.QUARTET)
def test_quartet_hybrid_marginals(name):
    engine, fx = gp_cases.run_populate(factory, name)
    lists = gp_cases.quartet_lists(fx)
    got = engine.calculate_quartet_hybrid_likelihoods(int(fx["quartet_central_gpcsp"]), *lists)
    gp_cases.close(got, fx["quartet_log_likelihoods"])
    for request in gp.split_hybrid_requests(fx["hybrid_requests"]):
        engine.process_quartet_hybrid_request(*request)
    gp_cases.close(engine.hybrid_marginal_log_likelihoods, fx["hybrid_marginals"])


def test_rescaling_thresholds_agree():
    """gp_doctest.cpp:243-253: the marginal does not depend on the threshold."""
    values = [load_fixture(f"gp_flua_threshold_{t}")["populated_log_marginal_likelihood"]
              for t in ("1e-40", "1e-4", "0.5")]
    assert max(values) - min(values) < 1e-10
    assert load_fixture("gp_flua_threshold_0.5")["populated_rescaling_counts"].max() >= 3


def test_transition_matrix():
    """gp_engine.hpp:216-225."""
    fx = load_fixture("gp_hello")
    engine = gp_cases.make_engine(factory, fx)
    matrix = engine.transition_matrix(0.75)
    assert abs(0.52590958087 - matrix[0, 0]) < 1e-10
    assert abs(0.1580301397 - matrix[0, 1]) < 1e-10


def test_log_add_matches_reference_branches():
    assert gp.log_add(-np.inf, -np.inf) == -np.inf
    assert gp.log_add(0.0, -100.0) == 0.0
    assert abs(gp.log_add(np.log(2), np.log(3)) - np.log(5)) < 1e-15
    a = np.array([-np.inf, 0.0, np.log(2), -np.inf])
    b = np.array([-np.inf, -100.0, np.log(3), 1.5])
    expect = [gp.log_add(x, y) for x, y in zip(a, b)]
    assert np.array_equal(gp.log_add_vectors(a, b), expect)


def test_asserts_surface():
    fx = load_fixture("gp_hello")
    engine = gp_cases.make_engine(factory, fx)
    engine.rescaling_counts[3] = 2
    with pytest.raises(gp.GPAssertion, match="dest_ rescaling too large"):
        engine.process_operations([2, 3, 0, 0])
