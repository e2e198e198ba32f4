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


# ---- the device's schedule of a program (sbnb_gp_schedule_program; host code of libsbn_b200.so)
FLAG = 1 << 30


def _records(program):
    """(kind, fields, batch length, flag) per record of a (scheduled) program."""
    size = {0: 2, 1: 3, 2: 4, 3: 4, 4: 4, 5: 4, 6: 3, 7: 1, 8: 4}
    program = [int(w) for w in program]
    pc, out = 0, []
    while pc < len(program):
        kind = program[pc] & 0xff
        words = size[kind] if kind in size else (3 + program[pc + 2] if kind == 9 else 3 + 2 * program[pc + 2])
        out.append((kind, tuple(program[pc + 1:pc + words]), (program[pc] >> 8) & 0xff, bool(program[pc] & FLAG)))
        pc += words
    return out


def _unfused(records):
    """The reference's records behind a schedule: fused accumulations split into their increments."""
    out = []
    for kind, fields, _, _ in records:
        if kind == 10:
            dest, count = fields[:2]
            out += [(2, (dest, fields[2 + 2 * i], fields[3 + 2 * i])) for i in range(count)]
        else:
            out.append((kind, fields))
    return out


@pytest.mark.parametrize("name,programs", [
    ("ds1_dag", ["program_populate_plvs", "program_compute_likelihoods", "program_branch_length_optimization",
                 "program_optimize_sbn_parameters", "program_marginal_likelihood"]),
    ("five_taxon", ["program_populate_plvs", "program_branch_length_optimization"]),
])
def test_schedule_is_a_batched_permutation_of_the_program(name, programs):
    from libsbn_b200.gp_engine import schedule_program
    fx = load_fixture("gp_" + name)
    for key in programs:
        scheduled = schedule_program(fx["plv_count"], fx["gpcsp_count"], fx[key])
        assert scheduled.size <= fx[key].size
        before, after = _records(fx[key]), _records(scheduled)
        assert sorted(_unfused(before)) == sorted(_unfused(after))
        # batch lengths: a tagged record is followed by length - 1 untagged records of its kind
        i = 0
        while i < len(after):
            kind, _, length, _ = after[i]
            assert length >= 1
            assert all(after[i + j][0] == kind and after[i + j][2] == 0 for j in range(1, length))
            i += length
    populate = _records(schedule_program(fx["plv_count"], fx["gpcsp_count"], fx["program_populate_plvs"]))
    assert sum(1 for record in populate if record[2] >= 1) < len(populate) / 2  # most ops ride in batches
    assert any(kind == 10 and flag for kind, _, _, flag in populate)  # accumulations fused, starting fresh
    assert any(kind == 0 and flag for kind, _, _, flag in populate)   # ... their ZeroPLVs reduced to the count


@pytest.mark.parametrize("name", ["five_taxon", "hello_two_trees", "seven_taxon_all_trees", "five_taxon_threshold_0.5"])
def test_scheduled_programs_compute_the_same_bits(name):
    """Ops of one dependency level are independent, a fused accumulation adds in the reference's order,
    and a PLV whose ZeroPLV was reduced to its count is overwritten before anybody reads it (the oracle
    poisons it): the scheduled program run record by record on the numpy oracle gives bit-identical
    PLVs, branch lengths, q and likelihoods."""
    from libsbn_b200.gp_engine import schedule_program
    fx = load_fixture("gp_" + name)
    engines = [gp_cases.make_engine(factory, fx), gp_cases.make_engine(factory, fx)]
    keys = ["program_populate_plvs", "program_compute_likelihoods", "program_branch_length_optimization",
            "program_populate_plvs", "program_compute_likelihoods", "program_optimize_sbn_parameters"]
    for key in keys:
        if key not in fx:
            continue
        engines[0].process_operations(fx[key])
        engines[1].process_operations(schedule_program(fx["plv_count"], fx["gpcsp_count"], fx[key]))
        assert np.array_equal(engines[0].plvs, engines[1].plvs)
        assert np.array_equal(engines[0].branch_lengths, engines[1].branch_lengths)
        assert np.array_equal(engines[0].q, engines[1].q)
        assert np.array_equal(engines[0].rescaling_counts, engines[1].rescaling_counts)
        assert np.array_equal(engines[0].log_likelihoods, engines[1].log_likelihoods, equal_nan=True)
        assert np.array_equal(engines[0].log_marginal_likelihood, engines[1].log_marginal_likelihood)


# ---- GP beyond JC69 (not in the reference): pinned through a single-tree DAG
GTR_PARAMS = np.array([0.05, 0.1, 0.15, 0.20, 0.25, 0.25, 0.1, 0.2, 0.3, 0.4])  # rooted_sbn_instance.hpp:336-337


def test_gtr_marginal_of_a_single_tree_dag_is_the_tree_likelihood():
    """data/hello_rooted.nwk = (jupiter,(mars,saturn)): with every branch at 0.1 the GP marginal of its
    one-tree DAG is the likelihood of the unrooted star tree with the outgroup's edge at 0.2 -- computed
    by the C restatement of the FatBeagle path (oracle/phylo.py, pinned to the reference's goldens)."""
    from oracle import phylo
    fx = load_fixture("gp_hello")
    weights = fx["pattern_weights"]
    for substitution, params in (("JC69", np.zeros(0)), ("GTR", GTR_PARAMS), ("HKY", np.array([0.1, 0.2, 0.3, 0.4, 2.0]))):
        engine = gp_cases.make_engine(factory, fx)
        engine.set_substitution_model(substitution, params)
        engine.set_branch_lengths(np.full(int(fx["gpcsp_count"]), 0.1))
        engine.process_operations(fx["program_populate_plvs"])
        engine.process_operations(fx["program_compute_likelihoods"])
        tree_params = params if substitution != "HKY" else np.concatenate([np.array([1, 2, 1, 1, 2, 1]) / 8, params[:4]])
        want = phylo.log_likelihoods("JC69" if substitution == "JC69" else "GTR", "constant", fx["tip_states"], weights,
                                     np.array([[3, 3, 3]]), np.array([[0.2, 0.1, 0.1, 0.0]]),
                                     params=tree_params[None, :] if tree_params.size else None)[0]
        assert abs(engine.get_log_marginal_likelihood() - want) < 1e-10 * abs(want), substitution


def test_rate_categories_on_a_single_tree_dag_are_the_tree_likelihood():
    """The same pin for rate categories in the GP engine (not in the reference): GTR + weibull+4 on
    the one-tree DAG equals FatBeagle's GTR + weibull+4 log likelihood of the tree."""
    from oracle import phylo
    fx = load_fixture("gp_hello")
    for site, shape in (("weibull+4", 0.5), ("weibull+2", 1.3)):
        engine = gp_cases.make_engine(factory, fx)
        engine.set_substitution_model("GTR", GTR_PARAMS)
        engine.set_site_model(site, [shape])
        engine.set_branch_lengths(np.full(int(fx["gpcsp_count"]), 0.1))
        engine.process_operations(fx["program_populate_plvs"])
        engine.process_operations(fx["program_compute_likelihoods"])
        want = phylo.log_likelihoods("GTR", site, fx["tip_states"], fx["pattern_weights"], np.array([[3, 3, 3]]),
                                     np.array([[0.2, 0.1, 0.1, 0.0]]), params=np.concatenate([GTR_PARAMS, [shape]])[None, :])[0]
        assert abs(engine.get_log_marginal_likelihood() - want) < 1e-10 * abs(want), site
