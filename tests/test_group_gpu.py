"""Device groups inside libsbn_b200.so (sbnb_engine_create_multi): what the
reference's Engine does with thread_count FatBeagles (engine.cpp:17-27), done with one
host thread + stream per GPU.  A device ordinal may repeat in the list, so the host
fan-out, the slicing of inputs / outputs and the cross-device sum are exercised on a
one-GPU box too; with two or more GPUs the same cases run across real devices.
"""
import numpy as np
import pytest

from conftest import load_fixture
import libsbn_b200 as sbn
from libsbn_b200 import _capi, trees

pytestmark = pytest.mark.gpu


def device_lists():
    count = _capi.load().sbnb_device_count()
    lists = [[0, 0], [0, 0, 0]]
    if count >= 2:
        lists.append(list(range(min(count, 8))))
    return lists


def engines(fx, devices, axis):
    spec = sbn.PhyloModelSpecification(fx["substitution"], fx["site"], fx["clock"])
    single = sbn.Engine(spec, fx["patterns"], fx["weights"], 0)
    group = sbn.Engine(spec, fx["patterns"], fx["weights"], devices=devices, shard_axis=axis)
    assert group.device_count == len(devices) and single.device_count == 1
    return single, group


def batch_of(fx):
    if fx["rooted"]:
        return sbn.TreeBatch(fx["parent_ids"], fx["branch_lengths"], fx["rates"], fx["node_heights"],
                             fx["node_bounds"], fx["height_ratios"], 1)
    return sbn.TreeBatch(fx["parent_ids"], fx["branch_lengths"])


def assert_same(got, want, exact):
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert set(g.gradient) == set(w.gradient)
        if exact:
            assert g.log_likelihood == w.log_likelihood
        else:
            assert abs(g.log_likelihood - w.log_likelihood) <= 1e-12 * abs(w.log_likelihood)
        for key in w.gradient:
            a, b = np.asarray(g.gradient[key]), np.asarray(w.gradient[key])
            if exact:
                assert np.array_equal(a, b), key
            else:
                # substitution-model entries are central differences of log-likelihoods: FD noise
                tol = 1e-4 if key == "substitution_model" else 1e-11
                assert np.max(np.abs(a - b)) <= tol * max(np.max(np.abs(b)), 1e-300), key


@pytest.mark.parametrize("devices", device_lists(), ids=lambda d: "x".join(map(str, d)))
@pytest.mark.parametrize("name", ["ds1_gtr_weibull4", "ds1_100_topologies_jc69", "flua_gtr_strict"])
def test_tree_axis_is_the_single_device_result(name, devices):
    """Trees are independent: every device evaluates a slice and the results are the
    single-device ones bit for bit."""
    fx = load_fixture(name)
    single, group = engines(fx, devices, "trees")
    batch = batch_of(fx)
    assert np.array_equal(group.log_likelihoods(batch, fx["params"], True, rooted=fx["rooted"]),
                          single.log_likelihoods(batch, fx["params"], True, rooted=fx["rooted"]))
    assert_same(group.gradients(batch, fx["params"], True, rooted=fx["rooted"]),
                single.gradients(batch, fx["params"], True, rooted=fx["rooted"]), exact=True)


@pytest.mark.parametrize("devices", device_lists(), ids=lambda d: "x".join(map(str, d)))
@pytest.mark.parametrize("name", ["ds1_gtr_weibull4", "ds1_jc69", "flua_jc69_weibull4_strict"])
def test_pattern_axis_sums_to_the_single_device_result(name, devices):
    """Every device walks all trees over its pattern range; the raw sums are added over
    peer memory in a fixed order (only the order of the additions differs)."""
    fx = load_fixture(name)
    single, group = engines(fx, devices, "patterns")
    batch = batch_of(fx)
    got = group.log_likelihoods(batch, fx["params"], False, rooted=fx["rooted"])
    want = single.log_likelihoods(batch, fx["params"], False, rooted=fx["rooted"])
    assert np.max(np.abs(got - want) / np.abs(want)) < 1e-12
    first = group.gradients(batch, fx["params"], True, rooted=fx["rooted"])
    assert_same(first, single.gradients(batch, fx["params"], True, rooted=fx["rooted"]), exact=False)
    # fixed order of the cross-device additions: bitwise reproducible
    assert_same(group.gradients(batch, fx["params"], True, rooted=fx["rooted"]), first, exact=True)


def test_more_devices_than_trees_and_empty_collections():
    fx = load_fixture("hello_jc69")
    single, group = engines(fx, [0, 0, 0], "trees")
    one = sbn.TreeBatch(fx["parent_ids"][:1], fx["branch_lengths"][:1])
    assert np.array_equal(group.log_likelihoods(one, fx["params"][:1]), single.log_likelihoods(one, fx["params"][:1]))
    empty = sbn.TreeBatch(fx["parent_ids"][:0], fx["branch_lengths"][:0])
    for engine in (single, group):
        assert engine.log_likelihoods(empty, fx["params"][:0]).shape == (0,)
        assert engine.gradients(empty, fx["params"][:0]) == []


def test_group_rejects_the_staged_entry_points():
    fx = load_fixture("hello_jc69")
    _, group = engines(fx, [0, 0], "trees")
    with pytest.raises(RuntimeError, match="one device"):
        group.stage(batch_of(fx), fx["params"])
    with pytest.raises(RuntimeError, match="Unknown shard axis|shard_axis"):
        sbn.Engine(sbn.PhyloModelSpecification("JC69", "constant", "none"), fx["patterns"], fx["weights"],
                   devices=[0], shard_axis="sites")


def test_config4_shape_over_a_group():
    """100 taxa x 20k patterns x 16 random trees, GTR + 4 categories: both axes against
    the single-device result."""
    taxa, patterns, tree_count = 100, 20000, 16
    states, weights = trees.random_alignment(taxa, patterns, seed=3)
    parent_ids, lengths = trees.random_tree_batch(taxa, tree_count, seed=4)
    params = np.tile(np.array([0.05, 0.1, 0.15, 0.20, 0.25, 0.25, 0.1, 0.2, 0.3, 0.4, 0.5]), (tree_count, 1))
    spec = sbn.PhyloModelSpecification("GTR", "weibull+4", "none")
    batch = sbn.TreeBatch(parent_ids, lengths)
    want = sbn.Engine(spec, states, weights).gradients(batch, params, True, substitution_gradient=False)
    devices = device_lists()[-1]
    by_tree = sbn.Engine(spec, states, weights, devices=devices, shard_axis="trees")
    # (how a tree's patterns are cut into chunks depends on how many trees a launch holds,
    #  so at this size only the order of the additions differs from the single device)
    assert_same(by_tree.gradients(batch, params, True, substitution_gradient=False), want, exact=False)
    by_pattern = sbn.Engine(spec, states, weights, devices=devices, shard_axis="patterns")
    assert_same(by_pattern.gradients(batch, params, True, substitution_gradient=False), want, exact=False)
