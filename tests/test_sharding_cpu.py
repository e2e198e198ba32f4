"""The N > 1 host logic on CPU: world_size-2 `gloo` process groups.

The CUDA engine cannot run here, so the per-rank compute is stood in for by the
CPU oracle (tests may use it); what is under test is libsbn_b200.sharding's
partitioning, collectives and host finishing (sbnb_finish_gradients), i.e.
everything between "this rank's raw sums" and "the PhyloGradients every rank
returns".  The same classes run over NCCL in tests/test_sharding_gpu.py.
"""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

import libsbn_b200 as sbn
from libsbn_b200 import sharding, trees


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def run_world(worker, world, *args):
    port = free_port()
    mp.spawn(_entry, args=(world, port, worker, args), nprocs=world, join=True)


def _entry(rank, world, port, worker, args):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        worker(rank, world, *args)
    finally:
        dist.destroy_process_group()


def test_shard_ranges_cover_and_balance():
    for count in (0, 1, 7, 1024, 100000):
        for world in (1, 2, 3, 8):
            ranges = [sharding.shard_range(r, world, count) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == count
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            sizes = [e - b for b, e in ranges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(RuntimeError):
        sharding.pattern_range(0, 8, 5)
    with pytest.raises(RuntimeError):
        sharding.shard_range(2, 2, 10)


def _workload(taxa=9, patterns=301, tree_count=5):
    states, weights = trees.random_alignment(taxa, patterns, seed=11, gap_fraction=0.05)
    weights = weights * np.arange(1, patterns + 1) % 4 + 1.0
    parent_ids, lengths = trees.random_tree_batch(taxa, tree_count, seed=12)
    return states, weights, parent_ids, lengths


def _pattern_worker(rank, world):
    from oracle import phylo
    states, weights, parent_ids, lengths = _workload()
    shape = np.full((len(parent_ids), 1), 0.7)
    want = phylo.gradients("JC69", "weibull+4", states, weights, parent_ids, lengths, shape, rescaling=True)
    begin, end = sharding.pattern_range(rank, world, states.shape[1])
    local = phylo.gradients("JC69", "weibull+4", states[:, begin:end], weights[begin:end], parent_ids, lengths,
                            shape, rescaling=True)
    logl, grad = local["log_likelihood"].copy(), local["branch"].copy()
    assert not np.allclose(logl, want["log_likelihood"])  # a shard alone is not the answer
    sharding.all_reduce_sum_host(logl, grad)
    np.testing.assert_allclose(logl, want["log_likelihood"], rtol=1e-12)
    # host finishing over the reduced sums (the site-model block is checked on the GPU,
    # where raw rate gradients exist; here the rate gradient is zero)
    spec = sbn.PhyloModelSpecification("JC69", "weibull+4", "none")
    finished = sharding.finish_gradients(spec, states.shape[0], sbn.TreeBatch(parent_ids, lengths), False, False,
                                         logl, grad, np.zeros_like(grad), 4)
    got = np.array([g.gradient["branch_lengths"] for g in finished])
    scale = np.abs(want["branch"]).max(axis=1, keepdims=True)
    assert np.max(np.abs(got - want["branch"]) / scale) < 1e-12
    assert all(g.gradient["site_model"][0] == 0.0 for g in finished)
    np.testing.assert_allclose([g.log_likelihood for g in finished], want["log_likelihood"], rtol=1e-12)


def test_pattern_sharding_world_2():
    run_world(_pattern_worker, 2)


def test_pattern_sharding_world_3():
    run_world(_pattern_worker, 3)


def _tree_worker(rank, world):
    from oracle import phylo
    states, weights, parent_ids, lengths = _workload(tree_count=5)
    want = phylo.gradients("JC69", "constant", states, weights, parent_ids, lengths, None, rescaling=False)
    begin, end = sharding.shard_range(rank, world, len(parent_ids))
    local = phylo.gradients("JC69", "constant", states, weights, parent_ids[begin:end], lengths[begin:end], None)
    logl = sharding.all_gather_rows(local["log_likelihood"][:, None], len(parent_ids))[:, 0]
    assert np.array_equal(logl, want["log_likelihood"])  # a gather moves bits, it does not round
    n = states.shape[0]
    packed = [sbn.PhyloGradient(float(l), {"branch_lengths": b}) for l, b in zip(local["log_likelihood"],
                                                                               local["branch"])]
    gathered = sharding.gather_gradients(packed, len(parent_ids), [("branch_lengths", 2 * n - 1)])
    assert len(gathered) == len(parent_ids)
    assert np.array_equal(np.array([g.gradient["branch_lengths"] for g in gathered]), want["branch"])


def test_tree_sharding_world_2():
    run_world(_tree_worker, 2)


def test_single_process_is_a_no_op():
    a = np.arange(6, dtype=np.float64)
    sharding.all_reduce_sum_host(a)
    assert np.array_equal(a, np.arange(6))
    assert sharding.all_gather_rows(a[:, None], 6).shape == (6, 1)


def test_finish_rooted_log_likelihoods_adds_the_jacobian():
    from conftest import load_fixture
    fx = load_fixture("flua_jc69_strict")
    batch = sbn.TreeBatch(fx["parent_ids"], fx["branch_lengths"], fx["rates"], fx["node_heights"],
                          fx["node_bounds"], fx["height_ratios"], 1)
    zero = np.zeros(batch.tree_count)
    jacobian = sharding.finish_log_likelihoods_rooted(fx["patterns"].shape[0], batch, zero.copy())
    assert np.all(np.isfinite(jacobian)) and np.all(jacobian != 0.0)
    # additive: finishing twice adds it twice
    twice = sharding.finish_log_likelihoods_rooted(fx["patterns"].shape[0], batch, jacobian.copy())
    np.testing.assert_allclose(twice, 2 * jacobian, rtol=1e-15)
