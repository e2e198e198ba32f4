"""Sharding on the GPU: the pattern-range arithmetic on one device, and (when
the box has >= 2 GPUs) the real one-process-per-GPU NCCL path via torchrun."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, load_fixture
import libsbn_b200 as sbn
from libsbn_b200 import _capi, sharding, trees

pytestmark = pytest.mark.gpu

GTR_ROW = [0.05, 0.1, 0.15, 0.20, 0.25, 0.25, 0.1, 0.2, 0.3, 0.4, 0.5]


def stack(gradients, key):
    return np.array([g.gradient[key] for g in gradients])


def raw_sums_over_ranges(engine, batch, params, ranges, rooted, fd):
    """What G ranks would all-reduce: raw results per pattern range, summed."""
    total = None
    for begin, end in ranges:
        engine.set_pattern_range(begin, end)
        staged = engine.stage(batch, params, rooted=rooted, substitution_fd=fd)
        staged.run(_capi.MODE_BRANCH_GRADIENT, True)
        parts = staged.fetch(gradients=True)
        staged.close()
        total = [p.copy() for p in parts] if total is None else [t + p for t, p in zip(total, parts)]
    engine.set_pattern_range(0, engine.pattern_count)
    return total


@pytest.mark.parametrize("world", [2, 3, 8])
def test_pattern_ranges_sum_to_the_whole_unrooted(world):
    taxa, patterns, tree_count = 14, 4000, 6
    states, weights = trees.random_alignment(taxa, patterns, seed=3, gap_fraction=0.02)
    parent_ids, lengths = trees.random_tree_batch(taxa, tree_count, seed=4)
    params = np.tile(np.array(GTR_ROW), (tree_count, 1))
    spec = sbn.PhyloModelSpecification("GTR", "weibull+4", "none")
    engine = sbn.Engine(spec, states, weights, 0)
    batch = sbn.TreeBatch(parent_ids, lengths)
    engine.set_substitution_gradient("fd")  # this test is about the finite-difference evaluations
    want = engine.gradients(batch, params, rescaling=True)
    ranges = [sharding.pattern_range(r, world, patterns) for r in range(world)]
    logl, grad, rgrad = raw_sums_over_ranges(engine, batch, params, ranges, False, True)
    got = sharding.finish_gradients(spec, taxa, batch, False, True, logl, grad, rgrad, engine.category_count)
    np.testing.assert_allclose([g.log_likelihood for g in got], [g.log_likelihood for g in want], rtol=1e-12)
    for key in ("branch_lengths", "site_model"):
        a, b = stack(got, key), stack(want, key)
        assert np.max(np.abs(a - b)) <= 1e-11 * np.max(np.abs(b)), key
    # central differences of log-likelihoods ~1e4 with delta 1e-6 amplify the ~1e-13
    # relative reordering noise of the sums by 1e6 / 2
    a, b = stack(got, "substitution_model"), stack(want, "substitution_model")
    assert np.max(np.abs(a - b)) <= np.abs(logl).max() * 1e-12 / 1e-6


@pytest.mark.parametrize("substitution", ["GTR", "HKY"])
def test_analytic_substitution_sums_add_up_over_pattern_ranges(substitution):
    """The 20 sums per tree of the analytic substitution gradient are sums over site
    patterns like the other raw results: ranks add them and finish once."""
    taxa, patterns, tree_count, world = 14, 4000, 6, 3
    states, weights = trees.random_alignment(taxa, patterns, seed=3, gap_fraction=0.02)
    parent_ids, lengths = trees.random_tree_batch(taxa, tree_count, seed=4)
    row = GTR_ROW if substitution == "GTR" else [0.1, 0.2, 0.3, 0.4, 2.0, 0.5]
    params = np.tile(np.array(row), (tree_count, 1))
    spec = sbn.PhyloModelSpecification(substitution, "weibull+4", "none")
    engine = sbn.Engine(spec, states, weights, 0)
    batch = sbn.TreeBatch(parent_ids, lengths)
    want = engine.gradients(batch, params, rescaling=True)
    total = None
    for rank in range(world):
        engine.set_pattern_range(*sharding.pattern_range(rank, world, patterns))
        staged = engine.stage(batch, params, substitution_analytic=True)
        staged.run(_capi.MODE_BRANCH_GRADIENT, True)
        parts = list(staged.fetch(gradients=True)) + [staged.fetch_substitution_sums()]
        staged.close()
        total = parts if total is None else [t + p for t, p in zip(total, parts)]
    engine.set_pattern_range(0, patterns)
    got = sharding.finish_gradients_analytic(spec, taxa, batch, False, params, *total, engine.category_count)
    np.testing.assert_allclose([g.log_likelihood for g in got], [g.log_likelihood for g in want], rtol=1e-12)
    for key in ("branch_lengths", "site_model", "substitution_model"):
        a, b = stack(got, key), stack(want, key)
        assert np.max(np.abs(a - b)) <= 1e-10 * np.max(np.abs(b)), key


def test_pattern_ranges_sum_to_the_whole_rooted():
    fx = load_fixture("flua_jc69_weibull4_strict")
    spec = sbn.PhyloModelSpecification(fx["substitution"], fx["site"], fx["clock"])
    engine = sbn.Engine(spec, fx["patterns"], fx["weights"], 0)
    batch = sbn.TreeBatch(fx["parent_ids"], fx["branch_lengths"], fx["rates"], fx["node_heights"],
                          fx["node_bounds"], fx["height_ratios"], 1)
    params = fx["params"]
    want = engine.gradients(batch, params, rescaling=False, rooted=True)
    patterns = engine.pattern_count
    ranges = [sharding.pattern_range(r, 4, patterns) for r in range(4)]
    logl, grad, rgrad = raw_sums_over_ranges(engine, batch, params, ranges, True, False)
    got = sharding.finish_gradients(spec, engine.taxon_count, batch, True, False, logl, grad, rgrad,
                                    engine.category_count)
    for key in ("ratios_root_height", "clock_model", "site_model"):
        a, b = stack(got, key), stack(want, key)
        assert np.max(np.abs(a - b)) <= 1e-10 * np.max(np.abs(b)), key
    # the log-det Jacobian is added once, after the reduction
    want_logl = engine.log_likelihoods(batch, params, rooted=True)
    finished = sharding.finish_log_likelihoods_rooted(engine.taxon_count, batch, logl[:batch.tree_count].copy())
    np.testing.assert_allclose(finished, want_logl, rtol=1e-12)


def test_world_of_one_is_the_plain_engine():
    taxa, patterns, tree_count = 10, 700, 3
    states, weights = trees.random_alignment(taxa, patterns, seed=5, gap_fraction=0.02)
    parent_ids, lengths = trees.random_tree_batch(taxa, tree_count, seed=6)
    spec = sbn.PhyloModelSpecification("JC69", "constant", "none")
    batch = sbn.TreeBatch(parent_ids, lengths)
    plain = sbn.Engine(spec, states, weights, 0).gradients(batch, None, rescaling=True)
    for cls in (sharding.PatternShardedEngine, sharding.TreeShardedEngine):
        got = cls(spec, states, weights, 0).gradients(batch, None, rescaling=True)
        assert np.array_equal(stack(got, "branch_lengths"), stack(plain, "branch_lengths"))
        assert [g.log_likelihood for g in got] == [g.log_likelihood for g in plain]


def test_nccl_two_ranks():
    """One process per GPU over NCCL (skipped on a 1-GPU box; bench.py --gpus N
    exercises the same classes)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    result = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
         "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "tests", "nccl_worker.py")],
        capture_output=True, text=True, timeout=600)
    assert result.returncode == 0, result.stdout[-3000:] + result.stderr[-3000:]
    assert "NCCL-SHARDING-OK" in result.stdout
