"""Site-pattern compression (include/sbn_b200_patterns.h).

CPU: the NumPy restatement (oracle/site_pattern.py) against the UNMODIFIED
reference's SitePattern::Compress on the reference's own alignments
(tests/golden/site_pattern_*.json.gz, made by make_site_pattern_fixtures.py) --
as a multiset of (pattern, weight) pairs, because the reference's order is an
accident of its hash map -- and the reference's unit test of the symbol table.
GPU: the CUDA path through the C ABI against the same fixtures (bit-exact) and
against the restatement on seeded random alignments, edge cases included.
"""
import ctypes
import gzip
import json
import os

import numpy as np
import pytest

from conftest import ROOT
from libsbn_b200 import _capi
from libsbn_b200.site_pattern import SitePattern

GOLDEN = os.path.join(ROOT, "tests", "golden")
FIXTURES = ["hello", "ds1", "flua", "five_taxon", "seven_taxon"]


def load(name):
    with gzip.open(os.path.join(GOLDEN, f"site_pattern_{name}.json.gz"), "rt") as handle:
        return json.load(handle)


def as_multiset(patterns, weights):
    patterns = np.asarray(patterns, dtype=np.uint8)
    return sorted((bytes(patterns[:, k]), float(w)) for k, w in enumerate(weights))


def random_alignment(taxa, sites, seed, distinct):
    """Columns drawn from `distinct` random columns, so patterns repeat."""
    rng = np.random.default_rng(seed)
    alphabet = np.frombuffer(b"ACGTacgt-N?RYKM", dtype=np.uint8)
    pool = alphabet[rng.integers(0, alphabet.size, size=(taxa, distinct))]
    columns = pool[:, rng.integers(0, distinct, size=sites)]
    return [bytes(row) for row in columns]


@pytest.fixture(scope="module")
def restatement():
    from oracle import site_pattern
    return site_pattern


@pytest.mark.parametrize("name", FIXTURES)
def test_restatement_matches_the_reference(restatement, name):
    fx = load(name)
    patterns, weights = restatement.compress(fx["sequences"])
    assert as_multiset(patterns, weights) == as_multiset(fx["patterns"], fx["weights"])
    assert weights.sum() == len(fx["sequences"][0])
    # first-appearance order: pattern k's first site precedes pattern k+1's
    table = restatement.symbol_table()
    symbols = table[np.array([np.frombuffer(s.encode(), np.uint8) for s in fx["sequences"]])]
    first = [int(np.argmax((symbols == patterns[:, [k]]).all(axis=0))) for k in range(patterns.shape[1])]
    assert first == sorted(first)


def test_symbol_table_is_the_reference_s(restatement):
    # site_pattern.hpp:57-62
    table = restatement.symbol_table()
    assert list(table[np.frombuffer(b"-tgcaTGCA?", np.uint8)]) == [4, 3, 2, 1, 0, 3, 2, 1, 0, 4]
    with pytest.raises(RuntimeError, match="Symbol 'Z' not known."):
        restatement.compress(["ACZT", "ACGT"])


def test_no_device_means_failure_not_fallback():
    if _capi.load().sbnb_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(RuntimeError) as info:
        SitePattern(["ACGT", "ACGA"])
    assert info.value.code == -2 and "no CPU fallback" in str(info.value)


@pytest.mark.gpu
@pytest.mark.parametrize("name", FIXTURES)
def test_device_matches_the_reference(restatement, name):
    fx = load(name)
    got = SitePattern(fx["sequences"])
    assert as_multiset(got.patterns, got.weights) == as_multiset(fx["patterns"], fx["weights"])
    want_patterns, want_weights = restatement.compress(fx["sequences"])
    assert np.array_equal(got.patterns, want_patterns) and np.array_equal(got.weights, want_weights)


@pytest.mark.gpu
@pytest.mark.parametrize("taxa,sites,distinct", [(1, 1, 1), (3, 15, 4), (2, 16, 16), (5, 17, 3), (27, 4097, 900),
                                                 (100, 100000, 70000), (7, 250001, 5), (300, 20000, 20000)])
def test_device_matches_the_restatement(restatement, taxa, sites, distinct):
    sequences = random_alignment(taxa, sites, seed=taxa * 7919 + sites, distinct=distinct)
    got = SitePattern(sequences)
    want_patterns, want_weights = restatement.compress(sequences)
    assert got.pattern_count == want_patterns.shape[1]
    assert np.array_equal(got.patterns, want_patterns) and np.array_equal(got.weights, want_weights)
    assert got.weights.sum() == sites
    again = SitePattern(sequences)  # deterministic
    assert np.array_equal(again.patterns, got.patterns) and np.array_equal(again.weights, got.weights)
    # the characters as one [taxon][site] array (no host copy)
    array = SitePattern(np.array([np.frombuffer(s, np.uint8) for s in sequences]))
    assert np.array_equal(array.patterns, got.patterns) and np.array_equal(array.weights, got.weights)


@pytest.mark.gpu
def test_colliding_keys_are_detected_not_merged(monkeypatch):
    """With column keys cut to 6 bits different columns share keys under every seed:
    the byte-for-byte verification must refuse to merge them."""
    sequences = random_alignment(6, 2000, seed=3, distinct=500)
    monkeypatch.setenv("SBNB_DEBUG_PATTERN_KEY_BITS", "6")
    with pytest.raises(RuntimeError, match="collided under four seeds"):
        SitePattern(sequences)
    monkeypatch.setenv("SBNB_DEBUG_PATTERN_KEY_BITS", "40")  # ample: no collision among 500 columns
    assert SitePattern(sequences).pattern_count <= 500


@pytest.mark.gpu
def test_device_edge_cases_and_errors(restatement):
    empty = SitePattern(["", "", ""])
    assert empty.pattern_count == 0 and empty.patterns.shape == (3, 0) and empty.weights.size == 0
    same = SitePattern(["A" * 1000, "c" * 1000])
    assert same.pattern_count == 1 and same.weights[0] == 1000 and list(same.patterns[:, 0]) == [0, 1]
    with pytest.raises(RuntimeError, match="Symbol 'Z' not known."):
        SitePattern(["ACGT" * 10, "ACGT" * 9 + "ACZT"])
    with pytest.raises(RuntimeError, match="same length"):
        SitePattern(["ACGT", "ACG"])
    # the compressed alignment drives the likelihood engine exactly like the host-compressed one
    import libsbn_b200 as sbn
    from libsbn_b200 import trees
    sequences = random_alignment(9, 3000, seed=5, distinct=200)
    pattern = SitePattern(sequences)
    table = restatement.symbol_table()
    raw = table[np.array([np.frombuffer(s, np.uint8) for s in sequences])]
    parent_ids, lengths = trees.random_tree_batch(9, 2, seed=1)
    spec = sbn.PhyloModelSpecification("JC69", "weibull+4", "none")
    params = np.full((2, 1), 0.7)
    compressed = sbn.Engine(spec, pattern.patterns, pattern.weights).log_likelihoods(
        sbn.TreeBatch(parent_ids, lengths), params)
    uncompressed = sbn.Engine(spec, raw, np.ones(raw.shape[1])).log_likelihoods(
        sbn.TreeBatch(parent_ids, lengths), params)
    assert np.max(np.abs(compressed - uncompressed)) <= 1e-11 * np.max(np.abs(uncompressed))
