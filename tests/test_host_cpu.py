"""CPU-only tests of the host side of libsbn_b200 (no compute call needs a GPU):

  * the C-ABI library loads and exports every symbol include/sbn_b200.h declares;
  * without a device the compute entry points fail loudly (no CPU fallback);
  * the model tables (substitution eigen-systems, Weibull rates) against the
    reference's own unit-test values (substitution_model.hpp:97-131,
    site_model.hpp:84-108) and against scipy's matrix exponential;
  * the traversal programs: a NumPy interpreter executes exactly the op lists
    the kernel would execute (same slots, same formulas) and must reproduce
    the oracle's log-likelihoods and branch gradients.
"""
import ctypes
import os
import re

import numpy as np
import pytest
from scipy.linalg import expm

from conftest import ROOT, load_fixture
from libsbn_b200 import _capi, trees
import libsbn_b200


def test_library_exports_every_declared_symbol():
    declared = set()
    for name in ("sbn_b200.h", "sbn_b200_gp.h", "sbn_b200_patterns.h"):
        header = open(os.path.join(ROOT, "include", name)).read()
        declared |= set(re.findall(r"\b(sbnb_[a-z_0-9]+)\s*\(", header))
    assert len(declared) >= 45
    lib = ctypes.CDLL(_capi.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/*.h but not exported"
    assert declared == set(_capi.SIGNATURES), "python binding and header disagree"


def test_no_device_means_failure_not_fallback():
    lib = _capi.load()
    if lib.sbnb_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(RuntimeError) as info:
        libsbn_b200.Engine(libsbn_b200.PhyloModelSpecification(), np.zeros((3, 5), np.uint8), np.ones(5))
    assert info.value.code == -2  # SBNB_ERR_NO_DEVICE
    assert "no CPU fallback" in str(info.value)


def test_gp_engine_without_device_fails_loudly():
    lib = _capi.load()
    if lib.sbnb_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(RuntimeError) as info:
        libsbn_b200.GPEngine(np.zeros((3, 5), np.uint8), np.ones(5), 5, 30, 5)
    assert info.value.code == -2  # SBNB_ERR_NO_DEVICE
    assert "no CPU fallback" in str(info.value)


def test_gp_operation_encoding_matches_the_reference_dump():
    """GPOperations.* encode the records oracle/gp_dump.cpp flattened from the
    reference's own GPOperationVector (hello DAG, gp_dag.cpp:255-263)."""
    from conftest import load_fixture
    ops = libsbn_b200.GPOperations
    fx = load_fixture("gp_hello")
    assert np.array_equal(ops.program([ops.ResetMarginalLikelihood(), ops.IncrementMarginalLikelihood(19, 0, 4)]),
                          fx["program_marginal_likelihood"])
    assert np.array_equal(ops.program([ops.UpdateSBNProbabilities(0, 1)]), fx["program_optimize_sbn_parameters"])
    head = ops.program([ops.ZeroPLV(3), ops.ZeroPLV(8)])
    assert np.array_equal(head, fx["program_populate_plvs"][:4])
    assert ops.PrepForMarginalization(9, [8, 7]) == [9, 9, 2, 8, 7]


def test_unknown_models_are_rejected_like_the_reference():
    lib = _capi.load()
    out = (ctypes.c_double * 16)()
    for spec, message in [((b"K80", b"constant", b"none"), "Substitution model not known"),
                          ((b"JC69", b"invariant", b"none"), "Site model not known"),
                          ((b"JC69", b"gammaX", b"none"), "Site model not known"),
                          ((b"JC69", b"constant", b"relaxed"), "Clock model not known")]:
        code = lib.sbnb_debug_model_tables(*spec, None, out, None, None, None, None, None, None, None)
        assert code == -1 and message in lib.sbnb_last_error().decode()


def model_tables(substitution, site, row):
    lib = _capi.load()
    categories = 1
    if site.startswith(("weibull", "gamma")):
        categories = int(site.split("+")[1]) if "+" in site else 4
    t = {k: np.zeros(s) for k, s in [("evec", 16), ("ivec", 16), ("eval", 4), ("freqs", 4), ("q", 16),
                                     ("rates", categories), ("weights", categories), ("drates", categories)]}
    row = np.ascontiguousarray(row, dtype=np.float64)
    _capi.check(lib.sbnb_debug_model_tables(substitution.encode(), site.encode(), b"none",
                                            _capi.as_double_ptr(row) if row.size else None,
                                            *[_capi.as_double_ptr(t[k]) for k in
                                              ("evec", "ivec", "eval", "freqs", "q", "rates", "weights", "drates")]))
    for k in ("evec", "ivec", "q"):
        t[k] = t[k].reshape(4, 4)
    return t


def test_gtr_eigensystem():
    # substitution_model.hpp:120-131: eigenvalues from R.
    rates = [0.060602, 0.402732, 0.028230, 0.047910, 0.407249, 0.053277]
    freqs = [0.479367, 0.172572, 0.140933, 0.207128]
    t = model_tables("GTR", "constant", rates + freqs)
    assert np.allclose(np.sort(t["eval"]), [-2.567992e+00, -1.760838e+00, -4.214918e-01, 0.0], atol=1e-4)
    q = t["q"]
    assert np.allclose(t["evec"] @ t["ivec"], np.eye(4), atol=1e-14)
    assert np.allclose(t["evec"] @ np.diag(t["eval"]) @ t["ivec"], q, atol=1e-14)
    assert np.allclose(q.sum(axis=1), 0, atol=1e-15)
    assert np.allclose(np.array(freqs) @ q, 0, atol=1e-15)
    assert abs(-(np.diag(q) * freqs).sum() - 1) < 1e-14  # unit expected rate
    for length in (1e-6, 0.03, 0.7, 5.0):
        p = t["evec"] @ np.diag(np.exp(t["eval"] * length)) @ t["ivec"]
        assert np.allclose(p, expm(q * length), atol=1e-14)


def test_jc69_equals_gtr_with_equal_parameters():
    jc = model_tables("JC69", "constant", [])
    gtr = model_tables("GTR", "constant", [1 / 6] * 6 + [0.25] * 4)
    assert np.allclose(jc["q"], gtr["q"], atol=1e-15)
    for length in (0.01, 0.75):
        pj = jc["evec"] @ np.diag(np.exp(jc["eval"] * length)) @ jc["ivec"]
        pg = gtr["evec"] @ np.diag(np.exp(gtr["eval"] * length)) @ gtr["ivec"]
        assert np.allclose(pj, pg, atol=1e-15)


def test_hky_is_gtr_with_kappa_on_transitions():
    freqs = [0.1, 0.2, 0.3, 0.4]
    hky = model_tables("HKY", "constant", freqs + [2.0])  # blocks: frequencies, kappa
    raw = np.array([1, 2, 1, 1, 2, 1.0])
    gtr = model_tables("GTR", "constant", list(raw / raw.sum()) + freqs)
    assert np.allclose(hky["q"], gtr["q"], atol=1e-15)


def test_gamma_rates_match_scipy():
    """Discrete Gamma is an addition (the reference has Weibull only): pinned to
    scipy.stats.gamma.ppf, median discretisation, rates normalised to mean 1."""
    from oracle import phylo
    for shape in (0.05, 0.1, 0.5, 1.0, 2.7, 10.0, 150.0):
        for categories in (1, 2, 4, 6, 16):
            t = model_tables("JC69", f"gamma+{categories}", [shape])
            rates, derivative = phylo.gamma_rates(shape, categories)
            assert np.allclose(t["rates"], rates, rtol=1e-11, atol=1e-300), (shape, categories)
            assert np.allclose(t["weights"], 1 / categories)
            assert abs(t["rates"] @ t["weights"] - 1) < 1e-14
            # the mean rate is pinned to 1, so the derivatives sum to zero
            assert abs(t["drates"].sum()) <= 1e-13 * categories * max(1.0, np.abs(t["drates"]).max()), (shape, categories)
            if shape >= 0.1:  # below, differencing scipy's ppf is the less accurate side
                scale = np.abs(derivative).max() + 1e-300
                assert np.abs(t["drates"] - derivative).max() <= 1e-6 * scale, (shape, categories)
    assert np.allclose(model_tables("GTR", "gamma", [1 / 6] * 6 + [0.25] * 4 + [0.5])["rates"],
                       phylo.gamma_rates(0.5, 4)[0], rtol=1e-11)
    with pytest.raises(_capi.SbnbError, match="Gamma shape must be positive"):
        model_tables("JC69", "gamma+4", [0.0])


def test_weibull_rates():
    t = model_tables("JC69", "weibull+4", [1.0])  # site_model.hpp:86-90
    assert np.allclose(t["rates"], [0.1457844, 0.5131316, 1.0708310, 2.2702530], atol=1e-4)
    t = model_tables("JC69", "weibull+4", [0.1])  # site_model.hpp:93-99
    assert np.allclose(t["rates"], [4.766392e-12, 1.391131e-06, 2.179165e-03, 3.997819e+00], atol=1e-4)
    assert np.allclose(t["weights"], 0.25)
    assert abs(t["rates"] @ t["weights"] - 1) < 1e-14
    # analytic d rate / d shape against central differences
    for shape in (0.1, 0.5, 2.0):
        h = 1e-6 * shape
        up, down = model_tables("JC69", "weibull+6", [shape + h]), model_tables("JC69", "weibull+6", [shape - h])
        numeric = (up["rates"] - down["rates"]) / (2 * h)
        assert np.allclose(model_tables("JC69", "weibull+6", [shape])["drates"], numeric, rtol=1e-6, atol=1e-9)


def test_parameter_sanity_failures():
    lib = _capi.load()
    bad = np.array([1 / 6] * 6 + [0.3, 0.3, 0.3, 0.3])
    out = np.zeros(16)
    code = lib.sbnb_debug_model_tables(b"GTR", b"constant", b"none", _capi.as_double_ptr(bad),
                                       _capi.as_double_ptr(out), None, None, None, None, None, None, None)
    assert code == -5 and "do not sum to 1" in lib.sbnb_last_error().decode()


# ---------------------------------------------------------------------------
# traversal programs

def tree_program(parent_ids, taxon_count):
    lib = _capi.load()
    parent_ids = np.ascontiguousarray(parent_ids, dtype=np.int32)
    post = np.zeros((taxon_count - 1, 8), dtype=np.int32)
    pre = np.zeros((taxon_count - 1, 8), dtype=np.int32)
    slots = np.zeros(2, dtype=np.int32)
    _capi.check(lib.sbnb_debug_tree_program(_capi.as_int32_ptr(parent_ids), len(parent_ids) + 1, taxon_count,
                                            _capi.as_int32_ptr(post), _capi.as_int32_ptr(pre),
                                            _capi.as_int32_ptr(slots)))
    return post, pre, slots


def interpret(parent_ids, lengths, patterns, weights, tables):
    """Executes the kernel's op lists with NumPy (all patterns at once): the
    same cur/stack discipline, the same formulas as TreeWalkKernel, no rescaling."""
    n, P = patterns.shape
    C = len(tables["rates"])
    post, pre, slots = tree_program(parent_ids, n)
    N = 2 * n - 1
    lengths = np.asarray(lengths, dtype=np.float64)
    if len(lengths) == N - 1:  # detrifurcate
        lengths = np.concatenate([lengths[:-1], [0.0, 0.0]])
    mats = np.zeros((N - 1, C, 4, 4))
    for e in range(N - 1):
        for c in range(C):
            mats[e, c] = tables["evec"] @ np.diag(np.exp(tables["eval"] * lengths[e] * tables["rates"][c])) @ tables["ivec"]
    one_hot = np.vstack([np.eye(4), np.ones((1, 4))])  # state 4 = gap

    def tip(taxon):  # [C][P][4]
        return np.broadcast_to(one_hot[np.minimum(patterns[taxon], 4)], (C, P, 4))

    # ---- post-order: one partial ("cur") in registers, the rest on a stack
    stack = [None] * int(max(slots))
    live = set()
    cur = None  # (node, partial)
    evolved = {}  # what the kernel streams to its scratch arena: P_x L_x of internal x
    log_likelihood = None
    seen = set()
    for node, a, b, push_slot, a_src, b_src, flags, _ in post:
        if push_slot >= 0:
            assert (flags & 3) == 3, "only a cherry may push: it must not read cur"
            assert cur is not None and push_slot not in live
            stack[push_slot] = cur
            live.add(push_slot)
            cur = None

        def child(idx, src, leaf):
            nonlocal cur
            if leaf:
                assert src == -1
                return tip(idx)
            if src == -2:
                assert cur is not None and cur[0] == idx, "child partial is not the previous result"
                out = cur[1]
                cur = None
                return out
            assert src in live and stack[src][0] == idx, "child partial not live in its slot"
            live.discard(src)
            return stack[src][1]
        la, lb = child(a, a_src, flags & 1), child(b, b_src, flags & 2)
        assert cur is None, "cur must have been consumed or pushed"
        ya = np.einsum("cij,ckj->cki", mats[a], la)
        yb = np.einsum("cij,ckj->cki", mats[b], lb)
        if not flags & 1:
            evolved[a] = ya
        if not flags & 2:
            evolved[b] = yb
        out = ya * yb
        seen.add(node)
        cur = (node, out)
        if flags & 4:
            site = np.einsum("c,cki,i->k", tables["weights"], out, tables["freqs"])
            log_likelihood = float(weights @ np.log(site))
    assert len(seen) == n - 1 and log_likelihood is not None and not live and cur[0] == 2 * n - 2

    # ---- pre-order, fused with the edge derivatives
    gradient = np.zeros(N)
    live = set()
    cur = None
    for node, a, b, pop_slot, a_dst, b_dst, flags, _ in pre:
        if flags & 4:
            pp = np.broadcast_to(tables["freqs"], (C, P, 4))
        elif pop_slot >= 0:
            assert cur is None and pop_slot in live and stack[pop_slot][0] == node
            pp = stack[pop_slot][1]
            live.discard(pop_slot)
        else:
            assert cur is not None and cur[0] == node, "pre-order partial is not the forwarded one"
            pp = cur[1]
        cur = None
        ya = np.einsum("cij,ckj->cki", mats[a], tip(a)) if flags & 1 else evolved[a]
        yb = np.einsum("cij,ckj->cki", mats[b], tip(b)) if flags & 2 else evolved[b]
        ta, tb = pp * yb, pp * ya
        den = np.einsum("c,cki,cki->k", tables["weights"], ta, ya)
        for child_id, t, y, dst, leaf in ((a, ta, ya, a_dst, flags & 1), (b, tb, yb, b_dst, flags & 2)):
            num = np.einsum("c,cki,ij,ckj->k", tables["weights"] * tables["rates"], t, tables["q"], y)
            gradient[child_id] = weights @ (num / den)
            if leaf:
                assert dst == -1
                continue
            pre_c = np.einsum("cij,cki->ckj", mats[child_id], t)
            if dst == -2:
                assert cur is None, "only one child may stay in cur"
                cur = (child_id, pre_c)
            else:
                assert dst >= 0 and dst not in live
                stack[dst] = (child_id, pre_c)
                live.add(dst)
    assert not live and cur is None
    return log_likelihood, gradient, slots


@pytest.mark.parametrize("taxa,seed", [(3, 0), (4, 1), (5, 2), (9, 3), (27, 4), (64, 5)])
def test_programs_reproduce_the_oracle(oracle, taxa, seed):
    rng = np.random.default_rng(seed)
    patterns, weights = trees.random_alignment(taxa, 40, seed, gap_fraction=0.05)
    weights = rng.integers(1, 5, size=40).astype(np.float64)
    rates = rng.dirichlet(np.ones(6))
    freqs = rng.dirichlet(np.ones(4) * 5)
    row = np.concatenate([rates, freqs, [0.7]])
    tables = model_tables("GTR", "weibull+3", row)
    for rooted in (False, True):
        parent_ids, lengths = trees.random_tree_batch(taxa, 3, seed, rooted=rooted)
        if rooted and taxa < 3:
            continue
        want = oracle.gradients("GTR", "weibull+3", patterns, weights, parent_ids, lengths,
                                np.tile(row, (3, 1)))
        for t in range(3):
            got_ll, got_grad, slots = interpret(parent_ids[t], lengths[t], patterns, weights, tables)
            assert abs(got_ll - want["log_likelihood"][t]) < 1e-10 * abs(want["log_likelihood"][t])
            expect = want["branch"][t].copy()
            if not rooted:
                got_grad[2 * taxa - 3] = 0.0  # the oracle zeroes the fixed node (fat_beagle.cpp:498-500)
            else:
                # Gradient(UnrootedTree) semantics are applied by the oracle to any tree it is given
                fixed = np.flatnonzero(expect == 0.0)
                got_grad[fixed] = 0.0
            assert np.max(np.abs(got_grad - expect)) < 1e-9 * np.max(np.abs(expect))
            assert slots.max() <= int(np.floor(np.log2(taxa)))


def test_stack_depth_is_logarithmic():
    """Strahler ordering: ladder trees need at most one stack slot, random trees
    of 1000 taxa at most floor(log2 n)."""
    for taxa in (10, 100, 1000):
        _, _, slots = tree_program(trees.ladder_topology(taxa), taxa)
        assert slots.max() <= 1
        rng = np.random.default_rng(taxa)
        for _ in range(5):
            _, _, slots = tree_program(trees.random_unrooted_topology(taxa, rng), taxa)
            assert 1 <= slots.max() <= int(np.floor(np.log2(taxa)))


def test_malformed_topologies_are_rejected():
    lib = _capi.load()
    scratch = np.zeros((8, 8), dtype=np.int32)
    slots = np.zeros(2, dtype=np.int32)
    for parent_ids, taxa in [([3, 3, 3, 4], 3),       # node count fits neither 2n-1 nor 2n-2
                             ([4, 4, 3, 4], 3),       # internal node 3 has one child
                             ([2, 3, 3], 3)]:         # a leaf as parent
        ids = np.array(parent_ids, dtype=np.int32)
        code = lib.sbnb_debug_tree_program(_capi.as_int32_ptr(ids), len(ids) + 1, taxa,
                                           _capi.as_int32_ptr(scratch), _capi.as_int32_ptr(scratch),
                                           _capi.as_int32_ptr(slots))
        assert code == -1, parent_ids


def test_fixture_topologies_build():
    fx = load_fixture("ds1_100_topologies_jc69")
    for ids in fx["parent_ids"]:
        post, pre, slots = tree_program(ids, 27)
        assert sorted(post[:, 0]) == list(range(27, 53)) and slots.max() <= 5


def test_fasta_parsing_and_symbol_table(tmp_path):
    from libsbn_b200 import alignment
    path = tmp_path / "a.fasta"
    path.write_text(">x\nACGTAC-N\n>y\nACGT\nACGT\n>z\naCGTAC?T\n")
    sequences = alignment.read_fasta(str(path))
    assert sequences == {"x": "ACGTAC-N", "y": "ACGTACGT", "z": "aCGTAC?T"}
    states = alignment.encode(sequences, ["z", "x", "y"])
    assert states.shape == (3, 8) and list(states[0]) == [0, 1, 2, 3, 0, 1, 4, 3] and states[1, 7] == 4
    assert alignment.sequences_in_leaf_order(sequences, ["y", "x"]) == ["ACGTACGT", "ACGTAC-N"]
    with pytest.raises(RuntimeError, match="Taxon w not found"):
        alignment.sequences_in_leaf_order(sequences, ["w"])
    with pytest.raises(RuntimeError):
        alignment.encode({"x": "ACZT"}, ["x"])


def _stick_breaking(y, size):
    x, stick = np.zeros(size), 1.0
    for k in range(size - 1):
        z = 1.0 / (1.0 + np.exp(-(y[k] - np.log(size - k - 1))))
        x[k] = stick * z
        stick -= x[k]
    x[-1] = stick
    return x


def _stick_breaking_inverse(x):
    size, y, total = len(x), np.zeros(len(x) - 1), 0.0
    for k in range(size - 1):
        z = x[k] / (1.0 - total)
        y[k] = np.log(z / (1.0 - z)) + np.log(size - k - 1)
        total += x[k]
    return y


@pytest.mark.parametrize("substitution", ["GTR", "HKY"])
def test_analytic_substitution_derivatives_match_central_differences_of_p(substitution):
    """Host part of the analytic substitution gradient: with
    B = V^-1 dQ/dtheta V from sbnb_debug_substitution_derivatives,
    d P(tau) / d theta = V (B o Phi(tau)) V^-1 must equal the central difference of P over the
    reference's own perturbation (stick-breaking coordinates +/- delta, fat_beagle.cpp:400-465)."""
    lib = _capi.load()
    rates = np.array([0.05, 0.1, 0.15, 0.20, 0.25, 0.25])
    freqs = np.array([0.1, 0.2, 0.3, 0.4])
    kappa = 2.3
    row = np.concatenate([rates, freqs]) if substitution == "GTR" else np.concatenate([freqs, [kappa]])
    b, dfreqs, count = np.zeros(8 * 16), np.zeros(8 * 4), ctypes.c_int32()
    _capi.check(lib.sbnb_debug_substitution_derivatives(substitution.encode(), b"constant", b"none",
                                                        _capi.as_double_ptr(row), _capi.as_double_ptr(b),
                                                        _capi.as_double_ptr(dfreqs), ctypes.byref(count)))
    assert count.value == (8 if substitution == "GTR" else 4)
    b, dfreqs = b.reshape(8, 4, 4), dfreqs.reshape(8, 4)
    base = model_tables(substitution, "constant", row)

    def perturbed(index, step):
        """The row with gradient coordinate `index` moved by `step`."""
        r, f, k = rates.copy(), freqs.copy(), kappa
        if substitution == "GTR" and index < 5:
            y = _stick_breaking_inverse(r)
            y[index] += step
            r = _stick_breaking(y, 6)
        elif substitution == "HKY" and index == 0:
            k += step
        else:
            y = _stick_breaking_inverse(f)
            y[index - (5 if substitution == "GTR" else 1)] += step
            f = _stick_breaking(y, 4)
        return np.concatenate([r, f]) if substitution == "GTR" else np.concatenate([f, [k]])

    def p_of(tables, tau):
        return tables["evec"] @ np.diag(np.exp(tables["eval"] * tau)) @ tables["ivec"]

    delta = 1e-6
    lam = base["eval"]
    for tau in (1e-3, 0.07, 0.9, 6.0):
        diff = lam[:, None] - lam[None, :]
        with np.errstate(divide="ignore", invalid="ignore"):
            phi = np.where(np.abs(diff) > 1e-12,
                           (np.exp(lam[:, None] * tau) - np.exp(lam[None, :] * tau)) / diff,
                           tau * np.exp(lam[:, None] * tau) * np.ones((4, 4)))
        for index in range(count.value):
            analytic = base["evec"] @ (b[index] * phi) @ base["ivec"]
            up = model_tables(substitution, "constant", perturbed(index, +delta))
            down = model_tables(substitution, "constant", perturbed(index, -delta))
            numeric = (p_of(up, tau) - p_of(down, tau)) / (2 * delta)
            assert np.max(np.abs(analytic - numeric)) < 1e-8, (tau, index)
            assert np.allclose(dfreqs[index], (up["freqs"] - down["freqs"]) / (2 * delta), atol=1e-9)
