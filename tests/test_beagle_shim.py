"""The inner drop-in boundary: libsbn_b200/lib/libhmsbeagle_b200.so, the 17
BEAGLE entry points of include/libhmsbeagle/beagle.h on the device
(libsbn_b200/csrc/beagle_shim.cu).

  * CPU: the library exports every function the header declares and fails loudly without a device;
  * GPU: the same call sequence -- the one fat_beagle.cpp makes for a log likelihood and a branch
    gradient (fat_beagle.cpp:50-70, 119-175) -- through the device library and through the CPU
    restatement (oracle/beagle_cpu.cpp, the test oracle pinned to the reference's goldens):
    compact tips and tip partials, gaps, 1 / 3 / 4 rate categories, rescaling on and off, a
    ragged pattern count;
  * GPU: the reference's own doctest suite, EVERY reference source unmodified (fat_beagle.cpp
    and engine.cpp included), linked against the device library (integration/Makefile).
"""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT
import test_integration_gpu as integration

SHIM = os.path.join(ROOT, "libsbn_b200", "lib", "libhmsbeagle_b200.so")
ORACLE = os.path.join(ROOT, "oracle", "_build", "libbeagle_oracle.so")
OP_NONE = -1
_I, _D = ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_double)


def _declared():
    header = open(os.path.join(ROOT, "include", "libhmsbeagle", "beagle.h")).read()
    return sorted(set(re.findall(r"^int (beagle[A-Za-z]+)\(", header, flags=re.M)))


def test_library_exports_every_declared_symbol():
    declared = _declared()
    assert len(declared) == 17
    lib = ctypes.CDLL(SHIM)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in beagle.h but not exported"


def test_no_device_means_no_resource_not_a_cpu_fallback():
    import libsbn_b200._capi as capi
    if capi.load().sbnb_device_count() > 0:
        pytest.skip("a CUDA device is present")
    lib = ctypes.CDLL(SHIM)
    handle = lib.beagleCreateInstance(3, 7, 3, 4, 10, 1, 10, 1, 8, None, 0, 0, 0, None)
    assert handle == -6  # BEAGLE_ERROR_NO_RESOURCE


class Beagle:
    """The BEAGLE calls of fat_beagle.cpp over one of the two libraries."""

    def __init__(self, path, n, P, C, use_tip_states):
        self.lib = ctypes.CDLL(path)
        self.n, self.P, self.C, self.N = n, P, C, 2 * n - 1
        partials = 3 * n - 2 + (0 if use_tip_states else n)  # fat_beagle.cpp:207-256
        self.handle = self.lib.beagleCreateInstance(n, partials, n if use_tip_states else 0, 4, P, 1, 2 * self.N, C,
                                                    partials + 1, None, 0, 0, 1 << 6, None)
        assert self.handle >= 0, self.handle

    def ok(self, code):
        assert code == 0, code

    def _d(self, a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        return a, a.ctypes.data_as(_D)

    def _i(self, a):
        a = np.ascontiguousarray(a, dtype=np.int32)
        return a, a.ctypes.data_as(_I)

    def set_tips(self, states, weights, use_tip_states):
        for tip in range(self.n):
            if use_tip_states:
                keep, ptr = self._i(states[tip])
                self.ok(self.lib.beagleSetTipStates(self.handle, tip, ptr))
            else:
                partial = np.zeros((self.P, 4))
                for k, s in enumerate(states[tip]):
                    partial[k, :] = 1.0 if s >= 4 else 0.0
                    if s < 4:
                        partial[k, s] = 1.0
                keep, ptr = self._d(partial)
                self.ok(self.lib.beagleSetTipPartials(self.handle, tip, ptr))
        keep, ptr = self._d(weights)
        self.ok(self.lib.beagleSetPatternWeights(self.handle, ptr))

    def set_model(self, evec, ivec, evals, freqs, rates, proportions):
        keeps = [self._d(x) for x in (evec, ivec, evals, freqs, rates, proportions)]
        self.ok(self.lib.beagleSetStateFrequencies(self.handle, 0, keeps[3][1]))
        self.ok(self.lib.beagleSetEigenDecomposition(self.handle, 0, keeps[0][1], keeps[1][1], keeps[2][1]))
        self.ok(self.lib.beagleSetCategoryWeights(self.handle, 0, keeps[5][1]))
        self.ok(self.lib.beagleSetCategoryRates(self.handle, keeps[4][1]))

    def log_likelihood_and_gradient(self, post_ops, pre_ops, lengths, q, rates, freqs, rescaling):
        """FatBeagle::BranchGradientInternals (fat_beagle.cpp:119-175)."""
        n, N = self.n, self.N
        keep_i, idx = self._i(np.arange(N - 1))
        keep_l, lens = self._d(lengths[:N - 1])
        self.ok(self.lib.beagleUpdateTransitionMatrices(self.handle, 0, idx, None, None, lens, N - 1))
        dq = np.stack([r * q for r in rates])
        keep_q, dq_ptr = self._d(dq)
        self.ok(self.lib.beagleSetDifferentialMatrix(self.handle, N - 1, dq_ptr))
        cumulative = 0 if rescaling else OP_NONE
        if rescaling:
            self.ok(self.lib.beagleResetScaleFactors(self.handle, 0))
        keep_a, ops = self._i(post_ops)
        self.ok(self.lib.beagleUpdatePartials(self.handle, ops, len(post_ops), cumulative))
        root_pre = np.tile(freqs, self.C * self.P)
        keep_r, root_ptr = self._d(root_pre)
        self.ok(self.lib.beagleSetPartials(self.handle, 2 * N - 1, root_ptr))
        keep_b, ops = self._i(pre_ops)
        self.ok(self.lib.beagleUpdatePrePartials(self.handle, ops, len(pre_ops), OP_NONE))
        keep_1, post_idx = self._i(np.arange(N - 1))
        keep_2, pre_idx = self._i(np.arange(N - 1) + N)
        keep_3, dm_idx = self._i(np.full(N - 1, N - 1))
        keep_4, zero = self._i([0])
        sums, squares = np.zeros(N - 1), np.zeros(N - 1)
        per_site = np.zeros((N - 1, self.P))
        self.ok(self.lib.beagleCalculateEdgeDerivatives(self.handle, post_idx, pre_idx, dm_idx, zero, N - 1,
                                                        per_site.ctypes.data_as(_D), sums.ctypes.data_as(_D),
                                                        squares.ctypes.data_as(_D)))
        keep_5, root = self._i([N - 1])
        keep_6, cum = self._i([cumulative])
        logl = ctypes.c_double()
        self.ok(self.lib.beagleCalculateRootLogLikelihoods(self.handle, root, zero, zero, cum, 1, ctypes.byref(logl)))
        return logl.value, sums, squares, per_site

    def close(self):
        self.ok(self.lib.beagleFinalizeInstance(self.handle))


def _random_tree_ops(n, rng, rescaling):
    """A random rooted binary tree in libsbn's numbering (leaves 0..n-1, internal nodes in post-order,
    root 2n-2) as the op lists of fat_beagle.cpp:327-362."""
    N = 2 * n - 1
    roots, children, next_id = list(range(n)), {}, n
    while len(roots) > 1:
        a, b = (roots.pop(int(rng.integers(len(roots)))) for _ in range(2))
        children[next_id] = (a, b)
        roots.append(next_id)
        next_id += 1
    post = [[v, (v - n + 1) if rescaling else OP_NONE, OP_NONE, a, a, b, b] for v, (a, b) in sorted(children.items())]
    pre = []
    for v in sorted(children, reverse=True):  # parents before children
        for child, sister in (children[v], children[v][::-1]):
            pre.append([child + N, (child + 1 + n - 1) if rescaling else OP_NONE, OP_NONE, v + N, child, sister, sister])
    return np.array(post, dtype=np.int32), np.array(pre, dtype=np.int32)


def _gtr():
    rates = np.array([0.05, 0.1, 0.15, 0.20, 0.25, 0.25])
    freqs = np.array([0.1, 0.2, 0.3, 0.4])
    q = np.zeros((4, 4))
    index = 0
    for i in range(4):
        for j in range(i + 1, 4):
            q[i, j], q[j, i] = rates[index] * freqs[j], rates[index] * freqs[i]
            index += 1
    q -= np.diag(q.sum(axis=1))
    q /= -np.sum(np.diag(q) * freqs)
    root = np.sqrt(freqs)
    values, vectors = np.linalg.eigh(root[:, None] * q / root[None, :])
    return vectors / root[:, None], vectors.T * root[None, :], values, freqs, q


@pytest.mark.gpu
@pytest.mark.parametrize("categories,use_tip_states,rescaling", [(1, True, False), (4, True, True), (4, False, False),
                                                                 (3, True, True), (4, True, False)])
def test_device_library_against_the_cpu_restatement(categories, use_tip_states, rescaling):
    rng = np.random.default_rng(categories * 10 + use_tip_states * 2 + rescaling)
    n, P = 13, 1003
    states = rng.integers(0, 4, size=(n, P)).astype(np.int32)
    states[rng.random(states.shape) < 0.05] = 4
    weights = rng.integers(1, 6, size=P).astype(np.float64)
    post, pre = _random_tree_ops(n, rng, rescaling)
    lengths = rng.exponential(0.1, size=2 * n - 1)
    evec, ivec, evals, freqs, q = _gtr()
    rates = np.sort(rng.gamma(2.0, 0.5, size=categories))
    rates /= rates.mean()
    proportions = np.full(categories, 1.0 / categories)
    results = []
    for path in (SHIM, ORACLE):
        beagle = Beagle(path, n, P, categories, use_tip_states)
        beagle.set_tips(states, weights, use_tip_states)
        beagle.set_model(evec, ivec, evals, freqs, rates, proportions)
        results.append(beagle.log_likelihood_and_gradient(post, pre, lengths, q, rates, freqs, rescaling))
        # a second evaluation on the same instance with other lengths (buffers are reused)
        results.append(beagle.log_likelihood_and_gradient(post, pre, lengths * 1.7, q, rates, freqs, rescaling))
        beagle.close()
    for ours, want in ((results[0], results[2]), (results[1], results[3])):
        assert abs(ours[0] - want[0]) <= 1e-12 * abs(want[0])
        np.testing.assert_allclose(ours[1], want[1], rtol=1e-10, atol=1e-9)
        np.testing.assert_allclose(ours[2], want[2], rtol=1e-10, atol=1e-9)
        np.testing.assert_allclose(ours[3], want[3], rtol=1e-10, atol=1e-11)


@pytest.mark.gpu
def test_reference_doctest_suite_passes_over_the_device_beagle_library():
    """reference src/doctest.cpp with EVERY reference source unmodified -- fat_beagle.cpp, engine.cpp,
    site_pattern.cpp, ... -- and libhmsbeagle replaced by libhmsbeagle_b200.so at link time."""
    integration._compare_with_reference_build("doctest_beagle", ("likelihood", "gradients", "time trees"), 42,
                                              reference_name="doctest")
    with open(integration._artefact("doctest_beagle"), "rb") as handle:
        binary = handle.read()
    assert b"beagleUpdatePrePartials" in binary  # the reference's FatBeagle is what runs
