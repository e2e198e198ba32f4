"""The drop-in proof: the UNMODIFIED reference test suites and python module,
built by integration/Makefile on top of libsbn_b200.so instead of BEAGLE (only
src/engine.{hpp,cpp} and src/gp_engine.{hpp,cpp} replaced by their counterparts in
integration/), run on the device.

The artefacts are built in the build container (where /root/reference is
mounted) and travel to the GPU box; the reference's input fixtures were staged
next to oracle/_ref by `make -C oracle ref`.  Nothing here reads /root/reference.
"""
import os
import re
import subprocess
import sys
import textwrap

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "integration", "_build")
REF_RUN_DIR = os.path.join(ROOT, "oracle", "_ref")  # holds data/ as the reference's tests expect
GOLDEN = os.path.join(ROOT, "tests", "golden")


def _artefact(name):
    path = os.path.join(BUILD, name)
    if not os.path.exists(path) or not os.path.isdir(os.path.join(REF_RUN_DIR, "data")):
        pytest.skip(f"{path} not built (make -C integration needs the reference sources)")
    return path


def _run_doctest(binary, *args):
    """Runs a doctest binary; returns (exit code, (total, passed, failed), names of
    the failed test cases, full text)."""
    os.makedirs(os.path.join(REF_RUN_DIR, "_ignore"), exist_ok=True)
    done = subprocess.run([binary, *args], cwd=REF_RUN_DIR, capture_output=True, text=True, timeout=900)
    text = (done.stdout + done.stderr).replace("\r", "\n")
    summary = re.search(r"test cases:\s+(\d+)\s+\|\s+(\d+) passed\s+\|\s+(\d+) failed", text)
    assert summary, text[-2000:]
    failed_cases = set(re.findall(r"^TEST CASE:\s+(.*)$", text, flags=re.M))
    return done.returncode, tuple(int(x) for x in summary.groups()), failed_cases, text


def _compare_with_reference_build(name, engine_cases, minimum_cases, reference_name=None):
    """The suite built over libsbn_b200.so must pass every test case the UNMODIFIED
    reference build (oracle/_ref, BEAGLE-equivalent CPU kernels) passes on this very
    host, and every case that touches the likelihood engine.  (Some reference cases
    fail on their own: they hard-code libstdc++ hash-iteration orders or hold 23 EM
    iterations to 1e-12 across libm versions -- SURVEY.md 8c; none touches the engine.)"""
    ours = _artefact(name)
    theirs = os.path.join(REF_RUN_DIR, reference_name or name)
    assert os.path.exists(theirs), theirs
    _, (total, passed, failed), our_failures, text = _run_doctest(ours)
    _, (ref_total, _, _), reference_failures, _ = _run_doctest(theirs)
    errors = "\n".join(line for line in text.split("\n") if "ERROR" in line)[:3000]
    assert total == ref_total and total >= minimum_cases
    assert our_failures <= reference_failures, (our_failures - reference_failures, errors)
    assert not [case for case in our_failures if any(key in case for key in engine_cases)], errors
    assert passed == total - len(our_failures)


@pytest.mark.gpu
def test_reference_doctest_suite_passes_on_the_device():
    """reference src/doctest.cpp: every test case, including the pybeagle / physher /
    phylotorch goldens of {un,}rooted_sbn_instance.hpp, through sbn_b200.h."""
    _compare_with_reference_build("doctest", ("likelihood", "gradients", "time trees"), 42)
    with open(_artefact("doctest"), "rb") as handle:
        # SitePattern::Compress is integration/site_pattern.cpp (device) in this build
        assert b"IntVectorHasher" not in handle.read(), "the reference SitePattern::Compress is still linked in"


@pytest.mark.gpu
def test_reference_gp_doctest_suite_passes_on_the_device():
    """reference src/gp_doctest.cpp with GPEngine replaced by integration/gp_engine.*
    (PLVs in HBM, op programs on the device through sbn_b200_gp.h): GP likelihoods,
    marginals, Brent branch-length optimisation, SBN parameter updates, rescaling
    and quartet hybrid marginals; its exact-marginal cross-pins also evaluate every
    topology through Engine (gp_doctest.cpp:110-156)."""
    _compare_with_reference_build(
        "gp_doctest", ("classical likelihood", "two tree marginal", "marginal likelihood on", "GPEngine",
                       "gradient calculation", "rescaling", "test populate PLV", "SBN root split",
                       "GPInstance: simplest hybrid marginal"), 50)
    with open(_artefact("gp_doctest"), "rb") as handle:
        assert b"BrentOptimization" not in handle.read(), "the reference GPEngine is still linked in"


_PYTHON_CASE = textwrap.dedent("""
    import sys, numpy as np
    sys.path.insert(0, sys.argv[1])
    import libsbn  # the reference's pybind11 module, linked against libsbn_b200.so
    inst = libsbn.unrooted_instance("ds1")
    inst.read_nexus_file("data/DS1.subsampled_10.t")
    inst.read_fasta_file("data/DS1.fasta")
    inst.prepare_for_phylo_likelihood(libsbn.PhyloModelSpecification("GTR", "weibull+4", "none"), 2, [], True)
    block_map = inst.get_phylo_model_param_block_map()
    block_map["GTR rates"][:] = [0.05, 0.1, 0.15, 0.20, 0.25, 0.25]
    block_map["frequencies"][:] = [0.1, 0.2, 0.3, 0.4]
    block_map["Weibull shape"][:] = 0.5
    out = {}
    for rescaling in (False, True):
        inst.set_rescaling(rescaling)
        tag = "_rescaled" if rescaling else ""
        out["log_likelihoods" + tag] = np.array(inst.log_likelihoods())
        gradients = inst.phylo_gradients()
        out["grad_log_likelihood" + tag] = np.array([g.log_likelihood for g in gradients])
        for key in gradients[0].gradient:
            out["grad_" + key + tag] = np.array([np.array(g.gradient[key]) for g in gradients])
    np.savez(sys.argv[2], **out)
""")


@pytest.mark.gpu
def test_reference_python_module_matches_the_reference_run_on_cpu(tmp_path):
    """`import libsbn` (reference src/pylibsbn.cpp, unchanged) on the device against
    the same calls made by the reference over its CPU path (tests/golden,
    make_fixtures.py `ds1_gtr_weibull4`): logL 1e-10, gradients 1e-8 relative."""
    _artefact("doctest")
    module = [f for f in os.listdir(BUILD) if f.startswith("libsbn") and f.endswith(".so")]
    assert module, "integration/_build/libsbn*.so missing"
    out = str(tmp_path / "out.npz")
    done = subprocess.run([sys.executable, "-c", _PYTHON_CASE, BUILD, out], cwd=REF_RUN_DIR,
                          capture_output=True, text=True, timeout=600)
    assert done.returncode == 0, done.stderr[-2000:]
    got = np.load(out)
    want = np.load(os.path.join(GOLDEN, "ds1_gtr_weibull4.npz"))
    for tag in ("", "_rescaled"):
        for key in ("log_likelihoods", "grad_log_likelihood"):
            np.testing.assert_allclose(got[key + tag], want[key + tag], rtol=1e-10)
        scale = np.max(np.abs(want["grad_branch_lengths" + tag]), axis=1, keepdims=True)
        assert np.max(np.abs(got["grad_branch_lengths" + tag] - want["grad_branch_lengths" + tag]) / scale) < 1e-8
        # Finite differences of logL with delta 1e-6 amplify 1e-12 differences in logL
        # by 5e5; the reference's own test holds this block to 1e-3 (rooted_sbn_instance.hpp:346-352).
        np.testing.assert_allclose(got["grad_substitution_model" + tag], want["grad_substitution_model" + tag],
                                   rtol=1e-4, atol=1e-3)
        # The reference evaluates this block on a model left perturbed by its
        # finite-difference loop ("reference quirk").
        np.testing.assert_allclose(got["grad_site_model" + tag], want["grad_site_model" + tag], rtol=1e-4)


def test_reference_doctest_fails_loudly_without_a_device():
    """No CPU fallback behind the reference's Engine either."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    code, (total, passed, failed), _, text = _run_doctest(
        _artefact("doctest"), "-tc=UnrootedSBNInstance: likelihood and gradient with Weibull")
    assert code != 0 and failed == 1
    assert "no CPU fallback" in text


_GIL_CASE = textwrap.dedent("""
    import sys, threading, time
    sys.path.insert(0, sys.argv[1])
    import libsbn
    inst = libsbn.unrooted_instance("ds1")
    inst.read_newick_file("data/DS1.100_topologies.nwk")
    inst.read_fasta_file("data/DS1.fasta")
    inst.prepare_for_phylo_likelihood(libsbn.PhyloModelSpecification("GTR", "weibull+4", "none"), 1, [], True)
    block_map = inst.get_phylo_model_param_block_map()
    block_map["GTR rates"][:] = [0.05, 0.1, 0.15, 0.20, 0.25, 0.25]
    block_map["frequencies"][:] = [0.1, 0.2, 0.3, 0.4]
    block_map["Weibull shape"][:] = 0.5
    inst.phylo_gradients()  # warm up
    ticks, inside, stop = [0, 0], [False], [False]
    def spin():
        while not stop[0]:
            ticks[1 if inside[0] else 0] += 1
            time.sleep(0)
    thread = threading.Thread(target=spin)
    thread.start()
    t0 = time.perf_counter()
    inside[0] = True
    for _ in range(200):
        inst.phylo_gradients()
    inside[0] = False
    elapsed = time.perf_counter() - t0
    stop[0] = True
    thread.join()
    print("TICKS_INSIDE", ticks[1], "ELAPSED", elapsed)
""")


@pytest.mark.gpu
def test_device_calls_release_the_gil():
    """a Python thread keeps running while the reference's (unchanged)
    python module is inside phylo_gradients() on the device -- the replacement Engine
    releases the GIL around every libsbn_b200 call (integration/scoped_gil_release.hpp)."""
    _artefact("doctest")
    done = subprocess.run([sys.executable, "-c", _GIL_CASE, BUILD], cwd=REF_RUN_DIR, capture_output=True,
                          text=True, timeout=600)
    assert done.returncode == 0, done.stderr[-2000:]
    found = re.search(r"TICKS_INSIDE (\d+) ELAPSED ([\d.]+)", done.stdout)
    assert found, done.stdout[-2000:]
    ticks, elapsed = int(found.group(1)), float(found.group(2))
    # with the GIL held across the calls the other thread only gets the gaps between
    # them (a few switch intervals); released, it spins the whole time
    assert ticks > 2000 * elapsed, (ticks, elapsed)
