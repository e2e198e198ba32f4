"""Generates tests/golden/*.npz by running the UNMODIFIED reference here.

Run in the build container only (needs /root/reference and the reference
python module built by `make -C oracle ref`, i.e. the reference host code
linked against oracle/beagle_cpu.cpp):

    python tests/golden/make_fixtures.py

Each fixture holds the flat inputs of one reference test scenario (site
patterns, weights, Node::ParentIdVector per tree, branch lengths, rooted-tree
fields, the phylo-model parameter matrix) together with what the reference's
own `log_likelihoods()` / `phylo_gradients()` returned for them, plus the
external golden numbers (pybeagle / physher / phylotorch) hard-coded in the
reference's doctests (src/unrooted_sbn_instance.hpp:206-335,
src/rooted_sbn_instance.hpp:246-378).  Nothing here runs on the GPU box.
"""
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF_BUILD = os.path.join(ROOT, "oracle", "_ref")
DATA = "/root/reference/data"
sys.path.insert(0, REF_BUILD)
sys.path.insert(0, ROOT)

import libsbn  # noqa: E402  (the reference's pybind11 module)
from libsbn_b200 import alignment  # noqa: E402
from oracle import site_pattern  # noqa: E402


def cpp_doubles(header, marker, occurrence=0):
    """The brace-initialised list of doubles that follows `marker` in a
    reference header: golden vectors are read from the reference's own doctest
    blocks at generation time instead of being retyped."""
    text = open(os.path.join("/root/reference/src", header)).read()
    start = -1
    for _ in range(occurrence + 1):
        start = text.index(marker, start + 1)
    open_brace = text.index("{", start)
    close_brace = text.index("}", open_brace)
    return np.array([float(x) for x in re.findall(r"-?\d+\.?\d*(?:[eE][-+]?\d+)?",
                                                  text[open_brace:close_brace])])


def vec(x):
    return np.array(x, dtype=np.float64, copy=True)


def collect_trees(inst, rooted):
    trees = inst.tree_collection.trees
    out = {
        "parent_ids": np.array([t.parent_id_vector() for t in trees], dtype=np.int32),
        "branch_lengths": np.array([vec(t.branch_lengths) for t in trees]),
    }
    if rooted:
        out["rates"] = np.array([vec(t.rates) for t in trees])
        out["node_heights"] = np.array([vec(t.node_heights) for t in trees])
        out["node_bounds"] = np.array([vec(t.node_bounds) for t in trees])
        out["height_ratios"] = np.array([vec(t.height_ratios) for t in trees])
    return out


def collect_gradients(gradients):
    out = {"grad_log_likelihood": np.array([g.log_likelihood for g in gradients])}
    for key in gradients[0].gradient:
        out["grad_" + key] = np.array([vec(g.gradient[key]) for g in gradients])
    return out


def save(name, inst, fasta, spec, rooted, extra):
    inst.process_loaded_trees()  # populates taxon_names() (leaf-id order)
    # (no GPU in the build container: the NumPy restatement of SitePattern::Compress)
    patterns, weights = site_pattern.compress(
        alignment.sequences_in_leaf_order(alignment.read_fasta(fasta), inst.taxon_names()))
    record = {
        "patterns": patterns,
        "weights": weights,
        "substitution": spec[0],
        "site": spec[1],
        "clock": spec[2],
        "rooted": rooted,
        "params": np.array(inst.get_phylo_model_params(), dtype=np.float64, copy=True),
    }
    record.update(collect_trees(inst, rooted))
    record.update(extra)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **record)
    print(name, {k: (v.shape if hasattr(v, "shape") else v) for k, v in record.items()})


def unrooted_case(name, newick=None, nexus=None, fasta=None, spec=("JC69", "constant", "none"),
                  set_params=None, set_lengths=None, goldens=None):
    inst = libsbn.unrooted_instance(name)
    if newick:
        inst.read_newick_file(newick)
    else:
        inst.read_nexus_file(nexus)
    inst.read_fasta_file(fasta)
    if set_lengths:
        set_lengths(inst)
    inst.prepare_for_phylo_likelihood(libsbn.PhyloModelSpecification(*spec), 2, [], True)
    if set_params:
        set_params(inst.get_phylo_model_param_block_map())
    extra = dict(goldens or {})
    for rescaling in (False, True):
        inst.set_rescaling(rescaling)
        tag = "_rescaled" if rescaling else ""
        extra["log_likelihoods" + tag] = vec(inst.log_likelihoods())
        for key, value in collect_gradients(inst.phylo_gradients()).items():
            extra[key + tag] = value
    save(name, inst, fasta, spec, False, extra)


def rooted_case(name, spec, set_params=None, relaxed=False, goldens=None):
    inst = libsbn.rooted_instance(name)
    inst.read_newick_file(f"{DATA}/fluA.tree")
    inst.parse_dates_from_taxon_names(True)
    inst.read_fasta_file(f"{DATA}/fluA.fa")
    inst.prepare_for_phylo_likelihood(libsbn.PhyloModelSpecification(*spec), 1, [], True)
    for tree in inst.tree_collection.trees:
        rates = np.array(tree.rates, copy=False)
        rates[:] = 0.001
        if relaxed:
            # rooted_sbn_instance.hpp:311-315.  rate_count_ is not exposed to
            # python, so the relaxed-clock *gradient* layout is exercised against
            # the flat oracle instead; the likelihood only sees the rates.
            rates *= (np.arange(rates.size) % 3 + 1.0)
    if set_params:
        set_params(inst.get_phylo_model_param_block_map())
    extra = dict(goldens or {})
    extra["log_likelihoods"] = vec(inst.log_likelihoods())
    extra.update(collect_gradients(inst.phylo_gradients()))
    save(name, inst, f"{DATA}/fluA.fa", spec, True, extra)


CASES = {}


def case(fn):
    CASES[fn.__name__] = fn
    return fn


def weibull_01(block_map):
    block_map["Weibull shape"][:] = 0.1


def main():
    # One reference scenario per subprocess: the reference's flex/bison driver
    # segfaults when a Nexus file is parsed after a Newick file in one process.
    if len(sys.argv) == 1:
        import subprocess
        for name in CASES:
            subprocess.run([sys.executable, os.path.abspath(__file__), name], check=True)
        return
    os.chdir(REF_BUILD)  # the reference prints/writes relative to cwd
    CASES[sys.argv[1]]()


@case
def unrooted_cases_newick():

    # hello (unrooted_sbn_instance.hpp:206-214): golden -84.852358
    unrooted_case("hello_jc69", newick=f"{DATA}/hello.nwk", fasta=f"{DATA}/hello.fasta",
                  goldens={"golden_log_likelihoods": np.array([-84.852358]), "golden_tol": 1e-6})



@case
def unrooted_cases_nexus():
    # DS1 x 10 JC69 (unrooted_sbn_instance.hpp:215-279): pybeagle logLs, physher
    # gradient of the last tree (sorted).
    pybeagle = cpp_doubles("unrooted_sbn_instance.hpp", "std::vector<double> pybeagle_likelihoods(")
    physher_gradient_sorted = cpp_doubles("unrooted_sbn_instance.hpp", "std::vector<double> physher_gradients = {")
    unrooted_case("ds1_jc69", nexus=f"{DATA}/DS1.subsampled_10.t", fasta=f"{DATA}/DS1.fasta",
                  goldens={"golden_log_likelihoods": pybeagle, "golden_tol": 1.1e-4,
                           "golden_last_gradient_sorted": physher_gradient_sorted})

    # DS1 x 10 JC69 + weibull4, shape 0.1 (unrooted_sbn_instance.hpp:284-335).
    physher_weibull = cpp_doubles("unrooted_sbn_instance.hpp", "std::vector<double> physher_likelihoods(")
    physher_weibull_grad0 = cpp_doubles("unrooted_sbn_instance.hpp", "std::vector<double> physher_gradients_bl0(")

    unrooted_case("ds1_jc69_weibull4", nexus=f"{DATA}/DS1.subsampled_10.t", fasta=f"{DATA}/DS1.fasta",
                  spec=("JC69", "weibull+4", "none"), set_params=weibull_01,
                  goldens={"golden_log_likelihoods": physher_weibull, "golden_tol": 1.1e-4,
                           "golden_first_branch_gradient": physher_weibull_grad0})

    # DS1 x 10 GTR + weibull4 (BASELINE config 2, pinned variant): no external golden.
    def gtr_weibull(block_map):
        block_map["GTR rates"][:] = [0.05, 0.1, 0.15, 0.20, 0.25, 0.25]
        block_map["frequencies"][:] = [0.1, 0.2, 0.3, 0.4]
        block_map["Weibull shape"][:] = 0.5

    unrooted_case("ds1_gtr_weibull4", nexus=f"{DATA}/DS1.subsampled_10.t", fasta=f"{DATA}/DS1.fasta",
                  spec=("GTR", "weibull+4", "none"), set_params=gtr_weibull)



@case
def unrooted_cases_newick_ds1():
    # DS1 x 100 topologies, all branch lengths 1 (BASELINE config 1).
    unrooted_case("ds1_100_topologies_jc69", newick=f"{DATA}/DS1.100_topologies.nwk",
                  fasta=f"{DATA}/DS1.fasta")



@case
def unrooted_cases_newick_ds1_config2():
    # BASELINE configs[1] exactly: DS1 x DS1.100_topologies.nwk with
    # branch lengths ~ Exp(mean 0.1) (numpy default_rng(1), tree-major), GTR + weibull4 with
    # the rooted_sbn_instance.hpp:336-337 rates / frequencies and shape 0.5: the
    # reference's log_likelihoods() and full phylo_gradients() for all 100 trees.
    def exponential_lengths(inst):
        rng = np.random.default_rng(1)
        for tree in inst.tree_collection.trees:
            lengths = np.array(tree.branch_lengths, copy=False)
            lengths[:-1] = np.maximum(rng.exponential(0.1, size=lengths.size - 1), 1e-6)

    def gtr_weibull(block_map):
        block_map["GTR rates"][:] = [0.05, 0.1, 0.15, 0.20, 0.25, 0.25]
        block_map["frequencies"][:] = [0.1, 0.2, 0.3, 0.4]
        block_map["Weibull shape"][:] = 0.5

    unrooted_case("ds1_100_topologies_gtr_weibull4", newick=f"{DATA}/DS1.100_topologies.nwk",
                  fasta=f"{DATA}/DS1.fasta", spec=("GTR", "weibull+4", "none"), set_params=gtr_weibull,
                  set_lengths=exponential_lengths)



@case
def unrooted_cases_nexus_reordered():
    # test/test_libsbn.py:95-118: JC69 == GTR(equal) on DS1 tree 0, all branches 0.1.
    def only_first_tree_01(inst):
        inst.tree_collection.erase(1, 10)
        np.array(inst.tree_collection.trees[0].branch_lengths, copy=False)[:] = 0.1

    def gtr_equal(block_map):
        block_map["GTR rates"][:] = 1.0 / 6
        block_map["frequencies"][:] = 0.25

    unrooted_case("ds1_tree0_gtr_equal", nexus=f"{DATA}/DS1.subsampled_10.t.reordered",
                  fasta=f"{DATA}/DS1.fasta", spec=("GTR", "constant", "none"),
                  set_params=gtr_equal, set_lengths=only_first_tree_01)



@case
def rooted_cases():
    # fluA rooted (rooted_sbn_instance.hpp:246-378).
    physher_ratio_gradient = cpp_doubles("rooted_sbn_instance.hpp", "std::vector<double> physher_gradients = {")
    rooted_case("flua_jc69_strict", ("JC69", "constant", "strict"),
                goldens={"golden_log_likelihood": -4777.616349, "golden_jacobian": -9.25135166,
                         "golden_ratio_gradient": physher_ratio_gradient, "golden_tol": 1e-4})
    rooted_case("flua_jc69_varied_rates", ("JC69", "constant", "strict"), relaxed=True)

    def gtr_flu(block_map):
        block_map["frequencies"][:] = [0.1, 0.2, 0.3, 0.4]
        block_map["GTR rates"][:] = [0.05, 0.1, 0.15, 0.20, 0.25, 0.25]

    phylotorch_gtr = cpp_doubles("rooted_sbn_instance.hpp", "std::vector<double> phylotorch_gradients = {")
    rooted_case("flua_gtr_strict", ("GTR", "constant", "strict"), set_params=gtr_flu,
                goldens={"golden_log_likelihood": -5221.438941335706, "golden_jacobian": -9.25135166,
                         "golden_substitution_gradient": phylotorch_gtr, "golden_tol": 1e-3})
    rooted_case("flua_jc69_weibull4_strict", ("JC69", "weibull+4", "strict"), set_params=weibull_01,
                goldens={"golden_log_likelihood": -4618.2062529058, "golden_jacobian": -9.25135166,
                         "golden_site_gradient": -5.231329, "golden_tol": 1e-3})


if __name__ == "__main__":
    main()
