// integration/engine.cpp -- replacement body of phylovi/libsbn's src/engine.cpp:
// each of the five Engine batch methods (reference src/engine.cpp:54-93) flattens
// its tree collection into an sbnb_tree_batch and makes ONE call into
// libsbn_b200.so (include/sbn_b200.h), where the reference fanned the trees over a
// thread pool of BEAGLE instances (FatBeagleParallelize, fat_beagle.hpp:119-149).
// Failures come back as the reference's own exception type: Failwith ->
// std::runtime_error -> Python RuntimeError (sugar.hpp:67-78).

#include "engine.hpp"

#include <algorithm>
#include <cstdlib>
#include <sstream>
#include <string>

#include "sbn_b200.h"
#include "scoped_gil_release.hpp"

namespace {

void Check(int status) {
  if (status != SBNB_OK) {
    Failwith(std::string("libsbn_b200: ") + sbnb_last_error());
  }
}

// A tree collection as the flat arrays of sbnb_tree_batch.  Owns the storage the
// struct points into.
struct FlatTrees {
  sbnb_tree_batch batch{};
  std::vector<int32_t> parent_ids;
  std::vector<double> branch_lengths, rates, node_heights, node_bounds, height_ratios;

  template <typename TTreeCollection>
  void AddTopologiesAndBranchLengths(const TTreeCollection &tree_collection) {
    const size_t tree_count = tree_collection.TreeCount();
    batch.tree_count = static_cast<int32_t>(tree_count);
    if (tree_count == 0) {
      return;
    }
    const size_t node_count = tree_collection.GetTree(0).BranchLengths().size();
    batch.node_count = static_cast<int32_t>(node_count);
    parent_ids.reserve(tree_count * (node_count - 1));
    branch_lengths.reserve(tree_count * node_count);
    for (size_t tree_idx = 0; tree_idx < tree_count; tree_idx++) {
      const auto &tree = tree_collection.GetTree(tree_idx);
      const auto tree_parent_ids = tree.ParentIdVector();
      if (tree.BranchLengths().size() != node_count ||
          tree_parent_ids.size() != node_count - 1) {
        Failwith("All trees of a collection must have the same number of nodes.");
      }
      for (const auto parent_id : tree_parent_ids) {
        parent_ids.push_back(static_cast<int32_t>(parent_id));
      }
      branch_lengths.insert(branch_lengths.end(), tree.BranchLengths().begin(),
                            tree.BranchLengths().end());
    }
    batch.parent_ids = parent_ids.data();
    batch.branch_lengths = branch_lengths.data();
  }

  // The time-tree extras of RootedTree (rooted_tree.hpp:86-102).  The getters
  // throw, as in the reference, when tip dates / the time tree were never set.
  void AddTimeTreeExtras(const RootedTreeCollection &tree_collection) {
    for (const auto &tree : tree_collection.Trees()) {
      const auto append = [](std::vector<double> &to, const std::vector<double> &from) {
        to.insert(to.end(), from.begin(), from.end());
      };
      append(rates, tree.GetRates());
      append(node_heights, tree.GetNodeHeights());
      append(node_bounds, tree.GetNodeBounds());
      append(height_ratios, tree.GetHeightRatios());
      batch.rate_count = static_cast<int32_t>(tree.RateCount());
    }
    batch.rates = rates.data();
    batch.node_heights = node_heights.data();
    batch.node_bounds = node_bounds.data();
    batch.height_ratios = height_ratios.data();
  }
};

// One contiguous row per tree (an Eigen::Ref may carry an outer stride).
std::vector<double> ContiguousParams(const EigenMatrixXdRef &params, size_t tree_count) {
  Assert(tree_count == static_cast<size_t>(params.rows()),
         "We param_matrix needs as many rows as we have trees.");
  std::vector<double> flat(static_cast<size_t>(params.rows() * params.cols()));
  for (Eigen::Index row = 0; row < params.rows(); row++) {
    for (Eigen::Index col = 0; col < params.cols(); col++) {
      flat[static_cast<size_t>(row * params.cols() + col)] = params(row, col);
    }
  }
  return flat;
}

// Per-tree result blocks of a gradient call, keyed as the reference's GradientMap
// (fat_beagle.cpp:467-545).
struct GradientBlocks {
  size_t tree_count;
  std::vector<double> log_likelihood;
  std::vector<std::pair<std::string, std::pair<size_t, std::vector<double>>>> blocks;
  sbnb_gradient_out out{};

  explicit GradientBlocks(size_t tree_count)
      : tree_count(tree_count), log_likelihood(tree_count, 0.) {
    blocks.reserve(8);
    out.log_likelihood = log_likelihood.data();
  }
  double *Add(const std::string &key, size_t width) {
    blocks.push_back({key, {width, std::vector<double>(tree_count * width, 0.)}});
    return blocks.back().second.second.data();
  }
  std::vector<PhyloGradient> Collect() const {
    std::vector<PhyloGradient> results(tree_count);
    for (size_t tree_idx = 0; tree_idx < tree_count; tree_idx++) {
      GradientMap gradient;
      for (const auto &[key, block] : blocks) {
        const auto &[width, values] = block;
        gradient[key] = std::vector<double>(values.begin() + tree_idx * width,
                                            values.begin() + (tree_idx + 1) * width);
      }
      results[tree_idx] = PhyloGradient(log_likelihood[tree_idx], gradient);
    }
    return results;
  }
};

}  // namespace

Engine::Engine(const EngineSpecification &engine_specification,
               const PhyloModelSpecification &specification, SitePattern site_pattern)
    : site_pattern_(std::move(site_pattern)),
      phylo_model_(PhyloModel::OfSpecification(specification)) {
  if (engine_specification.thread_count_ == 0) {
    Failwith("Thread count needs to be strictly positive.");
  }
  const auto &patterns = site_pattern_.GetPatterns();
  const size_t taxon_count = site_pattern_.SequenceCount();
  const size_t pattern_count = site_pattern_.PatternCount();
  std::vector<uint8_t> tip_states(taxon_count * pattern_count);
  for (size_t taxon = 0; taxon < taxon_count; taxon++) {
    for (size_t pattern = 0; pattern < pattern_count; pattern++) {
      const int symbol = patterns[taxon][pattern];
      tip_states[taxon * pattern_count + pattern] =
          static_cast<uint8_t>((symbol < 0 || symbol > 4) ? 4 : symbol);
    }
  }
  // thread_count_ was the number of BEAGLE instances the reference fanned a collection
  // over (engine.cpp:17-27); here it is the number of GPUs: the first
  // min(thread_count_, visible devices) ordinals, or the list in SBNB_DEVICES
  // ("0,2,3").  One device -> a plain engine; several -> a device group inside
  // libsbn_b200.so (one host thread + stream per GPU), sharded by tree, or by site
  // pattern with SBNB_SHARD_AXIS=patterns.
  std::vector<int32_t> devices;
  if (const char *list = std::getenv("SBNB_DEVICES")) {
    std::stringstream stream(list);
    std::string item;
    while (std::getline(stream, item, ',')) {
      if (!item.empty()) devices.push_back(static_cast<int32_t>(std::stoi(item)));
    }
  } else {
    const size_t visible = static_cast<size_t>(std::max(sbnb_device_count(), 1));
    for (size_t device = 0; device < std::min(engine_specification.thread_count_, visible);
         device++) {
      devices.push_back(static_cast<int32_t>(device));
    }
  }
  if (devices.empty()) {
    Failwith("SBNB_DEVICES names no CUDA device.");
  }
  const char *axis = std::getenv("SBNB_SHARD_AXIS");
  const bool by_pattern = axis != nullptr && std::string(axis) == "patterns";
  if (devices.size() == 1) {
    Check(sbnb_engine_create(specification.substitution_.c_str(), specification.site_.c_str(),
                             specification.clock_.c_str(), static_cast<int32_t>(taxon_count),
                             static_cast<int64_t>(pattern_count), tip_states.data(),
                             site_pattern_.GetWeights().data(), devices[0], &device_engine_));
  } else {
    Check(sbnb_engine_create_multi(
        specification.substitution_.c_str(), specification.site_.c_str(),
        specification.clock_.c_str(), static_cast<int32_t>(taxon_count),
        static_cast<int64_t>(pattern_count), tip_states.data(), site_pattern_.GetWeights().data(),
        devices.data(), static_cast<int32_t>(devices.size()),
        by_pattern ? SBNB_SHARD_PATTERNS : SBNB_SHARD_TREES, &device_engine_));
  }
  Assert(static_cast<size_t>(sbnb_engine_param_count(device_engine_)) ==
             phylo_model_->GetBlockSpecification().ParameterCount(),
         "The device engine and PhyloModel disagree about the parameter count.");
  if (!engine_specification.beagle_flag_vector_.empty()) {
    std::cout << "BEAGLE flags are ignored: likelihoods run on the libsbn_b200 CUDA engine."
              << std::endl;
  }
}

Engine::~Engine() { sbnb_engine_destroy(device_engine_); }

const BlockSpecification &Engine::GetPhyloModelBlockSpecification() const {
  return phylo_model_->GetBlockSpecification();
}

std::vector<double> Engine::LogLikelihoods(const UnrootedTreeCollection &tree_collection,
                                           const EigenMatrixXdRef phylo_model_params,
                                           const bool rescaling) const {
  FlatTrees trees;
  trees.AddTopologiesAndBranchLengths(tree_collection);
  const auto params = ContiguousParams(phylo_model_params, tree_collection.TreeCount());
  std::vector<double> results(tree_collection.TreeCount());
  int status = SBNB_OK;
  {
    ScopedGilRelease let_python_threads_run;
    status = sbnb_log_likelihoods_unrooted(device_engine_, &trees.batch, params.data(), rescaling,
                                      results.data());
  }
  Check(status);
  return results;
}

std::vector<double> Engine::LogLikelihoods(const RootedTreeCollection &tree_collection,
                                           const EigenMatrixXdRef phylo_model_params,
                                           const bool rescaling) const {
  FlatTrees trees;
  trees.AddTopologiesAndBranchLengths(tree_collection);
  trees.AddTimeTreeExtras(tree_collection);
  const auto params = ContiguousParams(phylo_model_params, tree_collection.TreeCount());
  std::vector<double> results(tree_collection.TreeCount());
  int status = SBNB_OK;
  {
    ScopedGilRelease let_python_threads_run;
    status = sbnb_log_likelihoods_rooted(device_engine_, &trees.batch, params.data(), rescaling,
                                    results.data());
  }
  Check(status);
  return results;
}

std::vector<double> Engine::UnrootedLogLikelihoods(
    const RootedTreeCollection &tree_collection, const EigenMatrixXdRef phylo_model_params,
    const bool rescaling) const {
  FlatTrees trees;
  trees.AddTopologiesAndBranchLengths(tree_collection);
  const auto params = ContiguousParams(phylo_model_params, tree_collection.TreeCount());
  std::vector<double> results(tree_collection.TreeCount());
  int status = SBNB_OK;
  {
    ScopedGilRelease let_python_threads_run;
    status = sbnb_unrooted_log_likelihoods_of_rooted(device_engine_, &trees.batch, params.data(),
                                                rescaling, results.data());
  }
  Check(status);
  return results;
}

std::vector<PhyloGradient> Engine::Gradients(const UnrootedTreeCollection &tree_collection,
                                             const EigenMatrixXdRef phylo_model_params,
                                             const bool rescaling) const {
  const size_t tree_count = tree_collection.TreeCount();
  FlatTrees trees;
  trees.AddTopologiesAndBranchLengths(tree_collection);
  const auto params = ContiguousParams(phylo_model_params, tree_count);
  GradientBlocks blocks(tree_count);
  // fat_beagle.cpp:479-500: which blocks exist depends on the model alone.
  const size_t substitution_rate_count = phylo_model_->GetSubstitutionModel()->GetRates().size();
  if (substitution_rate_count > 0) {
    blocks.out.substitution_model =
        blocks.Add("substitution_model",
                   substitution_rate_count - 1 +
                       phylo_model_->GetSubstitutionModel()->GetFrequencies().size() - 1);
  }
  if (phylo_model_->GetSiteModel()->GetCategoryCount() > 1) {
    blocks.out.site_model = blocks.Add("site_model", 1);
  }
  // Detrifurcate adds a node (unrooted_tree.cpp:27-37).
  blocks.out.branch_lengths =
      blocks.Add("branch_lengths", 2 * site_pattern_.SequenceCount() - 1);
  int status = SBNB_OK;
  {
    ScopedGilRelease let_python_threads_run;
    status = sbnb_gradients_unrooted(device_engine_, &trees.batch, params.data(), rescaling,
                                &blocks.out);
  }
  Check(status);
  return blocks.Collect();
}

std::vector<PhyloGradient> Engine::Gradients(const RootedTreeCollection &tree_collection,
                                             const EigenMatrixXdRef phylo_model_params,
                                             const bool rescaling) const {
  const size_t tree_count = tree_collection.TreeCount();
  FlatTrees trees;
  trees.AddTopologiesAndBranchLengths(tree_collection);
  trees.AddTimeTreeExtras(tree_collection);
  const auto params = ContiguousParams(phylo_model_params, tree_count);
  GradientBlocks blocks(tree_count);
  const size_t substitution_rate_count = phylo_model_->GetSubstitutionModel()->GetRates().size();
  if (substitution_rate_count > 0) {
    blocks.out.substitution_model =
        blocks.Add("substitution_model",
                   substitution_rate_count - 1 +
                       phylo_model_->GetSubstitutionModel()->GetFrequencies().size() - 1);
  }
  if (phylo_model_->GetSiteModel()->GetCategoryCount() > 1) {
    blocks.out.site_model = blocks.Add("site_model", 1);
  }
  blocks.out.ratios_root_height =
      blocks.Add("ratios_root_height", site_pattern_.SequenceCount() - 1);
  blocks.out.clock_model =
      blocks.Add("clock_model", static_cast<size_t>(std::max(trees.batch.rate_count, 0)));
  int status = SBNB_OK;
  {
    ScopedGilRelease let_python_threads_run;
    status = sbnb_gradients_rooted(device_engine_, &trees.batch, params.data(), rescaling,
                              &blocks.out);
  }
  Check(status);
  return blocks.Collect();
}
