// integration/engine.hpp -- the reference-side binding of libsbn_b200's outer
// boundary: a replacement for phylovi/libsbn's src/engine.hpp with the SAME public
// interface (EngineSpecification, Engine and its five batch methods,
// reference src/engine.hpp:20-53) whose implementation calls the C ABI of
// include/sbn_b200.h instead of a thread pool of BEAGLE-backed FatBeagles.
//
// A maintainer drops this file and engine.cpp over src/engine.{hpp,cpp}, removes
// src/fat_beagle.cpp from the build and links libsbn_b200.so instead of
// libhmsbeagle: the instances, pylibsbn.cpp and vip stay as they are.  INTEGRATION.md
// has the recipe; integration/Makefile builds the reference's doctest, gp_doctest
// and python module that way without modifying or copying a reference source.
#ifndef SRC_ENGINE_HPP_
#define SRC_ENGINE_HPP_

#include <memory>
#include <utility>
#include <vector>

#include "libhmsbeagle/beagle.h"  // only for the BeagleFlags enum of the python API
#include "phylo_model.hpp"
#include "rooted_tree_collection.hpp"
#include "site_pattern.hpp"
#include "task_processor.hpp"  // unused here; keeps its doctest case in the reference suite
#include "tree_gradient.hpp"
#include "unrooted_tree_collection.hpp"

struct sbnb_engine;

// Same three fields as the reference (generic_sbn_instance.hpp:252 brace-initialises
// it).  thread_count_ -- the reference's number of BEAGLE instances -- is the number of
// GPUs the collection is fanned over (engine.cpp); the BEAGLE flags are accepted and
// ignored; use_tip_states_ selects between two BEAGLE code paths with identical results,
// of which the device has only one.
struct EngineSpecification {
  const size_t thread_count_;
  const std::vector<BeagleFlags> &beagle_flag_vector_;
  const bool use_tip_states_;
};

class Engine {
 public:
  Engine(const EngineSpecification &engine_specification,
         const PhyloModelSpecification &specification, SitePattern site_pattern);
  ~Engine();
  Engine(const Engine &) = delete;
  Engine &operator=(const Engine &) = delete;

  const BlockSpecification &GetPhyloModelBlockSpecification() const;

  std::vector<double> LogLikelihoods(const UnrootedTreeCollection &tree_collection,
                                     const EigenMatrixXdRef phylo_model_params,
                                     const bool rescaling) const;
  std::vector<double> LogLikelihoods(const RootedTreeCollection &tree_collection,
                                     const EigenMatrixXdRef phylo_model_params,
                                     const bool rescaling) const;
  std::vector<double> UnrootedLogLikelihoods(const RootedTreeCollection &tree_collection,
                                             const EigenMatrixXdRef phylo_model_params,
                                             const bool rescaling) const;
  std::vector<PhyloGradient> Gradients(const UnrootedTreeCollection &tree_collection,
                                       const EigenMatrixXdRef phylo_model_params,
                                       const bool rescaling) const;
  std::vector<PhyloGradient> Gradients(const RootedTreeCollection &tree_collection,
                                       const EigenMatrixXdRef phylo_model_params,
                                       const bool rescaling) const;

 private:
  SitePattern site_pattern_;
  // Host-only: answers GetPhyloModelBlockSpecification exactly as the reference's
  // first FatBeagle did (engine.cpp:48-52).
  std::unique_ptr<PhyloModel> phylo_model_;
  sbnb_engine *device_engine_ = nullptr;
};

#endif  // SRC_ENGINE_HPP_
