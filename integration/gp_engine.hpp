// integration/gp_engine.hpp -- the reference-side binding of libsbn_b200's GP
// boundary: a replacement for phylovi/libsbn's src/gp_engine.hpp with the SAME class
// name, constructor and public methods (reference src/gp_engine.hpp:20-91) whose
// implementation forwards to the C ABI of include/sbn_b200_gp.h.  PLVs, branch
// lengths, q and the log-likelihood matrix live in HBM for the life of the engine
// (the reference keeps PLVs in an mmapped file, mmapped_plv.hpp); an operation
// vector is flattened to int32 records and runs as ONE kernel launch where the
// reference interpreted a std::variant at a time with Eigen
// (gp_engine.cpp:167-171).  GPInstance, GPDAG and pylibsbn.cpp stay as they are.
#ifndef SRC_GP_ENGINE_HPP_
#define SRC_GP_ENGINE_HPP_

#include <string>
#include <utility>
#include <vector>

#include "eigen_sugar.hpp"
#include "gp_operation.hpp"
#include "mmapped_plv.hpp"  // unused here; keeps its doctest cases in the reference suite
#include "numerical_utils.hpp"
#include "quartet_hybrid_request.hpp"
#include "rooted_tree_collection.hpp"
#include "sbn_maps.hpp"
#include "site_pattern.hpp"

struct sbnb_gp_engine;

class GPEngine {
 public:
  // mmap_file_path is accepted for signature compatibility and unused: nothing is
  // written to disk.
  GPEngine(SitePattern site_pattern, size_t plv_count, size_t gpcsp_count,
           const std::string &mmap_file_path, double rescaling_threshold,
           EigenVectorXd sbn_prior, EigenVectorXd unconditional_node_probabilities,
           EigenVectorXd inverted_sbn_prior);
  ~GPEngine();
  GPEngine(const GPEngine &) = delete;
  GPEngine &operator=(const GPEngine &) = delete;

  // One operation = a one-record program.
  void operator()(const GPOperations::ZeroPLV &op);
  void operator()(const GPOperations::SetToStationaryDistribution &op);
  void operator()(const GPOperations::IncrementWithWeightedEvolvedPLV &op);
  void operator()(const GPOperations::ResetMarginalLikelihood &op);
  void operator()(const GPOperations::IncrementMarginalLikelihood &op);
  void operator()(const GPOperations::Multiply &op);
  void operator()(const GPOperations::Likelihood &op);
  void operator()(const GPOperations::OptimizeBranchLength &op);
  void operator()(const GPOperations::UpdateSBNProbabilities &op);
  void operator()(const GPOperations::PrepForMarginalization &op);

  void ProcessOperations(GPOperationVector operations);

  void SetTransitionMatrixToHaveBranchLength(double branch_length);
  void SetTransitionAndDerivativeMatricesToHaveBranchLength(double branch_length);
  void SetTransitionMatrixToHaveBranchLengthAndTranspose(double branch_length);
  const Eigen::Matrix4d &GetTransitionMatrix() { return transition_matrix_; };

  void SetBranchLengths(EigenVectorXd branch_lengths);
  void SetBranchLengthsToConstant(double branch_length);
  void ResetLogMarginalLikelihood();
  double GetLogMarginalLikelihood() const;
  EigenVectorXd GetBranchLengths() const;
  EigenVectorXd GetPerGPCSPLogLikelihoods() const;
  EigenVectorXd GetPerGPCSPLogLikelihoods(size_t start, size_t length) const;
  EigenVectorXd GetPerGPCSPComponentsOfFullLogMarginal() const;
  // The three Ref getters return views of host copies refreshed by the call.
  EigenConstMatrixXdRef GetLogLikelihoodMatrix() const;
  EigenConstVectorXdRef GetHybridMarginals() const;
  EigenConstVectorXdRef GetSBNParameters() const;

  EigenVectorXd CalculateQuartetHybridLikelihoods(const QuartetHybridRequest &request);
  void ProcessQuartetHybridRequest(const QuartetHybridRequest &request);

  void PrintPLV(size_t plv_idx);

  void HotStartBranchLengths(const RootedTreeCollection &tree_collection,
                             const BitsetSizeMap &indexer);

  DoublePair LogLikelihoodAndDerivative(const GPOperations::OptimizeBranchLength &op);

  static constexpr double default_rescaling_threshold_ = 1e-40;
  static constexpr double default_branch_length_ = 0.1;

  // Bytes of PLV storage (device memory here).
  double PLVByteCount() const {
    return static_cast<double>(plv_count_ * pattern_count_ * 4 * sizeof(double));
  };

 private:
  size_t plv_count_;
  size_t gpcsp_count_;
  size_t pattern_count_;
  sbnb_gp_engine *device_engine_ = nullptr;
  Eigen::Matrix4d transition_matrix_;
  mutable EigenMatrixXd log_likelihoods_;
  mutable EigenVectorXd hybrid_marginal_log_likelihoods_;
  mutable EigenVectorXd q_;

  void Run(const std::vector<int32_t> &program);
};

#ifdef DOCTEST_LIBRARY_INCLUDED

TEST_CASE("GPEngine") {
  // JC69 at t = 0.75: 1/4 + 3/4 e^{-4t/3} on the diagonal, 1/4 - 1/4 e^{-4t/3} off it.
  EigenVectorXd nothing;
  GPEngine engine(SitePattern::HelloSitePattern(), 6 * 5, 5, "_ignore/mmapped_plv.data",
                  GPEngine::default_rescaling_threshold_, nothing, nothing, nothing);
  engine.SetTransitionMatrixToHaveBranchLength(0.75);
  CHECK(fabs(0.52590958087 - engine.GetTransitionMatrix()(0, 0)) < 1e-10);
  CHECK(fabs(0.1580301397 - engine.GetTransitionMatrix()(0, 1)) < 1e-10);
}

#endif  // DOCTEST_LIBRARY_INCLUDED

#endif  // SRC_GP_ENGINE_HPP_
