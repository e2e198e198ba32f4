// integration/gp_engine.cpp -- replacement body of phylovi/libsbn's
// src/gp_engine.cpp: every public method of GPEngine forwards to the C ABI of
// include/sbn_b200_gp.h.  The arithmetic of the reference's operator() overloads
// (gp_engine.cpp:48-165), Brent optimisation (gp_engine.cpp:326-345) and the quartet
// hybrid likelihoods (gp_engine.cpp:396-452) runs in the device interpreter of
// libsbn_b200/csrc/gp_engine.cu; what is left here is flattening the host objects.
// Failures come back as Failwith -> std::runtime_error (sugar.hpp:67-78), with the
// reference's own message where the failure is one of its Asserts.

#include "gp_engine.hpp"

#include <iostream>
#include <limits>

#include "sbn_b200_gp.h"
#include "scoped_gil_release.hpp"
#include "sugar.hpp"

namespace {

void Check(int status) {
  if (status != SBNB_OK) {
    const std::string message = sbnb_last_error();
    // Asserts of the reference's ops carry the reference's message verbatim.
    Failwith(status == SBNB_ERR_GP_ASSERT ? message : "libsbn_b200: " + message);
  }
}

int32_t Word(size_t value) {
  if (value > static_cast<size_t>(std::numeric_limits<int32_t>::max())) {
    Failwith("GP operation index does not fit the int32 program encoding.");
  }
  return static_cast<int32_t>(value);
}

// The record of each GPOperation (gp_operation.hpp:25-171) as documented in
// sbn_b200_gp.h: opcode, then the size_t fields in declaration order.
struct Encoder {
  std::vector<int32_t> &words;

  void operator()(const GPOperations::ZeroPLV &op) {
    words.insert(words.end(), {SBNB_GP_ZERO_PLV, Word(op.dest_)});
  }
  void operator()(const GPOperations::SetToStationaryDistribution &op) {
    words.insert(words.end(),
                 {SBNB_GP_SET_TO_STATIONARY, Word(op.dest_), Word(op.root_gpcsp_idx_)});
  }
  void operator()(const GPOperations::IncrementWithWeightedEvolvedPLV &op) {
    words.insert(words.end(), {SBNB_GP_INCREMENT_WITH_EVOLVED, Word(op.dest_),
                               Word(op.gpcsp_), Word(op.src_)});
  }
  void operator()(const GPOperations::ResetMarginalLikelihood &) {
    words.push_back(SBNB_GP_RESET_MARGINAL_LIKELIHOOD);
  }
  void operator()(const GPOperations::IncrementMarginalLikelihood &op) {
    words.insert(words.end(), {SBNB_GP_INCREMENT_MARGINAL, Word(op.stationary_times_prior_),
                               Word(op.rootsplit_), Word(op.p_)});
  }
  void operator()(const GPOperations::Multiply &op) {
    words.insert(words.end(),
                 {SBNB_GP_MULTIPLY, Word(op.dest_), Word(op.src1_), Word(op.src2_)});
  }
  void operator()(const GPOperations::Likelihood &op) {
    words.insert(words.end(),
                 {SBNB_GP_LIKELIHOOD, Word(op.dest_), Word(op.child_), Word(op.parent_)});
  }
  void operator()(const GPOperations::OptimizeBranchLength &op) {
    words.insert(words.end(), {SBNB_GP_OPTIMIZE_BRANCH_LENGTH, Word(op.leafward_),
                               Word(op.rootward_), Word(op.gpcsp_)});
  }
  void operator()(const GPOperations::UpdateSBNProbabilities &op) {
    words.insert(words.end(),
                 {SBNB_GP_UPDATE_SBN_PROBABILITIES, Word(op.start_), Word(op.stop_)});
  }
  void operator()(const GPOperations::PrepForMarginalization &op) {
    words.insert(words.end(), {SBNB_GP_PREP_FOR_MARGINALIZATION, Word(op.dest_),
                               Word(op.src_vector_.size())});
    for (const size_t src : op.src_vector_) {
      words.push_back(Word(src));
    }
  }
};

template <typename TOperation>
std::vector<int32_t> EncodeOne(const TOperation &op) {
  std::vector<int32_t> words;
  Encoder{words}(op);
  return words;
}

// QuartetTipVector -> {tip_node_id, plv_idx, gpcsp_idx} int32 triples.
std::vector<int32_t> FlatTips(const QuartetTipVector &tips) {
  std::vector<int32_t> flat;
  flat.reserve(3 * tips.size());
  for (const auto &tip : tips) {
    flat.insert(flat.end(), {Word(tip.tip_node_id_), Word(tip.plv_idx_), Word(tip.gpcsp_idx_)});
  }
  return flat;
}

const double *DataOrNull(const EigenVectorXd &vector) {
  return vector.size() == 0 ? nullptr : vector.data();
}

}  // namespace

GPEngine::GPEngine(SitePattern site_pattern, size_t plv_count, size_t gpcsp_count,
                   const std::string & /*mmap_file_path*/, double rescaling_threshold,
                   EigenVectorXd sbn_prior, EigenVectorXd unconditional_node_probabilities,
                   EigenVectorXd inverted_sbn_prior)
    : plv_count_(plv_count),
      gpcsp_count_(gpcsp_count),
      pattern_count_(site_pattern.PatternCount()) {
  const size_t taxon_count = site_pattern.SequenceCount();
  // Symbols as InitializePLVsWithSitePatterns reads them (gp_engine.cpp:268-286):
  // 0..3 a one-hot column, 4 (the gap) a column of ones.
  std::vector<uint8_t> tip_states;
  tip_states.reserve(taxon_count * pattern_count_);
  for (const auto &pattern : site_pattern.GetPatterns()) {
    for (const int symbol : pattern) {
      Assert(symbol >= 0, "Negative symbol!");
      Assert(symbol <= 4, "Symbol outside the nucleotide alphabet.");
      tip_states.push_back(static_cast<uint8_t>(symbol));
    }
  }
  Assert(sbn_prior.size() == 0 || static_cast<size_t>(sbn_prior.size()) == gpcsp_count,
         "The SBN prior needs one entry per GPCSP.");
  Assert(inverted_sbn_prior.size() == 0 ||
             static_cast<size_t>(inverted_sbn_prior.size()) == gpcsp_count,
         "The inverted SBN prior needs one entry per GPCSP.");
  Check(sbnb_gp_create(Word(taxon_count), static_cast<int64_t>(pattern_count_),
                       tip_states.data(), site_pattern.GetWeights().data(),
                       static_cast<int64_t>(site_pattern.SiteCount()), Word(plv_count),
                       Word(gpcsp_count), rescaling_threshold, DataOrNull(sbn_prior),
                       DataOrNull(unconditional_node_probabilities),
                       Word(static_cast<size_t>(unconditional_node_probabilities.size())),
                       DataOrNull(inverted_sbn_prior), /*device=*/0, &device_engine_));
  transition_matrix_.setZero();
}

GPEngine::~GPEngine() { sbnb_gp_destroy(device_engine_); }

void GPEngine::Run(const std::vector<int32_t> &program) {
  int status = SBNB_OK;
  {
    ScopedGilRelease let_python_threads_run;
    status = sbnb_gp_process_operations(device_engine_, program.data(),
                                        static_cast<int64_t>(program.size()));
  }
  Check(status);
}

void GPEngine::operator()(const GPOperations::ZeroPLV &op) { Run(EncodeOne(op)); }
void GPEngine::operator()(const GPOperations::SetToStationaryDistribution &op) {
  Run(EncodeOne(op));
}
void GPEngine::operator()(const GPOperations::IncrementWithWeightedEvolvedPLV &op) {
  Run(EncodeOne(op));
}
void GPEngine::operator()(const GPOperations::ResetMarginalLikelihood &op) {
  Run(EncodeOne(op));
}
void GPEngine::operator()(const GPOperations::IncrementMarginalLikelihood &op) {
  Run(EncodeOne(op));
}
void GPEngine::operator()(const GPOperations::Multiply &op) { Run(EncodeOne(op)); }
void GPEngine::operator()(const GPOperations::Likelihood &op) { Run(EncodeOne(op)); }
void GPEngine::operator()(const GPOperations::OptimizeBranchLength &op) {
  Run(EncodeOne(op));
}
void GPEngine::operator()(const GPOperations::UpdateSBNProbabilities &op) {
  Run(EncodeOne(op));
}
void GPEngine::operator()(const GPOperations::PrepForMarginalization &op) {
  Run(EncodeOne(op));
}

void GPEngine::ProcessOperations(GPOperationVector operations) {
  std::vector<int32_t> program;
  program.reserve(4 * operations.size());
  Encoder encoder{program};
  for (const auto &operation : operations) {
    std::visit(encoder, operation);
  }
  Run(program);
}

void GPEngine::SetTransitionMatrixToHaveBranchLength(double branch_length) {
  double row_major[16];
  Check(sbnb_gp_transition_matrix(device_engine_, branch_length, row_major));
  for (int row = 0; row < 4; row++) {
    for (int col = 0; col < 4; col++) {
      transition_matrix_(row, col) = row_major[4 * row + col];
    }
  }
}

// The derivative matrix has no getter in the reference either; it only feeds
// LogLikelihoodAndDerivative, which the device computes in one call.
void GPEngine::SetTransitionAndDerivativeMatricesToHaveBranchLength(double branch_length) {
  SetTransitionMatrixToHaveBranchLength(branch_length);
}

void GPEngine::SetTransitionMatrixToHaveBranchLengthAndTranspose(double branch_length) {
  SetTransitionMatrixToHaveBranchLength(branch_length);
  transition_matrix_.transposeInPlace();
}

void GPEngine::SetBranchLengths(EigenVectorXd branch_lengths) {
  Assert(static_cast<size_t>(branch_lengths.size()) == gpcsp_count_,
         "Size mismatch in GPEngine::SetBranchLengths.");
  Check(sbnb_gp_set_branch_lengths(device_engine_, branch_lengths.data()));
}

void GPEngine::SetBranchLengthsToConstant(double branch_length) {
  Check(sbnb_gp_set_branch_lengths_to_constant(device_engine_, branch_length));
}

void GPEngine::ResetLogMarginalLikelihood() {
  Check(sbnb_gp_reset_log_marginal_likelihood(device_engine_));
}

double GPEngine::GetLogMarginalLikelihood() const {
  double result = 0.;
  Check(sbnb_gp_get_log_marginal_likelihood(device_engine_, &result));
  return result;
}

EigenVectorXd GPEngine::GetBranchLengths() const {
  EigenVectorXd result(gpcsp_count_);
  Check(sbnb_gp_get_branch_lengths(device_engine_, result.data()));
  return result;
}

EigenVectorXd GPEngine::GetPerGPCSPLogLikelihoods() const {
  return GetPerGPCSPLogLikelihoods(0, gpcsp_count_);
}

EigenVectorXd GPEngine::GetPerGPCSPLogLikelihoods(size_t start, size_t length) const {
  EigenVectorXd result(length);
  Check(sbnb_gp_get_per_gpcsp_log_likelihoods(device_engine_, Word(start), Word(length),
                                              result.data()));
  return result;
}

EigenVectorXd GPEngine::GetPerGPCSPComponentsOfFullLogMarginal() const {
  EigenVectorXd result(gpcsp_count_);
  Check(sbnb_gp_get_per_gpcsp_components_of_full_log_marginal(device_engine_, result.data()));
  return result;
}

EigenConstMatrixXdRef GPEngine::GetLogLikelihoodMatrix() const {
  // EigenMatrixXd is RowMajor (eigen_sugar.hpp:16-17), like the device matrix.
  log_likelihoods_.resize(gpcsp_count_, pattern_count_);
  Check(sbnb_gp_get_log_likelihood_matrix(device_engine_, log_likelihoods_.data()));
  return log_likelihoods_;
}

EigenConstVectorXdRef GPEngine::GetHybridMarginals() const {
  hybrid_marginal_log_likelihoods_.resize(gpcsp_count_);
  Check(sbnb_gp_get_hybrid_marginals(device_engine_, hybrid_marginal_log_likelihoods_.data()));
  return hybrid_marginal_log_likelihoods_;
}

EigenConstVectorXdRef GPEngine::GetSBNParameters() const {
  q_.resize(gpcsp_count_);
  Check(sbnb_gp_get_sbn_parameters(device_engine_, q_.data()));
  return q_;
}

EigenVectorXd GPEngine::CalculateQuartetHybridLikelihoods(const QuartetHybridRequest &request) {
  const auto rootward = FlatTips(request.rootward_tips_), sister = FlatTips(request.sister_tips_),
             rotated = FlatTips(request.rotated_tips_), sorted = FlatTips(request.sorted_tips_);
  EigenVectorXd result(request.rootward_tips_.size() * request.sister_tips_.size() *
                       request.rotated_tips_.size() * request.sorted_tips_.size());
  if (result.size() == 0) {
    return result;
  }
  Check(sbnb_gp_quartet_hybrid_likelihoods(
      device_engine_, Word(request.central_gpcsp_idx_), rootward.data(),
      Word(request.rootward_tips_.size()), sister.data(), Word(request.sister_tips_.size()),
      rotated.data(), Word(request.rotated_tips_.size()), sorted.data(),
      Word(request.sorted_tips_.size()), result.data()));
  return result;
}

void GPEngine::ProcessQuartetHybridRequest(const QuartetHybridRequest &request) {
  if (!request.IsFullyFormed()) {
    return;
  }
  const auto rootward = FlatTips(request.rootward_tips_), sister = FlatTips(request.sister_tips_),
             rotated = FlatTips(request.rotated_tips_), sorted = FlatTips(request.sorted_tips_);
  Check(sbnb_gp_process_quartet_hybrid_request(
      device_engine_, Word(request.central_gpcsp_idx_), rootward.data(),
      Word(request.rootward_tips_.size()), sister.data(), Word(request.sister_tips_.size()),
      rotated.data(), Word(request.rotated_tips_.size()), sorted.data(),
      Word(request.sorted_tips_.size())));
}

void GPEngine::PrintPLV(size_t plv_idx) {
  // The device hands back [pattern][state]; print one state per line as the
  // reference's 4 x P matrix does.
  std::vector<double> plv(4 * pattern_count_);
  Check(sbnb_gp_get_plv(device_engine_, Word(plv_idx), plv.data()));
  for (size_t state = 0; state < 4; state++) {
    for (size_t pattern = 0; pattern < pattern_count_; pattern++) {
      std::cout << "[" << plv[4 * pattern + state] << "]";
    }
    std::cout << std::endl;
  }
  std::cout << std::endl;
}

// Mean branch length per GPCSP over the loaded trees, default where a GPCSP never
// occurs (reference gp_engine.cpp:362-394); the traversal and indexer are the
// reference's host objects, the result goes to the device in one copy.
void GPEngine::HotStartBranchLengths(const RootedTreeCollection &tree_collection,
                                     const BitsetSizeMap &indexer) {
  const size_t taxon_count = tree_collection.TaxonCount();
  const size_t absent = gpcsp_count_;
  std::vector<double> total(gpcsp_count_, 0.);
  std::vector<size_t> occurrences(gpcsp_count_, 0);
  for (const auto &tree : tree_collection.Trees()) {
    tree.Topology()->RootedPCSPPreorder(
        [&](const Node *sister, const Node *focal, const Node *child0, const Node *child1) {
          const Bitset pcsp = SBNMaps::PCSPBitsetOf(taxon_count, sister, false, focal, false,
                                                    child0, false, child1, false);
          const size_t gpcsp_idx = AtWithDefault(indexer, pcsp, absent);
          if (gpcsp_idx != absent) {
            total[gpcsp_idx] += tree.BranchLength(focal);
            occurrences[gpcsp_idx]++;
          }
        });
  }
  EigenVectorXd branch_lengths(gpcsp_count_);
  for (size_t gpcsp_idx = 0; gpcsp_idx < gpcsp_count_; gpcsp_idx++) {
    branch_lengths(gpcsp_idx) =
        occurrences[gpcsp_idx] == 0
            ? default_branch_length_
            : total[gpcsp_idx] / static_cast<double>(occurrences[gpcsp_idx]);
  }
  SetBranchLengths(std::move(branch_lengths));
}

DoublePair GPEngine::LogLikelihoodAndDerivative(const GPOperations::OptimizeBranchLength &op) {
  double result[2];
  Check(sbnb_gp_log_likelihood_and_derivative(device_engine_, Word(op.leafward_),
                                              Word(op.rootward_), Word(op.gpcsp_), result));
  return {result[0], result[1]};
}
