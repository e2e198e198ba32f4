// oracle/gp_dump.cpp -- TEST INFRASTRUCTURE (fixture generator).
//
// Drives the UNMODIFIED reference generalized-pruning code (GPInstance, GPDAG,
// GPEngine under /root/reference/src, compiled into oracle/_ref/obj by
// oracle/Makefile) through the scenarios of the reference's own gp_doctest.cpp
// and prints, as JSON, everything a replacement engine needs to be checked
// against it: the inputs GPInstance::MakeEngine hands to GPEngine
// (gp_instance.cpp:68-83), the GPOperation programs GPDAG schedules
// (gp_dag.cpp:94-263) flattened with the record layout of
// include/sbn_b200_gp.h, and the reference engine's state after each stage.
//
// Built with -fno-access-control so it can read GPEngine's private PLVs and
// rescaling counts; nothing of the reference is copied or modified.
//
// usage: gp_dump <scenario> <fasta> <newick> [options]   (see tests/golden/make_gp_fixtures.py)

#include <cmath>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <sstream>
#include <string>
#include <variant>
#include <vector>

#include "gp_instance.hpp"

namespace {

using Program = std::vector<int>;

// Record layout of include/sbn_b200_gp.h: opcode, then the op's size_t fields
// in declaration order (gp_operation.hpp:25-171).
struct Flatten {
  Program& out;
  void operator()(const GPOperations::ZeroPLV& op) { Put({0, op.dest_}); }
  void operator()(const GPOperations::SetToStationaryDistribution& op) {
    Put({1, op.dest_, op.root_gpcsp_idx_});
  }
  void operator()(const GPOperations::IncrementWithWeightedEvolvedPLV& op) {
    Put({2, op.dest_, op.gpcsp_, op.src_});
  }
  void operator()(const GPOperations::Multiply& op) { Put({3, op.dest_, op.src1_, op.src2_}); }
  void operator()(const GPOperations::Likelihood& op) { Put({4, op.dest_, op.child_, op.parent_}); }
  void operator()(const GPOperations::OptimizeBranchLength& op) {
    Put({5, op.leafward_, op.rootward_, op.gpcsp_});
  }
  void operator()(const GPOperations::UpdateSBNProbabilities& op) { Put({6, op.start_, op.stop_}); }
  void operator()(const GPOperations::ResetMarginalLikelihood&) { Put({7}); }
  void operator()(const GPOperations::IncrementMarginalLikelihood& op) {
    Put({8, op.stationary_times_prior_, op.rootsplit_, op.p_});
  }
  void operator()(const GPOperations::PrepForMarginalization& op) {
    Put({9, op.dest_, op.src_vector_.size()});
    for (size_t src : op.src_vector_) out.push_back(static_cast<int>(src));
  }
  void Put(std::initializer_list<size_t> words) {
    for (size_t w : words) out.push_back(static_cast<int>(w));
  }
};

Program FlattenProgram(const GPOperationVector& operations) {
  Program program;
  Flatten flatten{program};
  for (const auto& op : operations) std::visit(flatten, op);
  return program;
}

// ---- JSON printing -------------------------------------------------------------
bool first_key = true;
void Key(const std::string& name) {
  std::printf("%s\"%s\": ", first_key ? "" : ",\n", name.c_str());
  first_key = false;
}
void Number(double x) {
  if (std::isnan(x)) {
    std::printf("NaN");
  } else if (std::isinf(x)) {
    std::printf(x > 0 ? "Infinity" : "-Infinity");
  } else {
    std::printf("%.17g", x);
  }
}
template <class Vector>
void Doubles(const std::string& name, const Vector& v) {
  Key(name);
  std::printf("[");
  for (Eigen::Index i = 0; i < static_cast<Eigen::Index>(v.size()); i++) {
    if (i) std::printf(",");
    Number(v[i]);
  }
  std::printf("]");
}
void Scalar(const std::string& name, double x) {
  Key(name);
  Number(x);
}
template <class Vector>
void Ints(const std::string& name, const Vector& v) {
  Key(name);
  std::printf("[");
  for (size_t i = 0; i < static_cast<size_t>(v.size()); i++) std::printf("%s%ld", i ? "," : "", static_cast<long>(v[i]));
  std::printf("]");
}

void DumpEngineState(const std::string& prefix, GPEngine* engine, bool with_plvs) {
  Doubles(prefix + "branch_lengths", engine->GetBranchLengths());
  Doubles(prefix + "q", engine->GetSBNParameters());
  Scalar(prefix + "log_marginal_likelihood", engine->GetLogMarginalLikelihood());
  Doubles(prefix + "per_gpcsp_log_likelihoods", engine->GetPerGPCSPLogLikelihoods());
  Doubles(prefix + "per_gpcsp_components_of_full_log_marginal", engine->GetPerGPCSPComponentsOfFullLogMarginal());
  Ints(prefix + "rescaling_counts", engine->rescaling_counts_);
  if (with_plvs) {
    // [plv][pattern][state] -- the reference's memory order (mmapped_plv.hpp:15-41).
    std::vector<double> flat;
    for (const auto& plv : engine->plvs_)
      for (Eigen::Index k = 0; k < plv.cols(); k++)
        for (Eigen::Index i = 0; i < 4; i++) flat.push_back(plv(i, k));
    Doubles(prefix + "plvs", flat);
    const auto& matrix = engine->log_likelihoods_;
    std::vector<double> rows;
    for (Eigen::Index g = 0; g < matrix.rows(); g++)
      for (Eigen::Index k = 0; k < matrix.cols(); k++) rows.push_back(matrix(g, k));
    Doubles(prefix + "log_likelihood_matrix", rows);
  }
}

void TipList(std::vector<int>& out, const QuartetTipVector& tips) {
  for (const auto& tip : tips) {
    out.push_back(static_cast<int>(tip.tip_node_id_));
    out.push_back(static_cast<int>(tip.plv_idx_));
    out.push_back(static_cast<int>(tip.gpcsp_idx_));
  }
}

}  // namespace

int main(int argc, char** argv) {
  if (argc < 4) {
    std::fprintf(stderr, "usage: gp_dump <scenario> <fasta> <newick> [key=value ...]\n");
    return 2;
  }
  const std::string scenario = argv[1], fasta = argv[2], newick = argv[3];
  double threshold = GPEngine::default_rescaling_threshold_;
  double constant_branch_length = -1.0, tol = 1e-4;
  size_t max_iter = 100;
  bool with_plvs = false;
  std::vector<double> branch_lengths;
  std::vector<int> quartet;  // parent, rotated, child
  for (int i = 4; i < argc; i++) {
    std::string arg = argv[i];
    const auto eq = arg.find('=');
    const std::string key = arg.substr(0, eq), value = arg.substr(eq + 1);
    if (key == "threshold") threshold = std::stod(value);
    if (key == "constant") constant_branch_length = std::stod(value);
    if (key == "tol") tol = std::stod(value);
    if (key == "max_iter") max_iter = std::stoul(value);
    if (key == "plvs") with_plvs = value == "1";
    if (key == "branch_lengths" || key == "quartet") {
      std::stringstream stream(value);
      std::string item;
      while (std::getline(stream, item, ',')) {
        if (key == "quartet")
          quartet.push_back(std::stoi(item));
        else
          branch_lengths.push_back(std::stod(item));
      }
    }
  }

  std::stringstream sink;
  auto* saved = std::cout.rdbuf(sink.rdbuf());  // the reference chats on stdout
  GPInstance inst("/tmp/gp_dump_mmapped_plv.data");
  inst.ReadFastaFile(fasta);
  inst.ReadNewickFile(newick);
  inst.MakeEngine(threshold);
  GPEngine* engine = inst.GetEngine();
  if (constant_branch_length > 0) engine->SetBranchLengthsToConstant(constant_branch_length);
  if (!branch_lengths.empty()) {
    EigenVectorXd v(branch_lengths.size());
    for (size_t i = 0; i < branch_lengths.size(); i++) v[i] = branch_lengths[i];
    engine->SetBranchLengths(v);
  }
  GPDAG& dag = inst.dag_;  // non-const: BranchLengthOptimization() updates its clean/dirty marks

  std::printf("{");
  Key("scenario");
  std::printf("\"%s\"", scenario.c_str());
  // ---- what MakeEngine hands to the engine (gp_instance.cpp:68-83)
  const auto& site_pattern = engine->site_pattern_;
  std::vector<int> tips;
  for (const auto& row : site_pattern.GetPatterns())
    for (int symbol : row) tips.push_back(symbol);
  Ints("tip_states", tips);
  Scalar("taxon_count", static_cast<double>(site_pattern.SequenceCount()));
  Scalar("pattern_count", static_cast<double>(site_pattern.PatternCount()));
  Scalar("site_count", static_cast<double>(site_pattern.SiteCount()));
  Doubles("pattern_weights", site_pattern.GetWeights());
  Scalar("node_count", static_cast<double>(dag.NodeCount()));
  Scalar("plv_count", static_cast<double>(engine->plv_count_));
  Scalar("gpcsp_count", static_cast<double>(dag.GPCSPCountWithFakeSubsplits()));
  Scalar("rootsplit_count", static_cast<double>(dag.RootsplitCount()));
  Scalar("rescaling_threshold", threshold);
  Doubles("sbn_prior", engine->q_);
  Doubles("unconditional_node_probabilities", engine->unconditional_node_probabilities_);
  Doubles("inverted_sbn_prior", engine->inverted_sbn_prior_);
  Doubles("initial_branch_lengths", engine->GetBranchLengths());
  // ---- the programs (gp_dag.cpp)
  const GPOperationVector populate = dag.PopulatePLVs();
  const GPOperationVector likelihoods = dag.ComputeLikelihoods();
  const GPOperationVector marginal = dag.MarginalLikelihood();
  const GPOperationVector branch_optimization = dag.BranchLengthOptimization();
  const GPOperationVector sbn_optimization = dag.OptimizeSBNParameters();
  Ints("program_populate_plvs", FlattenProgram(populate));
  Ints("program_compute_likelihoods", FlattenProgram(likelihoods));
  Ints("program_marginal_likelihood", FlattenProgram(marginal));
  Ints("program_branch_length_optimization", FlattenProgram(branch_optimization));
  Ints("program_optimize_sbn_parameters", FlattenProgram(sbn_optimization));

  // ---- stage 1: PopulatePLVs + ComputeLikelihoods (gp_doctest.cpp:89-101)
  engine->ProcessOperations(populate);
  engine->ProcessOperations(likelihoods);
  DumpEngineState("populated_", engine, with_plvs);

  if (scenario == "gradient") {  // gp_doctest.cpp:218-241
    const size_t node_count = dag.NodeCount();
    const size_t leafward = GPDAG::GetPLVIndexStatic(GPDAG::PLVType::P, node_count, 0);
    const size_t rootward = GPDAG::GetPLVIndexStatic(GPDAG::PLVType::R, node_count, node_count - 1);
    GPOperations::OptimizeBranchLength op{leafward, rootward, 2};
    const auto [log_likelihood, derivative] = engine->LogLikelihoodAndDerivative(op);
    Ints("gradient_op", std::vector<int>{static_cast<int>(leafward), static_cast<int>(rootward), 2});
    Doubles("gradient_log_likelihood_and_derivative", std::vector<double>{log_likelihood, derivative});
  }

  if (scenario == "quartet") {  // gp_doctest.cpp:522-587
    const auto request = dag.QuartetHybridRequestOf(quartet.at(0), quartet.at(1) != 0, quartet.at(2));
    std::vector<int> request_tips;
    TipList(request_tips, request.rootward_tips_);
    TipList(request_tips, request.sister_tips_);
    TipList(request_tips, request.rotated_tips_);
    TipList(request_tips, request.sorted_tips_);
    Ints("quartet_tips", request_tips);
    Ints("quartet_counts", std::vector<int>{static_cast<int>(request.rootward_tips_.size()),
                                            static_cast<int>(request.sister_tips_.size()),
                                            static_cast<int>(request.rotated_tips_.size()),
                                            static_cast<int>(request.sorted_tips_.size())});
    Scalar("quartet_central_gpcsp", static_cast<double>(request.central_gpcsp_idx_));
    Doubles("quartet_log_likelihoods", engine->CalculateQuartetHybridLikelihoods(request));
    // CalculateHybridMarginals (gp_instance.cpp:184-192): every fully formed request of the DAG.
    std::vector<int> all_requests;  // records: central, 4 counts, then the tips
    dag.ReversePostorderIndexTraversal(
        [&](const size_t parent_id, const bool rotated, const size_t child_id, const size_t) {
          const auto r = dag.QuartetHybridRequestOf(parent_id, rotated, child_id);
          engine->ProcessQuartetHybridRequest(r);
          if (!r.IsFullyFormed()) return;
          all_requests.push_back(static_cast<int>(r.central_gpcsp_idx_));
          all_requests.push_back(static_cast<int>(r.rootward_tips_.size()));
          all_requests.push_back(static_cast<int>(r.sister_tips_.size()));
          all_requests.push_back(static_cast<int>(r.rotated_tips_.size()));
          all_requests.push_back(static_cast<int>(r.sorted_tips_.size()));
          TipList(all_requests, r.rootward_tips_);
          TipList(all_requests, r.sister_tips_);
          TipList(all_requests, r.rotated_tips_);
          TipList(all_requests, r.sorted_tips_);
        });
    Ints("hybrid_requests", all_requests);
    Doubles("hybrid_marginals", engine->GetHybridMarginals());
  }

  if (scenario == "estimate") {
    // GPInstance::EstimateBranchLengths (gp_instance.cpp:129-175), replayed from
    // the dumped programs so that the fixture and the programs agree exactly.
    engine->ProcessOperations(populate);
    engine->ProcessOperations(marginal);
    double current = engine->GetLogMarginalLikelihood();
    std::vector<double> trace{current};
    size_t iterations = 0;
    for (size_t i = 0; i < max_iter; i++) {
      engine->ProcessOperations(branch_optimization);
      engine->ProcessOperations(populate);
      engine->ProcessOperations(marginal);
      const double updated = engine->GetLogMarginalLikelihood();
      trace.push_back(updated);
      iterations++;
      if (i == 0) Doubles("first_iteration_branch_lengths", engine->GetBranchLengths());
      if (std::abs(current - updated) < tol) break;
      current = updated;
    }
    Doubles("estimate_marginal_trace", trace);
    Scalar("estimate_iterations", static_cast<double>(iterations));
    Scalar("estimate_tol", tol);
    Scalar("estimate_max_iter", static_cast<double>(max_iter));
    // the tail of TestCompositeMarginal (gp_doctest.cpp:177-182)
    engine->ProcessOperations(populate);
    engine->ProcessOperations(likelihoods);
    engine->ProcessOperations(marginal);
    DumpEngineState("estimated_", engine, false);
    // EstimateSBNParameters (gp_instance.cpp:177-182)
    engine->ProcessOperations(populate);
    engine->ProcessOperations(likelihoods);
    engine->ProcessOperations(sbn_optimization);
    Doubles("estimated_sbn_parameters", engine->GetSBNParameters());
    // cross-check against the reference's own driver loop on a fresh instance
    GPInstance check("/tmp/gp_dump_mmapped_plv_check.data");
    check.ReadFastaFile(fasta);
    check.ReadNewickFile(newick);
    check.MakeEngine(threshold);
    if (constant_branch_length > 0) check.GetEngine()->SetBranchLengthsToConstant(constant_branch_length);
    check.EstimateBranchLengths(tol, max_iter, true);
    Doubles("estimate_branch_lengths_by_reference_driver", check.GetEngine()->GetBranchLengths());
  }
  std::printf("}\n");
  std::cout.rdbuf(saved);
  return 0;
}
