// oracle/phylo_oracle.cpp -- TEST INFRASTRUCTURE (the oracle), never product code.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may call this file.  It is a CPU restatement, on flat
// arrays, of the reference's FatBeagle host logic: the exact sequence of BEAGLE
// calls that src/fat_beagle.cpp issues per tree, driven through the
// BEAGLE-equivalent restatement in oracle/beagle_cpu.cpp (linked into the same
// shared object).  Each function cites the reference lines it follows.
//
// Pinning: the unmodified reference (oracle/_ref, built by oracle/Makefile)
// was run here on the reference's own test inputs and its outputs are
// committed as tests/golden/*.npz (tests/golden/make_fixtures.py);
// tests/test_oracle.py requires this file to reproduce them, together with
// the external pybeagle / physher / phylotorch numbers those fixtures carry.
//
// Deliberate knob: `reference_quirks`.  After its finite-difference loop the
// reference leaves the substitution model at the last "minus delta"
// parameters (fat_beagle.cpp:433-436 call SetParameters with the still
// perturbed vector), so its site-model sweep (fat_beagle.cpp:488-496) runs on a
// model perturbed by ~1e-6.  quirks=1 reproduces that bit for bit; quirks=0
// evaluates the site-model gradient at the unperturbed parameters.

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <numeric>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "libhmsbeagle/beagle.h"

namespace {

thread_local std::string g_error;

[[noreturn]] void Failwith(const std::string& message) { throw std::runtime_error(message); }

// ---------------------------------------------------------------- models ----

struct Model {
  bool gtr = false;   // GTR (10 params) vs JC69 (0 params); HKY is expressed as GTR by the caller
  int categories = 1;  // weibull+K (1 param) vs constant
  bool weibull = false;
  // "rates+K": the row carries K category rates then K d rate / d shape values
  // (for site models the reference does not have; the caller discretises them).
  bool explicit_rates = false;
  // state
  double rates6[6] = {1 / 6., 1 / 6., 1 / 6., 1 / 6., 1 / 6., 1 / 6.};
  double freqs[4] = {0.25, 0.25, 0.25, 0.25};
  double evec[16], ivec[16], eval[4], q[16];
  std::vector<double> cat_rates, cat_weights, cat_rate_derivs;
  double shape = 1.0;
  std::vector<double> given_rates;

  int SubstitutionParamCount() const { return gtr ? 10 : 0; }
  int ParamCount() const {
    return SubstitutionParamCount() + (weibull ? 1 : 0) + (explicit_rates ? 2 * categories : 0);
  }
};

// Largest-off-diagonal Jacobi for a symmetric 4x4.
void SymmetricEigen(double a[4][4], double values[4], double vectors[4][4]) {
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) vectors[i][j] = i == j;
  for (int iteration = 0; iteration < 200; iteration++) {
    int p = 0, q = 1;
    double biggest = 0;
    for (int i = 0; i < 4; i++)
      for (int j = i + 1; j < 4; j++)
        if (std::fabs(a[i][j]) > biggest) biggest = std::fabs(a[i][j]), p = i, q = j;
    if (biggest < 1e-300) break;
    double scale = 0;
    for (int i = 0; i < 4; i++) scale = std::max(scale, std::fabs(a[i][i]));
    if (biggest <= 1e-22 * scale) break;
    const double phi = 0.5 * std::atan2(2 * a[p][q], a[q][q] - a[p][p]);
    const double c = std::cos(phi), s = std::sin(phi);
    for (int k = 0; k < 4; k++) {
      const double akp = a[k][p], akq = a[k][q];
      a[k][p] = c * akp - s * akq, a[k][q] = s * akp + c * akq;
    }
    for (int k = 0; k < 4; k++) {
      const double apk = a[p][k], aqk = a[q][k];
      a[p][k] = c * apk - s * aqk, a[q][k] = s * apk + c * aqk;
    }
    for (int k = 0; k < 4; k++) {
      const double vkp = vectors[k][p], vkq = vectors[k][q];
      vectors[k][p] = c * vkp - s * vkq, vectors[k][q] = s * vkp + c * vkq;
    }
  }
  for (int i = 0; i < 4; i++) values[i] = a[i][i];
}

// substitution_model.cpp:17-80 (GTRModel::SetParameters/UpdateQMatrix/Update),
// substitution_model.hpp:59-74 (JC69Model).
void UpdateSubstitution(Model& m) {
  if (!m.gtr) {
    const double evec[16] = {1.0, 2.0, 0.0, 0.5, 1.0, -2.0, 0.5, 0.0,
                             1.0, 2.0, 0.0, -0.5, 1.0, -2.0, -0.5, 0.0};
    const double ivec[16] = {0.25, 0.25, 0.25, 0.25, 0.125, -0.125, 0.125, -0.125,
                             0.0, 1.0, 0.0, -1.0, 1.0, 0.0, -1.0, 0.0};
    std::memcpy(m.evec, evec, sizeof evec);
    std::memcpy(m.ivec, ivec, sizeof ivec);
    const double eval[4] = {0.0, -1.3333333333333333, -1.3333333333333333, -1.3333333333333333};
    std::memcpy(m.eval, eval, sizeof eval);
    for (int i = 0; i < 4; i++)
      for (int j = 0; j < 4; j++) m.q[i * 4 + j] = i == j ? -1.0 : 1.0 / 3.0;
    for (double& f : m.freqs) f = 0.25;
    return;
  }
  double fsum = 0, rsum = 0;
  for (double f : m.freqs) fsum += f;
  for (double r : m.rates6) rsum += r;
  if (std::fabs(fsum - 1.) >= 0.001) Failwith("GTR frequencies do not sum to 1 +/- 0.001!");
  if (std::fabs(rsum - 1.) >= 0.001) Failwith("GTR rates do not sum to 1 +/- 0.001!");
  double Q[4][4];
  int rate_index = 0;
  for (int i = 0; i < 4; i++)
    for (int j = i + 1; j < 4; j++) {
      const double rate = m.rates6[rate_index++];
      Q[i][j] = rate * m.freqs[j];
      Q[j][i] = rate * m.freqs[i];
    }
  double total = 0;
  for (int i = 0; i < 4; i++) {
    double row_sum = 0;
    for (int j = 0; j < 4; j++)
      if (i != j) row_sum += Q[i][j];
    Q[i][i] = -row_sum;
    total += row_sum * m.freqs[i];
  }
  double S[4][4], values[4], vectors[4][4];
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) {
      Q[i][j] /= total;
      m.q[i * 4 + j] = Q[i][j];
    }
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) S[i][j] = std::sqrt(m.freqs[i]) * Q[i][j] / std::sqrt(m.freqs[j]);
  for (int i = 0; i < 4; i++)
    for (int j = i + 1; j < 4; j++) S[i][j] = S[j][i] = 0.5 * (S[i][j] + S[j][i]);
  SymmetricEigen(S, values, vectors);
  for (int i = 0; i < 4; i++) {
    m.eval[i] = values[i];
    for (int k = 0; k < 4; k++) {
      m.evec[i * 4 + k] = vectors[i][k] / std::sqrt(m.freqs[i]);
      m.ivec[k * 4 + i] = vectors[i][k] * std::sqrt(m.freqs[i]);
    }
  }
}

// site_model.cpp:37-62 (WeibullSiteModel::UpdateRates) / ConstantSiteModel.
void UpdateSite(Model& m) {
  const int C = m.categories;
  m.cat_rates.assign(C, 1.0);
  m.cat_weights.assign(C, 1.0 / C);
  m.cat_rate_derivs.assign(C, 0.0);
  if (m.explicit_rates) {
    if (static_cast<int>(m.given_rates.size()) == 2 * C) {
      m.cat_rates.assign(m.given_rates.begin(), m.given_rates.begin() + C);
      m.cat_rate_derivs.assign(m.given_rates.begin() + C, m.given_rates.end());
    }
    return;
  }
  if (!m.weibull) return;
  double mean_rate = 0, mean_derivative = 0;
  std::vector<double> unscaled(C);
  for (int i = 0; i < C; i++) {
    const double quantile = (2.0 * i + 1.0) / (2.0 * C);
    m.cat_rates[i] = std::pow(-std::log(1.0 - quantile), 1.0 / m.shape);
    mean_rate += m.cat_rates[i];
    unscaled[i] = -m.cat_rates[i] * std::log(-std::log(1.0 - quantile)) / (m.shape * m.shape);
    mean_derivative += unscaled[i];
  }
  mean_rate /= C;
  mean_derivative /= C;
  for (int i = 0; i < C; i++) {
    m.cat_rate_derivs[i] =
        (unscaled[i] * mean_rate - m.cat_rates[i] * mean_derivative) / (mean_rate * mean_rate);
    m.cat_rates[i] /= mean_rate;
  }
}

// phylo_model.cpp:26-31: row = [GTR rates(6), frequencies(4)] [Weibull shape] [clock rate].
void SetParameters(Model& m, const double* row) {
  if (m.gtr) {
    std::copy(row, row + 6, m.rates6);
    std::copy(row + 6, row + 10, m.freqs);
  }
  if (m.weibull) m.shape = row[m.SubstitutionParamCount()];
  if (m.explicit_rates)
    m.given_rates.assign(row + m.SubstitutionParamCount(), row + m.SubstitutionParamCount() + 2 * m.categories);
  UpdateSubstitution(m);
  UpdateSite(m);
}

// stick_breaking_transform.cpp:20-43
std::vector<double> StickForward(const std::vector<double>& y) {
  const size_t K = y.size() + 1;
  std::vector<double> x(K);
  double stick = 1.0;
  for (size_t k = 0; k < K - 1; k++) {
    const double z = 1.0 / (1 + std::exp(-(y[k] - std::log(double(K - k - 1)))));
    x[k] = stick * z;
    stick -= x[k];
  }
  x[K - 1] = stick;
  return x;
}
std::vector<double> StickInverse(const std::vector<double>& x) {
  const size_t K = x.size();
  std::vector<double> y(K - 1);
  double sum = 0;
  for (size_t k = 0; k < K - 1; k++) {
    const double z = x[k] / (1.0 - sum);
    y[k] = std::log(z / (1.0 - z)) + std::log(double(K - k - 1));
    sum += x[k];
  }
  return y;
}

// ----------------------------------------------------------------- trees ----

struct Topology {
  int n = 0, N = 0, root = 0;  // bifurcating after Detrifurcate: N = 2n-1
  std::vector<int> child0, child1;
  std::vector<double> lengths;  // [N]

  // node.cpp:137-141, 209-211: children[0] subtree, children[1] subtree, node.
  void Postorder(const std::function<void(int, int, int)>& f) const {
    std::vector<std::pair<int, int>> stack{{root, 0}};
    while (!stack.empty()) {
      auto& [id, stage] = stack.back();
      if (id < n) {
        stack.pop_back();
      } else if (stage == 0) {
        stage = 1;
        stack.push_back({child0[id], 0});
      } else if (stage == 1) {
        stage = 2;
        stack.push_back({child1[id], 0});
      } else {
        const int node = id;
        stack.pop_back();
        f(node, child0[node], child1[node]);
      }
    }
  }
  // node.cpp:226-261 TriplePreorderBifurcating: f(node, sister, parent) with
  // child0 first, then child0's subtree, then child1, then child1's subtree.
  void TriplePreorder(const std::function<void(int, int, int)>& f) const {
    // (iterative, to survive ladder trees)
    if (root < n) return;
    std::vector<std::pair<int, bool>> stack{{root, false}};
    while (!stack.empty()) {
      auto [id, visited] = stack.back();
      stack.pop_back();
      if (visited) {
        f(child1[id], child0[id], id);
        if (child1[id] >= n) stack.push_back({child1[id], false});
      } else {
        f(child0[id], child1[id], id);
        stack.push_back({id, true});
        if (child0[id] >= n) stack.push_back({child0[id], false});
      }
    }
  }
  // node.cpp: Preorder over internal nodes: node, child0 subtree, child1 subtree.
  void Preorder(const std::function<void(int, int, int)>& f) const {
    std::vector<int> stack{root};
    while (!stack.empty()) {
      const int id = stack.back();
      stack.pop_back();
      if (id < n) continue;
      f(id, child0[id], child1[id]);
      stack.push_back(child1[id]);
      stack.push_back(child0[id]);
    }
  }
};

// Node::OfParentIdVector + children sorted by max leaf id (node.cpp:36-44),
// then UnrootedTree::Detrifurcate (unrooted_tree.cpp:27-37) when the root is a
// trifurcation.
Topology BuildTopology(const int32_t* parent_ids, int node_count_in, int n,
                       const double* branch_lengths) {
  Topology t;
  t.n = n;
  t.N = 2 * n - 1;
  t.root = t.N - 1;
  const bool trifurcating = node_count_in == 2 * n - 2;
  if (!trifurcating && node_count_in != 2 * n - 1) Failwith("bad node count");
  const int input_root = node_count_in - 1;
  std::vector<std::vector<int>> children(t.N);
  for (int id = 0; id < input_root; id++) {
    const int parent = parent_ids[id];
    if (parent <= id || parent > input_root) Failwith("bad parent id vector");
    children[parent].push_back(id);
  }
  std::vector<int> max_leaf(t.N, -1);
  for (int id = 0; id < n; id++) max_leaf[id] = id;
  for (int id = n; id <= input_root; id++) {
    if (children[id].size() != ((trifurcating && id == input_root) ? 3u : 2u))
      Failwith("unexpected child count");
    std::sort(children[id].begin(), children[id].end(),
              [&](int a, int b) { return max_leaf[a] < max_leaf[b]; });
    max_leaf[id] = max_leaf[children[id].back()];
  }
  t.child0.assign(t.N, -1);
  t.child1.assign(t.N, -1);
  t.lengths.assign(branch_lengths, branch_lengths + node_count_in);
  for (int id = n; id <= input_root; id++) {
    if (trifurcating && id == input_root) {
      t.child0[id] = children[id][1];
      t.child1[id] = children[id][2];
      t.lengths[id] = 0.;
      t.child0[id + 1] = children[id][0];
      t.child1[id + 1] = id;
      t.lengths.push_back(0.);
    } else {
      t.child0[id] = children[id][0];
      t.child1[id] = children[id][1];
    }
  }
  return t;
}

// ------------------------------------------------------------- FatBeagle ----

// One BEAGLE instance + one model: fat_beagle.cpp:13-29, 207-300.
class FatBeagle {
 public:
  FatBeagle(const Model& model, int n, int P, const uint8_t* tips, const double* weights,
            bool use_tip_states)
      : model_(model), n_(n), P_(P) {
    int partials = 3 * n - 2;
    if (!use_tip_states) partials += n;
    BeagleInstanceDetails info;
    instance_ = beagleCreateInstance(n, partials, use_tip_states ? n : 0, 4, P, 1,
                                     2 * (2 * n - 1), model.categories, partials + 1, nullptr, 0,
                                     BEAGLE_FLAG_VECTOR_SSE, BEAGLE_FLAG_SCALING_MANUAL, &info);
    if (instance_ < 0) Failwith("beagleCreateInstance failed");
    std::vector<int> states(P);
    std::vector<double> tip_partials(static_cast<size_t>(P) * 4);
    for (int taxon = 0; taxon < n; taxon++) {
      const uint8_t* row = tips + static_cast<size_t>(taxon) * P;
      if (use_tip_states) {
        for (int k = 0; k < P; k++) states[k] = row[k];
        beagleSetTipStates(instance_, taxon, states.data());
      } else {
        // SitePattern::GetPartials (site_pattern.cpp:117-131)
        for (int k = 0; k < P; k++)
          for (int s = 0; s < 4; s++) tip_partials[k * 4 + s] = (row[k] >= 4 || row[k] == s) ? 1. : 0.;
        beagleSetTipPartials(instance_, taxon, tip_partials.data());
      }
    }
    beagleSetPatternWeights(instance_, weights);
    UpdateModelInBeagle();
  }
  ~FatBeagle() { beagleFinalizeInstance(instance_); }
  FatBeagle(const FatBeagle&) = delete;

  Model& model() { return model_; }
  void SetRescaling(bool rescaling) { rescaling_ = rescaling; }

  void SetParametersRow(const double* row) {  // fat_beagle.cpp:43-46
    SetParameters(model_, row);
    UpdateModelInBeagle();
  }
  void UpdateSubstitutionInBeagle() {  // fat_beagle.cpp:281-294
    beagleSetStateFrequencies(instance_, 0, model_.freqs);
    beagleSetEigenDecomposition(instance_, 0, model_.evec, model_.ivec, model_.eval);
  }
  void UpdateModelInBeagle() {  // fat_beagle.cpp:273-300
    beagleSetCategoryWeights(instance_, 0, model_.cat_weights.data());
    beagleSetCategoryRates(instance_, model_.cat_rates.data());
    UpdateSubstitutionInBeagle();
  }

  // fat_beagle.cpp:50-70
  double LogLikelihoodInternals(const Topology& tree, const std::vector<double>& lengths) {
    beagleResetScaleFactors(instance_, 0);
    std::vector<BeagleOperation> operations;
    tree.Postorder([&](int node, int c0, int c1) { operations.push_back(LowerOp(node, c0, c1)); });
    UpdateTransitionMatrices(tree, lengths);
    const int cumulative = rescaling_ ? 0 : BEAGLE_OP_NONE;
    beagleUpdatePartials(instance_, operations.data(), static_cast<int>(operations.size()), cumulative);
    return RootLogLikelihood(tree, cumulative);
  }

  // fat_beagle.cpp:119-175.  `scalers` multiplies Q per category (dQ).
  std::pair<double, std::vector<double>> BranchGradientInternals(
      const Topology& tree, const std::vector<double>& lengths, const std::vector<double>& scalers) {
    beagleResetScaleFactors(instance_, 0);
    UpdateTransitionMatrices(tree, lengths);
    // fat_beagle.cpp:316-325: root pre-order partial := frequencies
    std::vector<double> root_pre(static_cast<size_t>(P_) * model_.categories * 4);
    for (size_t x = 0; x < root_pre.size(); x++) root_pre[x] = model_.freqs[x % 4];
    beagleSetPartials(instance_, tree.root + tree.N, root_pre.data());
    // fat_beagle.cpp:107-117, 127-131: dQ_c = scaler_c * Q at matrix index N-1
    std::vector<double> dQ(static_cast<size_t>(model_.categories) * 16);
    for (int c = 0; c < model_.categories; c++)
      for (int x = 0; x < 16; x++) dQ[c * 16 + x] = model_.q[x] * scalers[c];
    const int derivative_matrix = tree.N - 1;
    beagleSetDifferentialMatrix(instance_, derivative_matrix, dQ.data());
    const int cumulative = rescaling_ ? 0 : BEAGLE_OP_NONE;
    std::vector<BeagleOperation> operations;
    tree.Postorder([&](int node, int c0, int c1) { operations.push_back(LowerOp(node, c0, c1)); });
    beagleUpdatePartials(instance_, operations.data(), static_cast<int>(operations.size()), cumulative);
    operations.clear();
    tree.TriplePreorder([&](int node, int sister, int parent) {
      // fat_beagle.cpp:344-362
      const int scale_write = rescaling_ ? node + 1 + (n_ - 1) : BEAGLE_OP_NONE;
      operations.push_back({node + tree.N, scale_write, BEAGLE_OP_NONE, parent + tree.N, node, sister, sister});
    });
    beagleUpdatePrePartials(instance_, operations.data(), static_cast<int>(operations.size()),
                            BEAGLE_OP_NONE);
    std::vector<double> gradient(tree.N, 0.);
    std::vector<int> post(tree.N - 1), pre(tree.N - 1), dmat(tree.N - 1, derivative_matrix);
    std::iota(post.begin(), post.end(), 0);
    std::iota(pre.begin(), pre.end(), tree.N);
    const int category_weights = 0;
    beagleCalculateEdgeDerivatives(instance_, post.data(), pre.data(), dmat.data(), &category_weights,
                                   tree.N - 1, nullptr, gradient.data(), nullptr);
    return {RootLogLikelihood(tree, cumulative), gradient};
  }

 private:
  BeagleOperation LowerOp(int node, int c0, int c1) const {  // fat_beagle.cpp:327-342
    const int scale_write = rescaling_ ? node - n_ + 1 : BEAGLE_OP_NONE;
    return {node, scale_write, BEAGLE_OP_NONE, c0, c0, c1, c1};
  }
  void UpdateTransitionMatrices(const Topology& tree, const std::vector<double>& lengths) {
    std::vector<int> indices(tree.N - 1);  // fat_beagle.cpp:304-314
    std::iota(indices.begin(), indices.end(), 0);
    beagleUpdateTransitionMatrices(instance_, 0, indices.data(), nullptr, nullptr, lengths.data(),
                                   tree.N - 1);
  }
  double RootLogLikelihood(const Topology& tree, int cumulative) {  // fat_beagle.cpp:65-69
    double log_like = 0.;
    const int zero = 0;
    beagleCalculateRootLogLikelihoods(instance_, &tree.root, &zero, &zero, &cumulative, 1, &log_like);
    return log_like;
  }

  Model model_;
  int n_, P_;
  int instance_ = -1;
  bool rescaling_ = false;
};

// ------------------------------------------------- rooted post-processing ----

struct RootedFields {
  const double* rates;
  const double* heights;
  const double* bounds;
  const double* ratios;
  int rate_count;
};

// fat_beagle.cpp:82-94
double LogDetJacobian(const Topology& tree, const RootedFields& r) {
  double total = 0;
  tree.TriplePreorder([&](int node, int, int parent) {
    if (node >= tree.n) total += std::log(r.heights[parent] - r.bounds[node]);
  });
  return total;
}

// rooted_gradient_transforms.cpp:17-170
std::vector<double> RatioGradient(const Topology& tree, const RootedFields& r,
                                  const std::vector<double>& branch_gradient) {
  const int n = tree.n, root = tree.root;
  std::vector<double> height_gradient(n - 1, 0);
  tree.Preorder([&](int node, int c0, int c1) {
    if (node != root) height_gradient[node - n] = -branch_gradient[node] * r.rates[node];
    height_gradient[node - n] += branch_gradient[c0] * r.rates[c0];
    height_gradient[node - n] += branch_gradient[c1] * r.rates[c1];
  });
  auto node_partial = [&](int node) { return (r.heights[node] - r.bounds[node]) / r.ratios[node - n]; };
  auto chain = [&](const std::vector<double>& gh) {
    std::vector<double> out(n - 1, 0.);
    auto epoch = [&](int node, int child) -> double {
      if (child < n) return 0.0;
      if (r.bounds[node] == r.bounds[child])
        return out[child - n] * r.ratios[child - n] / r.ratios[node - n];
      return out[child - n] * r.ratios[child - n] / (r.heights[node] - r.bounds[child]) *
             node_partial(node);
    };
    tree.Postorder([&](int node, int c0, int c1) {
      if (node != root) {
        out[node - n] += node_partial(node) * gh[node - n];
        out[node - n] += epoch(node, c0);
        out[node - n] += epoch(node, c1);
      }
    });
    return out;
  };
  auto root_height = [&](const std::vector<double>& g) {
    std::vector<double> multiplier(n - 1, 0.);
    multiplier[root - n] = 1.0;
    tree.Preorder([&](int node, int c0, int c1) {
      if (c0 >= n) multiplier[c0 - n] = r.ratios[c0 - n] * multiplier[node - n];
      if (c1 >= n) multiplier[c1 - n] = r.ratios[c1 - n] * multiplier[node - n];
    });
    double sum = 0;
    for (size_t i = 0; i < g.size(); i++) sum += g[i] * multiplier[i];
    return sum;
  };
  std::vector<double> result = chain(height_gradient);
  result[root - n] = root_height(height_gradient);
  std::vector<double> log_time(n - 1, 0);
  for (int i = 0; i < n - 2; i++) log_time[i] = 1.0 / (r.heights[n + i] - r.bounds[n + i]);
  std::vector<double> jac = chain(log_time);
  jac[root - n] = root_height(log_time);
  for (int i = 0; i < n - 2; i++) result[i] += jac[i] - 1.0 / r.ratios[i];
  result[root - n] += jac[root - n];
  return result;
}

// ----------------------------------------------------------- the C entry ----

struct Request {
  const char* substitution;
  const char* site;
  int n;
  int64_t P;
  const uint8_t* tips;
  const double* weights;
  int T, node_count;
  const int32_t* parent_ids;
  const double* branch_lengths;
  const double* params;
  int rescaling, use_tip_states, rooted;
  const double *rates, *heights, *bounds, *ratios;
  int rate_count, reference_quirks, threads;
};

Model ParseModel(const char* substitution, const char* site) {
  Model m;
  const std::string sub(substitution), st(site);
  if (sub == "GTR") m.gtr = true;
  else if (sub != "JC69") Failwith("Substitution model not known: " + sub);
  if (st.rfind("weibull", 0) == 0) {
    m.weibull = true;
    m.categories = 4;
    const auto plus = st.find('+');
    if (plus != std::string::npos) m.categories = std::stoi(st.substr(plus + 1));
  } else if (st.rfind("rates+", 0) == 0) {
    m.explicit_rates = true;
    m.categories = std::stoi(st.substr(6));
  } else if (st != "constant") {
    Failwith("Site model not known: " + st);
  }
  UpdateSubstitution(m);
  UpdateSite(m);
  return m;
}

// Per tree work, parallelised over trees like FatBeagleParallelize
// (fat_beagle.hpp:119-149): one FatBeagle per thread.
template <typename F>
void ForEachTree(const Request& rq, F&& per_tree) {
  const int threads = std::max(1, std::min(rq.threads, rq.T));
  const Model base = ParseModel(rq.substitution, rq.site);
  std::vector<std::string> errors(threads);
  auto worker = [&](int id) {
    try {
      FatBeagle beagle(base, rq.n, static_cast<int>(rq.P), rq.tips, rq.weights, rq.use_tip_states != 0);
      beagle.SetRescaling(rq.rescaling != 0);
      for (int t = id; t < rq.T; t += threads) {
        beagle.SetParametersRow(rq.params + static_cast<size_t>(t) * base.ParamCount());
        Topology tree = BuildTopology(rq.parent_ids + static_cast<size_t>(t) * (rq.node_count - 1),
                                      rq.node_count, rq.n,
                                      rq.branch_lengths + static_cast<size_t>(t) * rq.node_count);
        per_tree(beagle, tree, t);
      }
    } catch (const std::exception& e) {
      errors[id] = e.what();
    }
  };
  std::vector<std::thread> pool;
  for (int id = 1; id < threads; id++) pool.emplace_back(worker, id);
  worker(0);
  for (auto& th : pool) th.join();
  for (const auto& e : errors)
    if (!e.empty()) Failwith(e);
}

}  // namespace

extern "C" {

const char* sbno_last_error(void) { return g_error.c_str(); }

// `param_stride` doubles per row of `params` (>= the model's own parameter
// count: a trailing clock-rate column is ignored, as in the reference where
// clock models only hold parameters, fat_beagle.cpp:297).
//
// rooted = 0: FatBeagle::LogLikelihood(UnrootedTree) / UnrootedLogLikelihood(RootedTree)
//             (fat_beagle.cpp:72-80);
// rooted = 1: FatBeagle::LogLikelihood(RootedTree) (fat_beagle.cpp:96-104):
//             lengths *= rates, + LogDeterminantJacobian when heights are given.
int sbno_log_likelihoods(const char* substitution, const char* site, int n, int64_t P,
                         const uint8_t* tips, const double* weights, int T, int node_count,
                         const int32_t* parent_ids, const double* branch_lengths,
                         const double* params, int param_stride, int rescaling, int use_tip_states,
                         int rooted, const double* rates, const double* heights,
                         const double* bounds, int threads, double* out) {
  try {
    const Model shape = ParseModel(substitution, site);
    std::vector<double> packed(static_cast<size_t>(T) * std::max(shape.ParamCount(), 1));
    for (int t = 0; t < T; t++)
      for (int k = 0; k < shape.ParamCount(); k++)
        packed[static_cast<size_t>(t) * shape.ParamCount() + k] = params[static_cast<size_t>(t) * param_stride + k];
    Request rq{substitution, site, n, P, tips, weights, T, node_count, parent_ids, branch_lengths,
               packed.data(), rescaling, use_tip_states, rooted, rates, heights, bounds, nullptr, 1, 0,
               threads};
    ForEachTree(rq, [&](FatBeagle& beagle, const Topology& tree, int t) {
      std::vector<double> lengths = tree.lengths;
      double jacobian = 0;
      if (rooted) {
        const double* r = rates + static_cast<size_t>(t) * (tree.N - 1);
        for (int i = 0; i < tree.N - 1; i++) lengths[i] *= r[i];
        if (heights && bounds) {
          RootedFields fields{r, heights + static_cast<size_t>(t) * tree.N,
                              bounds + static_cast<size_t>(t) * tree.N, nullptr, 1};
          jacobian = LogDetJacobian(tree, fields);
        }
      }
      out[t] = beagle.LogLikelihoodInternals(tree, lengths) + jacobian;
    });
    return 0;
  } catch (const std::exception& e) {
    g_error = e.what();
    return -1;
  }
}

// FatBeagle::Gradient(UnrootedTree) (fat_beagle.cpp:467-503) when rooted = 0,
// FatBeagle::Gradient(RootedTree) (fat_beagle.cpp:505-545) when rooted = 1.
// Outputs (any may be NULL): log_likelihood [T]; branch [T][2n-1] (unrooted:
// fixed node zeroed; rooted: the raw branch gradient); substitution [T][8];
// site [T]; ratios [T][n-1]; clock [T][rate_count].
int sbno_gradients(const char* substitution, const char* site, int n, int64_t P,
                   const uint8_t* tips, const double* weights, int T, int node_count,
                   const int32_t* parent_ids, const double* branch_lengths, const double* params,
                   int param_stride, int rescaling, int use_tip_states, int rooted,
                   const double* rates, const double* heights, const double* bounds,
                   const double* ratios, int rate_count, int reference_quirks, int threads,
                   double* out_log_likelihood, double* out_branch, double* out_substitution,
                   double* out_site, double* out_ratios, double* out_clock) {
  try {
    const Model shape = ParseModel(substitution, site);
    const int K = shape.ParamCount();
    std::vector<double> packed(static_cast<size_t>(T) * std::max(K, 1));
    for (int t = 0; t < T; t++)
      for (int k = 0; k < K; k++)
        packed[static_cast<size_t>(t) * K + k] = params[static_cast<size_t>(t) * param_stride + k];
    Request rq{substitution, site, n, P, tips, weights, T, node_count, parent_ids, branch_lengths,
               packed.data(), rescaling, use_tip_states, rooted, rates, heights, bounds, ratios,
               rate_count, reference_quirks, threads};
    ForEachTree(rq, [&](FatBeagle& beagle, Topology& tree, int t) {
      const int N = tree.N;
      std::vector<double> lengths = tree.lengths;
      const double* row = packed.data() + static_cast<size_t>(t) * K;
      RootedFields fields{};
      if (rooted) {
        fields = RootedFields{rates + static_cast<size_t>(t) * (N - 1),
                              heights ? heights + static_cast<size_t>(t) * N : nullptr,
                              bounds ? bounds + static_cast<size_t>(t) * N : nullptr,
                              ratios ? ratios + static_cast<size_t>(t) * (n - 1) : nullptr, rate_count};
        for (int i = 0; i < N - 1; i++) lengths[i] *= fields.rates[i];
      } else {
        // Tree::SlideRootPosition (tree.cpp:72-78)
        const int fixed = tree.child1[tree.root], root_child = tree.child0[tree.root];
        lengths[root_child] += lengths[fixed];
        lengths[fixed] = 0.0;
      }
      Model& model = beagle.model();
      auto [log_likelihood, branch_gradient] =
          beagle.BranchGradientInternals(tree, lengths, model.cat_rates);

      // fat_beagle.cpp:400-465: central differences in stick-breaking space;
      // f = full LogLikelihood of the ORIGINAL tree (Detrifurcate only / rates + Jacobian).
      if (model.gtr) {
        std::vector<double> f_lengths = tree.lengths;
        double jacobian = 0;
        if (rooted) {
          f_lengths = lengths;
          if (fields.heights && fields.bounds) jacobian = LogDetJacobian(tree, fields);
        }
        const double delta = 1.e-6;
        std::vector<double> param_vector(row, row + 10);
        auto finite_difference = [&](int start, int length) {
          std::vector<double> work = param_vector;  // passed by value in the reference
          std::vector<double> x(work.begin() + start, work.begin() + start + length);
          std::vector<double> y = StickInverse(x);
          std::vector<double> gradient(y.size());
          auto evaluate = [&]() {
            std::vector<double> xx = StickForward(y);
            std::copy(xx.begin(), xx.end(), work.begin() + start);
            std::copy(work.begin(), work.begin() + 6, model.rates6);
            std::copy(work.begin() + 6, work.begin() + 10, model.freqs);
            UpdateSubstitution(model);
            beagle.UpdateSubstitutionInBeagle();
            return beagle.LogLikelihoodInternals(tree, f_lengths) + jacobian;
          };
          for (size_t idx = 0; idx < y.size(); idx++) {
            const double original = y[idx];
            y[idx] += delta;
            const double plus = evaluate();
            y[idx] = original - delta;
            const double minus = evaluate();
            gradient[idx] = (plus - minus) / (2. * delta);
            y[idx] = original;
            // The reference calls SetParameters(param_vector) with the
            // minus-delta vector here (fat_beagle.cpp:433-434).
          }
          return gradient;
        };
        const std::vector<double> freq_gradient = finite_difference(6, 4);
        const std::vector<double> rate_gradient = finite_difference(0, 6);
        if (out_substitution) {
          double* dst = out_substitution + static_cast<size_t>(t) * 8;
          std::copy(rate_gradient.begin(), rate_gradient.end(), dst);
          std::copy(freq_gradient.begin(), freq_gradient.end(), dst + 5);
        }
        if (!reference_quirks) {
          std::copy(row, row + 6, model.rates6);
          std::copy(row + 6, row + 10, model.freqs);
          UpdateSubstitution(model);
          beagle.UpdateSubstitutionInBeagle();
        }
      }
      if (model.categories > 1) {  // fat_beagle.cpp:488-496, 532-540
        auto [ignored, unscaled] = beagle.BranchGradientInternals(tree, lengths, model.cat_rate_derivs);
        (void)ignored;
        double site_gradient = 0;
        for (int node = 0; node < N - 1; node++) site_gradient += unscaled[node] * lengths[node];
        if (out_site) out_site[t] = site_gradient;
      }
      if (out_log_likelihood) out_log_likelihood[t] = log_likelihood;
      if (rooted) {
        if (out_ratios && fields.heights && fields.bounds && fields.ratios) {
          const std::vector<double> g = RatioGradient(tree, fields, branch_gradient);
          std::copy(g.begin(), g.end(), out_ratios + static_cast<size_t>(t) * (n - 1));
        }
        if (out_clock) {  // fat_beagle.cpp:367-387
          const double* time_lengths = branch_lengths + static_cast<size_t>(t) * node_count;
          std::vector<double> rate_gradient(N - 1);
          for (int i = 0; i < N - 1; i++) rate_gradient[i] = branch_gradient[i] * time_lengths[i];
          double* dst = out_clock + static_cast<size_t>(t) * rate_count;
          if (rate_count == 1) dst[0] = std::accumulate(rate_gradient.begin(), rate_gradient.end(), 0.0);
          else std::copy(rate_gradient.begin(), rate_gradient.end(), dst);
        }
        if (out_branch) std::copy(branch_gradient.begin(), branch_gradient.end(), out_branch + static_cast<size_t>(t) * N);
      } else if (out_branch) {
        branch_gradient[tree.child1[tree.root]] = 0.;  // fat_beagle.cpp:498-500
        std::copy(branch_gradient.begin(), branch_gradient.end(), out_branch + static_cast<size_t>(t) * N);
      }
    });
    return 0;
  } catch (const std::exception& e) {
    g_error = e.what();
    return -1;
  }
}

}  // extern "C"
