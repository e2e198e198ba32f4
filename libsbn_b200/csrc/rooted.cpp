#include "rooted.hpp"

#include <cmath>

#include "common.hpp"

namespace sbnb {

namespace {

// Internal node ids ascend in post-order (children before parents), so an
// ascending loop is a post-order pass and a descending loop a pre-order pass.

std::vector<int> ParentOf(const TreeProgram& tree) {
  std::vector<int> parent(tree.node_count, -1);
  for (int id = tree.taxon_count; id < tree.node_count; id++) {
    parent[tree.child0[id]] = id;
    parent[tree.child1[id]] = id;
  }
  return parent;
}

// (height - bound) / ratio for an internal node.
double NodePartial(int node, int n, const RootedView& v) {
  return (v.node_heights[node] - v.node_bounds[node]) / v.height_ratios[node - n];
}

// Contribution of `child` to d t / d ratio of `node` (epoch structure).
double EpochAddition(int node, int child, int n, const RootedView& v,
                     const std::vector<double>& ratio_gradient) {
  if (child < n) return 0.0;
  if (v.node_bounds[node] == v.node_bounds[child]) {
    return ratio_gradient[child - n] * v.height_ratios[child - n] / v.height_ratios[node - n];
  }
  return ratio_gradient[child - n] * v.height_ratios[child - n] /
         (v.node_heights[node] - v.node_bounds[child]) * NodePartial(node, n, v);
}

// Chain rule from node heights to height ratios for all non-root internal nodes.
std::vector<double> RatioChain(const TreeProgram& tree, const RootedView& v,
                               const std::vector<double>& height_gradient) {
  const int n = tree.taxon_count;
  std::vector<double> out(n - 1, 0.0);
  for (int node = n; node < tree.node_count; node++) {
    if (node == tree.root) continue;
    out[node - n] += NodePartial(node, n, v) * height_gradient[node - n];
    out[node - n] += EpochAddition(node, tree.child0[node], n, v, out);
    out[node - n] += EpochAddition(node, tree.child1[node], n, v, out);
  }
  return out;
}

// Derivative with respect to the root height: every height moves with the
// product of the ratios on its path to the root.
double RootHeightChain(const TreeProgram& tree, const RootedView& v,
                       const std::vector<double>& gradient) {
  const int n = tree.taxon_count;
  std::vector<double> multiplier(n - 1, 0.0);
  multiplier[tree.root - n] = 1.0;
  for (int node = tree.node_count - 1; node >= n; node--) {
    const int children[2] = {tree.child0[node], tree.child1[node]};
    for (int child : children)
      if (child >= n) multiplier[child - n] = v.height_ratios[child - n] * multiplier[node - n];
  }
  double sum = 0.0;
  for (int i = 0; i < n - 1; i++) sum += gradient[i] * multiplier[i];
  return sum;
}

}  // namespace

double LogDetJacobianHeightRatios(const TreeProgram& tree, const RootedView& v) {
  const std::vector<int> parent = ParentOf(tree);
  double total = 0.0;
  for (int node = tree.taxon_count; node < tree.node_count; node++) {
    if (node == tree.root) continue;
    total += std::log(v.node_heights[parent[node]] - v.node_bounds[node]);
  }
  return total;
}

std::vector<double> RatioGradientOfBranchGradient(const TreeProgram& tree, const RootedView& v,
                                                  const double* branch_gradient) {
  const int n = tree.taxon_count;
  // Height gradient: a node's height lengthens its children's branches and
  // shortens its own (rooted_gradient_transforms.cpp:17-37).
  std::vector<double> height_gradient(n - 1, 0.0);
  for (int node = n; node < tree.node_count; node++) {
    double g = 0.0;
    if (node != tree.root) g = -branch_gradient[node] * v.rates[node];
    g += branch_gradient[tree.child0[node]] * v.rates[tree.child0[node]];
    g += branch_gradient[tree.child1[node]] * v.rates[tree.child1[node]];
    height_gradient[node - n] = g;
  }
  std::vector<double> result = RatioChain(tree, v, height_gradient);
  result[tree.root - n] = RootHeightChain(tree, v, height_gradient);

  // Gradient of the log-det-Jacobian of the ratio transform
  // (rooted_gradient_transforms.cpp:132-161).
  std::vector<double> log_time(n - 1, 0.0);
  for (int i = 0; i < n - 2; i++)
    log_time[i] = 1.0 / (v.node_heights[n + i] - v.node_bounds[n + i]);
  std::vector<double> jacobian_gradient = RatioChain(tree, v, log_time);
  jacobian_gradient[tree.root - n] = RootHeightChain(tree, v, log_time);
  for (int i = 0; i < n - 2; i++) result[i] += jacobian_gradient[i] - 1.0 / v.height_ratios[i];
  result[tree.root - n] += jacobian_gradient[tree.root - n];
  return result;
}

std::vector<double> ClockGradient(const TreeProgram& tree, const RootedView& v,
                                  const double* branch_gradient) {
  const int edges = tree.node_count - 1;
  std::vector<double> rate_gradient(edges);
  for (int i = 0; i < edges; i++) rate_gradient[i] = branch_gradient[i] * v.branch_lengths[i];
  if (v.rate_count == 1) {
    double sum = 0.0;
    for (double g : rate_gradient) sum += g;
    return {sum};
  }
  if (v.rate_count == edges) return rate_gradient;
  Fail(SBNB_ERR_INVALID_ARGUMENT,
       "The number of rates should be equal to 1 (i.e. strict clock) or equal to the number of "
       "branches.");
}

double DiscreteSiteModelGradient(int node_count, const double* branch_lengths,
                                 const double* unscaled_category_gradient) {
  double total = 0.0;
  for (int node = 0; node < node_count - 1; node++)
    total += unscaled_category_gradient[node] * branch_lengths[node];
  return total;
}

}  // namespace sbnb
