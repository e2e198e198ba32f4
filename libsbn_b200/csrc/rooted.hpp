// Host-side O(n) post-processing of the edge gradient for rooted time trees
// (SURVEY.md row a13): what FatBeagle::Gradient(RootedTree) does after
// BranchGradientInternals (reference src/fat_beagle.cpp:505-545) --
// RatioGradientOfBranchGradient (src/rooted_gradient_transforms.cpp:17-170),
// ClockGradient / DiscreteSiteModelGradient (src/fat_beagle.cpp:367-398) and
// LogDeterminantJacobian (src/fat_beagle.cpp:82-94) -- on flat arrays.
#ifndef SBNB_ROOTED_HPP_
#define SBNB_ROOTED_HPP_

#include <vector>

#include "tree_program.hpp"

namespace sbnb {

// View of one RootedTree's time-tree parameterisation (rooted_tree.hpp).
struct RootedView {
  const double* branch_lengths;  // [2n-1] unscaled (time) branch lengths
  const double* rates;           // [2n-2]
  const double* node_heights;    // [2n-1]
  const double* node_bounds;     // [2n-1]
  const double* height_ratios;   // [n-1]
  int rate_count;
};

// sum over internal non-root nodes of log(height[parent] - bound[node]).
double LogDetJacobianHeightRatios(const TreeProgram& tree, const RootedView& view);

// d logL / d (height ratios, root height), including the log-det-Jacobian
// term; n-1 entries indexed by internal node id - n.
std::vector<double> RatioGradientOfBranchGradient(const TreeProgram& tree, const RootedView& view,
                                                  const double* branch_gradient);

// d logL / d clock rate(s): one entry (strict) or one per branch.
std::vector<double> ClockGradient(const TreeProgram& tree, const RootedView& view,
                                  const double* branch_gradient);

// sum over edges of unscaled_category_gradient[e] * branch_length[e].
double DiscreteSiteModelGradient(int node_count, const double* branch_lengths,
                                 const double* unscaled_category_gradient);

}  // namespace sbnb

#endif  // SBNB_ROOTED_HPP_
