#include "tree_program.hpp"

#include <algorithm>
#include <string>

#include "common.hpp"

namespace sbnb {

TreeProgram BuildTreeProgram(const int32_t* parent_ids, int node_count_in, int taxon_count) {
  const int n = taxon_count;
  Require(n >= 2, "A tree needs at least 2 taxa.");
  const bool bifurcating_input = (node_count_in == 2 * n - 1);
  const bool trifurcating_input = (node_count_in == 2 * n - 2) && n >= 3;
  Require(bifurcating_input || trifurcating_input,
          "node_count must be 2n-1 (bifurcating) or 2n-2 (trifurcation at the root) for n = " +
              std::to_string(n) + " taxa, got " + std::to_string(node_count_in));

  TreeProgram program;
  program.taxon_count = n;
  program.node_count = 2 * n - 1;
  program.root = 2 * n - 2;
  program.was_trifurcating = trifurcating_input;
  const int N = program.node_count;

  // Children lists from the parent vector.
  const int input_root = node_count_in - 1;
  std::vector<std::vector<int>> children(N);
  for (int id = 0; id < input_root; id++) {
    const int parent = parent_ids[id];
    Require(parent > id && parent <= input_root && parent >= n,
            "Malformed parent id vector: node " + std::to_string(id) + " has parent " +
                std::to_string(parent) + " (internal ids must follow their children).");
    children[parent].push_back(id);
  }
  // Max leaf id per node, to order children as Node's constructor does
  // (src/node.cpp:36-44).  Ids are post-ordered, so one ascending pass works.
  std::vector<int> max_leaf(N, -1);
  for (int id = 0; id < n; id++) max_leaf[id] = id;
  for (int id = n; id <= input_root; id++) {
    const size_t expected = (id == input_root && trifurcating_input) ? 3 : 2;
    Require(children[id].size() == expected,
            "Node " + std::to_string(id) + " has " + std::to_string(children[id].size()) +
                " children; expected " + std::to_string(expected) + ".");
    std::sort(children[id].begin(), children[id].end(),
              [&max_leaf](int x, int y) { return max_leaf[x] < max_leaf[y]; });
    max_leaf[id] = max_leaf[children[id].back()];
  }

  program.child0.assign(N, -1);
  program.child1.assign(N, -1);
  for (int id = n; id <= input_root; id++) {
    if (id == input_root && trifurcating_input) {
      // Detrifurcate: (c0, c1, c2) -> (c0, (c1, c2)); the inner node keeps the
      // old root id, the new root is id + 1.
      program.child0[id] = children[id][1];
      program.child1[id] = children[id][2];
      program.child0[id + 1] = children[id][0];
      program.child1[id + 1] = id;
    } else {
      program.child0[id] = children[id][0];
      program.child1[id] = children[id][1];
    }
  }

  auto is_leaf = [n](int id) { return id < n; };

  // ---- post-order program -------------------------------------------------
  // need[v] = stack depth used while evaluating v's subtree (cur excluded).
  // With two internal children the first result waits on the stack while the
  // second subtree is evaluated, so the child needing more goes first:
  // need = max(first, 1 + second).
  std::vector<int> need(N, 0);
  for (int id = n; id < N; id++) {
    const int c0 = program.child0[id], c1 = program.child1[id];
    if (!is_leaf(c0) && !is_leaf(c1)) {
      const int hi = std::max(need[c0], need[c1]), lo = std::min(need[c0], need[c1]);
      need[id] = std::max(hi, 1 + lo);
    } else {
      need[id] = std::max(need[c0], need[c1]);
    }
  }
  {
    int depth = 0, high_water = 0;
    int pending_push = -1;  // slot the next emitted op must push cur to
    std::vector<int> source(N, kFromLeaf);
    // Explicit DFS stack: (node, stage) where stage counts visited children.
    std::vector<std::pair<int, int>> stack;
    stack.push_back({program.root, 0});
    while (!stack.empty()) {
      auto [id, stage] = stack.back();
      const int c0 = program.child0[id], c1 = program.child1[id];
      const bool both = !is_leaf(c0) && !is_leaf(c1);
      const bool c0_first = is_leaf(c1) || (!is_leaf(c0) && need[c0] >= need[c1]);
      const int first = c0_first ? c0 : c1, second = c0_first ? c1 : c0;
      if (stage == 0) {
        stack.back().second = 1;
        if (!is_leaf(first)) stack.push_back({first, 0});
      } else if (stage == 1) {
        stack.back().second = 2;
        if (both) {
          // first's result is in cur; the first op of second's subtree (a
          // cherry, which does not read cur) pushes it.
          pending_push = depth;
          source[first] = depth;
          depth++;
          high_water = std::max(high_water, depth);
        }
        if (!is_leaf(second)) stack.push_back({second, 0});
      } else {
        stack.pop_back();
        PostOp op{};
        op.node = id;
        op.a = c0;
        op.b = c1;
        op.push_slot = pending_push;
        pending_push = -1;
        op.a_src = is_leaf(c0) ? kFromLeaf : source[c0];
        op.b_src = is_leaf(c1) ? kFromLeaf : source[c1];
        if (both) depth--;
        op.flags = (is_leaf(c0) ? kALeaf : 0) | (is_leaf(c1) ? kBLeaf : 0) |
                   (id == program.root ? kRoot : 0);
        source[id] = kFromCur;
        program.post.push_back(op);
      }
    }
    Require(depth == 0 && pending_push < 0, "internal error: unbalanced post-order walk");
    program.post_slots = high_water;
  }

  // ---- pre-order program --------------------------------------------------
  // pre_need[v] = stack depth used below v.  With two internal children one
  // pre-order partial stays in cur (its subtree is walked first) and the other
  // waits on the stack, so the child needing less goes first:
  // need = max(1 + first, second).
  std::vector<int> pre_need(N, 0);
  for (int id = n; id < N; id++) {
    const int c0 = program.child0[id], c1 = program.child1[id];
    if (!is_leaf(c0) && !is_leaf(c1)) {
      const int lo = std::min(pre_need[c0], pre_need[c1]);
      const int hi = std::max(pre_need[c0], pre_need[c1]);
      pre_need[id] = std::max(1 + lo, hi);
    } else {
      pre_need[id] = std::max(pre_need[c0], pre_need[c1]);
    }
  }
  {
    int high_water = 0;
    // (node, slot its pre-order partial is popped from or -1, stack depth on entry)
    struct Visit {
      int node, pop_slot, depth;
    };
    std::vector<Visit> stack;
    stack.push_back({program.root, -1, 0});
    while (!stack.empty()) {
      const Visit visit = stack.back();
      stack.pop_back();
      const int id = visit.node;
      const int c0 = program.child0[id], c1 = program.child1[id];
      const bool i0 = !is_leaf(c0), i1 = !is_leaf(c1);
      PreOp op{};
      op.node = id;
      op.a = c0;
      op.b = c1;
      op.pop_slot = visit.pop_slot;
      op.a_dst = kFromLeaf;
      op.b_dst = kFromLeaf;
      op.flags = (i0 ? 0 : kALeaf) | (i1 ? 0 : kBLeaf) | (id == program.root ? kRoot : 0);
      if (i0 && i1) {
        const bool c0_first = pre_need[c0] <= pre_need[c1];
        const int first = c0_first ? c0 : c1, second = c0_first ? c1 : c0;
        const int slot = visit.depth;
        high_water = std::max(high_water, slot + 1);
        (c0_first ? op.a_dst : op.b_dst) = kFromCur;
        (c0_first ? op.b_dst : op.a_dst) = slot;
        // LIFO: push the second so the first subtree is walked first.
        stack.push_back({second, slot, visit.depth});
        stack.push_back({first, -1, visit.depth + 1});
      } else if (i0 || i1) {
        (i0 ? op.a_dst : op.b_dst) = kFromCur;
        stack.push_back({i0 ? c0 : c1, -1, visit.depth});
      }
      program.pre.push_back(op);
    }
    program.pre_slots = high_water;
  }
  return program;
}

}  // namespace sbnb
