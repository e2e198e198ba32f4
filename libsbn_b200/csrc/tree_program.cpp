#include "tree_program.hpp"

#include <algorithm>
#include <string>

#include "common.hpp"

namespace sbnb {

namespace {

// LIFO pool of stack slots; the high-water mark is the stack depth.
class SlotPool {
 public:
  int Acquire() {
    if (!free_.empty()) {
      const int slot = free_.back();
      free_.pop_back();
      return slot;
    }
    return high_water_++;
  }
  void Release(int slot) {
    if (slot >= 0) free_.push_back(slot);
  }
  int HighWater() const { return high_water_; }

 private:
  std::vector<int> free_;
  int high_water_ = 0;
};

}  // namespace

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
  // need[v] = stack slots required to evaluate v's subtree, its own result
  // included; the child needing more goes first, and the destination reuses a
  // child's slot, so need = max(first, [first internal] + second, 1).
  std::vector<int> need(N, 0);
  for (int id = n; id < N; id++) {
    const int x = std::max(need[program.child0[id]], need[program.child1[id]]);
    const int y = std::min(need[program.child0[id]], need[program.child1[id]]);
    need[id] = std::max({x, (x > 0 ? 1 : 0) + y, 1});
  }
  {
    SlotPool pool;
    std::vector<int> slot_of(N, -1);
    // Explicit DFS stack: (node, stage) where stage counts visited children.
    std::vector<std::pair<int, int>> stack;
    stack.push_back({program.root, 0});
    while (!stack.empty()) {
      auto [id, stage] = stack.back();
      const int c0 = program.child0[id], c1 = program.child1[id];
      const bool c0_first = need[c0] >= need[c1];
      const int first = c0_first ? c0 : c1, second = c0_first ? c1 : c0;
      if (stage == 0) {
        stack.back().second = 1;
        if (!is_leaf(first)) stack.push_back({first, 0});
      } else if (stage == 1) {
        stack.back().second = 2;
        if (!is_leaf(second)) stack.push_back({second, 0});
      } else {
        stack.pop_back();
        PostOp op{};
        op.node = id;
        op.a = c0;
        op.b = c1;
        op.a_slot = slot_of[c0];
        op.b_slot = slot_of[c1];
        op.flags = (is_leaf(c0) ? kALeaf : 0) | (is_leaf(c1) ? kBLeaf : 0) |
                   (id == program.root ? kRoot : 0);
        if (op.a_slot >= 0) {
          op.dst_slot = op.a_slot;
          pool.Release(op.b_slot);
        } else if (op.b_slot >= 0) {
          op.dst_slot = op.b_slot;
        } else {
          op.dst_slot = pool.Acquire();
        }
        slot_of[id] = op.dst_slot;
        program.post.push_back(op);
      }
    }
    program.post_slots = pool.HighWater();
  }

  // ---- pre-order program --------------------------------------------------
  // pre_need[v] = slots needed below v given v's own pre-order partial holds
  // one; a single internal child overwrites the parent's slot, two internal
  // children cost one extra slot while the first subtree is walked.
  std::vector<int> pre_need(N, 0);
  for (int id = n; id < N; id++) {
    const int c0 = program.child0[id], c1 = program.child1[id];
    const bool i0 = !is_leaf(c0), i1 = !is_leaf(c1);
    if (i0 && i1) {
      const int lo = std::min(pre_need[c0], pre_need[c1]);
      const int hi = std::max(pre_need[c0], pre_need[c1]);
      pre_need[id] = std::max(1 + lo, hi);
    } else if (i0 || i1) {
      pre_need[id] = pre_need[i0 ? c0 : c1];
    } else {
      pre_need[id] = 1;
    }
  }
  {
    SlotPool pool;
    std::vector<int> slot_of(N, -1);
    std::vector<int> stack;
    stack.push_back(program.root);
    while (!stack.empty()) {
      const int id = stack.back();
      stack.pop_back();
      const int c0 = program.child0[id], c1 = program.child1[id];
      const bool i0 = !is_leaf(c0), i1 = !is_leaf(c1);
      PreOp op{};
      op.node = id;
      op.a = c0;
      op.b = c1;
      op.pre_slot = slot_of[id];
      op.a_dst_slot = -1;
      op.b_dst_slot = -1;
      op.flags = (i0 ? 0 : kALeaf) | (i1 ? 0 : kBLeaf) | (id == program.root ? kRoot : 0);
      // The node's own slot is dead once both children are computed; the
      // first-walked internal child inherits it.
      int inherited = op.pre_slot;
      if (i0 && i1) {
        const bool c0_first = pre_need[c0] <= pre_need[c1];
        const int first = c0_first ? c0 : c1, second = c0_first ? c1 : c0;
        slot_of[first] = inherited >= 0 ? inherited : pool.Acquire();
        slot_of[second] = pool.Acquire();
        // LIFO: push the second so the first subtree is walked first.
        stack.push_back(second);
        stack.push_back(first);
      } else if (i0 || i1) {
        const int only = i0 ? c0 : c1;
        slot_of[only] = inherited >= 0 ? inherited : pool.Acquire();
        stack.push_back(only);
      } else {
        pool.Release(inherited);
      }
      if (i0) op.a_dst_slot = slot_of[c0];
      if (i1) op.b_dst_slot = slot_of[c1];
      program.pre.push_back(op);
    }
    program.pre_slots = std::max(pool.HighWater(), 1);
  }
  if (program.post_slots < 1) program.post_slots = 1;
  return program;
}

}  // namespace sbnb
