// Host-side schedule generation: turns one topology (a Node::ParentIdVector)
// into the two flat "programs" the tree-walk kernel interprets.
//
// This replaces the reference's per-call op construction
// (Node::BinaryIdPostorder / TripleIdPreorderBifurcating with std::function
// callbacks, src/node.cpp:190-261; AddLowerPartialOperation /
// AddUpperPartialOperation, src/fat_beagle.cpp:327-362).  The op SET is the
// same -- one post-order op per internal node, one pre-order visit per
// internal node producing both children's pre-order partials -- but the ORDER
// is chosen so that a depth-first walk needs only O(log n) live partials
// (Strahler ordering), which is what lets a pattern tile keep its working set
// in shared memory instead of round-tripping every partial through HBM.
#ifndef SBNB_TREE_PROGRAM_HPP_
#define SBNB_TREE_PROGRAM_HPP_

#include <cstdint>
#include <vector>

namespace sbnb {

// One post-order op: dest = (P_a L_a) o (P_b L_b) for internal node `node`
// with children a, b (reference child order: sorted by max leaf id).
// 32 bytes, read by the kernel as two int4.
struct PostOp {
  int32_t node;      // destination node id (n..2n-2); its internal index is node-n
  int32_t a;         // child 0 node id (= its matrix index; a taxon id when a leaf)
  int32_t b;         // child 1 node id
  int32_t dst_slot;  // stack slot the result is written to
  int32_t a_slot;    // stack slot holding child 0's partial (-1 when a leaf)
  int32_t b_slot;    // stack slot holding child 1's partial (-1 when a leaf)
  int32_t flags;     // kALeaf | kBLeaf | kRoot
  int32_t pad;
};

// One pre-order visit of internal node `node` with pre-order partial in
// `pre_slot` (root: the stationary distribution, no slot): computes the
// children's pre-order partials, their edge derivatives, and pushes the
// pre-order partials of internal children.
struct PreOp {
  int32_t node;
  int32_t a;
  int32_t b;
  int32_t pre_slot;    // slot of this node's pre-order partial (-1 at the root)
  int32_t a_dst_slot;  // where child 0's pre-order partial goes (-1 when a leaf)
  int32_t b_dst_slot;
  int32_t flags;       // kALeaf | kBLeaf | kRoot
  int32_t pad;
};

enum : int32_t { kALeaf = 1, kBLeaf = 2, kRoot = 4 };

struct TreeProgram {
  int taxon_count = 0;
  int node_count = 0;  // 2n-1 after detrifurcation
  int root = 0;
  bool was_trifurcating = false;
  std::vector<int32_t> child0, child1;  // per node id, -1 for leaves
  std::vector<PostOp> post;             // n-1 ops
  std::vector<PreOp> pre;               // n-1 ops
  int post_slots = 0;                   // stack depth the post-order walk needs
  int pre_slots = 0;
};

// parent_ids: node_count_in-1 entries (the root has no entry).  Accepts a
// bifurcating tree (2n-1 nodes) or a tree with a trifurcation at the root
// (2n-2 nodes), which is detrifurcated as UnrootedTree::Detrifurcate does
// (src/unrooted_tree.cpp:27-37): children (c0,c1,c2) -> (c0,(c1,c2)), the new
// inner node takes the old root's id, the new root gets id+1.
TreeProgram BuildTreeProgram(const int32_t* parent_ids, int node_count_in, int taxon_count);

}  // namespace sbnb

#endif  // SBNB_TREE_PROGRAM_HPP_
