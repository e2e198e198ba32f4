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
// on chip instead of round-tripping every partial through HBM.
#ifndef SBNB_TREE_PROGRAM_HPP_
#define SBNB_TREE_PROGRAM_HPP_

#include <cstdint>
#include <vector>

namespace sbnb {

// The walk keeps ONE partial -- the result of the previous op -- in registers
// ("cur") and everything else that is live on a stack.  A depth-first walk
// consumes most partials immediately (a node's op directly follows the op of its
// last-visited internal child), so only nodes with two internal children touch
// the stack: one push and one pop each.

// Operand / destination codes of the op records below.
enum : int32_t { kFromLeaf = -1, kFromCur = -2 };

// One post-order op: cur = (P_a L_a) o (P_b L_b) for internal node `node` with
// children a, b (reference child order: sorted by max leaf id).
// 32 bytes; the debug API copies these records out as 8 x int32.
struct PostOp {
  int32_t node;       // destination node id (n..2n-2); its internal index is node-n
  int32_t a;          // child 0 node id (= its matrix index; a taxon id when a leaf)
  int32_t b;          // child 1 node id
  int32_t push_slot;  // >= 0: push cur to this stack slot BEFORE executing (else -1)
  int32_t a_src;      // kFromLeaf, kFromCur or the stack slot to pop child 0's partial from
  int32_t b_src;      // same for child 1
  int32_t flags;      // kALeaf | kBLeaf | kRoot
  int32_t pad;
};

// One pre-order visit of internal node `node`, whose pre-order partial is in
// cur (root: the stationary distribution) or is popped from `pop_slot` first:
// computes both children's edge derivatives and the pre-order partials of the
// internal children; one of them stays in cur, the other is pushed.
struct PreOp {
  int32_t node;
  int32_t a;
  int32_t b;
  int32_t pop_slot;  // >= 0: pop this node's pre-order partial from the slot first (else -1)
  int32_t a_dst;     // kFromLeaf (nothing to do), kFromCur (stays in cur) or the slot pushed to
  int32_t b_dst;
  int32_t flags;     // kALeaf | kBLeaf | kRoot
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
  int post_slots = 0;                   // stack depth the post-order walk needs (may be 0)
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
