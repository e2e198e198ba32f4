"""Host-side topology utilities: synthetic random trees in the reference's id
convention, and Newick export for the reference arm of the benchmark.

Id convention (reference src/node.cpp:341-357 `Node::Polish`): leaves are
0..n-1, internal nodes are numbered in post-order with the children of every
node sorted by their maximum leaf id (src/node.cpp:36-44), the root is last.
A topology is its `Node::ParentIdVector` (src/node.cpp:413-424).
"""
import numpy as np


def _canonical(children, root, taxon_count):
    """Relabels an arbitrary rooted tree (dict node -> child list, leaves are
    0..n-1) into the reference convention; returns (parent_ids, old->new map)."""
    max_leaf, order, stack = {}, [], [(root, False)]
    while stack:  # iterative post-order (ladder trees are deep)
        node, done = stack.pop()
        kids = children.get(node, [])
        if not kids:
            max_leaf[node] = node
        elif done:
            kids.sort(key=lambda k: max_leaf[k])
            max_leaf[node] = max_leaf[kids[-1]]
        else:
            stack.append((node, True))
            stack.extend((k, False) for k in kids)
    new_id, next_id, stack = {}, taxon_count, [(root, False)]
    while stack:
        node, done = stack.pop()
        kids = children.get(node, [])
        if not kids:
            new_id[node] = node
        elif done:
            new_id[node] = next_id
            next_id += 1
        else:
            stack.append((node, True))
            stack.extend((k, False) for k in reversed(kids))
    parent_ids = np.full(next_id - 1, -1, dtype=np.int32)
    for node, kids in children.items():
        for k in kids:
            parent_ids[new_id[k]] = new_id[node]
    return parent_ids, new_id


def random_unrooted_topology(taxon_count, rng):
    """Uniform random stepwise addition; trifurcating root; 2n-3 parent ids."""
    assert taxon_count >= 3
    root = taxon_count
    children = {root: [0, 1, 2]}
    parent = {0: root, 1: root, 2: root}
    next_label = taxon_count + 1
    edges = [0, 1, 2]  # an edge is named by its child node
    for leaf in range(3, taxon_count):
        target = edges[rng.integers(len(edges))]
        up = parent[target]
        inner = next_label
        next_label += 1
        children[up][children[up].index(target)] = inner
        children[inner] = [target, leaf]
        parent[inner], parent[target], parent[leaf] = up, inner, inner
        edges.extend([inner, leaf])
    return _canonical(children, root, taxon_count)[0]


def random_rooted_topology(taxon_count, rng):
    """Random bifurcating rooted topology; 2n-2 parent ids."""
    assert taxon_count >= 2
    root = taxon_count
    children = {root: [0, 1]}
    parent = {0: root, 1: root}
    next_label = taxon_count + 1
    edges = [0, 1]
    for leaf in range(2, taxon_count):
        target = edges[rng.integers(len(edges))]
        up = parent[target]
        inner = next_label
        next_label += 1
        children[up][children[up].index(target)] = inner
        children[inner] = [target, leaf]
        parent[inner], parent[target], parent[leaf] = up, inner, inner
        edges.extend([inner, leaf])
    return _canonical(children, root, taxon_count)[0]


def ladder_topology(taxon_count, unrooted=True):
    """Caterpillar tree: the deepest possible traversal (n-2 dependent levels)."""
    root = 10 * taxon_count
    if unrooted:
        children, spine = {root: [0, 1, None]}, root
        slot, first = 2, 2
    else:
        children, spine = {root: [0, None]}, root
        slot, first = 1, 1
    label = taxon_count
    for leaf in range(first, taxon_count - 1):
        children[spine][slot] = label
        children[label] = [leaf, None]
        spine, slot = label, 1
        label += 1
    children[spine][slot] = taxon_count - 1
    return _canonical(children, root, taxon_count)[0]


def random_tree_batch(taxon_count, tree_count, seed, mean_branch_length=0.1, rooted=False):
    """(parent_ids [T][nodes-1], branch_lengths [T][nodes]) ~ Exp(mean), clipped >= 1e-6."""
    rng = np.random.default_rng(seed)
    make = random_rooted_topology if rooted else random_unrooted_topology
    parent_ids = np.stack([make(taxon_count, rng) for _ in range(tree_count)])
    node_count = parent_ids.shape[1] + 1
    lengths = np.maximum(rng.exponential(mean_branch_length, size=(tree_count, node_count)), 1e-6)
    lengths[:, -1] = 0.0  # the root has no branch
    return parent_ids, lengths


def random_alignment(taxon_count, pattern_count, seed, gap_fraction=0.01):
    """iid uniform {A,C,G,T} with a fraction of gap states; every column is kept
    as its own pattern with weight 1 (BASELINE config 4/5 shape)."""
    rng = np.random.default_rng(seed)
    states = rng.integers(0, 4, size=(taxon_count, pattern_count), dtype=np.uint8)
    states[rng.random((taxon_count, pattern_count)) < gap_fraction] = 4
    return states, np.ones(pattern_count)


def newick(parent_ids, branch_lengths, taxon_names=None):
    """Newick string of one tree (children in id order, lengths by node id)."""
    node_count = len(parent_ids) + 1
    kids = [[] for _ in range(node_count)]
    for child, parent in enumerate(parent_ids):
        kids[parent].append(child)
    text = {}
    for node in range(node_count):  # children precede parents
        if kids[node]:
            label = "(" + ",".join(text.pop(k) for k in kids[node]) + ")"
        else:
            label = taxon_names[node] if taxon_names else f"t{node}"
        if node != node_count - 1:
            label += f":{float(branch_lengths[node])!r}"
        text[node] = label
    return text[node_count - 1] + ";"
