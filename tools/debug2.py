import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'tests')
import numpy as np
import libsbn_b200 as sbn
from libsbn_b200 import _capi, trees
from test_full_size_gpu import alignment, raw, GTR_ROW
taxa, patterns, T = 100, 100000, 3
states = alignment(taxa, patterns, 20261017); weights = np.ones(patterns)
parent_ids, lengths = trees.random_tree_batch(taxa, T, seed=4)
params = np.tile(GTR_ROW, (T, 1))
spec = sbn.PhyloModelSpecification("GTR", "weibull+4", "none")
batch = sbn.TreeBatch(parent_ids, lengths)
e1 = sbn.Engine(spec, states, weights)
r = raw(e1, batch, params, rescaling=True)
p = [raw(e1, batch, params, rescaling=False) for _ in range(6)]
print(os.environ.get("SBNB_LIBRARY", "default"), ["%.1e" % (np.abs(x[1] - r[1]).max() / np.abs(r[1]).max()) for x in p])
