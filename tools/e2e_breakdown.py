"""Host-side breakdown of one public gradients() call on the bench workload."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import libsbn_b200 as sbn
from libsbn_b200 import _capi, trees, sharding

taxa, patterns, T = 100, 100000, 1024
states, weights = trees.random_alignment(taxa, patterns, seed=20261017, gap_fraction=0.01)
parent_ids, lengths = trees.random_tree_batch(taxa, T, seed=4, mean_branch_length=0.1)
params = np.tile(np.array([0.05, 0.1, 0.15, 0.20, 0.25, 0.25, 0.1, 0.2, 0.3, 0.4, 0.5]), (T, 1))
spec = sbn.PhyloModelSpecification("GTR", "weibull+4", "none")
engine = sbn.Engine(spec, states, weights, 0)
batch = sbn.TreeBatch(parent_ids, lengths)
for _ in range(2):
    engine.gradients(batch, params, rescaling=True, substitution_gradient=False)
def clock():
    torch.cuda.synchronize()
    return time.perf_counter()
for rep in range(3):
    t0 = clock(); staged = engine.stage(batch, params)
    t1 = clock(); staged.run(_capi.MODE_BRANCH_GRADIENT, True)
    t2 = clock(); logl, grad, rgrad = staged.fetch(gradients=True)
    t3 = clock(); out = sharding.finish_gradients(spec, taxa, batch, False, False, logl, grad, rgrad, 4)
    t4 = clock(); staged.close()
    t5 = clock(); res = engine.gradients(batch, params, rescaling=True, substitution_gradient=False)
    t6 = clock()
    print(f"stage {1e3*(t1-t0):.1f} ms  run {1e3*(t2-t1):.1f}  fetch {1e3*(t3-t2):.1f}  finish(py) {1e3*(t4-t3):.1f}  "
          f"close {1e3*(t5-t4):.1f}  | one call {1e3*(t6-t5):.1f}")
