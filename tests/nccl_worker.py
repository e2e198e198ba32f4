"""torchrun worker of tests/test_sharding_gpu.py::test_nccl_two_ranks (also usable by
hand: `python -m torch.distributed.run --nproc-per-node N tests/nccl_worker.py`).
Checks both sharding axes over NCCL against the CPU oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import libsbn_b200 as sbn  # noqa: E402
from libsbn_b200 import sharding, trees  # noqa: E402
from oracle import phylo  # noqa: E402


def main():
    local_rank = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    taxa, patterns, tree_count = 12, 5003, 7
    states, weights = trees.random_alignment(taxa, patterns, seed=21, gap_fraction=0.02)
    parent_ids, lengths = trees.random_tree_batch(taxa, tree_count, seed=22)
    row = np.array([0.05, 0.1, 0.15, 0.20, 0.25, 0.25, 0.1, 0.2, 0.3, 0.4, 0.5])
    params = np.tile(row, (tree_count, 1))
    spec = sbn.PhyloModelSpecification("GTR", "weibull+4", "none")
    batch = sbn.TreeBatch(parent_ids, lengths)
    want = phylo.gradients("GTR", "weibull+4", states, weights, parent_ids, lengths, params, rescaling=True)
    for cls in (sharding.PatternShardedEngine, sharding.TreeShardedEngine):
        engine = cls(spec, states, weights, local_rank)
        got = engine.gradients(batch, params, rescaling=True, substitution_gradient=False)
        logl = np.array([g.log_likelihood for g in got])
        grad = np.array([g.gradient["branch_lengths"] for g in got])
        site = np.array([g.gradient["site_model"][0] for g in got])
        assert np.max(np.abs(logl - want["log_likelihood"]) / np.abs(want["log_likelihood"])) < 1e-10, cls
        assert np.max(np.abs(grad - want["branch"])) < 1e-8 * np.max(np.abs(want["branch"])), cls
        assert np.max(np.abs(site - want["site_model"])) < 1e-8 * np.max(np.abs(want["site_model"])), cls
        # the full call: the substitution block analytically, its sums reduced with the rest
        full = engine.gradients(batch, params, rescaling=True)
        sub = np.array([g.gradient["substitution_model"] for g in full])
        noise = np.abs(want["log_likelihood"]).max() * 1e-12 / 1e-6  # of the oracle's finite differences
        assert np.max(np.abs(sub - want["substitution_model"])) < 2 * noise, cls
        assert np.array_equal(np.array([g.gradient["branch_lengths"] for g in full]), grad), cls
        only_logl = engine.log_likelihoods(batch, params, rescaling=True)
        assert np.max(np.abs(only_logl - want["log_likelihood"]) / np.abs(want["log_likelihood"])) < 1e-10, cls
        # every rank holds the same bits (one collective, same reduction order everywhere)
        mine = torch.from_numpy(np.concatenate([logl, grad.ravel()])).cuda()
        others = [torch.empty_like(mine) for _ in range(dist.get_world_size())]
        dist.all_gather(others, mine)
        assert all(torch.equal(o, mine) for o in others), cls
    dist.barrier()
    if dist.get_rank() == 0:
        print("NCCL-SHARDING-OK", dist.get_world_size(), "ranks")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
