"""BASELINE.json configs[4]: synthetic 1000 taxa x 1M site patterns, HKY + 4 rate
categories, T trees, SITE PATTERNS sharded across the ranks (one process per GPU)
with one NCCL sum-all-reduce of the per-tree log-likelihoods and gradient sums
(libsbn_b200.sharding.PatternShardedEngine).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
      --master-port P tools/config5_bench.py [--taxa 1000 --patterns 1000000 --trees 8]

Every rank builds the same alignment (seeded) and keeps its contiguous range of
patterns on its GPU.  A step is one PatternShardedEngine.gradients() call on
the whole tree batch: staging, the tree walk over the local patterns, the
all-reduce on the device result arrays, and the O(n) host finishing.  Timed as
bench.py does (barrier + synchronize on both sides, max over ranks); rank 0
prints one JSON line."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import libsbn_b200 as sbn  # noqa: E402
from libsbn_b200 import sharding, trees  # noqa: E402


def alignment(taxa, patterns, seed):
    """iid uniform {A,C,G,T} with 1 % gap states, generated in byte arithmetic
    (trees.random_alignment draws float64 per cell: 8 GB at this size)."""
    rng = np.random.default_rng(seed)
    states = rng.integers(0, 4, size=(taxa, patterns), dtype=np.uint8)
    gaps = rng.integers(0, 100, size=(taxa, patterns), dtype=np.uint8) == 0
    states[gaps] = 4
    return states, np.ones(patterns)


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--taxa", type=int, default=1000)
    parser.add_argument("--patterns", type=int, default=1000000)
    parser.add_argument("--trees", type=int, default=8)
    parser.add_argument("--steps", type=int, default=3)
    parser.add_argument("--warmup", type=int, default=1)
    args = parser.parse_args()
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local_rank)
    sys.stdout.flush()
    json_fd = os.dup(1)  # NCCL prints its banner to descriptor 1: point that at stderr for the run
    os.dup2(2, 1)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    rank, world = dist.get_rank(), dist.get_world_size()

    states, weights = alignment(args.taxa, args.patterns, seed=5)
    parent_ids, lengths = trees.random_tree_batch(args.taxa, args.trees, seed=5)
    # row layout (blocks sorted by name, as the reference does): 4 frequencies, kappa; Weibull shape
    params = np.tile(np.array([0.1, 0.2, 0.3, 0.4, 2.0, 0.5]), (args.trees, 1))
    spec = sbn.PhyloModelSpecification("HKY", "weibull+4", "none")
    engine = sharding.PatternShardedEngine(spec, states, weights, local_rank)
    del states
    batch = sbn.TreeBatch(parent_ids, lengths)

    def step():
        return engine.gradients(batch, params, rescaling=True, substitution_gradient=False)

    for _ in range(args.warmup):
        result = step()
    engine.engine.walk_timing(reset=True)
    dist.barrier()
    torch.cuda.synchronize()
    start = time.perf_counter()
    for _ in range(args.steps):
        result = step()
    torch.cuda.synchronize()
    elapsed = torch.tensor([time.perf_counter() - start], device="cuda")
    dist.barrier()
    dist.all_reduce(elapsed, op=dist.ReduceOp.MAX)
    walk_ms, walk_samples = engine.engine.walk_timing(reset=True)
    kernel_ms = torch.tensor([walk_ms / max(walk_samples, 1)], device="cuda")
    dist.all_reduce(kernel_ms, op=dist.ReduceOp.MAX)
    # every rank must hold the same reduced results
    mine = torch.tensor([g.log_likelihood for g in result], device="cuda", dtype=torch.float64)
    others = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(others, mine)
    same = all(torch.equal(o, mine) for o in others)
    if rank == 0:
        seconds = elapsed.item() / args.steps
        n, local = args.taxa, engine.end - engine.begin
        unit_bytes = 32 * 4 * args.patterns  # one partial buffer over ALL patterns
        algorithmic = (10 * n - 14) * unit_bytes * args.trees
        peaks = os.path.join(ROOT, "MEASURED_PEAKS.json")
        peak = json.load(open(peaks))["hbm_gbs"] if os.path.exists(peaks) else 6650.0
        line = json.dumps({
            "metric": "tree logL+branch-gradient evals/sec", "value": args.trees / seconds, "unit": "evals/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": seconds * 1e3,
            "scaling": "strong (site patterns sharded, one NCCL all-reduce of [T] logL + [T x (2n-1)] sums)",
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"synthetic {n} taxa x {args.patterns} site patterns, HKY+weibull4, "
                                   f"{args.trees} trees, logL + branch gradients, rescaling on "
                                   "(BASELINE.json configs[4])",
                       "patterns_per_gpu": int(local)},
            "timing": "host clock around the public call (staging + walk + all-reduce + finishing), max over ranks",
            "kernel_ms_max_over_ranks": kernel_ms.item(),
            "algorithmic_GBps_all_gpus": algorithmic / (kernel_ms.item() * 1e-3) / 1e9,
            "frac_of_hbm_peak_per_gpu": algorithmic / world / (kernel_ms.item() * 1e-3) / 1e9 / peak,
            "ranks_agree_bitwise": bool(same),
            "mean_log_likelihood": float(np.mean([g.log_likelihood for g in result])),
        })
        os.write(json_fd, (line + "\n").encode())
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
