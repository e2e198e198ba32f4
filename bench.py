"""bench.py -- BASELINE.json's headline metric on BASELINE.json's config.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[3]): synthetic 100 taxa x 100,000 site patterns,
GTR + 4 rate categories (the reference's "weibull+4"), a batch of 1024 random
unrooted topologies PER GPU, rescaling on, sharded BY TREE across the ranks (no
data-path collective; weak scaling: every rank walks its own 1024-tree batch,
rank r's batch drawn with seed 4 + r, so N = 1 is exactly configs[3]).

A "step" is one pass of the hot path over the batch: per tree, the
log-likelihood and all 2n-2 branch-length derivatives, i.e. one
BranchGradientInternals-equivalent (reference src/fat_beagle.cpp:119-175).

  value : batch staged in HBM, only the kernels in the timed region
          (CUDA events on the engine's stream, max over ranks).
  e2e   : the public call -- host arrays in, PhyloGradient arrays out -- per
          step: host-side schedule generation, H2D from page-locked staging,
          kernels, D2H, host finishing.
  roofline: the tree-walk kernel's algorithmic bytes (SURVEY.md 8d:
          (10n-14) x 32 C P per tree) / its CUDA-event duration, against the
          measured HBM copy bandwidth (MEASURED_PEAKS.json).
  cpu_baseline / --impl reference: the UNMODIFIED reference host code
          (oracle/_ref, its own python module, thread pool = host cores) over
          the BEAGLE-equivalent CPU kernels of oracle/beagle_cpu.cpp, on a
          bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "tree logL+branch-gradient evals/sec"
UNIT = "evals/s"
GTR_ROW = [0.05, 0.1, 0.15, 0.20, 0.25, 0.25, 0.1, 0.2, 0.3, 0.4, 0.5]  # rates, freqs, Weibull shape


def workload(args, rank=0):
    from libsbn_b200 import trees
    states, weights = trees.random_alignment(args.taxa, args.patterns, seed=20261017, gap_fraction=0.01)
    parent_ids, lengths = trees.random_tree_batch(args.taxa, args.trees, seed=4 + rank, mean_branch_length=0.1)
    params = np.tile(np.array(GTR_ROW), (args.trees, 1))
    return states, weights, parent_ids, lengths, params


def config_of(args):
    return {
        "workload": f"synthetic {args.taxa} taxa x {args.patterns} site patterns, GTR+weibull4 (4 rate "
                    f"categories), {args.trees} random unrooted topologies per GPU, logL + branch gradients, "
                    "rescaling on (BASELINE.json configs[3])",
        "taxa": args.taxa, "patterns": args.patterns, "categories": 4, "trees_per_gpu": args.trees,
        "sharding": "trees across ranks (each rank its own batch of trees_per_gpu), no data-path collective",
        "l2": "no flush needed: each step streams the post-order scratch arena and per-tree matrices "
              "(>1 GB) through a 126 MB L2",
    }


# --------------------------------------------------------------------------- clocks

class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.rows, self.proc = device, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *exc):
        if self.proc:
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        sm, sm_max, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for row in self.rows:
            try:
                sm.append(float(row[1])), sm_max.append(float(row[2])), power.append(float(row[3]))
            except (ValueError, IndexError):
                continue
            for name, value in zip(names, row[5:9]):
                if value.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        busy = [c for c, p in zip(sm, power) if p > 0.5 * max(power)] or sm
        return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": max(sm_max), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power)}


# --------------------------------------------------------------------------- reference arm

# Site patterns of the bounded CPU sample: with one tree per core this is ~40 core-seconds
# per repeat for the reference's phylo_gradients() + log_likelihoods() pair (the cost of one
# BranchGradientInternals is a DIFFERENCE of the two timings, so the sample must not be tiny).
CPU_SAMPLE_PATTERNS = 10000


def reference_throughput(args, cores, sample_patterns, sample_trees, repeats=1):
    """Times the reference's own CPU path on a bounded sample of the workload.

    Prefers the unmodified reference (oracle/_ref: its pybind11 module, Engine
    thread pool of `cores` FatBeagles, stock code path) over the oracle port.
    The reference exposes BranchGradientInternals only inside phylo_gradients(),
    which for GTR + weibull runs 2 of them plus 16 full log-likelihood sweeps
    per tree (fat_beagle.cpp:467-503); both calls are timed and the cost of one
    BranchGradientInternals is (t_phylo_gradients - 16 t_log_likelihoods) / 2.
    Returns evals/s scaled to the full pattern count (cost is linear in patterns).
    """
    from libsbn_b200 import trees
    states, weights, parent_ids, lengths, params = workload(args)
    states, weights = states[:, :sample_patterns], weights[:sample_patterns]
    parent_ids, lengths, params = parent_ids[:sample_trees], lengths[:sample_trees], params[:sample_trees]
    scale = sample_patterns / args.patterns
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    detail = {"cores": cores, "sample": f"{sample_trees} trees x {sample_patterns} of {args.patterns} patterns"
                                        f" (throughput scaled by {scale:g}; cost is linear in patterns)"}
    try:
        sys.path.insert(0, ref_dir)
        import libsbn  # the reference's own python module
    except ImportError:
        libsbn = None
    if libsbn is not None:
        with tempfile.TemporaryDirectory() as tmp:
            names = [f"t{i}" for i in range(args.taxa)]
            with open(os.path.join(tmp, "a.fasta"), "w") as f:
                for name, row in zip(names, states):
                    f.write(f">{name}\n{''.join('ACGT-'[s] for s in row)}\n")
            with open(os.path.join(tmp, "t.nwk"), "w") as f:
                for ids, bl in zip(parent_ids, lengths):
                    f.write(trees.newick(ids, bl, names) + "\n")
            inst = libsbn.unrooted_instance("bench")
            inst.read_newick_file(os.path.join(tmp, "t.nwk"))
            inst.read_fasta_file(os.path.join(tmp, "a.fasta"))
            devnull = os.open(os.devnull, os.O_WRONLY)
            saved = os.dup(1)
            os.dup2(devnull, 1)  # the reference prints its BEAGLE banner per thread
            try:
                inst.prepare_for_phylo_likelihood(libsbn.PhyloModelSpecification("GTR", "weibull+4", "none"),
                                                  cores, [], True)
            finally:
                os.dup2(saved, 1)
                os.close(devnull)
            block = inst.get_phylo_model_param_block_map()
            block["GTR rates"][:] = GTR_ROW[:6]
            block["frequencies"][:] = GTR_ROW[6:10]
            block["Weibull shape"][:] = GTR_ROW[10]
            inst.set_rescaling(True)
            t_pg = t_ll = 0.0
            for _ in range(repeats):
                t0 = time.perf_counter()
                inst.phylo_gradients()
                t1 = time.perf_counter()
                inst.log_likelihoods()
                t2 = time.perf_counter()
                t_pg += t1 - t0
                t_ll += t2 - t1
        t_bgi = max((t_pg - 16.0 * t_ll) / 2.0, 1e-9)
        detail.update(kind="reference",
                      how="unmodified reference python module (oracle/_ref) over BEAGLE-equivalent CPU "
                          "kernels; one BranchGradientInternals = (t phylo_gradients - 16 t log_likelihoods)/2",
                      phylo_gradients_s=t_pg / repeats, log_likelihoods_s=t_ll / repeats,
                      full_phylo_gradients_trees_per_s=sample_trees * repeats * scale / t_pg)
        return sample_trees * repeats * scale / t_bgi, detail
    from oracle import phylo
    t0 = time.perf_counter()
    for _ in range(repeats):
        phylo.gradients("JC69", "weibull+4", states, weights, parent_ids, lengths,
                        np.full((sample_trees, 1), GTR_ROW[10]), rescaling=True, threads=cores)
    elapsed = time.perf_counter() - t0
    # JC69 + weibull: Gradient() = exactly 2 BranchGradientInternals, no finite differences.
    detail.update(kind="port", how="oracle port (oracle/phylo_oracle.cpp), JC69+weibull4 stand-in: 2 "
                                   "BranchGradientInternals per tree, same 4x4 kernels as GTR")
    return 2 * sample_trees * repeats * scale / elapsed, detail


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample_patterns = min(args.patterns, CPU_SAMPLE_PATTERNS)
    sample_trees = min(args.trees, max(cores, 8))
    for _ in range(args.warmup):
        reference_throughput(args, cores, sample_patterns, sample_trees)
    values, detail = [], {}
    t0 = time.perf_counter()
    for _ in range(args.steps):
        value, detail = reference_throughput(args, cores, sample_patterns, sample_trees)
        values.append(value)
    elapsed = time.perf_counter() - t0
    value = float(np.mean(values))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * elapsed / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": config_of(args),
        "cpu_baseline": dict(detail, value=value, unit=UNIT),
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------- our arm

def run_ours(args):
    import torch
    import torch.distributed as dist
    import libsbn_b200 as sbn
    from libsbn_b200 import _capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    # stdout carries exactly one JSON line.  Libraries write to file descriptor 1 behind
    # Python's back (NCCL prints its version banner there in this image), so for the
    # duration of the run descriptor 1 points at stderr and the line goes to the saved one.
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    states, weights, parent_ids, lengths, params = workload(args, rank)
    shard = sbn.TreeBatch(parent_ids, lengths)
    shard_params = params
    engine = sbn.Engine(sbn.PhyloModelSpecification("GTR", "weibull+4", "none"), states, weights, local_rank)
    stream = torch.cuda.ExternalStream(engine.stream, device=local_rank)
    total_trees = args.trees * world

    def timed(body, steps):
        """CUDA events on the engine's stream around `steps` calls of body()."""
        barrier()
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            start.record()
            for _ in range(steps):
                body()
            stop.record()
        stop.synchronize()
        barrier()
        timed.local_ms = start.elapsed_time(stop)
        return max_over_ranks(timed.local_ms)

    # ---- value: inputs resident in HBM, kernels only ---------------------------------
    staged = engine.stage(shard, shard_params)
    run = lambda: staged.run(_capi.MODE_BRANCH_GRADIENT, True)
    for _ in range(args.warmup):
        run()
    engine.walk_timing(reset=True)
    launches_before = engine.launch_count
    with ClockSampler(local_rank) as clocks:
        ms = timed(run, args.steps)
    ms_local = timed.local_ms
    launches = engine.launch_count - launches_before
    walk_ms, walk_samples = engine.walk_timing(reset=True)
    value = total_trees * args.steps / (ms * 1e-3)
    logl_check = staged.fetch()

    # roofline of the dominant kernel (TreeWalkLcKernel, gradient mode), this rank's launch
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_source = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_source = 6650.0, "fallback (B200_PROFILING.md)"
    bytes_per_launch = staged.algorithmic_bytes(_capi.MODE_BRANCH_GRADIENT)
    kernel_ms = walk_ms / max(walk_samples, 1)
    achieved = bytes_per_launch / (kernel_ms * 1e-3) / 1e9
    # What actually binds the fused walk: its real DRAM traffic (ncu, per tree) and
    # its fp64 work (26.1 kflop per pattern x category per tree) against the B200's
    # non-tensor fp64 rate (64 DFMA / clk / SM x 148 SMs x 1.965 GHz = 37.2 TFLOP/s).
    traffic_per_tree, profile_name = None, "r01_treewalk_v5_ncu_summary.json"
    profile = os.path.join(ROOT, "profiles", profile_name)
    if os.path.exists(profile):
        traffic_per_tree = json.load(open(profile)).get("dram_bytes_per_tree")
    flops_per_launch = 26.1e3 * args.patterns * 4 * args.trees
    fp64_peak = 148 * 64 * 2 * 1.965e9
    floors_ms = {"fp64": flops_per_launch / fp64_peak * 1e3}
    if traffic_per_tree and (args.taxa, args.patterns) == (100, 100000):
        floors_ms["hbm_real_traffic"] = traffic_per_tree * args.trees / (peak * 1e9) * 1e3
    binding = max(floors_ms, key=floors_ms.get)
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic_per_tree * args.trees if "hbm_real_traffic" in floors_ms else None,
                "kernel": "TreeWalkLcKernel<C=4,K=4,GRAD,RESCALE>", "kernel_ms": kernel_ms,
                "kernel_share_of_step": walk_ms / (ms_local if world > 1 else ms),
                "algorithmic_bytes_per_launch": bytes_per_launch, "peak_source": peak_source,
                "traffic_source": f"profiles/{profile_name} (ncu --set full, dram read+write per tree x trees)",
                "binding_bound": {"which": binding, "floor_ms": floors_ms,
                                  "frac_of_binding_bound": floors_ms[binding] / kernel_ms},
                "note": "achieved = ALGORITHMIC bytes of the reference's op-at-a-time schedule ((10n-14) x 32 C P "
                        "per tree, every partial through HBM) / kernel time; the fused walk keeps O(log n) "
                        "partials on chip and moves only `traffic` bytes, so frac > 1 is not an HBM measurement: "
                        "the honest ceiling is binding_bound (max of the real-traffic HBM floor and the fp64 floor)"}

    # ---- the same batch, log-likelihood only (post-order sweep + root reduction) -----
    run_logl = lambda: staged.run(_capi.MODE_LOG_LIKELIHOOD, True)
    for _ in range(2):
        run_logl()
    engine.walk_timing(reset=True)
    logl_ms = timed(run_logl, args.steps)
    logl_walk_ms, logl_samples = engine.walk_timing(reset=True)
    logl_bytes = staged.algorithmic_bytes(_capi.MODE_LOG_LIKELIHOOD)
    logl_only = {"value": total_trees * args.steps / (logl_ms * 1e-3), "unit": "logL evals/s",
                 "ms_per_step": logl_ms / args.steps,
                 "algorithmic_GBps": logl_bytes / (logl_walk_ms / max(logl_samples, 1) * 1e-3) / 1e9,
                 "frac_of_hbm_peak": logl_bytes / (logl_walk_ms / max(logl_samples, 1) * 1e-3) / 1e9 / peak}
    assert np.allclose(staged.fetch()[:args.trees], logl_check[:args.trees], rtol=1e-12)

    # ---- e2e: the public call with host buffers ----------------------------------------
    call = lambda: engine.gradients(shard, shard_params, rescaling=True, substitution_gradient=False)
    for _ in range(min(args.warmup, 2)):
        results = call()
    h2d_before, d2h_before = engine.transfer_bytes
    e2e_steps = max(1, min(args.steps, 5))
    e2e_ms = timed(call, e2e_steps)
    h2d_after, d2h_after = engine.transfer_bytes
    e2e_value = total_trees * e2e_steps / (e2e_ms * 1e-3)
    assert np.allclose([g.log_likelihood for g in results], logl_check[:len(results)], rtol=1e-12)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_of(args), "clocks": clocks.summary(),
        "e2e": {"value": e2e_value, "unit": UNIT, "steps": e2e_steps,
                "h2d_bytes_per_step": (h2d_after - h2d_before) // e2e_steps * world,
                "d2h_bytes_per_step": (d2h_after - d2h_before) // e2e_steps * world},
        "gpu_launches": int(launches) * world, "roofline": roofline, "log_likelihood_only": logl_only,
        "mean_log_likelihood": float(np.mean(logl_check)),
    }
    if world == 1 and rank == 0 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        cpu_value, detail = reference_throughput(args, cores, min(args.patterns, CPU_SAMPLE_PATTERNS),
                                                 min(args.trees, max(cores, 8)), repeats=2)
        line["cpu_baseline"] = dict(detail, value=cpu_value, unit=UNIT)
    if rank == 0:
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--gpus", type=int, default=1)
    parser.add_argument("--steps", type=int, default=5)
    parser.add_argument("--warmup", type=int, default=3)
    parser.add_argument("--impl", default="ours", choices=["ours", "reference"])
    parser.add_argument("--taxa", type=int, default=100)
    parser.add_argument("--patterns", type=int, default=100000)
    parser.add_argument("--trees", type=int, default=1024)
    parser.add_argument("--no-cpu-baseline", action="store_true")
    args = parser.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        run_reference(args)
    else:
        args.warmup = max(args.warmup, 3)
        run_ours(args)


if __name__ == "__main__":
    main()
