// libsbn_b200 engine: device buffers, staging, kernel launches and the C ABI of
// include/sbn_b200.h.  Replaces the reference's Engine / FatBeagle pair
// (src/engine.cpp, src/fat_beagle.cpp) and the BEAGLE instance they drive.
//
// There is deliberately no CPU path in this file: without a CUDA device every
// entry point fails.

#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <exception>
#include <mutex>
#include <thread>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "common.hpp"
#include "device_common.cuh"
#include "kernels.cuh"
#include "walk_lc.cuh"
#include "model.hpp"
#include "rooted.hpp"
#include "tree_program.hpp"

namespace sbnb {

namespace {

thread_local std::string g_last_error;

// Category counts the tree walk is instantiated for (the lanes of a pattern hold
// one category each, so powers of two); other counts run padded
// with zero-weight categories.
int PadCategories(int c) {
  for (int supported : {1, 2, 4, 8, 16})
    if (c <= supported) return supported;
  return 0;
}

int EnvInt(const char* name, int fallback) {
  const char* value = std::getenv(name);
  return value ? std::atoi(value) : fallback;
}

}  // namespace

void SetLastError(const std::string& message) { g_last_error = message; }

}  // namespace sbnb

using namespace sbnb;

struct sbnb_engine {
  ModelSpec spec;
  int device = 0;
  int sm_count = 0;
  int taxon_count = 0;
  int64_t pattern_count = 0;
  int64_t tip_pitch = 0;
  int64_t range_begin = 0, range_end = 0;
  int categories = 1;         // C
  int padded_categories = 1;  // instantiated count >= C (extra categories have weight 0)
  std::vector<uint8_t> host_tips;    // [taxon][pattern], states clamped to 0..4
  std::vector<double> host_weights;  // [pattern]
  cudaStream_t stream = nullptr;
  // CUDA-event pairs bracketing the base tree-walk launch of the last
  // kWalkRing runs; harvested into (walk_total_ms, walk_samples).
  static constexpr int kWalkRing = 32;
  cudaEvent_t walk_begin[kWalkRing] = {}, walk_end[kWalkRing] = {};
  bool walk_pending[kWalkRing] = {};
  int64_t walk_runs = 0;
  double walk_total_ms = 0.0, walk_last_ms = 0.0;
  int64_t walk_samples = 0;
  int64_t launch_count = 0;
  int64_t h2d_bytes = 0, d2h_bytes = 0;
  PinnedArena staging;
  DeviceArray<uint8_t> tips;
  DeviceArray<double> weights;
  DeviceArray<double2> scratch;  // evolved post-order partial arena, reused by every gradient run
  DeviceArray<double2> stack;    // per-CTA partial stacks
  DeviceArray<int32_t> stack_exps;

  // A destroyed batch parks its device arrays here so the next Stage() of a
  // same-sized collection (the one-call entry points stage per call) neither
  // cudaMallocs nor cudaFrees (cudaFree synchronises the device).
  sbnb_batch* spare = nullptr;

  ~sbnb_engine();
};

struct sbnb_batch {
  int tree_count = 0;
  int taxon_count = 0;
  int node_count = 0;  // 2n-1
  bool rooted = false;
  int fd_coords = 0;  // stick-breaking coordinates perturbed (0 if no FD staged)
  int vtree_count = 0;
  int slots = 1;
  int last_mode = -1;
  int64_t patterns = 0;  // pattern range of the engine when staged
  int categories = 1;
  // host side, kept for the O(n) finishing steps
  std::vector<TreeProgram> programs;
  std::vector<double> lengths;  // [T][2n-1] after detrifurcation / rate scaling / root slide
  // device side
  DeviceArray<WalkOp> ops;  // [tree][2(n-1)]: post-order ops then pre-order ops
  DeviceArray<int32_t> vtree_program, vtree_model, vtree_lengths;
  DeviceArray<ModelTables> models;
  DeviceArray<double> d_lengths, matrices;
  DeviceArray<double> logl_partial, grad_partial, rgrad_partial;
  DeviceArray<double> logl, grad, rgrad;
  // tiling of the last run
  int chunks = 1;
};

sbnb_engine::~sbnb_engine() {
  delete spare;
  for (int i = 0; i < kWalkRing; i++) {
    if (walk_begin[i]) cudaEventDestroy(walk_begin[i]);
    if (walk_end[i]) cudaEventDestroy(walk_end[i]);
  }
  if (stream) cudaStreamDestroy(stream);
}

namespace {

// Parks a finished batch in its engine for reuse (see sbnb_engine::spare).
void Recycle(sbnb_engine* e, sbnb_batch* batch) {
  if (!batch) return;
  if (e && !e->spare) {
    batch->programs.clear();
    batch->lengths.clear();
    batch->last_mode = -1;
    e->spare = batch;
  } else {
    delete batch;
  }
}
struct BatchRecycler {
  sbnb_engine* engine;
  void operator()(sbnb_batch* batch) const { Recycle(engine, batch); }
};
using BatchPtr = std::unique_ptr<sbnb_batch, BatchRecycler>;

struct LaunchPlan {
  int tiles_total, tiles_per_chunk, chunks, grid;
  size_t smem_bytes;
};

// Plans (and launches) TreeWalkLcKernel, whose tiles are kThreads / C * K patterns.
template <int C, int K, bool GRAD, bool RESCALE>
LaunchPlan PlanAndLaunchLc(sbnb_engine* e, WalkParams p, bool launch, int chunks_override) {
  auto kernel = TreeWalkLcKernel<C, K, GRAD, RESCALE>;
  const size_t smem = LcSmemBytes(C, K, GRAD);
  SBNB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 static_cast<int>(smem)));
  int per_sm = 0;
  SBNB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, smem));
  if (per_sm < 1) Fail(SBNB_ERR_CUDA, "TreeWalkLcKernel does not fit on an SM.");
  const int resident = per_sm * e->sm_count;
  constexpr int kTilePatterns = LcTilePatterns(C, K);
  LaunchPlan plan;
  const int64_t patterns = p.pattern_end - p.pattern_begin;
  plan.tiles_total = static_cast<int>((patterns + kTilePatterns - 1) / kTilePatterns);
  if (chunks_override > 0) {
    plan.chunks = chunks_override;
    plan.tiles_per_chunk = (plan.tiles_total + plan.chunks - 1) / plan.chunks;
  } else {
    // Enough (tree, chunk) work items to give every resident CTA ~4 of them, and
    // enough chunks per tree that the CTAs resident at any moment work on few
    // distinct trees: consecutive items are the chunks of one tree, and every tree
    // in flight keeps its ~0.4 MB of transition matrices in L2 next to the arena
    // stream (1024 trees at 2 chunks each had 222 trees -- 81 MB -- in flight).
    int64_t want = (4LL * resident + p.vtree_count - 1) / std::max(p.vtree_count, 1);
    const int in_flight = std::max(1, EnvInt("SBNB_TREES_IN_FLIGHT", 16));
    want = std::max<int64_t>(want, (resident + in_flight - 1) / in_flight);
    want = std::max<int64_t>(1, std::min<int64_t>(want, plan.tiles_total));
    plan.tiles_per_chunk = static_cast<int>((plan.tiles_total + want - 1) / want);
    plan.chunks = (plan.tiles_total + plan.tiles_per_chunk - 1) / plan.tiles_per_chunk;
  }
  const int64_t items = static_cast<int64_t>(p.vtree_count) * plan.chunks;
  plan.grid = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(items, resident)));
  plan.smem_bytes = smem;
  if (launch) {
    p.tiles_total = plan.tiles_total;
    p.tiles_per_chunk = plan.tiles_per_chunk;
    p.chunks = plan.chunks;
    const size_t block = static_cast<size_t>(K) * 2 * kThreads;  // double2 per partial block
    const size_t slots = static_cast<size_t>(std::max(p.slots, 1));
    e->stack.Reserve(static_cast<size_t>(plan.grid) * slots * block);
    p.stack = e->stack.get();
    if (RESCALE) {
      e->stack_exps.Reserve(static_cast<size_t>(plan.grid) * slots * K * kThreads);
      p.stack_exps = e->stack_exps.get();
    }
    if (GRAD) {
      e->scratch.Reserve(static_cast<size_t>(plan.grid) * (p.taxon_count - 1) * block);
      p.scratch = e->scratch.get();
    }
    kernel<<<plan.grid, kThreads, smem, e->stream>>>(p);
    SBNB_CUDA(cudaGetLastError());
    e->launch_count++;
  }
  return plan;
}

template <int C, int K, bool GRAD>
LaunchPlan DispatchRescale(sbnb_engine* e, const WalkParams& p, bool rescale, bool launch, int chunks_override) {
  return rescale ? PlanAndLaunchLc<C, K, GRAD, true>(e, p, launch, chunks_override)
                 : PlanAndLaunchLc<C, K, GRAD, false>(e, p, launch, chunks_override);
}

// K_GRAD / K_LOGL patterns per thread for the gradient / logL-only walk.
template <int C, int K_GRAD, int K_LOGL>
LaunchPlan DispatchModesLc(sbnb_engine* e, const WalkParams& p, bool grad, bool rescale, bool launch,
                           int chunks_override) {
  return grad ? DispatchRescale<C, K_GRAD, true>(e, p, rescale, launch, chunks_override)
              : DispatchRescale<C, K_LOGL, false>(e, p, rescale, launch, chunks_override);
}

// Patterns per thread K: more of them amortise the per-op overhead (operand requests,
// op decoding, barrier waits, reductions) but cost registers; the pre-order half of a
// gradient walk works through them two at a time (LcPreBatch).  Tiles are
// kThreads / C * K patterns and must stay a multiple of 16.
LaunchPlan Dispatch(sbnb_engine* e, const WalkParams& p, bool grad, bool rescale, bool launch,
                    int chunks_override) {
  switch (e->padded_categories) {
    case 1:
      return DispatchModesLc<1, 2, 4>(e, p, grad, rescale, launch, chunks_override);
    case 2:
      return DispatchModesLc<2, 2, 4>(e, p, grad, rescale, launch, chunks_override);
    case 4:  // (measured: gradient walks with 4 patterns per thread at 2 CTAs/SM beat 2 at 3 by 3 %)
      return DispatchModesLc<4, 4, 4>(e, p, grad, rescale, launch, chunks_override);
    case 8:
      return DispatchModesLc<8, 2, 2>(e, p, grad, rescale, launch, chunks_override);
    case 16:
      return DispatchModesLc<16, 2, 2>(e, p, grad, rescale, launch, chunks_override);
  }
  Fail(SBNB_ERR_INVALID_ARGUMENT, "Unsupported category count.");
}

// (Re)builds the device copy of the alignment for patterns [begin, end): the
// range starts at column 0 of the device arrays (TMA bulk copies need 16-byte
// aligned rows), padded with gap states / zero weights so that the last tile
// needs no bounds checks (a tile is at most kThreads * 2 patterns).
void UploadPatternRange(sbnb_engine* e, int64_t begin, int64_t end) {
  const int64_t count = end - begin;
  e->tip_pitch = ((count + 511) / 512) * 512 + 512;
  std::vector<uint8_t> tips(static_cast<size_t>(e->taxon_count) * e->tip_pitch, 4);
  for (int taxon = 0; taxon < e->taxon_count; taxon++)
    std::copy(e->host_tips.begin() + static_cast<size_t>(taxon) * e->pattern_count + begin,
              e->host_tips.begin() + static_cast<size_t>(taxon) * e->pattern_count + end,
              tips.begin() + static_cast<size_t>(taxon) * e->tip_pitch);
  std::vector<double> weights(e->tip_pitch, 0.0);
  std::copy(e->host_weights.begin() + begin, e->host_weights.begin() + end, weights.begin());
  SBNB_CUDA(cudaStreamSynchronize(e->stream));  // nothing in flight reads the old arrays
  e->h2d_bytes += e->tips.Upload(tips.data(), tips.size(), e->stream);
  e->h2d_bytes += e->weights.Upload(weights.data(), weights.size(), e->stream);
  SBNB_CUDA(cudaStreamSynchronize(e->stream));
  e->range_begin = begin;
  e->range_end = end;
}

void CheckTrees(const sbnb_engine* e, const sbnb_tree_batch* trees, bool rooted) {
  Require(trees != nullptr, "NULL tree batch.");
  Require(trees->tree_count >= 0, "Negative tree count.");
  Require(trees->tree_count == 0 || (trees->parent_ids && trees->branch_lengths),
          "NULL parent_ids / branch_lengths.");
  const int n = e->taxon_count;
  if (rooted) {
    Require(trees->node_count == 2 * n - 1,
            "Rooted trees must be bifurcating: node_count must be 2n-1.");
    Require(trees->tree_count == 0 || trees->rates != nullptr,
            "Rooted evaluation needs per-branch rates (RootedTree::rates_).");
  } else {
    Require(trees->node_count == 2 * n - 1 || trees->node_count == 2 * n - 2,
            "node_count must be 2n-2 (unrooted) or 2n-1 (bifurcating).");
  }
}

// Perturbed parameter rows of the finite-difference substitution gradient,
// in output order: GTR -> coordinates 0..4 = rates, 5..7 = frequencies
// (fat_beagle.cpp:455-464), each as (plus, minus).
void FiniteDifferenceRows(const ModelSpec& spec, const double* row, double delta,
                          std::vector<std::vector<double>>* rows) {
  rows->clear();
  auto perturb_simplex = [&](const std::string& key) {
    const auto [start, length] = spec.Block(key);
    std::vector<double> y(length - 1), x(length);
    StickBreakingInverse(row + start, length, y.data());
    for (int idx = 0; idx < length - 1; idx++) {
      for (int sign = +1; sign >= -1; sign -= 2) {
        std::vector<double> yy = y;
        yy[idx] += sign * delta;
        StickBreaking(yy.data(), length, x.data());
        std::vector<double> perturbed(row, row + spec.param_count);
        std::copy(x.begin(), x.end(), perturbed.begin() + start);
        rows->push_back(std::move(perturbed));
      }
    }
  };
  if (spec.substitution == SubstitutionKind::kGTR) {
    perturb_simplex("GTR rates");
    perturb_simplex("frequencies");
  } else if (spec.substitution == SubstitutionKind::kHKY) {
    const int kappa = spec.Block("kappa").first;
    for (int sign = +1; sign >= -1; sign -= 2) {
      std::vector<double> perturbed(row, row + spec.param_count);
      perturbed[kappa] += sign * delta;
      rows->push_back(std::move(perturbed));
    }
    perturb_simplex("frequencies");
  }
}

constexpr double kFiniteDifferenceDelta = 1.e-6;  // fat_beagle.cpp:454

// Host-side data parallelism over the trees of a batch (what the reference's
// TaskProcessor thread pool does, task_processor.hpp:43-112): contiguous chunks,
// one std::thread each; the first exception is rethrown on the calling thread.
template <typename F>
void ParallelOverTrees(int count, F&& body) {
  const int workers = std::max(1, std::min<int>({static_cast<int>(std::thread::hardware_concurrency()), 16,
                                                 count / 32}));
  if (workers <= 1) {
    for (int t = 0; t < count; t++) body(t);
    return;
  }
  std::vector<std::thread> threads;
  std::exception_ptr failure;
  std::mutex failure_mutex;
  for (int w = 0; w < workers; w++) {
    threads.emplace_back([&, w] {
      try {
        const int begin = static_cast<int>(static_cast<int64_t>(count) * w / workers);
        const int end = static_cast<int>(static_cast<int64_t>(count) * (w + 1) / workers);
        for (int t = begin; t < end; t++) body(t);
      } catch (...) {
        std::lock_guard<std::mutex> lock(failure_mutex);
        if (!failure) failure = std::current_exception();
      }
    });
  }
  for (auto& thread : threads) thread.join();
  if (failure) std::rethrow_exception(failure);
}

// The branch lengths one evaluation of tree t uses, indexed by node id of the
// bifurcating (2n-1 node) tree.
void EffectiveBranchLengths(const TreeProgram& program, const sbnb_tree_batch* trees, int t, bool rooted,
                            int N, double* out) {
  const double* in = trees->branch_lengths + static_cast<size_t>(t) * trees->node_count;
  std::copy(in, in + trees->node_count, out);
  if (program.was_trifurcating) {
    // Detrifurcate (unrooted_tree.cpp:31-35): the node that takes the old
    // root's id and the new root both get length 0.
    out[N - 2] = 0.0;
    out[N - 1] = 0.0;
  }
  if (rooted) {
    // fat_beagle.cpp:96-101, 507-511
    const double* rates = trees->rates + static_cast<size_t>(t) * (N - 1);
    for (int i = 0; i < N - 1; i++) out[i] *= rates[i];
  }
}

BatchPtr Stage(sbnb_engine* e, const sbnb_tree_batch* trees, const double* params, bool rooted,
               bool with_fd) {
  CheckTrees(e, trees, rooted);
  const ModelSpec& spec = e->spec;
  Require(spec.param_count == 0 || params != nullptr || trees->tree_count == 0,
          "NULL phylo model parameter matrix.");
  SBNB_CUDA(cudaSetDevice(e->device));
  BatchPtr batch(e->spare ? e->spare : new sbnb_batch(), BatchRecycler{e});
  e->spare = nullptr;
  const int T = trees->tree_count, n = e->taxon_count, N = 2 * n - 1;
  batch->tree_count = T;
  batch->taxon_count = n;
  batch->node_count = N;
  batch->rooted = rooted;
  batch->patterns = e->range_end - e->range_begin;
  batch->categories = e->categories;
  const int fd_evals = with_fd ? 2 * spec.SubstitutionGradientSize() : 0;
  batch->fd_coords = fd_evals / 2;
  batch->vtree_count = T * (1 + fd_evals);
  if (T == 0) return batch;

  // Everything that goes to the device is assembled in page-locked memory.
  const size_t op_count = static_cast<size_t>(T) * 2 * (n - 1);
  const size_t max_models = static_cast<size_t>(T) * (1 + fd_evals);
  e->staging.Reset(op_count * sizeof(WalkOp) + 3 * batch->vtree_count * sizeof(int32_t) +
                   max_models * sizeof(ModelTables) + static_cast<size_t>(T) * N * sizeof(double) + 16 * 256);
  WalkOp* ops = e->staging.Take<WalkOp>(op_count);
  int32_t* vtree_program = e->staging.Take<int32_t>(batch->vtree_count);
  int32_t* vtree_model = e->staging.Take<int32_t>(batch->vtree_count);
  int32_t* vtree_lengths = e->staging.Take<int32_t>(batch->vtree_count);
  ModelTables* models = e->staging.Take<ModelTables>(max_models);
  double* lengths = e->staging.Take<double>(static_cast<size_t>(T) * N);

  // Programs + branch lengths.
  batch->programs.resize(T);
  std::vector<int> tree_slots(T, 1);
  ParallelOverTrees(T, [&](int t) {
    TreeProgram program = BuildTreeProgram(
        trees->parent_ids + static_cast<size_t>(t) * (trees->node_count - 1), trees->node_count, n);
    EffectiveBranchLengths(program, trees, t, rooted, N, lengths + static_cast<size_t>(t) * N);
    // Pack both programs into the 16-byte records the kernel streams.
    auto slot_byte = [](int32_t slot) { return slot < 0 ? 0xff : (slot & 0xff); };
    Require(program.post_slots < 255 && program.pre_slots < 255 && N < (1 << 24), "Tree too large.");
    WalkOp* tree_ops = ops + static_cast<size_t>(t) * 2 * (n - 1);
    for (int o = 0; o < n - 1; o++) {
      const PostOp& op = program.post[o];
      const int post_flags = op.flags | (op.push_slot >= 0 ? kStackBefore : 0) |
                             (op.a_src == kFromCur ? kACur : 0) | (op.b_src == kFromCur ? kBCur : 0);
      tree_ops[o] = make_int4(op.a, op.b, op.node | (post_flags << 24),
                              slot_byte(op.push_slot) | (slot_byte(op.a_src) << 8) | (slot_byte(op.b_src) << 16));
      const PreOp& pre = program.pre[o];
      const int pre_flags = pre.flags | (pre.pop_slot >= 0 ? kStackBefore : 0) |
                            (pre.a_dst == kFromCur ? kACur : 0) | (pre.b_dst == kFromCur ? kBCur : 0);
      tree_ops[n - 1 + o] =
          make_int4(pre.a, pre.b, pre.node | (pre_flags << 24),
                    slot_byte(pre.pop_slot) | (slot_byte(pre.a_dst) << 8) | (slot_byte(pre.b_dst) << 16));
    }
    tree_slots[t] = std::max(program.post_slots, program.pre_slots);
    program.post.clear();
    program.post.shrink_to_fit();
    program.pre.clear();
    program.pre.shrink_to_fit();
    batch->programs[t] = std::move(program);
  });
  batch->slots = std::max(1, *std::max_element(tree_slots.begin(), tree_slots.end()));
  batch->lengths.assign(lengths, lengths + static_cast<size_t>(T) * N);

  // Models: one table per distinct consecutive parameter row (+ its FD rows).
  size_t model_count = 0;
  const int K = spec.param_count;
  int previous_base_model = -1;
  std::vector<std::vector<double>> fd_rows;
  for (int t = 0; t < T; t++) {
    const double* row = params + static_cast<size_t>(t) * K;
    const bool same_as_previous =
        t > 0 && (K == 0 || std::memcmp(row, row - K, sizeof(double) * K) == 0);
    if (!same_as_previous) {
      previous_base_model = static_cast<int>(model_count);
      BuildModelTables(spec, row, &models[model_count++]);
      if (fd_evals) {
        FiniteDifferenceRows(spec, row, kFiniteDifferenceDelta, &fd_rows);
        for (const auto& fd_row : fd_rows) BuildModelTables(spec, fd_row.data(), &models[model_count++]);
      }
    }
    vtree_program[t] = t;
    vtree_model[t] = previous_base_model;
    vtree_lengths[t] = t;
    for (int f = 0; f < fd_evals; f++) {
      const int v = T + t * fd_evals + f;
      vtree_program[v] = t;
      vtree_model[v] = previous_base_model + 1 + f;
      vtree_lengths[v] = t;
    }
  }

  cudaStream_t s = e->stream;
  e->h2d_bytes += batch->ops.Upload(ops, op_count, s);
  e->h2d_bytes += batch->vtree_program.Upload(vtree_program, batch->vtree_count, s);
  e->h2d_bytes += batch->vtree_model.Upload(vtree_model, batch->vtree_count, s);
  e->h2d_bytes += batch->vtree_lengths.Upload(vtree_lengths, batch->vtree_count, s);
  e->h2d_bytes += batch->models.Upload(models, model_count, s);
  e->h2d_bytes += batch->d_lengths.Upload(lengths, static_cast<size_t>(T) * N, s);
  batch->matrices.Reserve(static_cast<size_t>(batch->vtree_count) * (N - 1) * e->padded_categories *
                          kEdgeDoublesPerCategory);
  batch->logl.Reserve(batch->vtree_count);
  batch->grad.Reserve(static_cast<size_t>(T) * N);
  batch->rgrad.Reserve(static_cast<size_t>(T) * N);
  // The staging arena is reused by the next call.
  SBNB_CUDA(cudaStreamSynchronize(s));
  return batch;
}

// Folds one finished event pair into the running totals (waits for it).
void HarvestWalkTiming(sbnb_engine* e, int ring) {
  if (!e->walk_pending[ring]) return;
  SBNB_CUDA(cudaEventSynchronize(e->walk_end[ring]));
  float ms = 0.f;
  SBNB_CUDA(cudaEventElapsedTime(&ms, e->walk_begin[ring], e->walk_end[ring]));
  e->walk_total_ms += ms;
  e->walk_last_ms = ms;
  e->walk_samples++;
  e->walk_pending[ring] = false;
}

WalkParams BaseParams(sbnb_engine* e, sbnb_batch* b) {
  WalkParams p{};
  p.tips = e->tips.get();
  p.tip_pitch = e->tip_pitch;
  p.weights = e->weights.get();
  p.pattern_begin = 0;  // the device arrays start at the engine's pattern range
  p.pattern_end = e->range_end - e->range_begin;
  p.taxon_count = e->taxon_count;
  p.ops = b->ops.get();
  p.vtree_program = b->vtree_program.get();
  p.vtree_model = b->vtree_model.get();
  p.models = b->models.get();
  p.matrices = b->matrices.get();
  p.slots = b->slots;
  return p;
}

void Run(sbnb_engine* e, sbnb_batch* b, int mode, bool rescaling) {
  Require(mode == SBNB_MODE_LOG_LIKELIHOOD || mode == SBNB_MODE_BRANCH_GRADIENT, "Unknown mode.");
  SBNB_CUDA(cudaSetDevice(e->device));
  b->last_mode = mode;
  if (b->tree_count == 0) return;
  const bool grad = (mode == SBNB_MODE_BRANCH_GRADIENT);
  const int T = b->tree_count, N = b->node_count, C = e->padded_categories;
  cudaStream_t s = e->stream;
  const int vtrees = grad ? b->vtree_count : T;

  // K1: transition matrices for every (virtual tree, edge, category).
  {
    const int64_t total = static_cast<int64_t>(vtrees) * (N - 1) * C;
    const int block = 128;
    const int grid = static_cast<int>((total + block - 1) / block);
    TransitionMatrixKernel<<<grid, block, 0, s>>>(b->models.get(), b->vtree_model.get(),
                                                  b->vtree_lengths.get(), b->d_lengths.get(),
                                                  b->matrices.get(), vtrees, N - 1, N, C);
    SBNB_CUDA(cudaGetLastError());
    e->launch_count++;
  }

  WalkParams p = BaseParams(e, b);
  // Base trees: logL or gradient sweep.
  p.vtree_begin = 0;
  p.vtree_count = T;
  LaunchPlan plan = Dispatch(e, p, grad, rescaling, /*launch=*/false, 0);
  const int fd_vtrees = vtrees - T;
  LaunchPlan fd_plan{};
  if (fd_vtrees > 0) {
    WalkParams q = p;
    q.vtree_begin = T;
    q.vtree_count = fd_vtrees;
    fd_plan = Dispatch(e, q, false, rescaling, false, 0);
  }
  // Partial-sum rows: [vtree][chunk][warp]; both launches share one chunk count
  // so that a single row stride addresses the buffers.
  const int chunks = std::max(plan.chunks, fd_vtrees > 0 ? fd_plan.chunks : 1);
  b->chunks = chunks;
  const size_t rows = static_cast<size_t>(vtrees) * chunks * kWarps;
  b->logl_partial.Reserve(rows);
  SBNB_CUDA(cudaMemsetAsync(b->logl_partial.get(), 0, rows * sizeof(double), s));
  if (grad) {
    const size_t grad_rows = static_cast<size_t>(T) * chunks * kWarps * N;
    b->grad_partial.Reserve(grad_rows);
    SBNB_CUDA(cudaMemsetAsync(b->grad_partial.get(), 0, grad_rows * sizeof(double), s));
    if (C > 1) {
      b->rgrad_partial.Reserve(grad_rows);
      SBNB_CUDA(cudaMemsetAsync(b->rgrad_partial.get(), 0, grad_rows * sizeof(double), s));
    }
  }
  p.logl_partial = b->logl_partial.get();
  p.grad_partial = b->grad_partial.get();
  p.rgrad_partial = b->rgrad_partial.get();
  const int ring = static_cast<int>(e->walk_runs++ % sbnb_engine::kWalkRing);
  HarvestWalkTiming(e, ring);  // the slot about to be reused
  SBNB_CUDA(cudaEventRecord(e->walk_begin[ring], s));
  Dispatch(e, p, grad, rescaling, /*launch=*/true, chunks);
  SBNB_CUDA(cudaEventRecord(e->walk_end[ring], s));
  e->walk_pending[ring] = true;
  if (fd_vtrees > 0) {
    WalkParams q = p;
    q.vtree_begin = T;
    q.vtree_count = fd_vtrees;
    Dispatch(e, q, false, rescaling, true, chunks);
  }

  // K3: fixed-order reduction of the per-(chunk, warp) partial sums.
  {
    const int block = 128;
    ReducePartialsKernel<<<(vtrees + block - 1) / block, block, 0, s>>>(
        b->logl_partial.get(), b->logl.get(), 0, vtrees, chunks * kWarps, 1);
    SBNB_CUDA(cudaGetLastError());
    e->launch_count++;
    if (grad) {
      const int64_t total = static_cast<int64_t>(T) * N;
      const int grid = static_cast<int>((total + block - 1) / block);
      ReducePartialsKernel<<<grid, block, 0, s>>>(b->grad_partial.get(), b->grad.get(), 0, T,
                                                  chunks * kWarps, N);
      SBNB_CUDA(cudaGetLastError());
      e->launch_count++;
      if (C > 1) {
        ReducePartialsKernel<<<grid, block, 0, s>>>(b->rgrad_partial.get(), b->rgrad.get(), 0, T,
                                                    chunks * kWarps, N);
        SBNB_CUDA(cudaGetLastError());
        e->launch_count++;
      }
    }
  }
}

void Fetch(sbnb_engine* e, sbnb_batch* b, double* logl, double* grad, double* rgrad) {
  SBNB_CUDA(cudaSetDevice(e->device));
  Require(b->last_mode >= 0, "sbnb_batch_fetch called before sbnb_batch_run.");
  const bool was_grad = (b->last_mode == SBNB_MODE_BRANCH_GRADIENT);
  const int vtrees = was_grad ? b->vtree_count : b->tree_count;
  const size_t per_tree = static_cast<size_t>(b->node_count);
  cudaStream_t s = e->stream;
  if (b->tree_count > 0) {
    if (logl) {
      SBNB_CUDA(cudaMemcpyAsync(logl, b->logl.get(), vtrees * sizeof(double), cudaMemcpyDeviceToHost, s));
      e->d2h_bytes += vtrees * sizeof(double);
    }
    if (grad) {
      Require(was_grad, "No gradient available: the last run was a log-likelihood run.");
      SBNB_CUDA(cudaMemcpyAsync(grad, b->grad.get(), b->tree_count * per_tree * sizeof(double),
                                cudaMemcpyDeviceToHost, s));
      e->d2h_bytes += b->tree_count * per_tree * sizeof(double);
    }
    if (rgrad) {
      Require(was_grad, "No gradient available: the last run was a log-likelihood run.");
      if (e->padded_categories > 1) {
        SBNB_CUDA(cudaMemcpyAsync(rgrad, b->rgrad.get(), b->tree_count * per_tree * sizeof(double),
                                  cudaMemcpyDeviceToHost, s));
        e->d2h_bytes += b->tree_count * per_tree * sizeof(double);
      } else {
        std::fill(rgrad, rgrad + b->tree_count * per_tree, 0.0);
      }
    }
  }
  SBNB_CUDA(cudaStreamSynchronize(s));
}

RootedView ViewOf(const sbnb_tree_batch* trees, int t, int n) {
  const int N = 2 * n - 1;
  RootedView view{};
  view.branch_lengths = trees->branch_lengths + static_cast<size_t>(t) * N;
  view.rates = trees->rates + static_cast<size_t>(t) * (N - 1);
  view.node_heights = trees->node_heights ? trees->node_heights + static_cast<size_t>(t) * N : nullptr;
  view.node_bounds = trees->node_bounds ? trees->node_bounds + static_cast<size_t>(t) * N : nullptr;
  view.height_ratios =
      trees->height_ratios ? trees->height_ratios + static_cast<size_t>(t) * (n - 1) : nullptr;
  view.rate_count = trees->rate_count;
  return view;
}

// The one-call forms of Engine's five methods.
void LogLikelihoods(sbnb_engine* e, const sbnb_tree_batch* trees, const double* params,
                    bool rescaling, bool rooted_semantics, bool add_jacobian, double* out) {
  Require(out != nullptr || trees->tree_count == 0, "NULL output.");
  auto batch = Stage(e, trees, params, rooted_semantics, false);
  Run(e, batch.get(), SBNB_MODE_LOG_LIKELIHOOD, rescaling);
  Fetch(e, batch.get(), out, nullptr, nullptr);
  if (add_jacobian) {
    Require(trees->tree_count == 0 || (trees->node_heights && trees->node_bounds),
            "Rooted log likelihoods need node_heights and node_bounds.");
    for (int t = 0; t < trees->tree_count; t++)
      out[t] += LogDetJacobianHeightRatios(batch->programs[t], ViewOf(trees, t, e->taxon_count));
  }
}

// Host finishing of a gradient evaluation (the O(n) tail of FatBeagle::Gradient,
// fat_beagle.cpp:467-545) from the raw per-tree sums over site patterns:
// logl[T (+ T * 2 * fd_coords)], grad[T][N] edge derivatives, rgrad[T][N] the same
// with d rate_c / d shape as scalers.  Needs no device: every input is a sum over
// patterns, so under site-pattern sharding the ranks all-reduce the raw sums
// and each runs this once.
void FinishGradients(const ModelSpec& spec, int n, const sbnb_tree_batch* trees, bool rooted, int fd_coords,
                     const double* logl, const double* grad, const double* rgrad,
                     const sbnb_gradient_out* out, const std::vector<TreeProgram>* programs = nullptr) {
  Require(out != nullptr, "NULL gradient output.");
  const int T = trees->tree_count, N = 2 * n - 1;
  if (rooted)
    Require(T == 0 || (trees->node_heights && trees->node_bounds && trees->height_ratios),
            "Rooted gradients need node_heights, node_bounds and height_ratios.");
  const int categories = spec.category_count;
  ParallelOverTrees(T, [&](int t) {
    std::vector<double> g(N), lengths(N);
    TreeProgram rebuilt;
    if (!programs)
      rebuilt = BuildTreeProgram(trees->parent_ids + static_cast<size_t>(t) * (trees->node_count - 1),
                                 trees->node_count, n);
    const TreeProgram& tree = programs ? (*programs)[t] : rebuilt;
    std::copy(grad + static_cast<size_t>(t) * N, grad + static_cast<size_t>(t + 1) * N, g.begin());
    if (out->log_likelihood) out->log_likelihood[t] = logl[t];
    if (out->substitution_model && fd_coords > 0) {
      // Central differences; the rooted Jacobian term cancels (fat_beagle.cpp:431).
      const double* fd = logl + T + static_cast<size_t>(t) * 2 * fd_coords;
      for (int k = 0; k < fd_coords; k++)
        out->substitution_model[static_cast<size_t>(t) * fd_coords + k] =
            (fd[2 * k] - fd[2 * k + 1]) / (2. * kFiniteDifferenceDelta);
    }
    if (out->site_model && categories > 1) {
      EffectiveBranchLengths(tree, trees, t, rooted, N, lengths.data());
      out->site_model[t] =
          DiscreteSiteModelGradient(N, lengths.data(), rgrad + static_cast<size_t>(t) * N);  // fat_beagle.cpp:389-398
    }
    if (rooted) {
      const RootedView view = ViewOf(trees, t, n);
      if (out->ratios_root_height) {
        const std::vector<double> ratios = RatioGradientOfBranchGradient(tree, view, g.data());
        std::copy(ratios.begin(), ratios.end(), out->ratios_root_height + static_cast<size_t>(t) * (n - 1));
      }
      if (out->clock_model) {
        const std::vector<double> clock = ClockGradient(tree, view, g.data());
        std::copy(clock.begin(), clock.end(), out->clock_model + static_cast<size_t>(t) * view.rate_count);
      }
    } else if (out->branch_lengths) {
      // "We want the fixed node to have a zero gradient" (fat_beagle.cpp:498-500).
      g[tree.child1[tree.root]] = 0.0;
      std::copy(g.begin(), g.end(), out->branch_lengths + static_cast<size_t>(t) * N);
    }
  });
}

void Gradients(sbnb_engine* e, const sbnb_tree_batch* trees, const double* params, bool rescaling,
               bool rooted, const sbnb_gradient_out* out) {
  Require(out != nullptr, "NULL gradient output.");
  const int fd_coords = e->spec.SubstitutionGradientSize();
  auto batch = Stage(e, trees, params, rooted, fd_coords > 0 && out->substitution_model != nullptr);
  Run(e, batch.get(), SBNB_MODE_BRANCH_GRADIENT, rescaling);
  const int T = trees->tree_count, N = 2 * e->taxon_count - 1;
  std::vector<double> logl(batch->vtree_count), grad(static_cast<size_t>(T) * N),
      rgrad(static_cast<size_t>(T) * N);
  Fetch(e, batch.get(), logl.data(), grad.data(), rgrad.data());
  FinishGradients(e->spec, e->taxon_count, trees, rooted, batch->fd_coords, logl.data(), grad.data(),
                  rgrad.data(), out, &batch->programs);
}

}  // namespace

extern "C" {

const char* sbnb_last_error(void) { return g_last_error.c_str(); }

int sbnb_device_count(void) {
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return count;
}

int sbnb_engine_create(const char* substitution, const char* site, const char* clock,
                       int32_t taxon_count, int64_t pattern_count, const uint8_t* tip_states,
                       const double* pattern_weights, int32_t device, sbnb_engine** out) {
  return Guard([&] {
    Require(out != nullptr, "NULL output handle.");
    *out = nullptr;
    Require(substitution && site && clock, "NULL model specification string.");
    Require(taxon_count >= 2, "Need at least 2 taxa.");
    Require(pattern_count >= 1, "Need at least 1 site pattern.");
    Require(tip_states && pattern_weights, "NULL tip_states / pattern_weights.");
    auto engine = std::make_unique<sbnb_engine>();
    engine->spec = ModelSpec::Parse(substitution, site, clock);
    int device_count = 0;
    if (cudaGetDeviceCount(&device_count) != cudaSuccess || device_count < 1) {
      cudaGetLastError();
      Fail(SBNB_ERR_NO_DEVICE,
           "No CUDA device available: libsbn_b200 has no CPU fallback (cudaGetDeviceCount).");
    }
    Require(device >= 0 && device < device_count, "CUDA device ordinal out of range.");
    SBNB_CUDA(cudaSetDevice(device));
    engine->device = device;
    cudaDeviceProp prop{};
    SBNB_CUDA(cudaGetDeviceProperties(&prop, device));
    engine->sm_count = prop.multiProcessorCount;
    engine->taxon_count = taxon_count;
    engine->pattern_count = pattern_count;
    engine->range_begin = 0;
    engine->range_end = pattern_count;
    engine->categories = engine->spec.category_count;
    engine->padded_categories = PadCategories(engine->categories);
    Require(engine->padded_categories > 0, "At most 16 rate categories are supported.");
    SBNB_CUDA(cudaStreamCreateWithFlags(&engine->stream, cudaStreamNonBlocking));
    for (int i = 0; i < sbnb_engine::kWalkRing; i++) {
      SBNB_CUDA(cudaEventCreate(&engine->walk_begin[i]));
      SBNB_CUDA(cudaEventCreate(&engine->walk_end[i]));
    }
    engine->host_tips.resize(static_cast<size_t>(taxon_count) * pattern_count);
    for (size_t i = 0; i < engine->host_tips.size(); i++)
      engine->host_tips[i] = tip_states[i] < 4 ? tip_states[i] : 4;
    engine->host_weights.assign(pattern_weights, pattern_weights + pattern_count);
    UploadPatternRange(engine.get(), 0, pattern_count);
    *out = engine.release();
  });
}

void sbnb_engine_destroy(sbnb_engine* engine) {
  if (!engine) return;
  FloatingPointEnvironmentKeeper keep_caller_environment;
  cudaSetDevice(engine->device);
  delete engine;
}

int32_t sbnb_engine_param_count(const sbnb_engine* engine) {
  return engine ? engine->spec.param_count : -1;
}

int sbnb_engine_param_block(const sbnb_engine* engine, const char* key, int32_t* start,
                            int32_t* length) {
  return Guard([&] {
    Require(engine && key && start && length, "NULL argument.");
    const auto block = engine->spec.Block(key);
    *start = block.first;
    *length = block.second;
  });
}

int32_t sbnb_engine_category_count(const sbnb_engine* engine) {
  return engine ? engine->categories : -1;
}

int sbnb_log_likelihoods_unrooted(sbnb_engine* engine, const sbnb_tree_batch* trees,
                                  const double* params, int32_t rescaling, double* out) {
  return Guard([&] {
    Require(engine != nullptr, "NULL engine.");
    LogLikelihoods(engine, trees, params, rescaling != 0, false, false, out);
  });
}

int sbnb_log_likelihoods_rooted(sbnb_engine* engine, const sbnb_tree_batch* trees,
                                const double* params, int32_t rescaling, double* out) {
  return Guard([&] {
    Require(engine != nullptr, "NULL engine.");
    LogLikelihoods(engine, trees, params, rescaling != 0, true, true, out);
  });
}

int sbnb_unrooted_log_likelihoods_of_rooted(sbnb_engine* engine, const sbnb_tree_batch* trees,
                                            const double* params, int32_t rescaling, double* out) {
  return Guard([&] {
    Require(engine != nullptr, "NULL engine.");
    // fat_beagle.cpp:78-80: plain branch lengths, no rates, no Jacobian.
    Require(trees && trees->node_count == 2 * engine->taxon_count - 1,
            "Rooted trees must be bifurcating: node_count must be 2n-1.");
    LogLikelihoods(engine, trees, params, rescaling != 0, false, false, out);
  });
}

int sbnb_gradients_unrooted(sbnb_engine* engine, const sbnb_tree_batch* trees, const double* params,
                            int32_t rescaling, const sbnb_gradient_out* out) {
  return Guard([&] {
    Require(engine != nullptr, "NULL engine.");
    Gradients(engine, trees, params, rescaling != 0, false, out);
  });
}

int sbnb_gradients_rooted(sbnb_engine* engine, const sbnb_tree_batch* trees, const double* params,
                          int32_t rescaling, const sbnb_gradient_out* out) {
  return Guard([&] {
    Require(engine != nullptr, "NULL engine.");
    Gradients(engine, trees, params, rescaling != 0, true, out);
  });
}

int sbnb_finish_gradients(const char* substitution, const char* site, const char* clock,
                          int32_t taxon_count, const sbnb_tree_batch* trees, int32_t rooted,
                          int32_t with_substitution_fd, const double* log_likelihoods,
                          const double* branch_gradients, const double* rate_gradients,
                          const sbnb_gradient_out* out) {
  return Guard([&] {
    Require(substitution && site && clock, "NULL model specification string.");
    Require(trees != nullptr, "NULL tree batch.");
    Require(taxon_count >= 3, "Need at least 3 taxa.");
    const ModelSpec spec = ModelSpec::Parse(substitution, site, clock);
    Require(trees->tree_count == 0 || (trees->parent_ids && trees->branch_lengths),
            "NULL parent_ids / branch_lengths.");
    Require(trees->tree_count == 0 || (log_likelihoods && branch_gradients),
            "NULL log_likelihoods / branch_gradients.");
    Require(spec.category_count == 1 || trees->tree_count == 0 || rate_gradients != nullptr,
            "rate_gradients are required for a multi-category site model.");
    Require(!rooted || trees->tree_count == 0 || trees->rates != nullptr,
            "Rooted evaluation needs per-branch rates (RootedTree::rates_).");
    FinishGradients(spec, taxon_count, trees, rooted != 0,
                    with_substitution_fd ? spec.SubstitutionGradientSize() : 0, log_likelihoods,
                    branch_gradients, rate_gradients, out);
  });
}

int sbnb_finish_log_likelihoods_rooted(int32_t taxon_count, const sbnb_tree_batch* trees,
                                       double* log_likelihoods) {
  return Guard([&] {
    Require(trees != nullptr, "NULL tree batch.");
    Require(trees->tree_count == 0 || (log_likelihoods && trees->node_heights && trees->node_bounds),
            "Rooted log likelihoods need node_heights and node_bounds.");
    for (int t = 0; t < trees->tree_count; t++) {
      const TreeProgram tree = BuildTreeProgram(
          trees->parent_ids + static_cast<size_t>(t) * (trees->node_count - 1), trees->node_count, taxon_count);
      log_likelihoods[t] += LogDetJacobianHeightRatios(tree, ViewOf(trees, t, taxon_count));
    }
  });
}

int sbnb_batch_stage(sbnb_engine* engine, const sbnb_tree_batch* trees, const double* params,
                     int32_t stage_flags, sbnb_batch** out) {
  return Guard([&] {
    Require(engine && out, "NULL argument.");
    *out = nullptr;
    *out = Stage(engine, trees, params, (stage_flags & SBNB_STAGE_ROOTED) != 0,
                 (stage_flags & SBNB_STAGE_SUBSTITUTION_FD) != 0)
               .release();
  });
}

int sbnb_batch_run(sbnb_engine* engine, sbnb_batch* batch, int32_t mode, int32_t rescaling) {
  return Guard([&] {
    Require(engine && batch, "NULL argument.");
    Run(engine, batch, mode, rescaling != 0);
  });
}

int sbnb_batch_fetch(sbnb_engine* engine, sbnb_batch* batch, double* log_likelihoods,
                     double* branch_gradients, double* rate_gradients) {
  return Guard([&] {
    Require(engine && batch, "NULL argument.");
    Fetch(engine, batch, log_likelihoods, branch_gradients, rate_gradients);
  });
}

int sbnb_batch_device_results(sbnb_batch* batch, void** log_likelihoods, void** branch_gradients,
                              void** rate_gradients) {
  return Guard([&] {
    Require(batch != nullptr, "NULL batch.");
    if (log_likelihoods) *log_likelihoods = batch->logl.get();
    if (branch_gradients) *branch_gradients = batch->grad.get();
    if (rate_gradients) *rate_gradients = batch->rgrad.get();
  });
}

void sbnb_batch_destroy(sbnb_engine* engine, sbnb_batch* batch) {
  FloatingPointEnvironmentKeeper keep_caller_environment;
  if (engine) cudaSetDevice(engine->device);
  Recycle(engine, batch);
}

int32_t sbnb_batch_evaluation_count(const sbnb_batch* batch) {
  if (!batch) return -1;
  return batch->last_mode == SBNB_MODE_BRANCH_GRADIENT ? batch->vtree_count : batch->tree_count;
}

void* sbnb_engine_stream(sbnb_engine* engine) { return engine ? engine->stream : nullptr; }

int64_t sbnb_engine_launch_count(const sbnb_engine* engine) {
  return engine ? engine->launch_count : -1;
}

int sbnb_engine_transfer_bytes(const sbnb_engine* engine, int64_t* host_to_device,
                               int64_t* device_to_host) {
  return Guard([&] {
    Require(engine && host_to_device && device_to_host, "NULL argument.");
    *host_to_device = engine->h2d_bytes;
    *device_to_host = engine->d2h_bytes;
  });
}

int sbnb_engine_walk_timing(sbnb_engine* engine, double* total_ms, int64_t* samples,
                            int32_t reset) {
  return Guard([&] {
    Require(engine && total_ms && samples, "NULL argument.");
    SBNB_CUDA(cudaSetDevice(engine->device));
    // oldest first, so that walk_last_ms ends up being the newest run
    for (int i = 0; i < sbnb_engine::kWalkRing; i++)
      HarvestWalkTiming(engine, static_cast<int>((engine->walk_runs + i) % sbnb_engine::kWalkRing));
    *total_ms = engine->walk_total_ms;
    *samples = engine->walk_samples;
    if (reset) {
      engine->walk_total_ms = 0.0;
      engine->walk_samples = 0;
    }
  });
}

double sbnb_batch_algorithmic_bytes(const sbnb_batch* batch, int32_t mode) {
  // U = 32 C P bytes per partial; (2n-2) U per log-likelihood,
  // (10n-14) U per log-likelihood + branch gradient.
  if (!batch) return 0.0;
  const double n = batch->taxon_count;
  const double units = (mode == SBNB_MODE_BRANCH_GRADIENT) ? (10.0 * n - 14.0) : (2.0 * n - 2.0);
  return units * 32.0 * batch->categories * static_cast<double>(batch->patterns) * batch->tree_count;
}

int sbnb_engine_set_pattern_range(sbnb_engine* engine, int64_t begin, int64_t end) {
  return Guard([&] {
    Require(engine != nullptr, "NULL engine.");
    Require(0 <= begin && begin < end && end <= engine->pattern_count,
            "Pattern range must satisfy 0 <= begin < end <= pattern_count.");
    SBNB_CUDA(cudaSetDevice(engine->device));
    UploadPatternRange(engine, begin, end);
  });
}

int sbnb_debug_tree_program(const int32_t* parent_ids, int32_t node_count, int32_t taxon_count,
                            int32_t* post_ops, int32_t* pre_ops, int32_t* slots) {
  return Guard([&] {
    Require(parent_ids && post_ops && pre_ops && slots, "NULL argument.");
    const TreeProgram program = BuildTreeProgram(parent_ids, node_count, taxon_count);
    static_assert(sizeof(PostOp) == 32 && sizeof(PreOp) == 32, "ops are 8 x int32");
    std::memcpy(post_ops, program.post.data(), program.post.size() * sizeof(PostOp));
    std::memcpy(pre_ops, program.pre.data(), program.pre.size() * sizeof(PreOp));
    slots[0] = program.post_slots;
    slots[1] = program.pre_slots;
  });
}

int sbnb_debug_model_tables(const char* substitution, const char* site, const char* clock,
                            const double* param_row, double* eigenvectors,
                            double* inverse_eigenvectors, double* eigenvalues, double* frequencies,
                            double* q, double* category_rates, double* category_weights,
                            double* category_rate_derivatives) {
  return Guard([&] {
    Require(substitution && site && clock, "NULL model specification string.");
    const ModelSpec spec = ModelSpec::Parse(substitution, site, clock);
    Require(spec.param_count == 0 || param_row != nullptr, "NULL parameter row.");
    ModelTables tables;
    BuildModelTables(spec, param_row, &tables);
    if (eigenvectors) std::copy(tables.evec, tables.evec + 16, eigenvectors);
    if (inverse_eigenvectors) std::copy(tables.ivec, tables.ivec + 16, inverse_eigenvectors);
    if (eigenvalues) std::copy(tables.eval, tables.eval + 4, eigenvalues);
    if (frequencies) std::copy(tables.freqs, tables.freqs + 4, frequencies);
    if (q) std::copy(tables.q, tables.q + 16, q);
    const int C = spec.category_count;
    if (category_rates) std::copy(tables.rates, tables.rates + C, category_rates);
    if (category_weights) std::copy(tables.weights, tables.weights + C, category_weights);
    if (category_rate_derivatives) std::copy(tables.drates, tables.drates + C, category_rate_derivatives);
  });
}

}  // extern "C"
