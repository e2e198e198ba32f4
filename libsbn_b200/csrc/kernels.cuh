// Device code of libsbn_b200 (sm_100a, fp64, no tensor cores -- 4x4 mat-vecs).
//
// Kernels:
//   TransitionMatrixKernel  per (tree, edge, category): P = V diag(exp(lambda r_c t)) V^-1, its
//                           transpose, and (Q P)^T, in the layout the tree walk stages
//                           [replaces beagleUpdateTransitionMatrices, fat_beagle.cpp:304-314, and
//                            beagleSetDifferentialMatrix, fat_beagle.cpp:128-131]
//   TreeWalkKernel          one pass over a tree for a tile of site patterns: the
//                           post-order partial updates, per-pattern power-of-two
//                           rescaling, the root log-likelihood, and (gradient mode)
//                           the pre-order pass fused with all edge derivatives
//                           [replaces beagleUpdatePartials, beagleUpdatePrePartials,
//                            beagleCalculateEdgeDerivatives, beagleCalculateRootLogLikelihoods,
//                            beagleResetScaleFactors, beagleSetPartials(root pre := pi);
//                            fat_beagle.cpp:50-70, 119-175]
//   ReducePartialsKernel    fixed-order sum of the per-(chunk, warp) partial sums
//
// Design: site patterns are independent, so a thread owns K
// patterns -- all rate categories of them -- and walks the WHOLE tree for them.
// The walk order (host-generated, Strahler-ordered, tree_program.cpp) keeps the
// result of the previous op in registers ("cur"); only nodes with two internal
// children touch a small stack (global memory, L2 resident).  In gradient mode
// the evolved post-order partials P_x L_x are streamed to a per-CTA scratch
// arena (written once, read once by the pre-order pass).
//
// Everything an op needs besides partials -- the transition matrices of both
// child edges for every category and the tip states of the tile -- is brought
// into a shared-memory ring by TMA bulk copies (cp.async.bulk +
// mbarrier), issued two ops ahead by one elected thread; all lanes of a warp
// read the same matrix element, so matrix loads are shared-memory broadcasts.
#ifndef SBNB_KERNELS_CUH_
#define SBNB_KERNELS_CUH_

#include <cuda_runtime.h>
#include <stdint.h>

#include "model.hpp"
#include "tree_program.hpp"

namespace sbnb {

constexpr int kThreads = 128;  // 4 warps per CTA; warps only meet at the ring's mbarriers
constexpr int kWarps = kThreads / 32;
constexpr int kStages = 8;        // operand ring depth (one stage per op)
constexpr int kPrefetchOps = 5;   // ops in flight ahead of the one being computed
constexpr int kScratchStages = 4;   // scratch ring depth (one stage per (pre-order op, category))
constexpr int kScratchPrefetch = 2; // category steps in flight ahead

// Per (tree, edge) block written by TransitionMatrixKernel, in doubles, C = categories:
//   [0        .. 16C)  P_c            row-major, c = 0..C-1   (internal child: y = P L)
//   [16C      .. 36C)  per c: P_c^T (row s = column s of P) then 4 ones
//                                      (tip child: y = column of P; gap: y = 1)
//   [36C      .. 56C)  per c: (Q P_c)^T then 4 zeros
//                                      (tip child: Q y; gap: Q 1 = 0)
constexpr int kEdgeDoublesPerCategory = 56;
constexpr int kTipTableDoubles = 20;  // per category: 4 state rows + the gap row

// One op of the walk: 16 bytes.  Post-order op (first n-1 of a program) and
// pre-order op (last n-1) share the layout.
//   x = child 0 node id, y = child 1 node id,
//   z = node id | flags << 24
//   w = post: push_slot | a_slot << 8 | b_slot << 16
//       pre:  pop_slot  | a_push_slot << 8 | b_push_slot << 16      (0xff = none)
typedef int4 WalkOp;
// flags beyond kALeaf | kBLeaf | kRoot (tree_program.hpp)
enum : int32_t {
  kStackBefore = 8,  // post: push cur first; pre: pop cur first
  kACur = 16,        // post: child 0's partial is cur; pre: child 0's pre-order partial stays in cur
  kBCur = 32
};

struct WalkParams {
  // alignment (device)
  const uint8_t* tips;  // [taxon][tip_pitch], padded with gap states
  int64_t tip_pitch;
  const double* weights;  // [tip_pitch] padded with zeros
  int64_t pattern_begin, pattern_end;
  int32_t taxon_count;
  // programs (device): [program][2(n-1)] = post-order ops then pre-order ops
  const WalkOp* ops;
  // virtual trees: vtree v uses program vtree_program[v], model vtree_model[v]
  int32_t vtree_begin, vtree_count;
  const int32_t* vtree_program;
  const int32_t* vtree_model;
  const ModelTables* models;
  const double* matrices;  // [vtree][2n-2][56 C]
  // tiling
  int32_t tiles_total, tiles_per_chunk, chunks;
  int32_t slots;  // stack depth
  // per-CTA arenas + outputs
  double2* stack;         // [grid][slots][K][C][2][kThreads]
  int32_t* stack_exps;    // [grid][slots][K][kThreads]              (rescaling)
  double2* scratch;       // [grid][n-1][C][K][2][kThreads]          (gradient mode)
  double* logl_partial;   // [vtree][chunk][warp]
  double* grad_partial;   // [vtree][chunk][warp][2n-1]               (gradient mode)
  double* rgrad_partial;  // same, with d rate_c / d shape as the scalers (C > 1)
};

// Shared memory of one CTA (host and device agree through these).
// Per-item model constants: Q[16], p_c[16], p_c r_c[16], p_c dr_c/dshape[16], pi[4].
constexpr int kModelSmemDoubles = 16 + 3 * kMaxCategories + 4;
__host__ __device__ constexpr int StageChildDoubles(int C) { return 2 * kTipTableDoubles * C; }
__host__ __device__ constexpr int StageBytes(int C, int K) {
  return 2 * StageChildDoubles(C) * 8 + 2 * kThreads * K;
}
// One scratch-ring stage: the evolved partials of both children of a pre-order op
// for one category, [child][j][half][tid] double2.
__host__ __device__ constexpr int ScratchStageBytes(int K) { return 2 * K * 2 * kThreads * 16; }
__host__ __device__ constexpr size_t WalkSmemBytes(int C, int K, bool grad) {
  return static_cast<size_t>(kStages) * StageBytes(C, K) + 2 * (kStages + kScratchStages) * 8 +
         kModelSmemDoubles * 8 + (grad ? static_cast<size_t>(kScratchStages) * ScratchStageBytes(K) : 0);
}

// ---------------------------------------------------------------------------
// mbarrier / TMA bulk-copy helpers (PTX; see blackwell_cuda_programming.md)

__device__ __forceinline__ uint32_t SmemAddress(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void MbarInit(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(SmemAddress(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void MbarExpectTx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(SmemAddress(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void MbarArrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(SmemAddress(bar)) : "memory");
}
__device__ __forceinline__ void MbarWait(uint64_t* bar, uint32_t parity) {
  const uint32_t address = SmemAddress(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(address), "r"(parity)
        : "memory");
  } while (!done);
}
// global -> shared bulk copy (TMA, no tensor map); bytes and both addresses are
// multiples of 16; completion is signalled on `bar` as transaction bytes.
__device__ __forceinline__ void BulkCopy(void* smem_dst, const void* global_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   SmemAddress(smem_dst)),
               "l"(global_src), "r"(bytes), "r"(SmemAddress(bar))
               : "memory");
}

// ---------------------------------------------------------------------------
// small helpers (everything is fully unrolled; partials live in registers)

__device__ __forceinline__ void Load4(const double* src, double (&x)[4]) {
  const double2 v0 = reinterpret_cast<const double2*>(src)[0];
  const double2 v1 = reinterpret_cast<const double2*>(src)[1];
  x[0] = v0.x, x[1] = v0.y, x[2] = v1.x, x[3] = v1.y;
}

// y = M x, M row-major in shared memory at a warp-uniform address (broadcast loads)
__device__ __forceinline__ void MatVecShared(const double* m, const double (&x)[4], double (&y)[4]) {
#pragma unroll
  for (int i = 0; i < 4; i++) {
    double row[4];
    Load4(m + 4 * i, row);
    y[i] = fma(row[3], x[3], fma(row[2], x[2], fma(row[1], x[1], row[0] * x[0])));
  }
}
// x points at K vectors `stride` apart (in units of double[4]).
template <int K>
__device__ __forceinline__ void MatVecSharedK(const double* m, const double (*x)[4], double (&y)[K][4],
                                              int stride) {
#pragma unroll
  for (int i = 0; i < 4; i++) {
    double row[4];
    Load4(m + 4 * i, row);
#pragma unroll
    for (int j = 0; j < K; j++) {
      const double* v = x[j * stride];
      y[j][i] = fma(row[3], v[3], fma(row[2], v[2], fma(row[1], v[1], row[0] * v[0])));
    }
  }
}
// y = M^T x
template <int K>
__device__ __forceinline__ void MatTVecSharedK(const double* m, const double (&x)[K][4], double (&y)[K][4]) {
#pragma unroll
  for (int i = 0; i < 4; i++) {
    double row[4];
    Load4(m + 4 * i, row);
#pragma unroll
    for (int j = 0; j < K; j++)
#pragma unroll
      for (int s = 0; s < 4; s++) y[j][s] = (i == 0) ? row[s] * x[j][0] : fma(row[s], x[j][i], y[j][s]);
  }
}

__device__ __forceinline__ double Dot4(const double (&a)[4], const double (&b)[4]) {
  return fma(a[3], b[3], fma(a[2], b[2], fma(a[1], b[1], a[0] * b[0])));
}

__device__ __forceinline__ double WarpSum(double v) {
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
  return v;
}

// Per-pattern power-of-two normalisation over all C x 4 entries: when the
// largest entry has dropped below 2^-kLazyBits, scales by 2^-e (e = its
// exponent) and returns e.  Exact (no rounding), so rescaled and unrescaled
// runs agree bit for bit until the final log.
constexpr int kLazyBits = 128;
template <int C>
__device__ __forceinline__ int Normalize(double (&v)[C][4]) {
  int hi = __double2hiint(v[0][0]);
#pragma unroll
  for (int c = 0; c < C; c++)
#pragma unroll
    for (int i = 0; i < 4; i++) hi = max(hi, __double2hiint(v[c][i]));
  const int biased = (hi >> 20) & 0x7ff;
  // zero / subnormal / inf / nan, or still large enough: leave as is
  if (biased == 0 || biased >= 1023 - kLazyBits) return 0;
  const double scale = __hiloint2double((2046 - biased) << 20, 0);
#pragma unroll
  for (int c = 0; c < C; c++)
#pragma unroll
    for (int i = 0; i < 4; i++) v[c][i] *= scale;
  return biased - 1023;
}

// ---------------------------------------------------------------------------

// One thread per (virtual tree, edge, category).
__global__ void TransitionMatrixKernel(const ModelTables* __restrict__ models,
                                       const int32_t* __restrict__ vtree_model,
                                       const int32_t* __restrict__ vtree_lengths,
                                       const double* __restrict__ branch_lengths,
                                       double* __restrict__ matrices, int32_t vtree_count,
                                       int32_t edge_count, int32_t node_count, int32_t C) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t total = static_cast<int64_t>(vtree_count) * edge_count * C;
  if (idx >= total) return;
  const int c = static_cast<int>(idx % C);
  const int e = static_cast<int>((idx / C) % edge_count);
  const int v = static_cast<int>(idx / (static_cast<int64_t>(C) * edge_count));
  const ModelTables& model = models[vtree_model[v]];
  const double t =
      branch_lengths[static_cast<int64_t>(vtree_lengths[v]) * node_count + e] * model.rates[c];
  double ex[4];
#pragma unroll
  for (int k = 0; k < 4; k++) ex[k] = exp(model.eval[k] * t);
  double P[16];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) {
      double sum = 0.0;
#pragma unroll
      for (int k = 0; k < 4; k++) sum += (model.evec[i * 4 + k] * ex[k]) * model.ivec[k * 4 + j];
      P[i * 4 + j] = sum > 0.0 ? sum : 0.0;  // BEAGLE clamps round-off negatives
    }
  double* edge = matrices + (static_cast<int64_t>(v) * edge_count + e) * kEdgeDoublesPerCategory * C;
  double2* out = reinterpret_cast<double2*>(edge + 16 * c);
#pragma unroll
  for (int x = 0; x < 8; x++) out[x] = make_double2(P[2 * x], P[2 * x + 1]);
  out = reinterpret_cast<double2*>(edge + 16 * C + kTipTableDoubles * c);
#pragma unroll
  for (int s = 0; s < 4; s++) {  // P^T: row s = column s of P
    out[2 * s] = make_double2(P[s], P[4 + s]);
    out[2 * s + 1] = make_double2(P[8 + s], P[12 + s]);
  }
  out[8] = make_double2(1.0, 1.0);
  out[9] = make_double2(1.0, 1.0);
  out = reinterpret_cast<double2*>(edge + 36 * C + kTipTableDoubles * c);
#pragma unroll
  for (int s = 0; s < 4; s++) {  // (Q P)^T: row s = column s of Q P
    double col[4];
#pragma unroll
    for (int i = 0; i < 4; i++)
      col[i] = fma(model.q[i * 4 + 3], P[12 + s],
                   fma(model.q[i * 4 + 2], P[8 + s], fma(model.q[i * 4 + 1], P[4 + s], model.q[i * 4] * P[s])));
    out[2 * s] = make_double2(col[0], col[1]);
    out[2 * s + 1] = make_double2(col[2], col[3]);
  }
  out[8] = make_double2(0.0, 0.0);
  out[9] = make_double2(0.0, 0.0);
}

// ---------------------------------------------------------------------------
// Thread tid handles patterns pat0 + j * kThreads + tid, j < K, of its tile, all
// C categories.  Arena rows are [..][j][c][half][tid] double2 -> every global
// access is a coalesced 16 B per lane.

template <int C, int K, bool GRAD, bool RESCALE>
__global__ void __launch_bounds__(kThreads, (K == 1 && C <= 4) ? 3 : 2) TreeWalkKernel(const WalkParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int n = p.taxon_count;
  const int internal_count = n - 1;
  const int edge_count = 2 * n - 2;
  const int node_count = 2 * n - 1;
  const int ops_total = GRAD ? 2 * internal_count : internal_count;
  constexpr int kTilePatterns = kThreads * K;
  constexpr int kStage = StageBytes(C, K);
  constexpr int kChild = StageChildDoubles(C);
  constexpr int kRow = 2 * kThreads;  // double2 per (j, c) block: [half][tid]
  // Fetch category c + 1's stack / scratch operands while category c is computed
  // (only where the registers for it exist).
  constexpr bool kFetchAhead = (K == 1);

  // ---- shared memory carve-up ------------------------------------------------
  unsigned char* const ring = smem_raw;
  uint64_t* const full = reinterpret_cast<uint64_t*>(smem_raw + kStages * kStage);
  uint64_t* const empty = full + kStages;
  uint64_t* const scratch_full = empty + kStages;
  uint64_t* const scratch_empty = scratch_full + kScratchStages;
  double* const q_smem = reinterpret_cast<double*>(scratch_empty + kScratchStages);
  double* const cat_weight_smem = q_smem + 16;
  double* const rate_weight_smem = cat_weight_smem + kMaxCategories;
  double* const drate_weight_smem = rate_weight_smem + kMaxCategories;
  double* const freqs_smem = drate_weight_smem + kMaxCategories;
  // (16-byte aligned: every region before it is a multiple of 16 bytes)
  double2* const scratch_ring = reinterpret_cast<double2*>(freqs_smem + 4);
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kStages; s++) {
      MbarInit(full + s, 1);
      MbarInit(empty + s, kWarps);
    }
#pragma unroll
    for (int s = 0; s < kScratchStages; s++) {
      MbarInit(scratch_full + s, 1);
      MbarInit(scratch_empty + s, kWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  double2* const my_stack =
      p.stack + static_cast<size_t>(blockIdx.x) * p.slots * K * C * kRow + tid;
  int32_t* const my_stack_exps =
      RESCALE ? p.stack_exps + static_cast<size_t>(blockIdx.x) * p.slots * K * kThreads + tid : nullptr;
  double2* const my_scratch =
      GRAD ? p.scratch + static_cast<size_t>(blockIdx.x) * internal_count * K * C * kRow + tid : nullptr;
  auto block_ptr = [&](double2* base, int index, int j, int c) -> double2* {
    return base + (static_cast<size_t>(index) * K + j) * C * kRow + c * kRow;
  };
  // scratch arena: [node][c][j][half][tid], so one (node, category) block is contiguous
  auto scratch_ptr = [&](int index, int j, int c) -> double2* {
    return my_scratch + ((static_cast<size_t>(index) * C + c) * K + j) * kRow;
  };
  uint32_t scratch_sequence = 0;  // (pre-order op, category) steps this CTA has consumed

  uint32_t sequence = 0;  // ops this CTA has consumed; stage = sequence % kStages

  const int64_t total_items = static_cast<int64_t>(p.vtree_count) * p.chunks;
  for (int64_t item = blockIdx.x; item < total_items; item += gridDim.x) {
    const int vt = p.vtree_begin + static_cast<int>(item / p.chunks);
    const int chunk = static_cast<int>(item % p.chunks);
    const ModelTables& model = p.models[p.vtree_model[vt]];
    const double* mats = p.matrices + static_cast<size_t>(vt) * edge_count * kEdgeDoublesPerCategory * C;
    const WalkOp* ops = p.ops + static_cast<size_t>(p.vtree_program[vt]) * 2 * internal_count;
    const size_t out_row = (static_cast<size_t>(vt) * p.chunks + chunk) * kWarps + warp;
    double* grad_row = GRAD ? p.grad_partial + out_row * node_count : nullptr;
    double* rgrad_row = (GRAD && C > 1) ? p.rgrad_partial + out_row * node_count : nullptr;
    __syncthreads();  // every warp is done with the previous item's model constants
    if (tid < 16) {
      q_smem[tid] = model.q[tid];
      cat_weight_smem[tid] = model.weights[tid];
      rate_weight_smem[tid] = model.weights[tid] * model.rates[tid];     // p_c r_c
      drate_weight_smem[tid] = model.weights[tid] * model.drates[tid];  // p_c dr_c/dshape
      if (tid < 4) freqs_smem[tid] = model.freqs[tid];
    }
    __syncthreads();

    // Producer (thread 0): bulk copies of one op's operands into its ring stage.
    auto issue = [&](uint32_t seq, const WalkOp& op, bool is_pre, int64_t tile_pat0) {
      const int s = seq % kStages;
      if (seq >= kStages) MbarWait(empty + s, ((seq / kStages) - 1) & 1);
      unsigned char* stage = ring + s * kStage;
      const int flags = op.z >> 24;
      uint32_t bytes = 0;
#pragma unroll
      for (int child = 0; child < 2; child++) {
        const bool leaf = flags & (child ? kBLeaf : kALeaf);
        bytes += leaf ? (is_pre ? 2 : 1) * kTipTableDoubles * C * 8 + kTilePatterns : 16 * C * 8;
      }
      MbarExpectTx(full + s, bytes);
#pragma unroll
      for (int child = 0; child < 2; child++) {
        const int node = child ? op.y : op.x;
        const bool leaf = flags & (child ? kBLeaf : kALeaf);
        const double* edge = mats + static_cast<size_t>(node) * kEdgeDoublesPerCategory * C;
        double* dst = reinterpret_cast<double*>(stage) + child * kChild;
        if (leaf) {
          BulkCopy(dst, edge + 16 * C, (is_pre ? 2 : 1) * kTipTableDoubles * C * 8, full + s);
          BulkCopy(stage + 2 * kChild * 8 + child * kTilePatterns,
                   p.tips + static_cast<int64_t>(node) * p.tip_pitch + tile_pat0, kTilePatterns, full + s);
        } else {
          BulkCopy(dst, edge, 16 * C * 8, full + s);
        }
      }
    };

    // Producer (thread 0): one category step of a pre-order op -- the evolved
    // partials of its internal children, 2 K KB each -- into the scratch ring.
    auto issue_scratch = [&](uint32_t seq, int local_step) {
      const int s = seq % kScratchStages;
      if (seq >= kScratchStages) MbarWait(scratch_empty + s, ((seq / kScratchStages) - 1) & 1);
      const WalkOp op = __ldg(ops + internal_count + local_step / C);
      const int c = local_step % C;
      const int flags = op.z >> 24;
      constexpr uint32_t kBlockBytes = K * kRow * 16;
      const uint32_t bytes = ((flags & kALeaf) ? 0 : kBlockBytes) + ((flags & kBLeaf) ? 0 : kBlockBytes);
      if (bytes == 0) {
        MbarArrive(scratch_full + s);
        return;
      }
      MbarExpectTx(scratch_full + s, bytes);
      double2* stage = scratch_ring + static_cast<size_t>(s) * (2 * K * kRow);
      if (!(flags & kALeaf)) BulkCopy(stage, scratch_ptr(op.x - n, 0, c) - tid, kBlockBytes, scratch_full + s);
      if (!(flags & kBLeaf))
        BulkCopy(stage + K * kRow, scratch_ptr(op.y - n, 0, c) - tid, kBlockBytes, scratch_full + s);
    };

    double logl_acc = 0.0;
    const int tile_begin = chunk * p.tiles_per_chunk;
    const int tile_end = min(tile_begin + p.tiles_per_chunk, p.tiles_total);
    for (int tile = tile_begin; tile < tile_end; tile++) {
      const int64_t pat0 = p.pattern_begin + static_cast<int64_t>(tile) * kTilePatterns;
      double w[K];
#pragma unroll
      for (int j = 0; j < K; j++) {
        const int64_t pattern = pat0 + j * kThreads + tid;
        w[j] = (pattern < p.pattern_end) ? p.weights[pattern] : 0.0;
      }

      // ---- pipeline prologue: the first kPrefetchOps ops of this tile --------
      WalkOp ahead = make_int4(0, 0, 0, 0);  // record of op o + kPrefetchOps (thread 0)
      if (tid == 0) {
        for (int o = 0; o < kPrefetchOps && o < ops_total; o++)
          issue(sequence + o, __ldg(ops + o), GRAD && o >= internal_count, pat0);
        ahead = __ldg(ops + min(kPrefetchOps, ops_total - 1));
      }
      __syncwarp();
      WalkOp op_next = __ldg(ops);

      double cur[K][C][4];
      int cur_exp[K];
#pragma unroll
      for (int j = 0; j < K; j++) {
        cur_exp[j] = 0;
#pragma unroll
        for (int c = 0; c < C; c++)
#pragma unroll
          for (int i = 0; i < 4; i++) cur[j][c][i] = 0.0;
      }

      for (int o = 0; o < ops_total; o++, sequence++) {
        if (tid == 0 && o + kPrefetchOps < ops_total) {
          issue(sequence + kPrefetchOps, ahead, GRAD && (o + kPrefetchOps >= internal_count), pat0);
          ahead = __ldg(ops + min(o + kPrefetchOps + 1, ops_total - 1));
        }
        __syncwarp();
        const WalkOp op = op_next;
        op_next = __ldg(ops + min(o + 1, ops_total - 1));
        if (GRAD && o + 1 >= internal_count && o + 1 < ops_total) {
          // The next pre-order op reads its internal children's evolved partials
          // back from the scratch arena: start them on their way from HBM to L2 now.
          const int next_flags = op_next.z >> 24;
          constexpr int kBlockLines = K * C * kRow * 16 / 128;
#pragma unroll
          for (int child = 0; child < 2; child++) {
            if (next_flags & (child ? kBLeaf : kALeaf)) continue;
            const char* block = reinterpret_cast<const char*>(
                scratch_ptr((child ? op_next.y : op_next.x) - n, 0, 0) - tid);
#pragma unroll
            for (int line = 0; line < kBlockLines; line += kThreads)
              asm volatile("prefetch.global.L2 [%0];" ::"l"(block + static_cast<size_t>(line + tid) * 128));
          }
        }
        const int stage_index = sequence % kStages;
        MbarWait(full + stage_index, (sequence / kStages) & 1);
        const unsigned char* stage = ring + stage_index * kStage;
        const double* MA = reinterpret_cast<const double*>(stage);
        const double* MB = MA + kChild;
        const uint8_t* tips_a = stage + 2 * kChild * 8 + tid;
        const uint8_t* tips_b = tips_a + kTilePatterns;

        const int a = op.x, b = op.y;
        const int node = op.z & 0xffffff, flags = op.z >> 24;
        const int s0 = op.w & 0xff, s1 = (op.w >> 8) & 0xff, s2 = (op.w >> 16) & 0xff;
        const bool a_leaf = flags & kALeaf, b_leaf = flags & kBLeaf;

        if (!GRAD || o < internal_count) {
          // ======================= post-order op ===========================
          // cur = (P_a L_a) o (P_b L_b)
          if (flags & kStackBefore) {
#pragma unroll
            for (int j = 0; j < K; j++) {
#pragma unroll
              for (int c = 0; c < C; c++) {
                double2* dst = block_ptr(my_stack, s0, j, c);
                dst[0] = make_double2(cur[j][c][0], cur[j][c][1]);
                dst[kThreads] = make_double2(cur[j][c][2], cur[j][c][3]);
              }
              if (RESCALE) my_stack_exps[(s0 * K + j) * kThreads] = cur_exp[j];
            }
          }
          int tip_a[K], tip_b[K];
#pragma unroll
          for (int j = 0; j < K; j++) {
            tip_a[j] = a_leaf ? tips_a[j * kThreads] : 0;
            tip_b[j] = b_leaf ? tips_b[j * kThreads] : 0;
          }
          int popped_exp[K];
#pragma unroll
          for (int j = 0; j < K; j++) popped_exp[j] = 0;
          if (RESCALE) {
            const bool a_stack = !a_leaf && !(flags & kACur), b_stack = !b_leaf && !(flags & kBCur);
            if (a_stack || b_stack) {
              const int slot = a_stack ? s1 : s2;
#pragma unroll
              for (int j = 0; j < K; j++) popped_exp[j] = my_stack_exps[(slot * K + j) * kThreads];
            }
            const bool uses_cur = (flags & (kACur | kBCur)) != 0;
#pragma unroll
            for (int j = 0; j < K; j++) cur_exp[j] = (uses_cur ? cur_exp[j] : 0) + popped_exp[j];
          }
          // At most one operand comes off the stack (the other one is cur).
          // The category loop is a real loop -- unrolled it would not fit the
          // instruction cache -- so cur is rotated through its first slot: each
          // pass consumes cur[.][0] and appends the new value at the back; after C
          // passes every category is back in place.
          const bool a_pop = !a_leaf && !(flags & kACur), b_pop = !b_leaf && !(flags & kBCur);
          const double2* popped = block_ptr(my_stack, a_pop ? s1 : s2, 0, 0);
          const double* PA = MA;  // this category's block of child 0 / child 1
          const double* PB = MB;
#pragma unroll 1
          for (int c = 0; c < C; c++) {
            double ya[K][4], yb[K][4], x[K][4];
            if (a_pop || b_pop) {
#pragma unroll
              for (int j = 0; j < K; j++) {
                const double2 v0 = popped[j * C * kRow], v1 = popped[j * C * kRow + kThreads];
                x[j][0] = v0.x, x[j][1] = v0.y, x[j][2] = v1.x, x[j][3] = v1.y;
              }
              popped += kRow;
            }
            // ---- child 0
            if (a_leaf) {
#pragma unroll
              for (int j = 0; j < K; j++) Load4(PA + tip_a[j] * 4, ya[j]);
            } else {
              if (flags & kACur) {
                MatVecSharedK<K>(PA, cur[0], ya, C);
              } else {
                MatVecSharedK<K>(PA, x, ya, 1);
              }
              if (GRAD) {
#pragma unroll
                for (int j = 0; j < K; j++) {
                  double2* dst = scratch_ptr(a - n, j, c);
                  dst[0] = make_double2(ya[j][0], ya[j][1]);
                  dst[kThreads] = make_double2(ya[j][2], ya[j][3]);
                }
              }
            }
            // ---- child 1
            if (b_leaf) {
#pragma unroll
              for (int j = 0; j < K; j++) Load4(PB + tip_b[j] * 4, yb[j]);
            } else {
              if (flags & kBCur) {
                MatVecSharedK<K>(PB, cur[0], yb, C);
              } else {
                MatVecSharedK<K>(PB, x, yb, 1);
              }
              if (GRAD) {
#pragma unroll
                for (int j = 0; j < K; j++) {
                  double2* dst = scratch_ptr(b - n, j, c);
                  dst[0] = make_double2(yb[j][0], yb[j][1]);
                  dst[kThreads] = make_double2(yb[j][2], yb[j][3]);
                }
              }
            }
            PA += a_leaf ? kTipTableDoubles : 16;
            PB += b_leaf ? kTipTableDoubles : 16;
#pragma unroll
            for (int j = 0; j < K; j++) {
#pragma unroll
              for (int k = 0; k + 1 < C; k++)
#pragma unroll
                for (int i = 0; i < 4; i++) cur[j][k][i] = cur[j][k + 1][i];
#pragma unroll
              for (int i = 0; i < 4; i++) cur[j][C - 1][i] = ya[j][i] * yb[j][i];
            }
          }
          if (RESCALE) {
#pragma unroll
            for (int j = 0; j < K; j++) cur_exp[j] += Normalize<C>(cur[j]);
          }
          if (flags & kRoot) {
            // beagleCalculateRootLogLikelihoods: log sum_c p_c sum_i pi_i L[c,k,i] (+ scale)
            double freqs[4];
            Load4(freqs_smem, freqs);
#pragma unroll
            for (int j = 0; j < K; j++) {
              double site = 0.0;
#pragma unroll
              for (int c = 0; c < C; c++) site = fma(cat_weight_smem[c], Dot4(freqs, cur[j][c]), site);
              double log_site = log(site);
              if (RESCALE) log_site = fma(static_cast<double>(cur_exp[j]), 0.6931471805599453094, log_site);
              logl_acc = fma(w[j], (w[j] != 0.0) ? log_site : 0.0, logl_acc);
            }
          }
        } else {
          if (o == internal_count) {
            // The post-order pass wrote the scratch arena with ordinary stores; the
            // pre-order pass reads it back through TMA (the async proxy).
            asm volatile("fence.proxy.async;" ::: "memory");
            __syncthreads();
            if (tid == 0) {
              for (int step = 0; step < kScratchPrefetch && step < internal_count * C; step++)
                issue_scratch(scratch_sequence + step, step);
            }
            __syncwarp();
          }
          // ================ pre-order op + edge derivatives ================
          // cur = this node's pre-order partial pp (root: pi).  With y_x = P_x L_x
          // (read back from the scratch arena or looked up for a tip),
          // t_a = pp o y_b and t_b = pp o y_a, the children's pre-order partials are
          // P_a^T t_a and P_b^T t_b (beagleUpdatePrePartials), and because Q and P
          // commute the per-pattern derivative terms of edge a
          // (beagleCalculateEdgeDerivatives) are
          //   numerator   = pre_a^T Q L_a = t_a . (Q y_a)
          //   denominator = pre_a^T   L_a = t_a . y_a = pp . (y_a o y_b)   (shared by both edges)
          // so a tip edge needs no mat-vec at all: y_a and Q y_a are columns of P and Q P.
          if (flags & kRoot) {
            double freqs[4];
            Load4(freqs_smem, freqs);
#pragma unroll
            for (int j = 0; j < K; j++)
#pragma unroll
              for (int c = 0; c < C; c++)
#pragma unroll
                for (int i = 0; i < 4; i++) cur[j][c][i] = freqs[i];
          } else if (flags & kStackBefore) {
#pragma unroll
            for (int j = 0; j < K; j++)
#pragma unroll
              for (int c = 0; c < C; c++) {
                const double2* src = block_ptr(my_stack, s0, j, c);
                const double2 v0 = src[0], v1 = src[kThreads];
                cur[j][c][0] = v0.x, cur[j][c][1] = v0.y, cur[j][c][2] = v1.x, cur[j][c][3] = v1.y;
              }
          }
          if (RESCALE) {
#pragma unroll
            for (int j = 0; j < K; j++) Normalize<C>(cur[j]);
          }
          int tip_a[K], tip_b[K];
#pragma unroll
          for (int j = 0; j < K; j++) {
            tip_a[j] = a_leaf ? tips_a[j * kThreads] : 0;
            tip_b[j] = b_leaf ? tips_b[j * kThreads] : 0;
          }
          double den[K], num_a[K], num_b[K], rnum_a[K], rnum_b[K];
#pragma unroll
          for (int j = 0; j < K; j++) den[j] = num_a[j] = num_b[j] = rnum_a[j] = rnum_b[j] = 0.0;
          // A real loop over the categories, cur rotated through its first slot
          // (see the post-order op).
#pragma unroll 1
          for (int c = 0; c < C; c++) {
            // scratch ring: issue the step kScratchPrefetch ahead, wait for this one
            {
              const int local_step = (o - internal_count) * C + c;
              if (tid == 0 && local_step + kScratchPrefetch < internal_count * C)
                issue_scratch(scratch_sequence + kScratchPrefetch, local_step + kScratchPrefetch);
              __syncwarp();
            }
            const int scratch_stage = scratch_sequence % kScratchStages;
            // Waited for even when both children are tips (nothing was copied): a warp
            // that ran ahead through such steps could otherwise arrive twice in one
            // phase of scratch_empty.
            MbarWait(scratch_full + scratch_stage, (scratch_sequence / kScratchStages) & 1);
            const double2* evolved_a = scratch_ring + static_cast<size_t>(scratch_stage) * (2 * K * kRow) + tid;
            const double2* evolved_b = evolved_a + K * kRow;
            const double cat_weight = cat_weight_smem[c];
            const double rate_w = rate_weight_smem[c];
            const double drate_w = drate_weight_smem[c];
            const double* TA = MA + c * kTipTableDoubles;  // tip tables of this category
            const double* TB = MB + c * kTipTableDoubles;
            double ya[K][4], yb[K][4], da[K][4], db[K][4];
            if (a_leaf) {
#pragma unroll
              for (int j = 0; j < K; j++) {
                Load4(TA + tip_a[j] * 4, ya[j]);
                Load4(TA + C * kTipTableDoubles + tip_a[j] * 4, da[j]);
              }
            } else {
#pragma unroll
              for (int j = 0; j < K; j++) {
                const double2 v0 = evolved_a[j * kRow], v1 = evolved_a[j * kRow + kThreads];
                ya[j][0] = v0.x, ya[j][1] = v0.y, ya[j][2] = v1.x, ya[j][3] = v1.y;
              }
            }
            if (b_leaf) {
#pragma unroll
              for (int j = 0; j < K; j++) {
                Load4(TB + tip_b[j] * 4, yb[j]);
                Load4(TB + C * kTipTableDoubles + tip_b[j] * 4, db[j]);
              }
            } else {
#pragma unroll
              for (int j = 0; j < K; j++) {
                const double2 v0 = evolved_b[j * kRow], v1 = evolved_b[j * kRow + kThreads];
                yb[j][0] = v0.x, yb[j][1] = v0.y, yb[j][2] = v1.x, yb[j][3] = v1.y;
              }
            }
            // this stage's evolved partials are in registers: hand the stage back
            __syncwarp();
            if (lane == 0) MbarArrive(scratch_empty + scratch_stage);
            scratch_sequence++;
            if (!a_leaf) MatVecSharedK<K>(q_smem, ya, da, 1);
            if (!b_leaf) MatVecSharedK<K>(q_smem, yb, db, 1);
            double ta[K][4], tb[K][4];
#pragma unroll
            for (int j = 0; j < K; j++) {
#pragma unroll
              for (int i = 0; i < 4; i++) {
                ta[j][i] = cur[j][0][i] * yb[j][i];
                tb[j][i] = cur[j][0][i] * ya[j][i];
              }
              den[j] = fma(cat_weight, Dot4(ta[j], ya[j]), den[j]);
              const double na = Dot4(ta[j], da[j]), nb = Dot4(tb[j], db[j]);
              num_a[j] = fma(rate_w, na, num_a[j]);
              num_b[j] = fma(rate_w, nb, num_b[j]);
              if (C > 1) {
                rnum_a[j] = fma(drate_w, na, rnum_a[j]);
                rnum_b[j] = fma(drate_w, nb, rnum_b[j]);
              }
            }
            // children's pre-order partials: one stays in cur, the other is pushed
            double keep[K][4];
#pragma unroll
            for (int j = 0; j < K; j++)
#pragma unroll
              for (int i = 0; i < 4; i++) keep[j][i] = 0.0;
            if (!a_leaf) {
              double pre[K][4];
              MatTVecSharedK<K>(MA + 16 * c, ta, pre);
              if (flags & kACur) {
#pragma unroll
                for (int j = 0; j < K; j++)
#pragma unroll
                  for (int i = 0; i < 4; i++) keep[j][i] = pre[j][i];
              } else {
#pragma unroll
                for (int j = 0; j < K; j++) {
                  double2* dst = block_ptr(my_stack, s1, j, c);
                  dst[0] = make_double2(pre[j][0], pre[j][1]);
                  dst[kThreads] = make_double2(pre[j][2], pre[j][3]);
                }
              }
            }
            if (!b_leaf) {
              double pre[K][4];
              MatTVecSharedK<K>(MB + 16 * c, tb, pre);
              if (flags & kBCur) {
#pragma unroll
                for (int j = 0; j < K; j++)
#pragma unroll
                  for (int i = 0; i < 4; i++) keep[j][i] = pre[j][i];
              } else {
#pragma unroll
                for (int j = 0; j < K; j++) {
                  double2* dst = block_ptr(my_stack, s2, j, c);
                  dst[0] = make_double2(pre[j][0], pre[j][1]);
                  dst[kThreads] = make_double2(pre[j][2], pre[j][3]);
                }
              }
            }
#pragma unroll
            for (int j = 0; j < K; j++) {
#pragma unroll
              for (int k = 0; k + 1 < C; k++)
#pragma unroll
                for (int i = 0; i < 4; i++) cur[j][k][i] = cur[j][k + 1][i];
#pragma unroll
              for (int i = 0; i < 4; i++) cur[j][C - 1][i] = keep[j][i];
            }
          }
          // ---- per-pattern derivative terms, then one warp reduction per edge
          double ga = 0.0, gb = 0.0, ra = 0.0, rb = 0.0;
#pragma unroll
          for (int j = 0; j < K; j++) {
            const bool live = w[j] != 0.0;  // padding patterns contribute nothing (and may be 0/0)
            const double scale = live ? w[j] / den[j] : 0.0;
            ga = fma(scale, live ? num_a[j] : 0.0, ga);
            gb = fma(scale, live ? num_b[j] : 0.0, gb);
            if (C > 1) {
              ra = fma(scale, live ? rnum_a[j] : 0.0, ra);
              rb = fma(scale, live ? rnum_b[j] : 0.0, rb);
            }
          }
          ga = WarpSum(ga);
          gb = WarpSum(gb);
          if (C > 1) {
            ra = WarpSum(ra);
            rb = WarpSum(rb);
          }
          if (lane == 0) {
            // Single writer per (row, edge), in program order, so the sums are
            // deterministic; a reduction (no return value) keeps the round trip to
            // L2 off the warp's critical path.
            atomicAdd(grad_row + a, ga);
            atomicAdd(grad_row + b, gb);
            if (C > 1) {
              atomicAdd(rgrad_row + a, ra);
              atomicAdd(rgrad_row + b, rb);
            }
          }
        }
        __syncwarp();  // every lane is done reading this stage
        if (lane == 0) MbarArrive(empty + stage_index);
      }
    }
    // one partial per warp, in lane order
    logl_acc = WarpSum(logl_acc);
    if (lane == 0) p.logl_partial[out_row] = logl_acc;
  }
}

// out[v][e] = sum over the `parts` per-(chunk, warp) partial rows, in fixed order.
__global__ void ReducePartialsKernel(const double* __restrict__ partial, double* __restrict__ out,
                                     int32_t vtree_begin, int32_t vtree_count, int32_t parts,
                                     int32_t width) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<int64_t>(vtree_count) * width) return;
  const int64_t v = vtree_begin + idx / width;
  const int e = static_cast<int>(idx % width);
  const double* src = partial + v * parts * width + e;
  double sum = 0.0;
  for (int part = 0; part < parts; part++) sum += src[static_cast<int64_t>(part) * width];
  out[v * width + e] = sum;
}

}  // namespace sbnb

#endif  // SBNB_KERNELS_CUH_
