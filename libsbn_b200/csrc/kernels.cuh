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
// Design: site patterns are independent, so a warp owns a set of
// patterns and walks the WHOLE tree for them.  The walk order (host-generated,
// Strahler-ordered, tree_program.cpp) needs only O(log n) live partials, which
// live in a thread-private shared-memory stack; there is no __syncthreads
// anywhere.  In gradient mode the post-order partials are additionally streamed
// to a per-CTA scratch arena in global memory (written once, read once by the
// pre-order pass -- the only partial traffic that touches L2/HBM).
//
// The walk is a software pipeline: while op o is computed, the operands of op
// o+1 (transition matrices, scratch partials, tip states) are already in flight
// -- cp.async into a per-warp double-buffered staging area -- and the record of
// op o+2 is being loaded, so no global-memory latency sits on the critical path.
#ifndef SBNB_KERNELS_CUH_
#define SBNB_KERNELS_CUH_

#include <cuda_runtime.h>
#include <stdint.h>

#include "model.hpp"
#include "tree_program.hpp"

namespace sbnb {

constexpr int kThreads = 128;  // 4 independent warps per CTA
constexpr int kWarps = kThreads / 32;

// Per (tree, edge, category) block written by TransitionMatrixKernel, in doubles:
//   [ 0..15] P            row-major           (internal child: y = P L)
//   [16..31] P^T          row s = column s of P   (tip child: y = P[:, s])
//   [32..35] ones                              (tip with a gap: y = 1)
//   [36..51] (Q P)^T      row s = column s of Q P (tip child: dy = (Q P)[:, s])
//   [52..55] zeros                             (gap: Q P 1 = Q 1 = 0)
constexpr int kMatrixDoubles = 56;
constexpr int kMatrixChunks = kMatrixDoubles / 2;  // 16-byte chunks
constexpr int kTipBlockFirstChunk = 8;             // the tip part starts at P^T
// Staging stride per (child, category): 352 bytes = 88 words == 24 (mod 32), so the four
// category lanes of a warp hit disjoint 4-bank groups on every LDS.128.
constexpr int kStageDoubles = 44;

// One op of the walk: 16 bytes.  Post-order op (first n-1 of a program) and
// pre-order op (last n-1) share the layout.
//   x = child 0 node id, y = child 1 node id,
//   z = node id | flags << 24            (kALeaf | kBLeaf | kRoot)
//   w = post: dst_slot | child0_slot << 8 | child1_slot << 16
//       pre:  pre_slot | child0_dst_slot << 8 | child1_dst_slot << 16   (0xff = none)
typedef int4 WalkOp;

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
  const double* matrices;  // [vtree][2n-2][C][kMatrixDoubles]
  // tiling
  int32_t tiles_total, tiles_per_chunk, chunks;
  int32_t slots;  // shared-memory stack depth
  // scratch + outputs
  double2* scratch;       // [grid][n-1][K][2][kThreads]   (gradient mode)
  double* logl_partial;   // [vtree][chunk][warp]
  double* grad_partial;   // [vtree][chunk][warp][2n-1]     (gradient mode)
  double* rgrad_partial;  // same, with d rate_c / d shape as the scalers (C > 1)
};

// Dynamic shared memory of one CTA, in bytes (host and device agree through this).
__host__ __device__ constexpr size_t WalkStackBytes(int slots, int K) {
  return static_cast<size_t>(slots) * K * 2 * kThreads * sizeof(double2);
}
__host__ __device__ constexpr size_t WalkExpBytes(int slots, int K, bool rescale) {
  return rescale ? static_cast<size_t>(slots) * K * kThreads * sizeof(int) : 0;
}
__host__ __device__ constexpr size_t WalkMatStageBytes(int C) {
  return static_cast<size_t>(kWarps) * 2 * 2 * C * kStageDoubles * sizeof(double);
}
__host__ __device__ constexpr size_t WalkScratchStageBytes(int K, bool grad) {
  return grad ? static_cast<size_t>(2) * 2 * K * 2 * kThreads * sizeof(double2) : 0;
}
__host__ __device__ constexpr size_t WalkModelStageBytes(bool grad) {
  return grad ? static_cast<size_t>(kWarps) * 16 * sizeof(double) : 0;  // Q per warp
}
__host__ __device__ constexpr size_t WalkSmemBytes(int slots, int C, int K, bool grad, bool rescale) {
  return WalkStackBytes(slots, K) + WalkExpBytes(slots, K, rescale) + WalkMatStageBytes(C) +
         WalkScratchStageBytes(K, grad) + WalkModelStageBytes(grad);
}

// ---------------------------------------------------------------------------
// small helpers (everything is fully unrolled; matrices live in registers)

__device__ __forceinline__ void CpAsync16Cached(void* smem_dst, const void* global_src) {
  const unsigned dst = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(global_src) : "memory");
}
__device__ __forceinline__ void CpAsync16Streaming(void* smem_dst, const void* global_src) {
  const unsigned dst = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(global_src) : "memory");
}
__device__ __forceinline__ void CpAsyncCommit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void CpAsyncWaitAll() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ void Load4(const double* src, double (&x)[4]) {
  const double2 v0 = reinterpret_cast<const double2*>(src)[0];
  const double2 v1 = reinterpret_cast<const double2*>(src)[1];
  x[0] = v0.x, x[1] = v0.y, x[2] = v1.x, x[3] = v1.y;
}
__device__ __forceinline__ void Load16(const double* src, double (&m)[16]) {
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const double2 v = reinterpret_cast<const double2*>(src)[i];
    m[2 * i] = v.x, m[2 * i + 1] = v.y;
  }
}

// y = M x
__device__ __forceinline__ void MatVec(const double (&m)[16], const double (&x)[4],
                                       double (&y)[4]) {
#pragma unroll
  for (int i = 0; i < 4; i++)
    y[i] = fma(m[i * 4 + 3], x[3], fma(m[i * 4 + 2], x[2], fma(m[i * 4 + 1], x[1], m[i * 4] * x[0])));
}

// y = M^T x
__device__ __forceinline__ void MatTVec(const double (&m)[16], const double (&x)[4],
                                        double (&y)[4]) {
#pragma unroll
  for (int j = 0; j < 4; j++)
    y[j] = fma(m[12 + j], x[3], fma(m[8 + j], x[2], fma(m[4 + j], x[1], m[j] * x[0])));
}

__device__ __forceinline__ double Dot4(const double (&a)[4], const double (&b)[4]) {
  return fma(a[3], b[3], fma(a[2], b[2], fma(a[1], b[1], a[0] * b[0])));
}

// Per-pattern power-of-two normalisation shared by the C category lanes of a
// pattern: scales v by 2^-e with e = exponent of max over (category, state)
// and returns e.  Exact (no rounding), so rescaled and unrescaled runs agree
// bit for bit until the final log.
template <int C>
__device__ __forceinline__ int Normalize(double (&v)[4]) {
  int hi = max(max(__double2hiint(v[0]), __double2hiint(v[1])),
               max(__double2hiint(v[2]), __double2hiint(v[3])));
#pragma unroll
  for (int m = 1; m < C; m <<= 1) hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, m));
  const int biased = (hi >> 20) & 0x7ff;
  const bool ok = (biased != 0) && (biased != 0x7ff);  // zero / subnormal / inf / nan: leave as is
  const double scale = __hiloint2double(ok ? ((2046 - biased) << 20) : 0x3ff00000, 0);
#pragma unroll
  for (int i = 0; i < 4; i++) v[i] *= scale;
  return ok ? biased - 1023 : 0;
}

// Sum over the C category lanes of a pattern (lanes are adjacent).
template <int C>
__device__ __forceinline__ double SumCategories(double v) {
#pragma unroll
  for (int m = 1; m < C; m <<= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
  return v;
}

// Sum over the 32/C pattern groups of a warp (after SumCategories every lane of
// a group holds the same value, so stride-C butterflies suffice).
template <int C>
__device__ __forceinline__ double SumPatternGroups(double v) {
#pragma unroll
  for (int m = C; m < 32; m <<= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
  return v;
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
  double2* out = reinterpret_cast<double2*>(matrices + idx * kMatrixDoubles);
#pragma unroll
  for (int x = 0; x < 8; x++) out[x] = make_double2(P[2 * x], P[2 * x + 1]);
#pragma unroll
  for (int s = 0; s < 4; s++) {  // P^T: row s = column s of P
    out[8 + 2 * s] = make_double2(P[s], P[4 + s]);
    out[9 + 2 * s] = make_double2(P[8 + s], P[12 + s]);
  }
  out[16] = make_double2(1.0, 1.0);
  out[17] = make_double2(1.0, 1.0);
#pragma unroll
  for (int s = 0; s < 4; s++) {  // (Q P)^T: row s = column s of Q P
    double col[4];
#pragma unroll
    for (int i = 0; i < 4; i++)
      col[i] = fma(model.q[i * 4 + 3], P[12 + s],
                   fma(model.q[i * 4 + 2], P[8 + s], fma(model.q[i * 4 + 1], P[4 + s], model.q[i * 4] * P[s])));
    out[18 + 2 * s] = make_double2(col[0], col[1]);
    out[19 + 2 * s] = make_double2(col[2], col[3]);
  }
  out[26] = make_double2(0.0, 0.0);
  out[27] = make_double2(0.0, 0.0);
}

// ---------------------------------------------------------------------------
// Thread (pattern group g = tid / C, category c = tid % C) handles K patterns.
// Stack slot layout: [slot][j][half][tid] double2  -> conflict-free LDS.128/STS.128.

template <int C, int K, bool GRAD, bool RESCALE>
__global__ void __launch_bounds__(kThreads) TreeWalkKernel(const WalkParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  const int cat = tid % C;
  const int group = tid / C;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int n = p.taxon_count;
  const int internal_count = n - 1;
  const int edge_count = 2 * n - 2;
  const int node_count = 2 * n - 1;
  const int ops_total = GRAD ? 2 * internal_count : internal_count;
  constexpr int kTilePatterns = (kThreads / C) * K;

  // ---- shared memory carve-up ------------------------------------------------
  double2* const stack = reinterpret_cast<double2*>(smem_raw);
  int* const exps = reinterpret_cast<int*>(smem_raw + WalkStackBytes(p.slots, K));
  double* const mat_stage_warp =
      reinterpret_cast<double*>(smem_raw + WalkStackBytes(p.slots, K) + WalkExpBytes(p.slots, K, RESCALE)) +
      static_cast<size_t>(warp) * 2 * 2 * C * kStageDoubles;
  double2* const scratch_stage = reinterpret_cast<double2*>(
      smem_raw + WalkStackBytes(p.slots, K) + WalkExpBytes(p.slots, K, RESCALE) + WalkMatStageBytes(C));

  double* const q_stage =
      reinterpret_cast<double*>(smem_raw + WalkStackBytes(p.slots, K) + WalkExpBytes(p.slots, K, RESCALE) +
                                WalkMatStageBytes(C) + WalkScratchStageBytes(K, GRAD)) + warp * 16;

  auto slot_ptr = [&](int slot, int j, int half) -> double2* {
    return stack + (static_cast<size_t>(slot * K + j) * 2 + half) * kThreads + tid;
  };
  auto exp_ptr = [&](int slot, int j) -> int* { return exps + (slot * K + j) * kThreads + tid; };
  // staged matrices of (buffer, child) for this thread's category
  auto mat_ptr = [&](int buffer, int child) -> const double* {
    return mat_stage_warp + static_cast<size_t>((buffer * 2 + child) * C + cat) * kStageDoubles;
  };
  auto scratch_stage_ptr = [&](int buffer, int child, int j, int half) -> double2* {
    return scratch_stage + (static_cast<size_t>((buffer * 2 + child) * K + j) * 2 + half) * kThreads + tid;
  };
  double2* my_scratch = nullptr;
  if (GRAD)
    my_scratch = p.scratch + static_cast<size_t>(blockIdx.x) * internal_count * K * 2 * kThreads + tid;
  auto scratch_ptr = [&](int internal_index, int j, int half) -> double2* {
    return my_scratch + (static_cast<size_t>(internal_index * K + j) * 2 + half) * kThreads;
  };

  const int64_t total_items = static_cast<int64_t>(p.vtree_count) * p.chunks;
  for (int64_t item = blockIdx.x; item < total_items; item += gridDim.x) {
    const int vt = p.vtree_begin + static_cast<int>(item / p.chunks);
    const int chunk = static_cast<int>(item % p.chunks);
    const ModelTables& model = p.models[p.vtree_model[vt]];
    const double* mats = p.matrices + static_cast<size_t>(vt) * edge_count * C * kMatrixDoubles;
    const WalkOp* ops = p.ops + static_cast<size_t>(p.vtree_program[vt]) * 2 * internal_count;
    const double cat_weight = model.weights[cat];
    const double rate_w = cat_weight * model.rates[cat];    // p_c r_c
    const double drate_w = cat_weight * model.drates[cat];  // p_c dr_c/dshape
    double freqs[4];
#pragma unroll
    for (int i = 0; i < 4; i++) freqs[i] = model.freqs[i];
    const size_t out_row = (static_cast<size_t>(vt) * p.chunks + chunk) * kWarps + warp;
    double* grad_row = GRAD ? p.grad_partial + out_row * node_count : nullptr;
    double* rgrad_row = (GRAD && C > 1) ? p.rgrad_partial + out_row * node_count : nullptr;

    // Issues the asynchronous copies of one op's operands into staging `buffer`.
    auto prefetch = [&](const WalkOp& op, bool is_pre, int buffer) {
      const int flags = op.z >> 24;
#pragma unroll
      for (int child = 0; child < 2; child++) {
        const int node = child ? op.y : op.x;
        const bool leaf = flags & (child ? kBLeaf : kALeaf);
        // leaf: P^T | ones (| (QP)^T | zeros in the pre-order pass); internal: P
        const int first = leaf ? kTipBlockFirstChunk : 0;
        const int count = leaf ? (is_pre ? 20 : 10) : 8;
        const double* src = mats + static_cast<size_t>(node) * C * kMatrixDoubles;
        double* dst = mat_stage_warp + static_cast<size_t>(buffer * 2 + child) * C * kStageDoubles;
        for (int r = lane; r < C * 8; r += 32) {
          const int c = r >> 3, sub = r & 7;
          for (int ch = sub; ch < count; ch += 8)
            CpAsync16Cached(dst + c * kStageDoubles + ch * 2, src + c * kMatrixDoubles + (first + ch) * 2);
        }
        if (GRAD && is_pre && !leaf) {
#pragma unroll
          for (int j = 0; j < K; j++)
#pragma unroll
            for (int half = 0; half < 2; half++)
              CpAsync16Streaming(scratch_stage_ptr(buffer, child, j, half), scratch_ptr(node - n, j, half));
        }
      }
      CpAsyncCommit();
    };

    if (GRAD) {
      __syncwarp();
      if (lane < 16) q_stage[lane] = model.q[lane];  // this warp's copy of the rate matrix
      __syncwarp();
    }

    double logl_acc = 0.0;
    const int tile_begin = chunk * p.tiles_per_chunk;
    const int tile_end = min(tile_begin + p.tiles_per_chunk, p.tiles_total);
    for (int tile = tile_begin; tile < tile_end; tile++) {
      const int64_t pat0 =
          p.pattern_begin + static_cast<int64_t>(tile) * kTilePatterns + group * K;
      double w[K];
#pragma unroll
      for (int j = 0; j < K; j++) w[j] = (pat0 + j < p.pattern_end) ? p.weights[pat0 + j] : 0.0;
      const uint8_t* tip_base = p.tips + pat0;
      auto load_tips = [&](const WalkOp& op, int (&ta)[K], int (&tb)[K]) {
        const int flags = op.z >> 24;
#pragma unroll
        for (int j = 0; j < K; j++) {
          ta[j] = (flags & kALeaf) ? tip_base[static_cast<int64_t>(op.x) * p.tip_pitch + j] : 0;
          tb[j] = (flags & kBLeaf) ? tip_base[static_cast<int64_t>(op.y) * p.tip_pitch + j] : 0;
        }
      };

      // ---- pipeline prologue ----------------------------------------------
      WalkOp op_cur = __ldg(ops);
      WalkOp op_next = __ldg(ops + min(1, ops_total - 1));
      int tips_a[K], tips_b[K];
      load_tips(op_cur, tips_a, tips_b);
      __syncwarp();  // every lane is done with the previous tile's staging
      prefetch(op_cur, false, 0);

      for (int o = 0; o < ops_total; o++) {
        const int buffer = o & 1;
        const WalkOp op_after = __ldg(ops + min(o + 2, ops_total - 1));
        CpAsyncWaitAll();
        __syncwarp();  // staged operands of op o visible to all lanes; op o-1 fully retired
        int next_tips_a[K], next_tips_b[K];
        if (o + 1 < ops_total) {
          prefetch(op_next, GRAD && (o + 1 >= internal_count), buffer ^ 1);
          load_tips(op_next, next_tips_a, next_tips_b);
        } else {
#pragma unroll
          for (int j = 0; j < K; j++) next_tips_a[j] = next_tips_b[j] = 0;
        }

        const int a = op_cur.x, b = op_cur.y;
        const int node = op_cur.z & 0xffffff, flags = op_cur.z >> 24;
        const int s0 = op_cur.w & 0xff, s1 = (op_cur.w >> 8) & 0xff, s2 = (op_cur.w >> 16) & 0xff;
        const double* MA = mat_ptr(buffer, 0);
        const double* MB = mat_ptr(buffer, 1);

        if (!GRAD || o < internal_count) {
          // ======================= post-order op ===========================
          // dest = (P_a L_a) o (P_b L_b); s0 = dst slot, s1/s2 = child slots
          double ya[K][4], yb[K][4];
          int scale_exp[K];
#pragma unroll
          for (int j = 0; j < K; j++) scale_exp[j] = 0;
          if (flags & kALeaf) {
#pragma unroll
            for (int j = 0; j < K; j++) Load4(MA + tips_a[j] * 4, ya[j]);  // column of P (or ones)
          } else {
            double A[16];
            Load16(MA, A);
#pragma unroll
            for (int j = 0; j < K; j++) {
              const double2 v0 = *slot_ptr(s1, j, 0), v1 = *slot_ptr(s1, j, 1);
              const double x[4] = {v0.x, v0.y, v1.x, v1.y};
              MatVec(A, x, ya[j]);
              if (RESCALE) scale_exp[j] += *exp_ptr(s1, j);
            }
          }
          if (flags & kBLeaf) {
#pragma unroll
            for (int j = 0; j < K; j++) Load4(MB + tips_b[j] * 4, yb[j]);
          } else {
            double B[16];
            Load16(MB, B);
#pragma unroll
            for (int j = 0; j < K; j++) {
              const double2 v0 = *slot_ptr(s2, j, 0), v1 = *slot_ptr(s2, j, 1);
              const double x[4] = {v0.x, v0.y, v1.x, v1.y};
              MatVec(B, x, yb[j]);
              if (RESCALE) scale_exp[j] += *exp_ptr(s2, j);
            }
          }
#pragma unroll
          for (int j = 0; j < K; j++) {
            double out[4];
#pragma unroll
            for (int i = 0; i < 4; i++) out[i] = ya[j][i] * yb[j][i];
            if (RESCALE) scale_exp[j] += Normalize<C>(out);
            if (flags & kRoot) {
              // beagleCalculateRootLogLikelihoods: log sum_c p_c sum_i pi_i L[c,k,i] (+ scale)
              double site = SumCategories<C>(cat_weight * Dot4(freqs, out));
              double log_site = log(site);
              if (RESCALE) log_site = fma(static_cast<double>(scale_exp[j]), 0.6931471805599453094, log_site);
              logl_acc = fma(w[j], (cat == 0 && w[j] != 0.0) ? log_site : 0.0, logl_acc);
            } else {
              *slot_ptr(s0, j, 0) = make_double2(out[0], out[1]);
              *slot_ptr(s0, j, 1) = make_double2(out[2], out[3]);
              if (RESCALE) *exp_ptr(s0, j) = scale_exp[j];
              if (GRAD) {
                *scratch_ptr(node - n, j, 0) = make_double2(out[0], out[1]);
                *scratch_ptr(node - n, j, 1) = make_double2(out[2], out[3]);
              }
            }
          }
        } else {
          // ================ pre-order op + edge derivatives ================
          // s0 = this node's pre-order slot, s1/s2 = where the children's go.
          // With y_x = P_x L_x and t_a = pre o y_b (parent's pre-order partial times
          // the sister's contribution), the child's pre-order partial is P_a^T t_a
          // (beagleUpdatePrePartials), and because Q and P commute the per-pattern
          // derivative terms of edge a (beagleCalculateEdgeDerivatives) are
          //   numerator   = pre_a^T Q L_a = t_a . (Q y_a)
          //   denominator = pre_a^T   L_a = t_a . y_a
          // so a tip edge needs no mat-vec at all: y_a and Q y_a are columns of P and Q P.
          // Phases are ordered so that at most one 4x4 matrix is live in registers.
          double pp[K][4], ua[K][4], da[K][4], yb[K][4], db[K][4];
#pragma unroll
          for (int j = 0; j < K; j++) {
            if (flags & kRoot) {
#pragma unroll
              for (int i = 0; i < 4; i++) pp[j][i] = freqs[i];  // root pre-order partial := pi
            } else {
              const double2 v0 = *slot_ptr(s0, j, 0), v1 = *slot_ptr(s0, j, 1);
              pp[j][0] = v0.x, pp[j][1] = v0.y, pp[j][2] = v1.x, pp[j][3] = v1.y;
            }
          }
          // ---- child 0: y_a (in ua) and Q y_a (in da)
          if (flags & kALeaf) {
#pragma unroll
            for (int j = 0; j < K; j++) {
              Load4(MA + tips_a[j] * 4, ua[j]);
              Load4(MA + 20 + tips_a[j] * 4, da[j]);
            }
          } else {
            {
              double A[16];
              Load16(MA, A);
#pragma unroll
              for (int j = 0; j < K; j++) {
                const double2 v0 = *scratch_stage_ptr(buffer, 0, j, 0), v1 = *scratch_stage_ptr(buffer, 0, j, 1);
                const double x[4] = {v0.x, v0.y, v1.x, v1.y};
                MatVec(A, x, ua[j]);
              }
            }
            {
              double Q[16];
              Load16(q_stage, Q);
#pragma unroll
              for (int j = 0; j < K; j++) MatVec(Q, ua[j], da[j]);
            }
          }
          // fold the parent's pre-order partial in: ua = pre o y_a (= t_b), da = pre o (Q y_a)
#pragma unroll
          for (int j = 0; j < K; j++)
#pragma unroll
            for (int i = 0; i < 4; i++) {
              ua[j][i] *= pp[j][i];
              da[j][i] *= pp[j][i];
            }
          // ---- child 1: y_b, Q y_b, and its pre-order partial P_b^T t_b
          if (flags & kBLeaf) {
#pragma unroll
            for (int j = 0; j < K; j++) {
              Load4(MB + tips_b[j] * 4, yb[j]);
              Load4(MB + 20 + tips_b[j] * 4, db[j]);
            }
          } else {
            {
              double B[16];
              Load16(MB, B);
#pragma unroll
              for (int j = 0; j < K; j++) {
                const double2 v0 = *scratch_stage_ptr(buffer, 1, j, 0), v1 = *scratch_stage_ptr(buffer, 1, j, 1);
                const double x[4] = {v0.x, v0.y, v1.x, v1.y};
                MatVec(B, x, yb[j]);
                double pre_b[4];
                MatTVec(B, ua[j], pre_b);
                if (RESCALE) Normalize<C>(pre_b);
                *slot_ptr(s2, j, 0) = make_double2(pre_b[0], pre_b[1]);
                *slot_ptr(s2, j, 1) = make_double2(pre_b[2], pre_b[3]);
              }
            }
            {
              double Q[16];
              Load16(q_stage, Q);
#pragma unroll
              for (int j = 0; j < K; j++) MatVec(Q, yb[j], db[j]);
            }
          }
          // ---- per-pattern derivative terms; both edges share the denominator
          //      pre . (y_a o y_b) = the site likelihood seen from this node
          double ga = 0.0, gb = 0.0, ra = 0.0, rb = 0.0;
#pragma unroll
          for (int j = 0; j < K; j++) {
            const double den = Dot4(ua[j], yb[j]);
            const double num_a = Dot4(da[j], yb[j]);
            const double num_b = Dot4(ua[j], db[j]);
            const double wj = w[j];
            const bool live = wj != 0.0;  // padding patterns contribute nothing (and may be 0/0)
            const double inv = 1.0 / SumCategories<C>(cat_weight * den);
            const double term_a = SumCategories<C>(rate_w * num_a) * inv;
            const double term_b = SumCategories<C>(rate_w * num_b) * inv;
            ga = fma(wj, live ? term_a : 0.0, ga);
            gb = fma(wj, live ? term_b : 0.0, gb);
            if (C > 1) {
              const double rterm_a = SumCategories<C>(drate_w * num_a) * inv;
              const double rterm_b = SumCategories<C>(drate_w * num_b) * inv;
              ra = fma(wj, live ? rterm_a : 0.0, ra);
              rb = fma(wj, live ? rterm_b : 0.0, rb);
            }
          }
          if (!(flags & kALeaf)) {
            // child 0's pre-order partial = P_a^T (pre o y_b); P_a is re-read from staging
            double A[16];
            Load16(MA, A);
#pragma unroll
            for (int j = 0; j < K; j++) {
              double ta[4], pre_a[4];
#pragma unroll
              for (int i = 0; i < 4; i++) ta[i] = pp[j][i] * yb[j][i];
              MatTVec(A, ta, pre_a);
              if (RESCALE) Normalize<C>(pre_a);
              *slot_ptr(s1, j, 0) = make_double2(pre_a[0], pre_a[1]);
              *slot_ptr(s1, j, 1) = make_double2(pre_a[2], pre_a[3]);
            }
          }
          ga = SumPatternGroups<C>(ga);
          gb = SumPatternGroups<C>(gb);
          if (C > 1) {
            ra = SumPatternGroups<C>(ra);
            rb = SumPatternGroups<C>(rb);
          }
          if (lane == 0) {
            // single writer per (row, edge): plain read-modify-write, deterministic
            grad_row[a] += ga;
            grad_row[b] += gb;
            if (C > 1) {
              rgrad_row[a] += ra;
              rgrad_row[b] += rb;
            }
          }
        }

        op_cur = op_next;
        op_next = op_after;
#pragma unroll
        for (int j = 0; j < K; j++) {
          tips_a[j] = next_tips_a[j];
          tips_b[j] = next_tips_b[j];
        }
      }
    }
    // one partial per warp, in pattern-group order
    logl_acc = SumPatternGroups<C>(SumCategories<C>(logl_acc));
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
