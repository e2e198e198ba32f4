// Device code of libsbn_b200 (sm_100a, fp64, no tensor cores -- 4x4 mat-vecs).
//
// Kernels:
//   TransitionMatrixKernel  P_c(t_e) = V diag(exp(lambda r_c t_e)) V^-1 per (tree, edge, category)
//                           [replaces beagleUpdateTransitionMatrices, fat_beagle.cpp:304-314]
//   TreeWalkKernel          one pass over a tree for a tile of site patterns: the
//                           post-order partial updates, per-pattern power-of-two
//                           rescaling, the root log-likelihood, and (gradient mode)
//                           the pre-order pass fused with all edge derivatives
//                           [replaces beagleUpdatePartials, beagleUpdatePrePartials,
//                            beagleCalculateEdgeDerivatives, beagleCalculateRootLogLikelihoods,
//                            beagleResetScaleFactors, beagleSetPartials(root pre := pi),
//                            beagleSetDifferentialMatrix; fat_beagle.cpp:50-70, 119-175]
//   ReducePartialsKernel    fixed-order sum of the per-(chunk, warp) partial sums
//
// Design: site patterns are independent, so a warp owns a set of
// patterns and walks the WHOLE tree for them.  The walk order (host-generated,
// Strahler-ordered, tree_program.cpp) needs only O(log n) live partials, which
// live in a per-thread-private shared-memory stack; no __syncthreads anywhere.
// In gradient mode the post-order partials are additionally streamed to a
// per-CTA scratch arena in global memory (written once, read once by the
// pre-order pass -- the only partial traffic that touches L2/HBM).
#ifndef SBNB_KERNELS_CUH_
#define SBNB_KERNELS_CUH_

#include <cuda_runtime.h>
#include <stdint.h>

#include "model.hpp"
#include "tree_program.hpp"

namespace sbnb {

constexpr int kThreads = 128;  // 4 independent warps per CTA
constexpr int kWarps = kThreads / 32;

struct WalkParams {
  // alignment (device)
  const uint8_t* tips;  // [taxon][tip_pitch], padded with gap states
  int64_t tip_pitch;
  const double* weights;  // [tip_pitch] padded with zeros
  int64_t pattern_begin, pattern_end;
  int32_t taxon_count;
  // programs (device): [program][n-1]
  const PostOp* post_ops;
  const PreOp* pre_ops;
  // virtual trees: vtree v uses program vtree_program[v], model vtree_model[v]
  int32_t vtree_begin, vtree_count;
  const int32_t* vtree_program;
  const int32_t* vtree_model;
  const ModelTables* models;
  const double* matrices;  // [vtree][2n-2][C][16]
  // tiling
  int32_t tiles_total, tiles_per_chunk, chunks;
  int32_t slots;  // shared-memory stack depth
  // scratch + outputs
  double2* scratch;       // [grid][n-1][K][2][kThreads]   (gradient mode)
  double* logl_partial;   // [vtree][chunk][warp]
  double* grad_partial;   // [vtree][chunk][warp][2n-1]     (gradient mode)
  double* rgrad_partial;  // same, with d rate_c / d shape as the scalers (C > 1)
};

// ---------------------------------------------------------------------------
// small fp64 helpers (everything is fully unrolled; matrices live in registers)

__device__ __forceinline__ void LoadMatrix(const double* __restrict__ src, double (&m)[16]) {
  const double2* s = reinterpret_cast<const double2*>(src);
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const double2 v = __ldg(s + i);
    m[2 * i] = v.x;
    m[2 * i + 1] = v.y;
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

// Compact tip: state s < 4 selects column s of M, s >= 4 (gap) contributes 1.
__device__ __forceinline__ void TipColumn(const double (&m)[16], int s, double (&y)[4]) {
#pragma unroll
  for (int i = 0; i < 4; i++) {
    double v = 1.0;
    v = (s == 0) ? m[i * 4 + 0] : v;
    v = (s == 1) ? m[i * 4 + 1] : v;
    v = (s == 2) ? m[i * 4 + 2] : v;
    v = (s == 3) ? m[i * 4 + 3] : v;
    y[i] = v;
  }
}

// Compact tip as an explicit partial: one-hot, or all ones for a gap.
__device__ __forceinline__ void TipVector(int s, double (&x)[4]) {
#pragma unroll
  for (int i = 0; i < 4; i++) x[i] = (s >= 4 || s == i) ? 1.0 : 0.0;
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
  if (biased == 0 || biased == 0x7ff) return 0;  // zero / subnormal / inf / nan: leave as is
  const double scale = __hiloint2double((2046 - biased) << 20, 0);
#pragma unroll
  for (int i = 0; i < 4; i++) v[i] *= scale;
  return biased - 1023;
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

__device__ __forceinline__ void LoadOp(const void* src, int4& lo, int4& hi) {
  const int4* s = reinterpret_cast<const int4*>(src);
  lo = __ldg(s);
  hi = __ldg(s + 1);
}

// ---------------------------------------------------------------------------

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
  double* out = matrices + idx * 16;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    double row[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      double sum = 0.0;
#pragma unroll
      for (int k = 0; k < 4; k++) sum += (model.evec[i * 4 + k] * ex[k]) * model.ivec[k * 4 + j];
      row[j] = sum > 0.0 ? sum : 0.0;  // BEAGLE clamps round-off negatives
    }
    reinterpret_cast<double2*>(out + i * 4)[0] = make_double2(row[0], row[1]);
    reinterpret_cast<double2*>(out + i * 4)[1] = make_double2(row[2], row[3]);
  }
}

// ---------------------------------------------------------------------------
// Thread (pattern group g = tid / C, category c = tid % C) handles K patterns.
// Stack slot layout: [slot][j][half][tid] double2  -> conflict-free LDS.128/STS.128.

template <int C, int K, bool GRAD, bool RESCALE>
__global__ void __launch_bounds__(kThreads) TreeWalkKernel(const WalkParams p) {
  extern __shared__ double2 smem[];
  const int tid = threadIdx.x;
  const int cat = tid % C;
  const int group = tid / C;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int n = p.taxon_count;
  const int internal_count = n - 1;
  const int edge_count = 2 * n - 2;
  const int node_count = 2 * n - 1;
  constexpr int kTilePatterns = (kThreads / C) * K;
  int* exps = reinterpret_cast<int*>(smem + static_cast<size_t>(p.slots) * K * 2 * kThreads);
  double2* my_scratch = nullptr;
  if (GRAD)
    my_scratch = p.scratch + static_cast<size_t>(blockIdx.x) * internal_count * K * 2 * kThreads + tid;

  auto slot_ptr = [&](int slot, int j, int half) -> double2* {
    return smem + (static_cast<size_t>(slot * K + j) * 2 + half) * kThreads + tid;
  };
  auto exp_ptr = [&](int slot, int j) -> int* { return exps + (slot * K + j) * kThreads + tid; };
  auto scratch_ptr = [&](int internal_index, int j, int half) -> double2* {
    return my_scratch + (static_cast<size_t>(internal_index * K + j) * 2 + half) * kThreads;
  };

  const int64_t total_items = static_cast<int64_t>(p.vtree_count) * p.chunks;
  for (int64_t item = blockIdx.x; item < total_items; item += gridDim.x) {
    const int vt = p.vtree_begin + static_cast<int>(item / p.chunks);
    const int chunk = static_cast<int>(item % p.chunks);
    const ModelTables& model = p.models[p.vtree_model[vt]];
    const double* mats = p.matrices + static_cast<size_t>(vt) * edge_count * C * 16 + cat * 16;
    const PostOp* post = p.post_ops + static_cast<size_t>(p.vtree_program[vt]) * internal_count;
    const PreOp* pre = p.pre_ops + static_cast<size_t>(p.vtree_program[vt]) * internal_count;
    const double cat_weight = model.weights[cat];
    double freqs[4];
#pragma unroll
    for (int i = 0; i < 4; i++) freqs[i] = model.freqs[i];
    const size_t out_row = (static_cast<size_t>(vt) * p.chunks + chunk) * kWarps + warp;

    double logl_acc = 0.0;
    const int tile_begin = chunk * p.tiles_per_chunk;
    const int tile_end = min(tile_begin + p.tiles_per_chunk, p.tiles_total);
    for (int tile = tile_begin; tile < tile_end; tile++) {
      const int64_t pat0 =
          p.pattern_begin + static_cast<int64_t>(tile) * kTilePatterns + group * K;
      double w[K];
#pragma unroll
      for (int j = 0; j < K; j++) w[j] = (pat0 + j < p.pattern_end) ? p.weights[pat0 + j] : 0.0;

      // ------------------------- post-order sweep ---------------------------
      for (int o = 0; o < internal_count; o++) {
        int4 lo, hi;
        LoadOp(post + o, lo, hi);
        const int node = lo.x, a = lo.y, b = lo.z, dst_slot = lo.w;
        const int a_slot = hi.x, b_slot = hi.y, flags = hi.z;
        double A[16], B[16];
        LoadMatrix(mats + static_cast<size_t>(a) * C * 16, A);
        LoadMatrix(mats + static_cast<size_t>(b) * C * 16, B);
        const uint8_t* tip_a = p.tips + static_cast<int64_t>(a) * p.tip_pitch + pat0;
        const uint8_t* tip_b = p.tips + static_cast<int64_t>(b) * p.tip_pitch + pat0;
#pragma unroll
        for (int j = 0; j < K; j++) {
          double ya[4], yb[4], out[4];
          int scale_exp = 0;
          if (flags & kALeaf) {
            TipColumn(A, tip_a[j], ya);
          } else {
            const double2 v0 = *slot_ptr(a_slot, j, 0), v1 = *slot_ptr(a_slot, j, 1);
            const double x[4] = {v0.x, v0.y, v1.x, v1.y};
            MatVec(A, x, ya);
            if (RESCALE) scale_exp += *exp_ptr(a_slot, j);
          }
          if (flags & kBLeaf) {
            TipColumn(B, tip_b[j], yb);
          } else {
            const double2 v0 = *slot_ptr(b_slot, j, 0), v1 = *slot_ptr(b_slot, j, 1);
            const double x[4] = {v0.x, v0.y, v1.x, v1.y};
            MatVec(B, x, yb);
            if (RESCALE) scale_exp += *exp_ptr(b_slot, j);
          }
#pragma unroll
          for (int i = 0; i < 4; i++) out[i] = ya[i] * yb[i];
          if (RESCALE) scale_exp += Normalize<C>(out);
          if (flags & kRoot) {
            // beagleCalculateRootLogLikelihoods: log sum_c p_c sum_i pi_i L[c,k,i] (+ scale)
            double site = cat_weight * fma(freqs[3], out[3],
                                           fma(freqs[2], out[2], fma(freqs[1], out[1], freqs[0] * out[0])));
            site = SumCategories<C>(site);
            double log_site = log(site);
            if (RESCALE) log_site = fma(static_cast<double>(scale_exp), 0.6931471805599453094, log_site);
            if (cat == 0 && w[j] != 0.0) logl_acc = fma(w[j], log_site, logl_acc);
          } else {
            *slot_ptr(dst_slot, j, 0) = make_double2(out[0], out[1]);
            *slot_ptr(dst_slot, j, 1) = make_double2(out[2], out[3]);
            if (RESCALE) *exp_ptr(dst_slot, j) = scale_exp;
            if (GRAD) {
              *scratch_ptr(node - n, j, 0) = make_double2(out[0], out[1]);
              *scratch_ptr(node - n, j, 1) = make_double2(out[2], out[3]);
            }
          }
        }
      }

      if (!GRAD) continue;

      // ------------- pre-order sweep fused with edge derivatives -------------
      double Q[16];
#pragma unroll
      for (int i = 0; i < 16; i++) Q[i] = model.q[i];
      const double rate_w = cat_weight * model.rates[cat];    // p_c r_c
      const double drate_w = cat_weight * model.drates[cat];  // p_c dr_c/dshape
      double* grad_row = p.grad_partial + out_row * node_count;
      double* rgrad_row = (C > 1) ? p.rgrad_partial + out_row * node_count : nullptr;

      for (int o = 0; o < internal_count; o++) {
        int4 lo, hi;
        LoadOp(pre + o, lo, hi);
        const int a = lo.y, b = lo.z, pre_slot = lo.w;
        const int a_dst = hi.x, b_dst = hi.y, flags = hi.z;
        double A[16], B[16];
        LoadMatrix(mats + static_cast<size_t>(a) * C * 16, A);
        LoadMatrix(mats + static_cast<size_t>(b) * C * 16, B);
        const uint8_t* tip_a = p.tips + static_cast<int64_t>(a) * p.tip_pitch + pat0;
        const uint8_t* tip_b = p.tips + static_cast<int64_t>(b) * p.tip_pitch + pat0;
        double ga = 0.0, gb = 0.0, ra = 0.0, rb = 0.0;
#pragma unroll
        for (int j = 0; j < K; j++) {
          double pp[4], la[4], lb[4], ya[4], yb[4];
          if (flags & kRoot) {
#pragma unroll
            for (int i = 0; i < 4; i++) pp[i] = freqs[i];  // root pre-order partial := pi
          } else {
            const double2 v0 = *slot_ptr(pre_slot, j, 0), v1 = *slot_ptr(pre_slot, j, 1);
            pp[0] = v0.x, pp[1] = v0.y, pp[2] = v1.x, pp[3] = v1.y;
          }
          if (flags & kALeaf) {
            TipVector(tip_a[j], la);
          } else {
            const double2 v0 = *scratch_ptr(a - n, j, 0), v1 = *scratch_ptr(a - n, j, 1);
            la[0] = v0.x, la[1] = v0.y, la[2] = v1.x, la[3] = v1.y;
          }
          if (flags & kBLeaf) {
            TipVector(tip_b[j], lb);
          } else {
            const double2 v0 = *scratch_ptr(b - n, j, 0), v1 = *scratch_ptr(b - n, j, 1);
            lb[0] = v0.x, lb[1] = v0.y, lb[2] = v1.x, lb[3] = v1.y;
          }
          MatVec(A, la, ya);
          MatVec(B, lb, yb);
          double ta[4], tb[4], pre_a[4], pre_b[4];
#pragma unroll
          for (int i = 0; i < 4; i++) {
            ta[i] = pp[i] * yb[i];  // parent pre-order x sister's contribution
            tb[i] = pp[i] * ya[i];
          }
          MatTVec(A, ta, pre_a);  // own matrix, transposed (beagleUpdatePrePartials)
          MatTVec(B, tb, pre_b);
          if (RESCALE) {
            Normalize<C>(pre_a);
            Normalize<C>(pre_b);
          }
          // beagleCalculateEdgeDerivatives: per pattern
          //   [sum_c p_c pre^T (s_c Q) post] / [sum_c p_c pre^T post]
          double qa[4], qb[4];
          MatVec(Q, la, qa);
          MatVec(Q, lb, qb);
          double num_a = fma(pre_a[3], qa[3], fma(pre_a[2], qa[2], fma(pre_a[1], qa[1], pre_a[0] * qa[0])));
          double den_a = fma(pre_a[3], la[3], fma(pre_a[2], la[2], fma(pre_a[1], la[1], pre_a[0] * la[0])));
          double num_b = fma(pre_b[3], qb[3], fma(pre_b[2], qb[2], fma(pre_b[1], qb[1], pre_b[0] * qb[0])));
          double den_b = fma(pre_b[3], lb[3], fma(pre_b[2], lb[2], fma(pre_b[1], lb[1], pre_b[0] * lb[0])));
          const double sden_a = SumCategories<C>(cat_weight * den_a);
          const double sden_b = SumCategories<C>(cat_weight * den_b);
          const double snum_a = SumCategories<C>(rate_w * num_a);
          const double snum_b = SumCategories<C>(rate_w * num_b);
          if (w[j] != 0.0) {
            ga = fma(w[j], snum_a / sden_a, ga);
            gb = fma(w[j], snum_b / sden_b, gb);
          }
          if (C > 1) {
            const double rnum_a = SumCategories<C>(drate_w * num_a);
            const double rnum_b = SumCategories<C>(drate_w * num_b);
            if (w[j] != 0.0) {
              ra = fma(w[j], rnum_a / sden_a, ra);
              rb = fma(w[j], rnum_b / sden_b, rb);
            }
          }
          if (!(flags & kALeaf)) {
            *slot_ptr(a_dst, j, 0) = make_double2(pre_a[0], pre_a[1]);
            *slot_ptr(a_dst, j, 1) = make_double2(pre_a[2], pre_a[3]);
          }
          if (!(flags & kBLeaf)) {
            *slot_ptr(b_dst, j, 0) = make_double2(pre_b[0], pre_b[1]);
            *slot_ptr(b_dst, j, 1) = make_double2(pre_b[2], pre_b[3]);
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
