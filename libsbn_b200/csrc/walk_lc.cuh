// TreeWalkLcKernel: the tree walk, with LANES = (site pattern, rate category).
//
// One pass over a tree for a tile of site patterns -- post-order partial updates,
// per-pattern power-of-two rescaling, the root log-likelihood and (gradient mode) the
// pre-order pass fused with all edge derivatives [replaces beagleUpdatePartials,
// beagleUpdatePrePartials, beagleCalculateEdgeDerivatives,
// beagleCalculateRootLogLikelihoods, beagleResetScaleFactors and
// beagleSetPartials(root pre := pi); fat_beagle.cpp:50-70, 119-175].
//
// Site patterns are independent, so a warp owns a tile of them and walks the WHOLE
// tree for it.  The walk order (host-generated, Strahler-ordered, tree_program.cpp)
// keeps the result of the previous op in registers ("cur"); only nodes with two
// internal children touch a small thread-private stack (global memory, L2 resident).
//
// Work split: the C lanes of a pattern each own one category.
//   * A thread holds K patterns x 1 category x 4 states.  (An earlier version gave a
//     thread ALL categories of its patterns and looped over them, rotating the live
//     partial through registers: the rotation alone was ~45 % of its instructions.)
//   * Each lane reads ITS category's 4x4 matrix from the shared-memory operand stage.
//     Lanes are category-major, so the 8 lanes one 128-byte wavefront serves read the
//     same category's tables: a broadcast for a matrix, rows in different banks
//     (picked by tip state) for a tip table.
//   * The only cross-lane traffic is per op: the per-pattern denominator (sum over
//     categories, a log2(C)-step butterfly), the rescaling decision (only when some
//     lane's partial has dropped well below 2^-128), the root's category sum, and one
//     transposed 4-value warp reduction of the edge-derivative sums.
//   * The evolved post-order partials y = P L go to the warp's own arena block
//     (coalesced 16 B per lane, written once) and come back through per-warp TMA bulk
//     copies that complete on the warp's own mbarriers -- a warp only ever reads what
//     it wrote itself, so no CTA-wide barrier is involved.
// Warps meet only at the mbarriers of the operand ring: TMA bulk copies of the two
// child edges' matrix blocks and the tile's tip states, requested kLcPrefetchOps ops
// ahead by whichever warp reaches an op first (a shared-memory ticket).
#ifndef SBNB_WALK_LC_CUH_
#define SBNB_WALK_LC_CUH_

#include "kernels.cuh"

namespace sbnb {

constexpr int kLcStages = 8;  // operand ring depth (one stage per op)
constexpr int kLcPrefetchOps = 4;  // ops in flight ahead of the one being computed

__host__ __device__ constexpr int LcGroup(int C) { return kThreads / C; }  // patterns per j-slab
__host__ __device__ constexpr int LcTilePatterns(int C, int K) { return LcGroup(C) * K; }
__host__ __device__ constexpr int LcTipBytes(int C, int K) { return (LcTilePatterns(C, K) + 15) / 16 * 16; }
__host__ __device__ constexpr int LcChildDoubles(int C) { return 2 * kTipTableDoubles * C; }
__host__ __device__ constexpr int LcStageBytes(int C, int K) {
  return 2 * LcChildDoubles(C) * 8 + 2 * LcTipBytes(C, K);
}
// Read-back slots of one pre-order op: [child][warp][j][half][lane] double2.
__host__ __device__ constexpr int LcScratchBytes(int K) { return 2 * K * 2 * kThreads * 16; }
// The pre-order op works through its K patterns in sub-batches of at most 2: its
// live set is ~5 partial-sized arrays per pattern, and K = 4 in one go needs 255
// registers and spills.  The post-order half of the walk handles all K at once.
__host__ __device__ constexpr int LcPreBatch(int K) { return K > 2 ? 2 : K; }
__host__ __device__ constexpr size_t LcSmemBytes(int C, int K, bool grad) {
  return static_cast<size_t>(kLcStages) * LcStageBytes(C, K) + 2 * kLcStages * 8 + kModelSmemDoubles * 8 + 80 +
         (grad ? LcScratchBytes(K) : 0);
}
// Resident CTAs per SM the register allocation is held to.  Chosen spill-free: every
// build with local-memory traffic in the op loop ran 2-3x slower.
__host__ __device__ constexpr int LcMinBlocks(int K, bool grad) {
  return grad ? (K <= 2 ? 3 : 2) : (K <= 2 ? 5 : 3);
}

// y[j] = M x[j], M row-major in shared memory (per-lane address: the lane's category)
template <int K>
__device__ __forceinline__ void MatVecLc(const double* m, const double (&x)[K][4], double (&y)[K][4]) {
#pragma unroll
  for (int i = 0; i < 4; i++) {
    double row[4];
    Load4(m + 4 * i, row);
#pragma unroll
    for (int j = 0; j < K; j++)
      y[j][i] = fma(row[3], x[j][3], fma(row[2], x[j][2], fma(row[1], x[j][1], row[0] * x[j][0])));
  }
}
// out[j] = t[j] . (M x[j])
template <int K>
__device__ __forceinline__ void MatVecDotLc(const double* m, const double (&x)[K][4], const double (&t)[K][4],
                                            double (&out)[K]) {
#pragma unroll
  for (int i = 0; i < 4; i++) {
    double row[4];
    Load4(m + 4 * i, row);
#pragma unroll
    for (int j = 0; j < K; j++) {
      const double d = fma(row[3], x[j][3], fma(row[2], x[j][2], fma(row[1], x[j][1], row[0] * x[j][0])));
      out[j] = (i == 0) ? t[j][0] * d : fma(t[j][i], d, out[j]);
    }
  }
}

// 1 / x for a normal, positive x: hardware seed + two Newton steps (relative error
// ~1e-16; x is a per-pattern likelihood kept in range by the rescaling).
__device__ __forceinline__ double FastReciprocal(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  return fma(r, e, r);
}

// Per-pattern power-of-two normalisation (see Normalize in kernels.cuh) when the C
// lanes of a pattern hold one category each: a pattern is rescaled when its largest
// entry over all categories has dropped below 2^-kLazyBits.  Finding that maximum
// takes shuffles, so the warp first votes on a cheaper, per-lane test with a lower
// bar (2^-(kLazyBits + kLaneSlackBits)): a lane whose own category lags the
// pattern's maximum by less than 2^64 does not send the warp down the slow path at
// every op (measured: 24 % of ops took it with equal bars).  Until some lane trips
// the vote a pattern may sit between the two bars unrescaled -- far from underflow.
constexpr int kLaneSlackBits = 64;
template <int C, int K>
__device__ __forceinline__ void NormalizeLc(double (&v)[K][4], int (&exps)[K]) {
  int hi[K];
  bool low = false;
#pragma unroll
  for (int j = 0; j < K; j++) {
    hi[j] = max(max(__double2hiint(v[j][0]), __double2hiint(v[j][1])),
                max(__double2hiint(v[j][2]), __double2hiint(v[j][3])));
    low = low || (hi[j] < ((1023 - kLazyBits - kLaneSlackBits) << 20));
  }
  if (!__any_sync(0xffffffffu, low)) return;
#pragma unroll
  for (int j = 0; j < K; j++) {
    int m = hi[j];
#pragma unroll
    for (int s = 32 / C; s < 32; s <<= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, s));
    const int biased = (m >> 20) & 0x7ff;
    // zero / subnormal / inf / nan, or still large enough: leave as is
    if (biased == 0 || biased >= 1023 - kLazyBits) continue;
    const double scale = __hiloint2double((2046 - biased) << 20, 0);
#pragma unroll
    for (int i = 0; i < 4; i++) v[j][i] *= scale;
    exps[j] += biased - 1023;
  }
}

template <int C, int K, bool GRAD, bool RESCALE>
__global__ void __launch_bounds__(kThreads, LcMinBlocks(K, GRAD)) TreeWalkLcKernel(const WalkParams p) {
  static_assert((C & (C - 1)) == 0 && C >= 1 && C <= 16, "lanes per pattern must be a power of two");
  static_assert(LcTilePatterns(C, K) % 16 == 0, "tiles must start on 16-byte boundaries of the tip rows");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  // Category-major lanes: the 32 / C lanes of one category are adjacent, so the lanes
  // a shared-memory load serves together (8 per 128-byte wavefront) read the SAME
  // category's tables -- a broadcast for a matrix, and rows picked by tip state
  // (different banks for A, C, G, T) for a tip table.
  constexpr int kPerWarp = 32 / C;  // patterns of one j-slab held by a warp
  const int cat = lane / kPerWarp;
  const int pidx = warp * kPerWarp + lane % kPerWarp;
  const int n = p.taxon_count;
  const int internal_count = n - 1;
  const int edge_count = 2 * n - 2;
  const int node_count = 2 * n - 1;
  const int ops_total = GRAD ? 2 * internal_count : internal_count;
  constexpr int kGroup = LcGroup(C);
  constexpr int kTilePatterns = LcTilePatterns(C, K);
  constexpr int kTipBytes = LcTipBytes(C, K);
  constexpr int kStage = LcStageBytes(C, K);
  constexpr int kChild = LcChildDoubles(C);
  constexpr int kRow = 2 * kThreads;  // double2 per j: [half][tid]

  // ---- shared memory carve-up ------------------------------------------------
  unsigned char* const ring = smem_raw;
  uint64_t* const full = reinterpret_cast<uint64_t*>(smem_raw + kLcStages * kStage);
  uint64_t* const empty = full + kLcStages;
  double* const q_smem = reinterpret_cast<double*>(empty + kLcStages);
  double* const cat_weight_smem = q_smem + 16;
  double* const rate_weight_smem = cat_weight_smem + kMaxCategories;
  double* const drate_weight_smem = rate_weight_smem + kMaxCategories;
  double* const freqs_smem = drate_weight_smem + kMaxCategories;
  // Sequence number of the next op whose operands have not been requested yet.
  uint32_t* const ticket = reinterpret_cast<uint32_t*>(freqs_smem + 4);
  // Read-back slots of the current pre-order op, [child][warp][j][half][lane] double2,
  // filled per warp by bulk copies that complete on the warp's own barrier.
  constexpr int KP = LcPreBatch(K);    // patterns per pre-order sub-batch ...
  constexpr int kBatches = K / KP;     // ... and sub-batches per op, each with its own barrier
  static_assert(K % KP == 0 && kBatches <= 2, "pre-order sub-batches must tile K");
  uint64_t* const readback_full = reinterpret_cast<uint64_t*>(freqs_smem + 6) + warp * 2;
  double2* const readback_base = reinterpret_cast<double2*>(freqs_smem + 14);
  constexpr int kWarpRows = K * 2 * 32;  // double2 a warp owns per node: [j][half][lane]
  const double2* const readback = readback_base + warp * kWarpRows + lane;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kLcStages; s++) {
      MbarInit(full + s, 1);
      MbarInit(empty + s, kWarps);
    }
    for (int w = 0; w < 2 * kWarps; w++) MbarInit(reinterpret_cast<uint64_t*>(freqs_smem + 6) + w, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // Per-CTA arenas: coalesced 16 B per lane.  Stack rows are [..][j][half][tid]
  // double2; arena rows [node][warp][j][half][lane], so that what a warp wrote for a
  // node is one contiguous block its read-back bulk copy can fetch.
  double2* const my_stack = p.stack + static_cast<size_t>(blockIdx.x) * p.slots * K * kRow + tid;
  int32_t* const my_stack_exps =
      RESCALE ? p.stack_exps + static_cast<size_t>(blockIdx.x) * p.slots * K * kThreads + tid : nullptr;
  double2* const my_scratch =
      GRAD ? p.scratch + static_cast<size_t>(blockIdx.x) * internal_count * K * kRow + warp * kWarpRows : nullptr;
  uint32_t readback_sequence = 0;  // pre-order ops this warp has consumed

  uint32_t sequence = 0;  // ops this CTA has consumed; stage = sequence % kLcStages

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
      rate_weight_smem[tid] = model.weights[tid] * model.rates[tid];    // p_c r_c
      drate_weight_smem[tid] = model.weights[tid] * model.drates[tid];  // p_c dr_c/dshape
      if (tid < 4) freqs_smem[tid] = model.freqs[tid];
    }
    const int tile_begin = chunk * p.tiles_per_chunk;
    const int tile_end = min(tile_begin + p.tiles_per_chunk, p.tiles_total);
    // The item's ops form one stream g = 0 .. item_ops - 1 (tile-major).
    const int item_ops = max(tile_end - tile_begin, 0) * ops_total;

    // Bulk copies of one op's operands into its ring stage (one thread).
    auto issue = [&](uint32_t seq, const WalkOp& op, bool is_pre, int64_t tile_pat0) {
      const int s = seq % kLcStages;
      if (seq >= kLcStages) MbarWait(empty + s, ((seq / kLcStages) - 1) & 1);
      unsigned char* stage = ring + s * kStage;
      const int flags = op.z >> 24;
      const uint32_t leaf_bytes = (is_pre ? 2 : 1) * kTipTableDoubles * C * 8;
      uint32_t bytes = 0;
#pragma unroll
      for (int child = 0; child < 2; child++) {
        const bool leaf = flags & (child ? kBLeaf : kALeaf);
        bytes += leaf ? leaf_bytes + kTipBytes : kPStride * C * 8;
      }
      MbarExpectTx(full + s, bytes);
#pragma unroll
      for (int child = 0; child < 2; child++) {
        const int node = child ? op.y : op.x;
        const bool leaf = flags & (child ? kBLeaf : kALeaf);
        const double* edge = mats + static_cast<size_t>(node) * kEdgeDoublesPerCategory * C;
        double* dst = reinterpret_cast<double*>(stage) + child * kChild;
        if (leaf) {
          BulkCopy(dst, edge + kPStride * C, leaf_bytes, full + s);
          BulkCopy(stage + 2 * kChild * 8 + child * kTipBytes,
                   p.tips + static_cast<int64_t>(node) * p.tip_pitch + tile_pat0, kTipBytes, full + s);
        } else {
          BulkCopy(dst, edge, kPStride * C * 8, full + s);
        }
      }
    };

    // Lane 0 of every warp: start the bulk copies of one sub-batch of a pre-order
    // op's internal children's evolved partials (the warp's own arena blocks) into
    // its read-back slots.
    auto fetch_readback = [&](const WalkOp& op, int batch) {
      const int flags = op.z >> 24;
      constexpr uint32_t kBlockBytes = KP * 64 * 16;
      const uint32_t bytes = ((flags & kALeaf) ? 0 : kBlockBytes) + ((flags & kBLeaf) ? 0 : kBlockBytes);
      if (bytes == 0) {
        MbarArrive(readback_full + batch);
        return;
      }
      MbarExpectTx(readback_full + batch, bytes);
#pragma unroll
      for (int child = 0; child < 2; child++) {
        if (flags & (child ? kBLeaf : kALeaf)) continue;
        BulkCopy(readback_base + (child * kWarps + warp) * kWarpRows + batch * KP * 64,
                 my_scratch + static_cast<size_t>((child ? op.y : op.x) - n) * K * kRow + batch * KP * 64,
                 kBlockBytes, readback_full + batch);
      }
    };

    // Operands are requested kLcPrefetchOps ops ahead of the FIRST warp to get to
    // an op (warps drift apart; a fixed producer warp would let the others catch up
    // with its requests and then wait out the whole copy latency at every op): the
    // warp that wins the ticket for op g + kLcPrefetchOps issues its copies.
    if (tid == 0) {
      const int first = min(kLcPrefetchOps, item_ops);
      for (int g = 0; g < first; g++) {
        const int o = g % ops_total;
        issue(sequence + g, __ldg(ops + o), GRAD && o >= internal_count,
              p.pattern_begin + static_cast<int64_t>(tile_begin + g / ops_total) * kTilePatterns);
      }
      *ticket = sequence + first;
    }
    __syncthreads();
    const double cat_weight = cat_weight_smem[cat];
    const double rate_w = rate_weight_smem[cat];
    const double drate_w = drate_weight_smem[cat];
    int ahead_o = kLcPrefetchOps % ops_total;  // op g + kLcPrefetchOps: index within its tile ...
    int ahead_tile = tile_begin + kLcPrefetchOps / ops_total;  // ... and the tile
    int g = 0;

    double logl_acc = 0.0;
    for (int tile = tile_begin; tile < tile_end; tile++) {
      const int64_t pat0 = p.pattern_begin + static_cast<int64_t>(tile) * kTilePatterns;
      double w[K];
#pragma unroll
      for (int j = 0; j < K; j++) {
        const int64_t pattern = pat0 + j * kGroup + pidx;
        w[j] = (pattern < p.pattern_end) ? p.weights[pattern] : 0.0;
      }

      WalkOp op_next = __ldg(ops);

      double cur[K][4];
      int cur_exp[K];
#pragma unroll
      for (int j = 0; j < K; j++) {
        cur_exp[j] = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) cur[j][i] = 0.0;
      }

      for (int o = 0; o < ops_total; o++, sequence++, g++) {
        if (lane == 0 && g + kLcPrefetchOps < item_ops) {
          const uint32_t target = sequence + kLcPrefetchOps;
          // (plain read first: only a warp that can win goes through the atomic)
          if (*reinterpret_cast<volatile uint32_t*>(ticket) == target &&
              atomicCAS(ticket, target, target + 1) == target)
            issue(target, __ldg(ops + ahead_o), GRAD && ahead_o >= internal_count,
                  p.pattern_begin + static_cast<int64_t>(ahead_tile) * kTilePatterns);
        }
        if (++ahead_o == ops_total) ahead_o = 0, ahead_tile++;
        __syncwarp();
        const WalkOp op = op_next;
        op_next = __ldg(ops + min(o + 1, ops_total - 1));
        const int stage_index = sequence % kLcStages;
        MbarWait(full + stage_index, (sequence / kLcStages) & 1);
        const unsigned char* stage = ring + stage_index * kStage;
        const double* MA = reinterpret_cast<const double*>(stage);
        const double* MB = MA + kChild;
        const uint8_t* tips_a = stage + 2 * kChild * 8 + pidx;
        const uint8_t* tips_b = tips_a + kTipBytes;

        const int a = op.x, b = op.y;
        const int flags = op.z >> 24;
        const int s0 = op.w & 0xff, s1 = (op.w >> 8) & 0xff, s2 = (op.w >> 16) & 0xff;
        const bool a_leaf = flags & kALeaf, b_leaf = flags & kBLeaf;

        if (!GRAD || o < internal_count) {
          // ======================= post-order op ===========================
          // cur = (P_a L_a) o (P_b L_b)
          if (flags & kStackBefore) {
#pragma unroll
            for (int j = 0; j < K; j++) {
              double2* dst = my_stack + (static_cast<size_t>(s0) * K + j) * kRow;
              dst[0] = make_double2(cur[j][0], cur[j][1]);
              dst[kThreads] = make_double2(cur[j][2], cur[j][3]);
              if (RESCALE) my_stack_exps[(s0 * K + j) * kThreads] = cur_exp[j];
            }
          }
          // At most one operand comes off the stack (the other one is cur).  The pop is
          // done inside the branch that consumes it: hoisted, its loads were predicated
          // off but still issued at the ~75 % of ops that pop nothing.
          int popped_exp[K];
#pragma unroll
          for (int j = 0; j < K; j++) popped_exp[j] = 0;
          auto pop = [&](int slot, double (&x)[K][4]) {
#pragma unroll
            for (int j = 0; j < K; j++) {
              const double2* src = my_stack + (static_cast<size_t>(slot) * K + j) * kRow;
              const double2 v0 = src[0], v1 = src[kThreads];
              x[j][0] = v0.x, x[j][1] = v0.y, x[j][2] = v1.x, x[j][3] = v1.y;
              if (RESCALE) popped_exp[j] = my_stack_exps[(slot * K + j) * kThreads];
            }
          };
          double ya[K][4], yb[K][4];
          // ---- child 0
          if (a_leaf) {
#pragma unroll
            for (int j = 0; j < K; j++)
              Load4(MA + cat * kTipTableDoubles + tips_a[j * kGroup] * 4, ya[j]);
          } else {
            if (flags & kACur) {
              MatVecLc<K>(MA + cat * kPStride, cur, ya);
            } else {
              double x[K][4];
              pop(s1, x);
              MatVecLc<K>(MA + cat * kPStride, x, ya);
            }
            if (GRAD) {
              double2* dst = my_scratch + static_cast<size_t>(a - n) * K * kRow + lane;
#pragma unroll
              for (int j = 0; j < K; j++) {
                dst[j * 64] = make_double2(ya[j][0], ya[j][1]);
                dst[j * 64 + 32] = make_double2(ya[j][2], ya[j][3]);
              }
            }
          }
          // ---- child 1
          if (b_leaf) {
#pragma unroll
            for (int j = 0; j < K; j++)
              Load4(MB + cat * kTipTableDoubles + tips_b[j * kGroup] * 4, yb[j]);
          } else {
            if (flags & kBCur) {
              MatVecLc<K>(MB + cat * kPStride, cur, yb);
            } else {
              double x[K][4];
              pop(s2, x);
              MatVecLc<K>(MB + cat * kPStride, x, yb);
            }
            if (GRAD) {
              double2* dst = my_scratch + static_cast<size_t>(b - n) * K * kRow + lane;
#pragma unroll
              for (int j = 0; j < K; j++) {
                dst[j * 64] = make_double2(yb[j][0], yb[j][1]);
                dst[j * 64 + 32] = make_double2(yb[j][2], yb[j][3]);
              }
            }
          }
          if (RESCALE) {
            const bool uses_cur = (flags & (kACur | kBCur)) != 0;
#pragma unroll
            for (int j = 0; j < K; j++) cur_exp[j] = (uses_cur ? cur_exp[j] : 0) + popped_exp[j];
          }
#pragma unroll
          for (int j = 0; j < K; j++)
#pragma unroll
            for (int i = 0; i < 4; i++) cur[j][i] = ya[j][i] * yb[j][i];
          if (RESCALE) NormalizeLc<C, K>(cur, cur_exp);
          if (flags & kRoot) {
            // beagleCalculateRootLogLikelihoods: log sum_c p_c sum_i pi_i L[c,k,i] (+ scale)
            double freqs[4];
            Load4(freqs_smem, freqs);
#pragma unroll
            for (int j = 0; j < K; j++) {
              double site = cat_weight * Dot4(freqs, cur[j]);
#pragma unroll
              for (int s = kPerWarp; s < 32; s <<= 1) site += __shfl_xor_sync(0xffffffffu, site, s);
              double log_site = log(site);
              if (RESCALE) log_site = fma(static_cast<double>(cur_exp[j]), 0.6931471805599453094, log_site);
              // one lane per pattern carries the term
              logl_acc = fma(w[j], (w[j] != 0.0 && cat == 0) ? log_site : 0.0, logl_acc);
            }
            if (GRAD) {
              // The arena was written with ordinary stores; the pre-order pass reads it
              // back through bulk copies (the async proxy): the first op's operands.
              asm volatile("fence.proxy.async;" ::: "memory");
              __syncwarp();
              if (lane == 0) {
#pragma unroll
                for (int batch = 0; batch < kBatches; batch++) fetch_readback(op_next, batch);
              }
            }
          }
        } else {
          // ================ pre-order op + edge derivatives ================
          // cur = this node's pre-order partial pp (root: pi).  With y_x = P_x L_x
          // (read back from the arena, or looked up for a tip), t_a = pp o y_b and
          // t_b = pp o y_a, the children's pre-order partials are P_a^T t_a and
          // P_b^T t_b (beagleUpdatePrePartials), and because Q and P commute the
          // per-pattern derivative terms of edge a (beagleCalculateEdgeDerivatives) are
          //   numerator   = pre_a^T Q L_a = t_a . (Q y_a)
          //   denominator = pre_a^T   L_a = t_a . y_a = pp . (y_a o y_b)   (shared by both edges)
          // so a tip edge needs no mat-vec at all: y_a and Q y_a are columns of P and Q P.
          if (flags & kRoot) {
            double freqs[4];
            Load4(freqs_smem, freqs);
#pragma unroll
            for (int j = 0; j < K; j++)
#pragma unroll
              for (int i = 0; i < 4; i++) cur[j][i] = freqs[i];
          } else if (flags & kStackBefore) {
#pragma unroll
            for (int j = 0; j < K; j++) {
              const double2* src = my_stack + (static_cast<size_t>(s0) * K + j) * kRow;
              const double2 v0 = src[0], v1 = src[kThreads];
              cur[j][0] = v0.x, cur[j][1] = v0.y, cur[j][2] = v1.x, cur[j][3] = v1.y;
            }
          }
          if (RESCALE) {
            int ignored[K];
#pragma unroll
            for (int j = 0; j < K; j++) ignored[j] = 0;
            NormalizeLc<C, K>(cur, ignored);  // the scale cancels in numerator / denominator
          }
          double ga = 0.0, gb = 0.0;
#pragma unroll
          for (int batch = 0; batch < kBatches; batch++) {
            const int j0 = batch * KP;  // this sub-batch: patterns j0 .. j0 + KP - 1
            double(&pp)[KP][4] = *reinterpret_cast<double(*)[KP][4]>(&cur[j0]);
            // ---- operands: evolved partials of both children (+ Q y for tips)
            double ya[KP][4], yb[KP][4];
            double num_a[KP], num_b[KP];
            MbarWait(readback_full + batch, readback_sequence & 1);
            if (!a_leaf) {
#pragma unroll
              for (int j = 0; j < KP; j++) {
                const double2 v0 = readback[(j0 + j) * 64], v1 = readback[(j0 + j) * 64 + 32];
                ya[j][0] = v0.x, ya[j][1] = v0.y, ya[j][2] = v1.x, ya[j][3] = v1.y;
              }
            } else {
#pragma unroll
              for (int j = 0; j < KP; j++)
                Load4(MA + cat * kTipTableDoubles + tips_a[(j0 + j) * kGroup] * 4, ya[j]);
            }
            if (!b_leaf) {
#pragma unroll
              for (int j = 0; j < KP; j++) {
                const double2 v0 = readback[kWarps * kWarpRows + (j0 + j) * 64],
                              v1 = readback[kWarps * kWarpRows + (j0 + j) * 64 + 32];
                yb[j][0] = v0.x, yb[j][1] = v0.y, yb[j][2] = v1.x, yb[j][3] = v1.y;
              }
            } else {
#pragma unroll
              for (int j = 0; j < KP; j++)
                Load4(MB + cat * kTipTableDoubles + tips_b[(j0 + j) * kGroup] * 4, yb[j]);
            }
            // this sub-batch's read-back slots are free again: start the next op's copies
            __syncwarp();
            if (lane == 0 && o + 1 < ops_total) fetch_readback(op_next, batch);

            double ta[KP][4], tb[KP][4], den[KP];
#pragma unroll
            for (int j = 0; j < KP; j++) {
#pragma unroll
              for (int i = 0; i < 4; i++) {
                ta[j][i] = pp[j][i] * yb[j][i];
                tb[j][i] = pp[j][i] * ya[j][i];
              }
              den[j] = cat_weight * Dot4(ta[j], ya[j]);
            }
            // (pp is dead from here on: the kept child's pre-order partial is written into it)
            if (a_leaf) {
#pragma unroll
              for (int j = 0; j < KP; j++) {
                double da[4];
                Load4(MA + (C + cat) * kTipTableDoubles + tips_a[(j0 + j) * kGroup] * 4, da);
                num_a[j] = Dot4(ta[j], da);
              }
            } else {
              MatVecDotLc<KP>(q_smem, ya, ta, num_a);
            }
            if (b_leaf) {
#pragma unroll
              for (int j = 0; j < KP; j++) {
                double db[4];
                Load4(MB + (C + cat) * kTipTableDoubles + tips_b[(j0 + j) * kGroup] * 4, db);
                num_b[j] = Dot4(tb[j], db);
              }
            } else {
              MatVecDotLc<KP>(q_smem, yb, tb, num_b);
            }
            // children's pre-order partials: one stays in cur, the other is pushed
            if (!a_leaf) {
              if (flags & kACur) {
                MatTVecSharedK<KP>(MA + cat * kPStride, ta, pp);
              } else {
                double pre[KP][4];
                MatTVecSharedK<KP>(MA + cat * kPStride, ta, pre);
#pragma unroll
                for (int j = 0; j < KP; j++) {
                  double2* dst = my_stack + (static_cast<size_t>(s1) * K + j0 + j) * kRow;
                  dst[0] = make_double2(pre[j][0], pre[j][1]);
                  dst[kThreads] = make_double2(pre[j][2], pre[j][3]);
                }
              }
            }
            if (!b_leaf) {
              if (flags & kBCur) {
                MatTVecSharedK<KP>(MB + cat * kPStride, tb, pp);
              } else {
                double pre[KP][4];
                MatTVecSharedK<KP>(MB + cat * kPStride, tb, pre);
#pragma unroll
                for (int j = 0; j < KP; j++) {
                  double2* dst = my_stack + (static_cast<size_t>(s2) * K + j0 + j) * kRow;
                  dst[0] = make_double2(pre[j][0], pre[j][1]);
                  dst[kThreads] = make_double2(pre[j][2], pre[j][3]);
                }
              }
            }
            // ---- per-pattern ratio
#pragma unroll
            for (int j = 0; j < KP; j++) {
              double d = den[j];
#pragma unroll
              for (int s = kPerWarp; s < 32; s <<= 1) d += __shfl_xor_sync(0xffffffffu, d, s);
              // padding patterns contribute nothing (and may be 0/0)
              const double scale = (w[j0 + j] != 0.0) ? w[j0 + j] * FastReciprocal(d) : 0.0;
              ga = fma(scale, num_a[j], ga);
              gb = fma(scale, num_b[j], gb);
            }
          }
          readback_sequence++;
          // ---- one transposed warp reduction per op
          // Lanes 0 / 8 / 16 / 24 end up with the warp sums of
          //   rate_w ga, rate_w gb, drate_w ga, drate_w gb.
          double v0 = rate_w * ga, v1 = rate_w * gb, v2 = drate_w * ga, v3 = drate_w * gb;
          {
            const bool up16 = lane & 16, up8 = lane & 8;
            const double keep0 = up16 ? v2 : v0, keep1 = up16 ? v3 : v1;
            const double send0 = up16 ? v0 : v2, send1 = up16 ? v1 : v3;
            const double r0 = keep0 + __shfl_xor_sync(0xffffffffu, send0, 16);
            const double r1 = keep1 + __shfl_xor_sync(0xffffffffu, send1, 16);
            double r = (up8 ? r1 : r0) + __shfl_xor_sync(0xffffffffu, up8 ? r0 : r1, 8);
            r += __shfl_xor_sync(0xffffffffu, r, 4);
            r += __shfl_xor_sync(0xffffffffu, r, 2);
            r += __shfl_xor_sync(0xffffffffu, r, 1);
            // Single writer per (row, edge), in program order, so the sums are
            // deterministic; a reduction (no return value) keeps the round trip to
            // L2 off the warp's critical path.
            if ((lane & 7) == 0) {
              const int edge = up8 ? b : a;
              if (!up16) {
                atomicAdd(grad_row + edge, r);
              } else if (C > 1) {
                atomicAdd(rgrad_row + edge, r);
              }
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

}  // namespace sbnb

#endif  // SBNB_WALK_LC_CUH_
