// GP engine of libsbn_b200: the device-resident replacement for the reference's
// GPEngine (src/gp_engine.hpp, src/gp_engine.cpp) and the C ABI of
// include/sbn_b200_gp.h.
//
// Design.  Every GP operation except three acts on site patterns independently,
// so a thread owns a fixed set of patterns for the WHOLE program: one persistent
// kernel interprets the op stream, and all data hazards between ops (op i + 1
// reads the PLV op i wrote) are same-thread hazards that need no barrier.  The
// PLVs stay in HBM/L2 ([plv][pattern][state], the reference's memory order,
// mmapped_plv.hpp:15-41), 32 bytes per (PLV, pattern), one coalesced 32-byte
// access per lane.  The three cross-pattern couplings are reductions:
//   * Multiply's finite check, min/max scan and conditional rescale
//     (gp_engine.cpp:111-117, 288-320) -- one (max, min, flag) reduction per op
//     instead of the reference's three full passes;
//   * OptimizeBranchLength's objective (gp_engine.cpp:326-345): Brent's control
//     flow runs redundantly in every thread on identical, deterministically
//     reduced objective values, so no host round trip happens inside the search;
//   * UpdateSBNProbabilities' per-GPCSP weighted sums (gp_engine.cpp:136-153).
// Scalars (rescaling counts, branch lengths, q) are written redundantly by every
// thread with identical values, so they need no barrier either.
// One CTA (barrier = __syncthreads) serves up to 2048 patterns -- every DAG the
// reference's tests use; beyond that a cooperative grid with one grid-wide
// barrier per reduction.
//
// There is no CPU path in this file.

#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "../../include/sbn_b200_gp.h"
#include "common.hpp"
#include "device_common.cuh"
#include "model.hpp"

namespace cg = cooperative_groups;

constexpr int kGpProgramSmemWords = 8192;

namespace sbnb {

namespace {

constexpr int kGpMaxBlockThreads = 512;
constexpr int kGpSingleBlockPatterns = 2048;  // up to 4 patterns per thread in one CTA
constexpr int kGpGridBlockThreads = 256;

// Status word written by the interpreter when a reference Assert would fire.
enum GpFault : int {
  kGpOk = 0,
  kGpFaultDestRescaling = 1,     // gp_engine.cpp:70-71
  kGpFaultRescaledStationary = 2,  // gp_engine.cpp:89-90
  kGpFaultNotFinite = 3,         // gp_engine.cpp:115
  kGpFaultNegative = 4,          // gp_engine.cpp:300-301
  kGpFaultBadProgram = 5
};

struct GpParams {
  int64_t pattern_count;
  int32_t plv_count, gpcsp_count;
  double* plvs;             // [plv][pattern][4]
  int32_t* counts;          // [block][warp][plv]  rescaling counts, one identical copy per warp
  double* branch_lengths;   // [gpcsp]
  double* q;                // [gpcsp]
  const double* hybrid;     // [gpcsp]
  double* log_likelihoods;  // [gpcsp][pattern]
  double* log_marginal;     // [pattern]
  const double* weights;    // [pattern]
  const int32_t* program;
  int64_t word_count;
  double threshold, log_threshold;
  double* exchange;  // [2][blocks][4] cross-block reduction mailboxes
  int32_t* status;   // [2] = fault code, op index
  double evec[16], ivec[16], eval[4], freqs[4];
};

// gp_engine.hpp:88-98
constexpr double kMinLogBranchLength = -13.9;
constexpr double kMaxLogBranchLength = 1.1;
constexpr int kSignificantDigits = 6;
constexpr int kMaxBrentIterations = 1000;

__device__ __forceinline__ void LoadState(const double* plv, int64_t pattern, double (&x)[4]) {
  const double2* src = reinterpret_cast<const double2*>(plv + pattern * 4);
  const double2 v0 = src[0], v1 = src[1];
  x[0] = v0.x, x[1] = v0.y, x[2] = v1.x, x[3] = v1.y;
}
__device__ __forceinline__ void StoreState(double* plv, int64_t pattern, const double (&x)[4]) {
  double2* dst = reinterpret_cast<double2*>(plv + pattern * 4);
  dst[0] = make_double2(x[0], x[1]);
  dst[1] = make_double2(x[2], x[3]);
}

// P = V diag(exp(lambda t)) V^-1 (gp_engine.cpp:173-176); with `derivative`,
// V diag(lambda exp(lambda t)) V^-1 (gp_engine.cpp:178-185).
__device__ __forceinline__ void TransitionMatrix(const GpParams& p, double t, bool derivative,
                                                 double (&m)[16]) {
  double d[4];
#pragma unroll
  for (int k = 0; k < 4; k++) {
    d[k] = exp(t * p.eval[k]);
    if (derivative) d[k] *= p.eval[k];
  }
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) {
      double sum = 0.0;
#pragma unroll
      for (int k = 0; k < 4; k++) sum += (p.evec[i * 4 + k] * d[k]) * p.ivec[k * 4 + j];
      m[i * 4 + j] = sum;
    }
}

// a^T M b
__device__ __forceinline__ double Bilinear(const double (&a)[4], const double (&m)[16],
                                           const double (&b)[4]) {
  double total = 0.0;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const double row = fma(m[i * 4 + 3], b[3], fma(m[i * 4 + 2], b[2], fma(m[i * 4 + 1], b[1], m[i * 4] * b[0])));
    total = fma(a[i], row, total);
  }
  return total;
}

// numerical_utils.hpp:36-52
__device__ __forceinline__ double LogAdd(double x, double y) {
  if (y > x) {
    const double t = x;
    x = y;
    y = t;
  }
  if (x == -INFINITY) return x;
  const double neg_diff = y - x;
  if (neg_diff < -36.04365338911715 /* LOG_EPS = log(DBL_EPSILON) */) return x;
  return x + log(1.0 + exp(neg_diff));
}

// Deterministic reductions over every thread of the launch.  Each returns the
// same bits in every thread.
struct Reducer {
  const GpParams& p;
  bool multi_block;
  int round = 0;  // alternates the cross-block mailbox
  double (*smem)[4];

  // values[0..count): count <= 4 quantities reduced together; is_max[i] selects
  // max instead of sum.
  template <int COUNT>
  __device__ void All(double (&values)[COUNT], const bool (&is_max)[COUNT]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int i = 0; i < COUNT; i++) {
#pragma unroll
      for (int m = 16; m >= 1; m >>= 1) {
        const double other = __shfl_xor_sync(0xffffffffu, values[i], m);
        values[i] = is_max[i] ? fmax(values[i], other) : values[i] + other;
      }
    }
    __syncthreads();  // the previous reduction's readers are done with smem
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < COUNT; i++) smem[warp][i] = values[i];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < COUNT; i++) {
      double total = smem[0][i];
      for (int w = 1; w < warps; w++) total = is_max[i] ? fmax(total, smem[w][i]) : total + smem[w][i];
      values[i] = total;
    }
    if (multi_block) {
      double* mailbox = p.exchange + static_cast<size_t>(round & 1) * gridDim.x * 4;
      round++;
      if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < COUNT; i++) mailbox[blockIdx.x * 4 + i] = values[i];
      }
      __threadfence();
      cg::this_grid().sync();
#pragma unroll
      for (int i = 0; i < COUNT; i++) {
        double total = __ldcg(mailbox + i);
        for (unsigned b = 1; b < gridDim.x; b++) {
          const double other = __ldcg(mailbox + b * 4 + i);
          total = is_max[i] ? fmax(total, other) : total + other;
        }
        values[i] = total;
      }
    }
  }
  __device__ double Sum(double v) {
    double values[1] = {v};
    const bool is_max[1] = {false};
    All<1>(values, is_max);
    return values[0];
  }
};

__global__ void __launch_bounds__(kGpMaxBlockThreads, 1) GpInterpretKernel(const GpParams p) {
  __shared__ double reduce_smem[32][4];
  // The op program is a dependency chain: every op starts by reading its own words.  A
  // program of up to kGpProgramSmemWords words (the DS1 DAG's sweeps are ~5.5 k) is staged
  // in shared memory once, so that read is an LDS instead of a global round trip per op.
  __shared__ int32_t program_smem[kGpProgramSmemWords];
  const int32_t* program = p.program;
  if (p.word_count <= kGpProgramSmemWords) {
    for (int64_t w = threadIdx.x; w < p.word_count; w += blockDim.x) program_smem[w] = p.program[w];
    __syncthreads();
    program = program_smem;
  }
  Reducer reduce{p, gridDim.x > 1, 0, reduce_smem};
  const int64_t P = p.pattern_count;
  const int64_t first = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  // Warps drift apart between reductions, so a shared copy of the rescaling
  // counts could show a lagging warp a value from its future; every warp keeps
  // (and redundantly updates) its own copy.  Branch lengths and q are only
  // written right after a barrier of the same op, which orders them after every
  // older read.
  int32_t* counts = p.counts + (static_cast<size_t>(blockIdx.x) * (kGpMaxBlockThreads / 32) + (threadIdx.x >> 5)) *
                                   p.plv_count;
  auto plv = [&](int index) -> double* { return p.plvs + static_cast<size_t>(index) * P * 4; };
  auto fault = [&](int code, int64_t pc) {
    if (first == 0) {
      p.status[0] = code;
      p.status[1] = static_cast<int32_t>(pc);
    }
  };
  // sum_k w_k (log(rootward_k^T M leafward_k) + count_log) for OptimizeBranchLength
  auto edge_log_likelihood = [&](const double* rootward, const double* leafward, double t,
                                 double count_log) -> double {
    double m[16];
    TransitionMatrix(p, t, false, m);
    double local = 0.0;
    for (int64_t k = first; k < P; k += stride) {
      double r[4], l[4];
      LoadState(rootward, k, r);
      LoadState(leafward, k, l);
      local = fma(p.weights[k], log(Bilinear(r, m, l)) + count_log, local);
    }
    return reduce.Sum(local);
  };

  int64_t pc = 0;
  while (pc < p.word_count) {
    const int opcode = program[pc];
    switch (opcode) {
      case SBNB_GP_ZERO_PLV: {  // gp_engine.cpp:48-51
        const int dest = program[pc + 1];
        const double zero[4] = {0.0, 0.0, 0.0, 0.0};
        for (int64_t k = first; k < P; k += stride) StoreState(plv(dest), k, zero);
        counts[dest] = 0;
        pc += 2;
        break;
      }
      case SBNB_GP_SET_TO_STATIONARY: {  // gp_engine.cpp:53-62
        const int dest = program[pc + 1], root = program[pc + 2];
        const double prior = p.q[root];
        const double x[4] = {prior * p.freqs[0], prior * p.freqs[1], prior * p.freqs[2], prior * p.freqs[3]};
        for (int64_t k = first; k < P; k += stride) StoreState(plv(dest), k, x);
        counts[dest] = 0;
        pc += 3;
        break;
      }
      case SBNB_GP_INCREMENT_WITH_EVOLVED: {  // gp_engine.cpp:64-82
        const int dest = program[pc + 1], gpcsp = program[pc + 2], src = program[pc + 3];
        const int difference = counts[src] - counts[dest];
        if (difference < 0) {
          fault(kGpFaultDestRescaling, pc);
          return;
        }
        const double factor =
            (difference == 0 ? 1.0 : pow(p.threshold, static_cast<double>(difference))) * p.q[gpcsp];
        double m[16];
        TransitionMatrix(p, p.branch_lengths[gpcsp], false, m);
        for (int64_t k = first; k < P; k += stride) {
          double s[4], d[4];
          LoadState(plv(src), k, s);
          LoadState(plv(dest), k, d);
#pragma unroll
          for (int i = 0; i < 4; i++)
            d[i] += factor * fma(m[i * 4 + 3], s[3], fma(m[i * 4 + 2], s[2], fma(m[i * 4 + 1], s[1], m[i * 4] * s[0])));
          StoreState(plv(dest), k, d);
        }
        pc += 4;
        break;
      }
      case SBNB_GP_MULTIPLY: {  // gp_engine.cpp:111-117 + RescalePLVIfNeeded 298-320
        const int dest = program[pc + 1], src1 = program[pc + 2], src2 = program[pc + 3];
        int count = counts[src1] + counts[src2];
        double values[3] = {0.0, 0.0, 0.0};  // max entry, max of -entry, non-finite flag
        for (int64_t k = first; k < P; k += stride) {
          double a[4], b[4], d[4];
          LoadState(plv(src1), k, a);
          LoadState(plv(src2), k, b);
#pragma unroll
          for (int i = 0; i < 4; i++) {
            d[i] = a[i] * b[i];
            values[0] = fmax(values[0], d[i]);
            values[1] = fmax(values[1], -d[i]);
            if (!isfinite(d[i])) values[2] = 1.0;
          }
          StoreState(plv(dest), k, d);
        }
        const bool is_max[3] = {true, true, true};
        reduce.All<3>(values, is_max);
        if (values[2] != 0.0) {
          fault(kGpFaultNotFinite, pc);
          return;
        }
        if (values[1] > 0.0) {
          fault(kGpFaultNegative, pc);
          return;
        }
        double max_entry = values[0];
        if (max_entry != 0.0) {
          int rescaling = 0;
          while (max_entry < p.threshold) {
            max_entry /= p.threshold;
            rescaling++;
          }
          if (rescaling > 0) {
            const double divisor = pow(p.threshold, static_cast<double>(rescaling));
            for (int64_t k = first; k < P; k += stride) {
              double d[4];
              LoadState(plv(dest), k, d);
#pragma unroll
              for (int i = 0; i < 4; i++) d[i] /= divisor;
              StoreState(plv(dest), k, d);
            }
            count += rescaling;
          }
        }
        counts[dest] = count;
        pc += 4;
        break;
      }
      case SBNB_GP_LIKELIHOOD: {  // gp_engine.cpp:119-123, gp_engine.hpp:198-206
        const int dest = program[pc + 1], child = program[pc + 2], parent = program[pc + 3];
        double m[16];
        TransitionMatrix(p, p.branch_lengths[dest], false, m);
        const double count_log = static_cast<double>(counts[parent]) * p.log_threshold +
                                 static_cast<double>(counts[child]) * p.log_threshold;
        double* row = p.log_likelihoods + static_cast<size_t>(dest) * P;
        for (int64_t k = first; k < P; k += stride) {
          double a[4], b[4];
          LoadState(plv(parent), k, a);
          LoadState(plv(child), k, b);
          row[k] = log(Bilinear(a, m, b)) + count_log;
        }
        pc += 4;
        break;
      }
      case SBNB_GP_OPTIMIZE_BRANCH_LENGTH: {  // gp_engine.cpp:326-345, optimization.hpp:10-115
        const int leafward = program[pc + 1], rootward = program[pc + 2], gpcsp = program[pc + 3];
        const double count_log = static_cast<double>(counts[rootward]) * p.log_threshold +
                                 static_cast<double>(counts[leafward]) * p.log_threshold;
        auto f = [&](double log_branch_length) -> double {
          return -edge_log_likelihood(plv(rootward), plv(leafward), exp(log_branch_length), count_log);
        };
        const double current_log_branch_length = log(p.branch_lengths[gpcsp]);
        const double current_value = f(current_log_branch_length);
        // ---- Brent's minimiser, transcribed control flow (boost/math/tools/minima.hpp
        //      as copied into optimization.hpp)
        double min = kMinLogBranchLength, max = kMaxLogBranchLength;
        const double tolerance = ldexp(1.0, 1 - kSignificantDigits);
        double x, w, v, u, delta, delta2, fu, fv, fw, fx, mid, fract1, fract2;
        const double golden = 0.3819660f;
        x = w = v = max;
        fw = fv = fx = f(x);
        delta2 = delta = 0;
        int count = kMaxBrentIterations;
        do {
          mid = (min + max) / 2;
          fract1 = tolerance * fabs(x) + tolerance / 4;
          fract2 = 2 * fract1;
          if (fabs(x - mid) <= (fract2 - (max - min) / 2)) break;
          if (fabs(delta2) > fract1) {
            double r = (x - w) * (fx - fv);
            double qq = (x - v) * (fx - fw);
            double pp = (x - v) * qq - (x - w) * r;
            qq = 2 * (qq - r);
            if (qq > 0) pp = -pp;
            qq = fabs(qq);
            const double td = delta2;
            delta2 = delta;
            if ((fabs(pp) >= fabs(qq * td / 2)) || (pp <= qq * (min - x)) || (pp >= qq * (max - x))) {
              delta2 = (x >= mid) ? min - x : max - x;
              delta = golden * delta2;
            } else {
              delta = pp / qq;
              u = x + delta;
              if (((u - min) < fract2) || ((max - u) < fract2))
                delta = (mid - x) < 0 ? -fabs(fract1) : fabs(fract1);
            }
          } else {
            delta2 = (x >= mid) ? min - x : max - x;
            delta = golden * delta2;
          }
          u = (fabs(delta) >= fract1) ? (x + delta) : (delta > 0 ? (x + fabs(fract1)) : (x - fabs(fract1)));
          fu = f(u);
          if (fu <= fx) {
            if (u >= x)
              min = x;
            else
              max = x;
            v = w;
            w = x;
            x = u;
            fv = fw;
            fw = fx;
            fx = fu;
          } else {
            if (u < x)
              min = u;
            else
              max = u;
            if ((fu <= fw) || (w == x)) {
              v = w;
              w = u;
              fv = fw;
              fw = fu;
            } else if ((fu <= fv) || (v == x) || (v == w)) {
              v = u;
              fv = fu;
            }
          }
        } while (--count);
        // "Numerical optimization sometimes yields new nllk > current nllk."
        p.branch_lengths[gpcsp] = (fx > current_value) ? exp(current_log_branch_length) : exp(x);
        pc += 4;
        break;
      }
      case SBNB_GP_UPDATE_SBN_PROBABILITIES: {  // gp_engine.cpp:136-153
        const int start = program[pc + 1], stop = program[pc + 2];
        const int length = stop - start;
        if (length == 1) {
          __syncthreads();  // lagging warps may still be reading q in an older op
          if (gridDim.x > 1) cg::this_grid().sync();
          p.q[start] = 1.0;
        } else if (length > 1) {
          bool use_hybrid = true;
          for (int g = start; g < stop; g++) use_hybrid = use_hybrid && (p.hybrid[g] > -INFINITY);
          // log of the unnormalised posterior per GPCSP, folded with LogAdd in index order
          double log_norm = 0.0;
          for (int g = start; g < stop; g++) {
            double log_likelihood;
            if (use_hybrid) {
              log_likelihood = p.hybrid[g];
            } else {
              const double* row = p.log_likelihoods + static_cast<size_t>(g) * P;
              double local = 0.0;
              for (int64_t k = first; k < P; k += stride) local = fma(row[k], p.weights[k], local);
              log_likelihood = reduce.Sum(local);
            }
            const double value = log_likelihood + log(p.q[g]);
            log_norm = (g == start) ? value : LogAdd(log_norm, value);
          }
          // second pass recomputes the same values (identical bits) and normalises
          for (int g = start; g < stop; g++) {
            double log_likelihood;
            if (use_hybrid) {
              log_likelihood = p.hybrid[g];
            } else {
              const double* row = p.log_likelihoods + static_cast<size_t>(g) * P;
              double local = 0.0;
              for (int64_t k = first; k < P; k += stride) local = fma(row[k], p.weights[k], local);
              log_likelihood = reduce.Sum(local);
            }
            const double updated = exp(log_likelihood + log(p.q[g]) - log_norm);
            // every thread must have read q[g] before anyone overwrites it
            __syncthreads();
            if (gridDim.x > 1) cg::this_grid().sync();
            p.q[g] = updated;
          }
        }
        pc += 3;
        break;
      }
      case SBNB_GP_RESET_MARGINAL_LIKELIHOOD: {  // gp_engine.cpp:84-86
        for (int64_t k = first; k < P; k += stride) p.log_marginal[k] = -INFINITY;
        pc += 1;
        break;
      }
      case SBNB_GP_INCREMENT_MARGINAL: {  // gp_engine.cpp:88-109
        const int stationary = program[pc + 1], rootsplit = program[pc + 2], leafward = program[pc + 3];
        if (counts[stationary] != 0) {
          fault(kGpFaultRescaledStationary, pc);
          return;
        }
        const double count_log = static_cast<double>(counts[leafward]) * p.log_threshold;
        const double log_prior = log(p.q[rootsplit]);
        double* row = p.log_likelihoods + static_cast<size_t>(rootsplit) * P;
        for (int64_t k = first; k < P; k += stride) {
          double a[4], b[4];
          LoadState(plv(stationary), k, a);
          LoadState(plv(leafward), k, b);
          const double value = log(fma(a[3], b[3], fma(a[2], b[2], fma(a[1], b[1], a[0] * b[0])))) + count_log;
          p.log_marginal[k] = LogAdd(p.log_marginal[k], value);
          row[k] = value - log_prior;
        }
        pc += 4;
        break;
      }
      case SBNB_GP_PREP_FOR_MARGINALIZATION: {  // gp_engine.cpp:155-165
        const int dest = program[pc + 1], src_count = program[pc + 2];
        int minimum = counts[program[pc + 3]];
        for (int i = 1; i < src_count; i++) minimum = min(minimum, counts[program[pc + 3 + i]]);
        counts[dest] = minimum;
        pc += 3 + src_count;
        break;
      }
      default:
        fault(kGpFaultBadProgram, pc);
        return;
    }
  }
}

// GetLogMarginalLikelihood, GetPerGPCSPLogLikelihoods: rows[r] . weights, one block per row.
__global__ void GpWeightedRowSumsKernel(const double* rows, const double* weights, int64_t P, double* out) {
  __shared__ double smem[32];
  const double* row = rows + static_cast<size_t>(blockIdx.x) * P;
  double local = 0.0;
  for (int64_t k = threadIdx.x; k < P; k += blockDim.x) local = fma(row[k], weights[k], local);
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) local += __shfl_xor_sync(0xffffffffu, local, m);
  if ((threadIdx.x & 31) == 0) smem[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    double total = 0.0;
    for (int w = 0; w < (blockDim.x + 31) / 32; w++) total += smem[w];
    out[blockIdx.x] = total;
  }
}

// LogLikelihoodAndDerivative (gp_engine.cpp:244-266), one block.
__global__ void GpLogLikelihoodAndDerivativeKernel(const GpParams p, int leafward, int rootward, int gpcsp,
                                                   double* out) {
  __shared__ double smem[32][2];
  const int64_t P = p.pattern_count;
  double m[16], dm[16];
  const double t = p.branch_lengths[gpcsp];
  TransitionMatrix(p, t, false, m);
  TransitionMatrix(p, t, true, dm);
  const double count_log = static_cast<double>(p.counts[rootward]) * p.log_threshold +
                           static_cast<double>(p.counts[leafward]) * p.log_threshold;
  double log_likelihood = 0.0, derivative = 0.0;
  for (int64_t k = threadIdx.x; k < P; k += blockDim.x) {
    double r[4], l[4];
    LoadState(p.plvs + static_cast<size_t>(rootward) * P * 4, k, r);
    LoadState(p.plvs + static_cast<size_t>(leafward) * P * 4, k, l);
    const double likelihood = Bilinear(r, m, l);
    log_likelihood = fma(p.weights[k], log(likelihood) + count_log, log_likelihood);
    derivative = fma(p.weights[k], Bilinear(r, dm, l) / likelihood, derivative);
  }
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    log_likelihood += __shfl_xor_sync(0xffffffffu, log_likelihood, s);
    derivative += __shfl_xor_sync(0xffffffffu, derivative, s);
  }
  if ((threadIdx.x & 31) == 0) smem[threadIdx.x >> 5][0] = log_likelihood, smem[threadIdx.x >> 5][1] = derivative;
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int w = 0; w < (blockDim.x + 31) / 32; w++) a += smem[w][0], b += smem[w][1];
    out[0] = a;
    out[1] = b;
  }
}

__global__ void GpTransitionMatrixKernel(const GpParams p, double branch_length, double* out) {
  double m[16];
  TransitionMatrix(p, branch_length, false, m);
  if (threadIdx.x == 0)
    for (int i = 0; i < 16; i++) out[i] = m[i];
}

// CalculateQuartetHybridLikelihoods (gp_engine.cpp:396-452): one block per
// (rootward, sister, rotated, sorted) combination; every combination recomputes
// the short chain of 4 x 4 products for its patterns instead of materialising
// the reference's four scratch PLVs.
struct GpQuartetParams {
  const int32_t* tips;  // records of 3 int32: rootward, then sister, rotated, sorted lists
  int32_t rootward_count, sister_count, rotated_count, sorted_count;
  int32_t central_gpcsp;
  const double* unconditional_node_probabilities;
  const double* inverted_sbn_prior;
};
__global__ void GpQuartetKernel(const GpParams p, const GpQuartetParams qp, double* out, int32_t* status) {
  __shared__ double smem[32];
  const int64_t P = p.pattern_count;
  int index = blockIdx.x;
  const int sorted_i = index % qp.sorted_count;
  index /= qp.sorted_count;
  const int rotated_i = index % qp.rotated_count;
  index /= qp.rotated_count;
  const int sister_i = index % qp.sister_count;
  const int rootward_i = index / qp.sister_count;
  const int32_t* rootward = qp.tips + 3 * rootward_i;
  const int32_t* sister = qp.tips + 3 * (qp.rootward_count + sister_i);
  const int32_t* rotated = qp.tips + 3 * (qp.rootward_count + qp.sister_count + rotated_i);
  const int32_t* sorted = qp.tips + 3 * (qp.rootward_count + qp.sister_count + qp.rotated_count + sorted_i);
  if (p.counts[rootward[1]] != 0 || p.counts[sister[1]] != 0 || p.counts[rotated[1]] != 0 ||
      p.counts[sorted[1]] != 0) {
    if (threadIdx.x == 0) status[0] = 1;  // "Rescaling not implemented in CalculateQuartetHybridLikelihoods."
    return;
  }
  double m_rootward[16], m_sister[16], m_central[16], m_rotated[16], m_sorted[16];
  TransitionMatrix(p, p.branch_lengths[rootward[2]], false, m_rootward);
  TransitionMatrix(p, p.branch_lengths[sister[2]], false, m_sister);
  TransitionMatrix(p, p.branch_lengths[qp.central_gpcsp], false, m_central);
  TransitionMatrix(p, p.branch_lengths[rotated[2]], false, m_rotated);
  TransitionMatrix(p, p.branch_lengths[sorted[2]], false, m_sorted);
  auto apply = [](const double (&m)[16], const double (&x)[4], double (&y)[4]) {
#pragma unroll
    for (int i = 0; i < 4; i++)
      y[i] = fma(m[i * 4 + 3], x[3], fma(m[i * 4 + 2], x[2], fma(m[i * 4 + 1], x[1], m[i * 4] * x[0])));
  };
  const double log_rootward_tip_prior = log(qp.unconditional_node_probabilities[rootward[0]]);
  double local = 0.0;
  for (int64_t k = threadIdx.x; k < P; k += blockDim.x) {
    double x[4], root_plv[4], y[4], r_s[4], q_s[4], r_sorted[4];
    LoadState(p.plvs + static_cast<size_t>(rootward[1]) * P * 4, k, x);
    apply(m_rootward, x, root_plv);
    LoadState(p.plvs + static_cast<size_t>(sister[1]) * P * 4, k, x);
    apply(m_sister, x, y);
#pragma unroll
    for (int i = 0; i < 4; i++) r_s[i] = root_plv[i] * y[i];
    apply(m_central, r_s, q_s);
    LoadState(p.plvs + static_cast<size_t>(rotated[1]) * P * 4, k, x);
    apply(m_rotated, x, y);
#pragma unroll
    for (int i = 0; i < 4; i++) r_sorted[i] = q_s[i] * y[i];
    LoadState(p.plvs + static_cast<size_t>(sorted[1]) * P * 4, k, x);
    local = fma(p.weights[k], log(Bilinear(r_sorted, m_sorted, x)) - log_rootward_tip_prior, local);
  }
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) local += __shfl_xor_sync(0xffffffffu, local, s);
  if ((threadIdx.x & 31) == 0) smem[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    double total = 0.0;
    for (int w = 0; w < (blockDim.x + 31) / 32; w++) total += smem[w];
    const double non_sequence = log(qp.inverted_sbn_prior[rootward[2]] * p.q[sister[2]] * p.q[rotated[2]] *
                                    p.q[sorted[2]]);
    out[blockIdx.x] = non_sequence + total;
  }
}

const char* FaultMessage(int code) {
  switch (code) {
    case kGpFaultDestRescaling:
      return "dest_ rescaling too large in IncrementWithWeightedEvolvedPLV";
    case kGpFaultRescaledStationary:
      return "Surprise! Rescaled stationary distribution in IncrementMarginalLikelihood";
    case kGpFaultNotFinite:
      return "Multiply dest_ is not finite";
    case kGpFaultNegative:
      return "PLV with negative entry passed to RescalePLVIfNeeded";
    default:
      return "Malformed GP operation program";
  }
}

}  // namespace

}  // namespace sbnb

using namespace sbnb;

struct sbnb_gp_engine {
  int device = 0;
  int sm_count = 0;
  int32_t taxon_count = 0, plv_count = 0, gpcsp_count = 0, node_count = 0;
  int64_t pattern_count = 0, site_count = 0;
  int blocks = 1, threads = 32;
  cudaStream_t stream = nullptr;
  cudaEvent_t begin = nullptr, end = nullptr;
  GpParams params{};
  DeviceArray<double> plvs, branch_lengths, q, hybrid, log_likelihoods, log_marginal, weights, exchange, scalars,
      node_probabilities, inverted_prior;
  DeviceArray<int32_t> counts, program, status, tips;
  bool has_node_probabilities = false;
  int64_t launch_count = 0;
  double last_kernel_ms = 0.0;

  ~sbnb_gp_engine() {
    if (begin) cudaEventDestroy(begin);
    if (end) cudaEventDestroy(end);
    if (stream) cudaStreamDestroy(stream);
  }
};

namespace {

void Bind(sbnb_gp_engine* e) { SBNB_CUDA(cudaSetDevice(e->device)); }

void CopyOut(sbnb_gp_engine* e, double* host, const double* device, size_t count) {
  SBNB_CUDA(cudaMemcpyAsync(host, device, count * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
  SBNB_CUDA(cudaStreamSynchronize(e->stream));
}

void CheckPlv(const sbnb_gp_engine* e, int index) {
  Require(index >= 0 && index < e->plv_count, "PLV index out of range.");
}
void CheckGpcsp(const sbnb_gp_engine* e, int index) {
  Require(index >= 0 && index < e->gpcsp_count, "GPCSP index out of range.");
}

// Host-side validation of a program: every index in range, records complete.
void ValidateProgram(const sbnb_gp_engine* e, const int32_t* program, int64_t words) {
  int64_t pc = 0;
  while (pc < words) {
    const int opcode = program[pc];
    auto need = [&](int64_t count) { Require(pc + count <= words, "Truncated GP operation record."); };
    switch (opcode) {
      case SBNB_GP_ZERO_PLV:
        need(2);
        CheckPlv(e, program[pc + 1]);
        pc += 2;
        break;
      case SBNB_GP_SET_TO_STATIONARY:
        need(3);
        CheckPlv(e, program[pc + 1]);
        CheckGpcsp(e, program[pc + 2]);
        pc += 3;
        break;
      case SBNB_GP_INCREMENT_WITH_EVOLVED:
        need(4);
        CheckPlv(e, program[pc + 1]);
        CheckGpcsp(e, program[pc + 2]);
        CheckPlv(e, program[pc + 3]);
        pc += 4;
        break;
      case SBNB_GP_MULTIPLY:
        need(4);
        for (int i = 1; i <= 3; i++) CheckPlv(e, program[pc + i]);
        pc += 4;
        break;
      case SBNB_GP_LIKELIHOOD:
        need(4);
        CheckGpcsp(e, program[pc + 1]);
        CheckPlv(e, program[pc + 2]);
        CheckPlv(e, program[pc + 3]);
        pc += 4;
        break;
      case SBNB_GP_OPTIMIZE_BRANCH_LENGTH:
        need(4);
        CheckPlv(e, program[pc + 1]);
        CheckPlv(e, program[pc + 2]);
        CheckGpcsp(e, program[pc + 3]);
        pc += 4;
        break;
      case SBNB_GP_UPDATE_SBN_PROBABILITIES:
        need(3);
        Require(program[pc + 1] >= 0 && program[pc + 1] <= program[pc + 2] && program[pc + 2] <= e->gpcsp_count,
                "UpdateSBNProbabilities range out of bounds.");
        pc += 3;
        break;
      case SBNB_GP_RESET_MARGINAL_LIKELIHOOD:
        pc += 1;
        break;
      case SBNB_GP_INCREMENT_MARGINAL:
        need(4);
        CheckPlv(e, program[pc + 1]);
        CheckGpcsp(e, program[pc + 2]);
        CheckPlv(e, program[pc + 3]);
        pc += 4;
        break;
      case SBNB_GP_PREP_FOR_MARGINALIZATION: {
        need(3);
        CheckPlv(e, program[pc + 1]);
        const int count = program[pc + 2];
        Require(count > 0, "Empty src_vector in PrepForMarginalization");
        need(3 + count);
        for (int i = 0; i < count; i++) CheckPlv(e, program[pc + 3 + i]);
        pc += 3 + count;
        break;
      }
      default:
        Fail(SBNB_ERR_INVALID_ARGUMENT, "Unknown GP opcode " + std::to_string(opcode) + ".");
    }
  }
}

void WeightedRowSums(sbnb_gp_engine* e, const double* rows, int row_count, double* host_out) {
  if (row_count == 0) return;
  e->scalars.Reserve(std::max<size_t>(row_count, 64));
  GpWeightedRowSumsKernel<<<row_count, 256, 0, e->stream>>>(rows, e->weights.get(), e->pattern_count,
                                                            e->scalars.get());
  SBNB_CUDA(cudaGetLastError());
  e->launch_count++;
  CopyOut(e, host_out, e->scalars.get(), row_count);
}

std::vector<int32_t> PackTips(const int32_t* a, int na, const int32_t* b, int nb, const int32_t* c, int nc,
                              const int32_t* d, int nd) {
  std::vector<int32_t> out;
  auto append = [&](const int32_t* tips, int count) { out.insert(out.end(), tips, tips + 3 * count); };
  append(a, na), append(b, nb), append(c, nc), append(d, nd);
  return out;
}

void QuartetLikelihoods(sbnb_gp_engine* e, int32_t central, const int32_t* rootward, int32_t rootward_count,
                        const int32_t* sister, int32_t sister_count, const int32_t* rotated, int32_t rotated_count,
                        const int32_t* sorted, int32_t sorted_count, std::vector<double>* out) {
  Bind(e);
  CheckGpcsp(e, central);
  Require(e->has_node_probabilities,
          "Quartet hybrid likelihoods need unconditional_node_probabilities and inverted_sbn_prior.");
  Require(rootward_count >= 0 && sister_count >= 0 && rotated_count >= 0 && sorted_count >= 0,
          "Negative tip count.");
  const int64_t total = static_cast<int64_t>(rootward_count) * sister_count * rotated_count * sorted_count;
  out->assign(total, 0.0);
  if (total == 0) return;
  Require((rootward && sister && rotated && sorted), "NULL tip list.");
  const std::vector<int32_t> tips =
      PackTips(rootward, rootward_count, sister, sister_count, rotated, rotated_count, sorted, sorted_count);
  for (size_t i = 0; i < tips.size(); i += 3) {
    Require(tips[i] >= 0 && tips[i] < e->node_count, "Quartet tip node id out of range.");
    CheckPlv(e, tips[i + 1]);
    CheckGpcsp(e, tips[i + 2]);
  }
  e->tips.Upload(tips.data(), tips.size(), e->stream);
  e->scalars.Reserve(std::max<size_t>(total, 64));
  SBNB_CUDA(cudaMemsetAsync(e->status.get(), 0, 2 * sizeof(int32_t), e->stream));
  GpQuartetParams qp{e->tips.get(),  rootward_count, sister_count, rotated_count, sorted_count, central,
                     e->node_probabilities.get(), e->inverted_prior.get()};
  GpQuartetKernel<<<static_cast<unsigned>(total), 256, 0, e->stream>>>(e->params, qp, e->scalars.get(),
                                                                      e->status.get());
  SBNB_CUDA(cudaGetLastError());
  e->launch_count++;
  int32_t status[2] = {0, 0};
  SBNB_CUDA(cudaMemcpyAsync(status, e->status.get(), sizeof(status), cudaMemcpyDeviceToHost, e->stream));
  CopyOut(e, out->data(), e->scalars.get(), total);
  if (status[0] != 0)
    Fail(SBNB_ERR_GP_ASSERT, "Rescaling not implemented in CalculateQuartetHybridLikelihoods.");
}

}  // namespace

extern "C" {

int sbnb_gp_create(int32_t taxon_count, int64_t pattern_count, const uint8_t* tip_states,
                   const double* pattern_weights, int64_t site_count, int32_t plv_count, int32_t gpcsp_count,
                   double rescaling_threshold, const double* sbn_prior,
                   const double* unconditional_node_probabilities, int32_t node_count,
                   const double* inverted_sbn_prior, int32_t device, sbnb_gp_engine** out) {
  return Guard([&] {
    Require(out != nullptr, "NULL output handle.");
    *out = nullptr;
    Require(taxon_count >= 1 && pattern_count >= 1, "Need at least 1 taxon and 1 site pattern.");
    Require(tip_states && pattern_weights, "NULL tip_states / pattern_weights.");
    Require(plv_count >= taxon_count, "plv_count must cover the taxa (6 PLVs per DAG node).");
    Require(gpcsp_count >= 0, "Negative GPCSP count.");
    Require(rescaling_threshold > 0.0 && rescaling_threshold < 1.0, "rescaling_threshold must be in (0, 1).");
    int device_count = 0;
    if (cudaGetDeviceCount(&device_count) != cudaSuccess || device_count < 1) {
      cudaGetLastError();
      Fail(SBNB_ERR_NO_DEVICE, "No CUDA device available: libsbn_b200 has no CPU fallback (cudaGetDeviceCount).");
    }
    Require(device >= 0 && device < device_count, "CUDA device ordinal out of range.");
    auto e = std::make_unique<sbnb_gp_engine>();
    e->device = device;
    Bind(e.get());
    cudaDeviceProp prop{};
    SBNB_CUDA(cudaGetDeviceProperties(&prop, device));
    e->sm_count = prop.multiProcessorCount;
    e->taxon_count = taxon_count;
    e->pattern_count = pattern_count;
    e->site_count = site_count;
    e->plv_count = plv_count;
    e->gpcsp_count = gpcsp_count;
    e->node_count = node_count;
    SBNB_CUDA(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
    SBNB_CUDA(cudaEventCreate(&e->begin));
    SBNB_CUDA(cudaEventCreate(&e->end));
    const int64_t P = pattern_count;
    if (P <= kGpSingleBlockPatterns) {
      e->blocks = 1;
      e->threads = static_cast<int>(std::min<int64_t>((P + 31) / 32 * 32, kGpMaxBlockThreads));
    } else {
      e->threads = kGpGridBlockThreads;
      int per_sm = 0;
      SBNB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, GpInterpretKernel, e->threads, 0));
      const int64_t resident = static_cast<int64_t>(std::max(per_sm, 1)) * e->sm_count;
      e->blocks = static_cast<int>(std::min<int64_t>((P + e->threads - 1) / e->threads, resident));
    }
    const size_t gpcsps = std::max(gpcsp_count, 1);
    // PLVs: zero, then one-hot tips / all-ones gaps (gp_engine.cpp:268-286).
    {
      const size_t plv_doubles = static_cast<size_t>(plv_count) * P * 4;
      e->plvs.Reserve(plv_doubles);
      SBNB_CUDA(cudaMemsetAsync(e->plvs.get(), 0, plv_doubles * sizeof(double), e->stream));
      std::vector<double> tips(static_cast<size_t>(taxon_count) * P * 4, 0.0);
      for (int taxon = 0; taxon < taxon_count; taxon++)
        for (int64_t k = 0; k < P; k++) {
          const uint8_t symbol = tip_states[static_cast<size_t>(taxon) * P + k];
          double* x = tips.data() + (static_cast<size_t>(taxon) * P + k) * 4;
          if (symbol == 4) {
            x[0] = x[1] = x[2] = x[3] = 1.0;
          } else if (symbol < 4) {
            x[symbol] = 1.0;
          }  // symbols > 4 leave the column zero, as the reference does
        }
      SBNB_CUDA(cudaMemcpyAsync(e->plvs.get(), tips.data(), tips.size() * sizeof(double), cudaMemcpyHostToDevice,
                                e->stream));
      SBNB_CUDA(cudaStreamSynchronize(e->stream));
    }
    e->weights.Upload(pattern_weights, P, e->stream);
    std::vector<double> branch_lengths(gpcsps, 0.1);  // default_branch_length_, gp_engine.hpp:84
    e->branch_lengths.Upload(branch_lengths.data(), gpcsps, e->stream);
    std::vector<double> q(gpcsps, 0.0);
    if (sbn_prior) std::copy(sbn_prior, sbn_prior + gpcsp_count, q.begin());
    e->q.Upload(q.data(), gpcsps, e->stream);
    std::vector<double> minus_infinity(std::max<size_t>(gpcsps, P), -INFINITY);
    e->hybrid.Upload(minus_infinity.data(), gpcsps, e->stream);
    e->log_marginal.Upload(minus_infinity.data(), P, e->stream);
    e->log_likelihoods.Reserve(gpcsps * P);
    SBNB_CUDA(cudaMemsetAsync(e->log_likelihoods.get(), 0, gpcsps * P * sizeof(double), e->stream));
    const size_t count_copies = static_cast<size_t>(e->blocks) * (kGpMaxBlockThreads / 32);
    e->counts.Reserve(count_copies * plv_count);
    SBNB_CUDA(cudaMemsetAsync(e->counts.get(), 0, count_copies * plv_count * sizeof(int32_t), e->stream));
    e->exchange.Reserve(static_cast<size_t>(2) * e->blocks * 4);
    e->status.Reserve(2);
    if (unconditional_node_probabilities && inverted_sbn_prior && node_count > 0) {
      e->node_probabilities.Upload(unconditional_node_probabilities, node_count, e->stream);
      e->inverted_prior.Upload(inverted_sbn_prior, gpcsps == static_cast<size_t>(gpcsp_count) ? gpcsp_count : 0,
                               e->stream);
      e->has_node_probabilities = true;
    }
    SBNB_CUDA(cudaStreamSynchronize(e->stream));

    GpParams& p = e->params;
    p.pattern_count = P;
    p.plv_count = plv_count;
    p.gpcsp_count = gpcsp_count;
    p.plvs = e->plvs.get();
    p.counts = e->counts.get();
    p.branch_lengths = e->branch_lengths.get();
    p.q = e->q.get();
    p.hybrid = e->hybrid.get();
    p.log_likelihoods = e->log_likelihoods.get();
    p.log_marginal = e->log_marginal.get();
    p.weights = e->weights.get();
    p.threshold = rescaling_threshold;
    p.log_threshold = std::log(rescaling_threshold);
    p.exchange = e->exchange.get();
    p.status = e->status.get();
    ModelTables tables;
    BuildModelTables(ModelSpec::Parse("JC69", "constant", "none"), nullptr, &tables);
    std::copy(tables.evec, tables.evec + 16, p.evec);
    std::copy(tables.ivec, tables.ivec + 16, p.ivec);
    std::copy(tables.eval, tables.eval + 4, p.eval);
    std::copy(tables.freqs, tables.freqs + 4, p.freqs);
    *out = e.release();
  });
}

void sbnb_gp_destroy(sbnb_gp_engine* engine) {
  if (!engine) return;
  FloatingPointEnvironmentKeeper keep_caller_environment;
  cudaSetDevice(engine->device);
  delete engine;
}

int sbnb_gp_process_operations(sbnb_gp_engine* e, const int32_t* program, int64_t word_count) {
  return Guard([&] {
    Require(e != nullptr, "NULL GP engine.");
    Require(word_count >= 0 && (program != nullptr || word_count == 0), "NULL program.");
    if (word_count == 0) return;
    Bind(e);
    ValidateProgram(e, program, word_count);
    e->program.Upload(program, word_count, e->stream);
    SBNB_CUDA(cudaMemsetAsync(e->status.get(), 0, 2 * sizeof(int32_t), e->stream));
    GpParams p = e->params;
    p.program = e->program.get();
    p.word_count = word_count;
    SBNB_CUDA(cudaEventRecord(e->begin, e->stream));
    if (e->blocks == 1) {
      GpInterpretKernel<<<1, e->threads, 0, e->stream>>>(p);
    } else {
      void* args[] = {&p};
      SBNB_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(GpInterpretKernel), dim3(e->blocks),
                                            dim3(e->threads), args, 0, e->stream));
    }
    SBNB_CUDA(cudaGetLastError());
    SBNB_CUDA(cudaEventRecord(e->end, e->stream));
    e->launch_count++;
    int32_t status[2] = {0, 0};
    SBNB_CUDA(cudaMemcpyAsync(status, e->status.get(), sizeof(status), cudaMemcpyDeviceToHost, e->stream));
    SBNB_CUDA(cudaStreamSynchronize(e->stream));
    float ms = 0.f;
    SBNB_CUDA(cudaEventElapsedTime(&ms, e->begin, e->end));
    e->last_kernel_ms = ms;
    if (status[0] != kGpOk)
      Fail(status[0] == kGpFaultBadProgram ? SBNB_ERR_INVALID_ARGUMENT : SBNB_ERR_GP_ASSERT,
           std::string(FaultMessage(status[0])) + " (operation at word " + std::to_string(status[1]) + ")");
  });
}

int sbnb_gp_set_branch_lengths(sbnb_gp_engine* e, const double* branch_lengths) {
  return Guard([&] {
    Require(e && branch_lengths, "NULL argument.");
    Bind(e);
    SBNB_CUDA(cudaMemcpyAsync(e->branch_lengths.get(), branch_lengths, e->gpcsp_count * sizeof(double),
                              cudaMemcpyHostToDevice, e->stream));
    SBNB_CUDA(cudaStreamSynchronize(e->stream));
  });
}

int sbnb_gp_set_branch_lengths_to_constant(sbnb_gp_engine* e, double branch_length) {
  return Guard([&] {
    Require(e != nullptr, "NULL GP engine.");
    std::vector<double> values(e->gpcsp_count, branch_length);
    Bind(e);
    SBNB_CUDA(cudaMemcpyAsync(e->branch_lengths.get(), values.data(), values.size() * sizeof(double),
                              cudaMemcpyHostToDevice, e->stream));
    SBNB_CUDA(cudaStreamSynchronize(e->stream));
  });
}

int sbnb_gp_get_branch_lengths(sbnb_gp_engine* e, double* out) {
  return Guard([&] {
    Require(e && out, "NULL argument.");
    Bind(e);
    CopyOut(e, out, e->branch_lengths.get(), e->gpcsp_count);
  });
}

int sbnb_gp_reset_log_marginal_likelihood(sbnb_gp_engine* e) {
  return Guard([&] {
    Require(e != nullptr, "NULL GP engine.");
    Bind(e);
    std::vector<double> values(e->pattern_count, -INFINITY);
    SBNB_CUDA(cudaMemcpyAsync(e->log_marginal.get(), values.data(), values.size() * sizeof(double),
                              cudaMemcpyHostToDevice, e->stream));
    SBNB_CUDA(cudaStreamSynchronize(e->stream));
  });
}

int sbnb_gp_get_log_marginal_likelihood(sbnb_gp_engine* e, double* out) {
  return Guard([&] {
    Require(e && out, "NULL argument.");
    Bind(e);
    WeightedRowSums(e, e->log_marginal.get(), 1, out);
  });
}

int sbnb_gp_get_per_gpcsp_log_likelihoods(sbnb_gp_engine* e, int32_t start, int32_t length, double* out) {
  return Guard([&] {
    Require(e && out, "NULL argument.");
    Require(start >= 0 && length >= 0 && start + length <= e->gpcsp_count, "GPCSP range out of bounds.");
    Bind(e);
    WeightedRowSums(e, e->log_likelihoods.get() + static_cast<size_t>(start) * e->pattern_count, length, out);
  });
}

int sbnb_gp_get_per_gpcsp_components_of_full_log_marginal(sbnb_gp_engine* e, double* out) {
  return Guard([&] {
    Require(e && out, "NULL argument.");
    Bind(e);
    WeightedRowSums(e, e->log_likelihoods.get(), e->gpcsp_count, out);
    std::vector<double> q(e->gpcsp_count);
    CopyOut(e, q.data(), e->q.get(), e->gpcsp_count);
    for (int g = 0; g < e->gpcsp_count; g++) out[g] += static_cast<double>(e->site_count) * std::log(q[g]);
  });
}

int sbnb_gp_get_log_likelihood_matrix(sbnb_gp_engine* e, double* out) {
  return Guard([&] {
    Require(e && out, "NULL argument.");
    Bind(e);
    CopyOut(e, out, e->log_likelihoods.get(), static_cast<size_t>(e->gpcsp_count) * e->pattern_count);
  });
}

int sbnb_gp_get_sbn_parameters(sbnb_gp_engine* e, double* out) {
  return Guard([&] {
    Require(e && out, "NULL argument.");
    Bind(e);
    CopyOut(e, out, e->q.get(), e->gpcsp_count);
  });
}

int sbnb_gp_set_sbn_parameters(sbnb_gp_engine* e, const double* q) {
  return Guard([&] {
    Require(e && q, "NULL argument.");
    Bind(e);
    SBNB_CUDA(cudaMemcpyAsync(e->q.get(), q, e->gpcsp_count * sizeof(double), cudaMemcpyHostToDevice, e->stream));
    SBNB_CUDA(cudaStreamSynchronize(e->stream));
  });
}

int sbnb_gp_get_hybrid_marginals(sbnb_gp_engine* e, double* out) {
  return Guard([&] {
    Require(e && out, "NULL argument.");
    Bind(e);
    CopyOut(e, out, e->hybrid.get(), e->gpcsp_count);
  });
}

int sbnb_gp_set_hybrid_marginals(sbnb_gp_engine* e, const double* values) {
  return Guard([&] {
    Require(e && values, "NULL argument.");
    Bind(e);
    SBNB_CUDA(cudaMemcpyAsync(e->hybrid.get(), values, e->gpcsp_count * sizeof(double), cudaMemcpyHostToDevice,
                              e->stream));
    SBNB_CUDA(cudaStreamSynchronize(e->stream));
  });
}

int sbnb_gp_log_likelihood_and_derivative(sbnb_gp_engine* e, int32_t leafward, int32_t rootward, int32_t gpcsp,
                                          double* out) {
  return Guard([&] {
    Require(e && out, "NULL argument.");
    CheckPlv(e, leafward);
    CheckPlv(e, rootward);
    CheckGpcsp(e, gpcsp);
    Bind(e);
    e->scalars.Reserve(64);
    GpLogLikelihoodAndDerivativeKernel<<<1, 512, 0, e->stream>>>(e->params, leafward, rootward, gpcsp,
                                                                  e->scalars.get());
    SBNB_CUDA(cudaGetLastError());
    e->launch_count++;
    CopyOut(e, out, e->scalars.get(), 2);
  });
}

int sbnb_gp_transition_matrix(sbnb_gp_engine* e, double branch_length, double* out) {
  return Guard([&] {
    Require(e && out, "NULL argument.");
    Bind(e);
    e->scalars.Reserve(64);
    GpTransitionMatrixKernel<<<1, 32, 0, e->stream>>>(e->params, branch_length, e->scalars.get());
    SBNB_CUDA(cudaGetLastError());
    e->launch_count++;
    CopyOut(e, out, e->scalars.get(), 16);
  });
}

int sbnb_gp_quartet_hybrid_likelihoods(sbnb_gp_engine* e, int32_t central_gpcsp, const int32_t* rootward_tips,
                                       int32_t rootward_count, const int32_t* sister_tips, int32_t sister_count,
                                       const int32_t* rotated_tips, int32_t rotated_count,
                                       const int32_t* sorted_tips, int32_t sorted_count, double* out) {
  return Guard([&] {
    Require(e != nullptr, "NULL GP engine.");
    std::vector<double> values;
    QuartetLikelihoods(e, central_gpcsp, rootward_tips, rootward_count, sister_tips, sister_count, rotated_tips,
                       rotated_count, sorted_tips, sorted_count, &values);
    Require(out != nullptr || values.empty(), "NULL output.");
    std::copy(values.begin(), values.end(), out);
  });
}

int sbnb_gp_process_quartet_hybrid_request(sbnb_gp_engine* e, int32_t central_gpcsp,
                                           const int32_t* rootward_tips, int32_t rootward_count,
                                           const int32_t* sister_tips, int32_t sister_count,
                                           const int32_t* rotated_tips, int32_t rotated_count,
                                           const int32_t* sorted_tips, int32_t sorted_count) {
  return Guard([&] {
    Require(e != nullptr, "NULL GP engine.");
    // IsFullyFormed (quartet_hybrid_request.cpp): every tip list non-empty.
    if (rootward_count <= 0 || sister_count <= 0 || rotated_count <= 0 || sorted_count <= 0) return;
    std::vector<double> values;
    QuartetLikelihoods(e, central_gpcsp, rootward_tips, rootward_count, sister_tips, sister_count, rotated_tips,
                       rotated_count, sorted_tips, sorted_count, &values);
    // NumericalUtils::LogSum = left fold with LogAdd (numerical_utils.cpp:8)
    auto log_add = [](double x, double y) {
      if (y > x) std::swap(x, y);
      if (x == -INFINITY) return x;
      const double neg_diff = y - x;
      if (neg_diff < std::log(2.220446049250313e-16)) return x;
      return x + std::log(1.0 + std::exp(neg_diff));
    };
    double total = values[0];
    for (size_t i = 1; i < values.size(); i++) total = log_add(total, values[i]);
    SBNB_CUDA(cudaMemcpyAsync(e->hybrid.get() + central_gpcsp, &total, sizeof(double), cudaMemcpyHostToDevice,
                              e->stream));
    SBNB_CUDA(cudaStreamSynchronize(e->stream));
  });
}

int sbnb_gp_get_plv(sbnb_gp_engine* e, int32_t plv_idx, double* out) {
  return Guard([&] {
    Require(e && out, "NULL argument.");
    CheckPlv(e, plv_idx);
    Bind(e);
    CopyOut(e, out, e->plvs.get() + static_cast<size_t>(plv_idx) * e->pattern_count * 4, e->pattern_count * 4);
  });
}

int sbnb_gp_get_rescaling_counts(sbnb_gp_engine* e, int32_t* out) {
  return Guard([&] {
    Require(e && out, "NULL argument.");
    Bind(e);
    SBNB_CUDA(cudaMemcpyAsync(out, e->counts.get(), e->plv_count * sizeof(int32_t), cudaMemcpyDeviceToHost,
                              e->stream));
    SBNB_CUDA(cudaStreamSynchronize(e->stream));
  });
}

int64_t sbnb_gp_launch_count(const sbnb_gp_engine* e) { return e ? e->launch_count : -1; }
double sbnb_gp_last_kernel_ms(const sbnb_gp_engine* e) { return e ? e->last_kernel_ms : -1.0; }

}  // extern "C"
