// oracle/beagle_cpu.cpp -- TEST INFRASTRUCTURE (the oracle), never product code.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs may link or execute this file.  The product path
// (libsbn_b200/csrc) never calls into it and has no CPU fallback.
//
// What it is: a plain single-threaded fp64 CPU restatement of the 17 entry
// points of the third-party BEAGLE library that libsbn calls (all call sites in
// reference src/fat_beagle.cpp; prototypes in include/libhmsbeagle/beagle.h).
// BEAGLE itself (beagle-dev/beagle-lib, branch hmc-clock, unpinned) is absent
// from /root/reference and cannot be installed here, so its published
// algorithm is restated and parity is anchored on the reference's own call
// sites and golden vectors:  the UNMODIFIED reference host code
// is linked against this file (oracle/Makefile -> oracle/_ref/) and must
// reproduce the pybeagle / physher / phylotorch numbers hard-coded in
// src/unrooted_sbn_instance.hpp:206-335 and src/rooted_sbn_instance.hpp:246-378.
//
// Semantics restated:
//  * one shared index space for partial and compact buffers
//    (fat_beagle.cpp:207-256): tips 0..n-1, internal post-order n..2n-2,
//    pre-order N+id;
//  * partials are [category][pattern][state] (state fastest), matrices are
//    [category][i][j] row-major, eigenvectors row-major (eigen_sugar.hpp:16-17);
//  * P_c(t) = V diag(exp(lambda r_c t)) V^-1, negative entries clamped to 0;
//  * post-order:  dest = (P1 L1) o (P2 L2); compact state s<S selects column s,
//    s>=S contributes 1;
//  * pre-order (fat_beagle.cpp:344-362): child1 = parent's pre-order partial,
//    child1 matrix = the node's own matrix (applied TRANSPOSED), child2 = the
//    sister's post-order partial with the sister's matrix;
//  * manual rescaling with raw scalers: per pattern divide by the max over
//    (category, state) (0 -> 1), store the raw max in the write buffer and add
//    log(max) to the cumulative buffer when one is given;
//  * edge derivatives: sum_k w_k [sum_c p_c pre^T dQ_c post]/[sum_c p_c pre^T post];
//  * root logL: sum_k w_k (log sum_c p_c sum_i pi_i L[c,k,i] + cum_k).
//
// ORACLE_REAL (default double) is the internal arithmetic type; building with
// -DORACLE_REAL="long double" gives an extended-precision witness.

#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <mutex>
#include <vector>

#include "libhmsbeagle/beagle.h"

#ifndef ORACLE_REAL
#define ORACLE_REAL double
#endif

namespace {

typedef ORACLE_REAL real;

struct Instance {
  int tips = 0, buffers = 0, S = 0, P = 0, C = 0, matrices = 0, scalers = 0;
  std::vector<std::vector<real>> partials;  // [buffer][c][k][i] (empty if unused)
  std::vector<std::vector<int>> states;     // [buffer][k]       (empty if not compact)
  std::vector<std::vector<real>> matrix;    // [matrix][c][i][j]
  std::vector<std::vector<real>> scale;     // [scaler][k]; index 0 is used as cumulative (log)
  std::vector<real> evec, ivec, eval;       // row-major S x S, S x S, S
  std::vector<real> cat_rates, cat_weights, freqs, pattern_weights;

  size_t PartialSize() const { return static_cast<size_t>(C) * P * S; }
};

std::mutex g_mutex;
std::vector<std::unique_ptr<Instance>> g_instances;

Instance* Get(int handle) {
  std::lock_guard<std::mutex> lock(g_mutex);
  if (handle < 0 || handle >= static_cast<int>(g_instances.size())) return nullptr;
  return g_instances[handle].get();
}

char g_resource_name[] = "CPU (oracle)";
char g_impl_name[] = "libsbn-b200 oracle: BEAGLE-equivalent CPU restatement";
char g_impl_desc[] = "plain fp64, single thread per instance";

// dest[c,k,i] for one child: either a full partial (mat-vec) or a compact tip.
inline void ChildTerm(const Instance& inst, int buffer, const real* mat_c, int c, int k,
                      real* out) {
  const int S = inst.S;
  if (!inst.states[buffer].empty()) {
    const int s = inst.states[buffer][k];
    for (int i = 0; i < S; i++) out[i] = (s < S) ? mat_c[i * S + s] : real(1);
  } else {
    const real* L = &inst.partials[buffer][(static_cast<size_t>(c) * inst.P + k) * S];
    for (int i = 0; i < S; i++) {
      real sum = 0;
      for (int j = 0; j < S; j++) sum += mat_c[i * S + j] * L[j];
      out[i] = sum;
    }
  }
}

// Per-pattern max rescale of a freshly written partial (a6/a9).
void Rescale(Instance& inst, std::vector<real>& dest, int scale_write, int cumulative) {
  const int S = inst.S, P = inst.P, C = inst.C;
  for (int k = 0; k < P; k++) {
    real max = 0;
    for (int c = 0; c < C; c++)
      for (int i = 0; i < S; i++) {
        real v = dest[(static_cast<size_t>(c) * P + k) * S + i];
        if (v > max) max = v;
      }
    if (max == 0) max = 1;
    const real one_over_max = real(1) / max;
    for (int c = 0; c < C; c++)
      for (int i = 0; i < S; i++) dest[(static_cast<size_t>(c) * P + k) * S + i] *= one_over_max;
    inst.scale[scale_write][k] = max;
    if (cumulative != BEAGLE_OP_NONE) inst.scale[cumulative][k] += std::log(max);
  }
}

void EnsurePartial(Instance& inst, int buffer) {
  if (inst.partials[buffer].empty()) inst.partials[buffer].assign(inst.PartialSize(), 0);
  inst.states[buffer].clear();
}

// One-hot / all-ones expansion of a compact tip for the derivative kernel.
inline void PostVector(const Instance& inst, int buffer, int c, int k, real* out) {
  const int S = inst.S;
  if (!inst.states[buffer].empty()) {
    const int s = inst.states[buffer][k];
    for (int i = 0; i < S; i++) out[i] = (s >= S || s == i) ? real(1) : real(0);
  } else {
    const real* L = &inst.partials[buffer][(static_cast<size_t>(c) * inst.P + k) * S];
    for (int i = 0; i < S; i++) out[i] = L[i];
  }
}

}  // namespace

extern "C" {

int beagleCreateInstance(int tipCount, int partialsBufferCount, int compactBufferCount,
                         int stateCount, int patternCount, int eigenBufferCount,
                         int matrixBufferCount, int categoryCount, int scaleBufferCount,
                         int* /*resourceList*/, int /*resourceCount*/, long preferenceFlags,
                         long requirementFlags, BeagleInstanceDetails* returnInfo) {
  if (tipCount < 1 || stateCount < 1 || patternCount < 0 || categoryCount < 1 ||
      eigenBufferCount < 1)
    return BEAGLE_ERROR_OUT_OF_RANGE;
  auto inst = std::make_unique<Instance>();
  inst->tips = tipCount;
  inst->buffers = partialsBufferCount + compactBufferCount;
  inst->S = stateCount;
  inst->P = patternCount;
  inst->C = categoryCount;
  inst->matrices = matrixBufferCount;
  inst->scalers = scaleBufferCount;
  inst->partials.resize(inst->buffers);
  inst->states.resize(inst->buffers);
  inst->matrix.assign(matrixBufferCount,
                      std::vector<real>(static_cast<size_t>(categoryCount) * stateCount * stateCount, 0));
  inst->scale.assign(scaleBufferCount, std::vector<real>(patternCount, 0));
  inst->evec.assign(stateCount * stateCount, 0);
  inst->ivec.assign(stateCount * stateCount, 0);
  inst->eval.assign(stateCount, 0);
  inst->cat_rates.assign(categoryCount, 1);
  inst->cat_weights.assign(categoryCount, real(1) / categoryCount);
  inst->freqs.assign(stateCount, real(1) / stateCount);
  inst->pattern_weights.assign(patternCount, 1);
  if (returnInfo != nullptr) {
    returnInfo->resourceNumber = 0;
    returnInfo->resourceName = g_resource_name;
    returnInfo->implName = g_impl_name;
    returnInfo->implDescription = g_impl_desc;
    long vector_flag = (preferenceFlags & BEAGLE_FLAG_VECTOR_NONE) ? BEAGLE_FLAG_VECTOR_NONE
                                                                   : BEAGLE_FLAG_VECTOR_SSE;
    returnInfo->flags = BEAGLE_FLAG_PRECISION_DOUBLE | BEAGLE_FLAG_COMPUTATION_SYNCH |
                        BEAGLE_FLAG_EIGEN_REAL | BEAGLE_FLAG_SCALING_MANUAL |
                        BEAGLE_FLAG_SCALERS_RAW | vector_flag | BEAGLE_FLAG_THREADING_NONE |
                        BEAGLE_FLAG_PROCESSOR_CPU | BEAGLE_FLAG_FRAMEWORK_CPU |
                        BEAGLE_FLAG_INVEVEC_STANDARD | (requirementFlags & BEAGLE_FLAG_SCALING_MANUAL);
  }
  std::lock_guard<std::mutex> lock(g_mutex);
  for (size_t h = 0; h < g_instances.size(); h++) {
    if (!g_instances[h]) {
      g_instances[h] = std::move(inst);
      return static_cast<int>(h);
    }
  }
  g_instances.push_back(std::move(inst));
  return static_cast<int>(g_instances.size()) - 1;
}

int beagleFinalizeInstance(int instance) {
  std::lock_guard<std::mutex> lock(g_mutex);
  if (instance < 0 || instance >= static_cast<int>(g_instances.size()) || !g_instances[instance])
    return BEAGLE_ERROR_UNINITIALIZED_INSTANCE;
  g_instances[instance].reset();
  return BEAGLE_SUCCESS;
}

int beagleSetTipStates(int instance, int tipIndex, const int* inStates) {
  Instance* inst = Get(instance);
  if (!inst) return BEAGLE_ERROR_UNINITIALIZED_INSTANCE;
  if (tipIndex < 0 || tipIndex >= inst->tips) return BEAGLE_ERROR_OUT_OF_RANGE;
  inst->states[tipIndex].assign(inStates, inStates + inst->P);
  for (int& s : inst->states[tipIndex])
    if (s < 0 || s > inst->S) s = inst->S;
  inst->partials[tipIndex].clear();
  return BEAGLE_SUCCESS;
}

int beagleSetTipPartials(int instance, int tipIndex, const double* inPartials) {
  Instance* inst = Get(instance);
  if (!inst) return BEAGLE_ERROR_UNINITIALIZED_INSTANCE;
  if (tipIndex < 0 || tipIndex >= inst->tips) return BEAGLE_ERROR_OUT_OF_RANGE;
  EnsurePartial(*inst, tipIndex);
  const size_t block = static_cast<size_t>(inst->P) * inst->S;
  for (int c = 0; c < inst->C; c++)
    for (size_t x = 0; x < block; x++) inst->partials[tipIndex][c * block + x] = inPartials[x];
  return BEAGLE_SUCCESS;
}

int beagleSetPartials(int instance, int bufferIndex, const double* inPartials) {
  Instance* inst = Get(instance);
  if (!inst) return BEAGLE_ERROR_UNINITIALIZED_INSTANCE;
  if (bufferIndex < 0 || bufferIndex >= inst->buffers) return BEAGLE_ERROR_OUT_OF_RANGE;
  EnsurePartial(*inst, bufferIndex);
  for (size_t x = 0; x < inst->PartialSize(); x++) inst->partials[bufferIndex][x] = inPartials[x];
  return BEAGLE_SUCCESS;
}

int beagleSetPatternWeights(int instance, const double* inPatternWeights) {
  Instance* inst = Get(instance);
  if (!inst) return BEAGLE_ERROR_UNINITIALIZED_INSTANCE;
  inst->pattern_weights.assign(inPatternWeights, inPatternWeights + inst->P);
  return BEAGLE_SUCCESS;
}

int beagleSetCategoryWeights(int instance, int index, const double* inCategoryWeights) {
  Instance* inst = Get(instance);
  if (!inst) return BEAGLE_ERROR_UNINITIALIZED_INSTANCE;
  if (index != 0) return BEAGLE_ERROR_OUT_OF_RANGE;
  inst->cat_weights.assign(inCategoryWeights, inCategoryWeights + inst->C);
  return BEAGLE_SUCCESS;
}

int beagleSetCategoryRates(int instance, const double* inCategoryRates) {
  Instance* inst = Get(instance);
  if (!inst) return BEAGLE_ERROR_UNINITIALIZED_INSTANCE;
  inst->cat_rates.assign(inCategoryRates, inCategoryRates + inst->C);
  return BEAGLE_SUCCESS;
}

int beagleSetStateFrequencies(int instance, int index, const double* inStateFrequencies) {
  Instance* inst = Get(instance);
  if (!inst) return BEAGLE_ERROR_UNINITIALIZED_INSTANCE;
  if (index != 0) return BEAGLE_ERROR_OUT_OF_RANGE;
  inst->freqs.assign(inStateFrequencies, inStateFrequencies + inst->S);
  return BEAGLE_SUCCESS;
}

int beagleSetEigenDecomposition(int instance, int eigenIndex, const double* inEigenVectors,
                                const double* inInverseEigenVectors,
                                const double* inEigenValues) {
  Instance* inst = Get(instance);
  if (!inst) return BEAGLE_ERROR_UNINITIALIZED_INSTANCE;
  if (eigenIndex != 0) return BEAGLE_ERROR_OUT_OF_RANGE;
  const int S = inst->S;
  inst->evec.assign(inEigenVectors, inEigenVectors + S * S);
  inst->ivec.assign(inInverseEigenVectors, inInverseEigenVectors + S * S);
  inst->eval.assign(inEigenValues, inEigenValues + S);
  return BEAGLE_SUCCESS;
}

int beagleUpdateTransitionMatrices(int instance, int eigenIndex, const int* probabilityIndices,
                                   const int* firstDerivativeIndices,
                                   const int* secondDerivativeIndices,
                                   const double* edgeLengths, int count) {
  Instance* inst = Get(instance);
  if (!inst) return BEAGLE_ERROR_UNINITIALIZED_INSTANCE;
  if (eigenIndex != 0) return BEAGLE_ERROR_OUT_OF_RANGE;
  // The reference always passes NULL derivative index lists (fat_beagle.cpp:310-311).
  if (firstDerivativeIndices != nullptr || secondDerivativeIndices != nullptr)
    return BEAGLE_ERROR_NO_IMPLEMENTATION;
  const int S = inst->S;
  std::vector<real> tmp(S * S);
  for (int e = 0; e < count; e++) {
    const int m = probabilityIndices[e];
    if (m < 0 || m >= inst->matrices) return BEAGLE_ERROR_OUT_OF_RANGE;
    for (int c = 0; c < inst->C; c++) {
      const real scaled_t = real(edgeLengths[e]) * inst->cat_rates[c];
      for (int i = 0; i < S; i++)
        for (int k = 0; k < S; k++)
          tmp[i * S + k] = inst->evec[i * S + k] * std::exp(inst->eval[k] * scaled_t);
      real* out = &inst->matrix[m][static_cast<size_t>(c) * S * S];
      for (int i = 0; i < S; i++)
        for (int j = 0; j < S; j++) {
          real sum = 0;
          for (int k = 0; k < S; k++) sum += tmp[i * S + k] * inst->ivec[k * S + j];
          out[i * S + j] = sum > 0 ? sum : real(0);
        }
    }
  }
  return BEAGLE_SUCCESS;
}

int beagleSetDifferentialMatrix(int instance, int matrixIndex, const double* inMatrix) {
  Instance* inst = Get(instance);
  if (!inst) return BEAGLE_ERROR_UNINITIALIZED_INSTANCE;
  if (matrixIndex < 0 || matrixIndex >= inst->matrices) return BEAGLE_ERROR_OUT_OF_RANGE;
  std::vector<real>& m = inst->matrix[matrixIndex];
  for (size_t x = 0; x < m.size(); x++) m[x] = inMatrix[x];
  return BEAGLE_SUCCESS;
}

int beagleResetScaleFactors(int instance, int cumulativeScaleIndex) {
  Instance* inst = Get(instance);
  if (!inst) return BEAGLE_ERROR_UNINITIALIZED_INSTANCE;
  if (cumulativeScaleIndex < 0 || cumulativeScaleIndex >= inst->scalers)
    return BEAGLE_ERROR_OUT_OF_RANGE;
  std::fill(inst->scale[cumulativeScaleIndex].begin(), inst->scale[cumulativeScaleIndex].end(),
            real(0));
  return BEAGLE_SUCCESS;
}

int beagleUpdatePartials(int instance, const BeagleOperation* operations, int operationCount,
                         int cumulativeScaleIndex) {
  Instance* inst = Get(instance);
  if (!inst) return BEAGLE_ERROR_UNINITIALIZED_INSTANCE;
  const int S = inst->S, P = inst->P, C = inst->C;
  std::vector<real> a(S), b(S);
  for (int op = 0; op < operationCount; op++) {
    const BeagleOperation& o = operations[op];
    if (o.destinationPartials < 0 || o.destinationPartials >= inst->buffers ||
        o.child1Partials < 0 || o.child1Partials >= inst->buffers || o.child2Partials < 0 ||
        o.child2Partials >= inst->buffers || o.child1TransitionMatrix < 0 ||
        o.child1TransitionMatrix >= inst->matrices || o.child2TransitionMatrix < 0 ||
        o.child2TransitionMatrix >= inst->matrices)
      return BEAGLE_ERROR_OUT_OF_RANGE;
    EnsurePartial(*inst, o.destinationPartials);
    std::vector<real>& dest = inst->partials[o.destinationPartials];
    for (int c = 0; c < C; c++) {
      const real* m1 = &inst->matrix[o.child1TransitionMatrix][static_cast<size_t>(c) * S * S];
      const real* m2 = &inst->matrix[o.child2TransitionMatrix][static_cast<size_t>(c) * S * S];
      for (int k = 0; k < P; k++) {
        ChildTerm(*inst, o.child1Partials, m1, c, k, a.data());
        ChildTerm(*inst, o.child2Partials, m2, c, k, b.data());
        real* d = &dest[(static_cast<size_t>(c) * P + k) * S];
        for (int i = 0; i < S; i++) d[i] = a[i] * b[i];
      }
    }
    if (o.destinationScaleWrite >= 0) {
      if (o.destinationScaleWrite >= inst->scalers) return BEAGLE_ERROR_OUT_OF_RANGE;
      Rescale(*inst, dest, o.destinationScaleWrite, cumulativeScaleIndex);
    }
  }
  return BEAGLE_SUCCESS;
}

int beagleUpdatePrePartials(int instance, const BeagleOperation* operations, int operationCount,
                            int cumulativeScaleIndex) {
  Instance* inst = Get(instance);
  if (!inst) return BEAGLE_ERROR_UNINITIALIZED_INSTANCE;
  const int S = inst->S, P = inst->P, C = inst->C;
  std::vector<real> sis(S), tmp(S);
  for (int op = 0; op < operationCount; op++) {
    const BeagleOperation& o = operations[op];
    if (o.destinationPartials < 0 || o.destinationPartials >= inst->buffers ||
        o.child1Partials < 0 || o.child1Partials >= inst->buffers || o.child2Partials < 0 ||
        o.child2Partials >= inst->buffers || inst->partials[o.child1Partials].empty())
      return BEAGLE_ERROR_OUT_OF_RANGE;
    EnsurePartial(*inst, o.destinationPartials);
    std::vector<real>& dest = inst->partials[o.destinationPartials];
    const std::vector<real>& parent_pre = inst->partials[o.child1Partials];
    for (int c = 0; c < C; c++) {
      const real* own = &inst->matrix[o.child1TransitionMatrix][static_cast<size_t>(c) * S * S];
      const real* sm = &inst->matrix[o.child2TransitionMatrix][static_cast<size_t>(c) * S * S];
      for (int k = 0; k < P; k++) {
        ChildTerm(*inst, o.child2Partials, sm, c, k, sis.data());
        const real* pp = &parent_pre[(static_cast<size_t>(c) * P + k) * S];
        for (int i = 0; i < S; i++) tmp[i] = pp[i] * sis[i];
        real* d = &dest[(static_cast<size_t>(c) * P + k) * S];
        for (int j = 0; j < S; j++) {
          real sum = 0;
          for (int i = 0; i < S; i++) sum += own[i * S + j] * tmp[i];
          d[j] = sum;
        }
      }
    }
    if (o.destinationScaleWrite >= 0) {
      if (o.destinationScaleWrite >= inst->scalers) return BEAGLE_ERROR_OUT_OF_RANGE;
      Rescale(*inst, dest, o.destinationScaleWrite, cumulativeScaleIndex);
    }
  }
  return BEAGLE_SUCCESS;
}

int beagleCalculateEdgeDerivatives(int instance, const int* postBufferIndices,
                                   const int* preBufferIndices,
                                   const int* derivativeMatrixIndices,
                                   const int* categoryWeightsIndices, int count,
                                   double* outDerivatives, double* outSumDerivatives,
                                   double* outSumSquaredDerivatives) {
  Instance* inst = Get(instance);
  if (!inst) return BEAGLE_ERROR_UNINITIALIZED_INSTANCE;
  const int S = inst->S, P = inst->P, C = inst->C;
  std::vector<real> post(S);
  for (int e = 0; e < count; e++) {
    const int post_buf = postBufferIndices[e], pre_buf = preBufferIndices[e];
    const int dm = derivativeMatrixIndices[e];
    if (post_buf < 0 || post_buf >= inst->buffers || pre_buf < 0 || pre_buf >= inst->buffers ||
        dm < 0 || dm >= inst->matrices || categoryWeightsIndices[0] != 0 ||
        inst->partials[pre_buf].empty())
      return BEAGLE_ERROR_OUT_OF_RANGE;
    real sum = 0, sum_sq = 0;
    for (int k = 0; k < P; k++) {
      real numerator = 0, denominator = 0;
      for (int c = 0; c < C; c++) {
        const real* dq = &inst->matrix[dm][static_cast<size_t>(c) * S * S];
        const real* pre = &inst->partials[pre_buf][(static_cast<size_t>(c) * P + k) * S];
        PostVector(*inst, post_buf, c, k, post.data());
        real num_c = 0, den_c = 0;
        for (int i = 0; i < S; i++) {
          real dq_post = 0;
          for (int j = 0; j < S; j++) dq_post += dq[i * S + j] * post[j];
          num_c += pre[i] * dq_post;
          den_c += pre[i] * post[i];
        }
        numerator += inst->cat_weights[c] * num_c;
        denominator += inst->cat_weights[c] * den_c;
      }
      const real derivative = numerator / denominator;
      if (outDerivatives != nullptr)
        outDerivatives[static_cast<size_t>(e) * P + k] = static_cast<double>(derivative);
      sum += inst->pattern_weights[k] * derivative;
      sum_sq += inst->pattern_weights[k] * derivative * derivative;
    }
    if (outSumDerivatives != nullptr) outSumDerivatives[e] = static_cast<double>(sum);
    if (outSumSquaredDerivatives != nullptr)
      outSumSquaredDerivatives[e] = static_cast<double>(sum_sq);
  }
  return BEAGLE_SUCCESS;
}

int beagleCalculateRootLogLikelihoods(int instance, const int* bufferIndices,
                                      const int* categoryWeightsIndices,
                                      const int* stateFrequenciesIndices,
                                      const int* cumulativeScaleIndices, int count,
                                      double* outSumLogLikelihood) {
  Instance* inst = Get(instance);
  if (!inst) return BEAGLE_ERROR_UNINITIALIZED_INSTANCE;
  if (count != 1) return BEAGLE_ERROR_NO_IMPLEMENTATION;  // fat_beagle passes 1
  const int root = bufferIndices[0];
  if (root < 0 || root >= inst->buffers || inst->partials[root].empty() ||
      categoryWeightsIndices[0] != 0 || stateFrequenciesIndices[0] != 0)
    return BEAGLE_ERROR_OUT_OF_RANGE;
  const int S = inst->S, P = inst->P, C = inst->C;
  const int cumulative = cumulativeScaleIndices[0];
  if (cumulative != BEAGLE_OP_NONE && (cumulative < 0 || cumulative >= inst->scalers))
    return BEAGLE_ERROR_OUT_OF_RANGE;
  real total = 0;
  for (int k = 0; k < P; k++) {
    real site = 0;
    for (int i = 0; i < S; i++) {
      real over_categories = 0;
      for (int c = 0; c < C; c++)
        over_categories +=
            inst->cat_weights[c] * inst->partials[root][(static_cast<size_t>(c) * P + k) * S + i];
      site += inst->freqs[i] * over_categories;
    }
    real log_site = std::log(site);
    if (cumulative != BEAGLE_OP_NONE) log_site += inst->scale[cumulative][k];
    total += inst->pattern_weights[k] * log_site;
  }
  *outSumLogLikelihood = static_cast<double>(total);
  return std::isnan(static_cast<double>(total)) ? BEAGLE_ERROR_FLOATING_POINT : BEAGLE_SUCCESS;
}

}  // extern "C"
