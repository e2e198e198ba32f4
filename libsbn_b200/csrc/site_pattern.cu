// Site-pattern compression on the device: the C ABI of include/sbn_b200_patterns.h,
// replacing the reference's SitePattern::Compress (src/site_pattern.cpp:77-115) and
// its symbol table (site_pattern.cpp:15-45).
//
// HBM-bound byte/integer work, five kernels over an alignment stored [taxon][site]:
//   SymbolizeHashKernel  4 sites per thread (32-bit loads along a taxon row, 8 taxa in
//                        flight): characters -> symbols 0..4 through a shared-memory
//                        table, four 32-bit multiplicative hashes per column, mixed
//                        into one 64-bit key
//   InsertKernel         one thread per site: claim / find the column's slot in an
//                        open-addressing table (atomicCAS on the 64-bit key), then
//                        atomicMin of the site index (the pattern's first
//                        appearance) and atomicAdd of its multiplicity
//   FlagKernel           flags first appearances
//   ScanBlocksKernel / ScanSumsKernel
//                        exclusive prefix sum of the flags = pattern index in order
//                        of first appearance (deterministic; no sort)
//   EmitKernel           first appearances write their column, weight and (per
//                        slot) pattern index
//   VerifyKernel         every other site compares its column byte for byte with its
//                        pattern's emitted column (a row of the compact output is
//                        small when there are many duplicates, so the gathers hit
//                        L1/L2): a 64-bit key shared by two different columns is
//                        caught -- the host then retries with another seed --
//                        never silently merged
// There is no CPU path in this file.

#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/sbn_b200_patterns.h"
#include "common.hpp"
#include "device_common.cuh"

namespace sbnb {

namespace {

constexpr int kSitesPerThread = 4;
constexpr int kRowAlignment = 16;  // device rows start on 16-byte boundaries
constexpr int kHashThreads = 256;
constexpr int kScanThreads = 256;
constexpr int kScanItems = 4;  // flags per thread -> 1024 per block
constexpr int kScanBlock = kScanThreads * kScanItems;
constexpr uint64_t kEmptyKey = ~0ull;
constexpr uint8_t kUnknownSymbol = 0xff;

// GetSymbolTable (site_pattern.cpp:15-45): DNA, degenerate nucleotides are gaps.
void BuildSymbolTable(uint8_t (&table)[256]) {
  std::memset(table, kUnknownSymbol, sizeof(table));
  const char* acgt = "ACGT";
  for (int i = 0; i < 4; i++) {
    table[static_cast<unsigned char>(acgt[i])] = static_cast<uint8_t>(i);
    table[static_cast<unsigned char>(acgt[i] - 'A' + 'a')] = static_cast<uint8_t>(i);
  }
  for (const char* c = "-NX?BDHKMRSUVWY"; *c; c++) table[static_cast<unsigned char>(*c)] = 4;
}

struct SymbolTable {
  uint8_t map[256];
};

__device__ __forceinline__ uint64_t Mix(uint64_t x) {  // splitmix64 finaliser
  x ^= x >> 30;
  x *= 0xbf58476d1ce4e5b9ull;
  x ^= x >> 27;
  x *= 0x94d049bb133111ebull;
  return x ^ (x >> 31);
}

// status[0] = 1 + first unknown character seen (0 = none); status[1] = collision flag.
__global__ void __launch_bounds__(kHashThreads) SymbolizeHashKernel(
    const uint8_t* __restrict__ sequences, uint8_t* __restrict__ symbols, uint64_t* __restrict__ keys,
    int32_t taxon_count, int64_t site_count, int64_t pitch, uint64_t seed, uint64_t key_mask,
    const SymbolTable table, int32_t* __restrict__ status) {
  __shared__ uint8_t map[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) map[i] = table.map[i];
  __syncthreads();
  const int64_t group = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t site0 = group * kSitesPerThread;
  if (site0 >= site_count) return;
  // Four 32-bit multiplicative hashes per column (cheap integer multiply-adds; a
  // 64-bit multiply is three of them), mixed into the 64-bit key at the end.
  uint32_t ha[kSitesPerThread], hb[kSitesPerThread], hc[kSitesPerThread], hd[kSitesPerThread];
  const uint32_t seed_low = static_cast<uint32_t>(seed), seed_high = static_cast<uint32_t>(seed >> 32);
#pragma unroll
  for (int k = 0; k < kSitesPerThread; k++) {
    ha[k] = 0x811c9dc5u ^ seed_low;
    hb[k] = 0x9e3779b9u + seed_high;
    hc[k] = 0x85ebca6bu ^ seed_high;
    hd[k] = 0xc2b2ae35u + seed_low;
  }
  int bad = 0;
  // (unrolled: a thread has one 4-byte load per taxon, so the loads of several taxa
  //  must be in flight together to cover the HBM latency)
#pragma unroll 8
  for (int t = 0; t < taxon_count; t++) {
    const uint32_t word = *reinterpret_cast<const uint32_t*>(sequences + static_cast<int64_t>(t) * pitch + site0);
    uint32_t packed = 0;
#pragma unroll
    for (int k = 0; k < kSitesPerThread; k++) {
      const uint32_t c = (word >> (8 * k)) & 0xffu;
      const uint32_t s = map[c];
      if (s == kUnknownSymbol && site0 + k < site_count && bad == 0) bad = 1 + static_cast<int>(c);
      packed |= (s & 0xffu) << (8 * k);
      ha[k] = (ha[k] ^ (s + 1)) * 0x01000193u;  // FNV-1a
      hb[k] = hb[k] * 0xcc9e2d51u + s + 1;
      hc[k] = (hc[k] + s + 1) * 0x1b873593u;
      hd[k] = (hd[k] ^ ((s + 1) * 0x27d4eb2fu)) * 0x165667b1u;
    }
    *reinterpret_cast<uint32_t*>(symbols + static_cast<int64_t>(t) * pitch + site0) = packed;
  }
  if (bad) atomicCAS(&status[0], 0, bad);
#pragma unroll
  for (int k = 0; k < kSitesPerThread; k++) {
    if (site0 + k < site_count) {
      uint64_t key = Mix((static_cast<uint64_t>(ha[k]) << 32 | hb[k])) ^
                     Mix((static_cast<uint64_t>(hc[k]) << 32 | hd[k]) + 0x9e3779b97f4a7c15ull);
      key &= key_mask;  // all ones, except in the collision-handling test
      if (key == kEmptyKey) key = 0;
      keys[site0 + k] = key;
    }
  }
}

__global__ void InsertKernel(const uint64_t* __restrict__ keys, unsigned long long* __restrict__ table_keys,
                             uint32_t* __restrict__ table_first, uint32_t* __restrict__ table_count,
                             uint32_t* __restrict__ slot_of, int64_t site_count, uint32_t mask) {
  const int64_t site = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (site >= site_count) return;
  const uint64_t key = keys[site];
  uint32_t slot = static_cast<uint32_t>(Mix(key)) & mask;
  while (true) {
    const unsigned long long previous = atomicCAS(&table_keys[slot], kEmptyKey, key);
    if (previous == kEmptyKey || previous == key) break;
    slot = (slot + 1) & mask;
  }
  slot_of[site] = slot;
  atomicMin(&table_first[slot], static_cast<uint32_t>(site));
  atomicAdd(&table_count[slot], 1u);
}

__global__ void FlagKernel(const uint32_t* __restrict__ slot_of, const uint32_t* __restrict__ table_first,
                           uint32_t* __restrict__ flags, int64_t site_count) {
  const int64_t site = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (site < site_count) flags[site] = (table_first[slot_of[site]] == site) ? 1u : 0u;
}

// Block-local exclusive scan of 1024 flags; the block's total goes to block_sums.
__global__ void __launch_bounds__(kScanThreads) ScanBlocksKernel(const uint32_t* __restrict__ flags,
                                                                 uint32_t* __restrict__ index,
                                                                 uint32_t* __restrict__ block_sums,
                                                                 int64_t count) {
  __shared__ uint32_t warp_totals[kScanThreads / 32];
  const int64_t base = (static_cast<int64_t>(blockIdx.x) * kScanThreads + threadIdx.x) * kScanItems;
  uint32_t item[kScanItems], sum = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; k++) {
    item[k] = (base + k < count) ? flags[base + k] : 0u;
    sum += item[k];
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inclusive = sum;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t up = __shfl_up_sync(0xffffffffu, inclusive, d);
    if (lane >= d) inclusive += up;
  }
  if (lane == 31) warp_totals[warp] = inclusive;
  __syncthreads();
  uint32_t offset = 0, total = 0;
#pragma unroll
  for (int w = 0; w < kScanThreads / 32; w++) {
    if (w < warp) offset += warp_totals[w];
    total += warp_totals[w];
  }
  uint32_t running = offset + inclusive - sum;
#pragma unroll
  for (int k = 0; k < kScanItems; k++) {
    if (base + k < count) index[base + k] = running;
    running += item[k];
  }
  if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// One block: exclusive scan of the block totals in place, grand total to *total.
__global__ void __launch_bounds__(1024) ScanSumsKernel(uint32_t* __restrict__ block_sums, int64_t blocks,
                                                       unsigned long long* __restrict__ total) {
  __shared__ uint32_t warp_totals[32];
  __shared__ unsigned long long carry_shared;
  if (threadIdx.x == 0) carry_shared = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int64_t start = 0; start < blocks; start += 1024) {
    const int64_t i = start + threadIdx.x;
    const uint32_t value = (i < blocks) ? block_sums[i] : 0u;
    uint32_t inclusive = value;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t up = __shfl_up_sync(0xffffffffu, inclusive, d);
      if (lane >= d) inclusive += up;
    }
    if (lane == 31) warp_totals[warp] = inclusive;
    __syncthreads();
    uint32_t offset = 0, chunk_total = 0;
    for (int w = 0; w < 32; w++) {
      if (w < warp) offset += warp_totals[w];
      chunk_total += warp_totals[w];
    }
    const unsigned long long carry = carry_shared;
    if (i < blocks) block_sums[i] = static_cast<uint32_t>(carry + offset + inclusive - value);
    __syncthreads();
    if (threadIdx.x == 0) carry_shared = carry + chunk_total;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry_shared;
}

__global__ void EmitKernel(const uint8_t* __restrict__ symbols, const uint32_t* __restrict__ slot_of,
                           const uint32_t* __restrict__ table_count, const uint32_t* __restrict__ flags,
                           const uint32_t* __restrict__ index, const uint32_t* __restrict__ block_sums,
                           uint8_t* __restrict__ patterns, double* __restrict__ weights,
                           uint32_t* __restrict__ table_pattern, int32_t taxon_count, int64_t site_count,
                           int64_t pitch, int64_t pattern_count) {
  const int64_t site = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (site >= site_count || !flags[site]) return;
  const int64_t p = static_cast<int64_t>(index[site]) + block_sums[site / kScanBlock];
  const uint32_t slot = slot_of[site];
  weights[p] = static_cast<double>(table_count[slot]);
  table_pattern[slot] = static_cast<uint32_t>(p);
#pragma unroll 16
  for (int t = 0; t < taxon_count; t++)
    patterns[static_cast<int64_t>(t) * pattern_count + p] = symbols[static_cast<int64_t>(t) * pitch + site];
}

__global__ void VerifyKernel(const uint8_t* __restrict__ symbols, const uint32_t* __restrict__ slot_of,
                             const uint32_t* __restrict__ flags, const uint32_t* __restrict__ table_pattern,
                             const uint8_t* __restrict__ patterns, int32_t taxon_count, int64_t site_count,
                             int64_t pitch, int64_t pattern_count, int32_t* __restrict__ status) {
  const int64_t site = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (site >= site_count || flags[site]) return;
  const int64_t p = table_pattern[slot_of[site]];
  bool same = true;
#pragma unroll 16
  for (int t = 0; t < taxon_count; t++)
    same = same && (symbols[static_cast<int64_t>(t) * pitch + site] ==
                    patterns[static_cast<int64_t>(t) * pattern_count + p]);
  if (!same) status[1] = 1;
}

void Compress(int32_t n, int64_t S, const char* sequences, int32_t device, uint8_t* out_patterns,
              double* out_weights, int64_t* out_pattern_count, double* out_device_ms) {
  Require(n > 0, "Site pattern compression needs at least one sequence.");
  Require(S >= 0 && S < (1ll << 30), "Site count out of range [0, 2^30).");  // 32-bit site ids, 2S-slot table
  Require(out_pattern_count != nullptr, "out_pattern_count is NULL.");
  int device_count = 0;
  if (cudaGetDeviceCount(&device_count) != cudaSuccess || device_count == 0)
    Fail(SBNB_ERR_NO_DEVICE, "No usable CUDA device: libsbn_b200 has no CPU fallback.");
  Require(device >= 0 && device < device_count, "CUDA device index out of range.");
  SBNB_CUDA(cudaSetDevice(device));
  *out_pattern_count = 0;
  if (out_device_ms) *out_device_ms = 0.0;
  if (S == 0) return;
  Require(sequences && out_patterns && out_weights, "NULL buffer passed to sbnb_compress_site_patterns.");

  const int64_t pitch = (S + kRowAlignment - 1) / kRowAlignment * kRowAlignment;
  uint32_t table_size = 1024;
  while (table_size < 2ull * static_cast<uint64_t>(S)) table_size <<= 1;
  const int64_t scan_blocks = (S + kScanBlock - 1) / kScanBlock;

  DeviceArray<uint8_t> d_sequences, d_symbols, d_patterns;
  DeviceArray<uint64_t> d_keys;
  DeviceArray<unsigned long long> d_table_keys, d_total;
  DeviceArray<uint32_t> d_table_first, d_table_count, d_table_pattern, d_slot_of, d_flags, d_index, d_block_sums;
  DeviceArray<int32_t> d_status;
  DeviceArray<double> d_weights;
  d_sequences.Reserve(static_cast<size_t>(n) * pitch);
  d_symbols.Reserve(static_cast<size_t>(n) * pitch);
  d_keys.Reserve(S);
  d_table_keys.Reserve(table_size);
  d_table_first.Reserve(table_size);
  d_table_count.Reserve(table_size);
  d_table_pattern.Reserve(table_size);
  d_slot_of.Reserve(S);
  d_flags.Reserve(S);
  d_index.Reserve(S);
  d_block_sums.Reserve(scan_blocks);
  d_total.Reserve(1);
  d_status.Reserve(2);
  d_weights.Reserve(S);
  d_patterns.Reserve(static_cast<size_t>(n) * pitch);
  // the padding columns of the last row segment must hold valid characters
  SBNB_CUDA(cudaMemset(d_sequences.get(), 'A', static_cast<size_t>(n) * pitch));
  SBNB_CUDA(cudaMemcpy2D(d_sequences.get(), pitch, sequences, S, S, n, cudaMemcpyHostToDevice));

  SymbolTable table;
  BuildSymbolTable(table.map);
  cudaEvent_t begin, end;
  SBNB_CUDA(cudaEventCreate(&begin));
  SBNB_CUDA(cudaEventCreate(&end));
  struct EventGuard {
    cudaEvent_t a, b;
    ~EventGuard() {
      cudaEventDestroy(a);
      cudaEventDestroy(b);
    }
  } event_guard{begin, end};

  const int site_blocks = static_cast<int>((S + 255) / 256);
  // Test hook: keep only the low bits of the column keys, so that different columns
  // share keys and the verification / rehash path is exercised.
  uint64_t key_mask = ~0ull;
  if (const char* bits = std::getenv("SBNB_DEBUG_PATTERN_KEY_BITS"))
    key_mask = (1ull << std::max(1, std::min(63, std::atoi(bits)))) - 1;
  int64_t pattern_count = 0;
  double device_ms = 0.0;
  bool done = false;
  for (uint64_t seed = 0; seed < 4 && !done; seed++) {
    SBNB_CUDA(cudaMemset(d_status.get(), 0, 2 * sizeof(int32_t)));
    SBNB_CUDA(cudaMemset(d_table_keys.get(), 0xff, static_cast<size_t>(table_size) * sizeof(uint64_t)));
    SBNB_CUDA(cudaMemset(d_table_first.get(), 0xff, static_cast<size_t>(table_size) * sizeof(uint32_t)));
    SBNB_CUDA(cudaMemset(d_table_count.get(), 0, static_cast<size_t>(table_size) * sizeof(uint32_t)));
    SBNB_CUDA(cudaEventRecord(begin));
    const int64_t groups = pitch / kSitesPerThread;
    SymbolizeHashKernel<<<static_cast<int>((groups + kHashThreads - 1) / kHashThreads), kHashThreads>>>(
        d_sequences.get(), d_symbols.get(), d_keys.get(), n, S, pitch, seed * 0x9e3779b97f4a7c15ull, key_mask,
        table, d_status.get());
    InsertKernel<<<site_blocks, 256>>>(d_keys.get(), d_table_keys.get(), d_table_first.get(),
                                       d_table_count.get(), d_slot_of.get(), S, table_size - 1);
    FlagKernel<<<site_blocks, 256>>>(d_slot_of.get(), d_table_first.get(), d_flags.get(), S);
    ScanBlocksKernel<<<static_cast<int>(scan_blocks), kScanThreads>>>(d_flags.get(), d_index.get(),
                                                                      d_block_sums.get(), S);
    ScanSumsKernel<<<1, 1024>>>(d_block_sums.get(), scan_blocks, d_total.get());
    // The output rows are written `pitch` apart (the pattern count is not known on
    // the host yet), so the whole pipeline is queued without a round trip.
    EmitKernel<<<site_blocks, 256>>>(d_symbols.get(), d_slot_of.get(), d_table_count.get(), d_flags.get(),
                                     d_index.get(), d_block_sums.get(), d_patterns.get(), d_weights.get(),
                                     d_table_pattern.get(), n, S, pitch, pitch);
    // every site that repeats a pattern is held against it, byte for byte
    VerifyKernel<<<site_blocks, 256>>>(d_symbols.get(), d_slot_of.get(), d_flags.get(), d_table_pattern.get(),
                                       d_patterns.get(), n, S, pitch, pitch, d_status.get());
    SBNB_CUDA(cudaEventRecord(end));
    SBNB_CUDA(cudaGetLastError());
    SBNB_CUDA(cudaEventSynchronize(end));
    float ms = 0.f;
    SBNB_CUDA(cudaEventElapsedTime(&ms, begin, end));
    device_ms = ms;
    int32_t status[2];
    unsigned long long total = 0;
    SBNB_CUDA(cudaMemcpy(status, d_status.get(), sizeof(status), cudaMemcpyDeviceToHost));
    SBNB_CUDA(cudaMemcpy(&total, d_total.get(), sizeof(total), cudaMemcpyDeviceToHost));
    if (status[0] != 0) {
      // SymbolTableAt (site_pattern.cpp:47-55)
      Fail(SBNB_ERR_INVALID_ARGUMENT,
           std::string("Symbol '") + static_cast<char>(status[0] - 1) + "' not known.");
    }
    if (status[1] != 0) continue;  // two different columns shared a 64-bit key: rehash
    pattern_count = static_cast<int64_t>(total);
    done = true;
  }
  if (!done) Fail(SBNB_ERR_CUDA, "Site pattern hashing collided under four seeds.");
  SBNB_CUDA(cudaMemcpy2D(out_patterns, pattern_count, d_patterns.get(), pitch, pattern_count, n,
                         cudaMemcpyDeviceToHost));
  SBNB_CUDA(cudaMemcpy(out_weights, d_weights.get(), static_cast<size_t>(pattern_count) * sizeof(double),
                       cudaMemcpyDeviceToHost));
  *out_pattern_count = pattern_count;
  if (out_device_ms) *out_device_ms = device_ms;
}

}  // namespace

}  // namespace sbnb

extern "C" int sbnb_compress_site_patterns(int32_t taxon_count, int64_t site_count, const char* sequences,
                                           int32_t device, uint8_t* out_patterns, double* out_weights,
                                           int64_t* out_pattern_count, double* out_device_ms) {
  return sbnb::Guard([&] {
    sbnb::Compress(taxon_count, site_count, sequences, device, out_patterns, out_weights, out_pattern_count,
                   out_device_ms);
  });
}
