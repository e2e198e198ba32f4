// Host-side CUDA plumbing shared by engine.cu and gp_engine.cu: error mapping,
// device / page-locked buffers that grow on demand, and the C-ABI exception guard.
#ifndef SBNB_DEVICE_COMMON_CUH_
#define SBNB_DEVICE_COMMON_CUH_

#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>  // header-only; ranges cost nothing unless a profiler is attached

#include <algorithm>
#include <cfenv>
#include <cstdlib>
#include <new>
#include <string>

#include "common.hpp"

namespace sbnb {

#define SBNB_CUDA(call)                                                                     \
  do {                                                                                      \
    cudaError_t status_ = (call);                                                           \
    if (status_ != cudaSuccess) {                                                           \
      const int code_ = (status_ == cudaErrorMemoryAllocation)                              \
                            ? SBNB_ERR_OUT_OF_MEMORY                                        \
                            : ((status_ == cudaErrorNoDevice ||                             \
                                status_ == cudaErrorInsufficientDriver)                     \
                                   ? SBNB_ERR_NO_DEVICE                                     \
                                   : SBNB_ERR_CUDA);                                        \
      ::sbnb::Fail(code_, std::string(#call) + ": " + cudaGetErrorString(status_));         \
    }                                                                                       \
  } while (0)

// Device allocation that grows on demand and is reused between calls.
template <typename T>
class DeviceArray {
 public:
  DeviceArray() = default;
  DeviceArray(const DeviceArray&) = delete;
  DeviceArray& operator=(const DeviceArray&) = delete;
  ~DeviceArray() { cudaFree(ptr_); }
  void Reserve(size_t count) {
    if (count <= capacity_) return;
    cudaFree(ptr_);
    ptr_ = nullptr;
    capacity_ = 0;
    SBNB_CUDA(cudaMalloc(&ptr_, std::max<size_t>(count, 1) * sizeof(T)));
    capacity_ = count;
  }
  // Returns the number of bytes copied host -> device.
  size_t Upload(const T* host, size_t count, cudaStream_t stream) {
    Reserve(count);
    if (count) SBNB_CUDA(cudaMemcpyAsync(ptr_, host, count * sizeof(T), cudaMemcpyHostToDevice, stream));
    return count * sizeof(T);
  }
  T* get() const { return ptr_; }
  size_t capacity() const { return capacity_; }

 private:
  T* ptr_ = nullptr;
  size_t capacity_ = 0;
};

// Page-locked host staging area (grows on demand, reused between calls) so that
// the host -> device copies of a staged batch are real asynchronous DMA.
class PinnedArena {
 public:
  PinnedArena() = default;
  PinnedArena(const PinnedArena&) = delete;
  PinnedArena& operator=(const PinnedArena&) = delete;
  ~PinnedArena() { cudaFreeHost(base_); }
  void Reset(size_t bytes) {
    if (bytes > capacity_) {
      cudaFreeHost(base_);
      base_ = nullptr;
      capacity_ = 0;
      SBNB_CUDA(cudaMallocHost(&base_, bytes));
      capacity_ = bytes;
    }
    used_ = 0;
  }
  template <typename T>
  T* Take(size_t count) {
    used_ = (used_ + 255) / 256 * 256;
    T* out = reinterpret_cast<T*>(static_cast<char*>(base_) + used_);
    used_ += count * sizeof(T);
    if (used_ > capacity_) Fail(SBNB_ERR_OUT_OF_MEMORY, "pinned staging arena overflow");
    return out;
  }

 private:
  void* base_ = nullptr;
  size_t capacity_ = 0, used_ = 0;
};


// Integer tuning knob from the environment (development aid).
inline int EnvInt(const char* name, int fallback) {
  const char* value = std::getenv(name);
  return value ? std::atoi(value) : fallback;
}

// NVTX range over a host-side phase (stage / run / fetch / finish), for nsys / ncu timelines.
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};

// Every C-ABI entry point runs inside Guard: exceptions become error codes, and
// the caller's floating-point environment comes back exactly as it went in.  The
// reference branches on sticky FE_OVERFLOW / FE_UNDERFLOW flags
// (sbn_probability.cpp:278-281), so a flag raised in here (an exp() that underflows
// while building a model table, or inside the CUDA driver) would change what its
// next SBN training computes.
struct FloatingPointEnvironmentKeeper {
  std::fenv_t saved;
  FloatingPointEnvironmentKeeper() { std::fegetenv(&saved); }
  ~FloatingPointEnvironmentKeeper() { std::fesetenv(&saved); }
};

template <typename F>
int Guard(F&& body) {
  FloatingPointEnvironmentKeeper keep_caller_environment;
  try {
    body();
    return SBNB_OK;
  } catch (const Error& error) {
    SetLastError(error.what());
    return error.code();
  } catch (const std::bad_alloc&) {
    SetLastError("host allocation failed");
    return SBNB_ERR_OUT_OF_MEMORY;
  } catch (const std::exception& error) {
    SetLastError(error.what());
    return SBNB_ERR_INVALID_ARGUMENT;
  }
}


}  // namespace sbnb

#endif  // SBNB_DEVICE_COMMON_CUH_
