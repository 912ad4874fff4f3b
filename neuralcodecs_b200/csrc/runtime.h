// Small host-side runtime shared by the engine: status/error plumbing, device buffers,
// per-launch CUDA-event profiler and the launch context every kernel wrapper receives.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/neuralcodecs_cuda.h"

namespace nc {

struct Error : public std::runtime_error {
  nc_status status;
  Error(nc_status s, const std::string& m) : std::runtime_error(m), status(s) {}
};

#define NC_CUDA(expr)                                                                          \
  do {                                                                                         \
    cudaError_t e__ = (expr);                                                                  \
    if (e__ != cudaSuccess) {                                                                  \
      cudaGetLastError();                                                                      \
      throw ::nc::Error(e__ == cudaErrorMemoryAllocation ? NC_OUT_OF_MEMORY : NC_CUDA_ERROR,   \
                        std::string(#expr) + ": " + cudaGetErrorString(e__));                  \
    }                                                                                          \
  } while (0)

inline void check_launch(int rc, const char* what) {
  if (rc == 0) return;
  if (rc < 0) throw Error(NC_INTERNAL, std::string(what) + ": shape not supported by kernel");
  cudaGetLastError();
  throw Error(NC_CUDA_ERROR, std::string(what) + ": " + cudaGetErrorString((cudaError_t)rc));
}

// Growable device allocation (never shrinks); plain cudaMalloc, owned by a handle.
class DeviceBuffer {
 public:
  DeviceBuffer() = default;
  DeviceBuffer(const DeviceBuffer&) = delete;
  DeviceBuffer& operator=(const DeviceBuffer&) = delete;
  ~DeviceBuffer() { release(); }
  void* reserve(size_t bytes) {
    if (bytes > cap_) {
      release();
      NC_CUDA(cudaMalloc(&ptr_, bytes));
      cap_ = bytes;
    }
    return ptr_;
  }
  template <typename T>
  T* as() const { return reinterpret_cast<T*>(ptr_); }
  size_t capacity() const { return cap_; }
  void release() {
    if (ptr_) cudaFree(ptr_);
    ptr_ = nullptr;
    cap_ = 0;
  }

 private:
  void* ptr_ = nullptr;
  size_t cap_ = 0;
};

template <typename T>
T* upload(const std::vector<T>& host) {
  T* d = nullptr;
  if (host.empty()) return d;
  NC_CUDA(cudaMalloc(&d, host.size() * sizeof(T)));
  NC_CUDA(cudaMemcpy(d, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice));
  return d;
}

struct KernelStat {
  uint64_t launches = 0;
  double ms = 0, flops = 0, bytes = 0;
};

// Brackets each kernel launch with CUDA events on the launching stream (option profile=1).
class Profiler {
 public:
  bool enabled = false;
  bool by_layer = false;  // key the report by layer name instead of kernel name
  ~Profiler() {
    for (auto e : pool_) cudaEventDestroy(e);
  }
  int begin(cudaStream_t s) {
    if (!enabled) return -1;
    cudaEvent_t a = get(), b = get();
    cudaEventRecord(a, s);
    pending_.push_back({a, b, "", 0, 0});
    return (int)pending_.size() - 1;
  }
  void end(int id, cudaStream_t s, const std::string& name, double flops, double bytes) {
    if (id < 0) return;
    auto& p = pending_[id];
    p.name = name;
    p.flops = flops;
    p.bytes = bytes;
    cudaEventRecord(p.b, s);
  }
  void flush() {
    for (auto& p : pending_) {
      cudaEventSynchronize(p.b);
      float ms = 0;
      cudaEventElapsedTime(&ms, p.a, p.b);
      auto& st = stats_[p.name];
      st.launches++;
      st.ms += ms;
      st.flops += p.flops;
      st.bytes += p.bytes;
      pool_free_.push_back(p.a);
      pool_free_.push_back(p.b);
    }
    pending_.clear();
  }
  std::string report_json(bool reset) {
    flush();
    std::string out = "{";
    bool first = true;
    char buf[512];
    for (auto& kv : stats_) {
      snprintf(buf, sizeof buf, "%s\"%s\": {\"launches\": %llu, \"ms\": %.6f, \"flops\": %.6e, \"bytes\": %.6e}",
               first ? "" : ", ", kv.first.c_str(), (unsigned long long)kv.second.launches, kv.second.ms,
               kv.second.flops, kv.second.bytes);
      out += buf;
      first = false;
    }
    out += "}";
    if (reset) stats_.clear();
    return out;
  }

 private:
  struct Pending {
    cudaEvent_t a, b;
    std::string name;
    double flops, bytes;
  };
  cudaEvent_t get() {
    if (!pool_free_.empty()) {
      cudaEvent_t e = pool_free_.back();
      pool_free_.pop_back();
      return e;
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    pool_.push_back(e);
    return e;
  }
  std::vector<cudaEvent_t> pool_, pool_free_;
  std::vector<Pending> pending_;
  std::map<std::string, KernelStat> stats_;
};

struct LaunchCtx {
  cudaStream_t stream = nullptr;
  int num_sms = 148;
  Profiler* prof = nullptr;
  uint64_t* launches = nullptr;
  // per-handle kernel policies (engine options "fast_sin" / "fuse_ru"); never process-global
  int fast_sin = -1;   // Snake sin(): -1 = MUFU sin on the tf32 / bf16x3 / f16 tensor-core layers, precise elsewhere; 0 = always precise; 1 = always MUFU
  int fuse_ru = 1;     // 0: never use the fused ResidualUnit kernel

  int begin() const { return prof ? prof->begin(stream) : -1; }
  void end(int id, const std::string& name, double flops, double bytes, const std::string& layer = "") const {
    if (launches) ++*launches;
    if (prof) prof->end(id, stream, (prof->by_layer && !layer.empty()) ? layer + " [" + name + "]" : name, flops, bytes);
  }
};

}  // namespace nc
