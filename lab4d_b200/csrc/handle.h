// Per-device library handle (internal).  Small read-only tables (pack-slice lists, weight-gradient job lists) are
// cached on the device per key and uploaded once on the caller's stream; the host copy stays alive with the entry.
#pragma once
#include <cuda_runtime.h>

#include <map>
#include <string>
#include <vector>

#include "kernels.h"

struct b200r_handle {
  int device;
  int n_sm;
  std::string err;
  struct Table { std::vector<uint8_t> host; void* dev; };
  std::map<std::string, Table> tables;
  float* d_scale;  // device scalars of the backward: [0] gradient scale, [1] its inverse, [2] max bits; [4..6] the same for the eikonal backward
};

namespace b200r {

// restores the caller's current device when an entry point returns
struct DeviceGuard {
  int prev;
  bool ok;
  explicit DeviceGuard(int dev) : prev(-1), ok(false) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    ok = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

inline std::string table_key(const char* kind, const void* a, size_t na, const void* b = nullptr, size_t nb = 0) {
  std::string k(kind);
  k.append(reinterpret_cast<const char*>(a), na);
  if (b) k.append(reinterpret_cast<const char*>(b), nb);
  return k;
}

// device copy of `bytes` bytes of `data` under `key` (uploaded once, asynchronously on `stream`); nullptr on failure
inline void* cached_table(b200r_handle* h, const std::string& key, const void* data, size_t bytes, cudaStream_t stream, cudaError_t* err) {
  auto it = h->tables.find(key);
  if (it != h->tables.end()) return it->second.dev;
  b200r_handle::Table t;
  t.host.assign(reinterpret_cast<const uint8_t*>(data), reinterpret_cast<const uint8_t*>(data) + bytes);
  t.dev = nullptr;
  cudaError_t e = cudaMalloc(&t.dev, bytes > 0 ? bytes : 16);
  if (e != cudaSuccess) { *err = e; return nullptr; }
  auto& slot = h->tables[key];
  slot = std::move(t);
  e = cudaMemcpyAsync(slot.dev, slot.host.data(), bytes, cudaMemcpyHostToDevice, stream);
  if (e != cudaSuccess) { *err = e; return nullptr; }
  return slot.dev;
}

}  // namespace b200r
