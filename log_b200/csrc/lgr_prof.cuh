// Optional per-kernel CUDA-event timing (diagnostics for bench.py's roofline numbers).  Off by default; when on,
// every launcher brackets its kernel(s) with an event pair recorded on the launching stream.
#pragma once
#include <cuda_runtime.h>

namespace lgr {

enum KernelId { K_PROJECT_FWD = 0, K_TILE_SCAN, K_BIN_SCATTER, K_TILE_SORT, K_BLEND_FWD, K_BLEND_BWD, K_PROJECT_BWD,
                K_COMPUTE_RADIUS, K_SHARD_SEND, K_SHARD_RECV, K_SHARD_RETURN, K_SHARD_GATHER, K_COUNT };

struct Profiler {
  static constexpr int MAX_PAIRS = 8192;
  bool enabled = false;
  int used = 0;
  cudaEvent_t start[MAX_PAIRS], stop[MAX_PAIRS];
  int kid[MAX_PAIRS];
  int created = 0;
  int launches[K_COUNT] = {0};
  static Profiler& get() { static Profiler p; return p; }
  int begin(int k, cudaStream_t st) {
    if (!enabled || used >= MAX_PAIRS) return -1;
    if (used >= created) {
      if (cudaEventCreate(&start[created]) != cudaSuccess || cudaEventCreate(&stop[created]) != cudaSuccess) return -1;
      created++;
    }
    const int i = used++;
    kid[i] = k;
    cudaEventRecord(start[i], st);
    return i;
  }
  void end(int i, cudaStream_t st) { if (i >= 0) cudaEventRecord(stop[i], st); }
};

struct ProfScope {
  int idx; cudaStream_t st;
  ProfScope(int k, cudaStream_t s, int nlaunch = 1) : st(s) {
    Profiler& p = Profiler::get();
    idx = p.begin(k, s);
    if (p.enabled) p.launches[k] += nlaunch;
  }
  ~ProfScope() { Profiler::get().end(idx, st); }
};

}  // namespace lgr
