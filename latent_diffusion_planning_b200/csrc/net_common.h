// Host-side building blocks shared by planner.cu / idm.cu / vae.cu: device allocation arena, canonical
// weight-blob walker, weight packing for the tcgen05 path, and the DDPM coefficient table.
#pragma once
#include <map>
#include <memory>
#include <utility>
#include <vector>

#include "kernels.h"

namespace ldp {

// Owns device allocations; freed together at handle destruction.
class Arena {
 public:
  ~Arena() { release(); }
  int alloc(void** out, size_t bytes, bool zero = true) {
    void* p = nullptr;
    if (bytes == 0) bytes = 16;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) {
      set_last_error(std::string("cudaMalloc(") + std::to_string(bytes) + ") failed: " + cudaGetErrorString(e));
      return LDP_ERR_CUDA;
    }
    if (zero) {
      // The fill runs on the legacy stream and is asynchronous to the host; work queued afterwards on a NON-BLOCKING stream (the trainer's
      // side stream, the capture streams) is not ordered behind it, so wait for it here.  (Found as a 1-in-4 `illegal instruction`: a
      // tensor map copied on the side stream was zeroed by the fill of its own allocation.)  Allocations are create-time / first-step events.
      e = cudaMemset(p, 0, bytes);
      if (e == cudaSuccess) e = cudaStreamSynchronize(cudaStreamLegacy);
      if (e != cudaSuccess) {
        set_last_error(std::string("cudaMemset failed: ") + cudaGetErrorString(e));
        return LDP_ERR_CUDA;
      }
    }
    ptrs_.push_back(p);
    total_ += bytes;
    *out = p;
    return LDP_OK;
  }
  template <typename T>
  int alloc_t(T** out, size_t count, bool zero = true) {
    return alloc(reinterpret_cast<void**>(out), count * sizeof(T), zero);
  }
  void release() {
    for (void* p : ptrs_) cudaFree(p);
    ptrs_.clear();
    total_ = 0;
  }
  size_t total() const { return total_; }

 private:
  std::vector<void*> ptrs_;
  size_t total_ = 0;
};

// Walks the canonical float32 blob (params.py spec order): take(n) returns the device pointer of the next tensor.
struct BlobWalker {
  const float* base = nullptr;
  uint64_t pos = 0;
  const float* take(uint64_t n) {
    const float* p = base ? base + pos : nullptr;
    pos += n;
    return p;
  }
};

// A bf16 activation tensor, channels-last, viewed as (items, rows_per_item, channels) with row pitch ld.
struct ActBf16 {
  __nv_bfloat16* p = nullptr;
  int c = 0;    // channels
  int ld = 0;   // row pitch in elements (multiple of 8)
};

// One source of the K dimension of a packed weight: `taps` kernel taps over `c` channels that live at rows
// [row0 + tap*tap_stride, +c) of the Flax kernel viewed as [K][N].
struct PackedW {
  __nv_bfloat16* wt = nullptr;   // [n_pad][kp]
  int kp = 0, n_pad = 0;
  TcKBlock* kb_dev = nullptr;
  int num_kb = 0;
  TcRun* runs_dev = nullptr;
  int num_runs = 0;
  std::vector<TcRun> runs_host;
};

// Upload a stage table and its run-length form into `arena`.
inline int upload_stage_table(Arena& arena, const std::vector<TcStage>& st, PackedW* pw) {
  std::vector<TcRun> runs;
  for (const TcStage& e : st) {
    const int nw = (e.src_acc >> 16) & 0xff;
    if (!runs.empty()) {
      TcRun& r = runs.back();
      if (r.src_acc == e.src_acc && r.d12 == e.d12 && e.c0 == r.c0 + 64 * r.count && e.wk == r.wk + nw * r.count) {
        ++r.count;
        continue;
      }
    }
    TcRun r;
    r.src_acc = e.src_acc; r.c0 = e.c0; r.d12 = e.d12; r.wk = e.wk; r.count = 1;
    runs.push_back(r);
  }
  pw->num_kb = (int)st.size();
  pw->num_runs = (int)runs.size();
  pw->runs_host = runs;
  LDP_TRY(arena.alloc_t(&pw->kb_dev, st.size()));
  LDP_TRY(arena.alloc_t(&pw->runs_dev, runs.size()));
  LDP_CUDA_OK(cudaMemcpy(pw->kb_dev, st.data(), st.size() * sizeof(TcStage), cudaMemcpyHostToDevice));
  LDP_CUDA_OK(cudaMemcpy(pw->runs_dev, runs.data(), runs.size() * sizeof(TcRun), cudaMemcpyHostToDevice));
  return LDP_OK;
}

// Per-shape workspaces (activations, tensor maps, CUDA graphs) are cached in the handles; a caller that walks through many batch sizes
// would otherwise keep one of each alive for the life of the handle.  Keeps at most `cap` entries: the least recently used one is
// released (after a device synchronize - its buffers may still be in flight) before a new shape is added.
constexpr size_t LDP_MAX_CACHED_SHAPES = 8;
template <class Map>
inline void ws_evict_lru(Map& m, size_t cap = LDP_MAX_CACHED_SHAPES) {
  if (m.size() < cap) return;
  cudaDeviceSynchronize();
  auto victim = m.begin();
  for (auto it = m.begin(); it != m.end(); ++it)
    if (it->second->last_use < victim->second->last_use) victim = it;
  m.erase(victim);
}

// DDPM coefficient table [n][8] (see kernels.h DdpmStep), fp32 arithmetic in the reference's op order
// (diffusers FlaxDDPMScheduler.step / _get_variance).
void ddpm_schedule_host(int n, std::vector<float>& betas, std::vector<float>& alphas, std::vector<float>& acp);
void ddpm_coef_host(int n, std::vector<float>& coef);

}  // namespace ldp
