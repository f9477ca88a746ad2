// Shared host-side plumbing of the C-ABI library: context, streams, error capture,
// stream-ordered workspace allocation, launch accounting.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/zkb.h"

namespace zkb {

constexpr int kNumSideStreams = 6;
constexpr int kEventPool = 256;       // rotating hand-off events between a caller's stream and the bulk stream
constexpr int kSMs = 148;   // B200

struct NttDomain;           // ntt.cu
struct Groth16Stage;        // groth16.cu

}  // namespace zkb

struct zkb_ctx {
  int device = 0;
  cudaStream_t main = nullptr;
  cudaStream_t side[zkb::kNumSideStreams] = {};
  cudaEvent_t ev_fork = nullptr;
  cudaEvent_t ev_join[zkb::kNumSideStreams] = {};
  // Low-priority twins of main (index 0) and of the side streams (index i + 1) for the machine-filling
  // bucket-accumulation kernels: everything else (transforms, sorts, bucket reductions, proof assembly)
  // sits on the high-priority streams, so those small grids take SM slots as accumulation blocks retire
  // instead of queueing behind a whole accumulation kernel.
  cudaStream_t bulk[zkb::kNumSideStreams + 1] = {};
  cudaEvent_t ev_pool[zkb::kEventPool] = {};
  unsigned ev_next = 0;
  std::mutex mu;
  std::string err;
  uint64_t launches = 0;
  bool serial = false;      // zkb_set_serial: the prove path uses one stream (measurement aid)
  int sm_count = zkb::kSMs;
  std::map<int, zkb::NttDomain*> domains;   // key = curve * 64 + log_n
  zkb::Groth16Stage* stage = nullptr;
  // pinned bounce buffer for small results
  void* pinned = nullptr;
  size_t pinned_bytes = 0;
  // optional timing of the dominant kernel (bucket accumulation) with CUDA events on its stream
  bool prof_on = false;
  std::vector<cudaEvent_t> prof_events;   // start/stop pairs
  size_t prof_used = 0;
  double prof_alg_bytes = 0;
  // multi-GPU: one process per GPU, this ctx is rank `rank` of `n_ranks` (comm.cu); comm == nullptr until
  // zkb_comm_init -- the sharded entry points then fail instead of silently computing a partial result
  void* comm = nullptr;       // ncclComm_t
  int n_ranks = 1, rank = 0;
  void* gather = nullptr;     // device receive buffer of the partial-point all-gather
  size_t gather_bytes = 0;
  uint64_t collectives = 0;   // all-gathers enqueued so far (bench.py reports it)
};

namespace zkb {

inline int set_err(zkb_ctx* ctx, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (ctx) ctx->err = buf;
  return code;
}

#define ZKB_CUDA(ctx, expr)                                                                   \
  do {                                                                                        \
    cudaError_t e__ = (expr);                                                                 \
    if (e__ != cudaSuccess)                                                                   \
      return zkb::set_err(ctx, ZKB_E_CUDA, "%s failed at %s:%d: %s", #expr, __FILE__, __LINE__, \
                          cudaGetErrorString(e__));                                           \
  } while (0)

#define ZKB_TRY(expr)            \
  do {                           \
    int rc__ = (expr);           \
    if (rc__ != ZKB_OK) return rc__; \
  } while (0)

// kernel launch with accounting + error check
#define ZKB_LAUNCH(ctx, kernel, grid, block, smem, stream, ...)                               \
  do {                                                                                        \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                               \
    (ctx)->launches++;                                                                        \
    cudaError_t e__ = cudaGetLastError();                                                     \
    if (e__ != cudaSuccess)                                                                   \
      return zkb::set_err(ctx, ZKB_E_CUDA, "launch %s failed at %s:%d: %s", #kernel, __FILE__, \
                          __LINE__, cudaGetErrorString(e__));                                 \
  } while (0)

// stream-ordered scratch allocation (the default mempool keeps freed blocks cached)
struct Scratch {
  zkb_ctx* ctx;
  cudaStream_t stream;
  std::vector<void*> ptrs;
  Scratch(zkb_ctx* c, cudaStream_t s) : ctx(c), stream(s) {}
  ~Scratch() {
    for (void* p : ptrs) cudaFreeAsync(p, stream);
  }
  template <class T>
  int alloc(T** out, size_t count) {
    void* p = nullptr;
    size_t bytes = count * sizeof(T);
    if (bytes == 0) bytes = 16;
    cudaError_t e = cudaMallocAsync(&p, bytes, stream);
    if (e != cudaSuccess)
      return set_err(ctx, ZKB_E_CUDA, "cudaMallocAsync(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
    ptrs.push_back(p);
    *out = reinterpret_cast<T*>(p);
    return ZKB_OK;
  }
};

// bracket one launch of the profiled kernel; alg_bytes = algorithmic bytes that launch streams
inline void prof_begin(zkb_ctx* ctx, cudaStream_t st) {
  if (!ctx->prof_on) return;
  if (ctx->prof_used + 2 > ctx->prof_events.size()) {
    for (int i = 0; i < 2; i++) {
      cudaEvent_t e;
      if (cudaEventCreate(&e) != cudaSuccess) { ctx->prof_on = false; return; }
      ctx->prof_events.push_back(e);
    }
  }
  cudaEventRecord(ctx->prof_events[ctx->prof_used], st);
}
inline void prof_end(zkb_ctx* ctx, cudaStream_t st, double alg_bytes) {
  if (!ctx->prof_on) return;
  cudaEventRecord(ctx->prof_events[ctx->prof_used + 1], st);
  ctx->prof_used += 2;
  ctx->prof_alg_bytes += alg_bytes;
}

// Run what `enqueue` launches on the low-priority twin of `st`, ordered after the work already on `st` and before the
// work that follows on `st`.  (A rotating event is safe to reuse: cudaStreamWaitEvent captures the record
// that precedes it at call time.)
template <class Fn>
inline int on_bulk_stream(zkb_ctx* ctx, cudaStream_t st, Fn enqueue) {
  cudaStream_t bulk = st == ctx->main ? ctx->bulk[0] : nullptr;
  for (int i = 0; i < kNumSideStreams; i++)
    if (st == ctx->side[i]) bulk = ctx->bulk[i + 1];
  if (!bulk) return enqueue(st);
  cudaEvent_t e0 = ctx->ev_pool[ctx->ev_next++ % kEventPool], e1 = ctx->ev_pool[ctx->ev_next++ % kEventPool];
  ZKB_CUDA(ctx, cudaEventRecord(e0, st));
  ZKB_CUDA(ctx, cudaStreamWaitEvent(bulk, e0, 0));
  ZKB_TRY(enqueue(bulk));
  ZKB_CUDA(ctx, cudaEventRecord(e1, bulk));
  ZKB_CUDA(ctx, cudaStreamWaitEvent(st, e1, 0));
  return ZKB_OK;
}

// comm.cu: all-gather `bytes` bytes per rank from d_send into d_recv (n_ranks * bytes, rank order) on `st`;
// with n_ranks == 1 a device-to-device copy
int comm_allgather(zkb_ctx* ctx, cudaStream_t st, const void* d_send, void* d_recv, size_t bytes);
int comm_gather_buffer(zkb_ctx* ctx, size_t bytes, void** out);

inline unsigned ceil_div(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }
inline unsigned ceil_log2(size_t n) {
  unsigned l = 0;
  while ((size_t(1) << l) < n) l++;
  return l;
}

// fork side streams from main / join them back
inline int fork_streams(zkb_ctx* ctx, int n) {
  ZKB_CUDA(ctx, cudaEventRecord(ctx->ev_fork, ctx->main));
  for (int i = 0; i < n; i++) ZKB_CUDA(ctx, cudaStreamWaitEvent(ctx->side[i], ctx->ev_fork, 0));
  return ZKB_OK;
}
inline int join_streams(zkb_ctx* ctx, int n) {
  for (int i = 0; i < n; i++) {
    ZKB_CUDA(ctx, cudaEventRecord(ctx->ev_join[i], ctx->side[i]));
    ZKB_CUDA(ctx, cudaStreamWaitEvent(ctx->main, ctx->ev_join[i], 0));
  }
  return ZKB_OK;
}

}  // namespace zkb
