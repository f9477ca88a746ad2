// Multi-GPU plumbing of the C-ABI library: one process per GPU, NCCL over NVLink / NVSwitch.
//
// The prove path shards by partitioning the (scalar, base) pairs of every MSM across the ranks
// (an MSM is a plain sum over pairs: curve/src/lib.rs:38-45; the five Groth16 MSMs share only
// read-only inputs: groth16/src/prover.rs:164-190).  The only exchange step is ONE all-gather of
// a few partial group elements per rank (<= 1.2 KB), folded by a kernel in rank order on every
// rank -- NCCL has no user-defined reduction and EC addition is not ncclSum, so the "allreduce of
// bucket partials" is an all-gather plus a local fold.
//
// NCCL is bound at run time (dlopen of libnccl.so.2: the copy torch already mapped into the
// process when the host is Python, the system library for a Rust / C host), so libzkb.so keeps
// linking cudart only and single-GPU users need no NCCL at all.  The 128-byte ncclUniqueId is
// created by rank 0 (zkb_comm_unique_id) and distributed by the host's own rendezvous
// (torch.distributed broadcast in bench.py / the tests; MPI or a file for a Rust host).
#include <dlfcn.h>
#include <nccl.h>      // types only: no -lnccl

#include "common.cuh"

namespace zkb {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
  std::string err;
};

static NcclApi* nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, []() {
    // the already-mapped library first (torch bundles its own NCCL; two NCCL copies in one process must not mix)
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
      const char* e = dlerror();
      api.err = std::string("cannot load libnccl.so.2: ") + (e ? e : "unknown error");
      return;
    }
    api.handle = h;
    bool ok = true;
    auto sym = [&](const char* name) -> void* {
      void* p = dlsym(h, name);
      if (!p) { ok = false; api.err = std::string("libnccl lacks ") + name; }
      return p;
    };
    api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
    api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
    api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
    api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
    api.GetVersion = (decltype(api.GetVersion))sym("ncclGetVersion");
    if (!ok) api.handle = nullptr;
  });
  return &api;
}

#define ZKB_NCCL(ctx, api, expr)                                                                       \
  do {                                                                                                 \
    ncclResult_t r__ = (expr);                                                                         \
    if (r__ != ncclSuccess)                                                                            \
      return zkb::set_err(ctx, ZKB_E_CUDA, "%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,        \
                          (api)->GetErrorString(r__));                                                 \
  } while (0)

int comm_gather_buffer(zkb_ctx* ctx, size_t bytes, void** out) {
  if (ctx->gather_bytes < bytes) {
    if (ctx->gather) ZKB_CUDA(ctx, cudaFree(ctx->gather));
    ctx->gather = nullptr;
    ctx->gather_bytes = 0;
    size_t want = bytes < 4096 ? 4096 : bytes;
    ZKB_CUDA(ctx, cudaMalloc(&ctx->gather, want));
    ctx->gather_bytes = want;
  }
  *out = ctx->gather;
  return ZKB_OK;
}

int comm_allgather(zkb_ctx* ctx, cudaStream_t st, const void* d_send, void* d_recv, size_t bytes) {
  if (ctx->n_ranks == 1) {
    ZKB_CUDA(ctx, cudaMemcpyAsync(d_recv, d_send, bytes, cudaMemcpyDeviceToDevice, st));
    return ZKB_OK;
  }
  if (!ctx->comm) return set_err(ctx, ZKB_E_INVALID, "sharded call on rank %d of %d without zkb_comm_init", ctx->rank, ctx->n_ranks);
  NcclApi* api = nccl_api();
  ZKB_NCCL(ctx, api, api->AllGather(d_send, d_recv, bytes, ncclUint8, (ncclComm_t)ctx->comm, st));
  ctx->collectives++;
  return ZKB_OK;
}

}  // namespace zkb

using namespace zkb;

extern "C" {

int zkb_comm_unique_id(zkb_ctx* ctx, uint8_t id[ZKB_COMM_ID_BYTES]) {
  if (!ctx || !id) return ZKB_E_INVALID;
  static_assert(sizeof(ncclUniqueId) == ZKB_COMM_ID_BYTES, "ncclUniqueId size");
  std::lock_guard<std::mutex> lk(ctx->mu);
  NcclApi* api = nccl_api();
  if (!api->handle) return set_err(ctx, ZKB_E_INVALID, "%s", api->err.c_str());
  ncclUniqueId uid;
  ZKB_NCCL(ctx, api, api->GetUniqueId(&uid));
  memcpy(id, &uid, ZKB_COMM_ID_BYTES);
  return ZKB_OK;
}

int zkb_comm_init(zkb_ctx* ctx, int n_ranks, int rank, const uint8_t id[ZKB_COMM_ID_BYTES]) {
  if (!ctx || n_ranks < 1 || rank < 0 || rank >= n_ranks) return ZKB_E_INVALID;
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (ctx->comm) return set_err(ctx, ZKB_E_INVALID, "comm_init: communicator already initialised");
  ZKB_CUDA(ctx, cudaSetDevice(ctx->device));
  if (n_ranks > 1) {
    if (!id) return set_err(ctx, ZKB_E_INVALID, "comm_init: null unique id");
    NcclApi* api = nccl_api();
    if (!api->handle) return set_err(ctx, ZKB_E_INVALID, "%s", api->err.c_str());
    ncclUniqueId uid;
    memcpy(&uid, id, ZKB_COMM_ID_BYTES);
    ncclComm_t comm = nullptr;
    ZKB_NCCL(ctx, api, api->CommInitRank(&comm, n_ranks, uid, rank));
    ctx->comm = comm;
  }
  ctx->n_ranks = n_ranks;
  ctx->rank = rank;
  return ZKB_OK;
}

void zkb_comm_destroy(zkb_ctx* ctx) {
  if (!ctx) return;
  std::lock_guard<std::mutex> lk(ctx->mu);
  cudaSetDevice(ctx->device);
  if (ctx->comm) {
    cudaDeviceSynchronize();
    nccl_api()->CommDestroy((ncclComm_t)ctx->comm);
    ctx->comm = nullptr;
  }
  if (ctx->gather) cudaFree(ctx->gather);
  ctx->gather = nullptr;
  ctx->gather_bytes = 0;
  ctx->n_ranks = 1;
  ctx->rank = 0;
}

int zkb_comm_rank(zkb_ctx* ctx) { return ctx ? ctx->rank : -1; }
int zkb_comm_size(zkb_ctx* ctx) { return ctx ? ctx->n_ranks : 0; }
uint64_t zkb_comm_collectives(zkb_ctx* ctx) { return ctx ? ctx->collectives : 0; }

}  // extern "C"
