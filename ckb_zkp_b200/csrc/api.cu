// C-ABI entry points of libzkb (include/zkb.h): context, SRS residency, MSM, NTT, Fr helpers.
// The Groth16 entry points live in groth16_api.cu.  Device code only -- there is no CPU
// fallback: without a usable CUDA device zkb_init fails with ZKB_E_NO_DEVICE.
#include <cstring>
#include <vector>

#include "common.cuh"
#include "groth16.cuh"
#include "msm.cuh"
#include "ntt.cuh"

namespace zkb {

const GroupOps* group_ops_bn_g1();
const GroupOps* group_ops_bn_g2();
const GroupOps* group_ops_bls_g1();
const GroupOps* group_ops_bls_g2();

const GroupOps* group_ops(int curve, int group) {
  if (curve == ZKB_BN254) return group == ZKB_G1 ? group_ops_bn_g1() : group == ZKB_G2 ? group_ops_bn_g2() : nullptr;
  if (curve == ZKB_BLS12_381) return group == ZKB_G1 ? group_ops_bls_g1() : group == ZKB_G2 ? group_ops_bls_g2() : nullptr;
  return nullptr;
}

// zkb_msm_batch takes the thread-per-term path when a call holds at least this many MSMs, each this short
constexpr size_t kSmallBatchMinJobs = 32, kSmallBatchMaxTerms = 16;

// out[i] = into_repr(in[i]) (mode 0) / from_repr(in[i]) (mode 1)
template <class FrP>
__global__ void k_fr_convert(const Fp<FrP>* in, Fp<FrP>* out, size_t n, int mode) {   // in may alias out
  using Fr = Fp<FrP>;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    Fr v = ld_vec_rw(&in[i]);
    v = mode == 0 ? Fr::from_mont(v) : Fr::to_mont(v);
    st_vec(&out[i], v);
  }
}

int fr_convert_dev(zkb_ctx* ctx, cudaStream_t st, int curve, const void* d_in, void* d_out, size_t n, int mode) {
  if (n == 0) return ZKB_OK;
  unsigned blocks = ceil_div(n, 256);
  if (blocks > (unsigned)ctx->sm_count * 16) blocks = ctx->sm_count * 16;
  if (curve == ZKB_BLS12_381)
    ZKB_LAUNCH(ctx, (k_fr_convert<BlsFr>), blocks, 256, 0, st, (const Fp<BlsFr>*)d_in, (Fp<BlsFr>*)d_out, n, mode);
  else
    ZKB_LAUNCH(ctx, (k_fr_convert<BnFr>), blocks, 256, 0, st, (const Fp<BnFr>*)d_in, (Fp<BnFr>*)d_out, n, mode);
  return ZKB_OK;
}

static bool valid_curve(int curve) { return curve == ZKB_BN254 || curve == ZKB_BLS12_381; }

}  // namespace zkb

using namespace zkb;

extern "C" {

int zkb_init(int device, zkb_ctx** out) {
  if (!out) return ZKB_E_INVALID;
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0 || device < 0 || device >= count) return ZKB_E_NO_DEVICE;
  if (cudaSetDevice(device) != cudaSuccess) return ZKB_E_NO_DEVICE;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return ZKB_E_NO_DEVICE;
  if (prop.major != 10) return ZKB_E_NO_DEVICE;   // sm_100a cubins only
  zkb_ctx* ctx = new zkb_ctx();
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  int prio_least = 0, prio_greatest = 0;
  cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest);
  bool ok = cudaStreamCreateWithPriority(&ctx->main, cudaStreamNonBlocking, prio_greatest) == cudaSuccess;
  for (int i = 0; ok && i < kNumSideStreams; i++) {
    ok = cudaStreamCreateWithPriority(&ctx->side[i], cudaStreamNonBlocking, prio_greatest) == cudaSuccess &&
         cudaEventCreateWithFlags(&ctx->ev_join[i], cudaEventDisableTiming) == cudaSuccess;
  }
  // opt-in (ZKB_BULK=1): measured on the 2^20 proof the low-priority twins change nothing (42.8 vs 42.3 ms) --
  // the five MSMs already keep the multiplier pipe ~95 % busy, see DESIGN.md 4
  const char* bulk = getenv("ZKB_BULK");
  if (ok && prio_least != prio_greatest && bulk && atoi(bulk)) {
    for (int i = 0; ok && i <= kNumSideStreams; i++)
      ok = cudaStreamCreateWithPriority(&ctx->bulk[i], cudaStreamNonBlocking, prio_least) == cudaSuccess;
    for (int i = 0; ok && i < kEventPool; i++)
      ok = cudaEventCreateWithFlags(&ctx->ev_pool[i], cudaEventDisableTiming) == cudaSuccess;
  }
  ok = ok && cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming) == cudaSuccess;
  ctx->pinned_bytes = 1 << 16;
  ok = ok && cudaMallocHost(&ctx->pinned, ctx->pinned_bytes) == cudaSuccess;
  if (ok) {
    // keep freed scratch cached in the pool: the prove path allocates the same sizes every proof
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
      uint64_t thr = UINT64_MAX;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
  }
  if (!ok) {
    zkb_destroy(ctx);
    return ZKB_E_CUDA;
  }
  *out = ctx;
  return ZKB_OK;
}

void zkb_destroy(zkb_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  zkb_comm_destroy(ctx);
  groth16_free_stage(ctx);
  ntt_free_domains(ctx);
  if (ctx->pinned) cudaFreeHost(ctx->pinned);
  for (cudaEvent_t e : ctx->prof_events) cudaEventDestroy(e);
  for (int i = 0; i < kNumSideStreams; i++) {
    if (ctx->side[i]) cudaStreamDestroy(ctx->side[i]);
    if (ctx->ev_join[i]) cudaEventDestroy(ctx->ev_join[i]);
  }
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  for (cudaEvent_t e : ctx->ev_pool) if (e) cudaEventDestroy(e);
  for (cudaStream_t b : ctx->bulk) if (b) cudaStreamDestroy(b);
  if (ctx->main) cudaStreamDestroy(ctx->main);
  delete ctx;
}

const char* zkb_last_error(zkb_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
void* zkb_stream(zkb_ctx* ctx) { return ctx ? (void*)ctx->main : nullptr; }
uint64_t zkb_launch_count(zkb_ctx* ctx) { return ctx ? ctx->launches : 0; }

int zkb_sync(zkb_ctx* ctx) {
  if (!ctx) return ZKB_E_INVALID;
  std::lock_guard<std::mutex> lk(ctx->mu);
  ZKB_CUDA(ctx, cudaSetDevice(ctx->device));
  ZKB_CUDA(ctx, cudaStreamSynchronize(ctx->main));
  return ZKB_OK;
}

int zkb_set_serial(zkb_ctx* ctx, int on) {
  if (!ctx) return ZKB_E_INVALID;
  std::lock_guard<std::mutex> lk(ctx->mu);
  ctx->serial = on != 0;
  return ZKB_OK;
}

// ---- kernel timing (bench.py's roofline figure) ------------------------------------------------
int zkb_prof_enable(zkb_ctx* ctx, int on) {
  if (!ctx) return ZKB_E_INVALID;
  std::lock_guard<std::mutex> lk(ctx->mu);
  ctx->prof_on = on != 0;
  ctx->prof_used = 0;
  ctx->prof_alg_bytes = 0;
  return ZKB_OK;
}
int zkb_prof_read(zkb_ctx* ctx, double* ms_total, uint64_t* launches, double* alg_bytes_total) {
  if (!ctx || !ms_total || !launches || !alg_bytes_total) return ZKB_E_INVALID;
  std::lock_guard<std::mutex> lk(ctx->mu);
  ZKB_CUDA(ctx, cudaSetDevice(ctx->device));
  ZKB_CUDA(ctx, cudaDeviceSynchronize());
  double ms = 0;
  for (size_t i = 0; i + 1 < ctx->prof_used; i += 2) {
    float t = 0;
    ZKB_CUDA(ctx, cudaEventElapsedTime(&t, ctx->prof_events[i], ctx->prof_events[i + 1]));
    ms += t;
  }
  *ms_total = ms;
  *launches = ctx->prof_used / 2;
  *alg_bytes_total = ctx->prof_alg_bytes;
  return ZKB_OK;
}

// ---- SRS -----------------------------------------------------------------------------------
int zkb_srs_upload(zkb_ctx* ctx, int curve, int group, const uint64_t* xy_mont, const uint8_t* inf, size_t n,
                   unsigned flags, zkb_srs** out) {
  if (!ctx || !out) return ZKB_E_INVALID;
  std::lock_guard<std::mutex> lk(ctx->mu);
  *out = nullptr;
  const GroupOps* ops = group_ops(curve, group);
  if (!ops) return set_err(ctx, ZKB_E_INVALID, "srs_upload: unknown curve %d / group %d", curve, group);
  if (n && (!xy_mont || !inf)) return set_err(ctx, ZKB_E_INVALID, "srs_upload: null bases");
  ZKB_CUDA(ctx, cudaSetDevice(ctx->device));
  zkb_srs* srs = new zkb_srs();
  srs->ctx = ctx; srs->curve = curve; srs->group = group; srs->n = n;
  srs->table = nullptr; srs->inf = nullptr;
  srs->global_lo = 0; srs->global_n = n;
  int rc = ops->srs_build(ctx, srs, xy_mont, inf, flags);
  if (rc != ZKB_OK) {
    zkb_srs_free(srs);
    return rc;
  }
  *out = srs;
  return ZKB_OK;
}

void zkb_srs_free(zkb_srs* srs) {
  if (!srs) return;
  cudaSetDevice(srs->ctx->device);
  if (srs->table) cudaFree(srs->table);
  if (srs->inf) cudaFree(srs->inf);
  delete srs;
}

size_t zkb_srs_len(const zkb_srs* srs) { return srs ? srs->n : 0; }

int zkb_srs_upload_shard(zkb_ctx* ctx, int curve, int group, const uint64_t* xy_mont_local, const uint8_t* inf_local,
                         size_t n_local, size_t global_lo, size_t global_n, unsigned flags, zkb_srs** out) {
  if (!ctx || !out) return ZKB_E_INVALID;
  if (global_lo > global_n || n_local > global_n - global_lo) {
    std::lock_guard<std::mutex> lk(ctx->mu);
    return set_err(ctx, ZKB_E_INVALID, "srs_upload_shard: [%zu, %zu) outside the logical SRS of %zu bases", global_lo,
                   global_lo + n_local, global_n);
  }
  ZKB_TRY(zkb_srs_upload(ctx, curve, group, xy_mont_local, inf_local, n_local, flags, out));
  (*out)->global_lo = global_lo;
  (*out)->global_n = global_n;
  return ZKB_OK;
}

// ---- MSM -----------------------------------------------------------------------------------
static int msm_host(zkb_ctx* ctx, const zkb_srs* srs, size_t base_offset, const uint64_t* scalars, size_t n,
                    int mont, uint64_t* out_xy, uint8_t* out_inf) {
  if (!ctx || !srs || !out_xy || !out_inf || (n && !scalars)) return ZKB_E_INVALID;
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (srs->ctx != ctx) return set_err(ctx, ZKB_E_INVALID, "msm: srs belongs to another context");
  ZKB_CUDA(ctx, cudaSetDevice(ctx->device));
  // ark's multi_scalar_mul zips bases and scalars: the shorter one wins (groth16/src/prover.rs:187)
  size_t avail = base_offset <= srs->n ? srs->n - base_offset : 0;
  if (n > avail) n = avail;
  const GroupOps* ops = group_ops(srs->curve, srs->group);
  Scratch ws(ctx, ctx->main);
  uint32_t* d_scalars;
  ZKB_TRY(ws.alloc(&d_scalars, n * 8));
  if (n) ZKB_CUDA(ctx, cudaMemcpyAsync(d_scalars, scalars, n * 32, cudaMemcpyDefault, ctx->main));
  return ops->msm_to_host(ctx, srs, base_offset, d_scalars, n, mont, out_xy, out_inf);
}

int zkb_msm(zkb_ctx* ctx, const zkb_srs* srs, size_t base_offset, const uint64_t* scalars_canonical, size_t n,
            uint64_t* out_xy, uint8_t* out_inf) {
  return msm_host(ctx, srs, base_offset, scalars_canonical, n, 0, out_xy, out_inf);
}
int zkb_msm_mont(zkb_ctx* ctx, const zkb_srs* srs, size_t base_offset, const uint64_t* scalars_mont, size_t n,
                 uint64_t* out_xy, uint8_t* out_inf) {
  return msm_host(ctx, srs, base_offset, scalars_mont, n, 1, out_xy, out_inf);
}
int zkb_msm_dev(zkb_ctx* ctx, const zkb_srs* srs, size_t base_offset, const void* d_scalars_canonical, size_t n,
                uint64_t* out_xy, uint8_t* out_inf) {
  if (!ctx || !srs || !out_xy || !out_inf || (n && !d_scalars_canonical)) return ZKB_E_INVALID;
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (srs->ctx != ctx) return set_err(ctx, ZKB_E_INVALID, "msm: srs belongs to another context");
  ZKB_CUDA(ctx, cudaSetDevice(ctx->device));
  size_t avail = base_offset <= srs->n ? srs->n - base_offset : 0;
  if (n > avail) n = avail;
  return group_ops(srs->curve, srs->group)
      ->msm_to_host(ctx, srs, base_offset, (const uint32_t*)d_scalars_canonical, n, 0, out_xy, out_inf);
}

// k MSMs on the side streams, one conversion + copy of the k results at the end
int zkb_msm_batch(zkb_ctx* ctx, size_t k, const zkb_srs* const* srs, const size_t* base_offsets,
                  const uint64_t* const* scalars, const size_t* n, int scalars_mont, uint64_t* out_xy, uint8_t* out_inf) {
  if (!ctx || (k && (!srs || !base_offsets || !scalars || !n || !out_xy || !out_inf))) return ZKB_E_INVALID;
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (k == 0) return ZKB_OK;
  if (k > (size_t(1) << 20)) return set_err(ctx, ZKB_E_INVALID, "msm_batch: too many MSMs in one call");
  bool small = k >= kSmallBatchMinJobs;
  for (size_t i = 0; i < k && small; i++) small = n[i] <= kSmallBatchMaxTerms;
  if (!small && k > 4096) return set_err(ctx, ZKB_E_INVALID, "msm_batch: more than 4096 MSMs that are not all short");
  for (size_t i = 0; i < k; i++) {
    if (!srs[i] || (n[i] && !scalars[i])) return set_err(ctx, ZKB_E_INVALID, "msm_batch: null argument for MSM %zu", i);
    if (srs[i]->ctx != ctx) return set_err(ctx, ZKB_E_INVALID, "msm_batch: srs belongs to another context");
    if (srs[i]->curve != srs[0]->curve || srs[i]->group != srs[0]->group)
      return set_err(ctx, ZKB_E_INVALID, "msm_batch: all MSMs of one call must share curve and group");
  }
  ZKB_CUDA(ctx, cudaSetDevice(ctx->device));
  const GroupOps* ops = group_ops(srs[0]->curve, srs[0]->group);
  cudaStream_t st = ctx->main;
  Scratch ws(ctx, st);
  uint8_t *d_pts, *d_aff, *d_inf;
  ZKB_TRY(ws.alloc(&d_pts, k * ops->xyzz_bytes));
  ZKB_TRY(ws.alloc(&d_aff, k * ops->affine_bytes));
  ZKB_TRY(ws.alloc(&d_inf, k));
  if (small) {
    // thousands of 2..16-term MSMs (a batch verifier's g_ic): scalars gathered into ONE staging buffer and one copy,
    // one thread per term, one per sum (group_impl.cuh) -- the bucket pipeline has nothing to amortise at this size
    std::vector<SmallMsmJob> jobs(k);
    std::vector<uint32_t> term_job;
    std::vector<uint64_t> stage;
    uint32_t n_terms = 0;
    for (size_t i = 0; i < k; i++) {
      size_t avail = base_offsets[i] <= srs[i]->n ? srs[i]->n - base_offsets[i] : 0;
      uint32_t len = (uint32_t)(n[i] < avail ? n[i] : avail);
      jobs[i] = SmallMsmJob{srs[i]->table, srs[i]->inf, nullptr, (uint32_t)base_offsets[i], len, n_terms, 0};
      n_terms += len;
    }
    term_job.resize(n_terms);
    stage.resize((size_t)n_terms * 4);
    uint32_t* d_scalars;
    ZKB_TRY(ws.alloc(&d_scalars, (size_t)n_terms * 8 + 8));
    std::vector<char> on_dev(k, 0);
    bool any_host = false;
    for (size_t i = 0; i < k; i++) {
      const uint32_t len = jobs[i].len, t0 = jobs[i].term0;
      for (uint32_t j = 0; j < len; j++) term_job[t0 + j] = (uint32_t)i;
      jobs[i].scalars = d_scalars + (size_t)t0 * 8;
      if (!len) continue;
      cudaPointerAttributes attr;
      on_dev[i] = cudaPointerGetAttributes(&attr, scalars[i]) == cudaSuccess &&
                  (attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged);
      cudaGetLastError();
      if (!on_dev[i]) { memcpy(stage.data() + (size_t)t0 * 4, scalars[i], (size_t)len * 32); any_host = true; }
    }
    // host scalars travel in one copy of the staging buffer; device-resident ones are then copied over their slots
    if (any_host && n_terms) ZKB_CUDA(ctx, cudaMemcpyAsync(d_scalars, stage.data(), (size_t)n_terms * 32, cudaMemcpyDefault, st));
    for (size_t i = 0; i < k; i++)
      if (on_dev[i])
        ZKB_CUDA(ctx, cudaMemcpyAsync(d_scalars + (size_t)jobs[i].term0 * 8, scalars[i], (size_t)jobs[i].len * 32, cudaMemcpyDefault, st));
    SmallMsmJob* d_jobs;
    uint32_t* d_term_job;
    uint8_t* d_terms;
    ZKB_TRY(ws.alloc(&d_jobs, k));
    ZKB_TRY(ws.alloc(&d_term_job, (size_t)n_terms + 1));
    ZKB_TRY(ws.alloc(&d_terms, (size_t)n_terms * ops->xyzz_bytes + 16));
    ZKB_CUDA(ctx, cudaMemcpyAsync(d_jobs, jobs.data(), k * sizeof(SmallMsmJob), cudaMemcpyDefault, st));
    if (n_terms) ZKB_CUDA(ctx, cudaMemcpyAsync(d_term_job, term_job.data(), (size_t)n_terms * 4, cudaMemcpyDefault, st));
    ZKB_TRY(ops->small_msms(ctx, st, d_jobs, (uint32_t)k, d_term_job, n_terms, scalars_mont, d_terms, d_pts));
    ZKB_TRY(ops->to_affine(ctx, st, d_pts, k, d_aff, d_inf));
    ZKB_CUDA(ctx, cudaMemcpyAsync(out_xy, d_aff, k * ops->affine_bytes, cudaMemcpyDefault, st));
    ZKB_CUDA(ctx, cudaMemcpyAsync(out_inf, d_inf, k, cudaMemcpyDefault, st));
    ZKB_CUDA(ctx, cudaStreamSynchronize(st));         // also keeps jobs / term_job / stage alive until the copies ran
    return ZKB_OK;
  }
  // scalar copies on the main stream (stream-ordered scratch), the MSMs fan out over the side streams
  std::vector<uint32_t*> d_sc(k, nullptr);
  std::vector<size_t> len(k, 0);
  for (size_t i = 0; i < k; i++) {
    size_t avail = base_offsets[i] <= srs[i]->n ? srs[i]->n - base_offsets[i] : 0;
    len[i] = n[i] < avail ? n[i] : avail;
    ZKB_TRY(ws.alloc(&d_sc[i], len[i] * 8));
    if (len[i]) ZKB_CUDA(ctx, cudaMemcpyAsync(d_sc[i], scalars[i], len[i] * 32, cudaMemcpyDefault, st));
  }
  const int lanes = ctx->serial ? 0 : kNumSideStreams;
  if (lanes) ZKB_TRY(fork_streams(ctx, lanes));
  for (size_t i = 0; i < k; i++) {
    cudaStream_t s_i = lanes ? ctx->side[i % lanes] : st;
    ZKB_TRY(ops->msm_run(ctx, s_i, srs[i], base_offsets[i], d_sc[i], len[i], scalars_mont, d_pts + i * ops->xyzz_bytes));
  }
  if (lanes) ZKB_TRY(join_streams(ctx, lanes));
  ZKB_TRY(ops->to_affine(ctx, st, d_pts, k, d_aff, d_inf));
  ZKB_CUDA(ctx, cudaMemcpyAsync(out_xy, d_aff, k * ops->affine_bytes, cudaMemcpyDefault, st));
  ZKB_CUDA(ctx, cudaMemcpyAsync(out_inf, d_inf, k, cudaMemcpyDefault, st));
  ZKB_CUDA(ctx, cudaStreamSynchronize(st));
  return ZKB_OK;
}

// ---- sharded MSM (one process per GPU) ---------------------------------------------------------
// This rank's partial of multi_scalar_mul(&bases[base_offset .. base_offset + n), &scalars[..n)) over the LOGICAL
// SRS: the pairs whose base lies in this rank's shard [global_lo, global_lo + srs->n).  `scalars` points at the full
// n-element array (host or device); only the local slice is read.  Result: one XYZZ point in d_partial.
static int msm_shard_partial(zkb_ctx* ctx, cudaStream_t st, Scratch& ws, const zkb_srs* srs, size_t base_offset,
                             const uint64_t* scalars, size_t n, int mont, void* d_partial) {
  size_t avail = base_offset <= srs->global_n ? srs->global_n - base_offset : 0;
  if (n > avail) n = avail;                                  // zip semantics of multi_scalar_mul
  const size_t lo = srs->global_lo, hi = srs->global_lo + srs->n;
  size_t g0 = base_offset > lo ? base_offset : lo;
  size_t g1 = base_offset + n < hi ? base_offset + n : hi;
  size_t n_loc = g1 > g0 ? g1 - g0 : 0;
  const GroupOps* ops = group_ops(srs->curve, srs->group);
  uint32_t* d_scalars;
  ZKB_TRY(ws.alloc(&d_scalars, n_loc * 8));
  if (n_loc)
    ZKB_CUDA(ctx, cudaMemcpyAsync(d_scalars, scalars + 4 * (g0 - base_offset), n_loc * 32, cudaMemcpyDefault, st));
  return ops->msm_run(ctx, st, srs, n_loc ? g0 - lo : 0, d_scalars, n_loc, mont, d_partial);
}

// partial (device) -> all-gather -> fold -> host
static int msm_gather_fold(zkb_ctx* ctx, cudaStream_t st, Scratch& ws, const GroupOps* ops, const void* d_partial,
                           uint64_t* out_xy, uint8_t* out_inf) {
  void* d_all;
  ZKB_TRY(comm_gather_buffer(ctx, ops->xyzz_bytes * (size_t)ctx->n_ranks, &d_all));
  ZKB_TRY(comm_allgather(ctx, st, d_partial, d_all, ops->xyzz_bytes));
  uint8_t *d_aff, *d_inf;
  ZKB_TRY(ws.alloc(&d_aff, ops->affine_bytes));
  ZKB_TRY(ws.alloc(&d_inf, 16));
  ZKB_TRY(ops->fold(ctx, st, d_all, (uint32_t)ctx->n_ranks, (uint32_t)ops->xyzz_bytes, nullptr, d_aff, d_inf));
  ZKB_CUDA(ctx, cudaMemcpyAsync(out_xy, d_aff, ops->affine_bytes, cudaMemcpyDeviceToHost, st));
  ZKB_CUDA(ctx, cudaMemcpyAsync(out_inf, d_inf, 1, cudaMemcpyDeviceToHost, st));
  ZKB_CUDA(ctx, cudaStreamSynchronize(st));
  return ZKB_OK;
}

int zkb_msm_partial(zkb_ctx* ctx, const zkb_srs* srs, size_t base_offset, const uint64_t* scalars, size_t n,
                    int scalars_mont, void* partial_out) {
  if (!ctx || !srs || !partial_out || (n && !scalars)) return ZKB_E_INVALID;
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (srs->ctx != ctx) return set_err(ctx, ZKB_E_INVALID, "msm: srs belongs to another context");
  ZKB_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->main;
  const GroupOps* ops = group_ops(srs->curve, srs->group);
  Scratch ws(ctx, st);
  uint8_t* d_partial;
  ZKB_TRY(ws.alloc(&d_partial, ops->xyzz_bytes));
  ZKB_TRY(msm_shard_partial(ctx, st, ws, srs, base_offset, scalars, n, scalars_mont, d_partial));
  ZKB_CUDA(ctx, cudaMemcpyAsync(partial_out, d_partial, ops->xyzz_bytes, cudaMemcpyDeviceToHost, st));
  ZKB_CUDA(ctx, cudaStreamSynchronize(st));
  return ZKB_OK;
}

size_t zkb_partial_bytes(int curve, int group) {
  const GroupOps* ops = group_ops(curve, group);
  return ops ? ops->xyzz_bytes : 0;
}

int zkb_msm_fold(zkb_ctx* ctx, int curve, int group, const void* partials, size_t count, uint64_t* out_xy, uint8_t* out_inf) {
  if (!ctx || !partials || !out_xy || !out_inf || count == 0 || count > 4096) return ZKB_E_INVALID;
  std::lock_guard<std::mutex> lk(ctx->mu);
  const GroupOps* ops = group_ops(curve, group);
  if (!ops) return set_err(ctx, ZKB_E_INVALID, "msm_fold: unknown curve %d / group %d", curve, group);
  ZKB_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->main;
  Scratch ws(ctx, st);
  uint8_t *d_all, *d_aff, *d_inf;
  ZKB_TRY(ws.alloc(&d_all, ops->xyzz_bytes * count));
  ZKB_TRY(ws.alloc(&d_aff, ops->affine_bytes));
  ZKB_TRY(ws.alloc(&d_inf, 16));
  ZKB_CUDA(ctx, cudaMemcpyAsync(d_all, partials, ops->xyzz_bytes * count, cudaMemcpyDefault, st));
  ZKB_TRY(ops->fold(ctx, st, d_all, (uint32_t)count, (uint32_t)ops->xyzz_bytes, nullptr, d_aff, d_inf));
  ZKB_CUDA(ctx, cudaMemcpyAsync(out_xy, d_aff, ops->affine_bytes, cudaMemcpyDeviceToHost, st));
  ZKB_CUDA(ctx, cudaMemcpyAsync(out_inf, d_inf, 1, cudaMemcpyDeviceToHost, st));
  ZKB_CUDA(ctx, cudaStreamSynchronize(st));
  return ZKB_OK;
}

int zkb_msm_sharded(zkb_ctx* ctx, const zkb_srs* srs, size_t base_offset, const uint64_t* scalars, size_t n,
                    int scalars_mont, uint64_t* out_xy, uint8_t* out_inf) {
  if (!ctx || !srs || !out_xy || !out_inf || (n && !scalars)) return ZKB_E_INVALID;
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (srs->ctx != ctx) return set_err(ctx, ZKB_E_INVALID, "msm: srs belongs to another context");
  ZKB_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->main;
  const GroupOps* ops = group_ops(srs->curve, srs->group);
  Scratch ws(ctx, st);
  uint8_t* d_partial;
  ZKB_TRY(ws.alloc(&d_partial, ops->xyzz_bytes));
  ZKB_TRY(msm_shard_partial(ctx, st, ws, srs, base_offset, scalars, n, scalars_mont, d_partial));
  return msm_gather_fold(ctx, st, ws, ops, d_partial, out_xy, out_inf);
}

int zkb_msm_sharded_local(zkb_ctx* ctx, const zkb_srs* srs, const void* d_scalars_local, size_t n_local,
                          uint64_t* out_xy, uint8_t* out_inf) {
  if (!ctx || !srs || !out_xy || !out_inf || (n_local && !d_scalars_local)) return ZKB_E_INVALID;
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (srs->ctx != ctx) return set_err(ctx, ZKB_E_INVALID, "msm: srs belongs to another context");
  ZKB_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->main;
  const GroupOps* ops = group_ops(srs->curve, srs->group);
  if (n_local > srs->n) n_local = srs->n;
  Scratch ws(ctx, st);
  uint8_t* d_partial;
  ZKB_TRY(ws.alloc(&d_partial, ops->xyzz_bytes));
  ZKB_TRY(ops->msm_run(ctx, st, srs, 0, (const uint32_t*)d_scalars_local, n_local, 0, d_partial));
  return msm_gather_fold(ctx, st, ws, ops, d_partial, out_xy, out_inf);
}

// ---- NTT -----------------------------------------------------------------------------------
int zkb_ntt_dev(zkb_ctx* ctx, int curve, void* d_data_mont, unsigned log_n, unsigned flags) {
  if (!ctx || !d_data_mont) return ZKB_E_INVALID;
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (!valid_curve(curve)) return set_err(ctx, ZKB_E_INVALID, "ntt: unknown curve %d", curve);
  ZKB_CUDA(ctx, cudaSetDevice(ctx->device));
  NttDomain* dom;
  ZKB_TRY(ntt_get_domain(ctx, curve, log_n, &dom));
  Scratch ws(ctx, ctx->main);
  uint32_t* scratch;
  ZKB_TRY(ws.alloc(&scratch, dom->n * 8));
  return ntt_run(ctx, ctx->main, dom, d_data_mont, scratch, flags);
}

int zkb_ntt(zkb_ctx* ctx, int curve, uint64_t* data_mont, unsigned log_n, unsigned flags) {
  if (!ctx || !data_mont) return ZKB_E_INVALID;
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (!valid_curve(curve)) return set_err(ctx, ZKB_E_INVALID, "ntt: unknown curve %d", curve);
  ZKB_CUDA(ctx, cudaSetDevice(ctx->device));
  NttDomain* dom;
  ZKB_TRY(ntt_get_domain(ctx, curve, log_n, &dom));
  Scratch ws(ctx, ctx->main);
  uint32_t *data, *scratch;
  ZKB_TRY(ws.alloc(&data, dom->n * 8));
  ZKB_TRY(ws.alloc(&scratch, dom->n * 8));
  ZKB_CUDA(ctx, cudaMemcpyAsync(data, data_mont, dom->n * 32, cudaMemcpyDefault, ctx->main));
  ZKB_TRY(ntt_run(ctx, ctx->main, dom, data, scratch, flags));
  ZKB_CUDA(ctx, cudaMemcpyAsync(data_mont, data, dom->n * 32, cudaMemcpyDefault, ctx->main));
  ZKB_CUDA(ctx, cudaStreamSynchronize(ctx->main));
  return ZKB_OK;
}

// ---- fixed-base multiplication ---------------------------------------------------------------
int zkb_fixed_base_mul(zkb_ctx* ctx, int curve, int group, const uint64_t* base_xy_mont,
                       const uint64_t* scalars_canonical, size_t n, uint64_t* out_xy, uint8_t* out_inf) {
  if (!ctx || !base_xy_mont || (n && (!scalars_canonical || !out_xy || !out_inf))) return ZKB_E_INVALID;
  std::lock_guard<std::mutex> lk(ctx->mu);
  const GroupOps* ops = group_ops(curve, group);
  if (!ops) return set_err(ctx, ZKB_E_INVALID, "fixed_base_mul: unknown curve %d / group %d", curve, group);
  ZKB_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->main;
  Scratch ws(ctx, st);
  uint8_t *d_base, *d_out, *d_inf;
  uint32_t* d_sc;
  ZKB_TRY(ws.alloc(&d_base, ops->affine_bytes));
  ZKB_TRY(ws.alloc(&d_sc, n * 8));
  ZKB_TRY(ws.alloc(&d_out, n * ops->affine_bytes));
  ZKB_TRY(ws.alloc(&d_inf, n));
  ZKB_CUDA(ctx, cudaMemcpyAsync(d_base, base_xy_mont, ops->affine_bytes, cudaMemcpyDefault, st));
  if (n) {
    ZKB_CUDA(ctx, cudaMemcpyAsync(d_sc, scalars_canonical, n * 32, cudaMemcpyDefault, st));
    ZKB_TRY(ops->fixed_base_mul(ctx, st, d_base, d_sc, n, d_out, d_inf));
    ZKB_CUDA(ctx, cudaMemcpyAsync(out_xy, d_out, n * ops->affine_bytes, cudaMemcpyDefault, st));
    ZKB_CUDA(ctx, cudaMemcpyAsync(out_inf, d_inf, n, cudaMemcpyDefault, st));
  }
  ZKB_CUDA(ctx, cudaStreamSynchronize(st));
  return ZKB_OK;
}

// ---- key-file ingestion ------------------------------------------------------------------------
int zkb_points_decompress(zkb_ctx* ctx, int curve, int group, const uint8_t* compressed, size_t n, unsigned flags,
                          uint64_t* out_xy_mont, uint8_t* out_inf, uint8_t* out_status) {
  if (!ctx || (n && (!compressed || !out_xy_mont || !out_inf || !out_status))) return ZKB_E_INVALID;
  std::lock_guard<std::mutex> lk(ctx->mu);
  const GroupOps* ops = group_ops(curve, group);
  if (!ops) return set_err(ctx, ZKB_E_INVALID, "points_decompress: unknown curve %d / group %d", curve, group);
  if (n >= (size_t(1) << 31)) return set_err(ctx, ZKB_E_INVALID, "points_decompress: too many points");
  if (n == 0) return ZKB_OK;
  ZKB_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->main;
  const size_t in_bytes = ops->affine_bytes / 2;          // x only
  Scratch ws(ctx, st);
  uint8_t *d_in, *d_xy, *d_inf, *d_status;
  ZKB_TRY(ws.alloc(&d_in, n * in_bytes));
  ZKB_TRY(ws.alloc(&d_xy, n * ops->affine_bytes));
  ZKB_TRY(ws.alloc(&d_inf, n));
  ZKB_TRY(ws.alloc(&d_status, n));
  ZKB_CUDA(ctx, cudaMemcpyAsync(d_in, compressed, n * in_bytes, cudaMemcpyDefault, st));
  ZKB_TRY(ops->decompress(ctx, st, d_in, n, (flags & ZKB_DECOMPRESS_CHECK_SUBGROUP) ? 1 : 0, d_xy, d_inf, d_status));
  ZKB_CUDA(ctx, cudaMemcpyAsync(out_xy_mont, d_xy, n * ops->affine_bytes, cudaMemcpyDefault, st));
  ZKB_CUDA(ctx, cudaMemcpyAsync(out_inf, d_inf, n, cudaMemcpyDefault, st));
  ZKB_CUDA(ctx, cudaMemcpyAsync(out_status, d_status, n, cudaMemcpyDefault, st));
  ZKB_CUDA(ctx, cudaStreamSynchronize(st));
  return ZKB_OK;
}

// ---- Fr helpers ------------------------------------------------------------------------------
int zkb_fr_convert(zkb_ctx* ctx, int curve, const uint64_t* in, uint64_t* out, size_t n, int mode) {
  if (!ctx || (n && (!in || !out)) || (mode != 0 && mode != 1)) return ZKB_E_INVALID;
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (!valid_curve(curve)) return set_err(ctx, ZKB_E_INVALID, "fr_convert: unknown curve %d", curve);
  if (n == 0) return ZKB_OK;
  ZKB_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->main;
  Scratch ws(ctx, st);
  uint32_t* d;
  ZKB_TRY(ws.alloc(&d, n * 8));
  ZKB_CUDA(ctx, cudaMemcpyAsync(d, in, n * 32, cudaMemcpyDefault, st));
  ZKB_TRY(fr_convert_dev(ctx, st, curve, d, d, n, mode));
  ZKB_CUDA(ctx, cudaMemcpyAsync(out, d, n * 32, cudaMemcpyDefault, st));
  ZKB_CUDA(ctx, cudaStreamSynchronize(st));
  return ZKB_OK;
}

}  // extern "C"
