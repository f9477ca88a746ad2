// Instantiates the MSM engine for one (coordinate field, scalar field) pair and exposes it
// through the type-erased GroupOps table.  Included by exactly one .cu file per group so the
// four heavy instantiations compile in parallel.
#pragma once
#include "msm.cuh"
#include "serialize.cuh"

namespace zkb {

// out[i] = scalars[i] * base (double-and-add per thread), canonical affine output
template <class F>
__global__ void __launch_bounds__(128)
k_fixed_base_mul(const Affine<F>* __restrict__ base, const uint32_t* __restrict__ scalars, uint32_t n,
                 Affine<F>* __restrict__ out_xy, uint8_t* __restrict__ out_inf) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  XYZZ<F> p = XYZZ<F>::from_affine(ld_vec_rw(base));
  uint32_t k[kScalarLimbs];
#pragma unroll
  for (int j = 0; j < kScalarLimbs; j++) k[j] = scalars[(size_t)i * kScalarLimbs + j];
  XYZZ<F> r = XYZZ<F>::mul_limbs(p, k, kScalarLimbs);
  Affine<F> a;
  pt_to_affine(a, r);
  st_vec(&out_xy[i], a);
  out_inf[i] = r.is_inf() ? 1 : 0;
}

// ark-serialize compressed points -> affine Montgomery (serialize.cuh); one thread per point, the curve coefficient
// is derived once per block.  check_subgroup: r * P == 0 (255 doublings per point: only when the caller asks).
template <class F, class FrP>
__global__ void __launch_bounds__(128)
k_decompress(const uint8_t* __restrict__ in, uint32_t n, int check_subgroup, Affine<F>* __restrict__ out_xy,
             uint8_t* __restrict__ out_inf, uint8_t* __restrict__ out_status) {
  __shared__ F b_sh;
  if (threadIdx.x == 0) b_sh = curve_b((const F*)nullptr);
  __syncthreads();
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint8_t bytes[sizeof(F)];
  const uint4* src = reinterpret_cast<const uint4*>(in + (size_t)i * sizeof(F));      // sizeof(F) is a multiple of 16
#pragma unroll
  for (int k = 0; k < (int)(sizeof(F) / 16); k++) reinterpret_cast<uint4*>(bytes)[k] = __ldg(src + k);
  Affine<F> p;
  bool is_inf = false;
  uint8_t st = decompress_point(bytes, b_sh, p, is_inf);
  if (st == kDecompOk && !is_inf && check_subgroup) {
    uint32_t r[kScalarLimbs];
#pragma unroll
    for (int k = 0; k < kScalarLimbs; k++) r[k] = FrP::mod(k);
    XYZZ<F> q = XYZZ<F>::mul_limbs(XYZZ<F>::from_affine(p), r, kScalarLimbs);
    if (!q.is_inf()) st = kDecompNotInSubgroup;
  }
  if (st != kDecompOk) { p = Affine<F>::inf(); is_inf = true; }
  st_vec(&out_xy[i], p);
  out_inf[i] = is_inf ? 1 : 0;
  out_status[i] = st;
}

// zkb_msm_batch over MANY tiny MSMs (thousands of 2..16-term sums: the g_ic of a batch verifier, verifier.rs:27-30):
// the bucket method has nothing to amortise there, so one thread multiplies one (scalar, base) term by double-and-add
// and one thread per MSM adds its terms.
template <class F, class FrP>
__global__ void __launch_bounds__(64)
k_small_msm_batch_terms(const SmallMsmJob* __restrict__ jobs, const uint32_t* __restrict__ term_job, uint32_t n_terms,
                        int scalars_mont, XYZZ<F>* __restrict__ terms) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_terms) return;
  const SmallMsmJob job = jobs[term_job[t]];
  const uint32_t j = t - job.term0;
  XYZZ<F> r = XYZZ<F>::inf();
  if (!job.inf[job.base_offset + j]) {
    Fp<FrP> sc = ld_vec(reinterpret_cast<const Fp<FrP>*>(job.scalars) + j);
    if (scalars_mont) sc = Fp<FrP>::from_mont(sc);
    Affine<F> p = ld_vec(reinterpret_cast<const Affine<F>*>(job.table) + job.base_offset + j);   // window 0 of a table = the base
    r = XYZZ<F>::mul_limbs(XYZZ<F>::from_affine(p), sc.v, kScalarLimbs);
  }
  st_vec(&terms[t], r);
}

template <class F>
__global__ void __launch_bounds__(64)
k_small_msm_batch_sum(const SmallMsmJob* __restrict__ jobs, uint32_t n_jobs, const XYZZ<F>* __restrict__ terms,
                      XYZZ<F>* __restrict__ out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_jobs) return;
  const SmallMsmJob job = jobs[i];
  XYZZ<F> acc = XYZZ<F>::inf();
  for (uint32_t j = 0; j < job.len; j++) pt_add(acc, ld_vec(&terms[job.term0 + j]));
  st_vec(&out[i], acc);
}

template <class F, class FrP>
struct GroupImpl {
  using E = MsmEngine<F, FrP>;
  static int fixed_base_mul(zkb_ctx* ctx, cudaStream_t st, const void* d_base, const uint32_t* d_scalars, size_t n,
                            void* d_out, uint8_t* d_inf) {
    if (n == 0) return ZKB_OK;
    ZKB_LAUNCH(ctx, (k_fixed_base_mul<F>), ceil_div(n, 128), 128, 0, st, (const Affine<F>*)d_base, d_scalars,
               (uint32_t)n, (Affine<F>*)d_out, d_inf);
    return ZKB_OK;
  }
  static int fold(zkb_ctx* ctx, cudaStream_t st, const void* d_points, uint32_t count, uint32_t stride_bytes, void* d_out_pt,
                  void* d_out_affine, uint8_t* d_out_inf) {
    if (count == 0) return set_err(ctx, ZKB_E_INVALID, "fold: no points");
    ZKB_LAUNCH(ctx, (k_fold_points<F>), 1, 32, 0, st, (const XYZZ<F>*)d_points, count, stride_bytes, (XYZZ<F>*)d_out_pt,
               (Affine<F>*)d_out_affine, d_out_inf);
    return ZKB_OK;
  }
  static int decompress(zkb_ctx* ctx, cudaStream_t st, const uint8_t* d_in, size_t n, int check_subgroup, void* d_xy,
                        uint8_t* d_inf, uint8_t* d_status) {
    if (n == 0) return ZKB_OK;
    ZKB_LAUNCH(ctx, (k_decompress<F, FrP>), ceil_div(n, 128), 128, 0, st, d_in, (uint32_t)n, check_subgroup, (Affine<F>*)d_xy,
               d_inf, d_status);
    return ZKB_OK;
  }
  static int to_affine(zkb_ctx* ctx, cudaStream_t st, const void* d_points, size_t n, void* d_xy, uint8_t* d_inf) {
    if (n == 0) return ZKB_OK;
    ZKB_LAUNCH(ctx, (k_to_affine<F>), ceil_div(n, 32), 32, 0, st, (const XYZZ<F>*)d_points, (uint32_t)n, (Affine<F>*)d_xy, d_inf);
    return ZKB_OK;
  }
  static int small_msms(zkb_ctx* ctx, cudaStream_t st, const SmallMsmJob* d_jobs, uint32_t n_jobs, const uint32_t* d_term_job,
                        uint32_t n_terms, int scalars_mont, void* d_terms, void* d_out) {
    if (n_terms)
      ZKB_LAUNCH(ctx, (k_small_msm_batch_terms<F, FrP>), ceil_div(n_terms, 64), 64, 0, st, d_jobs, d_term_job, n_terms,
                 scalars_mont, (XYZZ<F>*)d_terms);
    ZKB_LAUNCH(ctx, (k_small_msm_batch_sum<F>), ceil_div(n_jobs, 64), 64, 0, st, d_jobs, n_jobs, (const XYZZ<F>*)d_terms,
               (XYZZ<F>*)d_out);
    return ZKB_OK;
  }
  static const GroupOps* ops() {
    static const GroupOps o = {sizeof(Affine<F>), sizeof(XYZZ<F>), &E::srs_build, &E::run, &E::run_to_host, &E::run_split,
                               &fixed_base_mul, &fold, &decompress, &to_affine, &small_msms};
    return &o;
  }
};

}  // namespace zkb
