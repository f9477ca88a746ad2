//! zkb-sys: Rust binding of `libzkb.so`, the B200 proving backend for ckb-zkp's Groth16 / Marlin prove path.
//!
//! `ffi` is generated from `include/zkb.h` (tools/gen_zkb_sys.py).  This file adds the thin safe layer the patched
//! crates use (bindings/patches/*.diff): a `Context` per GPU, bases resident in HBM (`Srs`), a resident Groth16
//! proving key (`ProvingKey`) and the error mapping onto `zkp_r1cs::SynthesisError`.
//!
//! Marshalling rules (ark 0.2 types are not `#[repr(C)]`):
//!   * `Fp256(BigInteger256([u64; 4]))` / `Fp384(BigInteger384([u64; 6]))` are 32 / 48 contiguous bytes holding the
//!     MONTGOMERY limbs -- exactly the ABI's "mont" layout, so `&[Fr]` is passed as `as_ptr() as *const u64` after the
//!     `size_of` assertion in `fr_slice_as_words`; `into_repr()` vectors are the ABI's "canonical" layout.
//!   * `GroupAffine { x, y, infinity }` is repacked ONCE per key into x || y limbs (Fq2 = c0 || c1) plus one infinity
//!     byte per point (`pack_g1` / `pack_g2` in the patched crates); results come back in the same layout.
//!
//! NOTE: no Rust toolchain exists in the image this repository is built in; the crate is shipped as source and its
//! declarations are kept in step with the header by tests/test_bindings.py.  The same ABI is exercised end to end
//! through ctypes (ckb_zkp_b200/_lib.py) and from C (tests/c/abi_dlopen.c).
#![allow(non_camel_case_types, non_snake_case, clippy::too_many_arguments, clippy::missing_safety_doc)]

pub mod ffi;
pub use ffi::*;

use std::ffi::CStr;
use std::os::raw::c_int;
use std::ptr;

/// Error of a libzkb call: the negative ZKB_E_* code and the library's message.
#[derive(Debug, Clone)]
pub struct ZkbError {
    pub code: c_int,
    pub message: String,
}

impl ZkbError {
    /// `ZKB_E_TOO_LARGE` is `EvaluationDomain::new(..) == None`, i.e. `SynthesisError::PolynomialDegreeTooLarge`
    /// (groth16/src/r1cs_to_qap.rs:123-125); everything else has no counterpart in the reference and is surfaced as-is.
    pub fn is_degree_too_large(&self) -> bool {
        self.code == ZKB_E_TOO_LARGE
    }
}

impl std::fmt::Display for ZkbError {
    fn fmt(&self, f: &mut std::fmt::Formatter<'_>) -> std::fmt::Result {
        write!(f, "zkb error {}: {}", self.code, self.message)
    }
}
impl std::error::Error for ZkbError {}

pub type Result<T> = std::result::Result<T, ZkbError>;

/// Which curve / group an array of points belongs to.
#[derive(Copy, Clone, Debug, PartialEq, Eq)]
pub enum Curve {
    Bn254 = 0,
    Bls12_381 = 1,
}
impl Curve {
    /// u64 limbs of one base-field element
    pub fn fq_limbs(self) -> usize {
        match self {
            Curve::Bn254 => 4,
            Curve::Bls12_381 => 6,
        }
    }
    pub fn g1_words(self) -> usize {
        2 * self.fq_limbs()
    }
    pub fn g2_words(self) -> usize {
        4 * self.fq_limbs()
    }
}

/// One context per GPU per host thread (the library serialises calls on a context internally).
pub struct Context {
    raw: *mut zkb_ctx,
}
unsafe impl Send for Context {}

impl Context {
    /// Fails with `ZKB_E_NO_DEVICE` when no sm_100 device is visible: there is no CPU fallback.
    pub fn new(device: i32) -> Result<Self> {
        let mut raw = ptr::null_mut();
        let rc = unsafe { zkb_init(device, &mut raw) };
        if rc != ZKB_OK {
            return Err(ZkbError { code: rc, message: format!("zkb_init(device = {}) failed: no usable B200", device) });
        }
        Ok(Context { raw })
    }
    pub fn as_ptr(&self) -> *mut zkb_ctx {
        self.raw
    }
    pub(crate) fn check(&self, rc: c_int) -> Result<()> {
        if rc == ZKB_OK {
            return Ok(());
        }
        let message = unsafe { CStr::from_ptr(zkb_last_error(self.raw)) }.to_string_lossy().into_owned();
        Err(ZkbError { code: rc, message })
    }

    /// Bases resident in HBM: the `&[G::Affine]` of `VariableBaseMSM::multi_scalar_mul`
    /// (groth16/src/prover.rs:187,190,220; marlin/src/pc/kzg10.rs:109,118,137,146; curve/src/lib.rs:44).
    pub fn srs_upload(&self, curve: Curve, g2: bool, xy_mont: &[u64], inf: &[u8], precompute: bool) -> Result<Srs<'_>> {
        let words = if g2 { curve.g2_words() } else { curve.g1_words() };
        assert_eq!(xy_mont.len(), inf.len() * words, "x || y limbs per point");
        let mut raw = ptr::null_mut();
        let flags = if precompute { ZKB_SRS_PRECOMPUTE } else { 0 };
        self.check(unsafe {
            zkb_srs_upload(self.raw, curve as c_int, if g2 { ZKB_G2 } else { ZKB_G1 }, xy_mont.as_ptr(), inf.as_ptr(), inf.len(),
                           flags, &mut raw)
        })?;
        Ok(Srs { ctx: self, raw, words })
    }

    /// `EvaluationDomain::{fft, ifft, coset_fft, coset_ifft}_in_place` (groth16/src/r1cs_to_qap.rs:144-169) on a slice of
    /// Montgomery field elements viewed as words.
    pub fn ntt(&self, curve: Curve, data_mont: &mut [u64], inverse: bool, coset: bool) -> Result<()> {
        let n = data_mont.len() / 4;
        assert!(n.is_power_of_two() && data_mont.len() == 4 * n);
        let flags = (if inverse { ZKB_NTT_INVERSE } else { 0 }) | (if coset { ZKB_NTT_COSET } else { 0 });
        self.check(unsafe { zkb_ntt(self.raw, curve as c_int, data_mont.as_mut_ptr(), n.trailing_zeros(), flags) })
    }

    /// k independent MSMs in one call (`zkb_msm_batch`): the commitment loops of `PC::commit`
    /// (marlin/src/pc/mod.rs:42-69) and of the `Curve::vartime_multiscalar_mul` consumers
    /// (spartan/src/commitments.rs:42-56).  jobs[i] = (bases, base_offset, scalars as words); all bases of one call
    /// belong to the same curve and group.  Returns (x || y limbs, is_identity) per job, in order.
    pub fn msm_batch(&self, jobs: &[(&Srs<'_>, usize, &[u64])], scalars_mont: bool) -> Result<Vec<(Vec<u64>, bool)>> {
        if jobs.is_empty() {
            return Ok(Vec::new());
        }
        let words = jobs[0].0.words;
        let srs: Vec<*const zkb_srs> = jobs.iter().map(|j| j.0.raw as *const zkb_srs).collect();
        let offs: Vec<usize> = jobs.iter().map(|j| j.1).collect();
        let ptrs: Vec<*const u64> = jobs.iter().map(|j| j.2.as_ptr()).collect();
        let lens: Vec<usize> = jobs.iter().map(|j| j.2.len() / 4).collect();
        let mut xy = vec![0u64; words * jobs.len()];
        let mut inf = vec![0u8; jobs.len()];
        self.check(unsafe {
            zkb_msm_batch(self.raw, jobs.len(), srs.as_ptr(), offs.as_ptr(), ptrs.as_ptr(), lens.as_ptr(),
                          scalars_mont as c_int, xy.as_mut_ptr(), inf.as_mut_ptr())
        })?;
        Ok(xy.chunks(words).zip(inf.iter()).map(|(p, &i)| (p.to_vec(), i != 0)).collect())
    }

    /// Products of pairings for MANY checks in one call (`zkb_multi_pairing`): group g = pairs
    /// [g * group_size, (g + 1) * group_size) and the result holds one GT element (12 Fq, Montgomery, ark-ff tower order)
    /// per group -- what `verify_proof` (groth16/src/verifier.rs:31-41: `E::miller_loop` over three pairs, then
    /// `E::final_exponentiation`) and `KZG10::check` (marlin/src/pc/kzg10.rs:170-172) compute one check at a time.
    /// The values are a fixed power of ark-ec's: compare them only with other outputs of this call (or with one).
    pub fn multi_pairing(&self, curve: Curve, g1_xy: &[u64], g2_xy: &[u64], group_size: usize) -> Result<Vec<u64>> {
        let (w1, w2) = (curve.g1_words(), curve.g2_words());
        assert!(group_size > 0 && g1_xy.len() % w1 == 0 && g2_xy.len() % w2 == 0, "whole points, non-empty groups");
        let n = g1_xy.len() / w1;
        assert!(g2_xy.len() / w2 == n && n % group_size == 0, "one G2 point per G1 point, whole groups");
        let mut gt = vec![0u64; (n / group_size) * 6 * w1];
        self.check(unsafe {
            zkb_multi_pairing(self.raw, curve as c_int, g1_xy.as_ptr(), std::ptr::null(), g2_xy.as_ptr(), std::ptr::null(),
                              n / group_size, group_size, gt.as_mut_ptr())
        })?;
        Ok(gt)
    }

    /// ark-serialize compressed points -> x || y Montgomery limbs + infinity bytes (`zkb_points_decompress`): what
    /// `Parameters::<E>::deserialize` does per point (groth16/src/lib.rs:81, cli/src/zkp_prove.rs:117-124), with the
    /// square roots on the device.  A status other than 0 is `SerializationError::InvalidData` for that point.
    pub fn points_decompress(&self, curve: Curve, g2: bool, compressed: &[u8], check_subgroup: bool)
                             -> Result<(Vec<u64>, Vec<u8>, Vec<u8>)> {
        let words = if g2 { curve.g2_words() } else { curve.g1_words() };
        let bytes_per_point = words * 4;
        assert_eq!(compressed.len() % bytes_per_point, 0, "compressed points are half an affine point each");
        let n = compressed.len() / bytes_per_point;
        let (mut xy, mut inf, mut status) = (vec![0u64; n * words], vec![0u8; n], vec![0u8; n]);
        let flags = if check_subgroup { ZKB_DECOMPRESS_CHECK_SUBGROUP } else { 0 };
        self.check(unsafe {
            zkb_points_decompress(self.raw, curve as c_int, if g2 { ZKB_G2 } else { ZKB_G1 }, compressed.as_ptr(), n, flags,
                                  xy.as_mut_ptr(), inf.as_mut_ptr(), status.as_mut_ptr())
        })?;
        Ok((xy, inf, status))
    }

    /// out[i] = in[0] * ... * in[i - 1], out[0] = 1 (`zkb_fr_prefix_product`): the accumulator z of PLONK's
    /// permutation argument (plonk/src/ahp/indexer/permutation.rs:111-118).
    pub fn fr_prefix_product(&self, curve: Curve, in_mont: &[u64]) -> Result<Vec<u64>> {
        let mut out = vec![0u64; in_mont.len()];
        self.check(unsafe { zkb_fr_prefix_product(self.raw, curve as c_int, in_mont.as_ptr(), out.as_mut_ptr(), in_mont.len() / 4) })?;
        Ok(out)
    }

    /// `R1CStoQAP::witness_map` + the `into_repr` sweep of prover.rs:161: h in canonical form.
    pub fn groth16_h(&self, curve: Curve, a: &Csr<'_>, b: &Csr<'_>, c: &Csr<'_>, z_mont: &[u64], n_inputs: usize,
                     n_aux: usize) -> Result<Vec<u64>> {
        assert_eq!(z_mont.len(), 4 * (n_inputs + n_aux));
        let domain = (a.raw.n_rows + n_inputs).next_power_of_two();
        let mut h = vec![0u64; 4 * domain];
        self.check(unsafe {
            zkb_groth16_h(self.raw, curve as c_int, &a.raw, &b.raw, &c.raw, z_mont.as_ptr(), n_inputs, n_aux, h.as_mut_ptr())
        })?;
        Ok(h)
    }
}

impl Drop for Context {
    fn drop(&mut self) {
        unsafe { zkb_destroy(self.raw) }
    }
}

/// Bases resident on one GPU (optionally with the window tables 2^(c j) P_i precomputed).
pub struct Srs<'c> {
    ctx: &'c Context,
    raw: *mut zkb_srs,
    words: usize,
}

impl<'c> Srs<'c> {
    pub fn len(&self) -> usize {
        unsafe { zkb_srs_len(self.raw) }
    }
    pub fn is_empty(&self) -> bool {
        self.len() == 0
    }
    /// `multi_scalar_mul(&bases[base_offset..], scalars)`; `scalars_mont`: the scalars are `Fr` values (Montgomery limbs,
    /// curve/src/lib.rs:38-45) rather than `into_repr()` integers.  Returns (x || y limbs, is_identity).
    pub fn msm(&self, base_offset: usize, scalars: &[u64], scalars_mont: bool) -> Result<(Vec<u64>, bool)> {
        let n = scalars.len() / 4;
        let mut xy = vec![0u64; self.words];
        let mut inf = 0u8;
        let f = if scalars_mont { zkb_msm_mont } else { zkb_msm };
        self.ctx.check(unsafe { f(self.ctx.raw, self.raw, base_offset, scalars.as_ptr(), n, xy.as_mut_ptr(), &mut inf) })?;
        Ok((xy, inf != 0))
    }
}

impl<'c> Drop for Srs<'c> {
    fn drop(&mut self) {
        unsafe { zkb_srs_free(self.raw) }
    }
}

/// Borrowed CSR view of `ProvingAssignment::{at, bt, ct}` (groth16/src/prover.rs:16-25): row i holds (coeff, column)
/// pairs with column = Input(i) -> i, Aux(i) -> num_inputs + i (groth16/src/r1cs_to_qap.rs:34-37).
pub struct Csr<'a> {
    pub raw: zkb_csr,
    _marker: std::marker::PhantomData<&'a ()>,
}

impl<'a> Csr<'a> {
    pub fn new(row_ptr: &'a [u32], col_idx: &'a [u32], coeff_mont: &'a [u64]) -> Self {
        assert!(!row_ptr.is_empty() && coeff_mont.len() == 4 * col_idx.len());
        assert_eq!(*row_ptr.last().unwrap() as usize, col_idx.len());
        Csr {
            raw: zkb_csr { n_rows: row_ptr.len() - 1, nnz: col_idx.len(), row_ptr: row_ptr.as_ptr(), col_idx: col_idx.as_ptr(),
                           coeff_mont: coeff_mont.as_ptr() },
            _marker: std::marker::PhantomData,
        }
    }
}

/// One query of `Parameters<E>` in the ABI layout.
pub struct Query<'a> {
    pub xy_mont: &'a [u64],
    pub inf: &'a [u8],
}

/// `Parameters<E>` (groth16/src/lib.rs:81-91) resident in HBM, created once and reused by every proof.
pub struct ProvingKey<'c> {
    ctx: &'c Context,
    raw: *mut zkb_pk,
    curve: Curve,
}

impl<'c> ProvingKey<'c> {
    /// singles: g1 = [alpha_g1, beta_g1, delta_g1], g2 = [beta_g2, delta_g2].  `shard = Some((n_ranks, rank))` keeps only
    /// this rank's slice of the pairs of each MSM (one process per GPU, `Context::comm_init` done beforehand).
    pub fn new(ctx: &'c Context, curve: Curve, a: Query<'_>, b_g1: Query<'_>, b_g2: Query<'_>, h: Query<'_>, l: Query<'_>,
               g1_singles: &[u64], g2_singles: &[u64], shard: Option<(i32, i32)>) -> Result<Self> {
        assert_eq!(g1_singles.len(), 3 * curve.g1_words());
        assert_eq!(g2_singles.len(), 2 * curve.g2_words());
        let mut raw = ptr::null_mut();
        let rc = unsafe {
            match shard {
                None => zkb_groth16_pk_create(ctx.raw, curve as c_int, a.xy_mont.as_ptr(), a.inf.as_ptr(), a.inf.len(),
                                              b_g1.xy_mont.as_ptr(), b_g1.inf.as_ptr(), b_g1.inf.len(), b_g2.xy_mont.as_ptr(),
                                              b_g2.inf.as_ptr(), b_g2.inf.len(), h.xy_mont.as_ptr(), h.inf.as_ptr(), h.inf.len(),
                                              l.xy_mont.as_ptr(), l.inf.as_ptr(), l.inf.len(), g1_singles.as_ptr(),
                                              g2_singles.as_ptr(), &mut raw),
                Some((n_ranks, rank)) => zkb_groth16_pk_create_sharded(
                    ctx.raw, curve as c_int, a.xy_mont.as_ptr(), a.inf.as_ptr(), a.inf.len(), b_g1.xy_mont.as_ptr(),
                    b_g1.inf.as_ptr(), b_g1.inf.len(), b_g2.xy_mont.as_ptr(), b_g2.inf.as_ptr(), b_g2.inf.len(), h.xy_mont.as_ptr(),
                    h.inf.as_ptr(), h.inf.len(), l.xy_mont.as_ptr(), l.inf.as_ptr(), l.inf.len(), g1_singles.as_ptr(),
                    g2_singles.as_ptr(), n_ranks, rank, &mut raw),
            }
        };
        ctx.check(rc)?;
        Ok(ProvingKey { ctx, raw, curve })
    }

    /// `create_proof` from "prover filled" (prover.rs:146) to "Proof assembled" (:206).  r, s: `into_repr()` limbs.
    /// Returns A (G1) || B (G2) || C (G1) limbs and the three infinity flags.
    pub fn prove(&self, a: &Csr<'_>, b: &Csr<'_>, c: &Csr<'_>, z_mont: &[u64], n_inputs: usize, n_aux: usize, r: &[u64; 4],
                 s: &[u64; 4], sharded: bool) -> Result<(Vec<u64>, [u8; 3])> {
        assert_eq!(z_mont.len(), 4 * (n_inputs + n_aux));
        let mut xy = vec![0u64; 2 * self.curve.g1_words() + self.curve.g2_words()];
        let mut inf = [0u8; 3];
        let f = if sharded { zkb_groth16_prove_sharded } else { zkb_groth16_prove };
        self.ctx.check(unsafe {
            f(self.ctx.raw, self.raw, &a.raw, &b.raw, &c.raw, z_mont.as_ptr(), n_inputs, n_aux, r.as_ptr(), s.as_ptr(),
              xy.as_mut_ptr(), inf.as_mut_ptr())
        })?;
        Ok((xy, inf))
    }
}

impl<'c> Drop for ProvingKey<'c> {
    fn drop(&mut self) {
        unsafe { zkb_groth16_pk_free(self.raw) }
    }
}

/// Multi-GPU rendezvous: rank 0 creates the id, the host distributes it (MPI broadcast, a file, a socket), every rank
/// calls `comm_init` -- collective.  NCCL is dlopen'ed by the library.
impl Context {
    pub fn comm_unique_id(&self) -> Result<[u8; ZKB_COMM_ID_BYTES]> {
        let mut id = [0u8; ZKB_COMM_ID_BYTES];
        self.check(unsafe { zkb_comm_unique_id(self.raw, id.as_mut_ptr()) })?;
        Ok(id)
    }
    pub fn comm_init(&self, n_ranks: i32, rank: i32, id: &[u8; ZKB_COMM_ID_BYTES]) -> Result<()> {
        self.check(unsafe { zkb_comm_init(self.raw, n_ranks, rank, id.as_ptr()) })
    }
}

/// View a slice of 32-byte field elements (`ark_ff::Fp256`) as u64 words.  The caller asserts the layout once:
/// `assert_eq!(core::mem::size_of::<Fr>(), 32); assert_eq!(core::mem::align_of::<Fr>(), 8);`
pub unsafe fn fr_slice_as_words<T>(v: &[T]) -> &[u64] {
    assert_eq!(std::mem::size_of::<T>(), 32, "Fp256 is four u64 limbs");
    std::slice::from_raw_parts(v.as_ptr() as *const u64, 4 * v.len())
}
