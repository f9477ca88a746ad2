//! groth16/src/zkb_backend.rs -- NEW FILE of the patched `zkp-groth16` crate (feature "zkb").
//!
//! Everything `create_proof` does between "prover filled" (prover.rs:146) and "Proof assembled" (:206) -- witness_map,
//! the into_repr sweeps, the five MSMs, the assembly -- is ONE call into libzkb.so.  This module only marshals:
//!   * `ProvingAssignment::{at, bt, ct}` -> CSR (`Index::Input(i)` -> i, `Index::Aux(i)` -> num_inputs + i, r1cs_to_qap.rs:34-37);
//!   * `Parameters<E>` -> x || y limb arrays + infinity bytes, once per key (cached in a side table keyed by the key's address);
//!   * the proof's limbs back into `GroupAffine`s.
//! Field elements cross the boundary as their in-memory Montgomery limbs (ark-ff 0.2 `Fp256(BigInteger256([u64; 4]))`),
//! r and s as `into_repr()` limbs.  Shipped as source: no Rust toolchain exists in the image libzkb is built in.
use ark_ec::PairingEngine;
use ark_ff::{BigInteger, PrimeField};
use std::collections::HashMap;
use std::sync::Mutex;
use zkb_sys::{Context, Csr, Curve, ProvingKey, Query, ZkbError};
use zkp_r1cs::{Index, SynthesisError};

use crate::{prover::ProvingAssignment, Parameters, Proof};

/// Implemented for the pairing engines the backend has field / curve constants for (BLS12-381 and BN254); the methods
/// touch the concrete `GroupAffine { x, y, infinity }` / `Fp2 { c0, c1 }` fields, which the generic traits do not expose.
pub trait ZkbEngine: PairingEngine {
    const CURVE: Curve;
    fn pack_g1(points: &[Self::G1Affine]) -> (Vec<u64>, Vec<u8>);
    fn pack_g2(points: &[Self::G2Affine]) -> (Vec<u64>, Vec<u8>);
    fn unpack_g1(limbs: &[u64], infinity: bool) -> Self::G1Affine;
    fn unpack_g2(limbs: &[u64], infinity: bool) -> Self::G2Affine;
}

macro_rules! impl_zkb_engine {
    ($engine:ty, $curve:expr, $fq:ty, $big:ty, $fq2:ty, $g1:ty, $g2:ty) => {
        impl ZkbEngine for $engine {
            const CURVE: Curve = $curve;
            fn pack_g1(points: &[Self::G1Affine]) -> (Vec<u64>, Vec<u8>) {
                let mut xy = Vec::with_capacity(points.len() * Self::CURVE.g1_words());
                let mut inf = Vec::with_capacity(points.len());
                for p in points {
                    xy.extend_from_slice(&(p.x.0).0);       // Montgomery limbs, as stored
                    xy.extend_from_slice(&(p.y.0).0);
                    inf.push(p.infinity as u8);
                }
                (xy, inf)
            }
            fn pack_g2(points: &[Self::G2Affine]) -> (Vec<u64>, Vec<u8>) {
                let mut xy = Vec::with_capacity(points.len() * Self::CURVE.g2_words());
                let mut inf = Vec::with_capacity(points.len());
                for p in points {
                    for c in [&p.x.c0, &p.x.c1, &p.y.c0, &p.y.c1].iter() {
                        xy.extend_from_slice(&(c.0).0);
                    }
                    inf.push(p.infinity as u8);
                }
                (xy, inf)
            }
            fn unpack_g1(limbs: &[u64], infinity: bool) -> Self::G1Affine {
                if infinity {
                    return <$g1 as ark_ec::AffineCurve>::zero();
                }
                let l = Self::CURVE.fq_limbs();
                let fq = |w: &[u64]| { let mut b = <$big>::default(); b.0.copy_from_slice(w); <$fq>::new(b) };   // limbs ARE the Montgomery residue
                <$g1>::new(fq(&limbs[..l]), fq(&limbs[l..2 * l]), false)
            }
            fn unpack_g2(limbs: &[u64], infinity: bool) -> Self::G2Affine {
                if infinity {
                    return <$g2 as ark_ec::AffineCurve>::zero();
                }
                let l = Self::CURVE.fq_limbs();
                let fq = |w: &[u64]| { let mut b = <$big>::default(); b.0.copy_from_slice(w); <$fq>::new(b) };
                let fq2 = |w: &[u64]| <$fq2>::new(fq(&w[..l]), fq(&w[l..2 * l]));
                <$g2>::new(fq2(&limbs[..2 * l]), fq2(&limbs[2 * l..4 * l]), false)
            }
        }
    };
}

#[cfg(feature = "zkb-bls12-381")]
impl_zkb_engine!(ark_bls12_381::Bls12_381, Curve::Bls12_381, ark_bls12_381::Fq, ark_ff::BigInteger384, ark_bls12_381::Fq2,
                 ark_bls12_381::G1Affine, ark_bls12_381::G2Affine);
#[cfg(feature = "zkb-bn254")]
impl_zkb_engine!(ark_bn254::Bn254, Curve::Bn254, ark_bn254::Fq, ark_ff::BigInteger256, ark_bn254::Fq2, ark_bn254::G1Affine,
                 ark_bn254::G2Affine);

/// `Vec<Vec<(Fr, Index)>>` -> CSR arrays (duplicate columns inside a row are kept: r1cs/src/impl_lc.rs:58-70)
pub(crate) fn to_csr<E: PairingEngine>(rows: &[Vec<(E::Fr, Index)>], num_inputs: usize) -> (Vec<u32>, Vec<u32>, Vec<u64>) {
    let nnz: usize = rows.iter().map(|r| r.len()).sum();
    let (mut ptr, mut col, mut val) = (Vec::with_capacity(rows.len() + 1), Vec::with_capacity(nnz), Vec::with_capacity(4 * nnz));
    ptr.push(0u32);
    for row in rows {
        for (coeff, index) in row {
            col.push(match index {
                Index::Input(i) => *i as u32,
                Index::Aux(i) => (num_inputs + *i) as u32,
            });
            val.extend_from_slice(unsafe { zkb_sys::fr_slice_as_words(core::slice::from_ref(coeff)) });
        }
        ptr.push(col.len() as u32);
    }
    (ptr, col, val)
}

struct Resident {
    ctx: &'static Context,
    pk: ProvingKey<'static>,
}
unsafe impl Send for Resident {}

lazy_static::lazy_static! {
    /// one context per process (device $ZKB_DEVICE, default 0) and the keys made resident so far
    static ref CONTEXT: Context = Context::new(std::env::var("ZKB_DEVICE").ok().and_then(|s| s.parse().ok()).unwrap_or(0))
        .expect("zkb: no usable B200 (there is no CPU fallback in the zkb build of zkp-groth16)");
    static ref KEYS: Mutex<HashMap<usize, Resident>> = Mutex::new(HashMap::new());
}

fn map_err(e: ZkbError) -> SynthesisError {
    if e.is_degree_too_large() {
        SynthesisError::PolynomialDegreeTooLarge          // EvaluationDomain::new(..) == None (r1cs_to_qap.rs:123-125)
    } else {
        eprintln!("{}", e);
        SynthesisError::Unsatisfiable
    }
}

/// prover.rs:148-210 on the GPU
pub(crate) fn prove<E: ZkbEngine>(params: &Parameters<E>, prover: &ProvingAssignment<E>, r: E::Fr, s: E::Fr)
                                  -> Result<Proof<E>, SynthesisError> {
    let ctx: &'static Context = &CONTEXT;
    let mut keys = KEYS.lock().unwrap();
    let resident = match keys.entry(params as *const _ as usize) {
        std::collections::hash_map::Entry::Occupied(e) => e.into_mut(),
        std::collections::hash_map::Entry::Vacant(v) => {
            let (a, a_inf) = E::pack_g1(params.get_a_query_full()?);
            let (b1, b1_inf) = E::pack_g1(params.get_b_g1_query_full()?);
            let (b2, b2_inf) = E::pack_g2(params.get_b_g2_query_full()?);
            let (h, h_inf) = E::pack_g1(params.get_h_query_full()?);
            let (l, l_inf) = E::pack_g1(params.get_l_query_full()?);
            let (g1s, _) = E::pack_g1(&[params.vk.alpha_g1, params.beta_g1, params.delta_g1]);
            let (g2s, _) = E::pack_g2(&[params.vk.beta_g2, params.vk.delta_g2]);
            let pk = ProvingKey::new(ctx, E::CURVE, Query { xy_mont: &a, inf: &a_inf }, Query { xy_mont: &b1, inf: &b1_inf },
                                     Query { xy_mont: &b2, inf: &b2_inf }, Query { xy_mont: &h, inf: &h_inf },
                                     Query { xy_mont: &l, inf: &l_inf }, &g1s, &g2s, None).map_err(map_err)?;
            v.insert(Resident { ctx, pk })
        }
    };
    let ni = prover.input_assignment.len();
    let (ap, ac, av) = to_csr::<E>(&prover.at, ni);
    let (bp, bc, bv) = to_csr::<E>(&prover.bt, ni);
    let (cp, cc, cv) = to_csr::<E>(&prover.ct, ni);
    let z: Vec<E::Fr> = prover.input_assignment.iter().chain(prover.aux_assignment.iter()).cloned().collect();
    let mut r_limbs = [0u64; 4];
    let mut s_limbs = [0u64; 4];
    r_limbs.copy_from_slice(r.into_repr().as_ref());
    s_limbs.copy_from_slice(s.into_repr().as_ref());
    let (xy, inf) = resident.pk.prove(&Csr::new(&ap, &ac, &av), &Csr::new(&bp, &bc, &bv), &Csr::new(&cp, &cc, &cv),
                                      unsafe { zkb_sys::fr_slice_as_words(&z) }, ni, prover.aux_assignment.len(), &r_limbs,
                                      &s_limbs, false).map_err(map_err)?;
    let (w1, w2) = (E::CURVE.g1_words(), E::CURVE.g2_words());
    Ok(Proof { a: E::unpack_g1(&xy[..w1], inf[0] != 0), b: E::unpack_g2(&xy[w1..w1 + w2], inf[1] != 0),
               c: E::unpack_g1(&xy[w1 + w2..], inf[2] != 0) })
}

/// Entry used by the patched `create_proof` (prover.rs): `Some(..)` when `E` is an engine the backend knows, else `None`
/// and the reference's CPU body runs.  `E` is matched by `TypeId`; inside a branch `E` IS the concrete engine, so the
/// pointer casts below only rename the type for the compiler.
pub(crate) fn try_prove<E: PairingEngine>(params: &Parameters<E>, prover: &ProvingAssignment<E>, r: E::Fr, s: E::Fr)
                                          -> Option<Result<Proof<E>, SynthesisError>> {
    use core::any::TypeId;
    use core::mem::{transmute_copy, ManuallyDrop};
    macro_rules! dispatch {
        ($engine:ty) => {
            if TypeId::of::<E>() == TypeId::of::<$engine>() {
                let out = prove::<$engine>(unsafe { &*(params as *const Parameters<E> as *const Parameters<$engine>) },
                                           unsafe { &*(prover as *const ProvingAssignment<E> as *const ProvingAssignment<$engine>) },
                                           unsafe { transmute_copy(&r) }, unsafe { transmute_copy(&s) });
                return Some(out.map(|p| unsafe { transmute_copy::<Proof<$engine>, Proof<E>>(&*ManuallyDrop::new(p)) }));
            }
        };
    }
    #[cfg(feature = "zkb-bls12-381")]
    dispatch!(ark_bls12_381::Bls12_381);
    #[cfg(feature = "zkb-bn254")]
    dispatch!(ark_bn254::Bn254);
    let _ = (params, prover, r, s);
    None
}
