//! marlin/src/pc/zkb_backend.rs -- NEW FILE of the patched `zkp-marlin` crate (feature "zkb").
//!
//! KZG10::commit / KZG10::open (pc/kzg10.rs:100-156) spend their time in `VariableBaseMSM::multi_scalar_mul` over
//! `powers_of_g[num_leading_zeros..]`.  The committer key is uploaded once (cached by the address of `powers_of_g`, which
//! lives as long as the `CommitterKey`); every commitment is then `zkb_msm(srs, base_offset = num_leading_zeros, coeffs)`
//! -- the shifted powers of a degree-bounded commitment (`CommitterKey::shifted_powers`, pc/data_structures.rs:87-99) are a
//! sub-slice of the same array, i.e. a larger base offset into the same resident key.
//! The hiding MSMs over `powers_of_gamma_g` (two terms) stay on the CPU.
use ark_ec::PairingEngine;
use ark_ff::PrimeField;
use std::collections::HashMap;
use std::sync::Mutex;
use zkb_sys::{Context, Srs};
use zkp_groth16::zkb_backend::ZkbEngine;      // the same per-engine point marshalling

use super::Powers;

lazy_static::lazy_static! {
    static ref CONTEXT: Context = Context::new(0).expect("zkb: no usable B200");
    /// resident keys: base address of the longest `powers_of_g` slice seen -> (SRS, number of points)
    static ref KEYS: Mutex<HashMap<usize, (Srs<'static>, usize)>> = Mutex::new(HashMap::new());
}

/// multi_scalar_mul(&ck.powers_of_g[num_leading_zeros..], coeffs) on the GPU
pub(crate) fn msm_g<E: PairingEngine + ZkbEngine>(ck: &Powers<'_, E>, num_leading_zeros: usize,
                                                  coeffs: &[<E::Fr as PrimeField>::BigInt]) -> E::G1Projective {
    let ctx: &'static Context = &CONTEXT;
    let all: &[E::G1Affine] = &ck.powers_of_g;
    let end = all.as_ptr() as usize + all.len() * core::mem::size_of::<E::G1Affine>();
    let mut keys = KEYS.lock().unwrap();
    // a shifted-powers slice ends where the full key ends: find the resident key that contains it
    let hit = keys.iter().find(|(base, (_, n))| **base <= all.as_ptr() as usize && **base + n * core::mem::size_of::<E::G1Affine>() == end)
        .map(|(base, _)| *base);
    let base = match hit {
        Some(b) => b,
        None => {
            let (xy, inf) = E::pack_g1(all);
            let srs = ctx.srs_upload(E::CURVE, false, &xy, &inf, true).expect("zkb_srs_upload");
            keys.insert(all.as_ptr() as usize, (srs, all.len()));
            all.as_ptr() as usize
        }
    };
    let offset = (all.as_ptr() as usize - base) / core::mem::size_of::<E::G1Affine>() + num_leading_zeros;
    let words = unsafe { core::slice::from_raw_parts(coeffs.as_ptr() as *const u64, 4 * coeffs.len()) };   // BigInteger256 = [u64; 4]
    let (xy, inf) = keys[&base].0.msm(offset, words, false).expect("zkb_msm");
    E::unpack_g1(&xy, inf).into()
}
