//! curve/src/zkb_backend.rs -- NEW FILE of the patched `zkp-curve` crate (feature "zkb").
//!
//! `Curve::vartime_multiscalar_mul(scalars, points)` (curve/src/lib.rs:38-45; note the argument order, reversed w.r.t.
//! `VariableBaseMSM::multi_scalar_mul`) for the consumers outside Groth16 / Marlin (spartan/src/commitments.rs:42-56,
//! bulletproofs, hyrax, libra, asvc/src/lib.rs:160-225): the points are uploaded (cached by slice address + length, the
//! commitment keys of those schemes are long-lived), the scalars travel as Montgomery limbs and `into_repr` is fused on
//! the device (zkb_msm_mont).  Small calls stay on the CPU: an upload costs more than a few hundred additions.
use crate::Curve;
use std::collections::HashMap;
use std::sync::Mutex;
use zkb_sys::{Context, Curve as ZkbCurve, Srs};

/// Point marshalling for the G1 group of an engine the backend knows (BLS12-381, BN254); see groth16's `ZkbEngine`.
pub trait ZkbPoints: Curve {
    const CURVE: ZkbCurve;
    fn pack(points: &[Self::Affine]) -> (Vec<u64>, Vec<u8>);
    fn unpack(limbs: &[u64], infinity: bool) -> Self::Affine;
}

pub const MIN_GPU_TERMS: usize = 1 << 12;

lazy_static::lazy_static! {
    static ref CONTEXT: Context = Context::new(0).expect("zkb: no usable B200");
    static ref BASES: Mutex<HashMap<(usize, usize), Srs<'static>>> = Mutex::new(HashMap::new());
}

pub fn msm<C: ZkbPoints>(scalars: &[C::Fr], points: &[C::Affine]) -> C::Projective {
    let n = core::cmp::min(scalars.len(), points.len());          // multi_scalar_mul zips: the shorter one wins
    let ctx: &'static Context = &CONTEXT;
    let mut cache = BASES.lock().unwrap();
    let srs = cache.entry((points.as_ptr() as usize, points.len())).or_insert_with(|| {
        let (xy, inf) = C::pack(points);
        ctx.srs_upload(C::CURVE, false, &xy, &inf, points.len() >= 1 << 16).expect("zkb_srs_upload")
    });
    let words = unsafe { zkb_sys::fr_slice_as_words(&scalars[..n]) };
    let (xy, inf) = srs.msm(0, words, true).expect("zkb_msm_mont");
    C::unpack(&xy, inf).into()
}

/// `None` when `C` has no marshalling impl or the call is too small to be worth a transfer.
pub fn try_msm<C: Curve>(scalars: &[C::Fr], points: &[C::Affine]) -> Option<C::Projective> {
    if core::cmp::min(scalars.len(), points.len()) < MIN_GPU_TERMS {
        return None;
    }
    crate::zkb_dispatch::dispatch::<C>(scalars, points)     // TypeId match onto the ZkbPoints impls, as in zkp-groth16
}
