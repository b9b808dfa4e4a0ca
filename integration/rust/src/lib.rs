//! Thin Rust side of the drop-in boundary (source only; not compiled here -- there is no Rust toolchain in the
//! environment this repository was built in). `ffi` is generated from include/bp_b200.h by tools/gen_rust_ffi.py.
//!
//! The two functions below are what the reference crate's call sites turn into:
//!   * `G::Group::msm(&bases, &scalars)`  (17 call sites)        -> [`msm`]
//!   * `Prover::prove(self, prng, &bp_gens)` (src/r1cs/prover.rs:444)               -> `bp_prover_prove` via the handle API
//! ark-ff `Fp256<MontBackend<_, 4>>` is `[u64; 4]` little-endian limbs in Montgomery form (R = 2^256) -- exactly the
//! 32-byte scalar / 2 x 32-byte affine point format of the C ABI, so marshalling is a copy of the limbs.
pub mod ffi;
pub mod r1cs;

use ark_ec::{short_weierstrass::Affine, AffineRepr};
use ark_ff::{BigInt, Fp256, MontBackend, MontConfig};
use core::ffi::c_int;

/// Raw Montgomery limbs of an ark-ff 256-bit field element.
pub(crate) fn limbs<P: MontConfig<4>>(x: &Fp256<MontBackend<P, 4>>) -> [u8; 32] {
    let BigInt(l) = x.0; // the in-memory (Montgomery) representation, not `into_bigint()`
    let mut out = [0u8; 32];
    for (i, w) in l.iter().enumerate() {
        out[8 * i..8 * i + 8].copy_from_slice(&w.to_le_bytes());
    }
    out
}

/// `<ark_secq256k1::Projective as VariableBaseMSM>::msm(bases, scalars)` on the GPU. Returns the affine sum
/// (identity = `Affine::identity()`), the caller's `.into_affine()` becomes a no-op.
pub fn msm(
    ctx: *mut ffi::BpCtx,
    bases: &[ark_secq256k1::Affine],
    scalars: &[ark_secq256k1::Fr],
) -> Result<ark_secq256k1::Affine, c_int> {
    assert_eq!(bases.len(), scalars.len()); // the reference: msm(..).unwrap() panics on a length mismatch
    let mut b = Vec::with_capacity(64 * bases.len());
    for p in bases {
        if p.is_zero() {
            b.extend_from_slice(&[0u8; 64]); // identity = (0, 0)
        } else {
            b.extend_from_slice(&limbs(&p.x));
            b.extend_from_slice(&limbs(&p.y));
        }
    }
    let mut s = Vec::with_capacity(32 * scalars.len());
    for k in scalars {
        s.extend_from_slice(&limbs(k));
    }
    let (mut out, mut ident) = ([0u8; 64], 0 as c_int);
    let rc = unsafe { ffi::bp_msm(ctx, b.as_ptr(), s.as_ptr(), bases.len(), out.as_mut_ptr(), &mut ident) };
    if rc != ffi::BP_OK {
        return Err(rc);
    }
    if ident != 0 {
        return Ok(Affine::identity());
    }
    let fq = |raw: &[u8]| {
        let mut l = [0u64; 4];
        for i in 0..4 {
            l[i] = u64::from_le_bytes(raw[8 * i..8 * i + 8].try_into().unwrap());
        }
        ark_ff::Fp::<MontBackend<ark_secq256k1::FqConfig, 4>, 4>(BigInt(l), core::marker::PhantomData)
    };
    Ok(Affine::new_unchecked(fq(&out[..32]), fq(&out[32..])))
}

/// ark-ff field element from the 32 raw Montgomery bytes of the C ABI (no conversion: `Fp(BigInt(limbs))`).
pub(crate) fn fp_from_limbs<P: MontConfig<4>>(raw: &[u8; 32]) -> Fp256<MontBackend<P, 4>> {
    let mut l = [0u64; 4];
    for i in 0..4 {
        l[i] = u64::from_le_bytes(raw[8 * i..8 * i + 8].try_into().unwrap());
    }
    ark_ff::Fp::<MontBackend<P, 4>, 4>(BigInt(l), core::marker::PhantomData)
}
/// `x || y` as raw Montgomery limbs, identity = 64 zero bytes.
pub(crate) fn affine_bytes(p: &ark_secq256k1::Affine) -> [u8; 64] {
    let mut out = [0u8; 64];
    if !p.is_zero() {
        out[..32].copy_from_slice(&limbs(&p.x));
        out[32..].copy_from_slice(&limbs(&p.y));
    }
    out
}
pub(crate) fn affine_from_bytes(b: &[u8; 64]) -> ark_secq256k1::Affine {
    if b.iter().all(|&x| x == 0) {
        return Affine::identity();
    }
    Affine::new_unchecked(fp_from_limbs(b[..32].try_into().unwrap()), fp_from_limbs(b[32..].try_into().unwrap()))
}
