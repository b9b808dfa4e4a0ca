//! The crate's R1CS types as thin wrappers over the handle API of libbp_b200.so (source only; not compiled in the
//! environment this repository was built in -- no Rust toolchain).
//!
//! Each type mirrors its counterpart in the reference and forwards to the C entry point that replaces its body:
//!   BulletproofGens::new / PedersenGens::default   src/generators.rs:47-66,174-183   -> bp_gens_create
//!   PedersenGens::commit                            src/generators.rs:39-44           -> bp_pedersen_commit
//!   Prover::new / commit / prove                    src/r1cs/prover.rs:291,327,444    -> bp_prover_new / _commit / _prove
//!   Verifier::new / commit / verify                 src/r1cs/verifier.rs:252,279,549  -> bp_verifier_new / _commit / _verify
//!   batch_verify                                    src/r1cs/verifier.rs:604-691      -> bp_batch_verify
//!   R1CSProof::to_bytes / from_bytes                src/r1cs/proof.rs:74-91           -> bp_proof_to_bytes / _from_bytes
//!   ConstraintSystem / Randomizable / Randomized    src/r1cs/constraint_system.rs     -> bp_cs_*
//! Scalars cross as the in-memory Montgomery limbs of `Fp256<MontBackend<_, 4>>` (32 bytes), points as `x || y` (64 bytes,
//! identity = zeros); see `crate::limbs`. Error codes map onto `R1CSError` (include/bp_b200.h: bp_status).
use crate::ffi::*;
use core::ffi::{c_int, c_void};
use core::marker::PhantomData;
use core::ptr;

/// `R1CSError` of the reference (src/errors.rs:118-143) as the library reports it.
#[derive(Debug, Clone, Copy, PartialEq, Eq)]
pub enum R1CSError {
    InvalidGeneratorsLength, // BP_ERR_GENS
    FormatError,             // BP_ERR_FORMAT
    VerificationError,       // BP_ERR_VERIFY
    MissingAssignment,       // BP_ERR_MISSING
    GadgetError(c_int),      // anything a randomized-constraint callback returned
    Backend(c_int),          // BP_ERR_CUDA, BP_ERR_NOGPU, BP_ERR_ARG ... (no counterpart in the CPU crate)
}
fn check(rc: c_int) -> Result<(), R1CSError> {
    match rc {
        0 => Ok(()),
        -4 => Err(R1CSError::InvalidGeneratorsLength),
        -8 => Err(R1CSError::FormatError),
        -7 => Err(R1CSError::VerificationError),
        -9 => Err(R1CSError::MissingAssignment),
        other => Err(R1CSError::Backend(other)),
    }
}

/// A curve the library instantiates: its id, and the byte views of its scalars and points.
pub trait Curve {
    const ID: c_int;
    type Scalar: Copy;
    type Affine: Copy;
    fn scalar_bytes(s: &Self::Scalar) -> [u8; 32];
    fn scalar_from(b: &[u8; 32]) -> Self::Scalar;
    fn point_bytes(p: &Self::Affine) -> [u8; 64];
    fn point_from(b: &[u8; 64]) -> Self::Affine;
}

/// secq256k1 (tests/r1cs_secq256k1.rs:5): `ark_secq256k1::{Affine, Fr}`.
pub struct Secq256k1;
impl Curve for Secq256k1 {
    const ID: c_int = BP_CURVE_SECQ256K1;
    type Scalar = ark_secq256k1::Fr;
    type Affine = ark_secq256k1::Affine;
    fn scalar_bytes(s: &Self::Scalar) -> [u8; 32] { crate::limbs(s) }
    fn scalar_from(b: &[u8; 32]) -> Self::Scalar { crate::fp_from_limbs(b) }
    fn point_bytes(p: &Self::Affine) -> [u8; 64] { crate::affine_bytes(p) }
    fn point_from(b: &[u8; 64]) -> Self::Affine { crate::affine_from_bytes(b) }
}

/// One GPU context (stream, scratch arena). `bp_ctx_create`; the reference has no counterpart (implicit CPU).
pub struct Context<C: Curve> { pub(crate) raw: *mut BpCtx, _c: PhantomData<C> }
impl<C: Curve> Context<C> {
    pub fn new(device: i32) -> Result<Self, R1CSError> {
        let mut raw = ptr::null_mut();
        check(unsafe { bp_ctx_create(C::ID, device, &mut raw) })?;
        Ok(Self { raw, _c: PhantomData })
    }
}
impl<C: Curve> Drop for Context<C> { fn drop(&mut self) { unsafe { bp_ctx_destroy(self.raw) } } }

/// `PedersenGens` + `BulletproofGens::new(capacity, 1)` resident in HBM (src/generators.rs:30-66,150-221).
pub struct BulletproofGens<'a, C: Curve> { pub(crate) raw: *mut BpGens, _ctx: &'a Context<C> }
impl<'a, C: Curve> BulletproofGens<'a, C> {
    /// `BulletproofGens::new(gens_capacity, 1)` together with `PedersenGens::default()`.
    pub fn new(ctx: &'a Context<C>, gens_capacity: usize) -> Result<Self, R1CSError> {
        let mut raw = ptr::null_mut();
        check(unsafe { bp_gens_create(ctx.raw, gens_capacity, &mut raw) })?;
        Ok(Self { raw, _ctx: ctx })
    }
    pub fn gens_capacity(&self) -> usize { unsafe { bp_gens_capacity(self.raw) } }
    /// `PedersenGens::commit(value, blinding)` (src/generators.rs:39-44).
    pub fn commit(&self, value: &C::Scalar, blinding: &C::Scalar) -> Result<C::Affine, R1CSError> {
        let mut out = [0u8; 64];
        check(unsafe { bp_pedersen_commit(self.raw, C::scalar_bytes(value).as_ptr(), C::scalar_bytes(blinding).as_ptr(), out.as_mut_ptr()) })?;
        Ok(C::point_from(&out))
    }
}
impl<'a, C: Curve> Drop for BulletproofGens<'a, C> { fn drop(&mut self) { unsafe { bp_gens_free(self.raw) } } }

/// `merlin::Transcript` with the same STROBE schedule.
pub struct Transcript { pub(crate) raw: *mut BpTranscript }
impl Transcript {
    pub fn new(label: &'static [u8]) -> Self { Self { raw: unsafe { bp_transcript_new(label.as_ptr(), label.len()) } } }
    pub fn append_message(&mut self, label: &'static [u8], msg: &[u8]) {
        unsafe { bp_transcript_append_message(self.raw, label.as_ptr(), label.len(), msg.as_ptr(), msg.len()) }
    }
    pub fn append_u64(&mut self, label: &'static [u8], v: u64) { unsafe { bp_transcript_append_u64(self.raw, label.as_ptr(), label.len(), v) } }
}
impl Clone for Transcript { fn clone(&self) -> Self { Self { raw: unsafe { bp_transcript_clone(self.raw) } } } }
impl Drop for Transcript { fn drop(&mut self) { unsafe { bp_transcript_free(self.raw) } } }

/// The caller's `RngCore + CryptoRng` behind `bp_rng_from_callbacks`.
pub struct RngAdapter<'r, R: rand_core::RngCore + rand_core::CryptoRng> { raw: *mut BpRng, _r: PhantomData<&'r mut R> }
unsafe extern "C" fn cb_u64<R: rand_core::RngCore>(u: *mut c_void) -> u64 { (*(u as *mut R)).next_u64() }
unsafe extern "C" fn cb_u32<R: rand_core::RngCore>(u: *mut c_void) -> u32 { (*(u as *mut R)).next_u32() }
unsafe extern "C" fn cb_fill<R: rand_core::RngCore>(u: *mut c_void, out: *mut u8, n: usize) {
    (*(u as *mut R)).fill_bytes(core::slice::from_raw_parts_mut(out, n))
}
impl<'r, R: rand_core::RngCore + rand_core::CryptoRng> RngAdapter<'r, R> {
    pub fn new(rng: &'r mut R) -> Self {
        let raw = unsafe { bp_rng_from_callbacks(rng as *mut R as *mut c_void, Some(cb_u64::<R>), Some(cb_u32::<R>), Some(cb_fill::<R>)) };
        Self { raw, _r: PhantomData }
    }
}
impl<'r, R: rand_core::RngCore + rand_core::CryptoRng> Drop for RngAdapter<'r, R> { fn drop(&mut self) { unsafe { bp_rng_free(self.raw) } } }

/// `r1cs::Variable` (src/r1cs/linear_combination.rs:14-27).
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub enum Variable { Committed(usize), MultiplierLeft(usize), MultiplierRight(usize), MultiplierOutput(usize), One() }
impl Variable {
    fn to_raw(self) -> BpVar {
        let (kind, index) = match self {
            Variable::Committed(i) => (0, i), Variable::MultiplierLeft(i) => (1, i), Variable::MultiplierRight(i) => (2, i),
            Variable::MultiplierOutput(i) => (3, i), Variable::One() => (4, 0),
        };
        BpVar { kind, reserved: 0, index: index as u64 }
    }
    fn from_raw(v: BpVar) -> Self {
        match v.kind {
            0 => Variable::Committed(v.index as usize), 1 => Variable::MultiplierLeft(v.index as usize),
            2 => Variable::MultiplierRight(v.index as usize), 3 => Variable::MultiplierOutput(v.index as usize), _ => Variable::One(),
        }
    }
}

/// `r1cs::LinearCombination` (src/r1cs/linear_combination.rs:85-88): a list of (Variable, coefficient) terms.
#[derive(Clone)]
pub struct LinearCombination<C: Curve> { pub terms: Vec<(Variable, C::Scalar)> }
impl<C: Curve> LinearCombination<C> {
    fn pack(&self) -> Vec<BpTerm> { self.terms.iter().map(|(v, c)| BpTerm { var: v.to_raw(), coeff: C::scalar_bytes(c) }).collect() }
}

/// `trait ConstraintSystem` + `RandomizableConstraintSystem` + `RandomizedConstraintSystem`
/// (src/r1cs/constraint_system.rs:19-135) over a `bp_cs` handle; `Prover` and `Verifier` both expose one.
pub struct Cs<C: Curve> { raw: *mut BpCs, _c: PhantomData<C> }
impl<C: Curve> Cs<C> {
    /// `multiply(left, right) -> (l, r, o)` (prover.rs:103-133 / verifier.rs:74-98).
    pub fn multiply(&mut self, left: &LinearCombination<C>, right: &LinearCombination<C>) -> Result<(Variable, Variable, Variable), R1CSError> {
        let (l, r) = (left.pack(), right.pack());
        let mut o = [BpVar { kind: 0, reserved: 0, index: 0 }; 3];
        check(unsafe { bp_cs_multiply(self.raw, l.as_ptr(), l.len(), r.as_ptr(), r.len(), o.as_mut_ptr()) })?;
        Ok((Variable::from_raw(o[0]), Variable::from_raw(o[1]), Variable::from_raw(o[2])))
    }
    /// `allocate(assignment)`; the verifier passes `None`.
    pub fn allocate(&mut self, assignment: Option<C::Scalar>) -> Result<Variable, R1CSError> {
        let b = assignment.map(|a| C::scalar_bytes(&a));
        let mut o = BpVar { kind: 0, reserved: 0, index: 0 };
        check(unsafe { bp_cs_allocate(self.raw, b.as_ref().map_or(ptr::null(), |x| x.as_ptr()), &mut o) })?;
        Ok(Variable::from_raw(o))
    }
    /// `allocate_multiplier(input_assignments)`.
    pub fn allocate_multiplier(&mut self, a: Option<(C::Scalar, C::Scalar)>) -> Result<(Variable, Variable, Variable), R1CSError> {
        let b = a.map(|(l, r)| (C::scalar_bytes(&l), C::scalar_bytes(&r)));
        let (pl, pr) = b.as_ref().map_or((ptr::null(), ptr::null()), |(l, r)| (l.as_ptr(), r.as_ptr()));
        let mut o = [BpVar { kind: 0, reserved: 0, index: 0 }; 3];
        check(unsafe { bp_cs_allocate_multiplier(self.raw, pl, pr, o.as_mut_ptr()) })?;
        Ok((Variable::from_raw(o[0]), Variable::from_raw(o[1]), Variable::from_raw(o[2])))
    }
    pub fn multipliers_len(&self) -> usize { unsafe { bp_cs_multipliers_len(self.raw) } }
    /// `constrain(lc)`: lc = 0.
    pub fn constrain(&mut self, lc: &LinearCombination<C>) -> Result<(), R1CSError> {
        let t = lc.pack();
        check(unsafe { bp_cs_constrain(self.raw, t.as_ptr(), t.len()) })
    }
    /// `specify_randomized_constraints(callback)`: the callback runs in the second phase with a system that can draw
    /// challenges (constraint_system.rs:99-135).
    pub fn specify_randomized_constraints<F>(&mut self, callback: F) -> Result<(), R1CSError>
    where F: FnMut(&mut Cs<C>) -> Result<(), R1CSError> + 'static {
        unsafe extern "C" fn tramp<C: Curve, F: FnMut(&mut Cs<C>) -> Result<(), R1CSError>>(cs: *mut BpCs, user: *mut c_void) -> c_int {
            let f = &mut *(user as *mut F);
            let mut inner = Cs::<C> { raw: cs, _c: PhantomData };
            match f(&mut inner) { Ok(()) => 0, Err(R1CSError::GadgetError(c)) | Err(R1CSError::Backend(c)) => c, Err(_) => -1 }
        }
        let boxed = Box::into_raw(Box::new(callback));     // lives until the prover / verifier is dropped
        check(unsafe { bp_cs_specify_randomized_constraints(self.raw, Some(tramp::<C, F>), boxed as *mut c_void) })
    }
    /// `challenge_scalar(label)` of the randomized phase.
    pub fn challenge_scalar(&mut self, label: &'static [u8]) -> Result<C::Scalar, R1CSError> {
        let mut out = [0u8; 32];
        check(unsafe { bp_cs_challenge_scalar(self.raw, label.as_ptr(), label.len(), out.as_mut_ptr()) })?;
        Ok(C::scalar_from(&out))
    }
}

/// `r1cs::R1CSProof` (src/r1cs/proof.rs:25-59).
pub struct R1CSProof<C: Curve> { pub(crate) raw: *mut BpProof, _c: PhantomData<C> }
impl<C: Curve> R1CSProof<C> {
    pub fn to_bytes(&self) -> Vec<u8> {
        let mut len = 0usize;
        unsafe { bp_proof_to_bytes(self.raw, ptr::null_mut(), 0, &mut len) };
        let mut out = vec![0u8; len];
        unsafe { bp_proof_to_bytes(self.raw, out.as_mut_ptr(), len, &mut len) };
        out
    }
    pub fn from_bytes(slice: &[u8]) -> Result<Self, R1CSError> {
        let mut raw = ptr::null_mut();
        check(unsafe { bp_proof_from_bytes(C::ID, slice.as_ptr(), slice.len(), &mut raw) })?;
        Ok(Self { raw, _c: PhantomData })
    }
}
impl<C: Curve> Drop for R1CSProof<C> { fn drop(&mut self) { unsafe { bp_proof_free(self.raw) } } }

/// `r1cs::Prover` (src/r1cs/prover.rs:30-45).
pub struct Prover<'g, 't, C: Curve> { raw: *mut BpProver, _g: PhantomData<&'g BulletproofGens<'g, C>>, _t: PhantomData<&'t mut Transcript> }
impl<'g, 't, C: Curve> Prover<'g, 't, C> {
    /// `Prover::new(pc_gens, transcript)` (prover.rs:291-308); the Bulletproof generators are given here, not at `prove`.
    pub fn new(ctx: &Context<C>, gens: &'g BulletproofGens<'g, C>, transcript: &'t mut Transcript) -> Result<Self, R1CSError> {
        let mut raw = ptr::null_mut();
        check(unsafe { bp_prover_new(ctx.raw, gens.raw, transcript.raw, &mut raw) })?;
        Ok(Self { raw, _g: PhantomData, _t: PhantomData })
    }
    /// `commit(v, v_blinding) -> (commitment, Variable)` (prover.rs:327-341).
    pub fn commit(&mut self, v: C::Scalar, v_blinding: C::Scalar) -> Result<(C::Affine, Variable), R1CSError> {
        let (mut out, mut var) = ([0u8; 64], BpVar { kind: 0, reserved: 0, index: 0 });
        check(unsafe { bp_prover_commit(self.raw, C::scalar_bytes(&v).as_ptr(), C::scalar_bytes(&v_blinding).as_ptr(), out.as_mut_ptr(), &mut var) })?;
        Ok((C::point_from(&out), Variable::from_raw(var)))
    }
    pub fn cs(&mut self) -> Cs<C> { Cs { raw: unsafe { bp_prover_cs(self.raw) }, _c: PhantomData } }
    /// `prove(self, prng, bp_gens)` (prover.rs:444): byte-identical proof for the same transcript, witness and RNG stream.
    pub fn prove<R: rand_core::RngCore + rand_core::CryptoRng>(self, prng: &mut R) -> Result<R1CSProof<C>, R1CSError> {
        let rng = RngAdapter::new(prng);
        let mut raw = ptr::null_mut();
        check(unsafe { bp_prover_prove(self.raw, rng.raw, &mut raw) })?;
        Ok(R1CSProof { raw, _c: PhantomData })
    }
}
impl<'g, 't, C: Curve> Drop for Prover<'g, 't, C> { fn drop(&mut self) { unsafe { bp_prover_free(self.raw) } } }

/// `r1cs::Verifier` (src/r1cs/verifier.rs:34-58).
pub struct Verifier<'t, C: Curve> { raw: *mut BpVerifier, _t: PhantomData<&'t mut Transcript>, _c: PhantomData<C> }
impl<'t, C: Curve> Verifier<'t, C> {
    pub fn new(ctx: &Context<C>, transcript: &'t mut Transcript) -> Result<Self, R1CSError> {
        let mut raw = ptr::null_mut();
        check(unsafe { bp_verifier_new(ctx.raw, transcript.raw, &mut raw) })?;
        Ok(Self { raw, _t: PhantomData, _c: PhantomData })
    }
    /// `commit(commitment) -> Variable` (verifier.rs:279-287). The point is validated (on curve, canonical, subgroup).
    pub fn commit(&mut self, commitment: C::Affine) -> Result<Variable, R1CSError> {
        let mut var = BpVar { kind: 0, reserved: 0, index: 0 };
        check(unsafe { bp_verifier_commit(self.raw, C::point_bytes(&commitment).as_ptr(), &mut var) })?;
        Ok(Variable::from_raw(var))
    }
    pub fn cs(&mut self) -> Cs<C> { Cs { raw: unsafe { bp_verifier_cs(self.raw) }, _c: PhantomData } }
    /// `verify(self, proof, pc_gens, bp_gens)` (verifier.rs:549-600).
    pub fn verify(self, proof: &R1CSProof<C>, gens: &BulletproofGens<C>) -> Result<(), R1CSError> {
        check(unsafe { bp_verifier_verify(self.raw, proof.raw, gens.raw) })
    }
}
impl<'t, C: Curve> Drop for Verifier<'t, C> { fn drop(&mut self) { unsafe { bp_verifier_free(self.raw) } } }

/// `batch_verify(prng, instances, pc_gens, bp_gens)` (src/r1cs/verifier.rs:604-691).
pub fn batch_verify<C: Curve, R: rand_core::RngCore + rand_core::CryptoRng>(
    ctx: &Context<C>, prng: &mut R, instances: Vec<(Verifier<C>, &R1CSProof<C>)>, gens: &BulletproofGens<C>,
) -> Result<(), R1CSError> {
    let rng = RngAdapter::new(prng);
    let vs: Vec<*mut BpVerifier> = instances.iter().map(|(v, _)| v.raw).collect();
    let ps: Vec<*const BpProof> = instances.iter().map(|(_, p)| p.raw as *const BpProof).collect();
    check(unsafe { bp_batch_verify(ctx.raw, rng.raw, vs.as_ptr(), ps.as_ptr(), vs.len(), gens.raw) })
}
