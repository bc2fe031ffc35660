//! Safe wrappers over `libb200zk.so` for a halo2_proofs fork (tag v2023_02_02), plus what the golden-vector dumper
//! needs: the ConstraintSystem -> blob serialiser and an rng wrapper that records every byte the prover draws.
//! Source only — see rust-shim/README.md.  Call sites in the reference: `/root/reference/src/circuits/utils.rs:28-63`.
pub mod dump;
pub mod sys;

use halo2_proofs::halo2curves::bn256::{Fr, G1Affine, G1};
use halo2_proofs::halo2curves::group::Curve;
use halo2_proofs::plonk::{Any, ConstraintSystem, Expression};
use rand_core::{CryptoRng, RngCore};
use std::ffi::CStr;
use std::os::raw::c_void;
use std::sync::OnceLock;

/// One context per process and device (`B200ZK_DEVICE`, default 0).  A ctx is used by one caller thread at a time:
/// halo2's prover calls best_multiexp / best_fft sequentially from one thread, rayon never wraps a commit.
pub struct Ctx(pub *mut sys::Ctx);
unsafe impl Send for Ctx {}
unsafe impl Sync for Ctx {}

pub fn ctx() -> *mut sys::Ctx {
    static CTX: OnceLock<Ctx> = OnceLock::new();
    CTX.get_or_init(|| {
        let dev: i32 = std::env::var("B200ZK_DEVICE").ok().and_then(|s| s.parse().ok()).unwrap_or(0);
        let mut p = std::ptr::null_mut();
        let rc = unsafe { sys::b200zk_ctx_create(dev, &mut p) };
        assert_eq!(rc, sys::B200ZK_OK, "b200zk_ctx_create({dev}) failed: {rc} (no sm_100 device? there is no CPU fallback)");
        Ctx(p)
    })
    .0
}

/// A non-zero status becomes the panic the upstream function would have raised.
pub fn check(rc: i32) {
    if rc != sys::B200ZK_OK {
        let msg = unsafe { CStr::from_ptr(sys::b200zk_last_error(ctx())) }.to_string_lossy().into_owned();
        panic!("b200zk error {rc}: {msg}");
    }
}

/// arithmetic::best_multiexp for bn256 (coeffs: &[Fr], bases: &[G1Affine]) -> G1
pub fn best_multiexp(coeffs: &[Fr], bases: &[G1Affine]) -> G1 {
    assert_eq!(coeffs.len(), bases.len());
    let mut out = G1::default();
    check(unsafe { sys::b200zk_msm(ctx(), coeffs.as_ptr() as *const c_void, bases.as_ptr() as *const c_void, coeffs.len(), &mut out as *mut G1 as *mut c_void) });
    out
}

/// arithmetic::best_fft for G = Fr
pub fn best_fft(a: &mut [Fr], omega: Fr, log_n: u32) {
    assert_eq!(a.len(), 1usize << log_n);
    check(unsafe { sys::b200zk_fft(ctx(), a.as_mut_ptr() as *mut c_void, &omega as *const Fr as *const c_void, log_n) });
}

/// ParamsKZG with the SRS resident on the device.
pub struct Params(pub *mut sys::Params, pub u32);
impl Params {
    /// ParamsKZG::setup: `s = Fr::random(rng)` stays on the Rust side, the 2n bases are generated on the device.
    pub fn setup(k: u32, s: &Fr) -> Self {
        let mut p = std::ptr::null_mut();
        check(unsafe { sys::b200zk_params_setup(ctx(), k, s as *const Fr as *const c_void, &mut p) });
        Params(p, k)
    }
    pub fn load(k: u32, g: &[G1Affine], g_lagrange: &[G1Affine]) -> Self {
        assert_eq!(g.len(), 1usize << k);
        assert_eq!(g_lagrange.len(), 1usize << k);
        let mut p = std::ptr::null_mut();
        check(unsafe { sys::b200zk_params_load(ctx(), k, g.as_ptr() as *const c_void, g_lagrange.as_ptr() as *const c_void, &mut p) });
        Params(p, k)
    }
    pub fn commit(&self, poly: &[Fr]) -> G1 {
        let mut out = G1::default();
        check(unsafe { sys::b200zk_commit(self.0, poly.as_ptr() as *const c_void, poly.len(), &mut out as *mut G1 as *mut c_void) });
        out
    }
    pub fn commit_lagrange(&self, poly: &[Fr]) -> G1 {
        let mut out = G1::default();
        check(unsafe { sys::b200zk_commit_lagrange(self.0, poly.as_ptr() as *const c_void, poly.len(), &mut out as *mut G1 as *mut c_void) });
        out
    }
    pub fn get_g(&self) -> (Vec<G1Affine>, Vec<G1Affine>) {
        let n = 1usize << self.1;
        let (mut g, mut gl) = (vec![G1Affine::default(); n], vec![G1Affine::default(); n]);
        check(unsafe { sys::b200zk_params_read(self.0, g.as_mut_ptr() as *mut c_void, gl.as_mut_ptr() as *mut c_void) });
        (g, gl)
    }
}
impl Drop for Params {
    fn drop(&mut self) {
        unsafe { sys::b200zk_params_destroy(self.0) }
    }
}

/// The proving key resident on the device (keygen_pk) and create_proof on it.
pub struct Pk(pub *mut sys::Pk);
impl Pk {
    /// `fixed`: pk.fixed_values (selectors compressed), `map_col` / `map_row`: permutation::keygen::Assembly::mapping (P x n)
    pub fn create(params: &Params, cs: &ConstraintSystem<Fr>, fixed: &[Vec<Fr>], map_col: &[u32], map_row: &[u32]) -> Self {
        let blob = cs_blob(cs, params.1);
        let cols: Vec<*const c_void> = fixed.iter().map(|c| c.as_ptr() as *const c_void).collect();
        let mut p = std::ptr::null_mut();
        check(unsafe { sys::b200zk_pk_create(params.0, blob.as_ptr(), blob.len(), cols.as_ptr(), map_col.as_ptr(), map_row.as_ptr(), &mut p) });
        Pk(p)
    }
    pub fn rng_draws(&self) -> usize {
        unsafe { sys::b200zk_pk_rng_draws(self.0) }
    }
    /// plonk::create_proof for one circuit: `advice` = the witness columns after batch_invert_assigned (unblinded),
    /// `rng_wide` = 64 bytes per Fr::random call in call order (rng.fill_bytes of 64 * rng_draws()), returns the bytes
    /// `transcript.finalize()` would.  Err(()) = Error::ConstraintSystemFailure / InstanceTooLarge upstream.
    pub fn create_proof(&self, advice: &[Vec<Fr>], instances: &[&[Fr]], rng_wide: &[u8], transcript_repr: &Fr) -> Result<Vec<u8>, ()> {
        assert!(rng_wide.len() >= 64 * self.rng_draws());
        let adv: Vec<*const c_void> = advice.iter().map(|c| c.as_ptr() as *const c_void).collect();
        let inst: Vec<*const c_void> = instances.iter().map(|c| c.as_ptr() as *const c_void).collect();
        let lens: Vec<u32> = instances.iter().map(|c| c.len() as u32).collect();
        let mut proof = vec![0u8; unsafe { sys::b200zk_pk_proof_size(self.0) }];
        let mut len = 0usize;
        let rc = unsafe {
            sys::b200zk_create_proof(self.0, adv.as_ptr(), inst.as_ptr(), lens.as_ptr(), rng_wide.as_ptr() as *const c_void,
                                     transcript_repr as *const Fr as *const c_void, proof.as_mut_ptr(), proof.len(), &mut len)
        };
        if rc == sys::B200ZK_ESYNTH {
            return Err(());
        }
        check(rc);
        proof.truncate(len);
        Ok(proof)
    }
}
impl Drop for Pk {
    fn drop(&mut self) {
        unsafe { sys::b200zk_pk_destroy(self.0) }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// ConstraintSystem -> the word format of halo2-experiments_b200/circuit.py::ConstraintSystem.to_blob (parsed by
// csrc/cs_desc.hpp::parse_cs).  Call it on the constraint system AFTER compress_selectors (vk.cs()).
const OP_CONST: u32 = 0;
const OP_FIXED: u32 = 1;
const OP_ADVICE: u32 = 2;
const OP_INSTANCE: u32 = 3;
const OP_NEG: u32 = 4;
const OP_ADD: u32 = 5;
const OP_MUL: u32 = 6;
const OP_SCALE: u32 = 7;

struct Emit {
    consts: Vec<Fr>,
    prog: Vec<u32>,
}
impl Emit {
    fn const_idx(&mut self, c: Fr) -> u32 {
        if let Some(i) = self.consts.iter().position(|x| *x == c) {
            return i as u32;
        }
        self.consts.push(c);
        (self.consts.len() - 1) as u32
    }
    /// postfix, children before the operator, left operand first (same order as circuit.py::emit)
    fn expr(&mut self, e: &Expression<Fr>) {
        match e {
            Expression::Constant(c) => {
                let i = self.const_idx(*c);
                self.prog.push(OP_CONST | (i << 8));
            }
            Expression::Selector(_) => panic!("selectors must be compressed into fixed columns before serialisation"),
            Expression::Fixed(q) => self.prog.push(OP_FIXED | ((q.index() as u32) << 8)),
            Expression::Advice(q) => self.prog.push(OP_ADVICE | ((q.index() as u32) << 8)),
            Expression::Instance(q) => self.prog.push(OP_INSTANCE | ((q.index() as u32) << 8)),
            Expression::Challenge(_) => panic!("multi-phase circuits are not supported by b200zk_create_proof"),
            Expression::Negated(a) => {
                self.expr(a);
                self.prog.push(OP_NEG);
            }
            Expression::Sum(a, b) => {
                self.expr(a);
                self.expr(b);
                self.prog.push(OP_ADD);
            }
            Expression::Product(a, b) => {
                self.expr(a);
                self.expr(b);
                self.prog.push(OP_MUL);
            }
            Expression::Scaled(a, c) => {
                self.expr(a);
                let i = self.const_idx(*c);
                self.prog.push(OP_SCALE | (i << 8));
            }
        }
    }
    fn range(&mut self, e: &Expression<Fr>) -> (u32, u32) {
        let off = self.prog.len() as u32;
        self.expr(e);
        (off, self.prog.len() as u32 - off)
    }
}

pub fn cs_blob(cs: &ConstraintSystem<Fr>, k: u32) -> Vec<u32> {
    let mut em = Emit { consts: vec![], prog: vec![] };
    let gates: Vec<(u32, u32)> = cs.gates().iter().flat_map(|g| g.polynomials().iter()).map(|p| em.range(p)).collect();
    let lookups: Vec<(Vec<(u32, u32)>, Vec<(u32, u32)>)> = cs
        .lookups()
        .iter()
        .map(|l| (l.input_expressions().iter().map(|e| em.range(e)).collect(), l.table_expressions().iter().map(|e| em.range(e)).collect()))
        .collect();
    let perm: Vec<(u32, u32)> = cs
        .permutation()
        .get_columns()
        .iter()
        .map(|c| (match c.column_type() { Any::Advice(_) => 0u32, Any::Fixed => 1, Any::Instance => 2 }, c.index() as u32))
        .collect();
    let mut w: Vec<u32> = vec![
        0x324B5A42, 1, k, cs.num_advice_columns() as u32, cs.num_fixed_columns() as u32, cs.num_instance_columns() as u32,
        cs.advice_queries().len() as u32, cs.fixed_queries().len() as u32, cs.instance_queries().len() as u32,
        gates.len() as u32, lookups.len() as u32, perm.len() as u32, em.consts.len() as u32, em.prog.len() as u32,
        cs.blinding_factors() as u32, cs.degree() as u32,
    ];
    for (c, r) in cs.advice_queries() { w.push(c.index() as u32); w.push(r.0 as u32); }
    for (c, r) in cs.fixed_queries() { w.push(c.index() as u32); w.push(r.0 as u32); }
    for (c, r) in cs.instance_queries() { w.push(c.index() as u32); w.push(r.0 as u32); }
    for (t, c) in &perm { w.push(*t); w.push(*c); }
    for (o, l) in &gates { w.push(*o); w.push(*l); }
    for (ins, tabs) in &lookups {
        w.push(ins.len() as u32);
        for (o, l) in ins.iter().chain(tabs.iter()) { w.push(*o); w.push(*l); }
    }
    for c in &em.consts {
        // canonical little-endian value, eight u32 words
        let repr = halo2_proofs::halo2curves::ff::PrimeField::to_repr(c);
        for chunk in repr.as_ref().chunks(4) { w.push(u32::from_le_bytes([chunk[0], chunk[1], chunk[2], chunk[3]])); }
    }
    w.extend_from_slice(&em.prog);
    w
}

// ---------------------------------------------------------------------------------------------------------------------
/// Wraps the prover's rng and keeps every byte it hands out: `Fr::random` consumes 64 bytes per call in halo2curves
/// (`from_bytes_wide` of `fill_bytes`, or 8 x `next_u64`: the same 64 bytes of a XorShiftRng), so the recorded stream is
/// exactly `rng_wide` of b200zk_create_proof.
pub struct RecordingRng<R: RngCore> {
    pub inner: R,
    pub bytes: Vec<u8>,
}
impl<R: RngCore> RecordingRng<R> {
    pub fn new(inner: R) -> Self {
        Self { inner, bytes: vec![] }
    }
}
impl<R: RngCore> RngCore for RecordingRng<R> {
    fn next_u32(&mut self) -> u32 {
        let v = self.inner.next_u32();
        self.bytes.extend_from_slice(&v.to_le_bytes());
        v
    }
    fn next_u64(&mut self) -> u64 {
        let v = self.inner.next_u64();
        self.bytes.extend_from_slice(&v.to_le_bytes());
        v
    }
    fn fill_bytes(&mut self, dest: &mut [u8]) {
        self.inner.fill_bytes(dest);
        self.bytes.extend_from_slice(dest);
    }
    fn try_fill_bytes(&mut self, dest: &mut [u8]) -> Result<(), rand_core::Error> {
        self.fill_bytes(dest);
        Ok(())
    }
}
impl<R: RngCore> CryptoRng for RecordingRng<R> {}

/// G1 -> affine for `transcript.write_point`, as `commit(..).to_affine()` upstream
pub fn to_affine(p: &G1) -> G1Affine {
    p.to_affine()
}
