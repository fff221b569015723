//! Thin Rust binding of include/tr_prover.h and drop-in replacements for the halo2_proofs 0.2.0 routines the
//! TinyRAM prover spends its time in (`arithmetic::best_multiexp`, `arithmetic::best_fft`, the
//! `EvaluationDomain` transforms).  See INTEGRATION.md for how the halo2 fork calls these.
//!
//! NOT COMPILED in the build image (no Rust toolchain): the C ABI underneath is what the parity tests
//! exercise.  What IS checked without rustc (tests/test_rust_shim_cpu.py): `ffi.rs` is regenerated from the header and
//! diffed, every symbol it declares is exported by libtrp.so, and build.rs compiles the units the Makefile links.
#![allow(non_camel_case_types)]
use std::{ffi::CStr, os::raw::{c_int, c_void}, ptr, sync::Mutex};

use group::{Curve, prime::PrimeCurveAffine};
use pasta_curves::arithmetic::{CurveAffine, FieldExt};

mod ffi;          // GENERATED from include/tr_prover.h by rust/gen_bindings.py: every entry point, argument for argument
pub use ffi::*;

/// One context per (device, curve); halo2 calls are synchronous, so a process-wide handle behind a mutex is enough.
pub struct Backend { ctx: *mut trp_ctx }
unsafe impl Send for Backend {}

impl Backend {
    pub fn new(device: i32, curve: c_int) -> Self {
        let mut ctx = ptr::null_mut();
        let rc = unsafe { trp_ctx_create(&mut ctx, device, curve) };
        assert_eq!(rc, 0, "trp_ctx_create failed ({rc}): no CUDA device / bad curve -- there is no CPU fallback");
        Backend { ctx }
    }
    fn check(&self, rc: c_int) {
        if rc != 0 {
            let msg = unsafe { CStr::from_ptr(trp_last_error(self.ctx)) }.to_string_lossy().into_owned();
            panic!("tr_prover: {msg} ({rc})");   // halo2's own behaviour on misuse is a panic (assert_eq!)
        }
    }
}
impl Drop for Backend { fn drop(&mut self) { unsafe { trp_ctx_destroy(self.ctx) } } }

/// pasta's `Fp`/`Fq` are `#[repr(transparent)]` wrappers of `[u64; 4]` in Montgomery form, so a scalar slice
/// crosses the boundary by pointer cast.  Affine points are `repr(Rust)`: marshal into the explicit 64-byte layout.
pub fn marshal_bases<C: CurveAffine>(bases: &[C]) -> Vec<u64> {
    let mut out = vec![0u64; bases.len() * 8];
    for (i, b) in bases.iter().enumerate() {
        if let Some(c) = Option::<pasta_curves::arithmetic::Coordinates<C>>::from(b.coordinates()) {
            // SAFETY: C::Base is Fp or Fq = repr(transparent) [u64; 4]
            let x: [u64; 4] = unsafe { std::mem::transmute_copy(c.x()) };
            let y: [u64; 4] = unsafe { std::mem::transmute_copy(c.y()) };
            out[8 * i..8 * i + 4].copy_from_slice(&x);
            out[8 * i + 4..8 * i + 8].copy_from_slice(&y);
        } // identity stays x = y = 0
    }
    out
}

/// Device-resident `Params.g` / `Params.g_lagrange ++ [w]`, uploaded once per `Params`.
pub struct Bases { h: *mut trp_bases, pub len: usize }
impl Bases {
    pub fn load<C: CurveAffine>(be: &Backend, bases: &[C]) -> Self {
        let xy = marshal_bases(bases);
        let mut h = ptr::null_mut();
        be.check(unsafe { trp_bases_load(be.ctx, xy.as_ptr(), bases.len(), &mut h) });
        Bases { h, len: bases.len() }
    }
}
impl Drop for Bases { fn drop(&mut self) { unsafe { trp_bases_free(self.h) } } }

/// == halo2_proofs::arithmetic::best_multiexp(coeffs, bases)
pub fn best_multiexp<C: CurveAffine>(be: &Backend, coeffs: &[C::Scalar], bases: &Bases) -> C::Curve {
    assert!(coeffs.len() <= bases.len);
    let mut out = [0u64; 12];
    be.check(unsafe { trp_msm(be.ctx, bases.h, coeffs.as_ptr() as *const u64, coeffs.len(), out.as_mut_ptr()) });
    if out[8..12] == [0, 0, 0, 0] { return C::Curve::identity(); }
    // result is normalised (z = 1): rebuild the affine point, then lift
    let x: C::Base = unsafe { std::mem::transmute_copy(&[out[0], out[1], out[2], out[3]]) };
    let y: C::Base = unsafe { std::mem::transmute_copy(&[out[4], out[5], out[6], out[7]]) };
    C::from_xy(x, y).unwrap().to_curve()
}

/// == halo2_proofs::arithmetic::best_fft(a, omega, log_n) for field elements
pub fn best_fft<F: FieldExt>(be: &Backend, a: &mut [F], omega: F, log_n: u32) {
    assert_eq!(a.len(), 1 << log_n);
    let om: [u64; 4] = unsafe { std::mem::transmute_copy(&omega) };
    be.check(unsafe { trp_ntt(be.ctx, a.as_mut_ptr() as *mut u64, 1, log_n, om.as_ptr()) });
}

/// Process-wide backend used by the patched halo2_proofs (Vesta commitments over Fp, src/test_utils.rs:12,21).
pub static VESTA: Mutex<Option<Backend>> = Mutex::new(None);
pub fn vesta() -> std::sync::MutexGuard<'static, Option<Backend>> {
    let mut g = VESTA.lock().unwrap();
    if g.is_none() { *g = Some(Backend::new(0, TRP_CURVE_VESTA)); }
    g
}
#[allow(dead_code)] fn _unused(_: *mut c_void) {}
