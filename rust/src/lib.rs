//! Thin Rust binding of include/tr_prover.h and drop-in replacements for the halo2_proofs 0.2.0 routines the
//! TinyRAM prover spends its time in (`arithmetic::best_multiexp`, `arithmetic::best_fft`, the
//! `EvaluationDomain` transforms).  See INTEGRATION.md for how the halo2 fork calls these.
//!
//! UNTESTED: the build image has no Rust toolchain.  The C ABI underneath is what the parity tests exercise.
#![allow(non_camel_case_types)]
use std::{ffi::CStr, os::raw::{c_char, c_int, c_uint, c_void}, ptr, sync::Mutex};

use group::{Curve, prime::PrimeCurveAffine};
use pasta_curves::arithmetic::{CurveAffine, FieldExt};

#[repr(C)] pub struct trp_ctx { _p: [u8; 0] }
#[repr(C)] pub struct trp_bases { _p: [u8; 0] }
#[repr(C)] pub struct trp_domain { _p: [u8; 0] }

pub const TRP_CURVE_PALLAS: c_int = 0;
pub const TRP_CURVE_VESTA: c_int = 1;

extern "C" {
    pub fn trp_ctx_create(out: *mut *mut trp_ctx, device: c_int, curve: c_int) -> c_int;
    pub fn trp_ctx_destroy(ctx: *mut trp_ctx);
    pub fn trp_last_error(ctx: *const trp_ctx) -> *const c_char;
    pub fn trp_ctx_sync(ctx: *mut trp_ctx) -> c_int;
    pub fn trp_bases_load(ctx: *mut trp_ctx, affine_xy: *const u64, n: usize, out: *mut *mut trp_bases) -> c_int;
    pub fn trp_bases_free(b: *mut trp_bases);
    pub fn trp_msm(ctx: *mut trp_ctx, bases: *const trp_bases, scalars: *const u64, n: usize, out: *mut u64) -> c_int;
    pub fn trp_msm_batch(ctx: *mut trp_ctx, bases: *const trp_bases, scalars: *const u64, n: usize, m: usize, out: *mut u64) -> c_int;
    pub fn trp_ntt(ctx: *mut trp_ctx, a: *mut u64, batch: usize, log_n: c_uint, omega: *const u64) -> c_int;
    pub fn trp_domain_create(ctx: *mut trp_ctx, k: c_uint, j: c_uint, out: *mut *mut trp_domain) -> c_int;
    pub fn trp_domain_free(d: *mut trp_domain);
    pub fn trp_lagrange_to_coeff(d: *mut trp_domain, cols: *mut u64, batch: usize) -> c_int;
    pub fn trp_coeff_to_extended(d: *mut trp_domain, coeff: *const u64, ext: *mut u64, batch: usize) -> c_int;
    pub fn trp_extended_to_coeff(d: *mut trp_domain, ext: *mut u64, out_coeff: *mut u64, divide: c_int) -> c_int;
    pub fn trp_quotient_eval(d: *mut trp_domain, prog: *const u32, prog_len: usize, consts: *const u64, n_consts: usize,
                             cols: *const *const u64, n_cols: usize, out_ext: *mut u64) -> c_int;
    // rows f1-f3 of the scope table: the callers either side of the hot kernels (host-pointer forms)
    pub fn trp_batch_invert(ctx: *mut trp_ctx, which_field: c_int, a: *mut u64, n: usize) -> c_int;
    pub fn trp_permutation_product(d: *mut trp_domain, values: *const *const u64, sigmas: *const *const u64, m: usize, beta: *const u64,
                                   gamma: *const u64, delta_beta: *const u64, last_z: *const u64, z: *mut u64) -> c_int;
    pub fn trp_lookup_product(d: *mut trp_domain, input: *const u64, table: *const u64, perm_input: *const u64, perm_table: *const u64,
                              beta: *const u64, gamma: *const u64, z: *mut u64, n_out: usize) -> c_int;
    pub fn trp_permute_expression_pair(ctx: *mut trp_ctx, input: *const u64, table: *const u64, rows: usize, perm_input: *mut u64,
                                       perm_table: *mut u64, all_found: *mut c_int) -> c_int;
    pub fn trp_eval_polynomial(ctx: *mut trp_ctx, which_field: c_int, coeffs: *const u64, n: usize, x: *const u64, out: *mut u64) -> c_int;
    // device-resident openings (include/tr_prover.h): m separately allocated polynomials at one point / one linear combination of them
    pub fn trp_dev_eval_polynomials_at(ctx: *mut trp_ctx, which_field: c_int, d_poly_ptrs: *const *const u64, n: usize, m: usize, x: *const u64, d_out: *mut u64) -> c_int;
    pub fn trp_dev_linear_combination(ctx: *mut trp_ctx, which_field: c_int, d_poly_ptrs: *const *const u64, scalars: *const u64, n: usize, m: usize, d_out: *mut u64) -> c_int;
    pub fn trp_compute_inner_product(ctx: *mut trp_ctx, which_field: c_int, a: *const u64, b: *const u64, n: usize, out: *mut u64) -> c_int;
    pub fn trp_kate_division(ctx: *mut trp_ctx, which_field: c_int, coeffs: *const u64, n: usize, b: *const u64, q: *mut u64) -> c_int;
    /// Params::new(k): g, g_lagrange (2^k affine points each, 8 x u64), w, u
    pub fn trp_params_new(ctx: *mut trp_ctx, k: c_uint, g: *mut u64, g_lagrange: *mut u64, w: *mut u64, u: *mut u64) -> c_int;
    pub fn trp_hash_to_curve(ctx: *mut trp_ctx, domain_prefix: *const c_char, messages: *const u8, msg_len: usize, n: usize, out: *mut u64) -> c_int;
    pub fn trp_group_fft(ctx: *mut trp_ctx, points: *mut u64, log_n: c_uint, omega: *const u64, scale: *const u64) -> c_int;
}

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
