// K2: Pallas / Vesta group arithmetic (y^2 = x^3 + 5, a = 0) for the MSM kernels.
//
// Replaces pasta_curves::curves::{Ep,Eq,EpAffine,EqAffine} (pasta_curves 0.4.1, Cargo.lock:847-849; the
// commitment curve is selected at /root/reference/src/test_utils.rs:12,21).  The CPU crate works in Jacobian
// coordinates; here bucket accumulators use extended Jacobian XYZZ (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2)
// because the mixed add is 8M+2S and needs no Z inversion tricks.  Group results are unique, so the
// coordinate system is invisible after normalisation to affine.
//
// C-ABI affine layout: { x[8 x u32], y[8 x u32] } Montgomery; identity encoded as x = y = 0 (not on the curve).
#pragma once
#include "ff.cuh"

namespace ec {
using namespace ff;

template <class PR> struct Affine { Fe<PR> x, y; };
template <class PR> struct XYZZ { Fe<PR> x, y, zz, zzz; };

template <class PR> FF_HD bool affine_is_identity(const Affine<PR>& p) { return fe_is_zero(p.x) && fe_is_zero(p.y); }
template <class PR> FF_HD XYZZ<PR> xyzz_identity() {
  XYZZ<PR> r; r.x = fe_zero<PR>(); r.y = fe_zero<PR>(); r.zz = fe_zero<PR>(); r.zzz = fe_zero<PR>(); return r;
}
template <class PR> FF_HD bool xyzz_is_identity(const XYZZ<PR>& p) { return fe_is_zero(p.zz); }
template <class PR> FF_HD XYZZ<PR> xyzz_from_affine(const Affine<PR>& p) {
  XYZZ<PR> r;
  if (affine_is_identity(p)) return xyzz_identity<PR>();
  r.x = p.x; r.y = p.y; r.zz = fe_one<PR>(); r.zzz = fe_one<PR>();
  return r;
}

// dbl-2008-s-1 (a = 0)
template <class PR> FF_HD void xyzz_dbl(XYZZ<PR>& p) {
  if (xyzz_is_identity(p)) return;
  Fe<PR> u = fe_dbl(p.y);
  Fe<PR> v = fe_sqr(u);
  Fe<PR> w = fe_mul(u, v);
  Fe<PR> s = fe_mul(p.x, v);
  Fe<PR> xx = fe_sqr(p.x);
  Fe<PR> m = fe_add(fe_dbl(xx), xx);
  Fe<PR> x3 = fe_sub(fe_sqr(m), fe_dbl(s));
  Fe<PR> y3 = fe_sub(fe_mul(m, fe_sub(s, x3)), fe_mul(w, p.y));
  p.x = x3; p.y = y3;
  p.zz = fe_mul(v, p.zz);
  p.zzz = fe_mul(w, p.zzz);
}

// mdbl-2008-s-1: 2 * (affine, not identity)
template <class PR> FF_HD XYZZ<PR> xyzz_dbl_affine(const Affine<PR>& q) {
  XYZZ<PR> r;
  Fe<PR> u = fe_dbl(q.y);
  r.zz = fe_sqr(u);
  r.zzz = fe_mul(u, r.zz);
  Fe<PR> s = fe_mul(q.x, r.zz);
  Fe<PR> xx = fe_sqr(q.x);
  Fe<PR> m = fe_add(fe_dbl(xx), xx);
  r.x = fe_sub(fe_sqr(m), fe_dbl(s));
  r.y = fe_sub(fe_mul(m, fe_sub(s, r.x)), fe_mul(r.zzz, q.y));
  return r;
}

// madd-2008-s: p += q (q affine); all special cases handled exactly
template <class PR> FF_HD void xyzz_add_mixed(XYZZ<PR>& p, const Affine<PR>& q) {
  if (affine_is_identity(q)) return;
  if (xyzz_is_identity(p)) { p.x = q.x; p.y = q.y; p.zz = fe_one<PR>(); p.zzz = fe_one<PR>(); return; }
  Fe<PR> u2 = fe_mul(q.x, p.zz);
  Fe<PR> s2 = fe_mul(q.y, p.zzz);
  Fe<PR> pp_ = fe_sub(u2, p.x);
  Fe<PR> r = fe_sub(s2, p.y);
  if (fe_is_zero(pp_)) {
    if (fe_is_zero(r)) p = xyzz_dbl_affine(q);
    else p = xyzz_identity<PR>();
    return;
  }
  Fe<PR> pp = fe_sqr(pp_);
  Fe<PR> ppp = fe_mul(pp_, pp);
  Fe<PR> qv = fe_mul(p.x, pp);
  Fe<PR> x3 = fe_sub(fe_sub(fe_sqr(r), ppp), fe_dbl(qv));
  Fe<PR> y3 = fe_sub(fe_mul(r, fe_sub(qv, x3)), fe_mul(p.y, ppp));
  p.x = x3; p.y = y3;
  p.zz = fe_mul(p.zz, pp);
  p.zzz = fe_mul(p.zzz, ppp);
}

// add-2008-s: p += q (both XYZZ)
template <class PR> FF_HD void xyzz_add(XYZZ<PR>& p, const XYZZ<PR>& q) {
  if (xyzz_is_identity(q)) return;
  if (xyzz_is_identity(p)) { p = q; return; }
  Fe<PR> u1 = fe_mul(p.x, q.zz);
  Fe<PR> u2 = fe_mul(q.x, p.zz);
  Fe<PR> s1 = fe_mul(p.y, q.zzz);
  Fe<PR> s2 = fe_mul(q.y, p.zzz);
  Fe<PR> pp_ = fe_sub(u2, u1);
  Fe<PR> r = fe_sub(s2, s1);
  if (fe_is_zero(pp_)) {
    if (fe_is_zero(r)) xyzz_dbl(p);
    else p = xyzz_identity<PR>();
    return;
  }
  Fe<PR> pp = fe_sqr(pp_);
  Fe<PR> ppp = fe_mul(pp_, pp);
  Fe<PR> qv = fe_mul(u1, pp);
  Fe<PR> x3 = fe_sub(fe_sub(fe_sqr(r), ppp), fe_dbl(qv));
  Fe<PR> y3 = fe_sub(fe_mul(r, fe_sub(qv, x3)), fe_mul(s1, ppp));
  p.x = x3; p.y = y3;
  p.zz = fe_mul(fe_mul(p.zz, q.zz), pp);
  p.zzz = fe_mul(fe_mul(p.zzz, q.zzz), ppp);
}

template <class PR> FF_HD XYZZ<PR> xyzz_neg(const XYZZ<PR>& p) { XYZZ<PR> r = p; r.y = fe_neg(p.y); return r; }

// normalise: x = X/ZZ, y = Y/ZZZ with one inversion; identity -> (0, 0)
template <class PR> FF_HD Affine<PR> xyzz_to_affine(const XYZZ<PR>& p) {
  Affine<PR> a;
  if (xyzz_is_identity(p)) { a.x = fe_zero<PR>(); a.y = fe_zero<PR>(); return a; }
  Fe<PR> inv = fe_inv(fe_mul(p.zz, p.zzz));
  a.x = fe_mul(p.x, fe_mul(inv, p.zzz));
  a.y = fe_mul(p.y, fe_mul(inv, p.zz));
  return a;
}

}  // namespace ec
