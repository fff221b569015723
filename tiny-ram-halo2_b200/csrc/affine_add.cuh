// Affine point addition split around ONE shared inversion (Montgomery's trick): the building block of a batched-affine stage of
// the MSM bucket accumulation (DESIGN.md section 9 costs that stage; NO kernel uses this header yet -- it is exercised on the
// host only, tests/test_ff_host.py).  P1 + P2 on y^2 = x^3 + 5 costs
// 1 inversion + 2M + 1S in affine coordinates; with the inversion shared by a batch it is 5M + 1S + the batch's share,
// against 8M + 2S for the mixed XYZZ add of ec.cuh.
//
//   pair_classify : which formula the pair needs and the value whose inverse that formula uses
//   pair_finish   : the sum, given that inverse
// Every special case is exact (identity operands, P + P, P + (-P)): the group element is what bit-parity is about.
// Host/device code like ff.cuh / ec.cuh, so tests/test_ff_host.py exercises it on the CPU box.
#pragma once
#include "ec.cuh"

namespace ec {

enum PairCase : int { PAIR_NONE = 0, PAIR_COPY1 = 1, PAIR_COPY2 = 2, PAIR_ADD = 3, PAIR_DBL = 4 };   // >= PAIR_ADD needs an inverse

template <class PR> FF_HD int pair_classify(const Affine<PR>& p1, const Affine<PR>& p2, Fe<PR>& den) {
  const bool id1 = affine_is_identity(p1), id2 = affine_is_identity(p2);
  if (id1 && id2) return PAIR_NONE;
  if (id1) return PAIR_COPY2;
  if (id2) return PAIR_COPY1;
  den = fe_sub(p2.x, p1.x);
  if (!fe_is_zero(den)) return PAIR_ADD;
  if (fe_eq(p1.y, p2.y)) { den = fe_dbl(p1.y); return PAIR_DBL; }    // y != 0: the group has odd order, no 2-torsion
  return PAIR_NONE;                                                  // P + (-P)
}

template <class PR> FF_HD Affine<PR> pair_finish(int cs, const Affine<PR>& p1, const Affine<PR>& p2, const Fe<PR>& dinv) {
  Affine<PR> r;
  if (cs == PAIR_COPY1) return p1;
  if (cs == PAIR_COPY2) return p2;
  if (cs == PAIR_NONE) { r.x = fe_zero<PR>(); r.y = fe_zero<PR>(); return r; }
  Fe<PR> lambda, x3;
  if (cs == PAIR_ADD) {
    lambda = fe_mul(fe_sub(p2.y, p1.y), dinv);
    x3 = fe_sub(fe_sub(fe_sqr(lambda), p1.x), p2.x);
  } else {                                                           // tangent: lambda = 3 x^2 / (2 y)
    Fe<PR> xx = fe_sqr(p1.x);
    lambda = fe_mul(fe_add(fe_dbl(xx), xx), dinv);
    x3 = fe_sub(fe_sqr(lambda), fe_dbl(p1.x));
  }
  r.x = x3;
  r.y = fe_sub(fe_mul(lambda, fe_sub(p1.x, x3)), p1.y);
  return r;
}

}  // namespace ec
