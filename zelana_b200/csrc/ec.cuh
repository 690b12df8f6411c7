// BN254 G1 / G2 point arithmetic in extended Jacobian ("XYZZ") coordinates, generic over the
// coordinate field (Fq for G1, Fq2 for G2).  Device-side replacement for ark-ec 0.5.0
// short_weierstrass::{Affine, Projective} as used by VariableBaseMSM::msm_bigint and
// ark-groth16's create_proof_with_assignment (entered from the reference at
// core/src/sequencer/settlement/prover.rs:408).  Only affine results cross the C ABI, and the affine
// form of a group element is unique, so any complete addition law gives arkworks-identical outputs.
#pragma once
#include "fp.cuh"

namespace zkb {

// ---------------------------------------------------------------------------------- Fq2 = Fq[u]/(u^2+1)
struct Fq2 {
  Fq c0, c1;
  static __device__ __forceinline__ Fq2 zero() { return {Fq::zero(), Fq::zero()}; }
  static __device__ __forceinline__ Fq2 one() { return {Fq::one(), Fq::zero()}; }
  __device__ __forceinline__ bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
  __device__ __forceinline__ bool operator==(const Fq2& o) const { return c0 == o.c0 && c1 == o.c1; }
  __device__ __forceinline__ bool operator!=(const Fq2& o) const { return !(*this == o); }
  friend __device__ __forceinline__ Fq2 operator+(const Fq2& a, const Fq2& b) { return {a.c0 + b.c0, a.c1 + b.c1}; }
  friend __device__ __forceinline__ Fq2 operator-(const Fq2& a, const Fq2& b) { return {a.c0 - b.c0, a.c1 - b.c1}; }
  friend __device__ __forceinline__ Fq2 operator*(const Fq2& a, const Fq2& b) {
    // c0 = a0 b0 - a1 b1, c1 = a0 b1 + a1 b0: each ONE Montgomery pass over two products (Fp::mul_add_mul, lazy reduction):
    // 2 x 200 multiply-accumulates and no additions, against Karatsuba's 3 x 136 plus five modular additions (~445).
    return {Fq::mul_add_mul(a.c0, b.c0, Fq::modulus_minus(a.c1), b.c1), Fq::mul_add_mul(a.c0, b.c1, a.c1, b.c0)};
  }
  __device__ __forceinline__ Fq2 sqr() const {
    Fq t = c0 * c1;
    return {(c0 + c1) * (c0 - c1), t + t};
  }
  __device__ __forceinline__ Fq2 dbl() const { return {c0.dbl(), c1.dbl()}; }
  __device__ __forceinline__ Fq2 neg() const { return {c0.neg(), c1.neg()}; }
  __device__ Fq2 inverse() const {
    Fq d = (c0.sqr() + c1.sqr()).inverse();
    return {c0 * d, (c1 * d).neg()};
  }
  __device__ Fq2 inverse_vartime() const {
    Fq d = (c0.sqr() + c1.sqr()).inverse_vartime();
    return {c0 * d, (c1 * d).neg()};
  }
  __device__ __forceinline__ Fq2 to_mont() const { return {c0.to_mont(), c1.to_mont()}; }
  __device__ __forceinline__ Fq2 from_mont() const { return {c0.from_mont(), c1.from_mont()}; }
};

// a b - c d.  Fq: one Montgomery pass over both products (Fp::mul_add_mul with p - c: lazy reduction, 200 multiply-accumulates
// instead of 272) -- the last line of every XYZZ formula, y3 = R (Q - x3) - y1 PPP, is of this shape.  Fq2: the Karatsuba
// products already share work; nothing to gain.
__device__ __forceinline__ Fq mul_sub_mul(const Fq& a, const Fq& b, const Fq& c, const Fq& d) {
  return Fq::mul_add_mul(a, b, Fq::modulus_minus(c), d);   // c = 0 gives p: allowed (operands <= p)
}
__device__ __forceinline__ Fq2 mul_sub_mul(const Fq2& a, const Fq2& b, const Fq2& c, const Fq2& d) { return a * b - c * d; }

// ---------------------------------------------------------------------------------- points
// Affine point, coordinates in Montgomery form.  Infinity is (0, 0) -- not on y^2 = x^3 + b for b != 0 --
// matching the all-zero byte encoding of prover/src/bin/convert_vk.rs:165-171.
template <class F>
struct Affine {
  F x, y;
  __device__ __forceinline__ bool is_inf() const { return x.is_zero() && y.is_zero(); }
  static __device__ __forceinline__ Affine inf() { return {F::zero(), F::zero()}; }
};

template <class F>
struct XYZZ {
  F x, y, zz, zzz;  // (X/ZZ, Y/ZZZ), ZZ^3 = ZZZ^2; infinity <=> zz == 0

  static __device__ __forceinline__ XYZZ inf() { return {F::zero(), F::zero(), F::zero(), F::zero()}; }
  __device__ __forceinline__ bool is_inf() const { return zz.is_zero(); }

  static __device__ __forceinline__ XYZZ from_affine(const Affine<F>& p) {
    if (p.is_inf()) return inf();
    return {p.x, p.y, F::one(), F::one()};
  }

  // 2 * affine (mdbl-2008-s-1, a = 0).  Not inlined: inside madd it is the (practically never taken) P == Q branch, and inlining
  // its seven products there bloats the hot loop's code (instruction-fetch stalls, visible in the G2 accumulate profile).
  static __device__ __noinline__ XYZZ dbl_affine(const Affine<F>& p) {
    if (p.is_inf() || p.y.is_zero()) return inf();
    F U = p.y.dbl();
    F V = U.sqr();
    F W = U * V;
    F S = p.x * V;
    F xx = p.x.sqr();
    F M = xx.dbl() + xx;
    XYZZ r;
    r.x = M.sqr() - S.dbl();
    r.y = mul_sub_mul(M, S - r.x, W, p.y);
    r.zz = V;
    r.zzz = W;
    return r;
  }

  // dbl-2008-s-1, a = 0.  Not inlined (like add below): both are used only by the cold reduction / table kernels, where a call
  // costs ~1 % and inlining them at every site multiplies the compile time of the G2 instantiation.
  __device__ __noinline__ XYZZ dbl() const {
    if (is_inf() || y.is_zero()) return inf();
    F U = y.dbl();
    F V = U.sqr();
    F W = U * V;
    F S = x * V;
    F xx = x.sqr();
    F M = xx.dbl() + xx;
    XYZZ r;
    r.x = M.sqr() - S.dbl();
    r.y = mul_sub_mul(M, S - r.x, W, y);
    r.zz = V * zz;
    r.zzz = W * zzz;
    return r;
  }

  // this += q (q affine, not infinity-checked by the caller): madd-2008-s, 8M + 2S
  __device__ __forceinline__ void madd(const Affine<F>& q) {
    if (q.is_inf()) return;
    if (is_inf()) {
      *this = from_affine(q);
      return;
    }
    F U2 = q.x * zz;
    F S2 = q.y * zzz;
    F P = U2 - x;
    F R = S2 - y;
    if (P.is_zero()) {
      if (R.is_zero()) *this = dbl_affine(q);
      else *this = inf();
      return;
    }
    F PP = P.sqr();
    F PPP = P * PP;
    F Q = x * PP;
    F X3 = R.sqr() - PPP - Q.dbl();
    y = mul_sub_mul(R, Q - X3, y, PPP);
    x = X3;
    zz = zz * PP;
    zzz = zzz * PPP;
  }

  // this += q: add-2008-s, 12M + 2S
  __device__ __noinline__ void add(const XYZZ& q) {
    if (q.is_inf()) return;
    if (is_inf()) {
      *this = q;
      return;
    }
    F U1 = x * q.zz;
    F U2 = q.x * zz;
    F S1 = y * q.zzz;
    F S2 = q.y * zzz;
    F P = U2 - U1;
    F R = S2 - S1;
    if (P.is_zero()) {
      if (R.is_zero()) *this = dbl();
      else *this = inf();
      return;
    }
    F PP = P.sqr();
    F PPP = P * PP;
    F Q = U1 * PP;
    F X3 = R.sqr() - PPP - Q.dbl();
    y = mul_sub_mul(R, Q - X3, S1, PPP);
    x = X3;
    zz = zz * q.zz * PP;
    zzz = zzz * q.zzz * PPP;
  }

  __device__ __forceinline__ XYZZ neg() const { return {x, y.neg(), zz, zzz}; }

  // k * this for a small unsigned k (MSB-first double-and-add)
  __device__ XYZZ mul_u32(uint32_t k) const {
    XYZZ r = inf();
    for (int bit = 31; bit >= 0; bit--) {
      r = r.dbl();
      if ((k >> bit) & 1u) r.add(*this);
    }
    return r;
  }

  // scalar given as 8 LE canonical words
  __device__ XYZZ mul_words(const uint32_t* k) const {
    XYZZ r = inf();
    for (int i = 7; i >= 0; i--)
      for (int bit = 31; bit >= 0; bit--) {
        r = r.dbl();
        if ((k[i] >> bit) & 1u) r.add(*this);
      }
    return r;
  }

  // same with the variable-time inversion: single-thread tails only
  __device__ Affine<F> to_affine_vartime() const {
    if (is_inf()) return Affine<F>::inf();
    F zi3 = zzz.inverse_vartime();
    F zi = zi3 * zz;
    F zi2 = zi.sqr();
    return {x * zi2, y * zi3};
  }

  __device__ Affine<F> to_affine() const {
    if (is_inf()) return Affine<F>::inf();
    // 1/zzz ; x/zz = x * zzz^2 / zz^3 ... simpler: zi3 = 1/zzz, zi2 = (zi3 * zz)^2 since zz^3 = zzz^2
    F zi3 = zzz.inverse();
    F zi = zi3 * zz;  // = zz / zzz = 1 / z
    F zi2 = zi.sqr();
    return {x * zi2, y * zi3};
  }
};

using G1Affine = Affine<Fq>;
using G2Affine = Affine<Fq2>;
using G1XYZZ = XYZZ<Fq>;
using G2XYZZ = XYZZ<Fq2>;

}  // namespace zkb
