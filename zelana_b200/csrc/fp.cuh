// BN254 prime-field arithmetic for sm_100a: 8 x 32-bit limbs, Montgomery form with R = 2^256.
//
// Replaces (on device) ark-ff 0.5.0 `Fp<MontBackend<_,4>>` for ark-bn254 Fr / Fq -- the inner loop of
// every routine below reference call site core/src/sequencer/settlement/prover.rs:408.  Same Montgomery
// radix as arkworks, so a value's eight limbs here are bit-identical to arkworks' four u64 limbs.
//
// Multiplication is an operand-scanning Montgomery product on two interleaved accumulators (one
// holding the 64-bit products that start on even 32-bit columns, one those that start on odd columns)
// so that every 32x32->64 multiply-accumulate is ONE `IMAD.WIDE.U32[.X]` with the carry riding in a
// predicate: 8*(8+8) = 128 wide multiply-adds + 8 low multiplies for the quotient digits = 136 per
// product (the figure SURVEY.md section 8d normalises the MSM roofline with).
#pragma once
#include <cstdint>

namespace zkb {

struct FqCfg {
  // onchain-programs/verifier/programs/onchain_verifier/src/lib.rs:9-10
  static constexpr uint32_t M0 = 0xd87cfd47u, M1 = 0x3c208c16u, M2 = 0x6871ca8du, M3 = 0x97816a91u,
                            M4 = 0x8181585du, M5 = 0xb85045b6u, M6 = 0xe131a029u, M7 = 0x30644e72u;
  static constexpr uint32_t INV = 0xe4866389u;  // -p^-1 mod 2^32
  static __host__ __device__ constexpr uint32_t r1(int i) {
    constexpr uint32_t t[8] = {0xc58f0d9du, 0xd35d438du, 0xf5c70b3du, 0x0a78eb28u,
                                     0x7879462cu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
    return t[i];
  }
  static __host__ __device__ constexpr uint32_t r2(int i) {
    constexpr uint32_t t[8] = {0x538afa89u, 0xf32cfc5bu, 0xd44501fbu, 0xb5e71911u,
                                     0x0a417ff6u, 0x47ab1effu, 0xcab8351fu, 0x06d89f71u};
    return t[i];
  }
};

struct FrCfg {
  // forge/crates/prover-worker/src/prover.rs:19-20
  static constexpr uint32_t M0 = 0xf0000001u, M1 = 0x43e1f593u, M2 = 0x79b97091u, M3 = 0x2833e848u,
                            M4 = 0x8181585du, M5 = 0xb85045b6u, M6 = 0xe131a029u, M7 = 0x30644e72u;
  static constexpr uint32_t INV = 0xefffffffu;
  static __host__ __device__ constexpr uint32_t r1(int i) {
    constexpr uint32_t t[8] = {0x4ffffffbu, 0xac96341cu, 0x9f60cd29u, 0x36fc7695u,
                                     0x7879462eu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
    return t[i];
  }
  static __host__ __device__ constexpr uint32_t r2(int i) {
    constexpr uint32_t t[8] = {0xae216da7u, 0x1bb8e645u, 0xe35c59e3u, 0x53fe3ab1u,
                                     0x53bb8085u, 0x8c49833du, 0x7f4e44a5u, 0x0216d0b1u};
    return t[i];
  }
};

template <class C>
struct alignas(16) Fp {
  uint32_t v[8];

  // ---- constants -------------------------------------------------------------------------------
  static __device__ __forceinline__ Fp zero() {
    Fp r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = 0;
    return r;
  }
  static __device__ __forceinline__ Fp one() {  // Montgomery 1 = R mod p
    Fp r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = C::r1(i);
    return r;
  }
  static __device__ __forceinline__ Fp r2() {
    Fp r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = C::r2(i);
    return r;
  }
  static __device__ __forceinline__ Fp modulus() {
    Fp r;
    r.v[0] = C::M0; r.v[1] = C::M1; r.v[2] = C::M2; r.v[3] = C::M3;
    r.v[4] = C::M4; r.v[5] = C::M5; r.v[6] = C::M6; r.v[7] = C::M7;
    return r;
  }

  __device__ __forceinline__ bool is_zero() const {
    return (v[0] | v[1] | v[2] | v[3] | v[4] | v[5] | v[6] | v[7]) == 0;
  }
  __device__ __forceinline__ bool operator==(const Fp& o) const {
    uint32_t d = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) d |= v[i] ^ o.v[i];
    return d == 0;
  }
  __device__ __forceinline__ bool operator!=(const Fp& o) const { return !(*this == o); }

  // ---- r = a - p if a >= p (a < 2p) --------------------------------------------------------------
  static __device__ __forceinline__ void reduce_once(Fp& a) {
    uint32_t t0, t1, t2, t3, t4, t5, t6, t7, brw;
    asm("sub.cc.u32 %0, %9, %17;\n\t"
        "subc.cc.u32 %1, %10, %18;\n\t"
        "subc.cc.u32 %2, %11, %19;\n\t"
        "subc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\t"
        "subc.cc.u32 %5, %14, %22;\n\t"
        "subc.cc.u32 %6, %15, %23;\n\t"
        "subc.cc.u32 %7, %16, %24;\n\t"
        "subc.u32 %8, 0, 0;"
        : "=r"(t0), "=r"(t1), "=r"(t2), "=r"(t3), "=r"(t4), "=r"(t5), "=r"(t6), "=r"(t7), "=r"(brw)
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]),
          "r"(a.v[7]), "n"(C::M0), "n"(C::M1), "n"(C::M2), "n"(C::M3), "n"(C::M4), "n"(C::M5),
          "n"(C::M6), "n"(C::M7));
    if (brw == 0) {  // no borrow: a >= p
      a.v[0] = t0; a.v[1] = t1; a.v[2] = t2; a.v[3] = t3;
      a.v[4] = t4; a.v[5] = t5; a.v[6] = t6; a.v[7] = t7;
    }
  }

  friend __device__ __forceinline__ Fp operator+(const Fp& a, const Fp& b) {
    Fp r;
    asm("add.cc.u32 %0, %8, %16;\n\t"
        "addc.cc.u32 %1, %9, %17;\n\t"
        "addc.cc.u32 %2, %10, %18;\n\t"
        "addc.cc.u32 %3, %11, %19;\n\t"
        "addc.cc.u32 %4, %12, %20;\n\t"
        "addc.cc.u32 %5, %13, %21;\n\t"
        "addc.cc.u32 %6, %14, %22;\n\t"
        "addc.u32 %7, %15, %23;"
        : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]),
          "=r"(r.v[6]), "=r"(r.v[7])
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]),
          "r"(a.v[7]), "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]),
          "r"(b.v[6]), "r"(b.v[7]));
    reduce_once(r);  // a + b < 2p < 2^255: no carry out of 256 bits
    return r;
  }

  friend __device__ __forceinline__ Fp operator-(const Fp& a, const Fp& b) {
    Fp r;
    uint32_t brw;
    asm("sub.cc.u32 %0, %9, %17;\n\t"
        "subc.cc.u32 %1, %10, %18;\n\t"
        "subc.cc.u32 %2, %11, %19;\n\t"
        "subc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\t"
        "subc.cc.u32 %5, %14, %22;\n\t"
        "subc.cc.u32 %6, %15, %23;\n\t"
        "subc.cc.u32 %7, %16, %24;\n\t"
        "subc.u32 %8, 0, 0;"
        : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]),
          "=r"(r.v[6]), "=r"(r.v[7]), "=r"(brw)
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]),
          "r"(a.v[7]), "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]),
          "r"(b.v[6]), "r"(b.v[7]));
    // brw is 0 or 0xffffffff: add back p & brw
    asm("add.cc.u32 %0, %0, %8;\n\t"
        "addc.cc.u32 %1, %1, %9;\n\t"
        "addc.cc.u32 %2, %2, %10;\n\t"
        "addc.cc.u32 %3, %3, %11;\n\t"
        "addc.cc.u32 %4, %4, %12;\n\t"
        "addc.cc.u32 %5, %5, %13;\n\t"
        "addc.cc.u32 %6, %6, %14;\n\t"
        "addc.u32 %7, %7, %15;"
        : "+r"(r.v[0]), "+r"(r.v[1]), "+r"(r.v[2]), "+r"(r.v[3]), "+r"(r.v[4]), "+r"(r.v[5]),
          "+r"(r.v[6]), "+r"(r.v[7])
        : "r"(C::M0 & brw), "r"(C::M1 & brw), "r"(C::M2 & brw), "r"(C::M3 & brw), "r"(C::M4 & brw),
          "r"(C::M5 & brw), "r"(C::M6 & brw), "r"(C::M7 & brw));
    return r;
  }

  __device__ __forceinline__ Fp neg() const { return is_zero() ? *this : modulus_minus(*this); }
  static __device__ __forceinline__ Fp modulus_minus(const Fp& a) {  // p - a, a in (0, p]
    Fp r;
    asm("sub.cc.u32 %0, %16, %8;\n\t"
        "subc.cc.u32 %1, %17, %9;\n\t"
        "subc.cc.u32 %2, %18, %10;\n\t"
        "subc.cc.u32 %3, %19, %11;\n\t"
        "subc.cc.u32 %4, %20, %12;\n\t"
        "subc.cc.u32 %5, %21, %13;\n\t"
        "subc.cc.u32 %6, %22, %14;\n\t"
        "subc.u32 %7, %23, %15;"
        : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]),
          "=r"(r.v[6]), "=r"(r.v[7])
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]),
          "r"(a.v[7]), "n"(C::M0), "n"(C::M1), "n"(C::M2), "n"(C::M3), "n"(C::M4), "n"(C::M5),
          "n"(C::M6), "n"(C::M7));
    return r;
  }
  __device__ __forceinline__ Fp dbl() const { return *this + *this; }

  // ---- Montgomery product ------------------------------------------------------------------------
  // One operand-scanning step.  X holds the even-column lanes, Y the odd-column lanes (weight 2^32).
  //   FIRST: X = a_even*b ; Y = a_odd*b
  //   else : (X, Y) <- ((X + Y'[1]) + a_even*b , (Y' >> 64) + a_odd*b) where Y' is the previous step's
  //          even accumulator whose low word is already zero (roles swap every step).
  // then the reduction: m = X[0]*INV ; X += m*p_even ; Y += m*p_odd  (now X[0] == 0).
  template <bool FIRST>
  static __device__ __forceinline__ void mont_prod(uint32_t* X, uint32_t* Y, const uint32_t* a, uint32_t b) {
    if (FIRST) {
      asm("mul.lo.u32 %0, %16, %24;\n\t mul.hi.u32 %1, %16, %24;\n\t"
          "mul.lo.u32 %2, %18, %24;\n\t mul.hi.u32 %3, %18, %24;\n\t"
          "mul.lo.u32 %4, %20, %24;\n\t mul.hi.u32 %5, %20, %24;\n\t"
          "mul.lo.u32 %6, %22, %24;\n\t mul.hi.u32 %7, %22, %24;\n\t"
          "mul.lo.u32 %8, %17, %24;\n\t mul.hi.u32 %9, %17, %24;\n\t"
          "mul.lo.u32 %10, %19, %24;\n\t mul.hi.u32 %11, %19, %24;\n\t"
          "mul.lo.u32 %12, %21, %24;\n\t mul.hi.u32 %13, %21, %24;\n\t"
          "mul.lo.u32 %14, %23, %24;\n\t mul.hi.u32 %15, %23, %24;"
          : "=r"(X[0]), "=r"(X[1]), "=r"(X[2]), "=r"(X[3]), "=r"(X[4]), "=r"(X[5]), "=r"(X[6]),
            "=r"(X[7]), "=r"(Y[0]), "=r"(Y[1]), "=r"(Y[2]), "=r"(Y[3]), "=r"(Y[4]), "=r"(Y[5]),
            "=r"(Y[6]), "=r"(Y[7])
          : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
            "r"(b));
    } else {
      asm("add.cc.u32 %0, %0, %9;\n\t"
          // odd lanes, reading the old even accumulator two words up (the >> 64)
          "madc.lo.cc.u32 %8, %17, %24, %10;\n\t madc.hi.cc.u32 %9, %17, %24, %11;\n\t"
          "madc.lo.cc.u32 %10, %19, %24, %12;\n\t madc.hi.cc.u32 %11, %19, %24, %13;\n\t"
          "madc.lo.cc.u32 %12, %21, %24, %14;\n\t madc.hi.cc.u32 %13, %21, %24, %15;\n\t"
          "madc.lo.cc.u32 %14, %23, %24, 0;\n\t madc.hi.u32 %15, %23, %24, 0;\n\t"
          // even lanes
          "mad.lo.cc.u32 %0, %16, %24, %0;\n\t madc.hi.cc.u32 %1, %16, %24, %1;\n\t"
          "madc.lo.cc.u32 %2, %18, %24, %2;\n\t madc.hi.cc.u32 %3, %18, %24, %3;\n\t"
          "madc.lo.cc.u32 %4, %20, %24, %4;\n\t madc.hi.cc.u32 %5, %20, %24, %5;\n\t"
          "madc.lo.cc.u32 %6, %22, %24, %6;\n\t madc.hi.cc.u32 %7, %22, %24, %7;\n\t"
          "addc.u32 %15, %15, 0;"
          : "+r"(X[0]), "+r"(X[1]), "+r"(X[2]), "+r"(X[3]), "+r"(X[4]), "+r"(X[5]), "+r"(X[6]),
            "+r"(X[7]), "+r"(Y[0]), "+r"(Y[1]), "+r"(Y[2]), "+r"(Y[3]), "+r"(Y[4]), "+r"(Y[5]),
            "+r"(Y[6]), "+r"(Y[7])
          : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
            "r"(b));
    }
  }
  // the reduction row: m = X[0]*INV ; X += m*p_even ; Y += m*p_odd  (now X[0] == 0)
  static __device__ __forceinline__ void mont_reduce_row(uint32_t* X, uint32_t* Y) {
    uint32_t m = X[0] * C::INV;
    asm("mad.lo.cc.u32 %8, %16, %18, %8;\n\t madc.hi.cc.u32 %9, %16, %18, %9;\n\t"
        "madc.lo.cc.u32 %10, %16, %20, %10;\n\t madc.hi.cc.u32 %11, %16, %20, %11;\n\t"
        "madc.lo.cc.u32 %12, %16, %22, %12;\n\t madc.hi.cc.u32 %13, %16, %22, %13;\n\t"
        "madc.lo.cc.u32 %14, %16, %24, %14;\n\t madc.hi.u32 %15, %16, %24, %15;\n\t"
        "mad.lo.cc.u32 %0, %16, %17, %0;\n\t madc.hi.cc.u32 %1, %16, %17, %1;\n\t"
        "madc.lo.cc.u32 %2, %16, %19, %2;\n\t madc.hi.cc.u32 %3, %16, %19, %3;\n\t"
        "madc.lo.cc.u32 %4, %16, %21, %4;\n\t madc.hi.cc.u32 %5, %16, %21, %5;\n\t"
        "madc.lo.cc.u32 %6, %16, %23, %6;\n\t madc.hi.cc.u32 %7, %16, %23, %7;\n\t"
        "addc.u32 %15, %15, 0;"
        : "+r"(X[0]), "+r"(X[1]), "+r"(X[2]), "+r"(X[3]), "+r"(X[4]), "+r"(X[5]), "+r"(X[6]),
          "+r"(X[7]), "+r"(Y[0]), "+r"(Y[1]), "+r"(Y[2]), "+r"(Y[3]), "+r"(Y[4]), "+r"(Y[5]),
          "+r"(Y[6]), "+r"(Y[7])
        : "r"(m), "n"(C::M0), "n"(C::M1), "n"(C::M2), "n"(C::M3), "n"(C::M4), "n"(C::M5), "n"(C::M6),
          "n"(C::M7));
  }
  template <bool FIRST>
  static __device__ __forceinline__ void mont_step(uint32_t* X, uint32_t* Y, const uint32_t* a, uint32_t b) {
    mont_prod<FIRST>(X, Y, a, b);
    mont_reduce_row(X, Y);
  }

  // after eight steps the roles are X = O (low word zero), Y = E: value = (X >> 32) + Y, below 2^256 by the magnitude bound
  static __device__ __forceinline__ Fp mont_merge(const uint32_t* E, const uint32_t* O) {
    Fp r;
    asm("add.cc.u32 %0, %8, %16;\n\t"
        "addc.cc.u32 %1, %9, %17;\n\t"
        "addc.cc.u32 %2, %10, %18;\n\t"
        "addc.cc.u32 %3, %11, %19;\n\t"
        "addc.cc.u32 %4, %12, %20;\n\t"
        "addc.cc.u32 %5, %13, %21;\n\t"
        "addc.cc.u32 %6, %14, %22;\n\t"
        "addc.u32 %7, %15, 0;"
        : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]),
          "=r"(r.v[6]), "=r"(r.v[7])
        : "r"(E[0]), "r"(E[1]), "r"(E[2]), "r"(E[3]), "r"(E[4]), "r"(E[5]), "r"(E[6]), "r"(E[7]),
          "r"(O[1]), "r"(O[2]), "r"(O[3]), "r"(O[4]), "r"(O[5]), "r"(O[6]), "r"(O[7]));
    return r;
  }

  friend __device__ __forceinline__ Fp operator*(const Fp& a, const Fp& b) {
    uint32_t E[8], O[8];
    mont_step<true>(E, O, a.v, b.v[0]);
    mont_step<false>(O, E, a.v, b.v[1]);
    mont_step<false>(E, O, a.v, b.v[2]);
    mont_step<false>(O, E, a.v, b.v[3]);
    mont_step<false>(E, O, a.v, b.v[4]);
    mont_step<false>(O, E, a.v, b.v[5]);
    mont_step<false>(E, O, a.v, b.v[6]);
    mont_step<false>(O, E, a.v, b.v[7]);
    Fp r = mont_merge(E, O);
    reduce_once(r);   // < 2p
    return r;
  }

  // ---- a*b + c*d in ONE Montgomery pass (lazy reduction) -------------------------------------------
  // (X, Y) += c * d_i on top of a step's product row, before its reduction row: even words of c into the even lanes (carry out
  // of word 7 into Y[7] = word 8), odd words into the odd lanes.  Magnitudes: with a, c <= p the running value stays below
  // 3p (1 + 1/2^32) and a step's sum below 3 * 2^32 * p < 2^288, the nine words the two accumulators span -- no carry is lost.
  static __device__ __forceinline__ void mont_prod_add(uint32_t* X, uint32_t* Y, const uint32_t* a, uint32_t b) {
    asm("mad.lo.cc.u32 %0, %16, %24, %0;\n\t madc.hi.cc.u32 %1, %16, %24, %1;\n\t"
        "madc.lo.cc.u32 %2, %18, %24, %2;\n\t madc.hi.cc.u32 %3, %18, %24, %3;\n\t"
        "madc.lo.cc.u32 %4, %20, %24, %4;\n\t madc.hi.cc.u32 %5, %20, %24, %5;\n\t"
        "madc.lo.cc.u32 %6, %22, %24, %6;\n\t madc.hi.cc.u32 %7, %22, %24, %7;\n\t"
        "addc.u32 %15, %15, 0;\n\t"
        "mad.lo.cc.u32 %8, %17, %24, %8;\n\t madc.hi.cc.u32 %9, %17, %24, %9;\n\t"
        "madc.lo.cc.u32 %10, %19, %24, %10;\n\t madc.hi.cc.u32 %11, %19, %24, %11;\n\t"
        "madc.lo.cc.u32 %12, %21, %24, %12;\n\t madc.hi.cc.u32 %13, %21, %24, %13;\n\t"
        "madc.lo.cc.u32 %14, %23, %24, %14;\n\t madc.hi.u32 %15, %23, %24, %15;"
        : "+r"(X[0]), "+r"(X[1]), "+r"(X[2]), "+r"(X[3]), "+r"(X[4]), "+r"(X[5]), "+r"(X[6]),
          "+r"(X[7]), "+r"(Y[0]), "+r"(Y[1]), "+r"(Y[2]), "+r"(Y[3]), "+r"(Y[4]), "+r"(Y[5]),
          "+r"(Y[6]), "+r"(Y[7])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
          "r"(b));
  }
  template <bool FIRST>
  static __device__ __forceinline__ void mont_step2(uint32_t* X, uint32_t* Y, const uint32_t* a, uint32_t b, const uint32_t* c,
                                                    uint32_t d) {
    mont_prod<FIRST>(X, Y, a, b);
    mont_prod_add(X, Y, c, d);
    mont_reduce_row(X, Y);
  }
  // (a*b + c*d) / R mod p for a, b, c, d <= p: two products, ONE reduction (200 multiply-accumulates instead of 272).
  // The XYZZ formulas end in  y3 = R (Q - x3) - y1 PPP = R (Q - x3) + (p - y1) PPP: that is this.
  static __device__ __forceinline__ Fp mul_add_mul(const Fp& a, const Fp& b, const Fp& c, const Fp& d) {
    uint32_t E[8], O[8];
    mont_step2<true>(E, O, a.v, b.v[0], c.v, d.v[0]);
    mont_step2<false>(O, E, a.v, b.v[1], c.v, d.v[1]);
    mont_step2<false>(E, O, a.v, b.v[2], c.v, d.v[2]);
    mont_step2<false>(O, E, a.v, b.v[3], c.v, d.v[3]);
    mont_step2<false>(E, O, a.v, b.v[4], c.v, d.v[4]);
    mont_step2<false>(O, E, a.v, b.v[5], c.v, d.v[5]);
    mont_step2<false>(E, O, a.v, b.v[6], c.v, d.v[6]);
    mont_step2<false>(O, E, a.v, b.v[7], c.v, d.v[7]);
    Fp r = mont_merge(E, O);
    reduce_once(r);   // < 3p (1 + 2^-32): two conditional subtractions
    reduce_once(r);
    return r;
  }

  __device__ __forceinline__ Fp sqr() const { return *this * *this; }

  // canonical (non-Montgomery) little-endian words <-> Montgomery form
  __device__ __forceinline__ Fp to_mont() const { return *this * r2(); }
  __device__ __forceinline__ Fp from_mont() const {
    Fp o = zero();
    o.v[0] = 1;
    return *this * o;
  }

  // a^e for a 256-bit exponent given as 8 LE words (square-and-multiply, MSB first)
  __device__ Fp pow_words(const uint32_t* e) const {
    Fp r = one();
    bool started = false;
    for (int i = 7; i >= 0; i--) {
      for (int bit = 31; bit >= 0; bit--) {
        if (started) r = r.sqr();
        if ((e[i] >> bit) & 1u) {
          r = started ? r * *this : *this;
          started = true;
        }
      }
    }
    return r;
  }

  // a^-1 by the binary extended Euclidean algorithm (variable time, ~10x fewer instructions than Fermat).  For the
  // single-thread tails (final affine conversion of an MSM, proof assembly) where latency, not divergence, matters.
  __device__ Fp inverse_vartime() const {
    if (is_zero()) return zero();
    // integers: u = a R (this representation), v = p ; x1 * (aR) = u, x2 * (aR) = v (mod p)
    Fp u = *this, v = modulus(), x1 = zero(), x2 = zero();
    x1.v[0] = 1;
    auto is_one = [](const Fp& a) { return a.v[0] == 1 && (a.v[1] | a.v[2] | a.v[3] | a.v[4] | a.v[5] | a.v[6] | a.v[7]) == 0; };
    auto shr1 = [](Fp& a) {
#pragma unroll
      for (int i = 0; i < 7; i++) a.v[i] = (a.v[i] >> 1) | (a.v[i + 1] << 31);
      a.v[7] >>= 1;
    };
    auto half_mod = [&](Fp& x) {  // x / 2 mod p, x < p
      if (x.v[0] & 1u) {
        unsigned long long c = 0;
        const Fp m = modulus();
#pragma unroll
        for (int i = 0; i < 8; i++) {
          c += (unsigned long long)x.v[i] + m.v[i];
          x.v[i] = (uint32_t)c;
          c >>= 32;
        }  // x + p < 2^255: no carry out
      }
      shr1(x);
    };
    auto geq = [](const Fp& a, const Fp& b) {
      for (int i = 7; i >= 0; i--) {
        if (a.v[i] > b.v[i]) return true;
        if (a.v[i] < b.v[i]) return false;
      }
      return true;
    };
    auto sub_raw = [](Fp& a, const Fp& b) {  // a -= b, a >= b
      long long br = 0;
#pragma unroll
      for (int i = 0; i < 8; i++) {
        long long d = (long long)a.v[i] - b.v[i] + br;
        a.v[i] = (uint32_t)d;
        br = d >> 32;
      }
    };
    while (!is_one(u) && !is_one(v)) {
      while (!(u.v[0] & 1u)) {
        shr1(u);
        half_mod(x1);
      }
      while (!(v.v[0] & 1u)) {
        shr1(v);
        half_mod(x2);
      }
      if (geq(u, v)) {
        sub_raw(u, v);
        x1 = x1 - x2;
      } else {
        sub_raw(v, u);
        x2 = x2 - x1;
      }
    }
    Fp t = is_one(u) ? x1 : x2;  // (a R)^-1 as an integer
    Fp r3 = r2() * r2();         // R^3 in Montgomery arithmetic: R^2 R^2 R^-1
    return t * r3;               // a^-1 R^-1 R^3 R^-1 = a^-1 R
  }

  // Fermat inverse a^(p-2); 0 -> 0
  __device__ Fp inverse() const {
    uint32_t e[8] = {C::M0 - 2u, C::M1, C::M2, C::M3, C::M4, C::M5, C::M6, C::M7};  // M0 >= 2 for both fields
    return pow_words(e);
  }
};

using Fq = Fp<FqCfg>;
using Fr = Fp<FrCfg>;

}  // namespace zkb
