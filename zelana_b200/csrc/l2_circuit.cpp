// Host side of the reference's L2 batch circuit: constraint synthesis, witness assignment and the off-circuit Poseidon
// roots, plus BatchProver::prove end to end (RNG -> assignment -> zkb_prove -> Solana bytes).
//
// Replaces, in the reference tree:
//   prover/src/l2_circuit.rs:68-83,147-170,179-505   get_poseidon_config, L2BlockCircuit::dummy, generate_constraints
//   prover/src/main.rs.bak:93-154                    calculate_new_root_offchain
//   core/src/sequencer/settlement/prover.rs:304-334,350-425   proof_to_solana_bytes, Groth16Prover::prove
// and the un-vendored crates those call (ark-relations / ark-r1cs-std / ark-crypto-primitives 0.5.0, rand 0.8.5 StdRng).
//
// This is host work in the reference too ("Host-side (not GPU): circuit synthesis (a2), RNG -> (r, s)"); it is
// native here because the reference's is.  Design, different from arkworks' symbolic-LC ConstraintSystem:
//   * ONE synthesis routine runs in two modes.  STRUCTURE (once per circuit shape) keeps every linear combination as a
//     sorted (column, coefficient) vector and emits the three CSR matrices directly -- no symbolic variables, no inlining
//     pass.  ASSIGN (once per proof) runs the same statements with values only: ~16 k field multiplications, no allocation
//     of linear combinations at all.  Constant-versus-variable tracking is shared, so both modes allocate identical
//     variables in identical order.
//   * the GPU never waits for it: zkb_l2_prove assigns on the host and hands z to zkb_prove.
#include <algorithm>
#include <array>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <cstdint>
#include <cstdlib>
#include <initializer_list>
#include <cstring>
#include <map>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/zkb200.h"

namespace l2 {

// ------------------------------------------------------------------------------------------ Fr on the host (4 x 64, Montgomery)
typedef unsigned __int128 u128;
struct Fr {
  uint64_t l[4];
  bool operator==(const Fr& o) const { return l[0] == o.l[0] && l[1] == o.l[1] && l[2] == o.l[2] && l[3] == o.l[3]; }
  bool operator!=(const Fr& o) const { return !(*this == o); }
  bool is_zero() const { return (l[0] | l[1] | l[2] | l[3]) == 0; }
};
static const Fr MOD = {{0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull}};
static const uint64_t NINV = 0xc2e1f593efffffffull;  // -r^-1 mod 2^64

static inline bool geq(const Fr& a, const Fr& b) {
  for (int i = 3; i >= 0; --i) {
    if (a.l[i] != b.l[i]) return a.l[i] > b.l[i];
  }
  return true;
}
static inline Fr raw_sub(const Fr& a, const Fr& b) {
  Fr r;
  u128 br = 0;
  for (int i = 0; i < 4; ++i) {
    u128 t = (u128)a.l[i] - b.l[i] - (uint64_t)br;
    r.l[i] = (uint64_t)t;
    br = (t >> 64) & 1;
  }
  return r;
}
static inline Fr add(const Fr& a, const Fr& b) {
  Fr r;
  u128 c = 0;
  for (int i = 0; i < 4; ++i) {
    c += (u128)a.l[i] + b.l[i];
    r.l[i] = (uint64_t)c;
    c >>= 64;
  }
  return geq(r, MOD) ? raw_sub(r, MOD) : r;  // r < 2^254 + 2^254, no carry out of 256 bits
}
static inline Fr sub(const Fr& a, const Fr& b) {
  if (geq(a, b)) return raw_sub(a, b);
  Fr t = raw_sub(MOD, b);
  Fr r;
  u128 c = 0;
  for (int i = 0; i < 4; ++i) {
    c += (u128)a.l[i] + t.l[i];
    r.l[i] = (uint64_t)c;
    c >>= 64;
  }
  return r;
}
static inline Fr neg(const Fr& a) { return a.is_zero() ? a : raw_sub(MOD, a); }
static inline Fr mont_mul(const Fr& a, const Fr& b) {
  // CIOS without the extra carry word: valid because the modulus leaves its top two bits clear (r < 2^254); with the
  // multiplicand a < r and ANY multiplier b < 2^256 every partial sum stays below 2 r (from_le_bytes_mod_order relies on it).
  uint64_t t0 = 0, t1 = 0, t2 = 0, t3 = 0;
#pragma GCC unroll 4
  for (int i = 0; i < 4; ++i) {
    const uint64_t bi = b.l[i];
    u128 c = (u128)a.l[0] * bi + t0;
    const uint64_t lo = (uint64_t)c;
    const uint64_t m = lo * NINV;
    u128 d = (u128)m * MOD.l[0] + lo;
    c = (u128)a.l[1] * bi + t1 + (uint64_t)(c >> 64);
    d = (u128)m * MOD.l[1] + (uint64_t)c + (uint64_t)(d >> 64);
    t0 = (uint64_t)d;
    c = (u128)a.l[2] * bi + t2 + (uint64_t)(c >> 64);
    d = (u128)m * MOD.l[2] + (uint64_t)c + (uint64_t)(d >> 64);
    t1 = (uint64_t)d;
    c = (u128)a.l[3] * bi + t3 + (uint64_t)(c >> 64);
    d = (u128)m * MOD.l[3] + (uint64_t)c + (uint64_t)(d >> 64);
    t2 = (uint64_t)d;
    t3 = (uint64_t)(c >> 64) + (uint64_t)(d >> 64);
  }
  Fr r = {{t0, t1, t2, t3}};
  return geq(r, MOD) ? raw_sub(r, MOD) : r;
}

struct Consts {
  Fr one, r2, zero;
  Consts() {
    zero = Fr{{0, 0, 0, 0}};
    Fr x = {{1, 0, 0, 0}};
    for (int i = 0; i < 256; ++i) x = add(x, x);  // 2^256 mod r
    one = x;
    for (int i = 0; i < 256; ++i) x = add(x, x);  // 2^512 mod r
    r2 = x;
  }
};
static const Consts K;
static inline Fr to_mont(const Fr& canonical) { return mont_mul(canonical, K.r2); }
static inline Fr from_mont(const Fr& a) { return mont_mul(a, Fr{{1, 0, 0, 0}}); }
static inline Fr from_u64(uint64_t v) { return to_mont(Fr{{v, 0, 0, 0}}); }
static Fr pow_limbs(const Fr& a, const Fr& e) {
  Fr r = K.one;
  for (int i = 255; i >= 0; --i) {
    r = mont_mul(r, r);
    if ((e.l[i >> 6] >> (i & 63)) & 1) r = mont_mul(r, a);
  }
  return r;
}
static Fr inverse(const Fr& a) {  // Fermat; a != 0
  Fr e = raw_sub(MOD, Fr{{2, 0, 0, 0}});
  return pow_limbs(a, e);
}
// Inverses of 1..SMALL_INV (Montgomery), by one Fermat inversion and Montgomery's trick.  The is_eq inside Boolean::kary_and
// inverts (count - sum of bits), always a small integer; a Fermat inversion each time was a third of an assignment pass.
constexpr uint64_t SMALL_INV = 512;
struct SmallInverses {
  Fr inv[SMALL_INV + 1];
  SmallInverses() {
    Fr prefix[SMALL_INV + 1];
    Fr k = K.one, acc = K.one;
    for (uint64_t i = 1; i <= SMALL_INV; ++i) {
      prefix[i] = acc;  // (i - 1)!
      acc = mont_mul(acc, k);
      k = add(k, K.one);
    }
    Fr inv_acc = inverse(acc);  // 1 / SMALL_INV!
    for (uint64_t i = SMALL_INV; i >= 1; --i) {
      inv[i] = mont_mul(inv_acc, prefix[i]);
      inv_acc = mont_mul(inv_acc, from_u64(i));
    }
    inv[0] = K.zero;
  }
};
static Fr inverse_fast(const Fr& a) {  // a != 0
  static const SmallInverses table;
  Fr c = from_mont(a);
  if ((c.l[1] | c.l[2] | c.l[3]) == 0 && c.l[0] <= SMALL_INV) return table.inv[c.l[0]];
  Fr n = from_mont(neg(a));
  if ((n.l[1] | n.l[2] | n.l[3]) == 0 && n.l[0] <= SMALL_INV) return neg(table.inv[n.l[0]]);
  return inverse(a);
}
// Fr::from_le_bytes_mod_order for up to 32 bytes (every use in the circuit): value < 2^256 = reduce by Montgomery product
static Fr from_le_bytes_mod_order(const uint8_t* b, size_t n) {
  Fr x = {{0, 0, 0, 0}};
  for (size_t i = 0; i < n && i < 32; ++i) x.l[i >> 3] |= (uint64_t)b[i] << (8 * (i & 7));
  return mont_mul(K.r2, x);  // (R^2 * x) / R = x R mod r; the multiplier operand may be any x < 2^256
}
static void to_le_bytes(const Fr& mont, uint8_t out[32]) {
  if (mont.is_zero() || mont == K.one) {  // most of an assignment is bits
    memset(out, 0, 32);
    out[0] = mont.is_zero() ? 0 : 1;
    return;
  }
  Fr c = from_mont(mont);
  for (int i = 0; i < 32; ++i) out[i] = (uint8_t)(c.l[i >> 3] >> (8 * (i & 7)));
}

// ------------------------------------------------------------------------------------------ Poseidon parameters (Grain LFSR)
// ark-crypto-primitives sponge/poseidon/grain_lfsr.rs + find_poseidon_ark_and_mds, for get_poseidon_config():
// 254-bit prime, state 3 (rate 2 + capacity 1), 8 full + 56 partial rounds, x^5.
constexpr int T = 3, RATE = 2, CAPACITY = 1, FULL = 8, PARTIAL = 56, ROUNDS = FULL + PARTIAL, PRIME_BITS = 254;
constexpr uint64_t ALPHA = 5;

struct Grain {
  bool st[80];
  int head = 0;
  Grain(uint64_t prime_bits, uint64_t t, uint64_t rf, uint64_t rp) {
    for (bool& b : st) b = false;
    st[1] = true;  // prime field; S-box x^alpha leaves bits 2..5 clear
    auto put = [&](int lo, int hi, uint64_t v) {
      for (int i = hi; i >= lo; --i) {
        st[i] = v & 1;
        v >>= 1;
      }
    };
    put(6, 17, prime_bits);
    put(18, 29, t);
    put(30, 39, rf);
    put(40, 49, rp);
    for (int i = 50; i < 80; ++i) st[i] = true;
    for (int i = 0; i < 160; ++i) update();
  }
  bool update() {
    auto at = [&](int k) { return st[(head + k) % 80]; };
    bool nb = at(62) ^ at(51) ^ at(38) ^ at(23) ^ at(13) ^ at(0);
    st[head] = nb;
    head = (head + 1) % 80;
    return nb;
  }
  bool next_bit() {  // pairs: the second bit counts only if the first is set
    bool nb = update();
    while (!nb) {
      update();
      nb = update();
    }
    return update();
  }
  Fr raw_element() {  // PRIME_BITS bits, most significant first, as a plain integer
    Fr x = {{0, 0, 0, 0}};
    for (int i = PRIME_BITS - 1; i >= 0; --i) {
      if (next_bit()) x.l[i >> 6] |= 1ull << (i & 63);
    }
    return x;
  }
  Fr rejection_sampled() {
    for (;;) {
      Fr x = raw_element();
      if (!geq(x, MOD)) return to_mont(x);
    }
  }
  Fr mod_p() {
    Fr x = raw_element();  // < 2^254 < 2 r
    if (geq(x, MOD)) x = raw_sub(x, MOD);
    return to_mont(x);
  }
};

struct PoseidonConfig {
  Fr ark[ROUNDS][T];
  Fr mds[T][T];
  PoseidonConfig() {
    Grain g(PRIME_BITS, T, FULL, PARTIAL);
    for (int r = 0; r < ROUNDS; ++r) {
      for (int i = 0; i < T; ++i) ark[r][i] = g.rejection_sampled();
    }
    Fr xs[T], ys[T];
    for (int i = 0; i < T; ++i) xs[i] = g.mod_p();
    for (int i = 0; i < T; ++i) ys[i] = g.mod_p();
    for (int i = 0; i < T; ++i) {
      for (int j = 0; j < T; ++j) mds[i][j] = inverse(add(xs[i], ys[j]));
    }
  }
};
static const PoseidonConfig& config() {
  static const PoseidonConfig c;
  return c;
}

// ------------------------------------------------------------------------------------------ constraint builder
struct Term {
  uint32_t col;
  Fr co;
};
typedef std::vector<Term> LC;  // sorted by column, no zero coefficients

struct Csr {
  std::vector<uint64_t> row_ptr;
  std::vector<uint32_t> col;
  std::vector<Fr> co;          // Montgomery while building
  std::vector<uint8_t> bytes;  // canonical LE, filled by seal()
  Csr() : row_ptr(1, 0) {}
  void push(const LC& lc) {
    for (const Term& t : lc) {
      col.push_back(t.col);
      co.push_back(t.co);
    }
    row_ptr.push_back(col.size());
  }
  void seal() {
    bytes.resize(co.size() * 32);
    for (size_t i = 0; i < co.size(); ++i) to_le_bytes(co[i], &bytes[32 * i]);
  }
};

constexpr uint32_t NUM_INSTANCE = 8;  // ONE + the seven public inputs, all allocated before the first witness

// One record per Poseidon hash / balance comparison, in call order: how many witnesses and constraints it allocates.  Taken
// in the STRUCTURE pass; lets a parallel ASSIGN walker step over a piece another thread computes.
struct CallRec {
  uint32_t nwit, ncons;
};
enum CallCat { CAT_LEAF = -1 };  // >= 0: the fold chain a hash belongs to (post state, shielded, withdrawals, batch, pre state)
constexpr int NUM_CHAINS = 5;

struct Builder {
  bool structure;  // keep linear combinations and emit matrices
  // full assignment [1, instance.., witness..], Montgomery: an owned vector, or (parallel ASSIGN) a shared array every walker
  // indexes identically and writes only the pieces it owns
  std::vector<Fr> zown;
  Fr* zpar = nullptr;
  size_t pos = 0;
  bool write = true;
  uint32_t n_inputs = 1;
  Csr A, B, C;
  uint64_t n_constraints = 0;
  std::vector<CallRec> recs_out;            // STRUCTURE: recorded
  // parallel ASSIGN walker
  const std::vector<CallRec>* recs = nullptr;
  Fr* hash_out = nullptr;                   // output value of every hash call, shared between the two passes
  size_t call_idx = 0, task_counter = 0;
  int pass = 0, part = 0, nparts = 1, chain = -2;

  explicit Builder(bool s) : structure(s) { emit(K.one); }
  Builder(Fr* shared, const std::vector<CallRec>* r, Fr* hout, int pass_, int part_, int nparts_, int chain_)
      : structure(false), zpar(shared), recs(r), hash_out(hout), pass(pass_), part(part_), nparts(nparts_), chain(chain_) {
    write = pass == 1 && part == 0;  // the loose witnesses (inputs, balances, keys, counts) are written once
    emit(K.one);
  }
  bool parallel() const { return zpar != nullptr; }
  void emit(const Fr& v) {
    if (zpar) {
      if (write) zpar[pos] = v;
    } else {
      zown.push_back(v);
    }
    ++pos;
  }
  uint32_t new_input(const Fr& v) {
    emit(v);
    return n_inputs++;
  }
  uint32_t new_witness(const Fr& v) {
    emit(v);
    return (uint32_t)pos - 1;
  }
  void enforce(const LC& a, const LC& b, const LC& c) {
    ++n_constraints;
    if (structure) {
      A.push(a);
      B.push(b);
      C.push(c);
    }
  }
  void skip() { ++n_constraints; }  // ASSIGN mode: the constraint exists, its rows are not needed
};

static LC lc_var(uint32_t col) { return LC{Term{col, K.one}}; }
static LC lc_one() { return lc_var(0); }
static LC lc_axpy(const LC& x, const LC& y, const Fr& k) {  // x + k y
  LC out;
  out.reserve(x.size() + y.size());
  size_t i = 0, j = 0;
  while (i < x.size() || j < y.size()) {
    if (j == y.size() || (i < x.size() && x[i].col < y[j].col)) {
      out.push_back(x[i++]);
    } else if (i == x.size() || y[j].col < x[i].col) {
      Fr c = mont_mul(y[j].co, k);
      if (!c.is_zero()) out.push_back(Term{y[j].col, c});
      ++j;
    } else {
      Fr c = add(x[i].co, mont_mul(y[j].co, k));
      if (!c.is_zero()) out.push_back(Term{x[i].col, c});
      ++i;
      ++j;
    }
  }
  return out;
}
static LC lc_scale(const LC& x, const Fr& k) {
  LC out;
  if (k.is_zero()) return out;
  out.reserve(x.size());
  for (const Term& t : x) out.push_back(Term{t.col, mont_mul(t.co, k)});
  return out;
}

// FpVar: Constant(v) or Var(value v, linear combination lc).  Boolean likewise.
struct FpVar {
  bool c;
  Fr v;
  LC lc;
};
struct BoolVar {
  bool c;
  bool v;
  LC lc;
};

struct Gadgets {
  Builder& b;
  const Fr MINUS_ONE = neg(K.one);
  const Fr TWO = add(K.one, K.one);
  explicit Gadgets(Builder& bb) : b(bb) {}
  bool S() const { return b.structure; }

  FpVar constant(const Fr& v) { return FpVar{true, v, LC()}; }
  FpVar input(const Fr& v) {
    uint32_t col = b.new_input(v);
    return FpVar{false, v, S() ? lc_var(col) : LC()};
  }
  FpVar witness(const Fr& v) {
    uint32_t col = b.new_witness(v);
    return FpVar{false, v, S() ? lc_var(col) : LC()};
  }
  LC as_lc(const FpVar& x) { return x.c ? lc_scale(lc_one(), x.v) : x.lc; }

  FpVar add_const(const FpVar& x, const Fr& k) {
    if (x.c) return constant(add(x.v, k));
    if (k.is_zero()) return x;
    return FpVar{false, add(x.v, k), S() ? lc_axpy(x.lc, lc_one(), k) : LC()};
  }
  FpVar fadd(const FpVar& x, const FpVar& y) {
    if (y.c) return add_const(x, y.v);
    if (x.c) return add_const(y, x.v);
    return FpVar{false, add(x.v, y.v), S() ? lc_axpy(x.lc, y.lc, K.one) : LC()};
  }
  FpVar fneg(const FpVar& x) {
    if (x.c) return constant(neg(x.v));
    return FpVar{false, neg(x.v), S() ? lc_scale(x.lc, MINUS_ONE) : LC()};
  }
  FpVar fsub(const FpVar& x, const FpVar& y) {
    if (y.c) return add_const(x, neg(y.v));
    if (x.c) return fneg(add_const(y, neg(x.v)));
    return FpVar{false, sub(x.v, y.v), S() ? lc_axpy(x.lc, y.lc, MINUS_ONE) : LC()};
  }
  FpVar fdouble(const FpVar& x) {
    if (x.c) return constant(add(x.v, x.v));
    return FpVar{false, add(x.v, x.v), S() ? lc_scale(x.lc, TWO) : LC()};
  }
  FpVar mul_const(const FpVar& x, const Fr& k) {
    if (x.c) return constant(mont_mul(x.v, k));
    return FpVar{false, mont_mul(x.v, k), S() ? lc_scale(x.lc, k) : LC()};
  }
  FpVar fmul(const FpVar& x, const FpVar& y) {
    if (y.c) return mul_const(x, y.v);
    if (x.c) return mul_const(y, x.v);
    FpVar p = witness(mont_mul(x.v, y.v));  // AllocatedFp::mul: new witness, x * y = p
    b.enforce(x.lc, y.lc, p.lc);
    return p;
  }
  FpVar pow_by_constant(const FpVar& x, uint64_t e) {
    FpVar res = constant(K.one);
    int top = 63;
    while (top > 0 && !((e >> top) & 1)) --top;
    for (int i = top; i >= 0; --i) {
      res = fmul(res, res);
      if ((e >> i) & 1) res = fmul(res, x);
    }
    return res;
  }
  void enforce_equal(const FpVar& x, const FpVar& y) {  // (x - y) * ONE = 0; a constant side goes first
    if (x.c && y.c) return;
    if (x.c || y.c) {
      const FpVar& k = x.c ? x : y;
      const FpVar& v = x.c ? y : x;
      if (S()) b.enforce(lc_axpy(as_lc(k), v.lc, MINUS_ONE), lc_one(), LC());
      else b.skip();
      return;
    }
    if (S()) b.enforce(lc_axpy(x.lc, y.lc, MINUS_ONE), lc_one(), LC());
    else b.skip();
  }

  // ---- Boolean
  BoolVar bconst(bool v) { return BoolVar{true, v, LC()}; }
  BoolVar bwitness(bool v) {  // AllocatedBool::new_witness: (1 - a) * a = 0
    uint32_t col = b.new_witness(v ? K.one : K.zero);
    BoolVar r{false, v, S() ? lc_var(col) : LC()};
    if (S()) b.enforce(lc_axpy(lc_one(), r.lc, MINUS_ONE), r.lc, LC());
    else b.skip();
    return r;
  }
  BoolVar bwitness_unchecked(bool v) {
    uint32_t col = b.new_witness(v ? K.one : K.zero);
    return BoolVar{false, v, S() ? lc_var(col) : LC()};
  }
  LC as_lc(const BoolVar& x) { return x.c ? (x.v ? lc_one() : LC()) : x.lc; }
  BoolVar bnot(const BoolVar& x) {
    if (x.c) return bconst(!x.v);
    return BoolVar{false, !x.v, S() ? lc_axpy(lc_one(), x.lc, MINUS_ONE) : LC()};
  }
  BoolVar band(const BoolVar& x, const BoolVar& y) {
    if (x.c) return x.v ? y : bconst(false);
    if (y.c) return y.v ? x : bconst(false);
    BoolVar r = bwitness_unchecked(x.v && y.v);  // a * b = c
    b.enforce(x.lc, y.lc, r.lc);
    return r;
  }
  BoolVar bor(const BoolVar& x, const BoolVar& y) {
    if (x.c) return x.v ? bconst(true) : y;
    if (y.c) return y.v ? bconst(true) : x;
    BoolVar r = bwitness_unchecked(x.v || y.v);  // (1 - a) * (1 - b) = (1 - c)
    if (S()) {
      b.enforce(lc_axpy(lc_one(), x.lc, MINUS_ONE), lc_axpy(lc_one(), y.lc, MINUS_ONE), lc_axpy(lc_one(), r.lc, MINUS_ONE));
    } else {
      b.skip();
    }
    return r;
  }
  // Boolean::enforce_equal(x, Constant(false)): x * ONE = 0.  Returns false when two constants disagree.
  bool enforce_false(const BoolVar& x) {
    if (x.c) return !x.v;
    b.enforce(x.lc, S() ? lc_one() : LC(), LC());
    return true;
  }
  // AllocatedFp::is_neq(self, other): witnesses is_not_equal (unchecked) and multiplier, two constraints
  BoolVar is_neq(const Fr& sv, const LC& slc, const Fr& ov, const LC& olc) {
    bool ne = sv != ov;
    BoolVar r = bwitness_unchecked(ne);
    uint32_t mcol = b.new_witness(ne ? inverse_fast(sub(sv, ov)) : K.one);
    if (S()) {
      LC diff = lc_axpy(slc, olc, MINUS_ONE);
      b.enforce(diff, lc_var(mcol), r.lc);
      b.enforce(diff, lc_axpy(lc_one(), r.lc, MINUS_ONE), LC());
    } else {
      b.skip();
      b.skip();
    }
    return r;
  }
  BoolVar is_eq(const FpVar& x, const FpVar& y) {
    if (x.c && y.c) return bconst(x.v == y.v);
    if (x.c || y.c) {
      const FpVar& k = x.c ? x : y;
      const FpVar& v = x.c ? y : x;
      return bnot(is_neq(k.v, S() ? as_lc(k) : LC(), v.v, v.lc));
    }
    return bnot(is_neq(x.v, x.lc, y.v, y.lc));
  }
  // Boolean::kary_and: at most three operands as a chain of ANDs, otherwise sum == count
  BoolVar kary_and(const std::vector<BoolVar>& bits) {
    if (bits.size() <= 3) {
      BoolVar cur = bits[0];
      for (size_t i = 1; i < bits.size(); ++i) cur = band(cur, bits[i]);
      return cur;
    }
    Fr consts = K.zero, val = K.zero;
    LC lc;
    for (const BoolVar& x : bits) {
      if (x.c) {
        if (x.v) consts = add(consts, K.one);
      } else {
        if (x.v) val = add(val, K.one);
        if (S()) lc = lc_axpy(lc, x.lc, K.one);
      }
    }
    FpVar sum = add_const(FpVar{false, val, lc}, consts);
    return is_eq(sum, constant(from_u64(bits.size())));
  }
  bool enforce_kary_nand(const std::vector<BoolVar>& bits) { return enforce_false(kary_and(bits)); }

  // Boolean::enforce_smaller_or_equal_than_le(bits (little-endian), element); returns false on a constant contradiction
  bool enforce_smaller_or_equal_than_le(const std::vector<BoolVar>& bits, const Fr& element /* plain integer */) {
    int ebits = 256;
    while (ebits > 0 && !((element.l[(ebits - 1) >> 6] >> ((ebits - 1) & 63)) & 1)) --ebits;
    size_t pos = bits.size();  // walks down: big-endian iteration
    BoolVar last_run = bconst(true);
    std::vector<BoolVar> current_run;
    bool ok = true;
    if ((int)bits.size() > ebits) {
      BoolVar or_result = bconst(false);
      for (size_t i = ebits; i < bits.size(); ++i) {
        or_result = bor(or_result, bits[i]);
        --pos;
      }
      ok &= enforce_false(or_result);
    }
    for (int e = ebits - 1; e >= 0 && pos > 0; --e) {
      const BoolVar& a = bits[--pos];
      if ((element.l[e >> 6] >> (e & 63)) & 1) {
        current_run.push_back(a);
      } else {
        if (!current_run.empty()) {
          current_run.push_back(last_run);
          last_run = kary_and(current_run);
          current_run.clear();
        }
        ok &= enforce_kary_nand({last_run, a});
      }
    }
    return ok;
  }
  std::vector<BoolVar> to_non_unique_bits_le(const FpVar& x) {
    Fr canon = from_mont(x.v);
    std::vector<BoolVar> bits;
    bits.reserve(PRIME_BITS);
    if (x.c) {
      for (int i = 0; i < PRIME_BITS; ++i) bits.push_back(bconst((canon.l[i >> 6] >> (i & 63)) & 1));
      return bits;
    }
    for (int i = 0; i < PRIME_BITS; ++i) bits.push_back(bwitness((canon.l[i >> 6] >> (i & 63)) & 1));
    if (S()) {
      LC lc;
      lc.reserve(PRIME_BITS + x.lc.size());
      Fr coeff = K.one;
      for (int i = 0; i < PRIME_BITS; ++i) {  // fresh witnesses: ascending columns, all above x's
        lc.push_back(Term{bits[i].lc[0].col, coeff});
        coeff = add(coeff, coeff);
      }
      b.enforce(LC(), LC(), lc_axpy(lc, x.lc, MINUS_ONE));
    } else {
      b.skip();
    }
    return bits;
  }
  std::vector<BoolVar> to_bits_le(const FpVar& x, bool* ok) {
    std::vector<BoolVar> bits = to_non_unique_bits_le(x);
    if (!x.c) *ok &= enforce_smaller_or_equal_than_le(bits, raw_sub(MOD, Fr{{1, 0, 0, 0}}));  // enforce_in_field_le
    return bits;
  }
  // FpVar::enforce_cmp(self, other, Ordering::Greater, should_also_allow_equality = true)  (fields/fp/cmp.rs)
  bool enforce_greater_or_equal(const FpVar& self, const FpVar& other) {
    Fr half = raw_sub(MOD, Fr{{1, 0, 0, 0}});  // (r - 1) / 2
    for (int i = 0; i < 4; ++i) half.l[i] = (half.l[i] >> 1) | (i < 3 ? half.l[i + 1] << 63 : 0);
    const FpVar& left = other;
    FpVar right = add_const(self, K.one);
    bool ok = true;
    ok &= enforce_smaller_or_equal_than_le(to_non_unique_bits_le(left), half);
    ok &= enforce_smaller_or_equal_than_le(to_non_unique_bits_le(right), half);
    std::vector<BoolVar> bits = to_bits_le(fdouble(fsub(left, right)), &ok);
    const BoolVar& is_smaller = bits[0];
    if (S()) b.enforce(as_lc(is_smaller), lc_one(), lc_one());
    else b.skip();
    return ok;
  }

  // ---- PoseidonSpongeVar: absorb `n` elements into a fresh sponge, squeeze one (every use in the circuit)
  // ASSIGN mode: the same permutation on bare values.  Allocation order is what the generic path produces: per S-box on a
  // variable lane the witnesses x^2, x^4, x^5 (pow_by_constant(5) = square, square, multiply) and three constraints; a
  // lane that is still a constant (only possible in the first round) allocates nothing.
  void permute_assign(FpVar st[T]) {
    const PoseidonConfig& cfg = config();
    Fr v[T];
    bool c[T];
    for (int i = 0; i < T; ++i) {
      v[i] = st[i].v;
      c[i] = st[i].c;
    }
    for (int r = 0; r < ROUNDS; ++r) {
      const int lanes = (r < FULL / 2 || r >= FULL / 2 + PARTIAL) ? T : 1;
      for (int i = 0; i < T; ++i) v[i] = add(v[i], cfg.ark[r][i]);
      for (int i = 0; i < lanes; ++i) {
        Fr x2 = mont_mul(v[i], v[i]);
        Fr x4 = mont_mul(x2, x2);
        Fr x5 = mont_mul(x4, v[i]);
        if (!c[i]) {
          b.emit(x2);
          b.emit(x4);
          b.emit(x5);
          b.n_constraints += 3;
        }
        v[i] = x5;
      }
      Fr nw[T];
      for (int i = 0; i < T; ++i) {
        nw[i] = add(add(mont_mul(v[0], cfg.mds[i][0]), mont_mul(v[1], cfg.mds[i][1])), mont_mul(v[2], cfg.mds[i][2]));
      }
      bool any_var = !(c[0] && c[1] && c[2]);  // the MDS layer mixes every lane into every lane (no zero entries)
      for (int i = 0; i < T; ++i) {
        v[i] = nw[i];
        c[i] = !any_var;
      }
    }
    for (int i = 0; i < T; ++i) st[i] = FpVar{c[i], v[i], LC()};
  }
  void permute(FpVar st[T]) {
    static_assert(ALPHA == 5, "permute_assign hard-codes x^5");
    if (!S()) return permute_assign(st);
    const PoseidonConfig& cfg = config();
    for (int r = 0; r < ROUNDS; ++r) {
      for (int i = 0; i < T; ++i) st[i] = add_const(st[i], cfg.ark[r][i]);
      if (r < FULL / 2 || r >= FULL / 2 + PARTIAL) {
        for (int i = 0; i < T; ++i) st[i] = pow_by_constant(st[i], ALPHA);
      } else {
        st[0] = pow_by_constant(st[0], ALPHA);
      }
      FpVar nw[T];
      for (int i = 0; i < T; ++i) {
        FpVar cur = constant(K.zero);
        for (int j = 0; j < T; ++j) cur = fadd(cur, mul_const(st[j], cfg.mds[i][j]));
        nw[i] = cur;
      }
      for (int i = 0; i < T; ++i) st[i] = nw[i];
    }
  }
  FpVar hash_impl(const std::vector<const FpVar*>& elems) {
    FpVar st[T] = {constant(K.zero), constant(K.zero), constant(K.zero)};
    size_t start = 0, done = 0;  // DuplexSpongeMode::Absorbing { next_absorb_index: start }
    while (done < elems.size()) {
      if (start == RATE) {
        permute(st);
        start = 0;
      }
      st[CAPACITY + start] = fadd(st[CAPACITY + start], *elems[done]);
      ++start;
      ++done;
    }
    permute(st);  // squeeze from the absorbing mode
    return st[CAPACITY];
  }
  // Runs `fn` (a hash or a comparison) as one recorded call.  Sequential modes: just run it (STRUCTURE records its size).
  // Parallel ASSIGN: pass 1 computes the independent pieces (leaf hashes, comparisons) round-robin over the walkers, pass 2
  // the fold chains, one walker per chain, reading the leaf outputs of pass 1; a piece another walker owns is stepped over.
  template <class Fn>
  FpVar call(int cat, bool is_hash, Fn&& fn) {
    if (!b.parallel()) {
      const size_t p0 = b.pos;
      const uint64_t c0 = b.n_constraints;
      FpVar out = fn();
      if (S()) b.recs_out.push_back(CallRec{uint32_t(b.pos - p0), uint32_t(b.n_constraints - c0)});
      return out;
    }
    const size_t idx = b.call_idx++;
    const CallRec& rec = (*b.recs)[idx];
    bool compute;
    if (rec.nwit == 0) compute = true;  // constants only (a hash of a domain separator): free, and its value may be needed
    else if (b.pass == 1) compute = cat == CAT_LEAF && int(b.task_counter++ % size_t(b.nparts)) == b.part;
    else compute = cat == b.chain;
    if (!compute) {
      b.pos += rec.nwit;
      b.n_constraints += rec.ncons;
      // a leaf's value matters to the chain that folds it (pass 2); anything else stepped over is never looked at
      return FpVar{false, (b.pass == 2 && cat == CAT_LEAF && is_hash) ? b.hash_out[idx] : K.zero, LC()};
    }
    const bool saved = b.write;
    b.write = true;
    FpVar out = fn();
    b.write = saved;
    if (is_hash && rec.nwit) b.hash_out[idx] = out.v;
    return out;
  }
  FpVar hash(const std::vector<const FpVar*>& elems, int cat) {
    return call(cat, true, [&] { return hash_impl(elems); });
  }
  bool cmp_greater_or_equal(const FpVar& self, const FpVar& other) {
    bool ok = true;
    call(CAT_LEAF, false, [&] {
      ok = enforce_greater_or_equal(self, other);
      return constant(K.zero);
    });
    return ok;
  }
};

// ------------------------------------------------------------------------------------------ the circuit
typedef std::array<uint8_t, 32> Key;
struct Witness {  // L2BlockCircuit's private fields, initial_accounts already in BTreeMap order
  std::vector<std::pair<Key, uint64_t>> accounts;
  struct Tx { Key sender, recipient; uint64_t amount; };
  std::vector<Tx> txs;
  std::vector<Key> commitments;
  std::vector<std::pair<Key, uint64_t>> withdrawals;
};

static const Fr& ds_accounts() {
  static const Fr v = from_le_bytes_mod_order((const uint8_t*)"zelana:accounts-fold:v1", 23);
  return v;
}
static const Fr& ds_withdrawals() {
  static const Fr v = from_le_bytes_mod_order((const uint8_t*)"zelana:withdrawals:v1", 21);
  return v;
}
static const Fr& ds_batch() {
  static const Fr v = from_le_bytes_mod_order((const uint8_t*)"zelana:batch-hash:v1", 20);
  return v;
}

// l2_circuit.rs:179-505.  Returns SYNTH_MISSING for SynthesisError::AssignmentMissing (unknown sender, :263-266).
enum SynthStatus { SYNTH_OK = 0, SYNTH_MISSING = 1, SYNTH_CONTRADICTION = 2 };
static SynthStatus generate_constraints(Gadgets& g, const zkb_l2_public_inputs& in, const Witness& w) {
  FpVar pre_state_root = g.input(from_le_bytes_mod_order(in.pre_state_root, 32));
  FpVar expected_post_state = g.input(from_le_bytes_mod_order(in.post_state_root, 32));
  FpVar pre_shielded_root = g.input(from_le_bytes_mod_order(in.pre_shielded_root, 32));
  FpVar expected_post_shielded = g.input(from_le_bytes_mod_order(in.post_shielded_root, 32));
  FpVar expected_withdrawal_root = g.input(from_le_bytes_mod_order(in.withdrawal_root, 32));
  FpVar expected_batch_hash = g.input(from_le_bytes_mod_order(in.batch_hash, 32));
  FpVar batch_id = g.input(from_u64(in.batch_id));

  std::map<Key, FpVar> account_vars;
  for (const auto& a : w.accounts) account_vars[a.first] = g.witness(from_u64(a.second));

  bool ok = true;
  std::map<Key, FpVar> current = account_vars;
  for (const Witness::Tx& tx : w.txs) {
    FpVar amount = g.witness(from_u64(tx.amount));
    auto s = current.find(tx.sender);
    if (s == current.end()) return SYNTH_MISSING;
    FpVar sender_bal = s->second;
    auto r = current.find(tx.recipient);
    FpVar recipient_bal = r == current.end() ? g.constant(K.zero) : r->second;
    ok &= g.cmp_greater_or_equal(sender_bal, amount);
    FpVar new_sender = g.fsub(sender_bal, amount);
    FpVar new_recipient = g.fadd(recipient_bal, amount);
    current[tx.sender] = new_sender;
    current[tx.recipient] = new_recipient;
  }

  FpVar domain_separator = g.constant(ds_accounts());
  auto fold_accounts = [&](const std::map<Key, FpVar>& accounts, int chain) {
    FpVar state = g.hash({&domain_separator, &batch_id}, chain);
    for (const auto& kv : accounts) {
      FpVar pk = g.witness(from_le_bytes_mod_order(kv.first.data(), 32));
      FpVar leaf = g.hash({&pk, &kv.second}, CAT_LEAF);
      state = g.hash({&state, &leaf}, chain);
    }
    FpVar count = g.witness(from_u64(accounts.size()));
    return g.hash({&state, &count}, chain);
  };
  g.enforce_equal(fold_accounts(current, 0), expected_post_state);

  FpVar shielded_state = g.hash({&pre_shielded_root}, 1);
  for (const Key& c : w.commitments) {
    FpVar cv = g.witness(from_le_bytes_mod_order(c.data(), 32));
    shielded_state = g.hash({&shielded_state, &cv}, 1);
  }
  if (w.commitments.empty()) g.enforce_equal(pre_shielded_root, expected_post_shielded);
  else g.enforce_equal(shielded_state, expected_post_shielded);

  FpVar wd_ds = g.constant(ds_withdrawals());
  FpVar wd_state = g.hash({&wd_ds}, 2);
  for (const auto& wd : w.withdrawals) {
    FpVar recipient = g.witness(from_le_bytes_mod_order(wd.first.data(), 32));
    FpVar amount = g.witness(from_u64(wd.second));
    FpVar leaf = g.hash({&recipient, &amount}, CAT_LEAF);
    wd_state = g.hash({&wd_state, &leaf}, 2);
  }
  FpVar wd_count = g.witness(from_u64(w.withdrawals.size()));
  g.enforce_equal(g.hash({&wd_state, &wd_count}, 2), expected_withdrawal_root);

  FpVar batch_ds = g.constant(ds_batch());
  FpVar batch_state = g.hash({&batch_ds, &batch_id}, 3);
  for (const Witness::Tx& tx : w.txs) {
    FpVar sender = g.witness(from_le_bytes_mod_order(tx.sender.data(), 32));
    FpVar recipient = g.witness(from_le_bytes_mod_order(tx.recipient.data(), 32));
    FpVar amount = g.witness(from_u64(tx.amount));
    FpVar tx_hash = g.hash({&sender, &recipient, &amount}, CAT_LEAF);
    batch_state = g.hash({&batch_state, &tx_hash}, 3);
  }
  FpVar tx_count = g.witness(from_u64(w.txs.size()));
  g.enforce_equal(g.hash({&batch_state, &tx_count}, 3), expected_batch_hash);

  g.enforce_equal(fold_accounts(account_vars, 4), pre_state_root);
  return ok ? SYNTH_OK : SYNTH_CONTRADICTION;
}

// the part of a witness the matrices depend on: counts and which account each transfer touches
static std::vector<int64_t> shape_of(const Witness& w) {
  std::vector<int64_t> s = {(int64_t)w.accounts.size(), (int64_t)w.txs.size(), (int64_t)w.commitments.size(),
                            (int64_t)w.withdrawals.size()};
  std::map<Key, int64_t> idx;
  for (const auto& a : w.accounts) idx.emplace(a.first, (int64_t)idx.size());
  for (const Witness::Tx& tx : w.txs) {
    auto si = idx.find(tx.sender);
    s.push_back(si == idx.end() ? -1 : si->second);
    auto ri = idx.find(tx.recipient);
    if (ri == idx.end()) ri = idx.emplace(tx.recipient, (int64_t)idx.size()).first;  // created by this transfer
    s.push_back(ri->second);
  }
  // accounts created by transfers land in the BTreeMap by key order, which changes the fold order: record the ranks
  std::vector<Key> keys;
  for (const auto& kv : idx) keys.push_back(kv.first);  // std::map = sorted
  for (const Key& k : keys) s.push_back(idx[k]);
  return s;
}

static bool read_witness(const zkb_l2_witness* in, Witness* w, std::string* err) {
  if (!in) {
    *err = "null witness";
    return false;
  }
  auto key = [](const uint8_t* p, size_t i) {
    Key k;
    memcpy(k.data(), p + 32 * i, 32);
    return k;
  };
  if ((in->n_accounts && (!in->account_pks || !in->account_balances)) ||
      (in->n_txs && (!in->tx_senders || !in->tx_recipients || !in->tx_amounts)) || (in->n_commitments && !in->commitments) ||
      (in->n_withdrawals && (!in->wd_recipients || !in->wd_amounts))) {
    *err = "null array with a non-zero count";
    return false;
  }
  std::map<Key, uint64_t> acc;  // BTreeMap::insert: a repeated key keeps the last balance
  for (size_t i = 0; i < in->n_accounts; ++i) acc[key(in->account_pks, i)] = in->account_balances[i];
  w->accounts.assign(acc.begin(), acc.end());
  for (size_t i = 0; i < in->n_txs; ++i) w->txs.push_back({key(in->tx_senders, i), key(in->tx_recipients, i), in->tx_amounts[i]});
  for (size_t i = 0; i < in->n_commitments; ++i) w->commitments.push_back(key(in->commitments, i));
  for (size_t i = 0; i < in->n_withdrawals; ++i) w->withdrawals.push_back({key(in->wd_recipients, i), in->wd_amounts[i]});
  return true;
}

// ------------------------------------------------------------------------------------------ native sponge (off-circuit roots)
static void permute_native(Fr st[T]) {
  const PoseidonConfig& cfg = config();
  for (int r = 0; r < ROUNDS; ++r) {
    for (int i = 0; i < T; ++i) st[i] = add(st[i], cfg.ark[r][i]);
    int n = (r < FULL / 2 || r >= FULL / 2 + PARTIAL) ? T : 1;
    for (int i = 0; i < n; ++i) {
      Fr x2 = mont_mul(st[i], st[i]);
      st[i] = mont_mul(mont_mul(x2, x2), st[i]);
    }
    Fr nw[T];
    for (int i = 0; i < T; ++i) {
      nw[i] = K.zero;
      for (int j = 0; j < T; ++j) nw[i] = add(nw[i], mont_mul(st[j], cfg.mds[i][j]));
    }
    for (int i = 0; i < T; ++i) st[i] = nw[i];
  }
}
static Fr hash_native(std::initializer_list<Fr> elems) {
  Fr st[T] = {K.zero, K.zero, K.zero};
  size_t start = 0;
  for (const Fr& e : elems) {
    if (start == RATE) {
      permute_native(st);
      start = 0;
    }
    st[CAPACITY + start] = add(st[CAPACITY + start], e);
    ++start;
  }
  permute_native(st);
  return st[CAPACITY];
}
static Fr accounts_root(uint64_t batch_id, const std::map<Key, uint64_t>& accounts) {  // main.rs.bak:114-154
  Fr state = hash_native({ds_accounts(), from_u64(batch_id)});
  for (const auto& kv : accounts) {
    Fr leaf = hash_native({from_le_bytes_mod_order(kv.first.data(), 32), from_u64(kv.second)});
    state = hash_native({state, leaf});
  }
  return hash_native({state, from_u64(accounts.size())});
}

// ------------------------------------------------------------------------------------------ rand 0.8.5 StdRng + Fr::rand
struct StdRng {  // ChaCha12, 4-block buffer; seed_from_u64 = rand_core 0.6.4 PCG32 expansion
  uint32_t key[8];
  uint64_t block = 0;
  uint32_t words[64];
  int pos = 64;
  explicit StdRng(uint64_t state) {
    for (int i = 0; i < 8; ++i) {
      state = state * 6364136223846793005ull + 11634580027462260723ull;
      uint32_t x = (uint32_t)(((state >> 18) ^ state) >> 27);
      uint32_t rot = (uint32_t)(state >> 59);
      key[i] = rot ? ((x >> rot) | (x << (32 - rot))) : x;
    }
  }
  static inline uint32_t rotl(uint32_t v, int n) { return (v << n) | (v >> (32 - n)); }
  void refill() {
    for (int blk = 0; blk < 4; ++blk) {
      uint64_t ctr = block + blk;
      uint32_t init[16] = {0x61707865u, 0x3320646Eu, 0x79622D32u, 0x6B206574u, key[0], key[1], key[2], key[3], key[4],
                           key[5],      key[6],      key[7],      (uint32_t)ctr, (uint32_t)(ctr >> 32), 0, 0};
      uint32_t x[16];
      memcpy(x, init, sizeof x);
      auto qr = [&](int a, int b, int c, int d) {
        x[a] += x[b]; x[d] = rotl(x[d] ^ x[a], 16);
        x[c] += x[d]; x[b] = rotl(x[b] ^ x[c], 12);
        x[a] += x[b]; x[d] = rotl(x[d] ^ x[a], 8);
        x[c] += x[d]; x[b] = rotl(x[b] ^ x[c], 7);
      };
      for (int i = 0; i < 6; ++i) {
        qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15);
        qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14);
      }
      for (int i = 0; i < 16; ++i) words[16 * blk + i] = x[i] + init[i];
    }
    block += 4;
  }
  uint64_t next_u64() {  // BlockRng::next_u64: low word first; a read straddling the buffer end refills in between
    uint64_t lo, hi;
    if (pos < 63) {
      lo = words[pos];
      hi = words[pos + 1];
      pos += 2;
    } else if (pos >= 64) {
      refill();
      lo = words[0];
      hi = words[1];
      pos = 2;
    } else {
      lo = words[63];
      refill();
      hi = words[0];
      pos = 1;
    }
    return (hi << 32) | lo;
  }
  Fr fr_rand() {  // ark-ff 0.5.0 UniformRand for Fp: four limbs, top two bits cleared, taken AS the Montgomery form
    for (;;) {
      Fr x;
      for (int i = 0; i < 4; ++i) x.l[i] = next_u64();
      x.l[3] &= ~0ull >> 2;
      if (!geq(x, MOD)) return x;
    }
  }
};

}  // namespace l2

// ================================================================================================ C ABI
struct zkb_l2_circuit {
  l2::Csr a, b, c;
  uint64_t num_constraints = 0, num_witness = 0;
  std::vector<int64_t> shape;
  std::vector<l2::CallRec> recs;  // sizes of the hashes / comparisons in call order (parallel assignment)
  std::string err;
};

// assignment threads: ZKB_L2_ASSIGN_THREADS (1..64), default min(8, hardware threads); circuits below PAR_MIN_WITNESS
// variables (the dummy shape: 0.6 ms on one thread) are not worth a thread start
static int l2_assign_threads() {
  static const int n = [] {
    if (const char* e = getenv("ZKB_L2_ASSIGN_THREADS")) {
      int v = atoi(e);
      if (v >= 1 && v <= 64) return v;
    }
    unsigned hw = std::thread::hardware_concurrency();
    return int(hw == 0 ? 1 : (hw > 8 ? 8 : hw));
  }();
  return n;
}
constexpr uint64_t PAR_MIN_WITNESS = 40000;

static thread_local std::string g_l2_error;
const char* zkb_l2_last_error(void) { return g_l2_error.c_str(); }

// Joins whatever was started when the scope ends, also while an exception unwinds: a std::thread destroyed joinable calls
// std::terminate, and emplace_back can throw std::system_error half way through starting the workers.
struct JoinGuard {
  std::vector<std::thread> th;
  ~JoinGuard() {
    for (auto& t : th)
      if (t.joinable()) t.join();
  }
};

// Nothing may unwind through the C ABI (the host is Rust: UB): every exported zkb_l2_* body ends in these handlers.
#define ZKB_L2_CATCH_ALL()                    \
  catch (const std::bad_alloc&) {             \
    g_l2_error = "out of host memory";        \
    return ZKB_ERR_OOM;                       \
  }                                           \
  catch (const std::exception& e) {           \
    g_l2_error = e.what();                    \
    return ZKB_ERR_INVALID_ARG;               \
  }                                           \
  catch (...) {                               \
    g_l2_error = "unknown C++ exception";     \
    return ZKB_ERR_INVALID_ARG;               \
  }

int zkb_l2_circuit_create(const zkb_l2_witness* shape, zkb_l2_circuit** out) {
  if (!out) return ZKB_ERR_INVALID_ARG;
  *out = nullptr;
  try {
    l2::Witness w;
    if (!l2::read_witness(shape, &w, &g_l2_error)) return ZKB_ERR_INVALID_ARG;
    zkb_l2_public_inputs zero_inputs;
    memset(&zero_inputs, 0, sizeof zero_inputs);
    l2::Builder bld(true);
    l2::Gadgets g(bld);
    // only a missing sender stops the synthesis; values (and contradictions among them) do not matter for the structure
    if (l2::generate_constraints(g, zero_inputs, w) == l2::SYNTH_MISSING || bld.n_inputs != l2::NUM_INSTANCE) {
      g_l2_error = "a transfer names a sender that is not among the accounts (AssignmentMissing, l2_circuit.rs:263-266)";
      return ZKB_ERR_SHAPE;
    }
    zkb_l2_circuit* c = new zkb_l2_circuit();
    c->a = std::move(bld.A);
    c->b = std::move(bld.B);
    c->c = std::move(bld.C);
    c->a.seal();
    c->b.seal();
    c->c.seal();
    c->num_constraints = bld.n_constraints;
    c->num_witness = bld.pos - l2::NUM_INSTANCE;
    c->recs = std::move(bld.recs_out);
    c->shape = l2::shape_of(w);
    *out = c;
    return ZKB_OK;
  } catch (const std::bad_alloc&) {
    g_l2_error = "out of host memory";
    return ZKB_ERR_OOM;
  } catch (const std::exception& e) {
    g_l2_error = e.what();
    return ZKB_ERR_INVALID_ARG;
  }
}

void zkb_l2_circuit_free(zkb_l2_circuit* c) { delete c; }

int zkb_l2_circuit_desc(const zkb_l2_circuit* c, zkb_r1cs_desc* out) {
  if (!c || !out) return ZKB_ERR_INVALID_ARG;
  out->num_constraints = c->num_constraints;
  out->num_instance = l2::NUM_INSTANCE;
  out->num_witness = c->num_witness;
  const l2::Csr* m[3] = {&c->a, &c->b, &c->c};
  zkb_csr* o[3] = {&out->a, &out->b, &out->c};
  for (int i = 0; i < 3; ++i) {
    o[i]->row_ptr = m[i]->row_ptr.data();
    o[i]->col = m[i]->col.data();
    o[i]->coeff = m[i]->bytes.data();
  }
  return ZKB_OK;
}

static int l2_assign(const zkb_l2_circuit* c, const zkb_l2_public_inputs* inputs, const zkb_l2_witness* witness,
                     std::vector<l2::Fr>* z) {
  if (!c || !inputs) return ZKB_ERR_INVALID_ARG;
  l2::Witness w;
  if (!l2::read_witness(witness, &w, &g_l2_error)) return ZKB_ERR_INVALID_ARG;
  if (l2::shape_of(w) != c->shape) {
    g_l2_error = "witness shape (account / transfer / commitment / withdrawal pattern) differs from the circuit the key was made for";
    return ZKB_ERR_SHAPE;
  }
  const size_t total = l2::NUM_INSTANCE + c->num_witness;
  const int nthreads = l2_assign_threads();
  if (nthreads > 1 && c->num_witness >= PAR_MIN_WITNESS) {
    // Two passes over the same statements, every walker stepping over the pieces it does not own (sizes from the structure
    // pass): 1. leaf hashes and balance comparisons, round-robin over `nthreads` walkers; 2. the five fold chains, one
    // walker each, reading the leaf outputs.  All walkers index one shared array and write disjoint pieces of it.
    z->assign(total, l2::K.zero);
    std::vector<l2::Fr> hash_out(c->recs.size(), l2::K.zero);
    std::vector<int> bad(size_t(nthreads > l2::NUM_CHAINS ? nthreads : l2::NUM_CHAINS), 0);
    auto walk = [&](int pass, int part, int nparts, int chain) {
      l2::Builder bld(z->data(), &c->recs, hash_out.data(), pass, part, nparts, chain);
      l2::Gadgets g(bld);
      l2::generate_constraints(g, *inputs, w);
      if (bld.n_constraints != c->num_constraints || bld.pos != total || bld.call_idx != c->recs.size()) bad[size_t(part)] = 1;
    };
    for (int pass = 1; pass <= 2; ++pass) {
      const int n = pass == 1 ? nthreads : l2::NUM_CHAINS;
      JoinGuard jg;
      for (int t = 1; t < n; ++t) jg.th.emplace_back(walk, pass, t, n, pass == 1 ? -2 : t);
      walk(pass, 0, n, pass == 1 ? -2 : 0);
    }
    for (int x : bad) {
      if (x) {
        g_l2_error = "internal: a parallel assignment walker disagrees with the structure pass";
        return ZKB_ERR_SHAPE;
      }
    }
    return ZKB_OK;
  }
  l2::Builder bld(false);
  bld.zown.reserve(total);
  l2::Gadgets g(bld);
  l2::generate_constraints(g, *inputs, w);
  if (bld.n_constraints != c->num_constraints || bld.pos != total) {
    g_l2_error = "internal: assignment pass disagrees with the structure pass";
    return ZKB_ERR_SHAPE;
  }
  *z = std::move(bld.zown);
  return ZKB_OK;
}

// z (Montgomery) -> canonical little-endian bytes, split over the assignment threads when it is long
static void l2_z_to_bytes(const std::vector<l2::Fr>& z, uint8_t* out) {
  const int nthreads = l2_assign_threads();
  auto conv = [&](size_t lo, size_t hi) {
    for (size_t i = lo; i < hi; ++i) l2::to_le_bytes(z[i], out + 32 * i);
  };
  if (nthreads <= 1 || z.size() < PAR_MIN_WITNESS) return conv(0, z.size());
  JoinGuard jg;
  const size_t step = (z.size() + size_t(nthreads) - 1) / size_t(nthreads);
  for (int t = 1; t < nthreads; ++t) jg.th.emplace_back(conv, std::min(z.size(), step * size_t(t)), std::min(z.size(), step * size_t(t + 1)));
  conv(0, std::min(z.size(), step));
}

int zkb_l2_circuit_assign(const zkb_l2_circuit* c, const zkb_l2_public_inputs* inputs, const zkb_l2_witness* witness,
                          uint8_t* z_out) {
  if (!z_out) return ZKB_ERR_INVALID_ARG;
  try {
    std::vector<l2::Fr> z;
    int rc = l2_assign(c, inputs, witness, &z);
    if (rc != ZKB_OK) return rc;
    l2_z_to_bytes(z, z_out);
    return ZKB_OK;
  }
  ZKB_L2_CATCH_ALL()
}

int zkb_l2_circuit_is_satisfied(const zkb_l2_circuit* c, const uint8_t* z, int* satisfied, uint64_t* first_bad_row) {
  if (!c || !z || !satisfied) return ZKB_ERR_INVALID_ARG;
  try {
  size_t nz = l2::NUM_INSTANCE + c->num_witness;
  std::vector<l2::Fr> zm(nz);
  for (size_t i = 0; i < nz; ++i) zm[i] = l2::from_le_bytes_mod_order(z + 32 * i, 32);
  auto dot = [&](const l2::Csr& m, uint64_t row) {
    l2::Fr acc = l2::K.zero;
    for (uint64_t k = m.row_ptr[row]; k < m.row_ptr[row + 1]; ++k) acc = l2::add(acc, l2::mont_mul(m.co[k], zm[m.col[k]]));
    return acc;
  };
  *satisfied = 1;
  for (uint64_t r = 0; r < c->num_constraints; ++r) {
    if (l2::mont_mul(dot(c->a, r), dot(c->b, r)) != dot(c->c, r)) {
      *satisfied = 0;
      if (first_bad_row) *first_bad_row = r;
      break;
    }
  }
  return ZKB_OK;
  }
  ZKB_L2_CATCH_ALL()
}

int zkb_l2_roots(const zkb_l2_witness* witness, uint64_t batch_id, const uint8_t pre_shielded_root[32],
                 zkb_l2_public_inputs* out) {
  if (!out || !pre_shielded_root) return ZKB_ERR_INVALID_ARG;
  try {
    l2::Witness w;
    if (!l2::read_witness(witness, &w, &g_l2_error)) return ZKB_ERR_INVALID_ARG;
    std::map<l2::Key, uint64_t> before(w.accounts.begin(), w.accounts.end()), after = before;
    for (const auto& tx : w.txs) {
      auto s = after.find(tx.sender);
      if (s == after.end() || s->second < tx.amount) {
        g_l2_error = "transfer from an unknown account or beyond its balance";
        return ZKB_ERR_INVALID_ARG;
      }
      // as the circuit does (l2_circuit.rs:263-300): both balances are read before either is written, so a transfer to
      // oneself ends with balance + amount (reference quirk, preserved)
      auto r = after.find(tx.recipient);
      uint64_t sender_bal = s->second, recipient_bal = r == after.end() ? 0 : r->second;
      if (recipient_bal + tx.amount < recipient_bal) {  // the circuit adds in the field; a u64 that wraps cannot match it
        g_l2_error = "a recipient balance exceeds 64 bits";
        return ZKB_ERR_INVALID_ARG;
      }
      after[tx.sender] = sender_bal - tx.amount;
      after[tx.recipient] = recipient_bal + tx.amount;
    }
    memset(out, 0, sizeof *out);
    out->batch_id = batch_id;
    l2::to_le_bytes(l2::accounts_root(batch_id, before), out->pre_state_root);
    l2::to_le_bytes(l2::accounts_root(batch_id, after), out->post_state_root);
    memcpy(out->pre_shielded_root, pre_shielded_root, 32);
    if (w.commitments.empty()) {
      memcpy(out->post_shielded_root, pre_shielded_root, 32);
    } else {
      l2::Fr st = l2::hash_native({l2::from_le_bytes_mod_order(pre_shielded_root, 32)});
      for (const auto& cm : w.commitments) st = l2::hash_native({st, l2::from_le_bytes_mod_order(cm.data(), 32)});
      l2::to_le_bytes(st, out->post_shielded_root);
    }
    l2::Fr wd = l2::hash_native({l2::ds_withdrawals()});
    for (const auto& x : w.withdrawals) {
      wd = l2::hash_native({wd, l2::hash_native({l2::from_le_bytes_mod_order(x.first.data(), 32), l2::from_u64(x.second)})});
    }
    l2::to_le_bytes(l2::hash_native({wd, l2::from_u64(w.withdrawals.size())}), out->withdrawal_root);
    l2::Fr bh = l2::hash_native({l2::ds_batch(), l2::from_u64(batch_id)});
    for (const auto& tx : w.txs) {
      l2::Fr h = l2::hash_native({l2::from_le_bytes_mod_order(tx.sender.data(), 32),
                                  l2::from_le_bytes_mod_order(tx.recipient.data(), 32), l2::from_u64(tx.amount)});
      bh = l2::hash_native({bh, h});
    }
    l2::to_le_bytes(l2::hash_native({bh, l2::from_u64(w.txs.size())}), out->batch_hash);
    return ZKB_OK;
  }
  ZKB_L2_CATCH_ALL()
}

int zkb_l2_poseidon_hash(const uint8_t* elems, size_t n, uint8_t out[32]) {
  if ((!elems && n) || !out || n > 3) return ZKB_ERR_INVALID_ARG;
  l2::Fr e[3];
  for (size_t i = 0; i < n; ++i) e[i] = l2::from_le_bytes_mod_order(elems + 32 * i, 32);
  l2::Fr h = n == 0 ? l2::hash_native({}) : n == 1 ? l2::hash_native({e[0]}) : n == 2 ? l2::hash_native({e[0], e[1]})
                                                                                      : l2::hash_native({e[0], e[1], e[2]});
  l2::to_le_bytes(h, out);
  return ZKB_OK;
}

// get_poseidon_config() (prover/src/l2_circuit.rs:68-83: find_poseidon_ark_and_mds for 254 bits, rate 2, 8 full + 56 partial
// rounds) as canonical bytes: ark_out = 64 rounds x 3 lanes x 32 B, mds_out = 3 x 3 x 32 B (row-major).  What the GPU hash
// kernel behind zkb_l2_poseidon_hash_batch uploads; also a parity hook against the oracle's parameter generator.
int zkb_l2_poseidon_params(uint8_t* ark_out, uint8_t* mds_out) {
  if (!ark_out || !mds_out) return ZKB_ERR_INVALID_ARG;
  try {
    const l2::PoseidonConfig& cfg = l2::config();
    for (int r = 0; r < l2::ROUNDS; ++r)
      for (int i = 0; i < l2::T; ++i) l2::to_le_bytes(cfg.ark[r][i], ark_out + 32 * (r * l2::T + i));
    for (int i = 0; i < l2::T; ++i)
      for (int j = 0; j < l2::T; ++j) l2::to_le_bytes(cfg.mds[i][j], mds_out + 32 * (i * l2::T + j));
    return ZKB_OK;
  }
  ZKB_L2_CATCH_ALL()
}

// n independent hashes on `threads` host threads with the native sponge the witness walkers use: the host side of the
// GPU-versus-host measurement of the leaf hashes (tools/poseidon_leaf_bench.py), and a batch convenience for hosts.
int zkb_l2_poseidon_hash_batch_host(int arity, const uint8_t* in, size_t n, int threads, uint8_t* out) {
  if (arity < 0 || arity > 3 || (n && (!out || (arity && !in))) || threads < 1 || threads > 256) return ZKB_ERR_INVALID_ARG;
  try {
    auto work = [&](size_t lo, size_t hi) {
      for (size_t i = lo; i < hi; ++i) {
        l2::Fr e[3];
        for (int k = 0; k < arity; ++k) e[k] = l2::from_le_bytes_mod_order(in + 32 * (i * size_t(arity) + k), 32);
        l2::Fr h = arity == 0 ? l2::hash_native({}) : arity == 1 ? l2::hash_native({e[0]})
                   : arity == 2 ? l2::hash_native({e[0], e[1]}) : l2::hash_native({e[0], e[1], e[2]});
        l2::to_le_bytes(h, out + 32 * i);
      }
    };
    JoinGuard jg;
    const size_t step = (n + size_t(threads) - 1) / size_t(threads);
    for (int t = 1; t < threads; ++t) jg.th.emplace_back(work, std::min(n, step * size_t(t)), std::min(n, step * size_t(t + 1)));
    work(0, std::min(n, step));
    return ZKB_OK;
  }
  ZKB_L2_CATCH_ALL()
}

int zkb_l2_prover_randomness(uint64_t batch_id, uint8_t r[32], uint8_t s[32]) {
  if (!r || !s) return ZKB_ERR_INVALID_ARG;
  l2::StdRng rng(batch_id);
  l2::to_le_bytes(rng.fr_rand(), r);
  l2::to_le_bytes(rng.fr_rand(), s);
  return ZKB_OK;
}

// proof_to_solana_bytes (prover.rs:304-334): -A || B || C, 32-byte LE coordinates
static void l2_format_solana(const uint8_t a[64], const uint8_t b[128], const uint8_t cc[64], uint8_t proof_out[256]) {
  static const uint64_t FQ[4] = {0x3c208c16d87cfd47ull, 0x97816a916871ca8dull, 0xb85045b68181585dull, 0x30644e72e131a029ull};
  bool a_inf = true;
  for (int i = 0; i < 64; ++i) a_inf &= a[i] == 0;
  memcpy(proof_out, a, 32);
  if (a_inf) {
    memset(proof_out + 32, 0, 32);
  } else {
    uint64_t y[4], o[4];
    memcpy(y, a + 32, 32);
    bool y_zero = (y[0] | y[1] | y[2] | y[3]) == 0;
    l2::u128 br = 0;
    for (int i = 0; i < 4; ++i) {
      l2::u128 t = (l2::u128)FQ[i] - y[i] - (uint64_t)br;
      o[i] = (uint64_t)t;
      br = (t >> 64) & 1;
    }
    if (y_zero) memset(o, 0, sizeof o);
    memcpy(proof_out + 32, o, 32);
  }
  memcpy(proof_out + 64, b, 128);
  memcpy(proof_out + 192, cc, 64);
}

int zkb_l2_prove(zkb_ctx* ctx, const zkb_pk* pk, const zkb_r1cs* m, const zkb_l2_circuit* c, const zkb_l2_public_inputs* inputs,
                 const zkb_l2_witness* witness, uint8_t proof_out[256]) {
  if (!ctx || !pk || !m || !proof_out) return ZKB_ERR_INVALID_ARG;
  try {
    std::vector<l2::Fr> z;
    int rc = l2_assign(c, inputs, witness, &z);
    if (rc != ZKB_OK) return rc;
    if (zkb_r1cs_num_variables(m) != z.size()) {  // m / pk made for another circuit shape: zkb_prove would read past zb
      g_l2_error = "zkb_l2_prove: the matrices have " + std::to_string(zkb_r1cs_num_variables(m)) + " variables, the circuit " +
                   std::to_string(z.size());
      return ZKB_ERR_SHAPE;
    }
    std::vector<uint8_t> zb(z.size() * 32);
    l2_z_to_bytes(z, zb.data());
    uint8_t r[32], s[32];
    zkb_l2_prover_randomness(inputs->batch_id, r, s);
    uint8_t a[64], b[128], cc[64];
    rc = zkb_prove(ctx, pk, m, zb.data(), r, s, a, b, cc);
    if (rc != ZKB_OK) {
      g_l2_error = zkb_last_error(ctx);
      return rc;
    }
    l2_format_solana(a, b, cc, proof_out);
    return ZKB_OK;
  }
  ZKB_L2_CATCH_ALL()
}

// ================================================================================================ batches of proofs
// BASELINE.json config 5 ("batch of 64 independent L2 proofs"; the forge coordinator's chunk-per-worker dispatch,
// forge/crates/prover-coordinator/src/dispatcher.rs:290-330, inside one process).  Round 1 ran many small proofs side by side,
// each a chain of ~100 short kernels at 0.02-0.1 waves.  Now a batch is cut into sub-batches of up to ZKB_L2_SUBBATCH (256)
// proofs; a sub-batch is ASSIGNED on the host by a pool of `lanes` worker threads (Poseidon folds, comparison bits: 0.6 ms
// per proof per core) straight into pinned memory and PROVED on the GPU by one zkb_prove_batch_begin: batched mat-vecs and
// NTTs, five batched MSMs sharing the key's window tables, one finishing kernel.  A few device slots (ZKB_L2_SLOTS, default 4) rotate, so the
// assignment of sub-batch j+1 runs while the GPU proves sub-batch j.  Key, matrices and circuit are shared, read-only.

namespace {

class WorkerPool {
 public:
  explicit WorkerPool(int nthreads) {
    for (int i = 1; i < nthreads; ++i) threads_.emplace_back([this] { loop(); });
  }
  ~WorkerPool() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_ = true;
    }
    cv_.notify_all();
    for (auto& t : threads_)
      if (t.joinable()) t.join();
  }
  // fn(i) for every i < count, on the pool's threads and the caller's; returns when all are done.  fn must not throw.
  void parallel_for(size_t count, const std::function<void(size_t)>& fn) {
    if (count == 0) return;
    {
      std::lock_guard<std::mutex> lk(mu_);
      fn_ = &fn;
      count_ = count;
      next_.store(0);
      pending_ = count;
      ++generation_;
    }
    cv_.notify_all();
    drain();
    std::unique_lock<std::mutex> lk(mu_);
    done_cv_.wait(lk, [this] { return pending_ == 0; });
    fn_ = nullptr;
  }

 private:
  void drain() {
    size_t finished = 0;
    for (;;) {
      size_t i = next_.fetch_add(1);
      if (i >= count_) break;
      (*fn_)(i);
      ++finished;
    }
    if (finished) {
      std::lock_guard<std::mutex> lk(mu_);
      pending_ -= finished;
      if (pending_ == 0) done_cv_.notify_all();
    }
  }
  void loop() {
    unsigned long long seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return stop_ || generation_ != seen; });
        if (stop_) return;
        seen = generation_;
      }
      drain();
    }
  }
  std::vector<std::thread> threads_;
  std::mutex mu_;
  std::condition_variable cv_, done_cv_;
  const std::function<void(size_t)>* fn_ = nullptr;
  size_t count_ = 0, pending_ = 0;
  std::atomic<size_t> next_{0};
  unsigned long long generation_ = 0;
  bool stop_ = false;
};

struct BatchSlot {
  zkb_ctx* ctx = nullptr;
  uint8_t* z = nullptr;       // pinned: count x nv x 32
  uint8_t* rs = nullptr;      // pinned: count x 64
  size_t z_cap = 0, rs_cap = 0;
  std::vector<uint8_t> out;   // count x 256 (A | B | C)
  size_t base = 0, count = 0;
  bool busy = false;
};

}  // namespace

constexpr int L2_MAX_SLOTS = 8;
constexpr size_t L2_SMALL_BATCH = 8;   // <= L2_MAX_SLOTS: calls of at most this many proofs run zkb_prove per proof, one context each

// device slots in flight per batch object: env ZKB_L2_SLOTS (2..8), default 4.  More than two, because a sub-batch ends in
// latency-bound kernels (bucket folds, the two scalar-multiplication chains) that only overlap with ANOTHER sub-batch's
// accumulation kernels.
static int l2_slot_count() {
  static const int n = [] {
    if (const char* e = getenv("ZKB_L2_SLOTS")) {
      int v = atoi(e);
      if (v >= 2 && v <= L2_MAX_SLOTS) return v;
    }
    return 4;
  }();
  return n;
}

struct zkb_l2_batch {
  int device = 0;
  int lanes = 1;
  int nslots = 2;
  std::unique_ptr<WorkerPool> pool;
  BatchSlot slots[L2_MAX_SLOTS];
  std::mutex mu;   // one zkb_l2_batch_prove at a time per batch object
};

int zkb_l2_batch_create(int device, int lanes, zkb_l2_batch** out) {
  if (!out || lanes < 1 || lanes > 64) return ZKB_ERR_INVALID_ARG;
  *out = nullptr;
  try {
    std::unique_ptr<zkb_l2_batch> b(new zkb_l2_batch());
    b->device = device;
    b->lanes = lanes;
    b->nslots = l2_slot_count();
    for (int k = 0; k < b->nslots; ++k) {
      BatchSlot& s = b->slots[k];
      int rc = zkb_ctx_create(device, &s.ctx);
      if (rc != ZKB_OK) {
        g_l2_error = "zkb_l2_batch_create: zkb_ctx_create failed";
        for (auto& t : b->slots)
          if (t.ctx) zkb_ctx_destroy(t.ctx);
        return rc;
      }
      zkb_ctx_set_blocking_sync(s.ctx, 1);  // the dispatcher sleeps, not spins, while the GPU works: the cores belong to the pool
    }
    b->pool.reset(new WorkerPool(lanes));
    *out = b.release();
    return ZKB_OK;
  }
  ZKB_L2_CATCH_ALL()
}

void zkb_l2_batch_destroy(zkb_l2_batch* b) {
  if (!b) return;
  b->pool.reset();
  for (auto& s : b->slots) {
    if (s.ctx) {
      zkb_ctx_synchronize(s.ctx);
      if (s.z) zkb_host_free_pinned(s.z);
      if (s.rs) zkb_host_free_pinned(s.rs);
      zkb_ctx_destroy(s.ctx);
    }
  }
  delete b;
}

int zkb_l2_batch_lanes(const zkb_l2_batch* b) { return b ? b->lanes : 0; }

static size_t l2_subbatch(size_t n) {
  static const size_t cap = [] {
    if (const char* e = getenv("ZKB_L2_SUBBATCH")) {
      long v = atol(e);
      if (v >= 1 && v <= 4096) return size_t(v);
    }
    return size_t(256);
  }();
  static const size_t split = [] {
    if (const char* e = getenv("ZKB_L2_SPLIT")) {
      long v = atol(e);
      if (v >= 1 && v <= 64) return size_t(v);
    }
    return size_t(4);
  }();
  size_t kb = (n + split - 1) / split;   // at least `split` (4) sub-batches per call, so that assignment and proving overlap
  if (kb < 16) kb = 16;
  if (kb > cap) kb = cap;
  return kb;
}

int zkb_l2_batch_prove(zkb_l2_batch* b, const zkb_pk* pk, const zkb_r1cs* m, const zkb_l2_circuit* c,
                       const zkb_l2_public_inputs* inputs, const zkb_l2_witness* witnesses, size_t n, uint8_t* proofs_out,
                       int* status_out) {
  if (!b || !pk || !m || !c || (n && (!inputs || !witnesses || !proofs_out))) return ZKB_ERR_INVALID_ARG;
  try {
    std::lock_guard<std::mutex> guard(b->mu);
    const size_t nv = l2::NUM_INSTANCE + c->num_witness;
    if (zkb_r1cs_num_variables(m) != nv) {
      g_l2_error = "zkb_l2_batch_prove: the matrices have " + std::to_string(zkb_r1cs_num_variables(m)) + " variables, the circuit " +
                   std::to_string(nv);
      return ZKB_ERR_SHAPE;
    }
    std::vector<int> status(n, ZKB_OK);
    std::mutex err_mu;
    int first_rc = ZKB_OK;
    std::string first_msg;
    auto note = [&](int rc, const std::string& msg) {
      std::lock_guard<std::mutex> lk(err_mu);
      if (first_rc == ZKB_OK) {
        first_rc = rc;
        first_msg = msg;
      }
    };
    auto finish = [&](BatchSlot& s) {
      if (!s.busy) return;
      s.busy = false;
      s.out.resize(s.count * 256);
      int rc = zkb_prove_batch_end(s.ctx, s.count, s.out.data());
      if (rc != ZKB_OK) {
        note(rc, zkb_last_error(s.ctx));
        for (size_t k = 0; k < s.count; ++k)
          if (status[s.base + k] == ZKB_OK) status[s.base + k] = rc;
        return;
      }
      for (size_t k = 0; k < s.count; ++k) {
        if (status[s.base + k] != ZKB_OK) continue;
        const uint8_t* r = s.out.data() + 256 * k;
        l2_format_solana(r, r + 64, r + 192, proofs_out + 256 * (s.base + k));
      }
    };
    if (n >= 1 && n <= L2_SMALL_BATCH) {
      // A handful of proofs: the batched kernels are latency-bound here (one sub-batch of 8 takes 6.9 ms, most of it the
      // 254-step finishing chains), while zkb_prove -- which trades those chains for two more small MSMs and replays a captured
      // graph -- takes 1.7 ms per proof and runs side by side on separate contexts: 8 proofs in 4.6 ms.  One context per proof.
      for (size_t k = 0; k < n; ++k) {
        if (b->slots[k].ctx) continue;
        int rc = zkb_ctx_create(b->device, &b->slots[k].ctx);
        if (rc != ZKB_OK) {
          g_l2_error = "zkb_l2_batch_prove: zkb_ctx_create failed";
          return rc;
        }
        zkb_ctx_set_blocking_sync(b->slots[k].ctx, 1);
      }
      std::function<void(size_t)> prove_one = [&](size_t i) {
        int rc;
        try {
          std::vector<l2::Fr> z;
          rc = l2_assign(c, inputs + i, witnesses + i, &z);
          if (rc == ZKB_OK && z.size() != nv) rc = ZKB_ERR_SHAPE;
          if (rc == ZKB_OK) {
            std::vector<uint8_t> zb(nv * 32);
            l2_z_to_bytes(z, zb.data());
            uint8_t r[32], sc[32], pa[64], pb[128], pc[64];
            zkb_l2_prover_randomness(inputs[i].batch_id, r, sc);
            rc = zkb_prove(b->slots[i].ctx, pk, m, zb.data(), r, sc, pa, pb, pc);
            if (rc == ZKB_OK) l2_format_solana(pa, pb, pc, proofs_out + 256 * i);
            else g_l2_error = zkb_last_error(b->slots[i].ctx);
          }
        } catch (const std::bad_alloc&) {
          rc = ZKB_ERR_OOM;
          g_l2_error = "out of host memory";
        } catch (...) {
          rc = ZKB_ERR_INVALID_ARG;
          g_l2_error = "exception in the witness assignment";
        }
        if (rc != ZKB_OK) {
          status[i] = rc;
          note(rc, g_l2_error);
        }
      };
      b->pool->parallel_for(n, prove_one);
      if (status_out)
        for (size_t i = 0; i < n; ++i) status_out[i] = status[i];
      if (first_rc != ZKB_OK) {
        g_l2_error = first_msg;
        return first_rc;
      }
      return ZKB_OK;
    }
    const size_t kb = l2_subbatch(n);
    for (size_t j = 0; j * kb < n; ++j) {
      BatchSlot& s = b->slots[j % size_t(b->nslots)];
      finish(s);
      s.base = j * kb;
      s.count = n - s.base < kb ? n - s.base : kb;
      if (s.z_cap < s.count * nv * 32) {
        if (s.z) zkb_host_free_pinned(s.z);
        s.z = nullptr;
        s.z_cap = 0;
        void* p = nullptr;
        int rc = zkb_host_alloc_pinned(kb * nv * 32, &p);
        if (rc != ZKB_OK) {
          g_l2_error = "zkb_l2_batch_prove: pinned host allocation failed";
          return rc;
        }
        s.z = static_cast<uint8_t*>(p);
        s.z_cap = kb * nv * 32;
      }
      if (s.rs_cap < s.count * 64) {
        if (s.rs) zkb_host_free_pinned(s.rs);
        s.rs = nullptr;
        s.rs_cap = 0;
        void* p = nullptr;
        int rc = zkb_host_alloc_pinned(kb * 64, &p);
        if (rc != ZKB_OK) {
          g_l2_error = "zkb_l2_batch_prove: pinned host allocation failed";
          return rc;
        }
        s.rs = static_cast<uint8_t*>(p);
        s.rs_cap = kb * 64;
      }
      // host: witness assignment of the sub-batch, one proof per task, straight into the pinned staging buffer
      std::function<void(size_t)> assign_one = [&](size_t k) {
        const size_t i = s.base + k;
        uint8_t* zdst = s.z + k * nv * 32;
        int rc;
        try {
          std::vector<l2::Fr> z;
          rc = l2_assign(c, inputs + i, witnesses + i, &z);
          if (rc == ZKB_OK) {
            // Montgomery limbs as they are (little-endian host: the same 32 bytes the device's 8 x 32-bit form reads):
            // zkb_prove_batch_begin_ex converts on the GPU, the host saves one Montgomery product per variable
            static_assert(sizeof(l2::Fr) == 32, "Fr is four 64-bit limbs");
            if (z.size() != nv) rc = ZKB_ERR_SHAPE;
            else memcpy(zdst, z.data(), nv * 32);
          }
        } catch (const std::bad_alloc&) {
          rc = ZKB_ERR_OOM;
          g_l2_error = "out of host memory";
        } catch (...) {
          rc = ZKB_ERR_INVALID_ARG;
          g_l2_error = "exception in the witness assignment";
        }
        if (rc != ZKB_OK) {
          status[i] = rc;
          note(rc, g_l2_error);           // thread-local message of this worker
          memset(zdst, 0, nv * 32);       // a valid (all-zero) assignment keeps the rest of the sub-batch provable
        }
        zkb_l2_prover_randomness(inputs[i].batch_id, s.rs + 64 * k, s.rs + 64 * k + 32);
      };
      b->pool->parallel_for(s.count, assign_one);
      int rc = zkb_prove_batch_begin_ex(s.ctx, pk, m, s.z, s.rs, s.count, ZKB_BATCH_Z_MONTGOMERY);
      if (rc != ZKB_OK) {
        note(rc, zkb_last_error(s.ctx));
        for (size_t k = 0; k < s.count; ++k)
          if (status[s.base + k] == ZKB_OK) status[s.base + k] = rc;
      } else {
        s.busy = true;
      }
    }
    {
      // drain in submission order: the slot after the last one used holds the oldest sub-batch
      const size_t nsub = (n + kb - 1) / kb;
      for (int k = 0; k < b->nslots; ++k) finish(b->slots[(nsub + size_t(k)) % size_t(b->nslots)]);
    }
    if (status_out)
      for (size_t i = 0; i < n; ++i) status_out[i] = status[i];
    if (first_rc != ZKB_OK) {
      g_l2_error = first_msg;
      return first_rc;  // per-proof codes are in status_out; the other proofs of the batch are valid
    }
    return ZKB_OK;
  }
  ZKB_L2_CATCH_ALL()
}
