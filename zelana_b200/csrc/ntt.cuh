// Radix-2 NTT / iNTT / coset-NTT / coset-iNTT over BN254 Fr for sm_100a.
//
// Replaces ark-poly 0.5.0 `Radix2EvaluationDomain::{fft,ifft}_in_place` and the coset variants
// (`domain.get_coset(Fr::GENERATOR)`), the seven transforms inside ark-groth16's
// `LibsnarkReduction::witness_map_from_matrices` that the reference reaches from
// core/src/sequencer/settlement/prover.rs:408.  Natural order in, natural order out; omega_n =
// ROOT_2^28 ^ (2^(28 - log n)), coset generator g = 5 -- arkworks' constants.
//
// Schedule: a Stockham (autosort) decomposition into ceil(log n / 8) passes.  A pass takes the current
// sub-transform length M with inner batch B (n = M*B, layout [j][b]), splits j = j_top*(M/R) + j_r,
// does the R-point DFT over j_top inside a thread block (shared memory, radix-2 DIF stages), multiplies
// by omega_M^(j_r*k) and writes layout [j_r][k][b].  Every pass reads and writes each element exactly
// once (64 B of HBM traffic per element per pass), 128 B-contiguous on both sides.  Because
// the transform is linear and every constant (twiddle, scale) is kept in Montgomery form, data in
// canonical form stays canonical and data in Montgomery form stays Montgomery: no conversion passes.
//
// Scaling (g^k before a forward coset transform; n^-1 or n^-1 g^-k after an inverse one) is fused into
// the first pass's loads / the last pass's stores through two-level power tables.
#pragma once
#include <cuda_runtime.h>

#include "fp.cuh"

namespace zkb {

// Montgomery-form constants of ark-bn254 Fr (computed with oracle/bn254.py)
__device__ __constant__ const uint32_t FR_ROOT_W[8] = {0x80d13d9cu, 0x636e7355u, 0x2445ffd6u, 0xa22bf374u,
                                                       0x1eb203d8u, 0x56452ac0u, 0x2963f9e7u, 0x1860ef94u};
__device__ __constant__ const uint32_t FR_ROOT_INV_W[8] = {0x584bb683u, 0x89bcc016u, 0x0164a50cu, 0xe8d9887fu,
                                                           0x795eda3du, 0x755e95cbu, 0x1323b130u, 0x0f572b87u};
__device__ __constant__ const uint32_t FR_GEN_W[8] = {0x9fffffe6u, 0x1b0d0ef9u, 0xa32a913fu, 0xeaba68a3u,
                                                      0xd8dd0689u, 0x47d8eb76u, 0x20f5bbc3u, 0x15d00855u};
__device__ __constant__ const uint32_t FR_GEN_INV_W[8] = {0x09999999u, 0xd7453974u, 0x83c3efa8u, 0xb4ada7d4u,
                                                          0xe57f3161u, 0xc49ca2f8u, 0xac156cb3u, 0x162a3754u};
__device__ __constant__ const uint32_t FR_INV2_W[8] = {0x1ffffffeu, 0x783c14d8u, 0x0c8d1eddu, 0xaf982f6fu,
                                                       0xfcfd4f45u, 0x8f5f7492u, 0x3d9cbfacu, 0x1f37631au};

enum FrBase { FRB_ROOT = 0, FRB_ROOT_INV = 1, FRB_GEN = 2, FRB_GEN_INV = 3, FRB_INV2 = 4 };

__device__ __forceinline__ Fr fr_base(int which) {
  const uint32_t* w = which == FRB_ROOT ? FR_ROOT_W
                      : which == FRB_ROOT_INV ? FR_ROOT_INV_W
                      : which == FRB_GEN ? FR_GEN_W
                      : which == FRB_GEN_INV ? FR_GEN_INV_W
                                             : FR_INV2_W;
  Fr r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = w[i];
  return r;
}

__device__ inline Fr fr_pow_u64(Fr b, unsigned long long e) {
  Fr r = Fr::one();
  while (e) {
    if (e & 1ull) r = r * b;
    b = b.sqr();
    e >>= 1;
  }
  return r;
}

// out[i] = base^(mult * i * stride) * sbase^sexp      (all Montgomery)
__global__ void fr_pow_table_kernel(Fr* out, uint32_t count, int base, unsigned long long mult,
                                    unsigned long long stride, int sbase, unsigned long long sexp) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  Fr b = fr_pow_u64(fr_base(base), mult);
  Fr v = fr_pow_u64(b, (unsigned long long)i * stride);
  if (sexp) v = v * fr_pow_u64(fr_base(sbase), sexp);
  out[i] = v;
}

// Two-level table of powers: value(e) = hi[e >> lo_bits] * lo[e & (2^lo_bits - 1)]
struct PowTable {
  const Fr* lo = nullptr;
  const Fr* hi = nullptr;
  int lo_bits = 0;   // < 0: constant table, value = lo[0]
  int scaled = 0;    // hi[] carries an extra constant factor: hi[0] != 1
};

__device__ __forceinline__ Fr load_fr(const Fr* p) {
  Fr r;
  const uint4* s = reinterpret_cast<const uint4*>(p);
  uint4 a = s[0], b = s[1];
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
__device__ __forceinline__ Fr ldg_fr(const Fr* p) {
  Fr r;
  const uint4* s = reinterpret_cast<const uint4*>(p);
  uint4 a = __ldg(s), b = __ldg(s + 1);
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
__device__ __forceinline__ void store_fr(Fr* p, const Fr& r) {
  uint4* d = reinterpret_cast<uint4*>(p);
  d[0] = make_uint4(r.v[0], r.v[1], r.v[2], r.v[3]);
  d[1] = make_uint4(r.v[4], r.v[5], r.v[6], r.v[7]);
}

__device__ __forceinline__ Fr pow_lookup(const PowTable& t, size_t e) {
  if (t.lo_bits < 0) return ldg_fr(t.lo);
  size_t h = e >> t.lo_bits, l = e & ((size_t(1) << t.lo_bits) - 1);
  if (l == 0) return ldg_fr(t.hi + h);  // lo[0] = 1
  if (h == 0 && !t.scaled) return ldg_fr(t.lo + l);
  return ldg_fr(t.hi + h) * ldg_fr(t.lo + l);
}

struct NttPassArgs {
  const Fr* in;
  Fr* out;
  int logn;      // n = 2^logn
  int logm;      // current sub-transform length M = 2^logm (B = n / M)
  const Fr* wr;  // omega_(2^LRMAX)^e, e < 2^(LRMAX-1), direction-specific
  int lrmax;
  PowTable tw;   // powers of omega_n (direction-specific)
  PowTable pre;  // optional scale of input element j by pre(j)   (first pass only)
  PowTable post; // optional scale of output element k by post(k) (last pass only)
  int has_pre, has_post;
};

__device__ __forceinline__ uint32_t bitrev(uint32_t x, int bits) { return __brev(x) >> (32 - bits); }

// One Stockham pass of radix R = 2^LR on TQ adjacent columns per block.
template <int LR, int TQ>
__global__ void __launch_bounds__((1 << LR) * TQ / 2 < 32 ? 32 : (1 << LR) * TQ / 2)
ntt_pass_kernel(NttPassArgs a) {
  constexpr int R = 1 << LR;
  constexpr int NE = R * TQ;                       // elements per block
  constexpr int NT = NE / 2 < 32 ? 32 : NE / 2;    // threads per block
  __shared__ uint4 p0[NE], p1[NE];                 // SoA halves: conflict-free 16 B accesses

  // blockIdx.y: which polynomial of a batch (each n elements long, back to back in both buffers)
  a.in += size_t(blockIdx.y) << a.logn;
  a.out += size_t(blockIdx.y) << a.logn;
  const size_t nq = size_t(1) << (a.logn - LR);    // columns = n / R
  const size_t q0 = size_t(blockIdx.x) * TQ;
  const int logB = a.logn - a.logm;
  const int tid = threadIdx.x;

  // ---- load [j_top][q] (q fastest: TQ * 32 B contiguous), optional pre-scale by pre(j)
  for (int e = tid; e < NE; e += NT) {
    int jt = e / TQ, qq = e % TQ;
    size_t q = q0 + qq;
    Fr x = Fr::zero();
    if (q < nq) {
      size_t j = (size_t(jt) << (a.logn - LR)) + q;
      x = load_fr(a.in + j);
      if (a.has_pre) x = x * pow_lookup(a.pre, j);
    }
    p0[e] = make_uint4(x.v[0], x.v[1], x.v[2], x.v[3]);
    p1[e] = make_uint4(x.v[4], x.v[5], x.v[6], x.v[7]);
  }
  __syncthreads();

  // ---- R-point DIF over j_top: stage s pairs rows (j, j + half)
#pragma unroll 1
  for (int s = 0; s < LR; s++) {
    const int half = R >> (s + 1);
    for (int u = tid; u < NE / 2; u += NT) {
      int qq = u % TQ, t = u / TQ;
      int grp = t / half, pos = t % half;
      int i0 = (grp * 2 * half + pos) * TQ + qq;
      int i1 = i0 + half * TQ;
      uint4 a0 = p0[i0], a1 = p1[i0], b0 = p0[i1], b1 = p1[i1];
      Fr x, y;
      x.v[0] = a0.x; x.v[1] = a0.y; x.v[2] = a0.z; x.v[3] = a0.w;
      x.v[4] = a1.x; x.v[5] = a1.y; x.v[6] = a1.z; x.v[7] = a1.w;
      y.v[0] = b0.x; y.v[1] = b0.y; y.v[2] = b0.z; y.v[3] = b0.w;
      y.v[4] = b1.x; y.v[5] = b1.y; y.v[6] = b1.z; y.v[7] = b1.w;
      Fr sum = x + y, dif = x - y;
      if (pos != 0) {  // omega_R^(pos << s) = wr[(pos << s) << (lrmax - LR)]
        dif = dif * ldg_fr(a.wr + ((size_t(pos) << s) << (a.lrmax - LR)));
      }
      p0[i0] = make_uint4(sum.v[0], sum.v[1], sum.v[2], sum.v[3]);
      p1[i0] = make_uint4(sum.v[4], sum.v[5], sum.v[6], sum.v[7]);
      p0[i1] = make_uint4(dif.v[0], dif.v[1], dif.v[2], dif.v[3]);
      p1[i1] = make_uint4(dif.v[4], dif.v[5], dif.v[6], dif.v[7]);
    }
    __syncthreads();
  }

  // ---- store layout [j_r][k][b]; the DIF left frequency k at row bitrev(k)
  const bool first_layout = (logB == 0);  // B = 1: address = q * R + k, make k the fastest index
  for (int e = tid; e < NE; e += NT) {
    int k, qq;
    if (first_layout) {
      k = e % R;
      qq = e / R;
    } else {
      qq = e % TQ;
      k = e / TQ;
    }
    size_t q = q0 + qq;
    if (q >= nq) continue;
    int src = int(bitrev(uint32_t(k), LR)) * TQ + qq;
    uint4 a0 = p0[src], a1 = p1[src];
    Fr y;
    y.v[0] = a0.x; y.v[1] = a0.y; y.v[2] = a0.z; y.v[3] = a0.w;
    y.v[4] = a1.x; y.v[5] = a1.y; y.v[6] = a1.z; y.v[7] = a1.w;
    size_t jr = q >> logB, b = q & ((size_t(1) << logB) - 1);
    if (a.logm > LR && jr != 0 && k != 0) {
      // omega_M^(jr*k) = omega_n^(B*jr*k)
      y = y * pow_lookup(a.tw, (jr * size_t(k)) << logB);
    }
    size_t o = (((jr << LR) + size_t(k)) << logB) + b;
    if (a.has_post) y = y * pow_lookup(a.post, o);
    store_fr(a.out + o, y);
  }
}

}  // namespace zkb
