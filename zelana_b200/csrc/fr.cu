// Fr-side kernels and drivers: the seven NTTs, the sparse mat-vecs and the pointwise QAP division of ark-groth16's
// LibsnarkReduction::witness_map_from_matrices, entered by the reference at
// core/src/sequencer/settlement/prover.rs:408.
#include "internal.h"
#include "ntt.cuh"

namespace zkb {
namespace {

constexpr int NTT_LRMAX = 8;
constexpr int NTT_FULL_TABLE_LOG = 24;

struct NttTables {
  Fr* mem = nullptr;  // one allocation
  PowTable fwd, inv, coset_pre, coset_inv_post;
  const Fr* ninv = nullptr;
  const Fr* zinv = nullptr;  // (g^n - 1)^-1
};

struct FrState {
  Fr* poseidon = nullptr;     // 64 x 3 round constants then the 3 x 3 MDS matrix of the L2 circuit's Poseidon (Montgomery)
  Fr* mimc_states = nullptr;  // [arity] = permute(arity): the state every hash_arity starts from (Montgomery), arity 0..7
  Fr* wr_fwd = nullptr;  // omega_(2^LRMAX)^e
  Fr* wr_inv = nullptr;
  std::map<int, NttTables> tables;
};

FrState* state(zkb_ctx* ctx) {
  if (!ctx->fr_state) ctx->fr_state = new FrState();
  return static_cast<FrState*>(ctx->fr_state);
}

__device__ __forceinline__ bool fr_is_canonical(const Fr& a) {
  Fr m = Fr::modulus();
  for (int i = 7; i >= 0; i--) {
    if (a.v[i] < m.v[i]) return true;
    if (a.v[i] > m.v[i]) return false;
  }
  return false;
}

__global__ void fr_field_op_kernel(int op, const Fr* a, const Fr* b, size_t n, Fr* out, int* bad) {
  size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fr x = a[i];
  if (!fr_is_canonical(x)) { atomicExch(bad, 1); return; }
  Fr y = Fr::zero();
  if (op <= 2) {
    y = b[i];
    if (!fr_is_canonical(y)) { atomicExch(bad, 1); return; }
  }
  Fr r;
  switch (op) {
    case 0: r = x + y; break;
    case 1: r = x - y; break;
    case 2: r = (x.to_mont() * y.to_mont()).from_mont(); break;
    case 3: r = x.to_mont().inverse().from_mont(); break;
    default: r = x.neg(); break;
  }
  out[i] = r;
}

// canonical bytes -> Montgomery; flags non-canonical input
__global__ void to_mont_kernel(const Fr* in, Fr* out, size_t n, int* bad) {
  size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fr x = in[i];
  if (!fr_is_canonical(x)) { atomicExch(bad, 1); return; }
  out[i] = x.to_mont();
}

// Montgomery limbs -> canonical; flags values >= r (the host's Montgomery form is the same 256-bit R: zkb_prove_batch_begin_ex)
__global__ void from_mont_kernel(const Fr* in, Fr* out, size_t n, int* bad) {
  size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fr x = in[i];
  if (!fr_is_canonical(x)) { atomicExch(bad, 1); return; }
  out[i] = x.from_mont();
}

// out[i] = <row_i, z> for i < nc ; optionally out[nc + j] = z[j] for j < ni ; zero up to n.
// One thread per row; a row of CSR_LONG_ROW terms or more (the 254-term bit-packing rows of ark-r1cs-std's comparison
// gadget, the ~60-term rows of Poseidon's partial rounds) is handed to the whole warp afterwards: 32 lanes stride over it
// and tree-add by shuffles, so that one long row does not hold its 31 neighbours for 254 dependent products.
constexpr uint64_t CSR_LONG_ROW = 48;
__device__ __forceinline__ Fr shfl_down_fr(const Fr& v, int d) {
  Fr r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = __shfl_down_sync(0xffffffffu, v.v[i], d);
  return r;
}
__global__ void csr_matvec_kernel(const uint64_t* __restrict__ row_ptr, const uint32_t* __restrict__ col,
                                  const Fr* __restrict__ coeff, const Fr* __restrict__ z, uint64_t nc, uint64_t ni,
                                  int append_instance, size_t n, Fr* __restrict__ out, size_t z_stride) {
  // blockIdx.y: which assignment of a batch (z_stride elements apart; outputs n apart)
  z += size_t(blockIdx.y) * z_stride;
  out += size_t(blockIdx.y) * n;
  size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  Fr acc = Fr::zero();
  uint64_t lo = 0, hi = 0;
  bool is_long = false;
  if (i < nc) {
    lo = row_ptr[i];
    hi = row_ptr[i + 1];
    is_long = hi - lo >= CSR_LONG_ROW;
    if (!is_long)
      for (uint64_t k = lo; k < hi; k++) acc = acc + coeff[k] * z[col[k]];
  } else if (append_instance && i < nc + ni) {
    acc = z[i - nc];
  }
  unsigned pending = __ballot_sync(0xffffffffu, is_long);   // blockDim is a multiple of 32: whole warps take part
  while (pending) {
    const int src = __ffs(pending) - 1;
    pending &= pending - 1;
    const uint64_t rlo = __shfl_sync(0xffffffffu, lo, src), rhi = __shfl_sync(0xffffffffu, hi, src);
    Fr part = Fr::zero();
    for (uint64_t k = rlo + lane; k < rhi; k += 32) part = part + coeff[k] * z[col[k]];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) part = part + shfl_down_fr(part, d);  // lane 0 ends with the row sum
    Fr total;
#pragma unroll
    for (int w = 0; w < 8; w++) total.v[w] = __shfl_sync(0xffffffffu, part.v[w], 0);
    if (lane == src) acc = total;
  }
  if (i < n) out[i] = acc;
}

// zinv = (g^n - 1)^-1, g = 5: the inverse of the vanishing polynomial on the coset (one thread, once per domain size)
__global__ void qap_zinv_kernel(int logn, Fr* out) {
  if (blockIdx.x || threadIdx.x) return;
  Fr gn = fr_base(FRB_GEN);
  for (int k = 0; k < logn; k++) gn = gn.sqr();
  *out = (gn - Fr::one()).inverse_vartime();
}

// ab[i] = (a[i] * b[i] - c[i]) * zinv ; a is in Montgomery form, b and c canonical, zinv Montgomery -> canonical
__global__ void qap_pointwise_kernel(const Fr* __restrict__ a, const Fr* __restrict__ b, const Fr* __restrict__ c,
                                     const Fr* __restrict__ zinv, size_t n, Fr* __restrict__ out) {
  size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = (load_fr(a + i) * load_fr(b + i) - load_fr(c + i)) * ldg_fr(zinv);
}

// scalars for the folded MSMs (all canonical):
//   za[0..nv-1) = z[1..nv) ; za[nv-1] = 1 ; za[nv] = 1 ; za[nv+1] = r
//   zl[0..nw) = z[ni..nv)  ; zl[nw] = -(r*s) mod r_mod
__global__ void prove_tail_scalars_kernel(const Fr* r, const Fr* s, Fr* za_tail, Fr* zl_tail) {
  if (blockIdx.x || threadIdx.x) return;
  Fr one = Fr::zero();
  one.v[0] = 1;
  za_tail[0] = one;
  za_tail[1] = one;
  za_tail[2] = *r;
  Fr rs = (r->to_mont() * s->to_mont()).from_mont();
  zl_tail[0] = rs.neg();
}

// out[i] = k * in[i] mod r, canonical in and out: (in R) * k / R
__global__ void fr_scale_kernel(const Fr* in, const Fr* k, Fr* out, size_t n) {
  size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  out[i] = in[i].to_mont() * (*k);
}

// ---- trusted setup (ark-groth16 generate_parameters_with_qap; prover/src/bin/keygen.rs:87-91) -----------------------------
// consts (Montgomery): [0] tau, [1] alpha, [2] beta, [3] gamma^-1, [4] delta^-1, [5] zt = tau^n - 1, [6] zt / n, [7] zt / delta
__global__ void setup_consts_kernel(const Fr* in /* canonical: tau, alpha, beta, gamma, delta */, int logn, Fr* c, int* bad) {
  if (blockIdx.x || threadIdx.x) return;
  Fr tau = in[0].to_mont(), gamma = in[3].to_mont(), delta = in[4].to_mont();
  for (int i = 0; i < 5; i++)
    if (!fr_is_canonical(in[i])) { atomicExch(bad, 1); return; }
  Fr tn = tau;
  for (int k = 0; k < logn; k++) tn = tn.sqr();
  Fr zt = tn - Fr::one();
  if (zt.is_zero() || gamma.is_zero() || delta.is_zero()) { atomicExch(bad, 5); return; }  // tau in the domain / zero trapdoor
  Fr ninv = fr_base(FRB_INV2);
  Fr nacc = Fr::one();
  for (int k = 0; k < logn; k++) nacc = nacc * ninv;
  Fr dinv = delta.inverse_vartime();
  c[0] = tau;
  c[1] = in[1].to_mont();
  c[2] = in[2].to_mont();
  c[3] = gamma.inverse_vartime();
  c[4] = dinv;
  c[5] = zt;
  c[6] = zt * nacc;
  c[7] = zt * dinv;
}

// u[i] = L_i(tau) = zt / n * omega^i / (tau - omega^i)     (Montgomery)
__global__ void setup_lagrange_kernel(PowTable omega, const Fr* __restrict__ c, size_t n, Fr* __restrict__ u) {
  size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fr w = i ? pow_lookup(omega, i) : Fr::one();
  Fr d = ldg_fr(c + 0) - w;
  store_fr(u + i, ldg_fr(c + 6) * w * d.inverse());
}

// Column j of the three matrices (CSC): a_j = sum_k A[k][j] u_k (+ u_{nc + j} for an instance column), b_j, c_j.
// Outputs, all CANONICAL (they become fixed-base scalars): a_j, b_j, and (beta a_j + alpha b_j + c_j) / gamma for instance
// columns (gamma_abc) or / delta for witness columns (l).
struct CscDev {
  const uint64_t* col_ptr;
  const uint32_t* row;
  const Fr* coeff;  // Montgomery
};
__device__ __forceinline__ Fr csc_dot(const CscDev& m, size_t j, const Fr* __restrict__ u) {
  Fr acc = Fr::zero();
  for (uint64_t k = m.col_ptr[j]; k < m.col_ptr[j + 1]; k++) acc = acc + ldg_fr(m.coeff + k) * ldg_fr(u + m.row[k]);
  return acc;
}
__global__ void setup_columns_kernel(CscDev A, CscDev B, CscDev C, const Fr* __restrict__ u, const Fr* __restrict__ c, uint64_t nc,
                                     uint64_t ni, size_t nv, Fr* __restrict__ a_out, Fr* __restrict__ b_out, Fr* __restrict__ abc_out) {
  size_t j = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (j >= nv) return;
  Fr a = csc_dot(A, j, u), b = csc_dot(B, j, u), cc = csc_dot(C, j, u);
  if (j < ni) a = a + ldg_fr(u + nc + j);
  Fr abc = (ldg_fr(c + 2) * a + ldg_fr(c + 1) * b + cc) * ldg_fr(c + (j < ni ? 3 : 4));
  store_fr(a_out + j, a.from_mont());
  store_fr(b_out + j, b.from_mont());
  store_fr(abc_out + j, abc.from_mont());
}

// h scalar i = tau^i * zt / delta, i < n - 1 (canonical); also the three constants alpha, beta, delta / gamma for the fixed points
__global__ void setup_h_scalars_kernel(const Fr* __restrict__ c, size_t count, Fr* __restrict__ out) {
  size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= count) return;
  store_fr(out + i, (fr_pow_u64(ldg_fr(c + 0), (unsigned long long)i) * ldg_fr(c + 7)).from_mont());
}

// ---- MiMC-7 over Fr: the hash of the forge stack -----------------------------------------------------------
// forge/circuits/zelana_lib/src/poseidon.nr:15-59 = core/src/sequencer/storage/account_tree.rs:48-90 (BigUint there):
//   round constant c_i = (i+1)^3 + (i+1);  permutation (key 0): 91 rounds x -> (x + c_i)^7;  sponge: state = permute(state + input),
//   starting from 0;  hash_N(v_1..v_N) absorbs the arity N first (poseidon.nr:62-94).
// 4 products per round, 364 per permutation, N + 1 permutations per hash -- and the first one (state 0 + arity) is the same
// for every hash of that arity: computed once per context (mimc_states_kernel), so a hash_2 costs two permutations, not three.
// Values are kept in Montgomery form between rounds.
constexpr int MIMC_ROUNDS = 91;

__device__ __forceinline__ void mimc_load_constants(Fr* rc) {   // rc: MIMC_ROUNDS entries of shared memory
  for (int i = threadIdx.x; i < MIMC_ROUNDS; i += blockDim.x) {
    const uint32_t k = uint32_t(i + 1);
    Fr c = Fr::zero();
    c.v[0] = k * k * k + k;   // <= 91^3 + 91: one limb
    rc[i] = c.to_mont();
  }
  __syncthreads();
}

__device__ __forceinline__ Fr mimc_permute0(Fr x, const Fr* rc) {
#pragma unroll 1
  for (int i = 0; i < MIMC_ROUNDS; i++) {
    Fr t = x + rc[i];
    Fr t2 = t * t;
    Fr t4 = t2 * t2;
    x = t4 * t2 * t;
  }
  return x;   // + key, and the key is 0
}

__device__ __forceinline__ Fr mimc_arity_state(int arity, const Fr* rc) {
  Fr a = Fr::zero();
  a.v[0] = uint32_t(arity);
  return mimc_permute0(a.to_mont(), rc);
}

// states[a] = permute(a) for a < 8: computed once per context (thread a)
__global__ void mimc_states_kernel(Fr* states) {
  __shared__ Fr rc[MIMC_ROUNDS];
  mimc_load_constants(rc);
  if (threadIdx.x < 8) states[threadIdx.x] = mimc_arity_state(int(threadIdx.x), rc);
}

// out[i] = hash_arity(in[i * arity .. (i + 1) * arity)), canonical in and out
__global__ void __launch_bounds__(128)
mimc_hash_kernel(const Fr* __restrict__ in, int arity, size_t n, const Fr* __restrict__ states, Fr* __restrict__ out, int* bad) {
  __shared__ Fr rc[MIMC_ROUNDS];
  mimc_load_constants(rc);
  size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fr state = ldg_fr(states + arity);
  for (int k = 0; k < arity; k++) {
    Fr v = load_fr(in + i * arity + k);
    if (!fr_is_canonical(v)) {
      atomicExch(bad, 1);
      return;
    }
    state = mimc_permute0(state + v.to_mont(), rc);
  }
  store_fr(out + i, state.from_mont());
}

// out[i] = the root reached from leaves[i] along siblings[i * depth ..] with index bits bits[i * depth ..] (1 = the running
// node is the RIGHT child): AccountMerklePath::compute_root (account_tree.rs:222-237) = compute_merkle_root (merkle.nr:29-52)
__global__ void __launch_bounds__(128)
mimc_merkle_root_kernel(const Fr* __restrict__ leaves, const Fr* __restrict__ siblings, const uint8_t* __restrict__ bits, size_t n,
                        int depth, const Fr* __restrict__ states, Fr* __restrict__ out, int* bad) {
  __shared__ Fr rc[MIMC_ROUNDS];
  mimc_load_constants(rc);
  size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Fr h2 = ldg_fr(states + 2);   // the state every hash_2 starts from
  Fr cur = load_fr(leaves + i);
  if (!fr_is_canonical(cur)) {
    atomicExch(bad, 1);
    return;
  }
  cur = cur.to_mont();
  for (int l = 0; l < depth; l++) {
    Fr sib = load_fr(siblings + i * depth + l);
    if (!fr_is_canonical(sib)) {
      atomicExch(bad, 1);
      return;
    }
    sib = sib.to_mont();
    const bool right = bits[i * depth + l] != 0;
    Fr st = mimc_permute0(h2 + (right ? sib : cur), rc);
    cur = mimc_permute0(st + (right ? cur : sib), rc);
  }
  store_fr(out + i, cur.from_mont());
}

// ---- Poseidon of the L2 batch circuit: n independent hashes (prover/src/l2_circuit.rs:315-330,477-490: the account, transfer and
// withdrawal LEAF hashes are independent; the folds over them are chains) ------------------------------------------------------
// ark-crypto-primitives 0.5.0 PoseidonSponge with get_poseidon_config() (l2_circuit.rs:68-83): state 3 = capacity 1 + rate 2,
// 8 full + 56 partial rounds of x^5, round = add constants, S-box (all lanes / lane 0), MDS.  hash(e_1..e_k), k <= 3: fresh
// sponge, absorb (a permutation when the rate is full), permute, squeeze state[1] -- zkb_l2_poseidon_hash on the host.
constexpr int POS_T = 3, POS_FULL = 8, POS_PARTIAL = 56, POS_ROUNDS = 64;

__device__ __forceinline__ void poseidon_permute(Fr st[POS_T], const Fr* __restrict__ ark, const Fr* __restrict__ mds) {
#pragma unroll 1
  for (int r = 0; r < POS_ROUNDS; r++) {
    const bool full = r < POS_FULL / 2 || r >= POS_FULL / 2 + POS_PARTIAL;
#pragma unroll
    for (int i = 0; i < POS_T; i++) st[i] = st[i] + ark[r * POS_T + i];
#pragma unroll
    for (int i = 0; i < POS_T; i++) {
      if (i == 0 || full) {
        Fr x2 = st[i] * st[i];
        st[i] = x2 * x2 * st[i];
      }
    }
    Fr nw[POS_T];
#pragma unroll
    for (int i = 0; i < POS_T; i++) nw[i] = st[0] * mds[i * POS_T] + st[1] * mds[i * POS_T + 1] + st[2] * mds[i * POS_T + 2];
#pragma unroll
    for (int i = 0; i < POS_T; i++) st[i] = nw[i];
  }
}

// in: n x arity canonical elements (values >= r are reduced, as Fr::from_le_bytes_mod_order does); out: n canonical hashes
__global__ void __launch_bounds__(128)
poseidon_hash_kernel(const Fr* __restrict__ in, int arity, size_t n, const Fr* __restrict__ params, Fr* __restrict__ out) {
  __shared__ Fr sp[POS_ROUNDS * POS_T + POS_T * POS_T];
  for (int k = threadIdx.x; k < POS_ROUNDS * POS_T + POS_T * POS_T; k += blockDim.x) sp[k] = params[k];
  __syncthreads();
  size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Fr* ark = sp;
  const Fr* mds = sp + POS_ROUNDS * POS_T;
  Fr st[POS_T] = {Fr::zero(), Fr::zero(), Fr::zero()};
  int start = 0;
  for (int k = 0; k < arity; k++) {
    if (start == 2) {
      poseidon_permute(st, ark, mds);
      start = 0;
    }
    Fr v = load_fr(in + i * arity + k);
    while (!fr_is_canonical(v)) v = v - Fr::modulus();   // < 2^256 < 6 r: mod-order reduction of a 32-byte value
    st[1 + start] = st[1 + start] + v.to_mont();
    start++;
  }
  poseidon_permute(st, ark, mds);
  store_fr(out + i, st[1].from_mont());
}

int ensure_wr(zkb_ctx* ctx, FrState* S) {
  if (S->wr_fwd) return ZKB_OK;
  const uint32_t cnt = 1u << (NTT_LRMAX - 1);
  CUDA_TRY(ctx, cudaMalloc(&S->wr_fwd, cnt * sizeof(Fr)));
  CUDA_TRY(ctx, cudaMalloc(&S->wr_inv, cnt * sizeof(Fr)));
  unsigned long long mult = 1ull << (28 - NTT_LRMAX);
  fr_pow_table_kernel<<<blocks_for(cnt, 64), 64, 0, ctx->stream>>>(S->wr_fwd, cnt, FRB_ROOT, mult, 1, 0, 0);
  fr_pow_table_kernel<<<blocks_for(cnt, 64), 64, 0, ctx->stream>>>(S->wr_inv, cnt, FRB_ROOT_INV, mult, 1, 0, 0);
  ctx->launches += 2;
  CUDA_TRY(ctx, cudaGetLastError());
  return ZKB_OK;
}

int ensure_ntt_tables(zkb_ctx* ctx, FrState* S, int logn, NttTables** out) {
  auto it = S->tables.find(logn);
  if (it != S->tables.end()) {
    *out = &it->second;
    return ZKB_OK;
  }
  NttTables t;
  // Up to 2^24 the power tables are FULL (one load per twiddle, 32 B x n per table: the transform is multiply-bound, so a
  // table read is cheaper than the extra product of a two-level lookup); beyond that, two levels of ~sqrt(n) entries.
  const bool full = logn <= NTT_FULL_TABLE_LOG;
  int lb = full ? logn : (logn + 1) / 2;
  uint32_t nlo = 1u << lb, nhi = 1u << (logn - lb);
  size_t per = size_t(nlo) + nhi;
  CUDA_TRY(ctx, cudaMalloc(&t.mem, (4 * per + 2) * sizeof(Fr)));
  Fr* p = t.mem;
  unsigned long long wmult = 1ull << (28 - logn);
  auto gen = [&](Fr* lo, int base, unsigned long long mult, int sbase, unsigned long long sexp) {
    Fr* hi = lo + nlo;
    // full table: the constant factor is baked into every entry (hi has the single entry lo[0]'s value)
    fr_pow_table_kernel<<<blocks_for(nlo, 64), 64, 0, ctx->stream>>>(lo, nlo, base, mult, 1, full ? sbase : 0, full ? sexp : 0);
    fr_pow_table_kernel<<<blocks_for(nhi, 64), 64, 0, ctx->stream>>>(hi, nhi, base, mult, 1ull << lb, sbase, sexp);
    ctx->launches += 2;
    PowTable pt;
    pt.lo = lo;
    pt.hi = hi;
    pt.lo_bits = lb;
    pt.scaled = (sexp && !full) ? 1 : 0;
    return pt;
  };
  t.fwd = gen(p, FRB_ROOT, wmult, 0, 0);
  t.inv = gen(p + per, FRB_ROOT_INV, wmult, 0, 0);
  t.coset_pre = gen(p + 2 * per, FRB_GEN, 1, 0, 0);
  t.coset_inv_post = gen(p + 3 * per, FRB_GEN_INV, 1, FRB_INV2, (unsigned long long)logn);  // n^-1 g^-k
  Fr* ninv = p + 4 * per;
  // single constant n^-1 = (1/2)^logn: i = 0 gives base^0 = 1, times sbase^sexp
  fr_pow_table_kernel<<<1, 32, 0, ctx->stream>>>(ninv, 1, FRB_INV2, 1, 1, FRB_INV2, (unsigned long long)logn);
  qap_zinv_kernel<<<1, 32, 0, ctx->stream>>>(logn, ninv + 1);
  ctx->launches += 2;
  t.ninv = ninv;
  t.zinv = ninv + 1;
  CUDA_TRY(ctx, cudaGetLastError());
  auto ins = S->tables.emplace(logn, t);
  *out = &ins.first->second;
  return ZKB_OK;
}

template <int LR>
void launch_pass(const NttPassArgs& a, int batch, cudaStream_t st) {
  constexpr int TQ = 4;
  constexpr int NE = (1 << LR) * TQ;
  constexpr int NT = NE / 2 < 32 ? 32 : NE / 2;
  size_t nq = size_t(1) << (a.logn - LR);
  unsigned blocks = unsigned((nq + TQ - 1) / TQ);
  ntt_pass_kernel<LR, TQ><<<dim3(blocks, unsigned(batch)), NT, 0, st>>>(a);
}

}  // namespace

void fr_state_free(zkb_ctx* ctx) {
  if (!ctx->fr_state) return;
  FrState* S = static_cast<FrState*>(ctx->fr_state);
  g_alloc_epoch.fetch_add(1, std::memory_order_relaxed);
  if (S->mimc_states) cudaFree(S->mimc_states);
  if (S->poseidon) cudaFree(S->poseidon);
  if (S->wr_fwd) cudaFree(S->wr_fwd);
  if (S->wr_inv) cudaFree(S->wr_inv);
  for (auto& kv : S->tables) cudaFree(kv.second.mem);
  delete S;
  ctx->fr_state = nullptr;
}

int fr_field_op(zkb_ctx* ctx, int op, const uint8_t* a, const uint8_t* b, size_t n, uint8_t* out) {
  size_t bytes = n * 32;
  CUDA_TRY(ctx, ctx->tmp0.reserve(bytes));
  CUDA_TRY(ctx, ctx->tmp1.reserve(bytes));
  CUDA_TRY(ctx, ctx->tmp2.reserve(bytes));
  ZKB_TRY(clear_flag(ctx));
  CUDA_TRY(ctx, cudaMemcpyAsync(ctx->tmp0.p, a, bytes, cudaMemcpyHostToDevice, ctx->stream));
  if (op <= 2) CUDA_TRY(ctx, cudaMemcpyAsync(ctx->tmp1.p, b, bytes, cudaMemcpyHostToDevice, ctx->stream));
  fr_field_op_kernel<<<blocks_for(n, 128), 128, 0, ctx->stream>>>(op, ctx->tmp0.as<Fr>(), ctx->tmp1.as<Fr>(), n, ctx->tmp2.as<Fr>(), ctx->flag.as<int>());
  ctx->launches++;
  CUDA_TRY(ctx, cudaGetLastError());
  CUDA_TRY(ctx, cudaMemcpyAsync(out, ctx->tmp2.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  return check_flag(ctx, "zkb_field_op");
}

int poseidon_hash_dev(zkb_ctx* ctx, const uint8_t* params_canonical, int arity, const Fr* in, size_t n, Fr* out) {
  FrState* S = state(ctx);
  constexpr size_t NP = POS_ROUNDS * POS_T + POS_T * POS_T;
  if (!S->poseidon) {
    if (!params_canonical) ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "poseidon: parameters not loaded");
    Fr* raw = nullptr;
    CUDA_TRY(ctx, cudaMalloc(&raw, NP * sizeof(Fr)));
    cudaError_t e = cudaMalloc(&S->poseidon, NP * sizeof(Fr));
    if (e != cudaSuccess) {
      cudaFree(raw);
      S->poseidon = nullptr;
      CUDA_TRY(ctx, e);
    }
    ZKB_TRY(clear_flag(ctx));
    cudaMemcpyAsync(raw, params_canonical, NP * sizeof(Fr), cudaMemcpyHostToDevice, ctx->stream);
    int rc = fr_to_mont(ctx, raw, S->poseidon, NP);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(raw);
    if (rc != ZKB_OK) return rc;
  }
  if (!n) return ZKB_OK;
  poseidon_hash_kernel<<<blocks_for(n, 128), 128, 0, ctx->stream>>>(in, arity, n, S->poseidon, out);
  ctx->launches++;
  CUDA_TRY(ctx, cudaGetLastError());
  return ZKB_OK;
}

static int ensure_mimc_states(zkb_ctx* ctx, const Fr** out) {
  FrState* S = state(ctx);
  if (!S->mimc_states) {
    CUDA_TRY(ctx, cudaMalloc(&S->mimc_states, 8 * sizeof(Fr)));
    mimc_states_kernel<<<1, 128, 0, ctx->stream>>>(S->mimc_states);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
  }
  *out = S->mimc_states;
  return ZKB_OK;
}

int mimc_hash_dev(zkb_ctx* ctx, int arity, const Fr* in, size_t n, Fr* out) {
  if (!n) return ZKB_OK;
  const Fr* states = nullptr;
  ZKB_TRY(ensure_mimc_states(ctx, &states));
  mimc_hash_kernel<<<blocks_for(n, 128), 128, 0, ctx->stream>>>(in, arity, n, states, out, ctx->flag.as<int>());
  ctx->launches++;
  CUDA_TRY(ctx, cudaGetLastError());
  return ZKB_OK;
}

int mimc_merkle_roots_dev(zkb_ctx* ctx, const Fr* leaves, const Fr* siblings, const uint8_t* bits, size_t n, int depth, Fr* out) {
  if (!n) return ZKB_OK;
  const Fr* states = nullptr;
  ZKB_TRY(ensure_mimc_states(ctx, &states));
  mimc_merkle_root_kernel<<<blocks_for(n, 128), 128, 0, ctx->stream>>>(leaves, siblings, bits, n, depth, states, out, ctx->flag.as<int>());
  ctx->launches++;
  CUDA_TRY(ctx, cudaGetLastError());
  return ZKB_OK;
}

int fr_to_mont(zkb_ctx* ctx, const Fr* in, Fr* out, size_t n) {
  if (!n) return ZKB_OK;
  to_mont_kernel<<<blocks_for(n, 128), 128, 0, ctx->stream>>>(in, out, n, ctx->flag.as<int>());
  ctx->launches++;
  CUDA_TRY(ctx, cudaGetLastError());
  return ZKB_OK;
}

int fr_from_mont(zkb_ctx* ctx, const Fr* in, Fr* out, size_t n) {
  if (!n) return ZKB_OK;
  from_mont_kernel<<<blocks_for(n, 128), 128, 0, ctx->stream>>>(in, out, n, ctx->flag.as<int>());
  ctx->launches++;
  CUDA_TRY(ctx, cudaGetLastError());
  return ZKB_OK;
}

int prove_tail_scalars(zkb_ctx* ctx, const Fr* r, const Fr* s, Fr* za_tail, Fr* zl_tail) {
  prove_tail_scalars_kernel<<<1, 32, 0, ctx->stream>>>(r, s, za_tail, zl_tail);
  ctx->launches++;
  CUDA_TRY(ctx, cudaGetLastError());
  return ZKB_OK;
}

int fr_scale(zkb_ctx* ctx, const Fr* in, const Fr* k_dev, Fr* out, size_t n) {
  if (!n) return ZKB_OK;
  fr_scale_kernel<<<blocks_for(n, 128), 128, 0, ctx->stream>>>(in, k_dev, out, n);
  ctx->launches++;
  CUDA_TRY(ctx, cudaGetLastError());
  return ZKB_OK;
}

int ntt_dev_impl(zkb_ctx* ctx, const Fr* in, Fr* out, int logn, int inverse, int coset) {
  return ntt_batch_dev_impl(ctx, in, out, logn, inverse, coset, 1);
}

// `batch` transforms of 2^logn elements each, back to back in `in` and `out`: one launch per pass for all of them
// (blockIdx.y = polynomial).  A batch of small proofs runs its 7 x K transforms of 2^13 this way.
int ntt_batch_dev_impl(zkb_ctx* ctx, const Fr* in, Fr* out, int logn, int inverse, int coset, int batch) {
  if (logn < 0 || logn > 28) ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "ntt: log_n %d outside [0, 28]", logn);
  if (batch < 1 || batch > 65535) ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "ntt: batch %d outside [1, 65535]", batch);
  size_t n = size_t(1) << logn;
  if (logn == 0) {
    // size-1 domain: identity (coset scale g^0 = 1, 1/n = 1)
    if (in != out) CUDA_TRY(ctx, cudaMemcpyAsync(out, in, sizeof(Fr) * batch, cudaMemcpyDeviceToDevice, ctx->stream));
    return ZKB_OK;
  }
  FrState* S = state(ctx);
  ZKB_TRY(ensure_wr(ctx, S));
  NttTables* T = nullptr;
  ZKB_TRY(ensure_ntt_tables(ctx, S, logn, &T));
  int npass = (logn + NTT_LRMAX - 1) / NTT_LRMAX;
  int base = logn / npass, extra = logn % npass;
  Fr* scratch[2] = {nullptr, nullptr};
  if (npass >= 2 || in == out) {
    CUDA_TRY(ctx, ctx->tmp0.reserve(n * sizeof(Fr) * batch));
    scratch[0] = ctx->tmp0.as<Fr>();
  }
  if (npass >= 3) {
    CUDA_TRY(ctx, ctx->tmp1.reserve(n * sizeof(Fr) * batch));
    scratch[1] = ctx->tmp1.as<Fr>();
  }
  const Fr* src = in;
  int logm = logn;
  PowTable ninv_tab;
  ninv_tab.lo = T->ninv;
  ninv_tab.hi = T->ninv;
  ninv_tab.lo_bits = -1;  // constant
  ProfScope ps(ctx, PH_NTT);
  for (int p = 0; p < npass; p++) {
    int lr = base + (p < extra ? 1 : 0);
    bool last = (p == npass - 1);
    Fr* dst;
    if (last) {
      dst = (npass == 1 && in == out) ? scratch[0] : out;
    } else {
      dst = scratch[p & 1];
    }
    NttPassArgs a;
    a.in = src;
    a.out = dst;
    a.logn = logn;
    a.logm = logm;
    a.wr = inverse ? S->wr_inv : S->wr_fwd;
    a.lrmax = NTT_LRMAX;
    a.tw = inverse ? T->inv : T->fwd;
    a.has_pre = (p == 0 && !inverse && coset) ? 1 : 0;
    a.pre = T->coset_pre;
    a.has_post = (last && inverse) ? 1 : 0;
    a.post = coset ? T->coset_inv_post : ninv_tab;
    switch (lr) {
      case 1: launch_pass<1>(a, batch, ctx->stream); break;
      case 2: launch_pass<2>(a, batch, ctx->stream); break;
      case 3: launch_pass<3>(a, batch, ctx->stream); break;
      case 4: launch_pass<4>(a, batch, ctx->stream); break;
      case 5: launch_pass<5>(a, batch, ctx->stream); break;
      case 6: launch_pass<6>(a, batch, ctx->stream); break;
      case 7: launch_pass<7>(a, batch, ctx->stream); break;
      default: launch_pass<8>(a, batch, ctx->stream); break;
    }
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    src = dst;
    logm -= lr;
  }
  if (npass == 1 && in == out) CUDA_TRY(ctx, cudaMemcpyAsync(out, scratch[0], n * sizeof(Fr) * batch, cudaMemcpyDeviceToDevice, ctx->stream));
  return ZKB_OK;
}

// Setup scalars on the device.  in5: canonical tau, alpha, beta, gamma, delta (device).  Outputs (device, canonical):
// a_out, b_out, abc_out (nv each), h_out (n - 1).  The CSC arrays are device pointers with Montgomery coefficients.
int setup_scalars_dev(zkb_ctx* ctx, const uint64_t* const col_ptr[3], const uint32_t* const row[3], const Fr* const coeff[3],
                      uint64_t nc, uint64_t ni, uint64_t nw, int logn, const Fr* in5, Fr* consts, Fr* u, Fr* a_out, Fr* b_out,
                      Fr* abc_out, Fr* h_out) {
  const size_t n = size_t(1) << logn, nv = ni + nw;
  FrState* S = state(ctx);
  NttTables* T = nullptr;
  ZKB_TRY(ensure_ntt_tables(ctx, S, logn, &T));
  ZKB_TRY(clear_flag(ctx));
  cudaStream_t st = ctx->stream;
  setup_consts_kernel<<<1, 32, 0, st>>>(in5, logn, consts, ctx->flag.as<int>());
  setup_lagrange_kernel<<<blocks_for(n, 128), 128, 0, st>>>(T->fwd, consts, n, u);
  CscDev A{col_ptr[0], row[0], coeff[0]}, B{col_ptr[1], row[1], coeff[1]}, C{col_ptr[2], row[2], coeff[2]};
  setup_columns_kernel<<<blocks_for(nv, 128), 128, 0, st>>>(A, B, C, u, consts, nc, ni, nv, a_out, b_out, abc_out);
  if (n > 1) setup_h_scalars_kernel<<<blocks_for(n - 1, 128), 128, 0, st>>>(consts, n - 1, h_out);
  ctx->launches += 4;
  CUDA_TRY(ctx, cudaGetLastError());
  return ZKB_OK;
}

// h (canonical, device, domain_size elements) <- witness_map_from_matrices(m, z canonical in w.z)
int witness_map_dev(zkb_ctx* ctx, const CsrDev& A, const CsrDev& B, const CsrDev& C, uint64_t nc, uint64_t ni, uint64_t nw,
                    int lg, const WitnessBufs& w, Fr* h_out) {
  const size_t n = size_t(1) << lg;
  const size_t nv = ni + nw;
  cudaStream_t st = ctx->stream;
  // a-chain in Montgomery form (so that a*b of a Montgomery and a canonical value is canonical)
  ZKB_TRY(clear_flag(ctx));
  {
    ProfScope ps(ctx, PH_MATVEC);
    ZKB_TRY(fr_to_mont(ctx, w.z, w.zm, nv));
    csr_matvec_kernel<<<blocks_for(n, 128), 128, 0, st>>>(A.row_ptr, A.col, A.coeff, w.zm, nc, ni, 1, n, w.wa, 0);
    csr_matvec_kernel<<<blocks_for(n, 128), 128, 0, st>>>(B.row_ptr, B.col, B.coeff, w.z, nc, ni, 0, n, w.wb, 0);
    csr_matvec_kernel<<<blocks_for(n, 128), 128, 0, st>>>(C.row_ptr, C.col, C.coeff, w.z, nc, ni, 0, n, w.wc, 0);
    ctx->launches += 3;
    CUDA_TRY(ctx, cudaGetLastError());
  }
  Fr* chains[3] = {w.wa, w.wb, w.wc};
  for (Fr* v : chains) {
    ZKB_TRY(ntt_dev_impl(ctx, v, v, lg, 1, 0));  // domain.ifft_in_place
    ZKB_TRY(ntt_dev_impl(ctx, v, v, lg, 0, 1));  // coset_domain.fft_in_place
  }
  {
    ProfScope ps(ctx, PH_POINTWISE);
    NttTables* T = nullptr;
    ZKB_TRY(ensure_ntt_tables(ctx, state(ctx), lg, &T));
    qap_pointwise_kernel<<<blocks_for(n, 128), 128, 0, st>>>(w.wa, w.wb, w.wc, T->zinv, n, w.wa);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
  }
  ZKB_TRY(ntt_dev_impl(ctx, w.wa, h_out, lg, 1, 1));  // coset_domain.ifft_in_place
  return ZKB_OK;
}

// K witness maps at once (a batch of proofs of ONE circuit).  z: K x nv canonical; w3: 3 K n elements of scratch laid out
// [a-chain of every proof | b-chains | c-chains] so that each NTT stage is ONE batched launch over 3 K polynomials;
// zm: K x nv scratch; h_out: K x n canonical coefficients of the K quotients.
int witness_map_batch_dev(zkb_ctx* ctx, const CsrDev& A, const CsrDev& B, const CsrDev& C, uint64_t nc, uint64_t ni, uint64_t nw,
                          int lg, int K, const Fr* z, Fr* zm, bool zm_ready, Fr* w3, Fr* h_out) {
  const size_t n = size_t(1) << lg;
  const size_t nv = ni + nw;
  cudaStream_t st = ctx->stream;
  Fr* wa = w3;
  Fr* wb = w3 + size_t(K) * n;
  Fr* wc = w3 + 2 * size_t(K) * n;
  {
    ProfScope ps(ctx, PH_MATVEC);
    if (!zm_ready) ZKB_TRY(fr_to_mont(ctx, z, zm, nv * size_t(K)));
    dim3 grid(blocks_for(n, 128), unsigned(K));
    csr_matvec_kernel<<<grid, 128, 0, st>>>(A.row_ptr, A.col, A.coeff, zm, nc, ni, 1, n, wa, nv);
    csr_matvec_kernel<<<grid, 128, 0, st>>>(B.row_ptr, B.col, B.coeff, z, nc, ni, 0, n, wb, nv);
    csr_matvec_kernel<<<grid, 128, 0, st>>>(C.row_ptr, C.col, C.coeff, z, nc, ni, 0, n, wc, nv);
    ctx->launches += 3;
    CUDA_TRY(ctx, cudaGetLastError());
  }
  ZKB_TRY(ntt_batch_dev_impl(ctx, w3, w3, lg, 1, 0, 3 * K));  // domain.ifft_in_place on a, b, c of every proof
  ZKB_TRY(ntt_batch_dev_impl(ctx, w3, w3, lg, 0, 1, 3 * K));  // coset_domain.fft_in_place
  {
    ProfScope ps(ctx, PH_POINTWISE);
    NttTables* T = nullptr;
    ZKB_TRY(ensure_ntt_tables(ctx, state(ctx), lg, &T));
    const size_t tot = n * size_t(K);   // the K a-chains, b-chains, c-chains are each contiguous: one flat pointwise pass
    qap_pointwise_kernel<<<blocks_for(tot, 128), 128, 0, st>>>(wa, wb, wc, T->zinv, tot, wa);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
  }
  ZKB_TRY(ntt_batch_dev_impl(ctx, wa, h_out, lg, 1, 1, K));  // coset_domain.ifft_in_place
  return ZKB_OK;
}

}  // namespace zkb
