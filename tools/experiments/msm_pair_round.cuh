// EXPERIMENT (not compiled into the library): batched-affine pairwise pre-reduction of the sorted MSM entry list.
// Measured on B200 at 2^24 points (profiles/r01_msm_pair_rounds_rejected.txt): bit-exact, but SLOWER than folding the
// entries directly with XYZZ mixed additions -- the two passes of Montgomery's trick gather every 64-byte point twice
// and each random gather costs a 128-byte DRAM fetch, so round 0 moves 72 GB (24.5 ms, DRAM-bound) to save ~10 ms of
// multiplications.  Kept for the record; see DESIGN.md section 5.
// ------------------------------------------------------------------------------------------- 2b. pairwise pre-reduction
// Batched AFFINE additions.  After the sort, entries 2s and 2s+1 ("slot" s) almost always fall in the same bucket (runs are
// tens to hundreds long), and P + Q in affine coordinates costs 1 inversion + 2M + 1S.  A thread takes M consecutive slots,
// shares ONE inversion among them with Montgomery's trick (prefix products parked in a scratch array, laid out so that a
// warp's accesses coalesce) and writes the half-as-long, still sorted list of sums: ~6.75 field products per addition
// instead of the 10 of an XYZZ mixed addition.  Two such rounds quarter the list that msm_accumulate_kernel then folds.
// Slots whose two entries differ in key are copied through; offsets come from an exclusive scan of the per-slot output counts.
static __global__ void msm_pair_count_kernel(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ d_n, uint32_t n_host,
                                             uint32_t sentinel, size_t nslots_bound, uint32_t* __restrict__ cnt) {
  size_t s = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (s >= nslots_bound) return;
  const uint32_t n = d_n ? *d_n : n_host;
  uint32_t k0 = 2 * s < n ? keys[2 * s] : sentinel;
  uint32_t k1 = 2 * s + 1 < n ? keys[2 * s + 1] : sentinel;
  uint32_t c = (k0 < sentinel ? 1u : 0u) + (k1 < sentinel ? 1u : 0u);
  if (k0 < sentinel && k0 == k1) c = 1;
  cnt[s] = c;
}

static __global__ void msm_pair_total_kernel(const uint32_t* __restrict__ cnt, const uint32_t* __restrict__ off, size_t nslots_bound,
                                             uint32_t* __restrict__ d_n_out) {
  if (blockIdx.x || threadIdx.x) return;
  *d_n_out = off[nslots_bound - 1] + cnt[nslots_bound - 1];
}

template <class F, bool FROM_TABLE>
struct PairSrc {
  const Affine<F>* table;  // FROM_TABLE: window tables, indexed through vals ; else: the previous round's points
  const uint32_t* vals;
  const uint32_t* keys;
  uint32_t n, sentinel;
  __device__ __forceinline__ uint32_t key(size_t j) const { return j < n ? keys[j] : sentinel; }
  __device__ __forceinline__ Affine<F> point(size_t j) const {
    if (FROM_TABLE) {
      uint32_t v = vals[j];
      Affine<F> p = load_affine(table + (v & 0x7fffffffu));
      if (v >> 31) p.y = p.y.neg();
      return p;
    }
    return load_affine(table + j);
  }
};

template <class F>
__device__ __forceinline__ F ld_field(const F* p) {
  F r;
  const uint4* src = reinterpret_cast<const uint4*>(p);
  uint4* dst = reinterpret_cast<uint4*>(&r);
#pragma unroll
  for (int k = 0; k < int(sizeof(F) / 16); k++) dst[k] = src[k];
  return r;
}
template <class F>
__device__ __forceinline__ void st_field(F* p, const F& v) {
  const uint4* src = reinterpret_cast<const uint4*>(&v);
  uint4* dst = reinterpret_cast<uint4*>(p);
#pragma unroll
  for (int k = 0; k < int(sizeof(F) / 16); k++) dst[k] = src[k];
}

// denominator of P + Q (both finite): x_Q - x_P, or 2 y_P for a doubling; returns false when P = -Q (sum is infinity)
template <class F>
__device__ __forceinline__ bool pair_denominator(const Affine<F>& P, const Affine<F>& Q, F& den, bool& dbl) {
  den = Q.x - P.x;
  dbl = false;
  if (den.is_zero()) {
    if (P.y == Q.y && !P.y.is_zero()) {
      den = P.y.dbl();
      dbl = true;
    } else {
      return false;
    }
  }
  return true;
}

template <class F, bool FROM_TABLE, int THREADS>
__global__ void __launch_bounds__(THREADS)
msm_pair_round_kernel(const Affine<F>* __restrict__ table, const uint32_t* __restrict__ vals, const uint32_t* __restrict__ keys,
                      const uint32_t* __restrict__ d_n, uint32_t n_host, uint32_t sentinel, const uint32_t* __restrict__ off,
                      int M, F* __restrict__ scratch, Affine<F>* __restrict__ out_pts, uint32_t* __restrict__ out_keys) {
  const size_t T = size_t(gridDim.x) * THREADS;
  const size_t t = size_t(blockIdx.x) * THREADS + threadIdx.x;
  PairSrc<F, FROM_TABLE> src{table, vals, keys, d_n ? *d_n : n_host, sentinel};
  const size_t nslots = (size_t(src.n) + 1) / 2;
  const size_t s0 = t * size_t(M);
  if (s0 >= nslots) return;
  const size_t s1 = s0 + M < nslots ? s0 + M : nslots;

  // pass 1: running product of the denominators, prefix parked in scratch[(s - s0) * T + t]
  F pr = F::one();
  for (size_t s = s0; s < s1; s++) {
    uint32_t k0 = src.key(2 * s), k1 = src.key(2 * s + 1);
    if (k0 < sentinel && k1 == k0) {
      Affine<F> P = src.point(2 * s), Q = src.point(2 * s + 1);
      if (!P.is_inf() && !Q.is_inf()) {
        F den;
        bool dbl;
        if (pair_denominator(P, Q, den, dbl)) {
          st_field(scratch + (s - s0) * T + t, pr);
          pr = pr * den;
        }
      }
    }
  }
  F inv = pr.inverse();
  // pass 2, backwards: peel the inverses off and finish the additions
  for (size_t s = s1; s-- > s0;) {
    uint32_t k0 = src.key(2 * s), k1 = src.key(2 * s + 1);
    if (k0 >= sentinel) continue;
    const uint32_t o = off[s];
    Affine<F> P = src.point(2 * s);
    if (k1 != k0) {
      store_affine(out_pts + o, P);
      out_keys[o] = k0;
      if (k1 < sentinel) {
        store_affine(out_pts + o + 1, src.point(2 * s + 1));
        out_keys[o + 1] = k1;
      }
      continue;
    }
    Affine<F> Q = src.point(2 * s + 1);
    Affine<F> Rr;
    if (P.is_inf()) {
      Rr = Q;
    } else if (Q.is_inf()) {
      Rr = P;
    } else {
      F den;
      bool dbl;
      if (!pair_denominator(P, Q, den, dbl)) {
        Rr = Affine<F>::inf();
      } else {
        F dinv = inv * ld_field(scratch + (s - s0) * T + t);
        inv = inv * den;
        F num;
        if (dbl) {
          F xx = P.x.sqr();
          num = xx.dbl() + xx;
        } else {
          num = Q.y - P.y;
        }
        F lam = num * dinv;
        Rr.x = lam.sqr() - P.x - Q.x;
        Rr.y = lam * (P.x - Rr.x) - P.y;
      }
    }
    store_affine(out_pts + o, Rr);
    out_keys[o] = k0;
  }
}

