// Pippenger variable-base MSM for BN254 G1 / G2 on sm_100a.
//
// Replaces ark-ec 0.5.0 `VariableBaseMSM::msm_bigint` (signed-digit bucket method) at the five call
// sites of ark-groth16's create_proof_with_assignment (h_query, l_query, a_query, b_g1_query in G1 and
// b_g2_query in G2), which the reference enters from core/src/sequencer/settlement/prover.rs:408.
//
// Pipeline (all on one stream, no host synchronisation):
//   1. msm_digits_kernel     canonical 254-bit scalars -> signed c-bit digits; one (key, value) entry
//                            per (point, window): key = window * 2^(c-1) + |digit| - 1, value =
//                            point index | sign << 31; zero digits get the sentinel key (sorts last).
//   2. cub::DeviceRadixSort  entries by key (bucket id).  Library sort, HBM-bound, a few % of the run.
//   3. msm_accumulate_kernel the hot loop.  The sorted entry list is cut into equal chunks, one per
//                            thread, so load balance does not depend on the scalar distribution.  A
//                            thread sums the runs in its chunk with XYZZ mixed additions (8M+2S) on
//                            gathered 64-byte affine bases; a run that starts inside the chunk is
//                            owned by the thread and stored to its bucket, the run that was already
//                            open at the chunk start goes to a per-thread "head" slot.
//   4. msm_heads_kernel      folds head partial sums into their buckets (one leader per bucket id).
//   5. msm_reduce_kernel     per window, per segment of buckets: running-sum sum_b (b+1) B_b.
//   6. msm_window_sum_kernel tree-sum of segment results per window.
//   7. msm_final_kernel      Horner over windows (c doublings each), XYZZ -> affine -> canonical bytes.
#pragma once
#include <cub/device/device_radix_sort.cuh>
#include <cuda_runtime.h>

#include "internal.h"

namespace zkb {

struct MsmPlan {
  int c;            // window bits
  int nwin;         // number of windows
  uint32_t nbuck;   // buckets per window = 2^(c-1)
  uint32_t sentinel;  // key of a zero digit
  int key_bits;
  int chunk;        // sorted entries per accumulate thread
  int seg;          // buckets per reduce thread
};

static inline int ilog2_ceil(size_t n) {
  int l = 0;
  while ((size_t(1) << l) < n) l++;
  return l;
}

static inline MsmPlan msm_make_plan(size_t n, int c_override, int sm_count, int threads_per_sm) {
  MsmPlan p;
  int lg = ilog2_ceil(n < 2 ? 2 : n);
  int c = lg - 3;
  if (c < 6) c = 6;
  if (c > 16) c = 16;
  if (c_override > 0) c = c_override;
  p.c = c;
  p.nwin = (255 + c - 1) / c;
  p.nbuck = 1u << (c - 1);
  p.sentinel = uint32_t(p.nwin) * p.nbuck;
  p.key_bits = ilog2_ceil(size_t(p.sentinel) + 1);
  // aim for ~8 waves of accumulate threads, chunks of at least 32 and at most 1024 entries
  size_t total = size_t(p.nwin) * n;
  size_t resident = size_t(sm_count) * threads_per_sm;
  size_t chunk = (total + resident * 8 - 1) / (resident * 8);
  if (chunk < 32) chunk = 32;
  if (chunk > 1024) chunk = 1024;
  p.chunk = int((chunk + 3) & ~size_t(3));
  p.seg = p.nbuck >= 4096 ? 16 : (p.nbuck >= 256 ? 8 : 4);
  if (uint32_t(p.seg) > p.nbuck) p.seg = int(p.nbuck);
  return p;
}

// ------------------------------------------------------------------------------------------- 1. digits
// scalars: n x 32 bytes canonical little-endian (NOT Montgomery): what msm_bigint receives.
static __global__ void msm_digits_kernel(const uint32_t* __restrict__ scalars, size_t n, int c, int nwin, uint32_t nbuck,
                                  uint32_t sentinel, uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint4* sp = reinterpret_cast<const uint4*>(scalars + i * 8);
  uint4 lo = sp[0], hi = sp[1];
  uint32_t s[9] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w, 0u};
  uint32_t carry = 0;
  const uint32_t mask = (1u << c) - 1u;
  const uint32_t half = 1u << (c - 1);
  for (int w = 0; w < nwin; w++) {
    int bit = w * c;
    int word = bit >> 5, sh = bit & 31;
    uint64_t two = (word < 8) ? (uint64_t(s[word]) | (uint64_t(s[word + 1]) << 32)) : 0ull;
    uint32_t v = (uint32_t(two >> sh) & mask) + carry;
    uint32_t neg = 0;
    if (v > half) {  // digit = v - 2^c, negative
      v = (1u << c) - v;
      neg = 1;
      carry = 1;
    } else {
      carry = 0;
    }
    size_t o = size_t(w) * n + i;
    keys[o] = v ? (uint32_t(w) * nbuck + v - 1u) : sentinel;
    vals[o] = uint32_t(i) | (neg << 31);
  }
}

// ------------------------------------------------------------------------------------------- 3. accumulate
template <class F>
__device__ __forceinline__ Affine<F> load_affine(const Affine<F>* __restrict__ p) {
  // 64 B (G1) / 128 B (G2) as 16-byte vector loads through the read-only path
  Affine<F> r;
  const uint4* src = reinterpret_cast<const uint4*>(p);
  uint4* dst = reinterpret_cast<uint4*>(&r);
#pragma unroll
  for (int k = 0; k < int(sizeof(Affine<F>) / 16); k++) dst[k] = __ldg(src + k);
  return r;
}

template <class F>
__device__ __forceinline__ void store_xyzz(XYZZ<F>* p, const XYZZ<F>& v) {
  const uint4* src = reinterpret_cast<const uint4*>(&v);
  uint4* dst = reinterpret_cast<uint4*>(p);
#pragma unroll
  for (int k = 0; k < int(sizeof(XYZZ<F>) / 16); k++) dst[k] = src[k];
}

template <class F>
__device__ __forceinline__ XYZZ<F> load_xyzz(const XYZZ<F>* p) {
  XYZZ<F> r;
  const uint4* src = reinterpret_cast<const uint4*>(p);
  uint4* dst = reinterpret_cast<uint4*>(&r);
#pragma unroll
  for (int k = 0; k < int(sizeof(XYZZ<F>) / 16); k++) dst[k] = src[k];
  return r;
}

template <class F, int THREADS>
__global__ void __launch_bounds__(THREADS)
msm_accumulate_kernel(const Affine<F>* __restrict__ bases, const uint32_t* __restrict__ keys,
                      const uint32_t* __restrict__ vals, size_t total, int chunk, uint32_t sentinel,
                      XYZZ<F>* __restrict__ buckets, XYZZ<F>* __restrict__ heads, uint32_t* __restrict__ head_keys) {
  size_t t = size_t(blockIdx.x) * THREADS + threadIdx.x;
  size_t start = t * size_t(chunk);
  if (start >= total) return;
  size_t end = start + chunk;
  if (end > total) end = total;

  uint32_t cur = keys[start];
  if (cur >= sentinel) {
    head_keys[t] = sentinel;
    return;
  }
  bool first_run = true;
  XYZZ<F> acc = XYZZ<F>::inf();
  // software pipeline: the base of entry j+1 is in flight while entry j is added
  uint32_t v = vals[start];
  Affine<F> nxt = load_affine(bases + (v & 0x7fffffffu));
  uint32_t nxt_neg = v >> 31;
  for (size_t j = start; j < end; j++) {
    Affine<F> pt = nxt;
    uint32_t neg = nxt_neg;
    uint32_t k = cur;
    bool more = false;
    if (j + 1 < end) {
      k = keys[j + 1];
      if (k < sentinel) {
        uint32_t v2 = vals[j + 1];
        nxt = load_affine(bases + (v2 & 0x7fffffffu));
        nxt_neg = v2 >> 31;
        more = true;
      }
    }
    if (neg) pt.y = pt.y.neg();
    acc.madd(pt);
    if (!more || k != cur) {
      // run ends here
      if (first_run) {
        store_xyzz(heads + t, acc);
        head_keys[t] = cur;
        first_run = false;
      } else {
        store_xyzz(buckets + cur, acc);
      }
      acc = XYZZ<F>::inf();
      cur = k;
      if (!more) break;
    }
  }
}

// ------------------------------------------------------------------------------------------- 4. heads
template <class F>
__global__ void msm_heads_kernel(const XYZZ<F>* __restrict__ heads, const uint32_t* __restrict__ head_keys,
                                 size_t nthreads, uint32_t sentinel, XYZZ<F>* __restrict__ buckets) {
  size_t t = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= nthreads) return;
  uint32_t k = head_keys[t];
  if (k >= sentinel) return;
  if (t > 0 && head_keys[t - 1] == k) return;  // not the leader of this bucket's heads
  XYZZ<F> acc = load_xyzz(heads + t);
  for (size_t u = t + 1; u < nthreads && head_keys[u] == k; u++) acc.add(load_xyzz(heads + u));
  XYZZ<F> b = load_xyzz(buckets + k);
  b.add(acc);
  store_xyzz(buckets + k, b);
}

// ------------------------------------------------------------------------------------------- 5. reduce
// thread (w, s): R = sum_{b in segment} (b+1) * B[w][b] = running-sum part + b0 * (segment sum)
template <class F>
__global__ void msm_reduce_kernel(const XYZZ<F>* __restrict__ buckets, int nwin, uint32_t nbuck, int seg,
                                  XYZZ<F>* __restrict__ seg_out) {
  uint32_t nseg = nbuck / seg;
  size_t t = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= size_t(nwin) * nseg) return;
  uint32_t w = uint32_t(t / nseg), s = uint32_t(t % nseg);
  uint32_t b0 = s * seg;
  const XYZZ<F>* B = buckets + size_t(w) * nbuck + b0;
  XYZZ<F> run = XYZZ<F>::inf(), tot = XYZZ<F>::inf();
  for (int b = seg - 1; b >= 0; b--) {
    run.add(load_xyzz(B + b));
    tot.add(run);
  }
  if (b0) tot.add(run.mul_u32(b0));
  store_xyzz(seg_out + t, tot);
}

// ------------------------------------------------------------------------------------------- 6. window sums
// one block per window: tree reduction of nseg partial sums through shared memory
template <class F, int THREADS>
__global__ void __launch_bounds__(THREADS)
msm_window_sum_kernel(const XYZZ<F>* __restrict__ seg_out, uint32_t nseg, XYZZ<F>* __restrict__ win_out) {
  __shared__ XYZZ<F> sh[THREADS];
  const XYZZ<F>* in = seg_out + size_t(blockIdx.x) * nseg;
  XYZZ<F> acc = XYZZ<F>::inf();
  for (uint32_t i = threadIdx.x; i < nseg; i += THREADS) acc.add(load_xyzz(in + i));
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int stride = THREADS / 2; stride > 0; stride >>= 1) {
    if (int(threadIdx.x) < stride) {
      XYZZ<F> a = sh[threadIdx.x];
      a.add(sh[threadIdx.x + stride]);
      sh[threadIdx.x] = a;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) store_xyzz(win_out + blockIdx.x, sh[0]);
}

// ------------------------------------------------------------------------------------------- 7. final
template <class F>
__device__ void store_affine_canonical(const Affine<F>& a, uint32_t* out);
template <>
__device__ inline void store_affine_canonical<Fq>(const Affine<Fq>& a, uint32_t* out) {
  Fq x = a.x.from_mont(), y = a.y.from_mont();
#pragma unroll
  for (int i = 0; i < 8; i++) {
    out[i] = x.v[i];
    out[8 + i] = y.v[i];
  }
}
template <>
__device__ inline void store_affine_canonical<Fq2>(const Affine<Fq2>& a, uint32_t* out) {
  Fq v[4] = {a.x.c0.from_mont(), a.x.c1.from_mont(), a.y.c0.from_mont(), a.y.c1.from_mont()};
  for (int k = 0; k < 4; k++)
    for (int i = 0; i < 8; i++) out[8 * k + i] = v[k].v[i];
}

// Horner over windows; writes the XYZZ sum (Montgomery, for multi-GPU combining) and the canonical affine bytes.
template <class F>
__global__ void msm_final_kernel(const XYZZ<F>* __restrict__ win, int nwin, int c, XYZZ<F>* __restrict__ out_xyzz,
                                 uint32_t* __restrict__ out_affine) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  XYZZ<F> acc = load_xyzz(win + (nwin - 1));
  for (int w = nwin - 2; w >= 0; w--) {
    for (int k = 0; k < c; k++) acc = acc.dbl();
    acc.add(load_xyzz(win + w));
  }
  if (out_xyzz) store_xyzz(out_xyzz, acc);
  if (out_affine) store_affine_canonical<F>(acc.to_affine(), out_affine);
}

// sum of k XYZZ points (multi-GPU combine: one partial per rank) -> canonical affine
template <class F>
__global__ void msm_combine_kernel(const XYZZ<F>* __restrict__ parts, int k, uint32_t* __restrict__ out_affine) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  XYZZ<F> acc = XYZZ<F>::inf();
  for (int i = 0; i < k; i++) acc.add(load_xyzz(parts + i));
  store_affine_canonical<F>(acc.to_affine(), out_affine);
}

// ------------------------------------------------------------------------------------------- driver
template <class F>
struct MsmTraits;
template <>
struct MsmTraits<Fq> {
  static constexpr int ACC_THREADS = 128;
  static constexpr int THREADS_PER_SM = 512;
};
template <>
struct MsmTraits<Fq2> {
  static constexpr int ACC_THREADS = 64;
  static constexpr int THREADS_PER_SM = 256;
};

// bases: device, Montgomery affine.  scalars: device, canonical LE.  out_xyzz / out_affine: device (either may be null).
template <class F>
cudaError_t msm_run(zkb_ctx* ctx, const Affine<F>* bases, const uint32_t* scalars, size_t n, XYZZ<F>* out_xyzz,
                    uint32_t* out_affine) {
  using T = MsmTraits<F>;
  constexpr int PH0 = GroupOf<F>::PH0;
  cudaStream_t st = ctx->stream;
  if (n == 0) {
    // empty sum = infinity = all-zero encoding
    if (out_xyzz) cudaMemsetAsync(out_xyzz, 0, sizeof(XYZZ<F>), st);
    if (out_affine) cudaMemsetAsync(out_affine, 0, sizeof(Affine<F>), st);
    return cudaGetLastError();
  }
  MsmPlan p = msm_make_plan(n, ctx->msm_c, ctx->sm_count, T::THREADS_PER_SM);
  size_t total = size_t(p.nwin) * n;
  if (total >= (size_t(1) << 31)) return cudaErrorInvalidValue;
  size_t nthreads = (total + p.chunk - 1) / p.chunk;
  size_t nbuckets = size_t(p.nwin) * p.nbuck;
  uint32_t nseg = p.nbuck / p.seg;

  size_t sort_tmp = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, sort_tmp, (uint32_t*)nullptr, (uint32_t*)nullptr, (uint32_t*)nullptr,
                                  (uint32_t*)nullptr, int(total), 0, p.key_bits, st);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off += align_up(bytes);
    return o;
  };
  size_t o_k0 = take(total * 4), o_v0 = take(total * 4), o_k1 = take(total * 4), o_v1 = take(total * 4);
  size_t o_tmp = take(sort_tmp);
  size_t o_buck = take(nbuckets * sizeof(XYZZ<F>));
  size_t o_heads = take(nthreads * sizeof(XYZZ<F>));
  size_t o_hk = take(nthreads * 4);
  size_t o_seg = take(size_t(p.nwin) * nseg * sizeof(XYZZ<F>));
  size_t o_win = take(size_t(p.nwin) * sizeof(XYZZ<F>));
  cudaError_t e = ctx->msm_ws.reserve(off);
  if (e != cudaSuccess) return e;
  char* base = static_cast<char*>(ctx->msm_ws.p);
  uint32_t *k0 = (uint32_t*)(base + o_k0), *v0 = (uint32_t*)(base + o_v0);
  uint32_t *k1 = (uint32_t*)(base + o_k1), *v1 = (uint32_t*)(base + o_v1);
  XYZZ<F>* buckets = (XYZZ<F>*)(base + o_buck);
  XYZZ<F>* heads = (XYZZ<F>*)(base + o_heads);
  uint32_t* head_keys = (uint32_t*)(base + o_hk);
  XYZZ<F>* seg_out = (XYZZ<F>*)(base + o_seg);
  XYZZ<F>* win_out = (XYZZ<F>*)(base + o_win);

  {
    ProfScope ps(ctx, PH0 + 0);
    msm_digits_kernel<<<unsigned((n + 255) / 256), 256, 0, st>>>(scalars, n, p.c, p.nwin, p.nbuck, p.sentinel, k0, v0);
    ctx->launches++;
  }
  {
    ProfScope ps(ctx, PH0 + 1);
    e = cub::DeviceRadixSort::SortPairs(base + o_tmp, sort_tmp, k0, k1, v0, v1, int(total), 0, p.key_bits, st);
    if (e != cudaSuccess) return e;
  }
  cudaMemsetAsync(buckets, 0, nbuckets * sizeof(XYZZ<F>), st);
  cudaMemsetAsync(head_keys, 0xff, nthreads * 4, st);
  {
    ProfScope ps(ctx, PH0 + 2);
    unsigned acc_blocks = unsigned((nthreads + T::ACC_THREADS - 1) / T::ACC_THREADS);
    msm_accumulate_kernel<F, T::ACC_THREADS><<<acc_blocks, T::ACC_THREADS, 0, st>>>(
        bases, k1, v1, total, p.chunk, p.sentinel, buckets, heads, head_keys);
    ctx->launches++;
  }
  {
    ProfScope ps(ctx, PH0 + 3);
    msm_heads_kernel<F><<<unsigned((nthreads + 63) / 64), 64, 0, st>>>(heads, head_keys, nthreads, p.sentinel, buckets);
    size_t rthreads = size_t(p.nwin) * nseg;
    msm_reduce_kernel<F><<<unsigned((rthreads + 31) / 32), 32, 0, st>>>(buckets, p.nwin, p.nbuck, p.seg, seg_out);
    msm_window_sum_kernel<F, 64><<<p.nwin, 64, 0, st>>>(seg_out, nseg, win_out);
    msm_final_kernel<F><<<1, 32, 0, st>>>(win_out, p.nwin, p.c, out_xyzz, out_affine);
    ctx->launches += 4;
  }
  return cudaGetLastError();
}

}  // namespace zkb
