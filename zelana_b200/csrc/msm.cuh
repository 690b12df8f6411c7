// Pippenger variable-base MSM for BN254 G1 / G2 on sm_100a.
//
// Replaces ark-ec 0.5.0 `VariableBaseMSM::msm_bigint` (signed-digit bucket method) at the five call
// sites of ark-groth16's create_proof_with_assignment (h_query, l_query, a_query, b_g1_query in G1 and
// b_g2_query in G2), which the reference enters from core/src/sequencer/settlement/prover.rs:408.
//
// B200 design.  A 254-bit Montgomery product costs ~570 scheduler cycles per warp whatever the occupancy
// (profiles/r01_kbench_field_mul_variants.txt), so the lever is the NUMBER of point additions:
//   * the bases are a proving key: they stay in HBM across proofs, so at load time every base P gets the
//     table 2^(c j) P, j < nwin (window_tables_kernel).  Digit j of a scalar then selects a point of table j
//     and ALL windows share one set of 2^(c-1) buckets: no per-window bucket reduction, no Horner doublings,
//     and c can grow to 22 (12 windows instead of 16) because the 2^21 buckets are reduced once, not 16 times.
//   * one (key, value) entry per (point, window): key = |digit| - 1, value = table index | sign << 31;
//     zero digits -- and every digit of a base that is the point at infinity (half of a real b_query) -- get the
//     sentinel key, sort last and never reach the accumulation.
// Pipeline (all on one stream, no host synchronisation; a host-scalar MSM uploads in slices that run it per slice):
//   1-2. sort.cuh               canonical scalars -> signed c-bit digits -> entries sorted by bucket: digit extraction fused
//                               into the first pass of an own LSD radix sort (7 bits per pass, decoupled look-back).
//   3. msm_accumulate_kernel    the hot loop.  The sorted entry list is cut into equal chunks, one per thread, so
//                               load balance does not depend on the scalar distribution.  A thread sums the runs
//                               in its chunk with XYZZ mixed additions (8M+2S) on gathered 64-byte affine points;
//                               a run that starts inside the chunk continues from the bucket's current content and is
//                               stored back (so upload slices accumulate on top of each other), the run that was
//                               already open at the chunk start goes to a per-thread "head" slot.  ncu: the fmaheavy
//                               pipe (IMAD.WIDE.U32.X) is 90 % busy.
//   4. msm_heads_warp_kernel    heads are again a sorted list: one lane per head, warp-segmented scan (5 shuffle
//                               steps), 32x shorter per level (a bucket holding millions of entries -- scalars 0/1 of
//                               real witnesses -- is folded in log_32 steps).
//   5. msm_bucket_seg_kernel    sum_b (b+1) B_b by segments of S = 32 buckets: W_s = local weighted sum, T_s = plain
//                               sum; sum_b (b+1) B_b = sum_s W_s + S * sum_s s T_s.  (S shrinks when there are few buckets.)
//      msm_rowcol_kernel        s = 256 hi + lo:  sum_s s T_s = 256 sum_hi hi Row_hi + sum_lo lo Col_lo (tree sums).
//      msm_plane_kernel         the two short weighted sums by bit planes: sum_i i A_i = sum_y 2^y (sum of A_i over i
//                               with bit y set); no scalar multiplications, no long dependent chains.
//   6. msm_final_kernel         one warp sums the planes with shuffles, XYZZ -> affine -> canonical bytes.
#pragma once
#include <cuda_runtime.h>

#include <atomic>

#include "internal.h"
#include "sort.cuh"

namespace zkb {

static inline int ilog2_ceil(size_t n) {
  int l = 0;
  while ((size_t(1) << l) < n) l++;
  return l;
}

constexpr int MSM_SEG_LOG_MAX = 5;  // bucket reduction: segments of up to 32 buckets (shorter when there are few buckets)
constexpr int MSM_COL_LOG = 8;   // segment totals viewed as rows x 256 columns

static inline int msm_windows_for(int c) { return (255 + c - 1) / c; }  // nwin * c >= 255: the top digit absorbs the carry

// Window width for a table of n bases: minimise  nwin(c) * n  (mixed additions)  +  4 * 2^(c-1)  (a bucket costs two full
// additions in the segment kernel plus its share of the reduction's later stages: ~4 mixed additions), subject to the table
// fitting `mem_budget` bytes and nwin * n < 2^31 entries.  Measured: 2^24 points -> 22 (21: +1.8 ms, 23: +1.4 ms), 2^21 -> 20
// (19: +0.3 ms, 21: +0.3 ms); the L2 key's 6-8 k point vectors -> 13 (a batch of 256 proofs: 5 653 proofs/s; 12: 5 570; 14: 5 069).
static inline int msm_choose_window(size_t n, size_t point_bytes, size_t mem_budget) {
  if (n == 0) return 8;
  int best = 0;
  double best_cost = 0;
  for (int c = 10; c <= 23; c++) {
    int nwin = msm_windows_for(c);
    if (double(nwin) * double(n) >= 2147483648.0) continue;
    if (double(nwin) * double(n) * double(point_bytes) > double(mem_budget)) continue;
    double cost = double(nwin) * double(n) + 4.0 * double(size_t(1) << (c - 1));
    if (!best || cost < best_cost) {
      best = c;
      best_cost = cost;
    }
  }
  return best;  // 0: no admissible width
}

// ------------------------------------------------------------------------------------------- window tables
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
// The same for the random gathers of the accumulation, with the L2 told to fetch only the 64 bytes around the address
// (ld.global.nc.L2::64B -> LDG.E.LTC64B): a G1 point is one aligned 64 B record, and without the hint every gather pulled its
// whole 128 B line from HBM -- 28.9 GB of DRAM reads per 2^24-point MSM against 14.5 GB algorithmic; with it 16.1 GB
// (profiles/r02_l2_fetch_granularity.txt; cudaLimitMaxL2FetchGranularity changes nothing).  The kernel is bound by the multiply
// pipe, so its time does not move (29.7 ms); the HBM energy and the bandwidth left to a concurrent upload do.
template <class F>
__device__ __forceinline__ Affine<F> gather_affine(const Affine<F>* __restrict__ p) {
  Affine<F> r;
  uint4* dst = reinterpret_cast<uint4*>(&r);
  const uint4* src = reinterpret_cast<const uint4*>(p);
#pragma unroll
  for (int k = 0; k < int(sizeof(Affine<F>) / 16); k++)
    asm volatile("ld.global.nc.L2::64B.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(dst[k].x), "=r"(dst[k].y), "=r"(dst[k].z), "=r"(dst[k].w) : "l"(src + k));
  return r;
}

template <class F>
__device__ __forceinline__ void store_affine(Affine<F>* p, const Affine<F>& v) {
  const uint4* src = reinterpret_cast<const uint4*>(&v);
  uint4* dst = reinterpret_cast<uint4*>(p);
#pragma unroll
  for (int k = 0; k < int(sizeof(Affine<F>) / 16); k++) dst[k] = src[k];
}

// table[j * n + i] = 2^(c j) * table[i] for 1 <= j < nwin (table[0..n) holds the bases).  One thread per base: the doubling
// chain stays in XYZZ, and the conversions back to affine share one inversion per group of up to 16 windows (Montgomery's
// trick on the ZZZ coordinates, kept in local memory).
template <class F>
__global__ void __launch_bounds__(128) window_tables_kernel(Affine<F>* table, size_t n, int c, int nwin) {
  constexpr int GROUP = 16;
  size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  XYZZ<F> cur = XYZZ<F>::from_affine(table[i]);
  XYZZ<F> pts[GROUP];
  F prefix[GROUP];
  for (int j0 = 1; j0 < nwin; j0 += GROUP) {
    const int g = nwin - j0 < GROUP ? nwin - j0 : GROUP;
    F pr = F::one();
    for (int k = 0; k < g; k++) {
      for (int d = 0; d < c; d++) cur = cur.dbl();
      pts[k] = cur;
      prefix[k] = pr;                       // product of the finite points' ZZZ before this one
      if (!cur.is_inf()) pr = pr * cur.zzz;
    }
    F inv = pr.inverse();
    for (int k = g - 1; k >= 0; k--) {
      Affine<F> a = Affine<F>::inf();
      if (!pts[k].is_inf()) {
        F zi3 = inv * prefix[k];            // 1 / zzz_k
        inv = inv * pts[k].zzz;
        F zi = zi3 * pts[k].zz;             // zz / zzz = 1 / z
        F zi2 = zi.sqr();
        a = {pts[k].x * zi2, pts[k].y * zi3};
      }
      store_affine(table + size_t(j0 + k) * n + i, a);
    }
  }
}

// ------------------------------------------------------------------------------------------- 1. digits
// (the digits themselves are extracted inside the sort: sort.cuh)
template <class F>
__global__ void inf_mask_kernel(const Affine<F>* __restrict__ bases, size_t n, uint8_t* __restrict__ mask) {
  size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) mask[i] = load_affine(bases + i).is_inf() ? 1 : 0;
}

// ------------------------------------------------------------------------------------------- 3. accumulate
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

// total_dev: the number of sorted entries (known only on the device: zero digits produce no entry).  The list is cut into
// `nthreads` equal chunks HERE, from that count, so every thread has work whatever the scalar distribution (witness-like
// scalars are half zeros); a chunk is at least `min_chunk` entries.
template <class F, int THREADS>
__global__ void __launch_bounds__(THREADS)
msm_accumulate_kernel(const Affine<F>* __restrict__ table, const uint32_t* __restrict__ keys,
                      const uint32_t* __restrict__ vals, const uint32_t* __restrict__ total_dev, size_t nthreads, int min_chunk,
                      uint32_t sentinel, XYZZ<F>* __restrict__ buckets, XYZZ<F>* __restrict__ heads,
                      uint32_t* __restrict__ head_keys) {
  size_t t = size_t(blockIdx.x) * THREADS + threadIdx.x;
  if (t >= nthreads) return;
  const size_t total = *total_dev;
  size_t chunk = (total + nthreads - 1) / nthreads;
  if (chunk < size_t(min_chunk)) chunk = size_t(min_chunk);
  size_t start = t * chunk;
  if (start >= total) {
    head_keys[t] = sentinel;
    return;
  }
  size_t end = start + chunk;
  if (end > total) end = total;

  uint32_t cur = keys[start];
  bool first_run = true;
  XYZZ<F> acc = XYZZ<F>::inf();
  // software pipeline: the point of entry j+1 is in flight while entry j is added
  uint32_t v = vals[start];
  Affine<F> nxt = gather_affine(table + (v & 0x7fffffffu));
  uint32_t nxt_neg = v >> 31;
  for (size_t j = start; j < end; j++) {
    Affine<F> pt = nxt;
    uint32_t neg = nxt_neg;
    uint32_t k = cur;
    const bool more = j + 1 < end;
    if (more) {
      k = keys[j + 1];
      uint32_t v2 = vals[j + 1];
      nxt = gather_affine(table + (v2 & 0x7fffffffu));
      nxt_neg = v2 >> 31;
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
      if (!more) break;
      // the next run starts inside this chunk: this thread owns bucket k.  It continues from the bucket's current content
      // (zero = infinity after the memset; the earlier slices' sum when a host-scalar MSM is uploaded in slices).
      acc = load_xyzz(buckets + k);
      cur = k;
    }
  }
}

// ------------------------------------------------------------------------------------------- 4. heads
template <class F>
__device__ __forceinline__ F shfl_up_field(const F& a, int d);
template <>
__device__ __forceinline__ Fq shfl_up_field<Fq>(const Fq& a, int d) {
  Fq r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = __shfl_up_sync(0xffffffffu, a.v[i], d);
  return r;
}
template <>
__device__ __forceinline__ Fq2 shfl_up_field<Fq2>(const Fq2& a, int d) {
  return {shfl_up_field<Fq>(a.c0, d), shfl_up_field<Fq>(a.c1, d)};
}
template <class F>
__device__ __forceinline__ XYZZ<F> shfl_up_xyzz(const XYZZ<F>& p, int d) {
  return {shfl_up_field<F>(p.x, d), shfl_up_field<F>(p.y, d), shfl_up_field<F>(p.zz, d), shfl_up_field<F>(p.zzz, d)};
}

// One level of head folding: the head list is sorted by key, one lane per head.  A warp does a segmented inclusive scan
// (5 shuffle steps, one addition each) so the last lane of every run holds the run's sum.  A run that starts inside the warp
// is ADDED to its bucket; the run that is open at lane 0 becomes a head of the next level (or, when `last`, also goes to
// its bucket).  Each level shrinks the list 32x with a latency of five additions.
template <class F>
__global__ void __launch_bounds__(64)
msm_heads_warp_kernel(const XYZZ<F>* __restrict__ in, const uint32_t* __restrict__ in_keys, size_t count, uint32_t sentinel,
                      XYZZ<F>* __restrict__ buckets, XYZZ<F>* __restrict__ out, uint32_t* __restrict__ out_keys, int last) {
  const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t warp = i >> 5;
  if ((warp << 5) >= count) return;  // whole warp beyond the list (grid rounding)
  const int lane = threadIdx.x & 31;
  uint32_t key = i < count ? in_keys[i] : sentinel;
  if (key > sentinel) key = sentinel;
  XYZZ<F> val = (i < count && key < sentinel) ? load_xyzz(in + i) : XYZZ<F>::inf();
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    XYZZ<F> o = shfl_up_xyzz(val, d);
    uint32_t ok = __shfl_up_sync(0xffffffffu, key, d);
    if (lane >= d && ok == key && key < sentinel) val.add(o);
  }
  uint32_t next_key = __shfl_down_sync(0xffffffffu, key, 1);
  uint32_t key0 = __shfl_sync(0xffffffffu, key, 0);
  bool run_end = (lane == 31) || next_key != key;
  if (key < sentinel && run_end) {
    if (key == key0 && !last) {
      store_xyzz(out + warp, val);
      out_keys[warp] = key;
    } else {
      XYZZ<F> b = load_xyzz(buckets + key);
      b.add(val);
      store_xyzz(buckets + key, b);
    }
  }
  if (lane == 0 && key >= sentinel && !last) out_keys[warp] = sentinel;
}

// ------------------------------------------------------------------------------------------- 5. bucket reduction
// R(A) = sum_{b < m} (b + 1) A[b].  Segment s of S = 2^seg_log entries: W[s] = sum_i (i + 1) A[sS + i], T[s] = sum_i A[sS + i];
// then R(A) = sum_s W[s] + S * R(T[1..]).
template <class F>
__global__ void __launch_bounds__(64)
msm_bucket_seg_kernel(const XYZZ<F>* __restrict__ in, size_t m, XYZZ<F>* __restrict__ W, XYZZ<F>* __restrict__ T, size_t nseg,
                      int seg_log) {
  size_t s = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (s >= nseg) return;
  const int S = 1 << seg_log;
  size_t b0 = s * S;
  XYZZ<F> run = XYZZ<F>::inf(), w = XYZZ<F>::inf();
  for (int i = S - 1; i >= 0; i--) {
    if (b0 + i < m) run.add(load_xyzz(in + b0 + i));
    w.add(run);
  }
  store_xyzz(W + s, w);
  store_xyzz(T + s, run);
}

// Block tree sum helper: every thread contributes `acc`; thread 0 returns the total.
template <class F, int THREADS>
__device__ __forceinline__ XYZZ<F> block_sum(XYZZ<F> acc, XYZZ<F>* sh) {
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
  return sh[0];
}

// sum_s s T[s] with s = hi * C + lo (C = 2^MSM_COL_LOG columns) = C * sum_hi hi * Row[hi] + sum_lo lo * Col[lo].
// blocks [0, rows): Row[hi] = sum_lo T[hi C + lo]; blocks [rows, 2 rows): WRow[hi] = sum_lo W[hi C + lo];
// blocks [2 rows, 2 rows + C): Col[lo] = sum_hi T[hi C + lo].   out = Row[rows] | WRow[rows] | Col[C]
// With few rows (MSMs of a few thousand points: rows <= MSM_COL_SEQ) a column is summed by ONE thread, THREADS columns per block, instead of a
// block per column tree-summing mostly empty slots: same depth, 1/64 of the resident blocks (small proofs run many at a time
// and are bound by block slots, not by arithmetic).
constexpr size_t MSM_COL_SEQ = 16;  // 32 rows (the H MSM of a 2^13 domain) measured slower this way: 218 us vs 77 us for the tree
template <int THREADS>
inline unsigned msm_rowcol_blocks(size_t rows) {
  constexpr size_t C = size_t(1) << MSM_COL_LOG;
  return unsigned(2 * rows + (rows <= MSM_COL_SEQ ? C / THREADS : C));
}
template <class F, int THREADS>
__global__ void __launch_bounds__(THREADS)
msm_rowcol_kernel(const XYZZ<F>* __restrict__ T, const XYZZ<F>* __restrict__ W, size_t nseg, size_t rows, XYZZ<F>* __restrict__ out) {
  __shared__ XYZZ<F> sh[THREADS];
  constexpr size_t C = size_t(1) << MSM_COL_LOG;
  const size_t b = blockIdx.x;
  XYZZ<F> acc = XYZZ<F>::inf();
  if (b < 2 * rows) {
    const XYZZ<F>* src = b < rows ? T : W;
    size_t hi = b < rows ? b : b - rows;
    for (size_t lo = threadIdx.x; lo < C; lo += THREADS)
      if (hi * C + lo < nseg) acc.add(load_xyzz(src + hi * C + lo));
  } else if (rows <= MSM_COL_SEQ) {  // block-uniform branch: these blocks never reach block_sum's barriers
    size_t lo = (b - 2 * rows) * THREADS + threadIdx.x;
    for (size_t hi = 0; hi < rows; hi++)
      if (hi * C + lo < nseg) acc.add(load_xyzz(T + hi * C + lo));
    store_xyzz(out + 2 * rows + lo, acc);
    return;
  } else {
    size_t lo = b - 2 * rows;
    for (size_t hi = threadIdx.x; hi < rows; hi += THREADS)
      if (hi * C + lo < nseg) acc.add(load_xyzz(T + hi * C + lo));
  }
  XYZZ<F> tot = block_sum<F, THREADS>(acc, sh);
  if (threadIdx.x == 0) store_xyzz(out + b, tot);
}

// Weighted sums of the short arrays by bit planes: block (y, a) with a in {0: Row, 1: Col}:
//   out[a * 16 + y] = 2^y * sum_{i : bit y of i} A_a[i]       (y < 16; the 2^y by y doublings in thread 0)
// and block (0, 2): out[32] = sum_i WRow[i].
template <class F, int THREADS>
__global__ void __launch_bounds__(THREADS)
msm_plane_kernel(const XYZZ<F>* __restrict__ rc, size_t rows, XYZZ<F>* __restrict__ out) {
  __shared__ XYZZ<F> sh[THREADS];
  constexpr size_t C = size_t(1) << MSM_COL_LOG;
  const int y = blockIdx.x, a = blockIdx.y;
  const XYZZ<F>* src = a == 0 ? rc : (a == 1 ? rc + 2 * rows : rc + rows);
  const size_t cnt = a == 1 ? C : rows;
  XYZZ<F> acc = XYZZ<F>::inf();
  if (a < 2 || y == 0)
    for (size_t i = threadIdx.x; i < cnt; i += THREADS)
      if (a == 2 || ((i >> y) & 1)) acc.add(load_xyzz(src + i));
  XYZZ<F> tot = block_sum<F, THREADS>(acc, sh);
  if (threadIdx.x == 0 && (a < 2 || y == 0)) {
    if (a < 2)
      for (int k = 0; k < y; k++) tot = tot.dbl();
    store_xyzz(out + (a == 2 ? 32 : a * 16 + y), tot);
  }
}

// ------------------------------------------------------------------------------------------- 6. final
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

// planes[0..16) = 2^y RowPlane_y, planes[16..32) = 2^y ColPlane_y, planes[32] = sum W:
//   result = sum_b (b+1) B_b = sum W + S * (C * sum_y planes[y] + sum_y planes[16 + y]).
// One warp: lanes tree-sum the two groups with shuffles, lane 0 finishes.  Writes the XYZZ sum (Montgomery, for multi-GPU
// combining) and the canonical affine bytes.
template <class F>
__global__ void __launch_bounds__(32)
msm_final_kernel(const XYZZ<F>* __restrict__ planes, int row_planes, int col_planes, int seg_log, XYZZ<F>* __restrict__ out_xyzz,
                 uint32_t* __restrict__ out_affine) {
  const int lane = threadIdx.x;
  const int y = lane & 15, grp = lane >> 4;
  XYZZ<F> v = XYZZ<F>::inf();
  if (y < (grp == 0 ? row_planes : col_planes)) v = load_xyzz(planes + grp * 16 + y);
#pragma unroll
  for (int d = 1; d < 16; d <<= 1) {  // after the loop lanes 15 and 31 hold their group's sum
    XYZZ<F> o = shfl_up_xyzz(v, d);
    if (y >= d) v.add(o);
  }
  XYZZ<F> row_total = shfl_up_xyzz(v, 16);  // lane 31 receives the row-group total held by lane 15
  if (lane != 31) return;
  XYZZ<F> acc = row_total;             // sum_y 2^y RowPlane_y
  for (int k = 0; k < MSM_COL_LOG; k++) acc = acc.dbl();
  acc.add(v);                          // + column part
  for (int k = 0; k < seg_log; k++) acc = acc.dbl();
  acc.add(load_xyzz(planes + 32));
  if (out_xyzz) store_xyzz(out_xyzz, acc);
  if (out_affine) store_affine_canonical<F>(acc.to_affine_vartime(), out_affine);
}

// sum of k XYZZ points (multi-GPU combine: one partial per rank) -> canonical affine
template <class F>
__global__ void msm_combine_kernel(const XYZZ<F>* __restrict__ parts, int k, uint32_t* __restrict__ out_affine) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  XYZZ<F> acc = XYZZ<F>::inf();
  for (int i = 0; i < k; i++) acc.add(load_xyzz(parts + i));
  store_affine_canonical<F>(acc.to_affine_vartime(), out_affine);
}

// ------------------------------------------------------------------------------------------- batched reduction
// K scalar vectors against ONE table (a batch of small proofs sharing a key): bucket array = K x nbuck, keys = p * nbuck + b.
// msm_bucket_seg_kernel runs over the flat array (S divides nbuck: segments never straddle two vectors); then the nsp = nbuck / S
// segment records of every vector are folded 32 at a time by warps (msm_fold_level_kernel), one or two levels:
//   R_p = sum_s W[p,s] + S * sum_s s T[p,s].
template <class F>
__device__ __forceinline__ F shfl_field(const F& a, int src);
template <>
__device__ __forceinline__ Fq shfl_field<Fq>(const Fq& a, int src) {
  Fq r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = __shfl_sync(0xffffffffu, a.v[i], src);
  return r;
}
template <>
__device__ __forceinline__ Fq2 shfl_field<Fq2>(const Fq2& a, int src) {
  return {shfl_field<Fq>(a.c0, src), shfl_field<Fq>(a.c1, src)};
}
template <class F>
__device__ __forceinline__ XYZZ<F> shfl_xyzz(const XYZZ<F>& p, int src) {
  return {shfl_field<F>(p.x, src), shfl_field<F>(p.y, src), shfl_field<F>(p.zz, src), shfl_field<F>(p.zzz, src)};
}
// lane 0 ends with the sum over all lanes
template <class F>
__device__ __forceinline__ XYZZ<F> warp_sum_xyzz(XYZZ<F> v, int lane) {
#pragma unroll 1
  for (int d = 16; d > 0; d >>= 1) {
    XYZZ<F> o = shfl_xyzz(v, (lane + d) & 31);
    if (lane < d) v.add(o);
  }
  return v;
}

// Fold records.  A record (T, B, W) stands for m = 2^log_m consecutive segments: T = sum of their segment totals,
// B = sum_i i T_i over them (local index), W = sum of their weighted sums.  Combining `cnt` consecutive records j = 0..cnt-1:
//   T' = sum_j T_j,   B' = m sum_j j T_j + sum_j B_j,   W' = sum_j W_j.
// At the first level the records are the segments themselves (Bin == nullptr: B = 0, m = 1).

// Sequential version, one THREAD per group of `cnt` (4 or 8) records: 2 additions per record for T' and sum_j j T_j (running
// sums), one each for W' and the incoming B_j, log_m doublings for the factor m.  Used (a) as a pre-fold that keeps the segment
// kernel's chains short (segments of 8 buckets instead of 32) without quadrupling the number of records the warp levels have to
// fold and (b), for batches of many vectors, for EVERY level: a chain of ~30 additions per level is slow (0.3-0.5 ms) but
// holds one lane per group, where the warp version holds three warps of ~200 registers per group -- measured in a pipeline of
// 256-proof sub-batches, the warp levels' 2.8 ms were not hidden behind the other sub-batches' accumulation at all (they occupy
// the register file while waiting on their own dependent chains); the sequential levels are.
// `finish`: the group is a whole vector; write R = W' + 2^seg_log B' to outT.
template <class F>
__global__ void __launch_bounds__(64)
msm_fold_seq_kernel(const XYZZ<F>* __restrict__ Tin, const XYZZ<F>* __restrict__ Bin, const XYZZ<F>* __restrict__ Win, size_t groups,
                    int cnt, int log_m, int finish, int seg_log, XYZZ<F>* __restrict__ outT, XYZZ<F>* __restrict__ outB,
                    XYZZ<F>* __restrict__ outW) {
  const size_t g = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (g >= groups) return;
  const XYZZ<F>* t = Tin + g * cnt;
  const XYZZ<F>* w = Win + g * cnt;
  XYZZ<F> run = XYZZ<F>::inf(), b = XYZZ<F>::inf(), ws = XYZZ<F>::inf();
  for (int j = cnt - 1; j >= 1; j--) {   // b = sum_j j T_j by running sums
    run.add(load_xyzz(t + j));
    b.add(run);
  }
  run.add(load_xyzz(t));
  for (int k = 0; k < log_m; k++) b = b.dbl();
  for (int j = 0; j < cnt; j++) {
    ws.add(load_xyzz(w + j));
    if (Bin) b.add(load_xyzz(Bin + g * cnt + j));
  }
  if (finish) {
    for (int k = 0; k < seg_log; k++) b = b.dbl();
    b.add(ws);
    store_xyzz(outT + g, b);
  } else {
    store_xyzz(outT + g, run);
    store_xyzz(outB + g, b);
    store_xyzz(outW + g, ws);
  }
}

// Warp version: a block of three warps combines `cnt` (<= 32) consecutive records, lane j of every warp holding record j.
// Warp 0: suffix sums of T over the lanes (5 shuffle steps); their sum over lanes 1..31 is sum_j j T_j and lane 0's suffix sum
// is T' -- no scalar multiplications.  Warps 1 and 2 tree-sum W and B meanwhile (three independent chains side by side
// instead of one after the other).  `finish`: the group is a whole vector; write R = W' + 2^seg_log B' to outT.
template <class F>
__global__ void __launch_bounds__(96)
msm_fold_level_kernel(const XYZZ<F>* __restrict__ Tin, const XYZZ<F>* __restrict__ Bin, const XYZZ<F>* __restrict__ Win, int cnt,
                      int log_m, int finish, int seg_log, XYZZ<F>* __restrict__ outT, XYZZ<F>* __restrict__ outB,
                      XYZZ<F>* __restrict__ outW) {
  __shared__ XYZZ<F> sh[2];
  const size_t g = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t item = g * size_t(cnt) + lane;
  XYZZ<F> tsum = XYZZ<F>::inf(), z = XYZZ<F>::inf();
  if (warp == 0) {
    XYZZ<F> sx = XYZZ<F>::inf();
    if (lane < cnt) sx = load_xyzz(Tin + item);
#pragma unroll 1
    for (int d = 1; d < 32; d <<= 1) {
      XYZZ<F> o = shfl_xyzz(sx, (lane + d) & 31);
      if (lane + d < 32) sx.add(o);
    }
    tsum = sx;   // lane 0: T'
    z = lane >= 1 ? sx : XYZZ<F>::inf();
    z = warp_sum_xyzz(z, lane);
  } else {
    const XYZZ<F>* src = warp == 1 ? Win : Bin;
    XYZZ<F> v = XYZZ<F>::inf();
    if (src && lane < cnt) v = load_xyzz(src + item);
    if (src) v = warp_sum_xyzz(v, lane);
    if (lane == 0) sh[warp - 1] = v;
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  for (int k = 0; k < log_m; k++) z = z.dbl();
  z.add(sh[1]);
  if (finish) {
    for (int k = 0; k < seg_log; k++) z = z.dbl();
    z.add(sh[0]);
    store_xyzz(outT + g, z);
  } else {
    store_xyzz(outT + g, tsum);
    store_xyzz(outB + g, z);
    store_xyzz(outW + g, sh[0]);
  }
}

// batches of at least this many vectors fold every level sequentially (throughput regime); smaller ones use the warp levels
// (latency regime: few groups, nothing else to fill the machine with)
constexpr int MSM_SEQ_FOLD_BATCH = 64;

// XYZZ -> canonical affine bytes, one thread per point (the K results of a batched MSM)
template <class F>
__global__ void xyzz_to_affine_bytes_kernel(const XYZZ<F>* __restrict__ in, size_t n, uint32_t* __restrict__ out) {
  size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  store_affine_canonical<F>(load_xyzz(in + i).to_affine_vartime(), out + i * (sizeof(Affine<F>) / 4));
}

// ------------------------------------------------------------------------------------------- comb tables (small keys)
// A batch of small proofs shares ONE key, and a B200 has 180 GB of HBM: for a key of a few thousand points the table can hold
// every multiple a signed c-bit digit can ask for,
//     comb[((i * nwin + w) << (c-1)) + d - 1] = d * 2^(c w) * P_i ,   1 <= d <= 2^(c-1),
// so that a (scalar, window) pair is ONE gathered affine point and an MSM is a plain sum of gathered points: no buckets, no
// sort, no bucket reduction, no heads -- those were half of the GPU time of a batch of L2-circuit proofs (profiles/
// r02_launches_batch256.txt).  c = 12 for the L2 circuit's five query vectors: 22 windows, ~110 GB.
// Built once per key from the window tables (2^(c w) P_i is already there): one thread per (point, window) walks the
// multiples by mixed additions and converts 16 at a time to affine with one shared inversion.
template <class F>
__global__ void __launch_bounds__(64)
comb_build_kernel(const Affine<F>* __restrict__ wtable, size_t n, int nwin, int c, Affine<F>* __restrict__ comb) {
  constexpr int GROUP = 16;
  const size_t t = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= n * size_t(nwin)) return;
  const size_t i = t / nwin;
  const int w = int(t - i * nwin);
  const Affine<F> base = load_affine(wtable + size_t(w) * n + i);
  Affine<F>* dst = comb + (t << (c - 1));
  const int count = 1 << (c - 1);
  XYZZ<F> cur = XYZZ<F>::inf();
  XYZZ<F> pts[GROUP];
  F prefix[GROUP];
  for (int d0 = 0; d0 < count; d0 += GROUP) {
    const int g = count - d0 < GROUP ? count - d0 : GROUP;
    F pr = F::one();
    for (int k = 0; k < g; k++) {
      cur.madd(base);                        // (d0 + k + 1) * base
      pts[k] = cur;
      prefix[k] = pr;
      if (!cur.is_inf()) pr = pr * cur.zzz;
    }
    F inv = pr.inverse();
    for (int k = g - 1; k >= 0; k--) {
      Affine<F> a = Affine<F>::inf();
      if (!pts[k].is_inf()) {
        F zi3 = inv * prefix[k];            // 1 / zzz_k
        inv = inv * pts[k].zzz;
        F zi = zi3 * pts[k].zz;             // 1 / z
        F zi2 = zi.sqr();
        a = {pts[k].x * zi2, pts[k].y * zi3};
      }
      store_affine(dst + d0 + k, a);
    }
  }
}

// One block sums  digit_w(s_i) * 2^(c w) * P_i  over `ppb` consecutive points of ONE scalar vector as gathered comb points.
// The work per point follows the scalar (a full-size field element has a non-zero digit in every window, a bit or a u64 amount
// in one or three), so the block first COMPACTS: every thread walks the digits of its points and appends one 32-bit entry per
// non-zero digit (local point, window, digit, sign) to a list in shared memory; then the list is dealt out evenly, entry k to
// thread k mod THREADS, each gather in flight while the previous point is added.  (A first version gave each thread a range
// of points: warps waited for their slowest lane, 4.2 G additions/s against 6.3 G/s for the bucket method's equal chunks.)
// grid (blocks per vector, vectors); dynamic shared memory: ppb * nwin entries.  partial[vector * gridDim.x + block] = the sum.
template <class F, int THREADS>
__global__ void __launch_bounds__(THREADS, THREADS >= 128 ? 4 : 1)   // G1: 128 registers, 4 blocks per SM like msm_accumulate_kernel
comb_accumulate_kernel(const Affine<F>* __restrict__ comb, int c, int nwin, const uint8_t* __restrict__ inf_mask, size_t first,
                       const uint32_t* __restrict__ scalars, size_t n, size_t stride, int ppb, XYZZ<F>* __restrict__ partial) {
  extern __shared__ uint4 comb_smem[];   // the entry list, then (once every thread is done with it) the tree sum's scratch
  uint32_t* comb_list = reinterpret_cast<uint32_t*>(comb_smem);
  __shared__ uint32_t s_count;
  const size_t p = blockIdx.y;
  const size_t lo = size_t(blockIdx.x) * size_t(ppb);
  size_t hi = lo + size_t(ppb);
  if (hi > n) hi = n;
  if (threadIdx.x == 0) s_count = 0;
  __syncthreads();
  const int dbits = c - 1;   // digit - 1 < 2^(c-1)
  for (size_t i = lo + threadIdx.x; i < hi; i += THREADS) {
    if (inf_mask && inf_mask[first + i]) continue;
    DigitWalker dw;
    dw.load(scalars + (p * stride + i) * 8);
    const uint32_t row = uint32_t(i - lo) * uint32_t(nwin);
    for (int w = 0; w < nwin; w++) {
      uint32_t neg;
      const uint32_t v = dw.next(w, c, neg);
      if (!v) continue;
      const uint32_t slot = atomicAdd(&s_count, 1u);
      comb_list[slot] = (((row + uint32_t(w)) << dbits) | (v - 1u)) | (neg << 31);   // (ppb * nwin) << (c - 1) < 2^31: checked by the host
    }
  }
  __syncthreads();
  const uint32_t count = s_count;
  const Affine<F>* base = comb + (((first + lo) * size_t(nwin)) << dbits);
  XYZZ<F> acc = XYZZ<F>::inf();
  uint32_t k = threadIdx.x;
  if (k < count) {
    uint32_t e = comb_list[k];
    Affine<F> nxt = load_affine(base + (e & 0x7fffffffu));
    uint32_t nxt_neg = e >> 31;
    for (;;) {
      Affine<F> pt = nxt;
      const uint32_t neg = nxt_neg;
      k += THREADS;
      const bool more = k < count;
      if (more) {
        e = comb_list[k];
        nxt = load_affine(base + (e & 0x7fffffffu));
        nxt_neg = e >> 31;
      }
      if (neg) pt.y = pt.y.neg();
      acc.madd(pt);
      if (!more) break;
    }
  }
  __syncthreads();
  XYZZ<F> tot = block_sum<F, THREADS>(acc, reinterpret_cast<XYZZ<F>*>(comb_smem));
  if (threadIdx.x == 0) store_xyzz(partial + p * gridDim.x + blockIdx.x, tot);
}

// out[p] = sum of the `nb` block sums of vector p: one warp per vector
template <class F>
__global__ void __launch_bounds__(32)
comb_finish_kernel(const XYZZ<F>* __restrict__ partial, int nb, XYZZ<F>* __restrict__ out) {
  const int lane = threadIdx.x;
  XYZZ<F> v = XYZZ<F>::inf();
  for (int k = lane; k < nb; k += 32) v.add(load_xyzz(partial + size_t(blockIdx.x) * nb + k));
  v = warp_sum_xyzz(v, lane);
  if (lane == 0) store_xyzz(out + blockIdx.x, v);
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

// Workspace layout of one MSM (or one batch of `batch` MSMs over the same table) whose entries arrive in up to `nslices`
// slices of at most `n_slice` points per scalar vector each.
template <class F>
struct MsmLayout {
  using P = XYZZ<F>;
  int c = 0, nwin = 0, nslices = 1, batch = 1, key_bits = 0, row_planes = 0, col_planes = 0, seg_log = 0, min_chunk = 16;
  uint32_t nbuck = 0;        // buckets per scalar vector
  uint32_t nbuck_all = 0;    // batch * nbuck: size of the bucket array; also the "no head" marker of the head lists
  SortLayout sort;
  size_t n_slice = 0, total = 0, nthreads = 0, nseg = 0, rows = 0, bytes = 0;
  size_t o_hdr, o_k0, o_v0, o_k1, o_v1, o_buck, o_h0, o_hk0, o_h1, o_hk1, o_W, o_T, o_part, o_planes;
};

template <class F>
static MsmLayout<F> msm_layout(int sm_count, int c, int nwin, size_t n_slice, int nslices, int batch = 1) {
  using T = MsmTraits<F>;
  using P = XYZZ<F>;
  MsmLayout<F> L;
  L.c = c;
  L.nwin = nwin;
  L.nslices = nslices;
  L.batch = batch;
  L.n_slice = n_slice;
  L.nbuck = 1u << (c - 1);
  L.nbuck_all = L.nbuck * uint32_t(batch);
  L.key_bits = ilog2_ceil(size_t(L.nbuck_all));
  L.total = size_t(nwin) * n_slice * size_t(batch);   // upper bound: zero digits produce no entry
  L.sort = sort_layout(L.key_bits, n_slice * size_t(batch), nwin);
  // The sorted list is cut into equal chunks, one per thread, with the thread count a WHOLE number of resident waves (no partial
  // last wave): at least 4 waves so that the block scheduler evens out per-thread variance, ~512 entries per thread when the
  // list is long (few heads: one per thread).  The chunk length itself is computed on the device from the real entry count.
  size_t resident = size_t(sm_count) * T::THREADS_PER_SM;
  size_t waves = (L.total + resident * 256) / (resident * 512);
  if (waves < 4) waves = 4;
  size_t chunk = (L.total + waves * resident - 1) / (waves * resident);
  if (chunk < size_t(L.min_chunk)) chunk = size_t(L.min_chunk);
  L.nthreads = (L.total + chunk - 1) / chunk;
  if (L.nthreads == 0) L.nthreads = 1;
  // Bucket reduction: segments of S buckets, then row / column sums of the segment totals and their bit planes.  S is chosen
  // so that ~2^17 segments exist (one thread each: enough to fill the GPU) before segments get long: a segment is a dependent
  // chain of 2 S additions, and at 8 GPUs (2^19 buckets per shard) the chain, not the arithmetic, was the reduction's time.
  L.seg_log = c - 1 - 17;
  L.seg_log = L.seg_log < 1 ? 1 : (L.seg_log > MSM_SEG_LOG_MAX ? MSM_SEG_LOG_MAX : L.seg_log);
  if (batch > 1) {
    // batched: segments of 8 buckets (a 16-addition chain per thread, batch * 2^(c-4) threads), then the per-vector folds
    L.seg_log = c - 1 < 3 ? c - 1 : 3;
  }
  if (c - 1 < L.seg_log) L.seg_log = c - 1;
  const size_t S = size_t(1) << L.seg_log;
  L.nseg = (size_t(L.nbuck_all) + S - 1) / S;
  L.rows = (L.nseg + (size_t(1) << MSM_COL_LOG) - 1) >> MSM_COL_LOG;
  L.row_planes = L.rows > 1 ? ilog2_ceil(L.rows) : 0;
  L.col_planes = L.nseg > 1 ? (L.nseg >= (size_t(1) << MSM_COL_LOG) ? MSM_COL_LOG : ilog2_ceil(L.nseg)) : 0;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off += align_up(bytes);
    return o;
  };
  L.o_hdr = take(L.sort.hdr_bytes);
  L.o_k0 = take(L.total * 4), L.o_v0 = take(L.total * 4), L.o_k1 = take(L.total * 4), L.o_v1 = take(L.total * 4);
  L.o_buck = take(size_t(L.nbuck_all) * sizeof(P));
  size_t nh1 = (L.nthreads + 31) / 32;
  L.o_h0 = take(L.nthreads * sizeof(P)), L.o_hk0 = take(L.nthreads * 4);
  L.o_h1 = take(nh1 * sizeof(P)), L.o_hk1 = take(nh1 * 4);
  L.o_W = take(L.nseg * sizeof(P));
  L.o_T = take(L.nseg * sizeof(P));
  L.o_part = take((2 * L.rows + (size_t(1) << MSM_COL_LOG)) * sizeof(P));
  L.o_planes = take(33 * sizeof(P));
  L.bytes = off;
  return L;
}

// Stages 1-4 for one slice: points [first, first + n) of the table, scalars (device, canonical LE; vector p of a batch at
// scalars + p * stride * 8 words), accumulated into the (shared) bucket array; `first_slice` clears it.
template <class F>
cudaError_t msm_accumulate_slice(zkb_ctx* ctx, const MsmLayout<F>& L, const Affine<F>* table, size_t table_n, const uint8_t* inf_mask,
                                 size_t first, const uint32_t* scalars, size_t n, size_t stride, bool first_slice) {
  using T = MsmTraits<F>;
  using P = XYZZ<F>;
  constexpr int PH0 = GroupOf<F>::PH0;
  cudaStream_t st = ctx->stream;
  char* base = static_cast<char*>(ctx->msm_ws.p);
  uint32_t* hdr = (uint32_t*)(base + L.o_hdr);
  uint32_t *k0 = (uint32_t*)(base + L.o_k0), *v0 = (uint32_t*)(base + L.o_v0);
  uint32_t *k1 = (uint32_t*)(base + L.o_k1), *v1 = (uint32_t*)(base + L.o_v1);
  P* buckets = (P*)(base + L.o_buck);
  P* hp[2] = {(P*)(base + L.o_h0), (P*)(base + L.o_h1)};
  uint32_t* hk[2] = {(uint32_t*)(base + L.o_hk0), (uint32_t*)(base + L.o_hk1)};
  const uint32_t *sk = nullptr, *sv = nullptr;
  {
    ProfScope ps(ctx, PH0 + 1);
    EntrySource src{scalars, n, stride, L.batch, L.c, L.nwin, L.nbuck, table_n, first, inf_mask};
    SortLayout sl = sort_layout(L.key_bits, n * size_t(L.batch), L.nwin);   // this slice's tile counts; buffers sized by L.sort
    cudaError_t e = msm_sort_entries(src, sl, ctx->sm_count, hdr, k0, v0, k1, v1, st, &sk, &sv, &ctx->launches);
    if (e != cudaSuccess) return e;
  }
  if (first_slice) cudaMemsetAsync(buckets, 0, size_t(L.nbuck_all) * sizeof(P), st);
  {
    ProfScope ps(ctx, PH0 + 2);
    unsigned acc_blocks = unsigned((L.nthreads + T::ACC_THREADS - 1) / T::ACC_THREADS);
    msm_accumulate_kernel<F, T::ACC_THREADS><<<acc_blocks, T::ACC_THREADS, 0, st>>>(
        table, sk, sv, hdr + SortHeader::TOTAL, L.nthreads, L.min_chunk, L.nbuck_all, buckets, hp[0], hk[0]);
    ctx->launches++;
  }
  {
    ProfScope ps(ctx, PH0 + 3);
    // heads: 32x per level; the last level (<= 32 heads) folds everything into the buckets
    size_t count = L.nthreads;
    int cur = 0;
    while (true) {
      int last = count <= 32 ? 1 : 0;
      size_t nw = (count + 31) / 32;
      msm_heads_warp_kernel<F><<<unsigned((nw * 32 + 63) / 64), 64, 0, st>>>(hp[cur], hk[cur], count, L.nbuck_all, buckets,
                                                                          hp[cur ^ 1], hk[cur ^ 1], last);
      ctx->launches++;
      if (last) break;
      count = nw;
      cur ^= 1;
    }
  }
  return cudaGetLastError();
}

// Stages 5-6: reduce the buckets, write the results.
template <class F>
cudaError_t msm_reduce(zkb_ctx* ctx, const MsmLayout<F>& L, XYZZ<F>* out_xyzz, uint32_t* out_affine) {
  using P = XYZZ<F>;
  constexpr int PH0 = GroupOf<F>::PH0;
  cudaStream_t st = ctx->stream;
  char* base = static_cast<char*>(ctx->msm_ws.p);
  P* buckets = (P*)(base + L.o_buck);
  P* W = (P*)(base + L.o_W);
  P* Tt = (P*)(base + L.o_T);
  P* part = (P*)(base + L.o_part);
  P* planes = (P*)(base + L.o_planes);
  ProfScope ps(ctx, PH0 + 3);
  msm_bucket_seg_kernel<F><<<unsigned((L.nseg + 63) / 64), 64, 0, st>>>(buckets, L.nbuck_all, W, Tt, L.nseg, L.seg_log);
  if (L.batch > 1) {
    // out_xyzz: `batch` partial sums; out_affine: `batch` canonical affine points
    P* res = out_xyzz ? out_xyzz : part;   // part holds 2 rows + 256 >= batch records only when batch <= 256: see msm_run_batch
    size_t cnt_items = size_t(L.nbuck >> L.seg_log);   // records per vector (a power of two)
    const P *tin = Tt, *bin = nullptr, *win = W;
    // intermediate records of a level: the buckets are dead once the segment kernel has run, so their array is the scratch
    P* scratch = buckets;
    int log_m = 0;
    ctx->launches++;
    if (L.batch >= MSM_SEQ_FOLD_BATCH) {   // many vectors: every level by threads, 8 records each (see msm_fold_seq_kernel)
      while (true) {
        const int cnt = cnt_items > 8 ? 8 : int(cnt_items);
        const size_t groups = size_t(L.batch) * (cnt_items / cnt);
        const int finish = cnt_items <= 8 ? 1 : 0;
        P* oT = finish ? res : scratch;
        msm_fold_seq_kernel<F><<<unsigned((groups + 63) / 64), 64, 0, st>>>(tin, bin, win, groups, cnt, log_m, finish, L.seg_log, oT,
                                                                           scratch + groups, scratch + 2 * groups);
        ctx->launches++;
        if (finish) break;
        tin = oT, bin = scratch + groups, win = scratch + 2 * groups;
        scratch += 3 * groups;
        cnt_items /= 8;
        log_m += 3;
      }
    } else {
      if (cnt_items >= 128) {   // sequential pre-fold of 4 records per thread: 4x fewer records for the warp levels
        const size_t groups = size_t(L.batch) * (cnt_items / 4);
        msm_fold_seq_kernel<F><<<unsigned((groups + 63) / 64), 64, 0, st>>>(tin, nullptr, win, groups, 4, 0, 0, L.seg_log, scratch,
                                                                           scratch + groups, scratch + 2 * groups);
        ctx->launches++;
        tin = scratch, bin = scratch + groups, win = scratch + 2 * groups;
        scratch += 3 * groups;
        cnt_items /= 4;
        log_m = 2;
      }
      while (true) {
        const int cnt = cnt_items > 32 ? 32 : int(cnt_items);
        const size_t groups = size_t(L.batch) * (cnt_items / cnt);
        const int finish = cnt_items <= 32 ? 1 : 0;
        P* oT = finish ? res : scratch;
        P* oB = scratch + groups;
        P* oW = scratch + 2 * groups;
        msm_fold_level_kernel<F><<<unsigned(groups), 96, 0, st>>>(tin, bin, win, cnt, log_m, finish, L.seg_log, oT, oB, oW);
        ctx->launches++;
        if (finish) break;
        tin = oT, bin = oB, win = oW;
        scratch += 3 * groups;
        cnt_items /= 32;
        log_m += 5;
      }
    }
    if (out_affine) {
      xyzz_to_affine_bytes_kernel<F><<<unsigned((L.batch + 63) / 64), 64, 0, st>>>(res, size_t(L.batch), out_affine);
      ctx->launches++;
    }
    return cudaGetLastError();
  }
  msm_rowcol_kernel<F, 64><<<msm_rowcol_blocks<64>(L.rows), 64, 0, st>>>(Tt, W, L.nseg, L.rows, part);
  const int nplanes = L.row_planes > L.col_planes ? L.row_planes : L.col_planes;  // planes beyond these sum nothing and are not read
  msm_plane_kernel<F, 64><<<dim3(nplanes > 0 ? nplanes : 1, 3), 64, 0, st>>>(part, L.rows, planes);
  msm_final_kernel<F><<<1, 32, 0, st>>>(planes, L.row_planes, L.col_planes, L.seg_log, out_xyzz, out_affine);
  ctx->launches += 4;
  return cudaGetLastError();
}

// table: device, Montgomery affine, nwin x table_n (window-major).  The MSM covers bases [first, first + n).
// scalars: device, canonical LE.  out_xyzz / out_affine: device (either may be null).
template <class F>
cudaError_t msm_run(zkb_ctx* ctx, const Affine<F>* table, size_t table_n, const uint8_t* inf_mask, int c, int nwin, size_t first,
                    const uint32_t* scalars, size_t n, XYZZ<F>* out_xyzz, uint32_t* out_affine) {
  cudaStream_t st = ctx->stream;
  if (n == 0) {
    // empty sum = infinity = all-zero encoding
    if (out_xyzz) cudaMemsetAsync(out_xyzz, 0, sizeof(XYZZ<F>), st);
    if (out_affine) cudaMemsetAsync(out_affine, 0, sizeof(Affine<F>), st);
    return cudaGetLastError();
  }
  if (size_t(nwin) * n >= (size_t(1) << 30) || size_t(nwin) * table_n >= (size_t(1) << 31)) return cudaErrorInvalidValue;
  MsmLayout<F> L = msm_layout<F>(ctx->sm_count, c, nwin, n, 1);
  cudaError_t e = ctx->msm_ws.reserve(L.bytes);
  if (e != cudaSuccess) return e;
  e = msm_accumulate_slice<F>(ctx, L, table, table_n, inf_mask, first, scalars, n, n, true);
  if (e != cudaSuccess) return e;
  return msm_reduce<F>(ctx, L, out_xyzz, out_affine);
}

// `batch` MSMs over the SAME bases [first, first + n): scalar vector p at scalars + p * stride * 8 words.  One sort, one
// accumulation, one reduction for all of them (a batch of small proofs sharing a proving key; per-vector bucket arrays).
// out_xyzz: batch x XYZZ (may be null); out_affine: batch x canonical affine (may be null).
template <class F>
cudaError_t msm_run_batch(zkb_ctx* ctx, const Affine<F>* table, size_t table_n, const uint8_t* inf_mask, int c, int nwin, size_t first,
                          const uint32_t* scalars, size_t n, size_t stride, int batch, XYZZ<F>* out_xyzz, uint32_t* out_affine) {
  cudaStream_t st = ctx->stream;
  if (batch < 1) return cudaErrorInvalidValue;
  if (batch == 1) return msm_run<F>(ctx, table, table_n, inf_mask, c, nwin, first, scalars, n, out_xyzz, out_affine);
  if (n == 0) {
    if (out_xyzz) cudaMemsetAsync(out_xyzz, 0, sizeof(XYZZ<F>) * batch, st);
    if (out_affine) cudaMemsetAsync(out_affine, 0, sizeof(Affine<F>) * batch, st);
    return cudaGetLastError();
  }
  if (size_t(nwin) * n * size_t(batch) >= (size_t(1) << 30) || size_t(nwin) * table_n >= (size_t(1) << 31) ||
      (size_t(batch) << (c - 1)) > (size_t(1) << 27) || (!out_xyzz && batch > 256))
    return cudaErrorInvalidValue;
  MsmLayout<F> L = msm_layout<F>(ctx->sm_count, c, nwin, n, 1, batch);
  cudaError_t e = ctx->msm_ws.reserve(L.bytes);
  if (e != cudaSuccess) return e;
  e = msm_accumulate_slice<F>(ctx, L, table, table_n, inf_mask, first, scalars, n, stride, true);
  if (e != cudaSuccess) return e;
  return msm_reduce<F>(ctx, L, out_xyzz, out_affine);
}

// `batch` MSMs over the comb table of the bases (see comb_build_kernel): out_xyzz: batch x XYZZ (required).
template <class F>
cudaError_t msm_run_comb(zkb_ctx* ctx, const Affine<F>* comb, int c, int nwin, const uint8_t* inf_mask, size_t first,
                         const uint32_t* scalars, size_t n, size_t stride, int batch, XYZZ<F>* out_xyzz) {
  using T = MsmTraits<F>;
  using P = XYZZ<F>;
  cudaStream_t st = ctx->stream;
  if (batch < 1 || !out_xyzz) return cudaErrorInvalidValue;
  if (n == 0) {
    cudaMemsetAsync(out_xyzz, 0, sizeof(P) * batch, st);
    return cudaGetLastError();
  }
  // points per block: ~88 additions per thread when every digit is non-zero (the block-wide tree sum at the end costs ~10),
  // bounded by the entry list's shared memory and its 31-bit entry format
  size_t ppb = size_t(T::ACC_THREADS) * 88 / size_t(nwin);
  if (ppb < 64) ppb = 64;
  while (ppb > 64 && (ppb * nwin * 4 > 96 * 1024 || ((ppb * size_t(nwin)) << (c - 1)) >= (size_t(1) << 31))) ppb /= 2;
  if (((ppb * size_t(nwin)) << (c - 1)) >= (size_t(1) << 31)) return cudaErrorInvalidValue;
  if (ppb > n) ppb = n;
  const size_t nb = (n + ppb - 1) / ppb;
  size_t smem = ppb * nwin * 4;
  if (smem < size_t(T::ACC_THREADS) * sizeof(P)) smem = size_t(T::ACC_THREADS) * sizeof(P);   // the tree sum reuses the list's memory
  static std::atomic<unsigned long long> attr_done{0};   // per device (see msm_sort_entries)
  const unsigned long long bit = 1ull << (ctx->device & 63);
  if (!(attr_done.load(std::memory_order_acquire) & bit)) {
    cudaError_t ea = cudaFuncSetAttribute(comb_accumulate_kernel<F, T::ACC_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    if (ea != cudaSuccess) return ea;
    attr_done.fetch_or(bit, std::memory_order_release);
  }
  cudaError_t e = ctx->msm_ws.reserve(size_t(batch) * nb * sizeof(P));
  if (e != cudaSuccess) return e;
  P* partial = static_cast<P*>(ctx->msm_ws.p);
  {
    ProfScope ps(ctx, GroupOf<F>::PH0 + 2);
    comb_accumulate_kernel<F, T::ACC_THREADS><<<dim3(unsigned(nb), unsigned(batch)), T::ACC_THREADS, smem, st>>>(
        comb, c, nwin, inf_mask, first, scalars, n, stride, int(ppb), partial);
  }
  {
    ProfScope ps(ctx, GroupOf<F>::PH0 + 3);
    comb_finish_kernel<F><<<unsigned(batch), 32, 0, st>>>(partial, int(nb), out_xyzz);
  }
  ctx->launches += 2;
  return cudaGetLastError();
}

// Host-scalar MSM pipelined against the PCIe copy: the scalars go up in `nslices` slices on a second stream; slice k is
// decomposed, sorted and accumulated (on top of the buckets the earlier slices left) while slice k+1 is still in flight.
template <class F>
cudaError_t msm_run_host_sliced(zkb_ctx* ctx, const Affine<F>* table, size_t table_n, const uint8_t* inf_mask, int c, int nwin, size_t first,
                                const uint8_t* scalars_host, uint32_t* scalars_dev, size_t n, int nslices,
                                XYZZ<F>* out_xyzz, uint32_t* out_affine) {
  if (size_t(nwin) * n >= (size_t(1) << 30) || size_t(nwin) * table_n >= (size_t(1) << 31)) return cudaErrorInvalidValue;
  // Slice k+1 is three times slice k: accumulating a point takes ~4x longer than copying its scalar, so every upload after
  // the (small) first one hides behind the previous slice's compute.
  if (nslices > ZKB_MAX_SLICES - 1) nslices = ZKB_MAX_SLICES - 1;
  size_t bound[ZKB_MAX_SLICES + 1];
  {
    double wsum = 0, w = 1;
    for (int k = 0; k < nslices; k++, w *= 3) wsum += w;
    double acc = 0;
    w = 1;
    bound[0] = 0;
    for (int k = 0; k < nslices; k++, w *= 3) {
      acc += w;
      size_t b = size_t(double(n) * acc / wsum);
      b = (b + 255) & ~size_t(255);
      bound[k + 1] = (k == nslices - 1 || b > n) ? n : b;
    }
  }
  size_t per = 0;
  for (int k = 0; k < nslices; k++)
    if (bound[k + 1] - bound[k] > per) per = bound[k + 1] - bound[k];
  MsmLayout<F> L = msm_layout<F>(ctx->sm_count, c, nwin, per, nslices);
  cudaError_t e = ctx->msm_ws.reserve(L.bytes);
  if (e != cudaSuccess) return e;
  if (!ctx->copy_stream) {
    e = cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) return e;
    for (int k = 0; k < ZKB_MAX_SLICES; k++) {
      e = cudaEventCreateWithFlags(&ctx->copy_done[k], cudaEventDisableTiming);
      if (e != cudaSuccess) return e;
    }
  }
  // the copy stream must not overwrite scalars_dev while earlier work on the compute stream still reads it
  cudaEventRecord(ctx->copy_done[ZKB_MAX_SLICES - 1], ctx->stream);
  cudaStreamWaitEvent(ctx->copy_stream, ctx->copy_done[ZKB_MAX_SLICES - 1], 0);
  for (int k = 0; k < nslices; k++) {
    size_t lo = bound[k], cnt = bound[k + 1] - bound[k];
    if (cnt) cudaMemcpyAsync(scalars_dev + lo * 8, scalars_host + lo * 32, cnt * 32, cudaMemcpyHostToDevice, ctx->copy_stream);
    cudaEventRecord(ctx->copy_done[k], ctx->copy_stream);
  }
  int used = 0;
  for (int k = 0; k < nslices; k++) {
    size_t lo = bound[k], cnt = bound[k + 1] - bound[k];
    cudaStreamWaitEvent(ctx->stream, ctx->copy_done[k], 0);
    if (!cnt) continue;
    e = msm_accumulate_slice<F>(ctx, L, table, table_n, inf_mask, first + lo, scalars_dev + lo * 8, cnt, cnt, used == 0);
    if (e != cudaSuccess) return e;
    used++;
  }
  if (!used) cudaMemsetAsync(static_cast<char*>(ctx->msm_ws.p) + L.o_buck, 0, size_t(L.nbuck_all) * sizeof(XYZZ<F>), ctx->stream);
  return msm_reduce<F>(ctx, L, out_xyzz, out_affine);
}

}  // namespace zkb
