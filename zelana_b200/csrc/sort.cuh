// Bucket sort of the MSM entries for sm_100a: signed-digit extraction fused with an LSD radix sort of (bucket, table index)
// pairs.  Replaces the digit kernel + cub::DeviceRadixSort of round 1 (a library sort on the hot path: 5.3 ms of a 39 ms
// 2^24 MSM) inside the replacement of ark-ec 0.5.0 `VariableBaseMSM::msm_bigint`, reached from
// core/src/sequencer/settlement/prover.rs:408.
//
// What is sorted.  Entry (point i, window w) has key = |digit| - 1 (+ proof * nbuck when a batch of scalar vectors shares one
// table) and value = table index | sign << 31.  Zero digits and the digits of bases at infinity produce NO entry (round 1
// carried them as sentinel keys through every pass).  Keys have key_bits <= 28 significant bits.
//
// Schedule (one stream, no host synchronisation; every count the host does not know stays on the device):
//   msm_entry_hist_kernel   reads the scalars once (32 B / point), histograms every radix digit of every key: hist[pass][bin]
//   sort_scan_kernel        exclusive scan per pass -> global bin bases; total = number of entries
//   msm_entry_pass0_kernel  reads the scalars again, extracts the digits in registers, ranks the tile's entries by the lowest
//                           radix digit (shared-memory atomics: the first pass need not be stable), stages them in shared
//                           memory in bin order and writes bin-contiguous runs: the unsorted key / value arrays never exist
//   sort_pass_kernel        the remaining passes: stable (warps find equal digits through shared-memory masks, warp-striped items), 8192 pairs per tile,
//                           staged through shared memory, coalesced runs out
// Tiles take their index from an atomic counter and chain their per-bin counts with a decoupled look-back (flag and value in
// one 32-bit word, so a single store publishes both): one read and one write of every pair per pass, 7 bits per pass.
#pragma once
#include <cuda_runtime.h>

#include <atomic>

#include <cstdint>

namespace zkb {

constexpr int SORT_RADIX_BITS = 7;                       // <= 128 bins per pass: long contiguous runs per bin per tile
constexpr int SORT_BINS = 1 << SORT_RADIX_BITS;
constexpr int SORT_MAX_PASSES = 4;                       // key_bits <= 28
constexpr int SORT_THREADS = 512;
constexpr int SORT_IPT = 16;                             // pairs per thread in a generic pass
constexpr int SORT_TILE = SORT_THREADS * SORT_IPT;       // 8192 pairs per tile, 2 resident tiles per SM: half as many look-back walks as 4096-pair tiles at the same occupancy (4.28 -> 4.24 ms at 2^24)
constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr int SORT_WSTRIDE = SORT_BINS + 1;              // per-warp counters: one extra bin for the padding of the last tile
constexpr uint32_t SORT_FLAG_AGG = 1u << 30, SORT_FLAG_INC = 2u << 30, SORT_VAL_MASK = (1u << 30) - 1;
constexpr int SORT_P0_THREADS = 256;                     // fused first pass: points per tile (one per thread); 128 when nwin > 64

struct SortPlan {
  int npass = 0;
  int shift[SORT_MAX_PASSES] = {0, 0, 0, 0};
  int bits[SORT_MAX_PASSES] = {0, 0, 0, 0};
};

// key_bits split evenly over ceil(key_bits / 7) passes, low digits first
static inline SortPlan sort_plan(int key_bits) {
  SortPlan p;
  if (key_bits < 1) key_bits = 1;
  p.npass = (key_bits + SORT_RADIX_BITS - 1) / SORT_RADIX_BITS;
  int base = key_bits / p.npass, extra = key_bits % p.npass, sh = 0;
  for (int i = 0; i < p.npass; i++) {
    p.bits[i] = base + (i < extra ? 1 : 0);
    p.shift[i] = sh;
    sh += p.bits[i];
  }
  return p;
}

// Device-side header of one sort: [0, npass*128) histograms -> bin bases, then the total, then one tile counter per pass.
struct SortHeader {
  static constexpr size_t HIST = 0;
  static constexpr size_t TOTAL = SORT_MAX_PASSES * SORT_BINS;
  static constexpr size_t TILE_CTR = TOTAL + 1;
  static constexpr size_t WORDS = TILE_CTR + SORT_MAX_PASSES + 3;  // padded to a multiple of 4 words
};

// What the entries are made from.
struct EntrySource {
  const uint32_t* scalars;      // canonical LE, 8 words each; vector p of the batch starts at scalars + p * stride * 8
  size_t n;                     // points per scalar vector
  size_t stride;                // scalars between consecutive vectors of a batch
  int batch;                    // number of scalar vectors sharing the table (1: a plain MSM)
  int c, nwin;
  uint32_t nbuck;               // 2^(c-1) buckets per vector
  size_t table_n, first;        // table index of (window w, point i) = w * table_n + first + i
  const uint8_t* inf_mask;      // inf_mask[first + i] != 0: base i is the point at infinity (may be null)
};

// Signed c-bit digits of a canonical scalar, low window first (the recoding of msm.cuh, round 1): v in [0, 2^(c-1)], carry out.
struct DigitWalker {
  uint32_t s[9];
  uint32_t carry = 0;
  __device__ __forceinline__ void load(const uint32_t* p) {
    const uint4* sp = reinterpret_cast<const uint4*>(p);
    uint4 lo = sp[0], hi = sp[1];
    s[0] = lo.x; s[1] = lo.y; s[2] = lo.z; s[3] = lo.w;
    s[4] = hi.x; s[5] = hi.y; s[6] = hi.z; s[7] = hi.w;
    s[8] = 0;
    carry = 0;
  }
  // digit of window w: returns |digit| (0 = no entry), sets neg
  __device__ __forceinline__ uint32_t next(int w, int c, uint32_t& neg) {
    const int bit = w * c;
    const int word = bit >> 5, sh = bit & 31;
    uint32_t lo = 0, hi = 0;
    // dynamic word index without local memory: the scalar stays in registers
#pragma unroll
    for (int k = 0; k < 8; k++)
      if (k == word) {
        lo = s[k];
        hi = s[k + 1];
      }
    uint32_t v = (__funnelshift_r(lo, hi, sh) & ((1u << c) - 1u)) + carry;
    neg = 0;
    if (v > (1u << (c - 1))) {
      v = (1u << c) - v;
      neg = 1;
      carry = 1;
    } else {
      carry = 0;
    }
    return v;
  }
};

// The same recoding with the scalar's words parked in shared memory, word-major (word k of thread t at w[k * stride + t]: bank
// = t, conflict-free): a window is two shared-memory loads instead of the sixteen predicated selects that keep a dynamically
// indexed scalar in registers -- the entry kernels were bound by exactly those ALU instructions (pass 0: ALU pipe 78 % busy,
// profiles/r02_ncu_sort_pass0.txt).
struct SmemDigitWalker {
  const uint32_t* w;
  int stride;
  uint32_t carry;
  __device__ __forceinline__ void park(uint32_t* base, int stride_, int t, const uint32_t* p) {
    const uint4* sp = reinterpret_cast<const uint4*>(p);
    uint4 lo = sp[0], hi = sp[1];
    uint32_t* d = base + t;
    d[0] = lo.x; d[stride_] = lo.y; d[2 * stride_] = lo.z; d[3 * stride_] = lo.w;
    d[4 * stride_] = hi.x; d[5 * stride_] = hi.y; d[6 * stride_] = hi.z; d[7 * stride_] = hi.w;
    d[8 * stride_] = 0;
    w = d;
    stride = stride_;
    carry = 0;
  }
  __device__ __forceinline__ uint32_t next(int win, int c, uint32_t& neg) {
    const int bit = win * c;
    const int word = bit >> 5, sh = bit & 31;   // word <= 7: the last window starts below bit 255
    const uint32_t lo = w[word * stride], hi = w[(word + 1) * stride];
    uint32_t v = (__funnelshift_r(lo, hi, sh) & ((1u << c) - 1u)) + carry;
    neg = 0;
    if (v > (1u << (c - 1))) {
      v = (1u << c) - v;
      neg = 1;
      carry = 1;
    } else {
      carry = 0;
    }
    return v;
  }
};

__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_volatile_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ------------------------------------------------------------------------------------------- histogram of every radix digit
static __global__ void __launch_bounds__(256)
msm_entry_hist_kernel(EntrySource src, SortPlan plan, uint32_t* __restrict__ hdr) {
  __shared__ uint32_t sh[SORT_MAX_PASSES * SORT_BINS];
  __shared__ uint32_t words[9 * 256];
  for (int i = threadIdx.x; i < SORT_MAX_PASSES * SORT_BINS; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  const size_t npts = size_t(src.batch) * src.n;
  for (size_t g = size_t(blockIdx.x) * blockDim.x + threadIdx.x; g < npts; g += size_t(gridDim.x) * blockDim.x) {
    const size_t p = g / src.n, i = g - p * src.n;
    if (src.inf_mask && src.inf_mask[src.first + i]) continue;
    SmemDigitWalker dw;   // each thread reads back only the words it parked itself: no barrier needed
    dw.park(words, 256, threadIdx.x, src.scalars + (p * src.stride + i) * 8);
    const uint32_t koff = uint32_t(p) * src.nbuck;
    for (int w = 0; w < src.nwin; w++) {
      uint32_t neg;
      uint32_t v = dw.next(w, src.c, neg);
      if (!v) continue;
      const uint32_t key = koff + v - 1u;
      for (int q = 0; q < plan.npass; q++) atomicAdd(&sh[q * SORT_BINS + ((key >> plan.shift[q]) & ((1u << plan.bits[q]) - 1u))], 1u);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < plan.npass * SORT_BINS; i += blockDim.x)
    if (sh[i]) atomicAdd(&hdr[SortHeader::HIST + i], sh[i]);
}

// hist -> exclusive bin bases per pass (in place); total = sum of one pass's histogram.  One warp per pass.
static __global__ void sort_scan_kernel(SortPlan plan, uint32_t* __restrict__ hdr) {
  const int q = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (q >= plan.npass) return;
  uint32_t* h = hdr + SortHeader::HIST + q * SORT_BINS;
  uint32_t run = 0;
  for (int b0 = 0; b0 < SORT_BINS; b0 += 32) {
    uint32_t v = h[b0 + lane], inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += t;
    }
    h[b0 + lane] = run + inc - v;
    run += __shfl_sync(0xffffffffu, inc, 31);
  }
  if (q == 0 && lane == 0) hdr[SortHeader::TOTAL] = run;
}

// Exclusive scan over the SORT_BINS per-bin tile counts (in cnt[]) -> start[]; every thread of the block calls it.
// tmp: SORT_BINS / 32 words of shared memory.
__device__ __forceinline__ void tile_bin_scan(const uint32_t* cnt, uint32_t* start, uint32_t* tmp) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint32_t v = tid < SORT_BINS ? cnt[tid] : 0u, inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += t;
  }
  if (tid < SORT_BINS && lane == 31) tmp[warp] = inc;
  __syncthreads();
  if (tid < SORT_BINS) {
    uint32_t off = 0;
    for (int k = 0; k < warp; k++) off += tmp[k];
    start[tid] = off + inc - v;
  }
}

// Decoupled look-back for bin `b` of tile `tile`: publishes this tile's count, returns the number of entries of this bin in
// all earlier tiles.  state: SORT_BINS words per tile, zero before the pass.
__device__ __forceinline__ uint32_t tile_lookback(uint32_t* state, uint32_t tile, int b, uint32_t count) {
  uint32_t* mine = state + size_t(tile) * SORT_BINS + b;
  if (tile == 0) {
    st_volatile_u32(mine, count | SORT_FLAG_INC);
    return 0;
  }
  st_volatile_u32(mine, count | SORT_FLAG_AGG);
  uint32_t prev = 0;
  // Walk back over the predecessors eight at a time: the eight loads are independent, so a walk of w tiles costs ~w / 8 memory
  // round trips instead of w (one thread per bin walks alone; with ~600 tiles in flight the serial walk was most of a tile's
  // time).  A word that is still unpublished (flag 0) is re-read until it is.
  long long t = (long long)tile - 1;
  bool done = false;
  while (!done) {
    uint32_t v[8];
    const int cnt = t + 1 < 8 ? int(t + 1) : 8;
#pragma unroll
    for (int k = 0; k < 8; k++)
      if (k < cnt) v[k] = ld_volatile_u32(state + size_t(t - k) * SORT_BINS + b);
#pragma unroll
    for (int k = 0; k < 8; k++) {
      if (k < cnt && !done) {
        while ((v[k] >> 30) == 0u) v[k] = ld_volatile_u32(state + size_t(t - k) * SORT_BINS + b);
        prev += v[k] & SORT_VAL_MASK;
        if ((v[k] >> 30) == 2u) done = true;
      }
    }
    t -= cnt;   // tile 0 publishes an inclusive word, so the walk always ends at or before it
  }
  st_volatile_u32(mine, (prev + count) | SORT_FLAG_INC);
  return prev;
}

// ------------------------------------------------------------------------------------------- pass 0, fused with the digits
// One point per thread, blockDim.x (>= SORT_BINS) points per tile.  Dynamic shared memory: blockDim.x * nwin (key, value) pairs.
static __global__ void __launch_bounds__(SORT_P0_THREADS)
msm_entry_pass0_kernel(EntrySource src, int shift, int bits, uint32_t* __restrict__ hdr, uint32_t* __restrict__ state,
                       uint32_t* __restrict__ out_keys, uint32_t* __restrict__ out_vals) {
  extern __shared__ uint2 p0_stage[];   // (key, value) pairs in bin order: one 8-byte access per pair
  __shared__ uint32_t cnt[SORT_BINS], cnt2[SORT_BINS], start[SORT_BINS], gofs[SORT_BINS], tmp[SORT_BINS / 32];
  __shared__ uint32_t words[9 * SORT_P0_THREADS];
  __shared__ uint32_t s_tile;
  const int tid = threadIdx.x;
  uint2* skv = p0_stage;
  if (tid == 0) s_tile = atomicAdd(&hdr[SortHeader::TILE_CTR + 0], 1u);
  if (tid < SORT_BINS) cnt[tid] = cnt2[tid] = 0;
  __syncthreads();
  const uint32_t tile = s_tile;
  const size_t npts = size_t(src.batch) * src.n;
  const size_t g = size_t(tile) * blockDim.x + tid;
  const uint32_t mask = (1u << bits) - 1u;
  bool live = g < npts;
  size_t p = 0, i = 0;
  if (live) {
    p = g / src.n;
    i = g - p * src.n;
    if (src.inf_mask && src.inf_mask[src.first + i]) live = false;
  }
  SmemDigitWalker dw;
  const uint32_t koff = uint32_t(p) * src.nbuck;
  if (live) {
    dw.park(words, SORT_P0_THREADS, tid, src.scalars + (p * src.stride + i) * 8);
    for (int w = 0; w < src.nwin; w++) {
      uint32_t neg;
      uint32_t v = dw.next(w, src.c, neg);
      if (v) atomicAdd(&cnt[((koff + v - 1u) >> shift) & mask], 1u);
    }
  }
  __syncthreads();
  tile_bin_scan(cnt, start, tmp);   // contains a barrier; start[] valid for tid < SORT_BINS afterwards
  if (tid < SORT_BINS) gofs[tid] = hdr[SortHeader::HIST + 0 * SORT_BINS + tid] + tile_lookback(state, tile, tid, cnt[tid]);
  __syncthreads();
  if (live) {
    dw.carry = 0;
    const uint32_t vbase = uint32_t(src.first + i);
    for (int w = 0; w < src.nwin; w++) {
      uint32_t neg;
      uint32_t v = dw.next(w, src.c, neg);
      if (!v) continue;
      const uint32_t key = koff + v - 1u;
      const uint32_t d = (key >> shift) & mask;
      const uint32_t pos = start[d] + atomicAdd(&cnt2[d], 1u);
      skv[pos] = make_uint2(key, (uint32_t(size_t(w) * src.table_n) + vbase) | (neg << 31));
    }
  }
  __syncthreads();
  const uint32_t items = start[SORT_BINS - 1] + cnt[SORT_BINS - 1];
  for (uint32_t j = tid; j < items; j += blockDim.x) {
    const uint2 kv = skv[j];
    const uint32_t d = (kv.x >> shift) & mask;
    const uint32_t dst = gofs[d] + (j - start[d]);
    out_keys[dst] = kv.x;
    out_vals[dst] = kv.y;
  }
}

// ------------------------------------------------------------------------------------------- generic stable pass
// Dynamic shared memory: 2 * SORT_TILE words (staged keys and values).
static __global__ void __launch_bounds__(SORT_THREADS, 2)
sort_pass_kernel(const uint32_t* __restrict__ in_keys, const uint32_t* __restrict__ in_vals, int pass, int shift, int bits,
                 uint32_t* __restrict__ hdr, uint32_t* __restrict__ state, uint32_t* __restrict__ out_keys,
                 uint32_t* __restrict__ out_vals) {
  extern __shared__ uint2 gp_stage[];   // (key, value) pairs in bin order
  __shared__ uint32_t wcnt[SORT_WARPS * SORT_WSTRIDE];
  __shared__ uint32_t wmask[2 * SORT_WARPS * SORT_WSTRIDE];
  __shared__ uint32_t cnt[SORT_BINS], start[SORT_BINS], gofs[SORT_BINS], tmp[SORT_BINS / 32];
  __shared__ uint32_t s_tile;
  uint2* skv = gp_stage;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_tile = atomicAdd(&hdr[SortHeader::TILE_CTR + pass], 1u);
  for (int k = tid; k < SORT_WARPS * SORT_WSTRIDE; k += SORT_THREADS) wcnt[k] = 0;
  for (int k = tid; k < 2 * SORT_WARPS * SORT_WSTRIDE; k += SORT_THREADS) wmask[k] = 0;
  __syncthreads();
  const uint32_t tile = s_tile;
  const uint32_t total = hdr[SortHeader::TOTAL];
  const uint32_t tile_base = tile * uint32_t(SORT_TILE);
  if (tile_base >= total) return;   // tiles are handed out in order: nothing behind this one has work either
  const uint32_t mask = (1u << bits) - 1u;
  const uint32_t wbase = tile_base + uint32_t(warp) * (32 * SORT_IPT) + lane;   // warp-striped: item r at wbase + 32 r

  uint32_t key[SORT_IPT];
  uint32_t rk[SORT_IPT / 2];   // two 16-bit warp-local ranks per word
#pragma unroll
  for (int r = 0; r < SORT_IPT; r++) {
    const uint32_t gi = wbase + 32u * r;
    key[r] = gi < total ? in_keys[gi] : 0xffffffffu;
  }
  // the values are only needed after the ranking and the look-back: ask L2 for their lines now (no registers, no shared memory)
#pragma unroll
  for (int r = 0; r < SORT_IPT; r++) {
    const uint32_t gi = wbase + 32u * r;
    if (gi < total) asm volatile("prefetch.global.L2 [%0];" ::"l"(in_vals + gi));
  }
  uint32_t* mycnt = wcnt + warp * SORT_WSTRIDE;
#pragma unroll
  for (int r = 0; r < SORT_IPT; r++) {
    const bool valid = wbase + 32u * r < total;
    const uint32_t d = valid ? ((key[r] >> shift) & mask) : uint32_t(SORT_BINS);
    // Which lanes of the warp hold the same digit?  match.any answers in one instruction but runs on the ADU pipe, which it
    // kept 75 % busy -- the kernel's limiter (profiles/r02_ncu_sort_pass.txt).  A shared-memory OR does the same on the
    // load/store pipe: every lane ORs its bit into the warp's mask word of its digit, a warp barrier, read it back.  Two mask
    // arrays alternate so that clearing a word (by the digit's lowest lane) never races with the next round's ORs.
    uint32_t* mym = wmask + ((r & 1) * SORT_WARPS + warp) * SORT_WSTRIDE;
    atomicOr(&mym[d], 1u << lane);
    __syncwarp();
    const uint32_t peers = mym[d];
    const uint32_t below = __popc(peers & ((1u << lane) - 1u));
    const uint32_t base = mycnt[d];
    __syncwarp();
    if (below == 0) {
      mycnt[d] = base + __popc(peers);
      mym[d] = 0;
    }
    const uint32_t rank = base + below;
    if (r & 1) rk[r >> 1] |= rank << 16;
    else rk[r >> 1] = rank;
  }
  __syncthreads();
  // per bin: exclusive offsets of the warps inside the tile, and the tile's count
  if (tid < SORT_BINS) {
    uint32_t run = 0;
#pragma unroll 4
    for (int w = 0; w < SORT_WARPS; w++) {
      uint32_t t = wcnt[w * SORT_WSTRIDE + tid];
      wcnt[w * SORT_WSTRIDE + tid] = run;
      run += t;
    }
    cnt[tid] = run;
  }
  __syncthreads();
  tile_bin_scan(cnt, start, tmp);
  if (tid < SORT_BINS) gofs[tid] = hdr[SortHeader::HIST + pass * SORT_BINS + tid] + tile_lookback(state, tile, tid, cnt[tid]);
  __syncthreads();
#pragma unroll
  for (int r = 0; r < SORT_IPT; r++) {
    const uint32_t gi = wbase + 32u * r;
    if (gi < total) {
      const uint32_t d = (key[r] >> shift) & mask;
      const uint32_t rank = (r & 1) ? (rk[r >> 1] >> 16) : (rk[r >> 1] & 0xffffu);
      const uint32_t pos = start[d] + mycnt[d] + rank;
      skv[pos] = make_uint2(key[r], in_vals[gi]);
    }
  }
  __syncthreads();
  const uint32_t items = total - tile_base < uint32_t(SORT_TILE) ? total - tile_base : uint32_t(SORT_TILE);
  for (uint32_t j = tid; j < items; j += SORT_THREADS) {
    const uint2 kv = skv[j];
    const uint32_t d = (kv.x >> shift) & mask;
    const uint32_t dst = gofs[d] + (j - start[d]);
    out_keys[dst] = kv.x;
    out_vals[dst] = kv.y;
  }
}

// ------------------------------------------------------------------------------------------- host driver
struct SortLayout {
  SortPlan plan;
  size_t max_entries = 0;      // upper bound the buffers are sized for
  size_t tiles0 = 0, tiles = 0;
  size_t state_words = 0;      // per pass
  size_t hdr_bytes = 0;        // header + npass tile-state arrays (one memset clears all of it)
};

static inline int sort_p0_threads(int nwin) { return nwin > 64 ? 128 : SORT_P0_THREADS; }

static inline SortLayout sort_layout(int key_bits, size_t npts, int nwin) {
  SortLayout L;
  L.plan = sort_plan(key_bits);
  L.max_entries = npts * size_t(nwin);
  L.tiles0 = (npts + sort_p0_threads(nwin) - 1) / sort_p0_threads(nwin);
  L.tiles = (L.max_entries + SORT_TILE - 1) / SORT_TILE;
  size_t t = L.tiles0 > L.tiles ? L.tiles0 : L.tiles;
  L.state_words = t * SORT_BINS;
  L.hdr_bytes = (SortHeader::WORDS + size_t(L.plan.npass) * L.state_words) * sizeof(uint32_t);
  return L;
}

static inline size_t sort_p0_smem(int nwin) { return size_t(sort_p0_threads(nwin)) * nwin * 2 * sizeof(uint32_t); }
constexpr size_t SORT_PASS_SMEM = size_t(SORT_TILE) * 2 * sizeof(uint32_t);

// Entries of `src` sorted by key.  hdr: L.hdr_bytes of scratch; (k0, v0), (k1, v1): two pairs of max_entries-word buffers.
// On return *out_keys / *out_vals name the buffers holding the sorted entries; hdr[SortHeader::TOTAL] is their number.
static inline cudaError_t msm_sort_entries(const EntrySource& src, const SortLayout& L, int sm_count, uint32_t* hdr, uint32_t* k0,
                                           uint32_t* v0, uint32_t* k1, uint32_t* v1, cudaStream_t st, const uint32_t** out_keys,
                                           const uint32_t** out_vals, unsigned long long* launches) {
  // Function attributes belong to the DEVICE that is current when they are set: once per device, not once per process (a
  // process driving several GPUs launched the 64 KB tiles on its second GPU without the opt-in).  Harmless if a racing
  // thread repeats it.
  static std::atomic<unsigned long long> attr_done{0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return cudaGetLastError();
  const unsigned long long bit = 1ull << (dev & 63);
  if (!(attr_done.load(std::memory_order_acquire) & bit)) {
    cudaError_t e = cudaFuncSetAttribute(sort_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(SORT_PASS_SMEM));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(msm_entry_pass0_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sort_p0_smem(64)));
    if (e != cudaSuccess) return e;
    attr_done.fetch_or(bit, std::memory_order_release);
  }
  if (src.nwin > 128) return cudaErrorInvalidValue;   // c >= 2
  if (L.max_entries >= (size_t(1) << 30)) return cudaErrorInvalidValue;   // look-back words carry 30-bit counts
  cudaError_t e = cudaMemsetAsync(hdr, 0, L.hdr_bytes, st);
  if (e != cudaSuccess) return e;
  const size_t npts = size_t(src.batch) * src.n;
  unsigned hist_blocks = unsigned((npts + 255) / 256);
  const unsigned cap = unsigned(sm_count) * 8u;
  if (hist_blocks > cap) hist_blocks = cap;
  if (hist_blocks == 0) hist_blocks = 1;
  msm_entry_hist_kernel<<<hist_blocks, 256, 0, st>>>(src, L.plan, hdr);
  sort_scan_kernel<<<1, 32 * SORT_MAX_PASSES, 0, st>>>(L.plan, hdr);
  uint32_t* state = hdr + SortHeader::WORDS;
  msm_entry_pass0_kernel<<<unsigned(L.tiles0 ? L.tiles0 : 1), sort_p0_threads(src.nwin), sort_p0_smem(src.nwin), st>>>(
      src, L.plan.shift[0], L.plan.bits[0], hdr, state, k0, v0);
  *launches += 3;
  uint32_t *ik = k0, *iv = v0, *ok = k1, *ov = v1;
  for (int q = 1; q < L.plan.npass; q++) {
    sort_pass_kernel<<<unsigned(L.tiles ? L.tiles : 1), SORT_THREADS, SORT_PASS_SMEM, st>>>(
        ik, iv, q, L.plan.shift[q], L.plan.bits[q], hdr, state + size_t(q) * L.state_words, ok, ov);
    (*launches)++;
    uint32_t* t = ik; ik = ok; ok = t;
    t = iv; iv = ov; ov = t;
  }
  *out_keys = ik;
  *out_vals = iv;
  return cudaGetLastError();
}

}  // namespace zkb
