// Internal (non-ABI) declarations shared by the translation units of libzkb200.so:
//   zkb200.cu  context, profiling, R1CS / proving-key handles, prove orchestration, the extern "C" surface
//   fr.cu      Fr kernels: NTT passes, CSR mat-vec, QAP pointwise, witness map
//   g1.cu      G1 (Fq)  instantiation of group_impl.cuh: point import/export, MSM, fixed-base generation
//   g2.cu      G2 (Fq2) instantiation of the same
// Split so that nvcc compiles them in parallel (one TU took ~6 minutes).
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/zkb200.h"
#include "ec.cuh"

constexpr int ZKB_MAX_SLICES = 8;

namespace zkb {

// Bumped whenever device memory that a captured prove graph may have baked in is (re)allocated or freed: scratch buffers
// growing, keys / matrices / tables released.  A cached graph is replayed only while the epoch it was captured at is current.
inline std::atomic<unsigned long long> g_alloc_epoch{0};

// Scratch buffer.  With ZKB_GUARD=1 in the environment (checked when a buffer is allocated) every allocation is wrapped in two
// 4 KiB canary regions filled with 0xA5; zkb_debug_check_guards() verifies them.  compute-sanitizer is closed on this GPU pool,
// so this is the out-of-bounds check the scratch layouts (sort buffers, bucket arrays, head lists, batch buffers) run under in
// tests/test_gpu_parity.py::test_scratch_guards_stay_intact.
struct DevBuf {
  static constexpr size_t GUARD = 4096;
  void* p = nullptr;      // what users see
  void* base = nullptr;   // what cudaMalloc returned (== p without guards)
  size_t cap = 0;
  bool guarded = false;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    g_alloc_epoch.fetch_add(1, std::memory_order_relaxed);
    if (base) cudaFree(base);
    p = base = nullptr;
    cap = 0;
    if (bytes == 0) return cudaSuccess;
    const char* g = getenv("ZKB_GUARD");
    guarded = g && g[0] == '1';
    const size_t padded = (bytes + 255) / 256 * 256;
    cudaError_t e = cudaMalloc(&base, guarded ? padded + 2 * GUARD : bytes);
    if (e != cudaSuccess) {
      base = nullptr;
      return e;
    }
    if (guarded) {
      cudaMemset(base, 0xA5, GUARD);
      cudaMemset(static_cast<char*>(base) + GUARD + padded, 0xA5, GUARD);
      p = static_cast<char*>(base) + GUARD;
    } else {
      p = base;
    }
    cap = guarded ? padded : bytes;
    return e;
  }
  void release() {
    if (base) {
      g_alloc_epoch.fetch_add(1, std::memory_order_relaxed);
      cudaFree(base);
    }
    p = base = nullptr;
    cap = 0;
  }
  // 0: intact (or unguarded); 1: the canary before the buffer was overwritten; 2: the one after it
  int check_guards() const {
    if (!base || !guarded) return 0;
    unsigned char h[GUARD];
    for (int side = 0; side < 2; side++) {
      const char* src = side == 0 ? static_cast<const char*>(base) : static_cast<const char*>(base) + GUARD + cap;
      if (cudaMemcpy(h, src, GUARD, cudaMemcpyDeviceToHost) != cudaSuccess) return side + 1;
      for (size_t i = 0; i < GUARD; i++)
        if (h[i] != 0xA5) return side + 1;
    }
    return 0;
  }
  template <class T>
  T* as() const { return static_cast<T*>(p); }
};

// ---- per-phase device timing (CUDA events on the context's stream) ------------------------------------
enum Phase {
  PH_G1_DIGITS = 0, PH_G1_SORT, PH_G1_ACCUM, PH_G1_REDUCE,
  PH_G2_DIGITS, PH_G2_SORT, PH_G2_ACCUM, PH_G2_REDUCE,
  PH_NTT, PH_MATVEC, PH_POINTWISE, PH_ASSEMBLE,
  PH_COUNT
};

struct Prof {
  bool on = false;
  struct Span { int phase; cudaEvent_t a, b; };
  std::vector<Span> pending;
  std::vector<cudaEvent_t> pool;
  double ms[PH_COUNT] = {0};
  unsigned long long cnt[PH_COUNT] = {0};
  cudaEvent_t get() {
    if (!pool.empty()) {
      cudaEvent_t e = pool.back();
      pool.pop_back();
      return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
  }
};

}  // namespace zkb

struct zkb_ctx {
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  std::string err;
  unsigned long long launches = 0;
  int msm_c = 0;
  int msm_slices = 3;                             // host-scalar MSMs >= 2^20: upload slices (env ZKB_MSM_SLICES, 1..8)
  cudaStream_t copy_stream = nullptr;             // H2D slices of a host-scalar MSM, overlapped with compute
  cudaEvent_t copy_done[8] = {nullptr};
  zkb::Prof prof;
  zkb::DevBuf msm_ws;                             // MSM scratch (keys, sort space, buckets)
  // a prove runs its MSMs side by side: B2, A, B1, L on four auxiliary lanes (own stream + MSM scratch), the witness map
  // and H on the main stream
  struct AuxLane {
    cudaStream_t stream = nullptr;
    zkb::DevBuf ws;
    cudaEvent_t done = nullptr;
  } aux[5];                                       // [4]: s*A as its own MSM for small circuits (prove_device_part)
  cudaEvent_t ev_inputs = nullptr;
  zkb::DevBuf scal, res, tmp0, tmp1, tmp2, flag;  // staging
  zkb::DevBuf pz, pzm, pwa, pwb, pwc, ph, pza, pzb, pzl, prs, ppts, pzsa, pzrb;  // prove scratch
  // The device part of a prove (everything between the upload of z, r, s and the download of A, B, C: ~140 launches on five
  // streams) captured once per (key, matrices) as a CUDA graph and replayed with one cudaGraphLaunch.  Small proofs are
  // launch-bound: several contexts proving side by side serialise on the driver's launch path, not on the SMs.
  struct ProveGraph {
    const void* pk;
    const void* m;
    bool partial;
    unsigned long long epoch;
    cudaGraphExec_t exec;      // nullptr: seen once, uncaptured (the warm-up run that sizes every buffer)
    unsigned long long kernels;  // launches one replay stands for
    int failures;
  };
  std::vector<ProveGraph> graphs;
  bool graphs_on = true;
  // Wait for results by sleeping on an event instead of spinning in cudaStreamSynchronize: batch lanes outnumber the host
  // cores (8 GPUs x 16 lanes on one box) and a spinning waiter steals the core another lane's witness assignment needs.
  bool blocking_sync = false;
  cudaEvent_t sync_ev = nullptr;
  // 512 B of pinned host memory: proof bytes [0, 256) and the status flag [256, 260) come back here.  A device-to-host copy
  // into PAGEABLE memory blocks the calling thread (spinning in the driver) until everything queued before it has run.
  uint8_t* pinned = nullptr;
  unsigned long long graph_replays = 0, graph_captures = 0;
  zkb::DevBuf bz, bzm, bw3, bh, brs, bpart, bout;   // batched prove (zkb_prove_batch): K assignments, 3 K chains, K quotients
  uint8_t* bpinned = nullptr;                     // K x 256 B of results, pinned
  size_t bpinned_cap = 0;
  void* fr_state = nullptr;                       // NTT tables, owned by fr.cu
  void* g1_table = nullptr;                       // fixed-base tables, owned by g1.cu / g2.cu
  void* g2_table = nullptr;
};

// Device-resident bases with their window tables: p[j * n + i] = 2^(c j) * base_i, j < nwin (msm.cuh).
// inf_mask[i] != 0: base i is the point at infinity (a variable absent from that query's matrix: half of b_query in
// practice) -- the MSM drops its entries before the sort instead of carrying them as idle lanes through the accumulation.
// comb / comb_c / comb_nwin: the optional table of ALL digit multiples (msm.cuh comb_build_kernel), built for small keys by the
// first batched prove.
struct zkb_g1_bases {
  int device;
  zkb::Affine<zkb::Fq>* p;
  size_t n;
  int c, nwin;
  uint8_t* inf_mask;
  zkb::Affine<zkb::Fq>* comb = nullptr;
  int comb_c = 0, comb_nwin = 0;
};
struct zkb_g2_bases {
  int device;
  zkb::Affine<zkb::Fq2>* p;
  size_t n;
  int c, nwin;
  uint8_t* inf_mask;
  zkb::Affine<zkb::Fq2>* comb = nullptr;
  int comb_c = 0, comb_nwin = 0;
};

namespace zkb {

struct ProfScope {
  zkb_ctx* ctx;
  cudaEvent_t b = nullptr;
  int phase;
  ProfScope(zkb_ctx* c, int ph) : ctx(c), phase(ph) {
    if (!c->prof.on) return;
    cudaEvent_t a = c->prof.get();
    b = c->prof.get();
    cudaEventRecord(a, c->stream);
    c->prof.pending.push_back({ph, a, b});
  }
  ~ProfScope() {
    if (b) cudaEventRecord(b, ctx->stream);
  }
};

#define ZKB_FAIL(ctx, code, ...)           \
  do {                                     \
    char _b[512];                          \
    snprintf(_b, sizeof(_b), __VA_ARGS__); \
    (ctx)->err = _b;                       \
    return (code);                         \
  } while (0)

#define CUDA_TRY(ctx, expr)                                                                                  \
  do {                                                                                                       \
    cudaError_t _e = (expr);                                                                                 \
    if (_e != cudaSuccess) {                                                                                 \
      cudaGetLastError();                                                                                    \
      ZKB_FAIL(ctx, _e == cudaErrorMemoryAllocation ? ZKB_ERR_OOM : ZKB_ERR_CUDA, "%s: %s (%s:%d)", #expr,   \
               cudaGetErrorString(_e), __FILE__, __LINE__);                                                  \
    }                                                                                                        \
  } while (0)

#define ZKB_TRY(expr)            \
  do {                           \
    int _s = (expr);             \
    if (_s != ZKB_OK) return _s; \
  } while (0)

inline unsigned blocks_for(size_t n, int threads) { return unsigned((n + threads - 1) / threads); }
inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// zkb200.cu
int set_device(zkb_ctx* ctx);

// Every entry point runs on its context's device and puts the calling thread's current device back when it returns: a host
// that drives several GPUs from one thread (or mixes this library with another CUDA library, as the tests do with torch) must
// not find its current device changed behind its back.
struct DeviceGuard {
  int prev = -1;
  DeviceGuard() {
    if (cudaGetDevice(&prev) != cudaSuccess) {
      cudaGetLastError();
      prev = -1;
    }
  }
  explicit DeviceGuard(int device) : DeviceGuard() { cudaSetDevice(device); }
  ~DeviceGuard() {
    int cur = -1;
    if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
  }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};
#define ZKB_ON_DEVICE(ctx)   \
  zkb::DeviceGuard _zkb_dg;  \
  ZKB_TRY(zkb::set_device(ctx))
int clear_flag(zkb_ctx* ctx);
int check_flag(zkb_ctx* ctx, const char* what);  // synchronises the stream

template <class F> struct GroupOf;
template <> struct GroupOf<Fq> { using Bases = zkb_g1_bases; static constexpr int PH0 = PH_G1_DIGITS; };
template <> struct GroupOf<Fq2> { using Bases = zkb_g2_bases; static constexpr int PH0 = PH_G2_DIGITS; };

// group_impl.cuh, instantiated for Fq in g1.cu and Fq2 in g2.cu
template <class F> int import_points(zkb_ctx* ctx, const uint8_t* host, size_t n, int validate, Affine<F>* dst);
template <class F> int bases_load_impl(zkb_ctx* ctx, const uint8_t* host, size_t n, int validate, typename GroupOf<F>::Bases** out);
template <class F> int bases_load_compressed_impl(zkb_ctx* ctx, const uint8_t* host, size_t n, int validate, typename GroupOf<F>::Bases** out);
template <class F> int bases_generate_impl(zkb_ctx* ctx, const void* k_dev, size_t n, typename GroupOf<F>::Bases** out);
template <class F> int bases_read_impl(zkb_ctx* ctx, const typename GroupOf<F>::Bases* b, size_t offset, size_t n, uint8_t* out_host);
template <class F> void bases_free_impl(typename GroupOf<F>::Bases* b);
template <class F> int scalar_mul_impl(zkb_ctx* ctx, const uint8_t* points, const uint8_t* scalars, size_t n, uint8_t* out);
template <class F> int point_sum_impl(zkb_ctx* ctx, const uint8_t* points, size_t n, uint8_t* out);
template <class F> int msm_dev_impl(zkb_ctx* ctx, const typename GroupOf<F>::Bases* bases, size_t offset, const void* scalars_dev,
                                    size_t n, void* out_affine_dev, void* out_partial_dev);
template <class F> int msm_host_impl(zkb_ctx* ctx, const typename GroupOf<F>::Bases* bases, size_t offset, const uint8_t* scalars_host,
                                     size_t n, uint8_t* out);
template <class F> int msm_combine_impl(zkb_ctx* ctx, const void* parts, int k, void* out);
template <class F> int msm_host_partial_impl(zkb_ctx* ctx, const typename GroupOf<F>::Bases* bases, size_t offset,
                                             const uint8_t* scalars_host, size_t n, void* out_partial_dev);
template <class F> int msm_multi_impl(zkb_ctx* const* ctxs, const typename GroupOf<F>::Bases* const* bases, int n_gpus,
                                      const uint8_t* scalars_host, size_t n, uint8_t* out);
template <class F> int msm_batch_dev_impl(zkb_ctx* ctx, const typename GroupOf<F>::Bases* bases, size_t offset, const void* scalars_dev,
                                          size_t n, size_t stride, int batch, void* out_affine_dev, void* out_partial_dev);
template <class F> int msm_entries_debug_impl(zkb_ctx* ctx, const typename GroupOf<F>::Bases* bases, size_t offset, const void* scalars_dev,
                                              size_t n, size_t stride, int batch, void* out_keys, void* out_vals, void* out_count);
// comb table of every digit multiple for window width c (needs the window tables of the SAME c: the handle is rebuilt if not);
// comb_msm: `batch` MSMs through it
template <class F> int bases_build_comb(zkb_ctx* ctx, typename GroupOf<F>::Bases* bases, int c);
template <class F> size_t bases_comb_bytes(const typename GroupOf<F>::Bases* bases, int c);
template <class F> int msm_comb_dev_impl(zkb_ctx* ctx, const typename GroupOf<F>::Bases* bases, size_t offset, const void* scalars_dev,
                                         size_t n, size_t stride, int batch, void* out_partial_dev, void* out_affine_dev = nullptr);
template <class F> int fixed_table_for_base(zkb_ctx* ctx, const typename GroupOf<F>::Bases* bases, size_t idx, void** out_table);
template <class F> void fixed_table_free(zkb_ctx* ctx);
// out_host[i] = scalars_dev[i] * generator  (canonical 32 B scalars on the device; raw canonical affine out; zero -> infinity)
template <class F> int fixed_base_batch(zkb_ctx* ctx, const uint8_t* generator_raw, const void* scalars_dev, size_t n, uint8_t* out_host);

// g1.cu
int fq_field_op(zkb_ctx* ctx, int op, const uint8_t* a, const uint8_t* b, size_t n, uint8_t* out);
// C = s*A + r*B1 + L + H (projective partial sums on device) -> canonical affine
int prove_assemble_c(zkb_ctx* ctx, const void* pA, const void* pB1, const void* pL, const void* pH, const void* r_dev,
                     const void* s_dev, void* out_c_dev);
// C = SA + RB1 + L + H where the scalars were already folded into the MSMs (small circuits)
int prove_assemble_sum(zkb_ctx* ctx, const void* pSA, const void* pRB1, const void* pL, const void* pH, void* out_c_dev);
// sharded prove: sum `world` partial records [A | B1 | L | H (XYZZ G1) | B2 (XYZZ G2)] and finish the proof (g1.cu / g2.cu)
int prove_combine_g1(zkb_ctx* ctx, const void* parts, int world, size_t stride, const void* r_dev, const void* s_dev,
                     void* out_a_dev, void* out_c_dev);
int prove_combine_g2(zkb_ctx* ctx, const void* parts, int world, size_t stride, void* out_b_dev);
// Batched prove, last step (K proofs of one key).  G1: A = PA + a_query[0] + alpha + r delta, B1 = PB1 + b_query[0] + beta + s delta,
// C = s A + r B1 - r s delta + PL + PH; out: K x 256 B records, A at +0 and C at +192.  G2: B = PB2 + b0 + beta + s delta at +64.
// tails: the three points (query[0], alpha|beta, delta) = the last three bases of a_ext / b1_ext / b2_ext; fb_*: fixed-base
// tables of delta (fixed_table_for_base); rs: K x (r, s) canonical.
int prove_batch_finish_g1(zkb_ctx* ctx, int K, const void* PA, const void* PB1, const void* PL, const void* PH, const void* rs,
                          const void* a_tail, const void* b1_tail, const void* fb_delta1, void* out);
int prove_batch_finish_g2(zkb_ctx* ctx, int K, const void* PB2, const void* rs, const void* b2_tail, const void* fb_delta2, void* out);

// fr.cu
struct CsrDev {
  uint64_t* row_ptr = nullptr;
  uint32_t* col = nullptr;
  Fr* coeff = nullptr;  // Montgomery
  size_t nnz = 0;
};
int fr_field_op(zkb_ctx* ctx, int op, const uint8_t* a, const uint8_t* b, size_t n, uint8_t* out);
int fr_to_mont(zkb_ctx* ctx, const Fr* in, Fr* out, size_t n);
int fr_from_mont(zkb_ctx* ctx, const Fr* in, Fr* out, size_t n);   // flags values >= r like fr_to_mont
// MiMC-7 (forge stack): n hashes of `arity` elements; n Merkle roots along depth-long paths.  Flag non-canonical input in ctx->flag.
// Poseidon of the L2 circuit: params_canonical = 64 x 3 round constants | 3 x 3 MDS (canonical bytes; needed on the first call)
int poseidon_hash_dev(zkb_ctx* ctx, const uint8_t* params_canonical, int arity, const Fr* in, size_t n, Fr* out);
int mimc_hash_dev(zkb_ctx* ctx, int arity, const Fr* in, size_t n, Fr* out);
int mimc_merkle_roots_dev(zkb_ctx* ctx, const Fr* leaves, const Fr* siblings, const uint8_t* bits, size_t n, int depth, Fr* out);  // flags non-canonical input in ctx->flag
int ntt_dev_impl(zkb_ctx* ctx, const Fr* in, Fr* out, int logn, int inverse, int coset);
int ntt_batch_dev_impl(zkb_ctx* ctx, const Fr* in, Fr* out, int logn, int inverse, int coset, int batch);
void fr_state_free(zkb_ctx* ctx);
struct WitnessBufs {
  Fr *z, *zm, *wa, *wb, *wc;
};
int witness_map_dev(zkb_ctx* ctx, const CsrDev& A, const CsrDev& B, const CsrDev& C, uint64_t nc, uint64_t ni, uint64_t nw,
                    int log_domain, const WitnessBufs& w, Fr* h_out);
int witness_map_batch_dev(zkb_ctx* ctx, const CsrDev& A, const CsrDev& B, const CsrDev& C, uint64_t nc, uint64_t ni, uint64_t nw,
                          int log_domain, int K, const Fr* z, Fr* zm, bool zm_ready, Fr* w3, Fr* h_out);
int prove_tail_scalars(zkb_ctx* ctx, const Fr* r, const Fr* s, Fr* za_tail, Fr* zl_tail);
int fr_scale(zkb_ctx* ctx, const Fr* in, const Fr* k_dev, Fr* out, size_t n);  // out[i] = k * in[i], all canonical
int setup_scalars_dev(zkb_ctx* ctx, const uint64_t* const col_ptr[3], const uint32_t* const row[3], const Fr* const coeff[3],
                      uint64_t nc, uint64_t ni, uint64_t nw, int logn, const Fr* in5, Fr* consts, Fr* u, Fr* a_out, Fr* b_out,
                      Fr* abc_out, Fr* h_out);

}  // namespace zkb

struct zkb_r1cs {
  int device;
  uint64_t nc, ni, nw;
  int log_domain;
  zkb::CsrDev a, b, c;
};

struct zkb_pk {
  int device;
  size_t nv;  // num_instance + num_witness
  size_t nw;  // l_query length
  size_t nh;  // h_query length
  // each query with the constant terms appended so that the whole coefficient is ONE msm:
  //   a_ext  = a_query[1..]  || a_query[0]  || alpha_g1 || delta_g1      scalars: z[1..] || 1 || 1 || r
  //   b1_ext = b_g1_query[1..] || b_g1_query[0] || beta_g1 || delta_g1   scalars: z[1..] || 1 || 1 || s
  //   b2_ext = b_g2_query[1..] || b_g2_query[0] || beta_g2 || delta_g2   scalars: z[1..] || 1 || 1 || s
  //   l_ext  = l_query || delta_g1                                       scalars: aux   || -(r s)
  zkb_g1_bases *a_ext = nullptr, *b1_ext = nullptr, *l_ext = nullptr, *h = nullptr;
  zkb_g2_bases* b2_ext = nullptr;
  // A shard of a key (zkb_pk_load_shard): the handles above hold only the range [off, off + len) of each extended vector;
  // nv / nw / nh keep describing the WHOLE key.  Unsharded: offsets 0.
  size_t off_a = 0, off_l = 0, off_h = 0;   // a_ext, b1_ext and b2_ext share off_a
  int shard = 0, world = 1;
  int comb_state = 0;   // 0: not decided, 1: comb tables built for all five query vectors, -1: too large, bucket path
  void *fb_delta1 = nullptr, *fb_delta2 = nullptr;  // fixed-base tables of delta_g1 / delta_g2 (built by the first batched prove)
};
