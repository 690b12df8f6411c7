// Group-generic kernels and host drivers (G1 over Fq, G2 over Fq2).  Included by g1.cu and g2.cu only; each
// ends with explicit instantiations of the host templates declared in internal.h.
#pragma once
#include "internal.h"
#include "msm.cuh"

namespace zkb {

template <class F>
__device__ __forceinline__ bool fp_is_canonical(const F& a) {
  F m = F::modulus();
  for (int i = 7; i >= 0; i--) {
    if (a.v[i] < m.v[i]) return true;
    if (a.v[i] > m.v[i]) return false;
  }
  return false;  // equal to the modulus
}

template <class F> struct CurveB;
template <> struct CurveB<Fq> {
  static __device__ __forceinline__ Fq b() {  // 3 in Montgomery form
    Fq r;
    const uint32_t w[8] = {0x50ad28d7u, 0x7a17caa9u, 0xe15521b9u, 0x1f6ac17au, 0x696bd284u, 0x334bea4eu, 0xce179d8eu, 0x2a1f6744u};
    for (int i = 0; i < 8; i++) r.v[i] = w[i];
    return r;
  }
};
template <> struct CurveB<Fq2> {
  static __device__ __forceinline__ Fq2 b() {  // b' = 3/(9+u) in Montgomery form (oracle/bn254.py B_G2)
    const uint32_t w[16] = {0x77b802a8u, 0x3bf938e3u, 0x3633535du, 0x020b1b27u, 0x49755260u, 0x26b7edf0u,
                            0x4384a86du, 0x2514c632u, 0xd1dcff67u, 0x38e7ecccu, 0x93ce0d3eu, 0x65f0b37du,
                            0x22ac00aau, 0xd749d0ddu, 0x4a688d4du, 0x0141b9ceu};
    Fq2 r;
    for (int i = 0; i < 8; i++) { r.c0.v[i] = w[i]; r.c1.v[i] = w[8 + i]; }
    return r;
  }
};

template <class F> struct FieldIO;
template <> struct FieldIO<Fq> {
  static constexpr int WORDS = 8;
  static __device__ __forceinline__ bool load(const uint32_t* w, Fq& out) {
    for (int i = 0; i < 8; i++) out.v[i] = w[i];
    return fp_is_canonical(out);
  }
};
template <> struct FieldIO<Fq2> {
  static constexpr int WORDS = 16;
  static __device__ __forceinline__ bool load(const uint32_t* w, Fq2& out) {
    for (int i = 0; i < 8; i++) { out.c0.v[i] = w[i]; out.c1.v[i] = w[8 + i]; }
    return fp_is_canonical(out.c0) && fp_is_canonical(out.c1);
  }
};

template <class F> struct SubgroupCheck;
template <> struct SubgroupCheck<Fq> { static constexpr bool NEEDED = false; };
template <> struct SubgroupCheck<Fq2> { static constexpr bool NEEDED = true; };

// canonical affine bytes -> Montgomery affine, optional validation (on curve; G2: in the prime-order subgroup)
template <class F>
__global__ void affine_import_kernel(const uint32_t* in, Affine<F>* out, size_t n, int validate, int* bad) {
  size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  constexpr int W = FieldIO<F>::WORDS;
  F x, y;
  bool ok = FieldIO<F>::load(in + i * 2 * W, x);
  ok = FieldIO<F>::load(in + i * 2 * W + W, y) && ok;
  if (!ok) { atomicExch(bad, 1); return; }
  Affine<F> p{x.to_mont(), y.to_mont()};
  if (validate && !p.is_inf()) {
    F lhs = p.y.sqr();
    F rhs = p.x.sqr() * p.x + CurveB<F>::b();
    if (lhs != rhs) { atomicExch(bad, 2); return; }
    // same meaning of `validate` as the compressed loader and arkworks' Validate::Yes: G2 has a cofactor, so an on-curve
    // point must also be killed by r (G1 has cofactor 1: nothing to check)
    if (SubgroupCheck<F>::NEEDED && validate != 2) {   // validate == 2: on-curve only (the curve-arithmetic parity hooks)
      const uint32_t r[8] = {FrCfg::M0, FrCfg::M1, FrCfg::M2, FrCfg::M3, FrCfg::M4, FrCfg::M5, FrCfg::M6, FrCfg::M7};
      if (!XYZZ<F>::from_affine(p).mul_words(r).is_inf()) { atomicExch(bad, 4); return; }
    }
  }
  out[i] = p;
}

template <class F>
__global__ void affine_export_kernel(const Affine<F>* in, uint32_t* out, size_t n) {
  size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  store_affine_canonical<F>(in[i], out + i * (sizeof(Affine<F>) / 4));
}

// ---- ark-serialize COMPRESSED points (Groth16Prover::from_bytes -> ProvingKey::deserialize_compressed, prover.rs:263-277)
// Fq: p = 3 mod 4, sqrt(a) = a^((p+1)/4).  Values in and out are in Montgomery form.
__device__ inline bool fq_sqrt(const Fq& a, Fq& out) {
  const uint32_t e[8] = {0xb61f3f52u, 0x4f082305u, 0x5a1c72a3u, 0x65e05aa4u, 0xa0605617u, 0x6e14116du, 0xb84c680au, 0x0c19139cu};
  out = a.pow_words(e);
  return out.sqr() == a;
}
// "y is the larger of {y, -y}" on canonical integers: y > (p - 1) / 2
__device__ inline bool fq_is_larger(const Fq& y_mont) {
  const uint32_t h[8] = {0x6c3e7ea3u, 0x9e10460bu, 0xb438e546u, 0xcbc0b548u, 0x40c0ac2eu, 0xdc2822dbu, 0x7098d014u, 0x18322739u};
  Fq y = y_mont.from_mont();
  for (int i = 7; i >= 0; i--) {
    if (y.v[i] > h[i]) return true;
    if (y.v[i] < h[i]) return false;
  }
  return false;
}
// Fq2 square root, complex method (any root; the caller picks the sign)
__device__ inline bool fq2_sqrt(const Fq2& a, Fq2& out) {
  if (a.c1.is_zero()) {
    Fq s;
    if (fq_sqrt(a.c0, s)) {
      out = {s, Fq::zero()};
      return true;
    }
    if (fq_sqrt(a.c0.neg(), s)) {
      out = {Fq::zero(), s};
      return true;
    }
    return false;
  }
  Fq alpha;
  if (!fq_sqrt(a.c0.sqr() + a.c1.sqr(), alpha)) return false;
  Fq half;
  const uint32_t hw[8] = {0x4f060572u, 0x87bee7d2u, 0x2f1c6ae5u, 0xd0fd2addu, 0xfcfd4f44u, 0x8f5f7492u, 0x3d9cbfacu, 0x1f37631au};
  for (int i = 0; i < 8; i++) half.v[i] = hw[i];  // 1/2 in Montgomery form
  Fq x0;
  if (!fq_sqrt((a.c0 + alpha) * half, x0)) {
    if (!fq_sqrt((a.c0 - alpha) * half, x0)) return false;
  }
  Fq x1 = a.c1 * x0.dbl().inverse();
  out = {x0, x1};
  return out.sqr() == a;
}

template <class F> struct Compressed;
template <> struct Compressed<Fq> {
  static constexpr int WORDS = 8;
  static __device__ __forceinline__ bool load_x(const uint32_t* w, Fq& x) {
    for (int i = 0; i < 8; i++) x.v[i] = w[i];
    x.v[7] &= 0x3fffffffu;
    return fp_is_canonical(x);
  }
  static __device__ __forceinline__ bool sqrt(const Fq& a, Fq& out) { return fq_sqrt(a, out); }
  static __device__ __forceinline__ bool is_larger(const Fq& y) { return fq_is_larger(y); }
  static constexpr bool SUBGROUP_CHECK = false;  // cofactor 1
};
template <> struct Compressed<Fq2> {
  static constexpr int WORDS = 16;
  static __device__ __forceinline__ bool load_x(const uint32_t* w, Fq2& x) {
    for (int i = 0; i < 8; i++) { x.c0.v[i] = w[i]; x.c1.v[i] = w[8 + i]; }
    x.c1.v[7] &= 0x3fffffffu;
    return fp_is_canonical(x.c0) && fp_is_canonical(x.c1);
  }
  static __device__ __forceinline__ bool sqrt(const Fq2& a, Fq2& out) { return fq2_sqrt(a, out); }
  // Fq2 ordering compares c1 first, then c0
  static __device__ __forceinline__ bool is_larger(const Fq2& y) { return y.c1.is_zero() ? fq_is_larger(y.c0) : fq_is_larger(y.c1); }
  static constexpr bool SUBGROUP_CHECK = true;
};

// x || flags (bit 7 of the last byte: y is the larger root, bit 6: infinity) -> Montgomery affine.
// bad: 1 non-canonical x, 2 x not on the curve, 3 both flags set, 4 not in the prime-order subgroup.
template <class F>
__global__ void affine_decompress_kernel(const uint32_t* __restrict__ in, Affine<F>* __restrict__ out, size_t n, int validate,
                                         int* bad) {
  size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  constexpr int W = Compressed<F>::WORDS;
  const uint32_t* w = in + i * W;
  const uint32_t flags = w[W - 1] >> 30;
  if (flags == 3u) { atomicExch(bad, 3); return; }
  if (flags & 1u) {  // infinity
    out[i] = Affine<F>::inf();
    return;
  }
  F x;
  if (!Compressed<F>::load_x(w, x)) { atomicExch(bad, 1); return; }
  x = x.to_mont();
  F y;
  if (!Compressed<F>::sqrt(x.sqr() * x + CurveB<F>::b(), y)) { atomicExch(bad, 2); return; }
  if (Compressed<F>::is_larger(y) != bool(flags & 2u)) y = y.neg();
  Affine<F> p{x, y};
  if (validate && Compressed<F>::SUBGROUP_CHECK) {
    const uint32_t r[8] = {FrCfg::M0, FrCfg::M1, FrCfg::M2, FrCfg::M3, FrCfg::M4, FrCfg::M5, FrCfg::M6, FrCfg::M7};
    if (!XYZZ<F>::from_affine(p).mul_words(r).is_inf()) { atomicExch(bad, 4); return; }
  }
  out[i] = p;
}

template <class F>
__global__ void scalar_mul_kernel(const Affine<F>* pts, const uint32_t* scalars, size_t n, uint32_t* out) {
  size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t k[8];
  for (int j = 0; j < 8; j++) k[j] = scalars[i * 8 + j];
  XYZZ<F> r = XYZZ<F>::from_affine(pts[i]).mul_words(k);
  store_affine_canonical<F>(r.to_affine(), out + i * (sizeof(Affine<F>) / 4));
}

template <class F>
__global__ void point_sum_kernel(const Affine<F>* pts, size_t n, uint32_t* out) {
  if (blockIdx.x || threadIdx.x) return;
  XYZZ<F> acc = XYZZ<F>::inf();
  for (size_t i = 0; i < n; i++) acc.madd(pts[i]);
  store_affine_canonical<F>(acc.to_affine(), out);
}

// table[w * 255 + d - 1] = d * 2^(8w) * G
template <class F>
__global__ void fixed_base_table_kernel(Affine<F> gen, Affine<F>* table) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 32 * 255) return;
  int w = t / 255, d = t % 255 + 1;
  uint32_t k[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  k[w / 4] = uint32_t(d) << (8 * (w % 4));
  table[t] = XYZZ<F>::from_affine(gen).mul_words(k).to_affine();
}

template <class F>
__global__ void fixed_base_mul_kernel(const Affine<F>* __restrict__ table, const uint32_t* __restrict__ scalars,
                                      size_t n, Affine<F>* __restrict__ out) {
  size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  XYZZ<F> acc = XYZZ<F>::inf();
  for (int w = 0; w < 32; w++) {
    uint32_t d = (scalars[i * 8 + w / 4] >> (8 * (w % 4))) & 0xffu;
    if (d) acc.madd(table[w * 255 + d - 1]);
  }
  out[i] = acc.to_affine();
}

// k * G for a fixed_base_table_kernel table of G (8-bit windows: at most 32 mixed additions, no doublings); k: 8 canonical words
template <class F>
__device__ __forceinline__ XYZZ<F> fixed_table_mul(const Affine<F>* __restrict__ table, const uint32_t* k) {
  XYZZ<F> acc = XYZZ<F>::inf();
  for (int w = 0; w < 32; w++) {
    uint32_t d = (k[w >> 2] >> (8 * (w & 3))) & 0xffu;
    if (d) acc.madd(load_affine(table + w * 255 + d - 1));
  }
  return acc;
}

// same table, generator read from device memory (a key's delta)
template <class F>
__global__ void fixed_base_table_dev_kernel(const Affine<F>* gen, Affine<F>* table) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 32 * 255) return;
  int w = t / 255, d = t % 255 + 1;
  uint32_t k[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  k[w / 4] = uint32_t(d) << (8 * (w % 4));
  table[t] = XYZZ<F>::from_affine(*gen).mul_words(k).to_affine();
}

// ------------------------------------------------------------------------------------------- host drivers
template <class F>
int import_points(zkb_ctx* ctx, const uint8_t* host, size_t n, int validate, Affine<F>* dst) {
  if (n == 0) return ZKB_OK;
  size_t bytes = n * sizeof(Affine<F>);
  CUDA_TRY(ctx, ctx->tmp0.reserve(bytes));
  ZKB_TRY(clear_flag(ctx));
  CUDA_TRY(ctx, cudaMemcpyAsync(ctx->tmp0.p, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
  affine_import_kernel<F><<<blocks_for(n, 128), 128, 0, ctx->stream>>>(ctx->tmp0.as<uint32_t>(), dst, n, validate, ctx->flag.as<int>());
  ctx->launches++;
  CUDA_TRY(ctx, cudaGetLastError());
  return check_flag(ctx, "point import");
}

template <class F>
int scalar_mul_impl(zkb_ctx* ctx, const uint8_t* points, const uint8_t* scalars, size_t n, uint8_t* out) {
  if (n == 0) return ZKB_OK;
  size_t pbytes = n * sizeof(Affine<F>);
  CUDA_TRY(ctx, ctx->tmp1.reserve(pbytes));
  // on-curve only: keygen clears the G2 cofactor through this hook, i.e. multiplies points that are NOT in the subgroup yet
  ZKB_TRY(import_points<F>(ctx, points, n, 2, ctx->tmp1.as<Affine<F>>()));
  CUDA_TRY(ctx, ctx->scal.reserve(n * 32));
  CUDA_TRY(ctx, ctx->tmp2.reserve(pbytes));
  CUDA_TRY(ctx, cudaMemcpyAsync(ctx->scal.p, scalars, n * 32, cudaMemcpyHostToDevice, ctx->stream));
  scalar_mul_kernel<F><<<blocks_for(n, 64), 64, 0, ctx->stream>>>(ctx->tmp1.as<Affine<F>>(), ctx->scal.as<uint32_t>(), n, ctx->tmp2.as<uint32_t>());
  ctx->launches++;
  CUDA_TRY(ctx, cudaGetLastError());
  CUDA_TRY(ctx, cudaMemcpyAsync(out, ctx->tmp2.p, pbytes, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return ZKB_OK;
}

template <class F>
int point_sum_impl(zkb_ctx* ctx, const uint8_t* points, size_t n, uint8_t* out) {
  size_t pbytes = n * sizeof(Affine<F>);
  CUDA_TRY(ctx, ctx->tmp1.reserve(pbytes + sizeof(Affine<F>)));
  ZKB_TRY(import_points<F>(ctx, points, n, 2, ctx->tmp1.as<Affine<F>>()));
  CUDA_TRY(ctx, ctx->res.reserve(512));
  point_sum_kernel<F><<<1, 32, 0, ctx->stream>>>(ctx->tmp1.as<Affine<F>>(), n, ctx->res.as<uint32_t>());
  ctx->launches++;
  CUDA_TRY(ctx, cudaGetLastError());
  CUDA_TRY(ctx, cudaMemcpyAsync(out, ctx->res.p, sizeof(Affine<F>), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return ZKB_OK;
}

// Allocates the handle and the whole window table (nwin x n points); the bases go to its first n entries.
// The window width is fixed here, from n (or the context's override), and kept for every MSM on this handle.
template <class F>
int bases_alloc(zkb_ctx* ctx, size_t n, typename GroupOf<F>::Bases** out) {
  using H = typename GroupOf<F>::Bases;
  size_t free_b = 0, total_b = 0;
  CUDA_TRY(ctx, cudaMemGetInfo(&free_b, &total_b));
  int c = ctx->msm_c > 0 ? ctx->msm_c : msm_choose_window(n, sizeof(Affine<F>), free_b / 2);
  if (c <= 0 || double(msm_windows_for(c)) * double(n) >= 2147483648.0)
    ZKB_FAIL(ctx, ZKB_ERR_OOM, "bases: no window width fits %zu points in %zu MB of free device memory", n, free_b >> 20);
  int nwin = msm_windows_for(c);
  H* h = new (std::nothrow) H();
  if (h) {
    h->device = ctx->device;
    h->p = nullptr;
    h->n = n;
    h->c = c;
    h->nwin = nwin;
    h->inf_mask = nullptr;
  }
  if (!h) ZKB_FAIL(ctx, ZKB_ERR_OOM, "bases: host allocation failed");
  if (n) {
    cudaError_t e = cudaMalloc(&h->p, size_t(nwin) * n * sizeof(Affine<F>));
    if (e == cudaSuccess) {
      e = cudaMalloc(&h->inf_mask, n);
      if (e != cudaSuccess) cudaFree(h->p);
    }
    if (e != cudaSuccess) {
      cudaGetLastError();
      delete h;
      ZKB_FAIL(ctx, ZKB_ERR_OOM, "bases: cudaMalloc(%zu) failed: %s", size_t(nwin) * n * sizeof(Affine<F>), cudaGetErrorString(e));
    }
  }
  *out = h;
  return ZKB_OK;
}

template <class F>
int build_window_tables(zkb_ctx* ctx, typename GroupOf<F>::Bases* h) {
  if (h->n == 0) return ZKB_OK;
  inf_mask_kernel<F><<<blocks_for(h->n, 256), 256, 0, ctx->stream>>>(h->p, h->n, h->inf_mask);
  ctx->launches++;
  if (h->nwin <= 1) return ZKB_OK;
  window_tables_kernel<F><<<blocks_for(h->n, 128), 128, 0, ctx->stream>>>(h->p, h->n, h->c, h->nwin);
  ctx->launches++;
  CUDA_TRY(ctx, cudaGetLastError());
  return ZKB_OK;
}

template <class F>
int bases_load_impl(zkb_ctx* ctx, const uint8_t* host, size_t n, int validate, typename GroupOf<F>::Bases** out) {
  using H = typename GroupOf<F>::Bases;
  if (!ctx) return ZKB_ERR_INVALID_ARG;
  if (!out || (!host && n)) ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "bases_load: bad argument");
  *out = nullptr;
  ZKB_ON_DEVICE(ctx);
  H* h = nullptr;
  ZKB_TRY(bases_alloc<F>(ctx, n, &h));
  if (n) {
    int s = import_points<F>(ctx, host, n, validate, h->p);
    if (s == ZKB_OK) s = build_window_tables<F>(ctx, h);
    if (s != ZKB_OK) {
      cudaFree(h->p);
      cudaFree(h->inf_mask);
      delete h;
      return s;
    }
  }
  *out = h;
  return ZKB_OK;
}

// Same from ark-serialize compressed encodings (32 B per G1 point, 64 B per G2 point).
template <class F>
int bases_load_compressed_impl(zkb_ctx* ctx, const uint8_t* host, size_t n, int validate, typename GroupOf<F>::Bases** out) {
  using H = typename GroupOf<F>::Bases;
  if (!ctx) return ZKB_ERR_INVALID_ARG;
  if (!out || (!host && n)) ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "bases_load_compressed: bad argument");
  *out = nullptr;
  ZKB_ON_DEVICE(ctx);
  H* h = nullptr;
  ZKB_TRY(bases_alloc<F>(ctx, n, &h));
  if (n) {
    const size_t bytes = n * sizeof(Affine<F>) / 2;
    int s = ZKB_OK;
    auto run = [&]() -> int {
      CUDA_TRY(ctx, ctx->tmp0.reserve(bytes));
      ZKB_TRY(clear_flag(ctx));
      CUDA_TRY(ctx, cudaMemcpyAsync(ctx->tmp0.p, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
      affine_decompress_kernel<F><<<blocks_for(n, 64), 64, 0, ctx->stream>>>(ctx->tmp0.as<uint32_t>(), h->p, n, validate, ctx->flag.as<int>());
      ctx->launches++;
      CUDA_TRY(ctx, cudaGetLastError());
      ZKB_TRY(check_flag(ctx, "compressed point"));
      return build_window_tables<F>(ctx, h);
    };
    s = run();
    if (s != ZKB_OK) {
      cudaFree(h->p);
      cudaFree(h->inf_mask);
      delete h;
      return s;
    }
  }
  *out = h;
  return ZKB_OK;
}

template <class F> struct Generator;
template <> struct Generator<Fq> {
  static void canonical(uint8_t* out) {  // (1, 2)
    memset(out, 0, 64);
    out[0] = 1;
    out[32] = 2;
  }
  static void** slot(zkb_ctx* ctx) { return &ctx->g1_table; }
};
template <> struct Generator<Fq2> {
  static void canonical(uint8_t* out) {  // ark-bn254 g2::G2_GENERATOR_{X,Y}
    static const uint32_t w[32] = {
        0xd992f6edu, 0x46debd5cu, 0xf75edaddu, 0x674322d4u, 0x5e5c4479u, 0x426a0066u, 0x121f1e76u, 0x1800deefu,
        0xaef312c2u, 0x97e485b7u, 0x35a9e712u, 0xf1aa4933u, 0x31fb5d25u, 0x7260bfb7u, 0x920d483au, 0x198e9393u,
        0x66fa7daau, 0x4ce6cc01u, 0x0c43d37bu, 0xe3d1e769u, 0x8dcb408fu, 0x4aab7180u, 0xdb8c6debu, 0x12c85ea5u,
        0xd122975bu, 0x55acdadcu, 0x70b38ef3u, 0xbc4b3133u, 0x690c3395u, 0xec9e99adu, 0x585ff075u, 0x090689d0u};
    memcpy(out, w, 128);
  }
  static void** slot(zkb_ctx* ctx) { return &ctx->g2_table; }
};

template <class F>
int ensure_fixed_table(zkb_ctx* ctx) {
  void** slot = Generator<F>::slot(ctx);
  if (*slot) return ZKB_OK;
  Affine<F>* table = nullptr;
  CUDA_TRY(ctx, cudaMalloc(&table, 32 * 255 * sizeof(Affine<F>)));
  uint8_t gen[128];
  Generator<F>::canonical(gen);
  CUDA_TRY(ctx, ctx->tmp1.reserve(sizeof(Affine<F>)));
  int s = import_points<F>(ctx, gen, 1, 1, ctx->tmp1.as<Affine<F>>());
  if (s != ZKB_OK) {
    cudaFree(table);
    return s;
  }
  Affine<F> g;
  CUDA_TRY(ctx, cudaMemcpy(&g, ctx->tmp1.p, sizeof(g), cudaMemcpyDeviceToHost));
  fixed_base_table_kernel<F><<<blocks_for(32 * 255, 64), 64, 0, ctx->stream>>>(g, table);
  ctx->launches++;
  CUDA_TRY(ctx, cudaGetLastError());
  *slot = table;
  return ZKB_OK;
}

template <class F>
void fixed_table_free(zkb_ctx* ctx) {
  void** slot = Generator<F>::slot(ctx);
  if (*slot) cudaFree(*slot);
  *slot = nullptr;
}

// Fixed-base batch multiplication by an ARBITRARY generator (the setup's random g1 / g2): window table for that point, one
// thread per scalar, results exported as canonical affine bytes to the host.
template <class F>
int fixed_base_batch(zkb_ctx* ctx, const uint8_t* generator_raw, const void* scalars_dev, size_t n, uint8_t* out_host) {
  if (n == 0) return ZKB_OK;
  Affine<F>* table = nullptr;
  Affine<F>* pts = nullptr;
  uint32_t* raw = nullptr;
  auto cleanup = [&]() {
    if (table) cudaFree(table);
    if (pts) cudaFree(pts);
    if (raw) cudaFree(raw);
  };
  auto run = [&]() -> int {
    CUDA_TRY(ctx, cudaMalloc(&table, 32 * 255 * sizeof(Affine<F>)));
    CUDA_TRY(ctx, cudaMalloc(&pts, n * sizeof(Affine<F>)));
    CUDA_TRY(ctx, cudaMalloc(&raw, n * sizeof(Affine<F>)));
    CUDA_TRY(ctx, ctx->tmp1.reserve(sizeof(Affine<F>)));
    ZKB_TRY(import_points<F>(ctx, generator_raw, 1, 1, ctx->tmp1.as<Affine<F>>()));
    Affine<F> g;
    CUDA_TRY(ctx, cudaMemcpy(&g, ctx->tmp1.p, sizeof(g), cudaMemcpyDeviceToHost));
    fixed_base_table_kernel<F><<<blocks_for(32 * 255, 64), 64, 0, ctx->stream>>>(g, table);
    fixed_base_mul_kernel<F><<<blocks_for(n, 64), 64, 0, ctx->stream>>>(table, static_cast<const uint32_t*>(scalars_dev), n, pts);
    affine_export_kernel<F><<<blocks_for(n, 128), 128, 0, ctx->stream>>>(pts, raw, n);
    ctx->launches += 3;
    CUDA_TRY(ctx, cudaGetLastError());
    CUDA_TRY(ctx, cudaMemcpyAsync(out_host, raw, n * sizeof(Affine<F>), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return ZKB_OK;
  };
  int s = run();
  cleanup();
  return s;
}

template <class F>
int bases_generate_impl(zkb_ctx* ctx, const void* k_dev, size_t n, typename GroupOf<F>::Bases** out) {
  using H = typename GroupOf<F>::Bases;
  if (!ctx) return ZKB_ERR_INVALID_ARG;
  if (!out || (!k_dev && n)) ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "bases_generate: bad argument");
  *out = nullptr;
  ZKB_ON_DEVICE(ctx);
  ZKB_TRY(ensure_fixed_table<F>(ctx));
  H* h = nullptr;
  ZKB_TRY(bases_alloc<F>(ctx, n, &h));
  if (n) {
    fixed_base_mul_kernel<F><<<blocks_for(n, 64), 64, 0, ctx->stream>>>(
        static_cast<const Affine<F>*>(*Generator<F>::slot(ctx)), static_cast<const uint32_t*>(k_dev), n, h->p);
    ctx->launches++;
    cudaError_t e2 = cudaGetLastError();
    if (e2 != cudaSuccess) {
      cudaFree(h->p);
      cudaFree(h->inf_mask);
      delete h;
      ZKB_FAIL(ctx, ZKB_ERR_CUDA, "fixed_base_mul_kernel: %s", cudaGetErrorString(e2));
    }
    int s = build_window_tables<F>(ctx, h);
    if (s != ZKB_OK) {
      cudaFree(h->p);
      cudaFree(h->inf_mask);
      delete h;
      return s;
    }
  }
  *out = h;
  return ZKB_OK;
}

template <class F>
int bases_read_impl(zkb_ctx* ctx, const typename GroupOf<F>::Bases* b, size_t offset, size_t n, uint8_t* out_host) {
  if (!ctx) return ZKB_ERR_INVALID_ARG;
  if (!b || !out_host || offset + n > b->n) ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "bases_read: bad range");
  if (n == 0) return ZKB_OK;
  ZKB_ON_DEVICE(ctx);
  CUDA_TRY(ctx, ctx->tmp0.reserve(n * sizeof(Affine<F>)));
  affine_export_kernel<F><<<blocks_for(n, 128), 128, 0, ctx->stream>>>(b->p + offset, ctx->tmp0.as<uint32_t>(), n);
  ctx->launches++;
  CUDA_TRY(ctx, cudaGetLastError());
  CUDA_TRY(ctx, cudaMemcpyAsync(out_host, ctx->tmp0.p, n * sizeof(Affine<F>), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return ZKB_OK;
}

template <class F>
void bases_free_impl(typename GroupOf<F>::Bases* b) {
  if (!b) return;
  DeviceGuard dg(b->device);
  g_alloc_epoch.fetch_add(1, std::memory_order_relaxed);  // captured prove graphs hold these pointers
  if (b->p) cudaFree(b->p);
  if (b->inf_mask) cudaFree(b->inf_mask);
  if (b->comb) cudaFree(b->comb);
  delete b;
}

template <class F>
int msm_dev_impl(zkb_ctx* ctx, const typename GroupOf<F>::Bases* bases, size_t offset, const void* scalars_dev, size_t n,
                 void* out_affine_dev, void* out_partial_dev) {
  if (!ctx) return ZKB_ERR_INVALID_ARG;
  if (!bases || offset + n > bases->n || (!scalars_dev && n)) ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "msm: bad bases range or scalars");
  if (bases->device != ctx->device) ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "msm: bases live on device %d, ctx on %d", bases->device, ctx->device);
  ZKB_ON_DEVICE(ctx);
  cudaError_t e = msm_run<F>(ctx, bases->p, bases->n, bases->inf_mask, bases->c, bases->nwin, offset, static_cast<const uint32_t*>(scalars_dev), n,
                             static_cast<XYZZ<F>*>(out_partial_dev), static_cast<uint32_t*>(out_affine_dev));
  if (e != cudaSuccess) {
    cudaGetLastError();
    ZKB_FAIL(ctx, e == cudaErrorMemoryAllocation ? ZKB_ERR_OOM : ZKB_ERR_CUDA, "msm_run: %s", cudaGetErrorString(e));
  }
  return ZKB_OK;
}

// `batch` scalar vectors (vector p at scalars_dev + p * stride * 32 bytes) against bases [offset, offset + n): one sort /
// accumulation / reduction for all (msm_run_batch).  out_partial_dev: batch x XYZZ; out_affine_dev: batch x canonical affine.
template <class F>
int msm_batch_dev_impl(zkb_ctx* ctx, const typename GroupOf<F>::Bases* bases, size_t offset, const void* scalars_dev, size_t n,
                       size_t stride, int batch, void* out_affine_dev, void* out_partial_dev) {
  if (!ctx) return ZKB_ERR_INVALID_ARG;
  if (!bases || offset + n > bases->n || (!scalars_dev && n) || batch < 1 || stride < n)
    ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "msm_batch: bad bases range, scalars or batch");
  if (bases->device != ctx->device) ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "msm_batch: bases live on device %d, ctx on %d", bases->device, ctx->device);
  ZKB_ON_DEVICE(ctx);
  cudaError_t e = msm_run_batch<F>(ctx, bases->p, bases->n, bases->inf_mask, bases->c, bases->nwin, offset,
                                   static_cast<const uint32_t*>(scalars_dev), n, stride, batch,
                                   static_cast<XYZZ<F>*>(out_partial_dev), static_cast<uint32_t*>(out_affine_dev));
  if (e != cudaSuccess) {
    cudaGetLastError();
    ZKB_FAIL(ctx, e == cudaErrorMemoryAllocation ? ZKB_ERR_OOM : ZKB_ERR_CUDA, "msm_run_batch: %s", cudaGetErrorString(e));
  }
  return ZKB_OK;
}

// parity hook for the entry sort (sort.cuh): the sorted (key, value) lists of a batched MSM front end, copied to out_keys /
// out_vals (device, capacity nwin * n * batch words each); *out_count (device) = their number.
template <class F>
int msm_entries_debug_impl(zkb_ctx* ctx, const typename GroupOf<F>::Bases* bases, size_t offset, const void* scalars_dev, size_t n,
                           size_t stride, int batch, void* out_keys, void* out_vals, void* out_count) {
  if (!ctx) return ZKB_ERR_INVALID_ARG;
  if (!bases || offset + n > bases->n || !scalars_dev || batch < 1 || n == 0) ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "msm_entries_debug: bad argument");
  ZKB_ON_DEVICE(ctx);
  MsmLayout<F> L = msm_layout<F>(ctx->sm_count, bases->c, bases->nwin, n, 1, batch);
  CUDA_TRY(ctx, ctx->msm_ws.reserve(L.bytes));
  char* base = static_cast<char*>(ctx->msm_ws.p);
  uint32_t* hdr = (uint32_t*)(base + L.o_hdr);
  const uint32_t *sk = nullptr, *sv = nullptr;
  EntrySource src{static_cast<const uint32_t*>(scalars_dev), n, stride, batch, L.c, L.nwin, L.nbuck, bases->n, offset, bases->inf_mask};
  cudaError_t e = msm_sort_entries(src, L.sort, ctx->sm_count, hdr, (uint32_t*)(base + L.o_k0), (uint32_t*)(base + L.o_v0),
                                   (uint32_t*)(base + L.o_k1), (uint32_t*)(base + L.o_v1), ctx->stream, &sk, &sv, &ctx->launches);
  if (e != cudaSuccess) ZKB_FAIL(ctx, ZKB_ERR_CUDA, "msm_sort_entries: %s", cudaGetErrorString(e));
  CUDA_TRY(ctx, cudaMemcpyAsync(out_keys, sk, L.total * 4, cudaMemcpyDeviceToDevice, ctx->stream));
  CUDA_TRY(ctx, cudaMemcpyAsync(out_vals, sv, L.total * 4, cudaMemcpyDeviceToDevice, ctx->stream));
  CUDA_TRY(ctx, cudaMemcpyAsync(out_count, hdr + SortHeader::TOTAL, 4, cudaMemcpyDeviceToDevice, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return ZKB_OK;
}

template <class F>
size_t bases_comb_bytes(const typename GroupOf<F>::Bases* b, int c) {
  return b ? b->n * size_t(msm_windows_for(c)) * (size_t(1) << (c - 1)) * sizeof(Affine<F>) : 0;
}

// Builds the comb table for window width c.  The window tables 2^(c w) P_i must exist for the same c: when the handle was
// built with another width they are rebuilt here into a temporary.
template <class F>
int bases_build_comb(zkb_ctx* ctx, typename GroupOf<F>::Bases* b, int c) {
  if (!b || c < 2 || c > 16) ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "bases_build_comb: bad argument");
  if (b->comb && b->comb_c == c) return ZKB_OK;
  if (b->n == 0) return ZKB_OK;
  ZKB_ON_DEVICE(ctx);
  const int nwin = msm_windows_for(c);
  Affine<F>* wt = nullptr;     // window tables for width c
  bool own_wt = false;
  if (b->c == c) {
    wt = b->p;
  } else {
    CUDA_TRY(ctx, cudaMalloc(&wt, size_t(nwin) * b->n * sizeof(Affine<F>)));
    own_wt = true;
    CUDA_TRY(ctx, cudaMemcpyAsync(wt, b->p, b->n * sizeof(Affine<F>), cudaMemcpyDeviceToDevice, ctx->stream));
    if (nwin > 1) window_tables_kernel<F><<<blocks_for(b->n, 128), 128, 0, ctx->stream>>>(wt, b->n, c, nwin);
    ctx->launches++;
  }
  Affine<F>* comb = nullptr;
  cudaError_t e = cudaMalloc(&comb, bases_comb_bytes<F>(b, c));
  if (e != cudaSuccess) {
    cudaGetLastError();
    if (own_wt) cudaFree(wt);
    ZKB_FAIL(ctx, ZKB_ERR_OOM, "bases_build_comb: cudaMalloc(%zu) failed", bases_comb_bytes<F>(b, c));
  }
  comb_build_kernel<F><<<blocks_for(b->n * size_t(nwin), 64), 64, 0, ctx->stream>>>(wt, b->n, nwin, c, comb);
  ctx->launches++;
  e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (own_wt) cudaFree(wt);
  if (e != cudaSuccess) {
    cudaFree(comb);
    ZKB_FAIL(ctx, ZKB_ERR_CUDA, "comb_build_kernel: %s", cudaGetErrorString(e));
  }
  if (b->comb) cudaFree(b->comb);
  b->comb = comb;
  b->comb_c = c;
  b->comb_nwin = nwin;
  return ZKB_OK;
}

template <class F>
int msm_comb_dev_impl(zkb_ctx* ctx, const typename GroupOf<F>::Bases* bases, size_t offset, const void* scalars_dev, size_t n,
                      size_t stride, int batch, void* out_partial_dev, void* out_affine_dev) {
  if (!ctx) return ZKB_ERR_INVALID_ARG;
  if (!bases || !bases->comb || offset + n > bases->n || (!scalars_dev && n) || batch < 1 || stride < n || !out_partial_dev)
    ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "msm_comb: bad bases range, scalars or batch (or no comb table)");
  if (bases->device != ctx->device) ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "msm_comb: bases live on device %d, ctx on %d", bases->device, ctx->device);
  ZKB_ON_DEVICE(ctx);
  cudaError_t e = msm_run_comb<F>(ctx, bases->comb, bases->comb_c, bases->comb_nwin, bases->inf_mask, offset,
                                  static_cast<const uint32_t*>(scalars_dev), n, stride, batch, static_cast<XYZZ<F>*>(out_partial_dev));
  if (e != cudaSuccess) {
    cudaGetLastError();
    ZKB_FAIL(ctx, e == cudaErrorMemoryAllocation ? ZKB_ERR_OOM : ZKB_ERR_CUDA, "msm_run_comb: %s", cudaGetErrorString(e));
  }
  if (out_affine_dev) {
    xyzz_to_affine_bytes_kernel<F><<<blocks_for(size_t(batch), 64), 64, 0, ctx->stream>>>(static_cast<const XYZZ<F>*>(out_partial_dev),
                                                                                       size_t(batch), static_cast<uint32_t*>(out_affine_dev));
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
  }
  return ZKB_OK;
}

// 32 x 255 fixed-base table (8-bit windows) of bases->p[idx]: delta_g1 / delta_g2 of a key, for the per-proof r, s multiples
template <class F>
int fixed_table_for_base(zkb_ctx* ctx, const typename GroupOf<F>::Bases* bases, size_t idx, void** out_table) {
  if (!bases || idx >= bases->n) ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "fixed_table_for_base: bad index");
  Affine<F>* table = nullptr;
  CUDA_TRY(ctx, cudaMalloc(&table, 32 * 255 * sizeof(Affine<F>)));
  fixed_base_table_dev_kernel<F><<<blocks_for(32 * 255, 64), 64, 0, ctx->stream>>>(bases->p + idx, table);
  ctx->launches++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    cudaFree(table);
    ZKB_FAIL(ctx, ZKB_ERR_CUDA, "fixed_base_table_dev_kernel: %s", cudaGetErrorString(e));
  }
  *out_table = table;
  return ZKB_OK;
}

// Host scalars -> queued MSM.  out_affine_dev / out_partial_dev: device memory, either may be null.  Large inputs upload in
// slices on a second stream and accumulate each slice while the next one is in flight (msm_run_host_sliced).
template <class F>
int msm_host_enqueue(zkb_ctx* ctx, const typename GroupOf<F>::Bases* bases, size_t offset, const uint8_t* scalars_host, size_t n,
                     void* out_affine_dev, void* out_partial_dev) {
  CUDA_TRY(ctx, ctx->scal.reserve(n * 32 + 32));
  if (n >= (size_t(1) << 20) && ctx->msm_slices > 1) {
    if (!bases || offset + n > bases->n) ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "msm: bad bases range or scalars");
    if (bases->device != ctx->device) ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "msm: bases live on device %d, ctx on %d", bases->device, ctx->device);
    cudaError_t e = msm_run_host_sliced<F>(ctx, bases->p, bases->n, bases->inf_mask, bases->c, bases->nwin, offset, scalars_host,
                                           ctx->scal.as<uint32_t>(), n, ctx->msm_slices, static_cast<XYZZ<F>*>(out_partial_dev),
                                           static_cast<uint32_t*>(out_affine_dev));
    if (e != cudaSuccess) {
      cudaGetLastError();
      ZKB_FAIL(ctx, e == cudaErrorMemoryAllocation ? ZKB_ERR_OOM : ZKB_ERR_CUDA, "msm_run_host_sliced: %s", cudaGetErrorString(e));
    }
    return ZKB_OK;
  }
  if (n) CUDA_TRY(ctx, cudaMemcpyAsync(ctx->scal.p, scalars_host, n * 32, cudaMemcpyHostToDevice, ctx->stream));
  return msm_dev_impl<F>(ctx, bases, offset, ctx->scal.p, n, out_affine_dev, out_partial_dev);
}

template <class F>
int msm_host_impl(zkb_ctx* ctx, const typename GroupOf<F>::Bases* bases, size_t offset, const uint8_t* scalars_host, size_t n,
                  uint8_t* out) {
  if (!ctx) return ZKB_ERR_INVALID_ARG;
  if (!out || (!scalars_host && n)) ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "msm: null argument");
  ZKB_ON_DEVICE(ctx);
  CUDA_TRY(ctx, ctx->res.reserve(512));
  ZKB_TRY((msm_host_enqueue<F>(ctx, bases, offset, scalars_host, n, ctx->res.p, nullptr)));
  CUDA_TRY(ctx, cudaMemcpyAsync(out, ctx->res.p, sizeof(Affine<F>), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return ZKB_OK;
}

// One rank's share of a range-sharded MSM from HOST scalars: the sliced upload pipeline, result = the projective partial sum
// in device memory (for an all-gather) -- asynchronous on ctx's stream.
template <class F>
int msm_host_partial_impl(zkb_ctx* ctx, const typename GroupOf<F>::Bases* bases, size_t offset, const uint8_t* scalars_host, size_t n,
                          void* out_partial_dev) {
  if (!ctx) return ZKB_ERR_INVALID_ARG;
  if (!out_partial_dev || (!scalars_host && n)) ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "msm_partial: null argument");
  ZKB_ON_DEVICE(ctx);
  return msm_host_enqueue<F>(ctx, bases, offset, scalars_host, n, nullptr, out_partial_dev);
}

// One process, one context per GPU: GPU i holds bases[i] = range i of the points (in order); the n host scalars are split the
// same way.  Every GPU runs its partial MSM (own host thread, sliced upload), the n_gpus partial sums (128 / 256 B each) are
// gathered through host memory and added on ctxs[0].
template <class F>
int msm_multi_impl(zkb_ctx* const* ctxs, const typename GroupOf<F>::Bases* const* bases, int n_gpus, const uint8_t* scalars_host,
                   size_t n, uint8_t* out) {
  if (!ctxs || n_gpus < 1 || !ctxs[0]) return ZKB_ERR_INVALID_ARG;
  zkb_ctx* c0 = ctxs[0];
  if (!bases || !out || (!scalars_host && n) || n_gpus > 64) ZKB_FAIL(c0, ZKB_ERR_INVALID_ARG, "msm_multi: bad argument");
  size_t total = 0;
  for (int i = 0; i < n_gpus; i++) {
    if (!ctxs[i] || !bases[i]) ZKB_FAIL(c0, ZKB_ERR_INVALID_ARG, "msm_multi: null context or bases for GPU %d", i);
    total += bases[i]->n;
  }
  if (total != n) ZKB_FAIL(c0, ZKB_ERR_SHAPE, "msm_multi: the bases handles hold %zu points, %zu scalars given", total, n);
  constexpr size_t PB = sizeof(XYZZ<F>);
  std::vector<uint8_t> parts(size_t(n_gpus) * PB);
  std::vector<int> rc(size_t(n_gpus), ZKB_OK);
  auto work = [&](int i, size_t off) {
    zkb_ctx* ctx = ctxs[i];
    rc[size_t(i)] = [&]() -> int {
      ZKB_ON_DEVICE(ctx);
      CUDA_TRY(ctx, ctx->res.reserve(512));
      ZKB_TRY((msm_host_enqueue<F>(ctx, bases[i], 0, scalars_host + off * 32, bases[i]->n, nullptr, ctx->res.p)));
      CUDA_TRY(ctx, cudaMemcpyAsync(parts.data() + size_t(i) * PB, ctx->res.p, PB, cudaMemcpyDeviceToHost, ctx->stream));
      CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
      return ZKB_OK;
    }();
  };
  try {
    struct Joiner {
      std::vector<std::thread> th;
      ~Joiner() {
        for (auto& t : th)
          if (t.joinable()) t.join();
      }
    } jn;
    size_t off = bases[0]->n;
    for (int i = 1; i < n_gpus; i++) {
      jn.th.emplace_back(work, i, off);
      off += bases[i]->n;
    }
    work(0, 0);
  } catch (...) {
    ZKB_FAIL(c0, ZKB_ERR_OOM, "msm_multi: could not start the per-GPU host threads");
  }
  for (int i = 0; i < n_gpus; i++)
    if (rc[size_t(i)] != ZKB_OK) {
      if (i) c0->err = ctxs[i]->err;
      return rc[size_t(i)];
    }
  ZKB_ON_DEVICE(c0);
  CUDA_TRY(c0, c0->tmp1.reserve(parts.size()));
  CUDA_TRY(c0, c0->res.reserve(512));
  CUDA_TRY(c0, cudaMemcpyAsync(c0->tmp1.p, parts.data(), parts.size(), cudaMemcpyHostToDevice, c0->stream));
  ZKB_TRY((msm_combine_impl<F>(c0, c0->tmp1.p, n_gpus, c0->res.p)));
  CUDA_TRY(c0, cudaMemcpyAsync(out, c0->res.p, sizeof(Affine<F>), cudaMemcpyDeviceToHost, c0->stream));
  CUDA_TRY(c0, cudaStreamSynchronize(c0->stream));
  return ZKB_OK;
}

template <class F>
int msm_combine_impl(zkb_ctx* ctx, const void* parts, int k, void* out) {
  if (!ctx) return ZKB_ERR_INVALID_ARG;
  if (!parts || k <= 0 || !out) ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "msm_combine: bad argument");
  ZKB_ON_DEVICE(ctx);
  msm_combine_kernel<F><<<1, 32, 0, ctx->stream>>>(static_cast<const XYZZ<F>*>(parts), k, static_cast<uint32_t*>(out));
  ctx->launches++;
  CUDA_TRY(ctx, cudaGetLastError());
  return ZKB_OK;
}

#define ZKB_INSTANTIATE_GROUP(F)                                                                                            \
  template int import_points<F>(zkb_ctx*, const uint8_t*, size_t, int, Affine<F>*);                                        \
  template int bases_load_impl<F>(zkb_ctx*, const uint8_t*, size_t, int, GroupOf<F>::Bases**);                              \
  template int bases_load_compressed_impl<F>(zkb_ctx*, const uint8_t*, size_t, int, GroupOf<F>::Bases**);                   \
  template int bases_generate_impl<F>(zkb_ctx*, const void*, size_t, GroupOf<F>::Bases**);                                  \
  template int bases_read_impl<F>(zkb_ctx*, const GroupOf<F>::Bases*, size_t, size_t, uint8_t*);                            \
  template void bases_free_impl<F>(GroupOf<F>::Bases*);                                                                     \
  template int scalar_mul_impl<F>(zkb_ctx*, const uint8_t*, const uint8_t*, size_t, uint8_t*);                              \
  template int point_sum_impl<F>(zkb_ctx*, const uint8_t*, size_t, uint8_t*);                                               \
  template int msm_dev_impl<F>(zkb_ctx*, const GroupOf<F>::Bases*, size_t, const void*, size_t, void*, void*);              \
  template int msm_host_impl<F>(zkb_ctx*, const GroupOf<F>::Bases*, size_t, const uint8_t*, size_t, uint8_t*);              \
  template int msm_combine_impl<F>(zkb_ctx*, const void*, int, void*);                                                      \
  template int msm_host_partial_impl<F>(zkb_ctx*, const GroupOf<F>::Bases*, size_t, const uint8_t*, size_t, void*);          \
  template int msm_multi_impl<F>(zkb_ctx* const*, const GroupOf<F>::Bases* const*, int, const uint8_t*, size_t, uint8_t*);   \
  template int msm_batch_dev_impl<F>(zkb_ctx*, const GroupOf<F>::Bases*, size_t, const void*, size_t, size_t, int, void*, void*); \
  template int fixed_table_for_base<F>(zkb_ctx*, const GroupOf<F>::Bases*, size_t, void**);                                  \
  template int bases_build_comb<F>(zkb_ctx*, GroupOf<F>::Bases*, int);                                                        \
  template size_t bases_comb_bytes<F>(const GroupOf<F>::Bases*, int);                                                         \
  template int msm_comb_dev_impl<F>(zkb_ctx*, const GroupOf<F>::Bases*, size_t, const void*, size_t, size_t, int, void*, void*); \
  template int msm_entries_debug_impl<F>(zkb_ctx*, const GroupOf<F>::Bases*, size_t, const void*, size_t, size_t, int, void*, void*, void*); \
  template void fixed_table_free<F>(zkb_ctx*);                                                                             \
  template int fixed_base_batch<F>(zkb_ctx*, const uint8_t*, const void*, size_t, uint8_t*);

}  // namespace zkb
