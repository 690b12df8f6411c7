// C ABI of the B200-native Groth16/BN254 backend (declared in include/zkb200.h).
// Everything mathematical runs in the CUDA kernels of fp.cuh / ec.cuh / msm.cuh / ntt.cuh and the
// small kernels below; the host code here only moves bytes, sizes workspaces and orders launches.
#include "../../include/zkb200.h"

#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <map>
#include <new>
#include <string>
#include <vector>

#include "ec.cuh"
#include "msm.cuh"
#include "ntt.cuh"

using namespace zkb;

// =============================================================================================== helpers
namespace {

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    if (bytes == 0) return cudaSuccess;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e == cudaSuccess) cap = bytes;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <class T>
  T* as() const { return static_cast<T*>(p); }
};

struct NttTables {
  Fr* mem = nullptr;  // one allocation
  PowTable fwd, inv, coset_pre, coset_inv_post;
  const Fr* ninv = nullptr;
};

constexpr int NTT_LRMAX = 8;

}  // namespace

struct zkb_ctx {
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  std::string err;
  LaunchCounter lc;
  int msm_c = 0;
  MsmWorkspace msm_ws;
  DevBuf scal, res, tmp0, tmp1, tmp2, flag;
  Fr* wr_fwd = nullptr;  // omega_(2^LRMAX)^e
  Fr* wr_inv = nullptr;
  std::map<int, NttTables> ntt_tables;
  Affine<Fq>* g1_table = nullptr;   // 32 windows x 255 multiples of G
  Affine<Fq2>* g2_table = nullptr;
};

struct zkb_g1_bases {
  int device;
  Affine<Fq>* p;
  size_t n;
};
struct zkb_g2_bases {
  int device;
  Affine<Fq2>* p;
  size_t n;
};

struct CsrDev {
  uint64_t* row_ptr = nullptr;
  uint32_t* col = nullptr;
  Fr* coeff = nullptr;  // Montgomery
  size_t nnz = 0;
};
struct zkb_r1cs {
  int device;
  uint64_t nc, ni, nw;
  int log_domain;
  CsrDev a, b, c;
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
};

#define ZKB_FAIL(ctx, code, ...)                         \
  do {                                                   \
    char _b[512];                                        \
    snprintf(_b, sizeof(_b), __VA_ARGS__);               \
    (ctx)->err = _b;                                     \
    return (code);                                       \
  } while (0)

#define CUDA_TRY(ctx, expr)                                                                            \
  do {                                                                                                 \
    cudaError_t _e = (expr);                                                                           \
    if (_e != cudaSuccess) {                                                                           \
      cudaGetLastError();                                                                              \
      ZKB_FAIL(ctx, _e == cudaErrorMemoryAllocation ? ZKB_ERR_OOM : ZKB_ERR_CUDA, "%s: %s (%s:%d)", #expr, \
               cudaGetErrorString(_e), __FILE__, __LINE__);                                            \
    }                                                                                                  \
  } while (0)

#define ZKB_TRY(expr)          \
  do {                         \
    int _s = (expr);           \
    if (_s != ZKB_OK) return _s; \
  } while (0)

// =============================================================================================== small kernels
namespace {

template <class F>
__device__ __forceinline__ bool fp_is_canonical(const F& a) {
  F m = F::modulus();
  for (int i = 7; i >= 0; i--) {
    if (a.v[i] < m.v[i]) return true;
    if (a.v[i] > m.v[i]) return false;
  }
  return false;  // equal to the modulus
}

template <class F>
__global__ void field_op_kernel(int op, const F* a, const F* b, size_t n, F* out, int* bad) {
  size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  F x = a[i];
  if (!fp_is_canonical(x)) { atomicExch(bad, 1); return; }
  F y = F::zero();
  if (op <= 2) {
    y = b[i];
    if (!fp_is_canonical(y)) { atomicExch(bad, 1); return; }
  }
  F r;
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
template <class F>
__global__ void to_mont_kernel(const F* in, F* out, size_t n, int* bad) {
  size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  F x = in[i];
  if (!fp_is_canonical(x)) { atomicExch(bad, 1); return; }
  out[i] = x.to_mont();
}

template <class F>
__global__ void from_mont_kernel(const F* in, F* out, size_t n) {
  size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = in[i].from_mont();
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
// b' = 3/(9+u) in Montgomery form (oracle/bn254.py B_G2)
__device__ const uint32_t G2_B_MONT[16] = {0x77b802a8u, 0x3bf938e3u, 0x3633535du, 0x020b1b27u, 0x49755260u, 0x26b7edf0u,
                                           0x4384a86du, 0x2514c632u, 0xd1dcff67u, 0x38e7ecccu, 0x93ce0d3eu, 0x65f0b37du,
                                           0x22ac00aau, 0xd749d0ddu, 0x4a688d4du, 0x0141b9ceu};
template <> struct CurveB<Fq2> {
  static __device__ __forceinline__ Fq2 b() {
    Fq2 r;
    for (int i = 0; i < 8; i++) { r.c0.v[i] = G2_B_MONT[i]; r.c1.v[i] = G2_B_MONT[8 + i]; }
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

// canonical affine bytes -> Montgomery affine, optional on-curve validation
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
  }
  out[i] = p;
}

template <class F>
__global__ void affine_export_kernel(const Affine<F>* in, uint32_t* out, size_t n) {
  size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  store_affine_canonical<F>(in[i], out + i * (sizeof(Affine<F>) / 4));
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

// ---- R1CS ---------------------------------------------------------------------------------------
// out[i] = <row_i, z> for i < nc ; optionally out[nc + j] = z[j] for j < ni ; zero up to n
__global__ void csr_matvec_kernel(const uint64_t* __restrict__ row_ptr, const uint32_t* __restrict__ col,
                                  const Fr* __restrict__ coeff, const Fr* __restrict__ z, uint64_t nc, uint64_t ni,
                                  int append_instance, size_t n, Fr* __restrict__ out) {
  size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fr acc = Fr::zero();
  if (i < nc) {
    for (uint64_t k = row_ptr[i]; k < row_ptr[i + 1]; k++) acc = acc + coeff[k] * z[col[k]];
  } else if (append_instance && i < nc + ni) {
    acc = z[i - nc];
  }
  out[i] = acc;
}

// ab[i] = (a[i] * b[i] - c[i]) * zinv ; a is in Montgomery form, b and c canonical, zinv Montgomery -> canonical
__global__ void qap_pointwise_kernel(const Fr* __restrict__ a, const Fr* __restrict__ b, const Fr* __restrict__ c,
                                     int logn, size_t n, Fr* __restrict__ out) {
  size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  __shared__ Fr zinv_sh;
  if (threadIdx.x == 0) {
    // (g^n - 1)^-1, g = 5
    Fr gn = fr_base(FRB_GEN);
    for (int k = 0; k < logn; k++) gn = gn.sqr();
    zinv_sh = (gn - Fr::one()).inverse();
  }
  __syncthreads();
  if (i >= n) return;
  out[i] = (a[i] * b[i] - c[i]) * zinv_sh;
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

// C = s*A + r*B1 + L + H   (L already contains -(r s) delta_1)
__global__ void prove_assemble_c_kernel(const XYZZ<Fq>* A, const XYZZ<Fq>* B1, const XYZZ<Fq>* L, const XYZZ<Fq>* H,
                                        const uint32_t* r, const uint32_t* s, uint32_t* out_c) {
  __shared__ XYZZ<Fq> part[2];
  int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  if (lane == 0 && warp < 2) {
    uint32_t k[8];
    for (int j = 0; j < 8; j++) k[j] = warp == 0 ? s[j] : r[j];
    part[warp] = (warp == 0 ? *A : *B1).mul_words(k);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    XYZZ<Fq> acc = part[0];
    acc.add(part[1]);
    acc.add(*L);
    acc.add(*H);
    store_affine_canonical<Fq>(acc.to_affine(), out_c);
  }
}

inline unsigned blocks_for(size_t n, int threads) { return unsigned((n + threads - 1) / threads); }

int check_flag(zkb_ctx* ctx, const char* what) {
  int h = 0;
  CUDA_TRY(ctx, cudaMemcpyAsync(&h, ctx->flag.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  if (h == 1) ZKB_FAIL(ctx, ZKB_ERR_NOT_CANONICAL, "%s: field element >= modulus", what);
  if (h == 2) ZKB_FAIL(ctx, ZKB_ERR_NOT_CANONICAL, "%s: point not on curve", what);
  return ZKB_OK;
}

int clear_flag(zkb_ctx* ctx) {
  CUDA_TRY(ctx, ctx->flag.reserve(sizeof(int)));
  CUDA_TRY(ctx, cudaMemsetAsync(ctx->flag.p, 0, sizeof(int), ctx->stream));
  return ZKB_OK;
}

int set_device(zkb_ctx* ctx) {
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  return ZKB_OK;
}

}  // namespace

// =============================================================================================== context
extern "C" const char* zkb_version(void) { return "zkb200 0.1 (sm_100a)"; }

extern "C" int zkb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

extern "C" int zkb_ctx_create(int device, zkb_ctx** out) {
  if (!out) return ZKB_ERR_INVALID_ARG;
  *out = nullptr;
  int n = zkb_device_count();
  if (n <= 0 || device < 0 || device >= n) return ZKB_ERR_NO_DEVICE;
  zkb_ctx* ctx = new (std::nothrow) zkb_ctx();
  if (!ctx) return ZKB_ERR_OOM;
  ctx->device = device;
  if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
    cudaGetLastError();
    delete ctx;
    return ZKB_ERR_CUDA;
  }
  ctx->own_stream = true;
  cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
  *out = ctx;
  return ZKB_OK;
}

extern "C" int zkb_ctx_set_stream(zkb_ctx* ctx, void* cuda_stream) {
  if (!ctx) return ZKB_ERR_INVALID_ARG;
  ZKB_TRY(set_device(ctx));
  if (ctx->own_stream && ctx->stream) {
    cudaStreamSynchronize(ctx->stream);
    cudaStreamDestroy(ctx->stream);
  }
  if (cuda_stream) {
    ctx->stream = static_cast<cudaStream_t>(cuda_stream);
    ctx->own_stream = false;
  } else {
    CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    ctx->own_stream = true;
  }
  return ZKB_OK;
}

extern "C" int zkb_ctx_synchronize(zkb_ctx* ctx) {
  if (!ctx) return ZKB_ERR_INVALID_ARG;
  ZKB_TRY(set_device(ctx));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return ZKB_OK;
}

extern "C" void zkb_ctx_destroy(zkb_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  ctx->msm_ws.release();
  ctx->scal.release(); ctx->res.release(); ctx->tmp0.release(); ctx->tmp1.release(); ctx->tmp2.release();
  ctx->flag.release();
  if (ctx->wr_fwd) cudaFree(ctx->wr_fwd);
  if (ctx->wr_inv) cudaFree(ctx->wr_inv);
  for (auto& kv : ctx->ntt_tables) cudaFree(kv.second.mem);
  if (ctx->g1_table) cudaFree(ctx->g1_table);
  if (ctx->g2_table) cudaFree(ctx->g2_table);
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

extern "C" const char* zkb_last_error(zkb_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
extern "C" unsigned long long zkb_launch_count(zkb_ctx* ctx) { return ctx ? ctx->lc.n : 0; }
extern "C" int zkb_ctx_set_msm_window(zkb_ctx* ctx, int c) {
  if (!ctx || c < 0 || c > 16 || (c > 0 && c < 2)) return ZKB_ERR_INVALID_ARG;
  ctx->msm_c = c;
  return ZKB_OK;
}

// =============================================================================================== field / curve hooks
extern "C" int zkb_field_op(zkb_ctx* ctx, int field, int op, const uint8_t* a, const uint8_t* b, size_t n, uint8_t* out) {
  if (!ctx) return ZKB_ERR_INVALID_ARG;
  if (field < 0 || field > 1 || op < 0 || op > 4 || !a || !out || (op <= 2 && !b)) ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "zkb_field_op: bad argument");
  if (n == 0) return ZKB_OK;
  ZKB_TRY(set_device(ctx));
  size_t bytes = n * 32;
  CUDA_TRY(ctx, ctx->tmp0.reserve(bytes));
  CUDA_TRY(ctx, ctx->tmp1.reserve(bytes));
  CUDA_TRY(ctx, ctx->tmp2.reserve(bytes));
  ZKB_TRY(clear_flag(ctx));
  CUDA_TRY(ctx, cudaMemcpyAsync(ctx->tmp0.p, a, bytes, cudaMemcpyHostToDevice, ctx->stream));
  if (op <= 2) CUDA_TRY(ctx, cudaMemcpyAsync(ctx->tmp1.p, b, bytes, cudaMemcpyHostToDevice, ctx->stream));
  if (field == 0)
    field_op_kernel<Fr><<<blocks_for(n, 128), 128, 0, ctx->stream>>>(op, ctx->tmp0.as<Fr>(), ctx->tmp1.as<Fr>(), n, ctx->tmp2.as<Fr>(), ctx->flag.as<int>());
  else
    field_op_kernel<Fq><<<blocks_for(n, 128), 128, 0, ctx->stream>>>(op, ctx->tmp0.as<Fq>(), ctx->tmp1.as<Fq>(), n, ctx->tmp2.as<Fq>(), ctx->flag.as<int>());
  ctx->lc.n++;
  CUDA_TRY(ctx, cudaGetLastError());
  CUDA_TRY(ctx, cudaMemcpyAsync(out, ctx->tmp2.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  return check_flag(ctx, "zkb_field_op");
}

namespace {

template <class F>
int import_points(zkb_ctx* ctx, const uint8_t* host, size_t n, int validate, Affine<F>* dst) {
  if (n == 0) return ZKB_OK;
  size_t bytes = n * sizeof(Affine<F>);
  CUDA_TRY(ctx, ctx->tmp0.reserve(bytes));
  ZKB_TRY(clear_flag(ctx));
  CUDA_TRY(ctx, cudaMemcpyAsync(ctx->tmp0.p, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
  affine_import_kernel<F><<<blocks_for(n, 128), 128, 0, ctx->stream>>>(ctx->tmp0.as<uint32_t>(), dst, n, validate, ctx->flag.as<int>());
  ctx->lc.n++;
  CUDA_TRY(ctx, cudaGetLastError());
  return check_flag(ctx, "point import");
}

template <class F>
int scalar_mul_impl(zkb_ctx* ctx, const uint8_t* points, const uint8_t* scalars, size_t n, uint8_t* out) {
  if (n == 0) return ZKB_OK;
  size_t pbytes = n * sizeof(Affine<F>);
  CUDA_TRY(ctx, ctx->tmp1.reserve(pbytes));
  ZKB_TRY(import_points<F>(ctx, points, n, 1, ctx->tmp1.as<Affine<F>>()));
  CUDA_TRY(ctx, ctx->scal.reserve(n * 32));
  CUDA_TRY(ctx, ctx->tmp2.reserve(pbytes));
  CUDA_TRY(ctx, cudaMemcpyAsync(ctx->scal.p, scalars, n * 32, cudaMemcpyHostToDevice, ctx->stream));
  scalar_mul_kernel<F><<<blocks_for(n, 64), 64, 0, ctx->stream>>>(ctx->tmp1.as<Affine<F>>(), ctx->scal.as<uint32_t>(), n, ctx->tmp2.as<uint32_t>());
  ctx->lc.n++;
  CUDA_TRY(ctx, cudaGetLastError());
  CUDA_TRY(ctx, cudaMemcpyAsync(out, ctx->tmp2.p, pbytes, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return ZKB_OK;
}

template <class F>
int point_sum_impl(zkb_ctx* ctx, const uint8_t* points, size_t n, uint8_t* out) {
  size_t pbytes = n * sizeof(Affine<F>);
  CUDA_TRY(ctx, ctx->tmp1.reserve(pbytes + sizeof(Affine<F>)));
  ZKB_TRY(import_points<F>(ctx, points, n, 1, ctx->tmp1.as<Affine<F>>()));
  CUDA_TRY(ctx, ctx->res.reserve(256));
  point_sum_kernel<F><<<1, 32, 0, ctx->stream>>>(ctx->tmp1.as<Affine<F>>(), n, ctx->res.as<uint32_t>());
  ctx->lc.n++;
  CUDA_TRY(ctx, cudaGetLastError());
  CUDA_TRY(ctx, cudaMemcpyAsync(out, ctx->res.p, sizeof(Affine<F>), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return ZKB_OK;
}

}  // namespace

extern "C" int zkb_scalar_mul(zkb_ctx* ctx, int group, const uint8_t* points, const uint8_t* scalars, size_t n, uint8_t* out) {
  if (!ctx) return ZKB_ERR_INVALID_ARG;
  if ((group != 1 && group != 2) || !points || !scalars || !out) ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "zkb_scalar_mul: bad argument");
  ZKB_TRY(set_device(ctx));
  return group == 1 ? scalar_mul_impl<Fq>(ctx, points, scalars, n, out) : scalar_mul_impl<Fq2>(ctx, points, scalars, n, out);
}

extern "C" int zkb_point_sum(zkb_ctx* ctx, int group, const uint8_t* points, size_t n, uint8_t* out) {
  if (!ctx) return ZKB_ERR_INVALID_ARG;
  if ((group != 1 && group != 2) || (!points && n) || !out) ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "zkb_point_sum: bad argument");
  ZKB_TRY(set_device(ctx));
  return group == 1 ? point_sum_impl<Fq>(ctx, points, n, out) : point_sum_impl<Fq2>(ctx, points, n, out);
}

// =============================================================================================== bases
namespace {

template <class F, class H>
int bases_load_impl(zkb_ctx* ctx, const uint8_t* host, size_t n, int validate, H** out) {
  if (!ctx) return ZKB_ERR_INVALID_ARG;
  if (!out || (!host && n)) ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "bases_load: bad argument");
  *out = nullptr;
  ZKB_TRY(set_device(ctx));
  H* h = new (std::nothrow) H{ctx->device, nullptr, n};
  if (!h) ZKB_FAIL(ctx, ZKB_ERR_OOM, "bases_load: host allocation failed");
  if (n) {
    cudaError_t e = cudaMalloc(&h->p, n * sizeof(Affine<F>));
    if (e != cudaSuccess) {
      cudaGetLastError();
      delete h;
      ZKB_FAIL(ctx, ZKB_ERR_OOM, "bases_load: cudaMalloc(%zu) failed: %s", n * sizeof(Affine<F>), cudaGetErrorString(e));
    }
    int s = import_points<F>(ctx, host, n, validate, h->p);
    if (s != ZKB_OK) {
      cudaFree(h->p);
      delete h;
      return s;
    }
  }
  *out = h;
  return ZKB_OK;
}

// G2 generator (ark-bn254 g2::G2_GENERATOR_{X,Y}), canonical words
const uint32_t G2_GEN_CANON[32] = {
    0xd992f6edu, 0x46debd5cu, 0xf75edaddu, 0x674322d4u, 0x5e5c4479u, 0x426a0066u, 0x121f1e76u, 0x1800deefu,
    0xaef312c2u, 0x97e485b7u, 0x35a9e712u, 0xf1aa4933u, 0x31fb5d25u, 0x7260bfb7u, 0x920d483au, 0x198e9393u,
    0x66fa7daau, 0x4ce6cc01u, 0x0c43d37bu, 0xe3d1e769u, 0x8dcb408fu, 0x4aab7180u, 0xdb8c6debu, 0x12c85ea5u,
    0xd122975bu, 0x55acdadcu, 0x70b38ef3u, 0xbc4b3133u, 0x690c3395u, 0xec9e99adu, 0x585ff075u, 0x090689d0u};

template <class F>
int ensure_fixed_table(zkb_ctx* ctx, Affine<F>** slot);

template <>
int ensure_fixed_table<Fq>(zkb_ctx* ctx, Affine<Fq>** slot) {
  if (*slot) return ZKB_OK;
  CUDA_TRY(ctx, cudaMalloc(slot, 32 * 255 * sizeof(Affine<Fq>)));
  uint8_t gen[64] = {0};
  gen[0] = 1;
  gen[32] = 2;
  CUDA_TRY(ctx, ctx->tmp1.reserve(sizeof(Affine<Fq>)));
  ZKB_TRY(import_points<Fq>(ctx, gen, 1, 1, ctx->tmp1.as<Affine<Fq>>()));
  Affine<Fq> g;
  CUDA_TRY(ctx, cudaMemcpy(&g, ctx->tmp1.p, sizeof(g), cudaMemcpyDeviceToHost));
  fixed_base_table_kernel<Fq><<<blocks_for(32 * 255, 64), 64, 0, ctx->stream>>>(g, *slot);
  ctx->lc.n++;
  CUDA_TRY(ctx, cudaGetLastError());
  return ZKB_OK;
}

template <>
int ensure_fixed_table<Fq2>(zkb_ctx* ctx, Affine<Fq2>** slot) {
  if (*slot) return ZKB_OK;
  CUDA_TRY(ctx, cudaMalloc(slot, 32 * 255 * sizeof(Affine<Fq2>)));
  CUDA_TRY(ctx, ctx->tmp1.reserve(sizeof(Affine<Fq2>)));
  ZKB_TRY(import_points<Fq2>(ctx, reinterpret_cast<const uint8_t*>(G2_GEN_CANON), 1, 1, ctx->tmp1.as<Affine<Fq2>>()));
  Affine<Fq2> g;
  CUDA_TRY(ctx, cudaMemcpy(&g, ctx->tmp1.p, sizeof(g), cudaMemcpyDeviceToHost));
  fixed_base_table_kernel<Fq2><<<blocks_for(32 * 255, 64), 64, 0, ctx->stream>>>(g, *slot);
  ctx->lc.n++;
  CUDA_TRY(ctx, cudaGetLastError());
  return ZKB_OK;
}

template <class F, class H>
int bases_generate_impl(zkb_ctx* ctx, Affine<F>** table_slot, const void* k_dev, size_t n, H** out) {
  if (!ctx) return ZKB_ERR_INVALID_ARG;
  if (!out || (!k_dev && n)) ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "bases_generate: bad argument");
  *out = nullptr;
  ZKB_TRY(set_device(ctx));
  ZKB_TRY(ensure_fixed_table<F>(ctx, table_slot));
  H* h = new (std::nothrow) H{ctx->device, nullptr, n};
  if (!h) ZKB_FAIL(ctx, ZKB_ERR_OOM, "bases_generate: host allocation failed");
  if (n) {
    cudaError_t e = cudaMalloc(&h->p, n * sizeof(Affine<F>));
    if (e != cudaSuccess) {
      cudaGetLastError();
      delete h;
      ZKB_FAIL(ctx, ZKB_ERR_OOM, "bases_generate: cudaMalloc failed: %s", cudaGetErrorString(e));
    }
    fixed_base_mul_kernel<F><<<blocks_for(n, 64), 64, 0, ctx->stream>>>(*table_slot, static_cast<const uint32_t*>(k_dev), n, h->p);
    ctx->lc.n++;
    cudaError_t e2 = cudaGetLastError();
    if (e2 != cudaSuccess) {
      cudaFree(h->p);
      delete h;
      ZKB_FAIL(ctx, ZKB_ERR_CUDA, "fixed_base_mul_kernel: %s", cudaGetErrorString(e2));
    }
  }
  *out = h;
  return ZKB_OK;
}

template <class F, class H>
int bases_read_impl(zkb_ctx* ctx, const H* b, size_t offset, size_t n, uint8_t* out_host) {
  if (!ctx) return ZKB_ERR_INVALID_ARG;
  if (!b || !out_host || offset + n > b->n) ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "bases_read: bad range");
  if (n == 0) return ZKB_OK;
  ZKB_TRY(set_device(ctx));
  CUDA_TRY(ctx, ctx->tmp0.reserve(n * sizeof(Affine<F>)));
  affine_export_kernel<F><<<blocks_for(n, 128), 128, 0, ctx->stream>>>(b->p + offset, ctx->tmp0.as<uint32_t>(), n);
  ctx->lc.n++;
  CUDA_TRY(ctx, cudaGetLastError());
  CUDA_TRY(ctx, cudaMemcpyAsync(out_host, ctx->tmp0.p, n * sizeof(Affine<F>), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return ZKB_OK;
}

}  // namespace

extern "C" int zkb_g1_bases_load(zkb_ctx* ctx, const uint8_t* h, size_t n, int validate, zkb_g1_bases** out) {
  return bases_load_impl<Fq, zkb_g1_bases>(ctx, h, n, validate, out);
}
extern "C" int zkb_g2_bases_load(zkb_ctx* ctx, const uint8_t* h, size_t n, int validate, zkb_g2_bases** out) {
  return bases_load_impl<Fq2, zkb_g2_bases>(ctx, h, n, validate, out);
}
extern "C" int zkb_g1_bases_generate(zkb_ctx* ctx, const void* k_dev, size_t n, zkb_g1_bases** out) {
  if (!ctx) return ZKB_ERR_INVALID_ARG;
  return bases_generate_impl<Fq, zkb_g1_bases>(ctx, &ctx->g1_table, k_dev, n, out);
}
extern "C" int zkb_g2_bases_generate(zkb_ctx* ctx, const void* k_dev, size_t n, zkb_g2_bases** out) {
  if (!ctx) return ZKB_ERR_INVALID_ARG;
  return bases_generate_impl<Fq2, zkb_g2_bases>(ctx, &ctx->g2_table, k_dev, n, out);
}
extern "C" size_t zkb_g1_bases_len(const zkb_g1_bases* b) { return b ? b->n : 0; }
extern "C" size_t zkb_g2_bases_len(const zkb_g2_bases* b) { return b ? b->n : 0; }
extern "C" int zkb_g1_bases_read(zkb_ctx* ctx, const zkb_g1_bases* b, size_t off, size_t n, uint8_t* o) {
  return bases_read_impl<Fq, zkb_g1_bases>(ctx, b, off, n, o);
}
extern "C" int zkb_g2_bases_read(zkb_ctx* ctx, const zkb_g2_bases* b, size_t off, size_t n, uint8_t* o) {
  return bases_read_impl<Fq2, zkb_g2_bases>(ctx, b, off, n, o);
}
extern "C" void zkb_g1_bases_free(zkb_g1_bases* b) {
  if (!b) return;
  cudaSetDevice(b->device);
  if (b->p) cudaFree(b->p);
  delete b;
}
extern "C" void zkb_g2_bases_free(zkb_g2_bases* b) {
  if (!b) return;
  cudaSetDevice(b->device);
  if (b->p) cudaFree(b->p);
  delete b;
}

// =============================================================================================== MSM
namespace {

template <class F, class H>
int msm_dev_impl(zkb_ctx* ctx, const H* bases, size_t offset, const void* scalars_dev, size_t n, void* out_affine_dev,
                 void* out_partial_dev) {
  if (!ctx) return ZKB_ERR_INVALID_ARG;
  if (!bases || offset + n > bases->n || (!scalars_dev && n)) ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "msm: bad bases range or scalars");
  if (bases->device != ctx->device) ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "msm: bases live on device %d, ctx on %d", bases->device, ctx->device);
  ZKB_TRY(set_device(ctx));
  cudaError_t e = msm_run<F>(ctx->msm_ws, ctx->lc, ctx->sm_count, ctx->msm_c, bases->p + offset,
                             static_cast<const uint32_t*>(scalars_dev), n, static_cast<XYZZ<F>*>(out_partial_dev),
                             static_cast<uint32_t*>(out_affine_dev), ctx->stream);
  if (e != cudaSuccess) {
    cudaGetLastError();
    ZKB_FAIL(ctx, e == cudaErrorMemoryAllocation ? ZKB_ERR_OOM : ZKB_ERR_CUDA, "msm_run: %s", cudaGetErrorString(e));
  }
  return ZKB_OK;
}

template <class F, class H>
int msm_host_impl(zkb_ctx* ctx, const H* bases, size_t offset, const uint8_t* scalars_host, size_t n, uint8_t* out) {
  if (!ctx) return ZKB_ERR_INVALID_ARG;
  if (!out || (!scalars_host && n)) ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "msm: null argument");
  ZKB_TRY(set_device(ctx));
  CUDA_TRY(ctx, ctx->scal.reserve(n * 32 + 32));
  CUDA_TRY(ctx, ctx->res.reserve(512));
  if (n) CUDA_TRY(ctx, cudaMemcpyAsync(ctx->scal.p, scalars_host, n * 32, cudaMemcpyHostToDevice, ctx->stream));
  ZKB_TRY((msm_dev_impl<F, H>(ctx, bases, offset, ctx->scal.p, n, ctx->res.p, nullptr)));
  CUDA_TRY(ctx, cudaMemcpyAsync(out, ctx->res.p, sizeof(Affine<F>), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return ZKB_OK;
}

}  // namespace

extern "C" int zkb_msm_g1(zkb_ctx* ctx, const zkb_g1_bases* b, size_t off, const uint8_t* s, size_t n, uint8_t out[64]) {
  return msm_host_impl<Fq, zkb_g1_bases>(ctx, b, off, s, n, out);
}
extern "C" int zkb_msm_g2(zkb_ctx* ctx, const zkb_g2_bases* b, size_t off, const uint8_t* s, size_t n, uint8_t out[128]) {
  return msm_host_impl<Fq2, zkb_g2_bases>(ctx, b, off, s, n, out);
}
extern "C" int zkb_msm_g1_dev(zkb_ctx* ctx, const zkb_g1_bases* b, size_t off, const void* s, size_t n, void* oa, void* op) {
  return msm_dev_impl<Fq, zkb_g1_bases>(ctx, b, off, s, n, oa, op);
}
extern "C" int zkb_msm_g2_dev(zkb_ctx* ctx, const zkb_g2_bases* b, size_t off, const void* s, size_t n, void* oa, void* op) {
  return msm_dev_impl<Fq2, zkb_g2_bases>(ctx, b, off, s, n, oa, op);
}
extern "C" int zkb_msm_g1_combine(zkb_ctx* ctx, const void* parts, int k, void* out) {
  if (!ctx) return ZKB_ERR_INVALID_ARG;
  if (!parts || k <= 0 || !out) ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "msm_combine: bad argument");
  ZKB_TRY(set_device(ctx));
  msm_combine_kernel<Fq><<<1, 32, 0, ctx->stream>>>(static_cast<const XYZZ<Fq>*>(parts), k, static_cast<uint32_t*>(out));
  ctx->lc.n++;
  CUDA_TRY(ctx, cudaGetLastError());
  return ZKB_OK;
}
extern "C" int zkb_msm_g2_combine(zkb_ctx* ctx, const void* parts, int k, void* out) {
  if (!ctx) return ZKB_ERR_INVALID_ARG;
  if (!parts || k <= 0 || !out) ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "msm_combine: bad argument");
  ZKB_TRY(set_device(ctx));
  msm_combine_kernel<Fq2><<<1, 32, 0, ctx->stream>>>(static_cast<const XYZZ<Fq2>*>(parts), k, static_cast<uint32_t*>(out));
  ctx->lc.n++;
  CUDA_TRY(ctx, cudaGetLastError());
  return ZKB_OK;
}

// =============================================================================================== NTT
namespace {

int ensure_wr(zkb_ctx* ctx) {
  if (ctx->wr_fwd) return ZKB_OK;
  const uint32_t cnt = 1u << (NTT_LRMAX - 1);
  CUDA_TRY(ctx, cudaMalloc(&ctx->wr_fwd, cnt * sizeof(Fr)));
  CUDA_TRY(ctx, cudaMalloc(&ctx->wr_inv, cnt * sizeof(Fr)));
  unsigned long long mult = 1ull << (28 - NTT_LRMAX);
  fr_pow_table_kernel<<<blocks_for(cnt, 64), 64, 0, ctx->stream>>>(ctx->wr_fwd, cnt, FRB_ROOT, mult, 1, 0, 0);
  fr_pow_table_kernel<<<blocks_for(cnt, 64), 64, 0, ctx->stream>>>(ctx->wr_inv, cnt, FRB_ROOT_INV, mult, 1, 0, 0);
  ctx->lc.n += 2;
  CUDA_TRY(ctx, cudaGetLastError());
  return ZKB_OK;
}

int ensure_ntt_tables(zkb_ctx* ctx, int logn, NttTables** out) {
  auto it = ctx->ntt_tables.find(logn);
  if (it != ctx->ntt_tables.end()) {
    *out = &it->second;
    return ZKB_OK;
  }
  NttTables t;
  int lb = (logn + 1) / 2;
  uint32_t nlo = 1u << lb, nhi = 1u << (logn - lb);
  size_t per = size_t(nlo) + nhi;
  CUDA_TRY(ctx, cudaMalloc(&t.mem, (4 * per + 1) * sizeof(Fr)));
  Fr* p = t.mem;
  unsigned long long wmult = 1ull << (28 - logn);
  auto gen = [&](Fr* lo, int base, unsigned long long mult, int sbase, unsigned long long sexp) {
    Fr* hi = lo + nlo;
    fr_pow_table_kernel<<<blocks_for(nlo, 64), 64, 0, ctx->stream>>>(lo, nlo, base, mult, 1, 0, 0);
    fr_pow_table_kernel<<<blocks_for(nhi, 64), 64, 0, ctx->stream>>>(hi, nhi, base, mult, 1ull << lb, sbase, sexp);
    ctx->lc.n += 2;
    PowTable pt;
    pt.lo = lo;
    pt.hi = hi;
    pt.lo_bits = lb;
    pt.scaled = sexp ? 1 : 0;
    return pt;
  };
  t.fwd = gen(p, FRB_ROOT, wmult, 0, 0);
  t.inv = gen(p + per, FRB_ROOT_INV, wmult, 0, 0);
  t.coset_pre = gen(p + 2 * per, FRB_GEN, 1, 0, 0);
  t.coset_inv_post = gen(p + 3 * per, FRB_GEN_INV, 1, FRB_INV2, (unsigned long long)logn);  // n^-1 g^-k
  Fr* ninv = p + 4 * per;
  // single constant n^-1 = (1/2)^logn: i = 0 gives base^0 = 1, times sbase^sexp
  fr_pow_table_kernel<<<1, 32, 0, ctx->stream>>>(ninv, 1, FRB_INV2, 1, 1, FRB_INV2, (unsigned long long)logn);
  ctx->lc.n++;
  t.ninv = ninv;
  CUDA_TRY(ctx, cudaGetLastError());
  auto ins = ctx->ntt_tables.emplace(logn, t);
  *out = &ins.first->second;
  return ZKB_OK;
}

template <int LR>
void launch_pass(const NttPassArgs& a, cudaStream_t st) {
  constexpr int TQ = 4;
  constexpr int NE = (1 << LR) * TQ;
  constexpr int NT = NE / 2 < 32 ? 32 : NE / 2;
  size_t nq = size_t(1) << (a.logn - LR);
  unsigned blocks = unsigned((nq + TQ - 1) / TQ);
  ntt_pass_kernel<LR, TQ><<<blocks, NT, 0, st>>>(a);
}

// scale_mode for the inverse: the post table multiplies by n^-1 (plain) or n^-1 g^-k (coset)
int ntt_dev_impl(zkb_ctx* ctx, const Fr* in, Fr* out, int logn, int inverse, int coset) {
  if (logn < 0 || logn > 28) ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "ntt: log_n %d outside [0, 28]", logn);
  size_t n = size_t(1) << logn;
  if (logn == 0) {
    // size-1 domain: identity (coset scale g^0 = 1, 1/n = 1)
    if (in != out) CUDA_TRY(ctx, cudaMemcpyAsync(out, in, sizeof(Fr), cudaMemcpyDeviceToDevice, ctx->stream));
    return ZKB_OK;
  }
  ZKB_TRY(ensure_wr(ctx));
  NttTables* T = nullptr;
  ZKB_TRY(ensure_ntt_tables(ctx, logn, &T));
  int npass = (logn + NTT_LRMAX - 1) / NTT_LRMAX;
  int base = logn / npass, extra = logn % npass;
  Fr* scratch[2] = {nullptr, nullptr};
  if (npass >= 2 || in == out) {
    CUDA_TRY(ctx, ctx->tmp0.reserve(n * sizeof(Fr)));
    scratch[0] = ctx->tmp0.as<Fr>();
  }
  if (npass >= 3) {
    CUDA_TRY(ctx, ctx->tmp1.reserve(n * sizeof(Fr)));
    scratch[1] = ctx->tmp1.as<Fr>();
  }
  const Fr* src = in;
  int logm = logn;
  PowTable ninv_tab;
  ninv_tab.lo = T->ninv;
  ninv_tab.hi = T->ninv;
  ninv_tab.lo_bits = -1;  // constant
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
    a.wr = inverse ? ctx->wr_inv : ctx->wr_fwd;
    a.lrmax = NTT_LRMAX;
    a.tw = inverse ? T->inv : T->fwd;
    a.has_pre = (p == 0 && !inverse && coset) ? 1 : 0;
    a.pre = T->coset_pre;
    a.has_post = (last && inverse) ? 1 : 0;
    a.post = coset ? T->coset_inv_post : ninv_tab;
    switch (lr) {
      case 1: launch_pass<1>(a, ctx->stream); break;
      case 2: launch_pass<2>(a, ctx->stream); break;
      case 3: launch_pass<3>(a, ctx->stream); break;
      case 4: launch_pass<4>(a, ctx->stream); break;
      case 5: launch_pass<5>(a, ctx->stream); break;
      case 6: launch_pass<6>(a, ctx->stream); break;
      case 7: launch_pass<7>(a, ctx->stream); break;
      default: launch_pass<8>(a, ctx->stream); break;
    }
    ctx->lc.n++;
    CUDA_TRY(ctx, cudaGetLastError());
    src = dst;
    logm -= lr;
  }
  if (npass == 1 && in == out) CUDA_TRY(ctx, cudaMemcpyAsync(out, scratch[0], n * sizeof(Fr), cudaMemcpyDeviceToDevice, ctx->stream));
  return ZKB_OK;
}

}  // namespace

extern "C" int zkb_ntt_dev(zkb_ctx* ctx, const void* in_dev, void* out_dev, int log_n, int direction, int coset) {
  if (!ctx) return ZKB_ERR_INVALID_ARG;
  if (!in_dev || !out_dev) ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "ntt: null buffer");
  ZKB_TRY(set_device(ctx));
  return ntt_dev_impl(ctx, static_cast<const Fr*>(in_dev), static_cast<Fr*>(out_dev), log_n, direction, coset);
}

extern "C" int zkb_ntt(zkb_ctx* ctx, const uint8_t* in_host, uint8_t* out_host, int log_n, int direction, int coset) {
  if (!ctx) return ZKB_ERR_INVALID_ARG;
  if (!in_host || !out_host) ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "ntt: null buffer");
  if (log_n < 0 || log_n > 28) ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "ntt: log_n %d outside [0, 28]", log_n);
  ZKB_TRY(set_device(ctx));
  size_t bytes = (size_t(1) << log_n) * 32;
  CUDA_TRY(ctx, ctx->scal.reserve(bytes));
  CUDA_TRY(ctx, ctx->tmp2.reserve(bytes));
  CUDA_TRY(ctx, cudaMemcpyAsync(ctx->scal.p, in_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
  ZKB_TRY(ntt_dev_impl(ctx, ctx->scal.as<Fr>(), ctx->tmp2.as<Fr>(), log_n, direction, coset));
  CUDA_TRY(ctx, cudaMemcpyAsync(out_host, ctx->tmp2.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return ZKB_OK;
}

// =============================================================================================== R1CS + witness map
namespace {

void csr_free(CsrDev& c) {
  if (c.row_ptr) cudaFree(c.row_ptr);
  if (c.col) cudaFree(c.col);
  if (c.coeff) cudaFree(c.coeff);
  c = CsrDev();
}

int csr_upload(zkb_ctx* ctx, const zkb_csr& h, uint64_t nc, uint64_t nvars, CsrDev& d) {
  if (!h.row_ptr) ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "r1cs: null row_ptr");
  if (h.row_ptr[0] != 0) ZKB_FAIL(ctx, ZKB_ERR_SHAPE, "r1cs: row_ptr[0] != 0");
  for (uint64_t i = 0; i < nc; i++)
    if (h.row_ptr[i + 1] < h.row_ptr[i]) ZKB_FAIL(ctx, ZKB_ERR_SHAPE, "r1cs: row_ptr not monotone at row %llu", (unsigned long long)i);
  size_t nnz = h.row_ptr[nc];
  if (nnz && (!h.col || !h.coeff)) ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "r1cs: null col/coeff");
  for (size_t k = 0; k < nnz; k++)
    if (h.col[k] >= nvars) ZKB_FAIL(ctx, ZKB_ERR_SHAPE, "r1cs: column %u >= %llu variables", h.col[k], (unsigned long long)nvars);
  d.nnz = nnz;
  CUDA_TRY(ctx, cudaMalloc(&d.row_ptr, (nc + 1) * sizeof(uint64_t)));
  CUDA_TRY(ctx, cudaMemcpyAsync(d.row_ptr, h.row_ptr, (nc + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
  if (nnz) {
    CUDA_TRY(ctx, cudaMalloc(&d.col, nnz * sizeof(uint32_t)));
    CUDA_TRY(ctx, cudaMalloc(&d.coeff, nnz * sizeof(Fr)));
    CUDA_TRY(ctx, cudaMemcpyAsync(d.col, h.col, nnz * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, ctx->tmp0.reserve(nnz * sizeof(Fr)));
    ZKB_TRY(clear_flag(ctx));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->tmp0.p, h.coeff, nnz * sizeof(Fr), cudaMemcpyHostToDevice, ctx->stream));
    to_mont_kernel<Fr><<<blocks_for(nnz, 128), 128, 0, ctx->stream>>>(ctx->tmp0.as<Fr>(), d.coeff, nnz, ctx->flag.as<int>());
    ctx->lc.n++;
    CUDA_TRY(ctx, cudaGetLastError());
    ZKB_TRY(check_flag(ctx, "r1cs coefficients"));
  }
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return ZKB_OK;
}

}  // namespace

extern "C" int zkb_r1cs_load(zkb_ctx* ctx, const zkb_r1cs_desc* desc, zkb_r1cs** out) {
  if (!ctx) return ZKB_ERR_INVALID_ARG;
  if (!desc || !out) ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "zkb_r1cs_load: null argument");
  *out = nullptr;
  if (desc->num_instance < 1) ZKB_FAIL(ctx, ZKB_ERR_SHAPE, "r1cs: num_instance must include the constant ONE");
  uint64_t dom = desc->num_constraints + desc->num_instance;
  int lg = 0;
  while ((uint64_t(1) << lg) < dom) lg++;
  if (lg > 28) ZKB_FAIL(ctx, ZKB_ERR_SHAPE, "r1cs: domain 2^%d exceeds Fr two-adicity 28", lg);
  ZKB_TRY(set_device(ctx));
  zkb_r1cs* m = new (std::nothrow) zkb_r1cs();
  if (!m) ZKB_FAIL(ctx, ZKB_ERR_OOM, "r1cs: host allocation failed");
  m->device = ctx->device;
  m->nc = desc->num_constraints;
  m->ni = desc->num_instance;
  m->nw = desc->num_witness;
  m->log_domain = lg;
  uint64_t nv = m->ni + m->nw;
  int s = csr_upload(ctx, desc->a, m->nc, nv, m->a);
  if (s == ZKB_OK) s = csr_upload(ctx, desc->b, m->nc, nv, m->b);
  if (s == ZKB_OK) s = csr_upload(ctx, desc->c, m->nc, nv, m->c);
  if (s != ZKB_OK) {
    zkb_r1cs_free(m);
    return s;
  }
  *out = m;
  return ZKB_OK;
}

extern "C" void zkb_r1cs_free(zkb_r1cs* m) {
  if (!m) return;
  cudaSetDevice(m->device);
  csr_free(m->a);
  csr_free(m->b);
  csr_free(m->c);
  delete m;
}

extern "C" int zkb_r1cs_log_domain(const zkb_r1cs* m) { return m ? m->log_domain : -1; }

namespace {

// Buffers used by the witness map (all domain_size Fr): wa, wb, wc + z (canonical) and z_mont.
struct WitnessBufs {
  Fr *z, *zm, *wa, *wb, *wc;
};

// h (canonical, device, domain_size elements) <- witness_map_from_matrices(m, z_dev canonical)
int witness_map_dev(zkb_ctx* ctx, const zkb_r1cs* m, const WitnessBufs& w, Fr* h_out) {
  const int lg = m->log_domain;
  const size_t n = size_t(1) << lg;
  const size_t nv = m->ni + m->nw;
  cudaStream_t st = ctx->stream;
  // a-chain in Montgomery form (so that a*b of a Montgomery and a canonical value is canonical)
  ZKB_TRY(clear_flag(ctx));
  to_mont_kernel<Fr><<<blocks_for(nv, 128), 128, 0, st>>>(w.z, w.zm, nv, ctx->flag.as<int>());
  csr_matvec_kernel<<<blocks_for(n, 128), 128, 0, st>>>(m->a.row_ptr, m->a.col, m->a.coeff, w.zm, m->nc, m->ni, 1, n, w.wa);
  csr_matvec_kernel<<<blocks_for(n, 128), 128, 0, st>>>(m->b.row_ptr, m->b.col, m->b.coeff, w.z, m->nc, m->ni, 0, n, w.wb);
  csr_matvec_kernel<<<blocks_for(n, 128), 128, 0, st>>>(m->c.row_ptr, m->c.col, m->c.coeff, w.z, m->nc, m->ni, 0, n, w.wc);
  ctx->lc.n += 4;
  CUDA_TRY(ctx, cudaGetLastError());
  Fr* chains[3] = {w.wa, w.wb, w.wc};
  for (Fr* v : chains) {
    ZKB_TRY(ntt_dev_impl(ctx, v, v, lg, 1, 0));  // domain.ifft_in_place
    ZKB_TRY(ntt_dev_impl(ctx, v, v, lg, 0, 1));  // coset_domain.fft_in_place
  }
  qap_pointwise_kernel<<<blocks_for(n, 128), 128, 0, st>>>(w.wa, w.wb, w.wc, lg, n, w.wa);
  ctx->lc.n++;
  CUDA_TRY(ctx, cudaGetLastError());
  ZKB_TRY(ntt_dev_impl(ctx, w.wa, h_out, lg, 1, 1));  // coset_domain.ifft_in_place
  return ZKB_OK;
}

}  // namespace

struct ProveBufs {
  DevBuf z, zm, wa, wb, wc, h, za, zl, rs, pts;
};

// one lazily created scratch set per context (kept in a side table to keep zkb_ctx POD-ish)
static std::map<zkb_ctx*, ProveBufs>& prove_bufs() {
  static std::map<zkb_ctx*, ProveBufs> m;
  return m;
}

extern "C" int zkb_witness_map(zkb_ctx* ctx, const zkb_r1cs* m, const uint8_t* z_host, uint8_t* h_out_host) {
  if (!ctx) return ZKB_ERR_INVALID_ARG;
  if (!m || !z_host || !h_out_host) ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "zkb_witness_map: null argument");
  if (m->device != ctx->device) ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "zkb_witness_map: matrices on another device");
  ZKB_TRY(set_device(ctx));
  ProveBufs& pb = prove_bufs()[ctx];
  const size_t n = size_t(1) << m->log_domain, nv = m->ni + m->nw;
  CUDA_TRY(ctx, pb.z.reserve(nv * 32));
  CUDA_TRY(ctx, pb.zm.reserve(nv * 32));
  CUDA_TRY(ctx, pb.wa.reserve(n * 32));
  CUDA_TRY(ctx, pb.wb.reserve(n * 32));
  CUDA_TRY(ctx, pb.wc.reserve(n * 32));
  CUDA_TRY(ctx, pb.h.reserve(n * 32));
  CUDA_TRY(ctx, cudaMemcpyAsync(pb.z.p, z_host, nv * 32, cudaMemcpyHostToDevice, ctx->stream));
  WitnessBufs w{pb.z.as<Fr>(), pb.zm.as<Fr>(), pb.wa.as<Fr>(), pb.wb.as<Fr>(), pb.wc.as<Fr>()};
  ZKB_TRY(witness_map_dev(ctx, m, w, pb.h.as<Fr>()));
  CUDA_TRY(ctx, cudaMemcpyAsync(h_out_host, pb.h.p, n * 32, cudaMemcpyDeviceToHost, ctx->stream));
  return check_flag(ctx, "witness assignment");
}

// =============================================================================================== proving key + prove
namespace {

template <class F, class H>
int load_ext(zkb_ctx* ctx, const uint8_t* q, size_t qlen, bool rotate_first, const uint8_t* const* extra, int nextra,
             int validate, H** out) {
  // host-side concatenation: q[1..] || q[0] || extra...   (or q || extra... when !rotate_first)
  const size_t sz = sizeof(Affine<F>);
  std::vector<uint8_t> buf;
  try {
    buf.resize((qlen + nextra) * sz);
  } catch (...) {
    ZKB_FAIL(ctx, ZKB_ERR_OOM, "pk_load: host staging allocation failed");
  }
  size_t o = 0;
  if (rotate_first && qlen) {
    memcpy(buf.data(), q + sz, (qlen - 1) * sz);
    o = (qlen - 1) * sz;
    memcpy(buf.data() + o, q, sz);
    o += sz;
  } else if (qlen) {
    memcpy(buf.data(), q, qlen * sz);
    o = qlen * sz;
  }
  for (int i = 0; i < nextra; i++) {
    memcpy(buf.data() + o, extra[i], sz);
    o += sz;
  }
  return bases_load_impl<F, H>(ctx, buf.data(), qlen + nextra, validate, out);
}

}  // namespace

extern "C" void zkb_pk_free(zkb_pk* pk) {
  if (!pk) return;
  zkb_g1_bases_free(pk->a_ext);
  zkb_g1_bases_free(pk->b1_ext);
  zkb_g1_bases_free(pk->l_ext);
  zkb_g1_bases_free(pk->h);
  zkb_g2_bases_free(pk->b2_ext);
  delete pk;
}

extern "C" int zkb_pk_load(zkb_ctx* ctx, const zkb_pk_desc* d, int validate, zkb_pk** out) {
  if (!ctx) return ZKB_ERR_INVALID_ARG;
  if (!d || !out) ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "zkb_pk_load: null argument");
  *out = nullptr;
  if (!d->alpha_g1 || !d->beta_g1 || !d->beta_g2 || !d->delta_g1 || !d->delta_g2 || !d->a_query || !d->b_g1_query ||
      !d->b_g2_query || (!d->h_query && d->h_len) || (!d->l_query && d->l_len))
    ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "zkb_pk_load: null key component");
  if (d->a_len < 1 || d->b_g1_len != d->a_len || d->b_g2_len != d->a_len)
    ZKB_FAIL(ctx, ZKB_ERR_SHAPE, "zkb_pk_load: a/b_g1/b_g2 query lengths %zu/%zu/%zu disagree", d->a_len, d->b_g1_len, d->b_g2_len);
  if (d->l_len > d->a_len - 1) ZKB_FAIL(ctx, ZKB_ERR_SHAPE, "zkb_pk_load: l_query longer than the witness");
  ZKB_TRY(set_device(ctx));
  zkb_pk* pk = new (std::nothrow) zkb_pk();
  if (!pk) ZKB_FAIL(ctx, ZKB_ERR_OOM, "zkb_pk_load: host allocation failed");
  pk->device = ctx->device;
  pk->nv = d->a_len;
  pk->nw = d->l_len;
  pk->nh = d->h_len;
  const uint8_t* ea[2] = {d->alpha_g1, d->delta_g1};
  const uint8_t* eb1[2] = {d->beta_g1, d->delta_g1};
  const uint8_t* eb2[2] = {d->beta_g2, d->delta_g2};
  const uint8_t* el[1] = {d->delta_g1};
  int s = load_ext<Fq, zkb_g1_bases>(ctx, d->a_query, d->a_len, true, ea, 2, validate, &pk->a_ext);
  if (s == ZKB_OK) s = load_ext<Fq, zkb_g1_bases>(ctx, d->b_g1_query, d->b_g1_len, true, eb1, 2, validate, &pk->b1_ext);
  if (s == ZKB_OK) s = load_ext<Fq2, zkb_g2_bases>(ctx, d->b_g2_query, d->b_g2_len, true, eb2, 2, validate, &pk->b2_ext);
  if (s == ZKB_OK) s = load_ext<Fq, zkb_g1_bases>(ctx, d->l_query, d->l_len, false, el, 1, validate, &pk->l_ext);
  if (s == ZKB_OK) s = load_ext<Fq, zkb_g1_bases>(ctx, d->h_query, d->h_len, false, nullptr, 0, validate, &pk->h);
  if (s != ZKB_OK) {
    zkb_pk_free(pk);
    return s;
  }
  *out = pk;
  return ZKB_OK;
}

extern "C" int zkb_prove(zkb_ctx* ctx, const zkb_pk* pk, const zkb_r1cs* m, const uint8_t* z_host, const uint8_t r[32],
                         const uint8_t s[32], uint8_t out_a[64], uint8_t out_b[128], uint8_t out_c[64]) {
  if (!ctx) return ZKB_ERR_INVALID_ARG;
  if (!pk || !m || !z_host || !r || !s || !out_a || !out_b || !out_c) ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "zkb_prove: null argument");
  if (pk->device != ctx->device || m->device != ctx->device) ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "zkb_prove: key/matrices on another device");
  const size_t nv = m->ni + m->nw, nw = m->nw, ni = m->ni;
  const size_t n = size_t(1) << m->log_domain;
  if (pk->nv != nv) ZKB_FAIL(ctx, ZKB_ERR_SHAPE, "zkb_prove: key has %zu variables, circuit has %zu", pk->nv, nv);
  if (pk->nw != nw) ZKB_FAIL(ctx, ZKB_ERR_SHAPE, "zkb_prove: l_query has %zu points, circuit has %zu witness variables", pk->nw, nw);
  ZKB_TRY(set_device(ctx));
  cudaStream_t st = ctx->stream;
  ProveBufs& pb = prove_bufs()[ctx];
  CUDA_TRY(ctx, pb.z.reserve(nv * 32));
  CUDA_TRY(ctx, pb.zm.reserve(nv * 32));
  CUDA_TRY(ctx, pb.wa.reserve(n * 32));
  CUDA_TRY(ctx, pb.wb.reserve(n * 32));
  CUDA_TRY(ctx, pb.wc.reserve(n * 32));
  CUDA_TRY(ctx, pb.h.reserve(n * 32));
  CUDA_TRY(ctx, pb.za.reserve((nv + 2) * 32));
  CUDA_TRY(ctx, pb.zl.reserve((nw + 1) * 32));
  CUDA_TRY(ctx, pb.rs.reserve(64));
  CUDA_TRY(ctx, pb.pts.reserve(4 * sizeof(XYZZ<Fq>) + 64 + 128 + 64));
  Fr* rs = pb.rs.as<Fr>();
  CUDA_TRY(ctx, cudaMemcpyAsync(pb.z.p, z_host, nv * 32, cudaMemcpyHostToDevice, st));
  CUDA_TRY(ctx, cudaMemcpyAsync(rs, r, 32, cudaMemcpyHostToDevice, st));
  CUDA_TRY(ctx, cudaMemcpyAsync(rs + 1, s, 32, cudaMemcpyHostToDevice, st));

  // h = witness_map_from_matrices
  WitnessBufs w{pb.z.as<Fr>(), pb.zm.as<Fr>(), pb.wa.as<Fr>(), pb.wb.as<Fr>(), pb.wc.as<Fr>()};
  ZKB_TRY(witness_map_dev(ctx, m, w, pb.h.as<Fr>()));

  // scalar vectors of the folded MSMs
  Fr* za = pb.za.as<Fr>();
  Fr* zl = pb.zl.as<Fr>();
  if (nv > 1) CUDA_TRY(ctx, cudaMemcpyAsync(za, pb.z.as<Fr>() + 1, (nv - 1) * 32, cudaMemcpyDeviceToDevice, st));
  if (nw) CUDA_TRY(ctx, cudaMemcpyAsync(zl, pb.z.as<Fr>() + ni, nw * 32, cudaMemcpyDeviceToDevice, st));
  prove_tail_scalars_kernel<<<1, 32, 0, st>>>(rs, rs + 1, za + (nv - 1), zl + nw);
  ctx->lc.n++;
  CUDA_TRY(ctx, cudaGetLastError());

  char* pts = static_cast<char*>(pb.pts.p);
  XYZZ<Fq>* pA = reinterpret_cast<XYZZ<Fq>*>(pts);
  XYZZ<Fq>* pB1 = pA + 1;
  XYZZ<Fq>* pL = pA + 2;
  XYZZ<Fq>* pH = pA + 3;
  uint32_t* oA = reinterpret_cast<uint32_t*>(pts + 4 * sizeof(XYZZ<Fq>));
  uint32_t* oB = oA + 16;
  uint32_t* oC = oB + 32;

  // A = MSM(a_ext, z[1..] || 1 || 1 || r)
  ZKB_TRY((msm_dev_impl<Fq, zkb_g1_bases>(ctx, pk->a_ext, 0, za, nv + 2, oA, pA)));
  // H: msm_bigint truncates to the shorter of (h_query, h)
  size_t hn = pk->nh < n ? pk->nh : n;
  ZKB_TRY((msm_dev_impl<Fq, zkb_g1_bases>(ctx, pk->h, 0, pb.h.p, hn, nullptr, pH)));
  // L = MSM(l_query || delta_1, aux || -(r s))
  ZKB_TRY((msm_dev_impl<Fq, zkb_g1_bases>(ctx, pk->l_ext, 0, zl, nw + 1, nullptr, pL)));
  // B: same scalars with s in the last slot
  CUDA_TRY(ctx, cudaMemcpyAsync(za + (nv + 1), rs + 1, 32, cudaMemcpyDeviceToDevice, st));
  ZKB_TRY((msm_dev_impl<Fq, zkb_g1_bases>(ctx, pk->b1_ext, 0, za, nv + 2, nullptr, pB1)));
  ZKB_TRY((msm_dev_impl<Fq2, zkb_g2_bases>(ctx, pk->b2_ext, 0, za, nv + 2, oB, nullptr)));
  // C = s A + r B1 + L + H
  prove_assemble_c_kernel<<<1, 64, 0, st>>>(pA, pB1, pL, pH, reinterpret_cast<uint32_t*>(rs), reinterpret_cast<uint32_t*>(rs + 1), oC);
  ctx->lc.n++;
  CUDA_TRY(ctx, cudaGetLastError());
  CUDA_TRY(ctx, cudaMemcpyAsync(out_a, oA, 64, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(ctx, cudaMemcpyAsync(out_b, oB, 128, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(ctx, cudaMemcpyAsync(out_c, oC, 64, cudaMemcpyDeviceToHost, st));
  return check_flag(ctx, "witness assignment");
}
