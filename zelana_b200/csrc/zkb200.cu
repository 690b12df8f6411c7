// C ABI of the B200-native Groth16/BN254 backend (declared in include/zkb200.h).
// Everything mathematical runs in the CUDA kernels of fp.cuh / ec.cuh / msm.cuh / ntt.cuh (compiled in fr.cu,
// g1.cu, g2.cu); the host code here only moves bytes, sizes workspaces and orders launches.
#include "internal.h"

#include <cstdlib>
#include <initializer_list>
#include <mutex>
#include <utility>

using namespace zkb;

// =============================================================================================== helpers
namespace zkb {

static int ensure_pinned(zkb_ctx* ctx) {
  if (!ctx->pinned) CUDA_TRY(ctx, cudaHostAlloc(reinterpret_cast<void**>(&ctx->pinned), 512, cudaHostAllocDefault));
  return ZKB_OK;
}

int check_flag(zkb_ctx* ctx, const char* what) {
  ZKB_ON_DEVICE(ctx);   // the blocking-sync event below belongs to the device that is current when it is created
  ZKB_TRY(ensure_pinned(ctx));
  int* hf = reinterpret_cast<int*>(ctx->pinned + 256);
  CUDA_TRY(ctx, cudaMemcpyAsync(hf, ctx->flag.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  if (ctx->blocking_sync) {
    if (!ctx->sync_ev) CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->sync_ev, cudaEventBlockingSync | cudaEventDisableTiming));
    CUDA_TRY(ctx, cudaEventRecord(ctx->sync_ev, ctx->stream));
    CUDA_TRY(ctx, cudaEventSynchronize(ctx->sync_ev));
  } else {
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  }
  const int h = *hf;
  if (h == 1) ZKB_FAIL(ctx, ZKB_ERR_NOT_CANONICAL, "%s: field element >= modulus", what);
  if (h == 2) ZKB_FAIL(ctx, ZKB_ERR_NOT_CANONICAL, "%s: point not on curve", what);
  if (h == 3) ZKB_FAIL(ctx, ZKB_ERR_NOT_CANONICAL, "%s: both the infinity and the sign flag are set", what);
  if (h == 4) ZKB_FAIL(ctx, ZKB_ERR_NOT_CANONICAL, "%s: point not in the prime-order subgroup", what);
  if (h != 0) ZKB_FAIL(ctx, ZKB_ERR_CUDA, "%s: unexpected device status flag %d", what, h);
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

}  // namespace zkb

// =============================================================================================== context
extern "C" const char* zkb_version(void) { return "zkb200 0.2 (sm_100a)"; }

// A prove runs on up to six streams and batches use several lanes per GPU: with the default 8 hardware queues, streams alias
// and a long single-block kernel of one proof stalls the others.  CUDA reads CUDA_DEVICE_MAX_CONNECTIONS once, when the
// process creates its CUDA context, so the host should export it before its first CUDA call (for Python hosts
// zelana_b200/__init__.py does).  As a convenience the FIRST zkb_ctx_create of a process sets it if the user has not -- once,
// under std::call_once (setenv is not safe against concurrent getenv in a threaded host, so never from a query function).
static void max_connections_once() {
  static std::once_flag once;
  std::call_once(once, [] { setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0); });
}

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
  max_connections_once();
  int n = zkb_device_count();
  if (n <= 0 || device < 0 || device >= n) return ZKB_ERR_NO_DEVICE;
  zkb_ctx* ctx = new (std::nothrow) zkb_ctx();
  if (!ctx) return ZKB_ERR_OOM;
  ctx->device = device;
  DeviceGuard dg;
  if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
    cudaGetLastError();
    delete ctx;
    return ZKB_ERR_CUDA;
  }
  ctx->own_stream = true;
  cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
  if (const char* fg = getenv("ZKB_L2_FETCH_GRANULARITY")) {
    // experiment knob: the L2's DRAM fetch granularity for this device, 32 / 64 / 128 bytes; a hint the driver may ignore
    int v = atoi(fg);
    if (v == 32 || v == 64 || v == 128) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, size_t(v));
    cudaGetLastError();
  }
  if (const char* sl = getenv("ZKB_MSM_SLICES")) {
    int v = atoi(sl);
    if (v >= 1 && v <= ZKB_MAX_SLICES - 1) ctx->msm_slices = v;
  }
  *out = ctx;
  return ZKB_OK;
}

extern "C" int zkb_ctx_set_stream(zkb_ctx* ctx, void* cuda_stream) {
  if (!ctx) return ZKB_ERR_INVALID_ARG;
  ZKB_ON_DEVICE(ctx);
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
  ZKB_ON_DEVICE(ctx);
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return ZKB_OK;
}

extern "C" void zkb_ctx_destroy(zkb_ctx* ctx) {
  if (!ctx) return;
  DeviceGuard dg(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (auto& g : ctx->graphs)
    if (g.exec) cudaGraphExecDestroy(g.exec);
  ctx->graphs.clear();
  ctx->msm_ws.release();
  ctx->pzb.release();
  for (auto& lane : ctx->aux) {
    if (lane.stream) {
      cudaStreamSynchronize(lane.stream);
      cudaStreamDestroy(lane.stream);
      cudaEventDestroy(lane.done);
    }
    lane.ws.release();
  }
  if (ctx->ev_inputs) cudaEventDestroy(ctx->ev_inputs);
  if (ctx->sync_ev) cudaEventDestroy(ctx->sync_ev);
  if (ctx->pinned) cudaFreeHost(ctx->pinned);
  if (ctx->bpinned) cudaFreeHost(ctx->bpinned);
  for (DevBuf* b : {&ctx->bz, &ctx->bzm, &ctx->bw3, &ctx->bh, &ctx->brs, &ctx->bpart, &ctx->bout}) b->release();
  DevBuf* bufs[] = {&ctx->scal, &ctx->res, &ctx->tmp0, &ctx->tmp1, &ctx->tmp2, &ctx->flag, &ctx->pz, &ctx->pzm, &ctx->pwa,
                    &ctx->pwb, &ctx->pwc, &ctx->ph, &ctx->pza, &ctx->pzl, &ctx->prs, &ctx->ppts, &ctx->pzsa, &ctx->pzrb};
  for (DevBuf* b : bufs) b->release();
  fr_state_free(ctx);
  fixed_table_free<Fq>(ctx);
  fixed_table_free<Fq2>(ctx);
  for (auto& sp : ctx->prof.pending) {
    cudaEventDestroy(sp.a);
    cudaEventDestroy(sp.b);
  }
  for (cudaEvent_t e : ctx->prof.pool) cudaEventDestroy(e);
  if (ctx->copy_stream) {
    cudaStreamSynchronize(ctx->copy_stream);
    cudaStreamDestroy(ctx->copy_stream);
    for (int k = 0; k < ZKB_MAX_SLICES; k++)
      if (ctx->copy_done[k]) cudaEventDestroy(ctx->copy_done[k]);
  }
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

// Page-locked host memory for buffers the GPU reads or writes asynchronously (assignments of a batch, results).
extern "C" int zkb_host_alloc_pinned(size_t bytes, void** out) {
  if (!out) return ZKB_ERR_INVALID_ARG;
  *out = nullptr;
  if (cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) {
    cudaGetLastError();
    *out = nullptr;
    return ZKB_ERR_OOM;
  }
  return ZKB_OK;
}
extern "C" void zkb_host_free_pinned(void* p) {
  if (p) cudaFreeHost(p);
}

// Verifies the canaries around every scratch buffer of the context (allocated while ZKB_GUARD=1 was set): ZKB_OK if intact.
extern "C" int zkb_debug_check_guards(zkb_ctx* ctx) {
  if (!ctx) return ZKB_ERR_INVALID_ARG;
  ZKB_ON_DEVICE(ctx);
  CUDA_TRY(ctx, cudaDeviceSynchronize());
  struct Named { const char* name; const DevBuf* b; };
  const Named bufs[] = {{"msm_ws", &ctx->msm_ws}, {"lane0", &ctx->aux[0].ws}, {"lane1", &ctx->aux[1].ws}, {"lane2", &ctx->aux[2].ws},
                        {"lane3", &ctx->aux[3].ws}, {"lane4", &ctx->aux[4].ws}, {"scal", &ctx->scal}, {"res", &ctx->res},
                        {"tmp0", &ctx->tmp0}, {"tmp1", &ctx->tmp1}, {"tmp2", &ctx->tmp2}, {"flag", &ctx->flag}, {"pz", &ctx->pz},
                        {"pzm", &ctx->pzm}, {"pwa", &ctx->pwa}, {"pwb", &ctx->pwb}, {"pwc", &ctx->pwc}, {"ph", &ctx->ph},
                        {"pza", &ctx->pza}, {"pzb", &ctx->pzb}, {"pzl", &ctx->pzl}, {"prs", &ctx->prs}, {"ppts", &ctx->ppts},
                        {"pzsa", &ctx->pzsa}, {"pzrb", &ctx->pzrb}, {"bz", &ctx->bz}, {"bzm", &ctx->bzm}, {"bw3", &ctx->bw3},
                        {"bh", &ctx->bh}, {"brs", &ctx->brs}, {"bpart", &ctx->bpart}, {"bout", &ctx->bout}};
  int guarded = 0;
  for (const Named& n : bufs) {
    if (n.b->guarded && n.b->base) guarded++;
    int rc = n.b->check_guards();
    if (rc) ZKB_FAIL(ctx, ZKB_ERR_CUDA, "scratch buffer %s: the canary %s it was overwritten", n.name, rc == 1 ? "before" : "after");
  }
  if (!guarded) ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "no guarded buffer: set ZKB_GUARD=1 before the context allocates");
  return ZKB_OK;
}

extern "C" const char* zkb_last_error(zkb_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
extern "C" unsigned long long zkb_launch_count(zkb_ctx* ctx) { return ctx ? ctx->launches : 0; }
extern "C" int zkb_ctx_set_graphs(zkb_ctx* ctx, int on) {
  if (!ctx) return ZKB_ERR_INVALID_ARG;
  ctx->graphs_on = on != 0;
  return ZKB_OK;
}
extern "C" int zkb_ctx_set_blocking_sync(zkb_ctx* ctx, int on) {
  if (!ctx) return ZKB_ERR_INVALID_ARG;
  ctx->blocking_sync = on != 0;
  return ZKB_OK;
}
extern "C" int zkb_graph_stats(zkb_ctx* ctx, unsigned long long* captures, unsigned long long* replays) {
  if (!ctx) return ZKB_ERR_INVALID_ARG;
  if (captures) *captures = ctx->graph_captures;
  if (replays) *replays = ctx->graph_replays;
  return ZKB_OK;
}
extern "C" int zkb_ctx_set_msm_window(zkb_ctx* ctx, int c) {
  if (!ctx || c < 0 || c > 23 || (c > 0 && c < 2)) return ZKB_ERR_INVALID_ARG;
  ctx->msm_c = c;
  return ZKB_OK;
}

// ---- per-phase device timing ---------------------------------------------------------------------------
static const char* const kPhaseNames[PH_COUNT] = {
    "msm_g1_digits", "msm_g1_sort", "msm_g1_accumulate", "msm_g1_reduce",
    "msm_g2_digits", "msm_g2_sort", "msm_g2_accumulate", "msm_g2_reduce",
    "ntt", "matvec", "qap_pointwise", "assemble"};

extern "C" int zkb_prof_phase_count(void) { return PH_COUNT; }
extern "C" const char* zkb_prof_phase_name(int phase) { return phase >= 0 && phase < PH_COUNT ? kPhaseNames[phase] : ""; }

extern "C" int zkb_prof_enable(zkb_ctx* ctx, int on) {
  if (!ctx) return ZKB_ERR_INVALID_ARG;
  ctx->prof.on = on != 0;
  return ZKB_OK;
}

static int prof_collect(zkb_ctx* ctx) {
  ZKB_ON_DEVICE(ctx);
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  for (auto& sp : ctx->prof.pending) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, sp.a, sp.b) == cudaSuccess) {
      ctx->prof.ms[sp.phase] += ms;
      ctx->prof.cnt[sp.phase]++;
    } else {
      cudaGetLastError();
    }
    ctx->prof.pool.push_back(sp.a);
    ctx->prof.pool.push_back(sp.b);
  }
  ctx->prof.pending.clear();
  return ZKB_OK;
}

extern "C" int zkb_prof_reset(zkb_ctx* ctx) {
  if (!ctx) return ZKB_ERR_INVALID_ARG;
  ZKB_TRY(prof_collect(ctx));
  for (int i = 0; i < PH_COUNT; i++) {
    ctx->prof.ms[i] = 0;
    ctx->prof.cnt[i] = 0;
  }
  return ZKB_OK;
}

extern "C" int zkb_prof_read(zkb_ctx* ctx, int phase, double* total_ms, unsigned long long* spans) {
  if (!ctx) return ZKB_ERR_INVALID_ARG;
  if (phase < 0 || phase >= PH_COUNT) ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "zkb_prof_read: phase %d out of range", phase);
  ZKB_TRY(prof_collect(ctx));
  if (total_ms) *total_ms = ctx->prof.ms[phase];
  if (spans) *spans = ctx->prof.cnt[phase];
  return ZKB_OK;
}

// =============================================================================================== field / curve hooks
extern "C" int zkb_field_op(zkb_ctx* ctx, int field, int op, const uint8_t* a, const uint8_t* b, size_t n, uint8_t* out) {
  if (!ctx) return ZKB_ERR_INVALID_ARG;
  if (field < 0 || field > 1 || op < 0 || op > 4 || !a || !out || (op <= 2 && !b)) ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "zkb_field_op: bad argument");
  if (n == 0) return ZKB_OK;
  ZKB_ON_DEVICE(ctx);
  return field == 0 ? fr_field_op(ctx, op, a, b, n, out) : fq_field_op(ctx, op, a, b, n, out);
}

extern "C" int zkb_scalar_mul(zkb_ctx* ctx, int group, const uint8_t* points, const uint8_t* scalars, size_t n, uint8_t* out) {
  if (!ctx) return ZKB_ERR_INVALID_ARG;
  if ((group != 1 && group != 2) || !points || !scalars || !out) ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "zkb_scalar_mul: bad argument");
  ZKB_ON_DEVICE(ctx);
  return group == 1 ? scalar_mul_impl<Fq>(ctx, points, scalars, n, out) : scalar_mul_impl<Fq2>(ctx, points, scalars, n, out);
}

extern "C" int zkb_point_sum(zkb_ctx* ctx, int group, const uint8_t* points, size_t n, uint8_t* out) {
  if (!ctx) return ZKB_ERR_INVALID_ARG;
  if ((group != 1 && group != 2) || (!points && n) || !out) ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "zkb_point_sum: bad argument");
  ZKB_ON_DEVICE(ctx);
  return group == 1 ? point_sum_impl<Fq>(ctx, points, n, out) : point_sum_impl<Fq2>(ctx, points, n, out);
}

// =============================================================================================== bases
extern "C" int zkb_g1_bases_load(zkb_ctx* ctx, const uint8_t* h, size_t n, int validate, zkb_g1_bases** out) {
  return bases_load_impl<Fq>(ctx, h, n, validate, out);
}
extern "C" int zkb_g2_bases_load(zkb_ctx* ctx, const uint8_t* h, size_t n, int validate, zkb_g2_bases** out) {
  return bases_load_impl<Fq2>(ctx, h, n, validate, out);
}
extern "C" int zkb_g1_bases_generate(zkb_ctx* ctx, const void* k_dev, size_t n, zkb_g1_bases** out) {
  return bases_generate_impl<Fq>(ctx, k_dev, n, out);
}
extern "C" int zkb_g2_bases_generate(zkb_ctx* ctx, const void* k_dev, size_t n, zkb_g2_bases** out) {
  return bases_generate_impl<Fq2>(ctx, k_dev, n, out);
}
extern "C" int zkb_g1_bases_window(const zkb_g1_bases* b, int* c, int* nwin) {
  if (!b) return ZKB_ERR_INVALID_ARG;
  if (c) *c = b->c;
  if (nwin) *nwin = b->nwin;
  return ZKB_OK;
}
extern "C" int zkb_g2_bases_window(const zkb_g2_bases* b, int* c, int* nwin) {
  if (!b) return ZKB_ERR_INVALID_ARG;
  if (c) *c = b->c;
  if (nwin) *nwin = b->nwin;
  return ZKB_OK;
}
extern "C" size_t zkb_g1_bases_len(const zkb_g1_bases* b) { return b ? b->n : 0; }
extern "C" size_t zkb_g2_bases_len(const zkb_g2_bases* b) { return b ? b->n : 0; }
extern "C" int zkb_g1_bases_read(zkb_ctx* ctx, const zkb_g1_bases* b, size_t off, size_t n, uint8_t* o) {
  return bases_read_impl<Fq>(ctx, b, off, n, o);
}
extern "C" int zkb_g2_bases_read(zkb_ctx* ctx, const zkb_g2_bases* b, size_t off, size_t n, uint8_t* o) {
  return bases_read_impl<Fq2>(ctx, b, off, n, o);
}
extern "C" void zkb_g1_bases_free(zkb_g1_bases* b) { bases_free_impl<Fq>(b); }
extern "C" void zkb_g2_bases_free(zkb_g2_bases* b) { bases_free_impl<Fq2>(b); }

// =============================================================================================== MSM
extern "C" int zkb_msm_g1(zkb_ctx* ctx, const zkb_g1_bases* b, size_t off, const uint8_t* s, size_t n, uint8_t out[64]) {
  return msm_host_impl<Fq>(ctx, b, off, s, n, out);
}
extern "C" int zkb_msm_g2(zkb_ctx* ctx, const zkb_g2_bases* b, size_t off, const uint8_t* s, size_t n, uint8_t out[128]) {
  return msm_host_impl<Fq2>(ctx, b, off, s, n, out);
}
extern "C" int zkb_msm_g1_dev(zkb_ctx* ctx, const zkb_g1_bases* b, size_t off, const void* s, size_t n, void* oa, void* op) {
  return msm_dev_impl<Fq>(ctx, b, off, s, n, oa, op);
}
extern "C" int zkb_msm_g2_dev(zkb_ctx* ctx, const zkb_g2_bases* b, size_t off, const void* s, size_t n, void* oa, void* op) {
  return msm_dev_impl<Fq2>(ctx, b, off, s, n, oa, op);
}
extern "C" int zkb_msm_g1_partial(zkb_ctx* ctx, const zkb_g1_bases* b, size_t off, const uint8_t* s, size_t n, void* op) {
  return msm_host_partial_impl<Fq>(ctx, b, off, s, n, op);
}
extern "C" int zkb_msm_g2_partial(zkb_ctx* ctx, const zkb_g2_bases* b, size_t off, const uint8_t* s, size_t n, void* op) {
  return msm_host_partial_impl<Fq2>(ctx, b, off, s, n, op);
}
extern "C" int zkb_msm_g1_multi(zkb_ctx* const* ctxs, const zkb_g1_bases* const* bases, int n_gpus, const uint8_t* s, size_t n,
                                uint8_t out[64]) {
  return msm_multi_impl<Fq>(ctxs, bases, n_gpus, s, n, out);
}
extern "C" int zkb_msm_g2_multi(zkb_ctx* const* ctxs, const zkb_g2_bases* const* bases, int n_gpus, const uint8_t* s, size_t n,
                                uint8_t out[128]) {
  return msm_multi_impl<Fq2>(ctxs, bases, n_gpus, s, n, out);
}
extern "C" int zkb_msm_g1_combine(zkb_ctx* ctx, const void* parts, int k, void* out) { return msm_combine_impl<Fq>(ctx, parts, k, out); }
extern "C" int zkb_msm_g2_combine(zkb_ctx* ctx, const void* parts, int k, void* out) { return msm_combine_impl<Fq2>(ctx, parts, k, out); }

// parity hook for the batched MSM (msm_run_batch): `batch` scalar vectors, vector p at scalars_dev + p * stride * 32 bytes,
// against bases [offset, offset + n) -> batch canonical affine points (device).  group: 1 = G1, 2 = G2.
extern "C" int zkb_debug_msm_batch(zkb_ctx* ctx, int group, const void* bases, size_t offset, const void* scalars_dev, size_t n,
                                   size_t stride, int batch, void* out_affine_dev) {
  if (!ctx) return ZKB_ERR_INVALID_ARG;
  if ((group != 1 && group != 2) || !bases || !out_affine_dev) ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "zkb_debug_msm_batch: bad argument");
  ZKB_ON_DEVICE(ctx);
  void* part = nullptr;
  const size_t psz = group == 1 ? sizeof(XYZZ<Fq>) : sizeof(XYZZ<Fq2>);
  CUDA_TRY(ctx, ctx->bpart.reserve(size_t(batch > 0 ? batch : 1) * psz));
  part = ctx->bpart.p;
  return group == 1 ? msm_batch_dev_impl<Fq>(ctx, static_cast<const zkb_g1_bases*>(bases), offset, scalars_dev, n, stride, batch, out_affine_dev, part)
                    : msm_batch_dev_impl<Fq2>(ctx, static_cast<const zkb_g2_bases*>(bases), offset, scalars_dev, n, stride, batch, out_affine_dev, part);
}

// parity hook for the comb-table MSM (msm.cuh comb_*): builds the table of every digit multiple for window width c on `bases`
// (kept on the handle) and runs `batch` MSMs through it -> batch canonical affine points (device).
extern "C" int zkb_debug_msm_comb(zkb_ctx* ctx, int group, void* bases, size_t offset, const void* scalars_dev, size_t n, size_t stride,
                                  int batch, int c, void* out_affine_dev) {
  if (!ctx) return ZKB_ERR_INVALID_ARG;
  if ((group != 1 && group != 2) || !bases || !out_affine_dev || batch < 1) ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "zkb_debug_msm_comb: bad argument");
  ZKB_ON_DEVICE(ctx);
  const size_t psz = group == 1 ? sizeof(XYZZ<Fq>) : sizeof(XYZZ<Fq2>);
  CUDA_TRY(ctx, ctx->bpart.reserve(size_t(batch) * psz));
  if (group == 1) {
    ZKB_TRY(bases_build_comb<Fq>(ctx, static_cast<zkb_g1_bases*>(bases), c));
    return msm_comb_dev_impl<Fq>(ctx, static_cast<zkb_g1_bases*>(bases), offset, scalars_dev, n, stride, batch, ctx->bpart.p, out_affine_dev);
  }
  ZKB_TRY(bases_build_comb<Fq2>(ctx, static_cast<zkb_g2_bases*>(bases), c));
  return msm_comb_dev_impl<Fq2>(ctx, static_cast<zkb_g2_bases*>(bases), offset, scalars_dev, n, stride, batch, ctx->bpart.p, out_affine_dev);
}

extern "C" int zkb_debug_msm_entries(zkb_ctx* ctx, const zkb_g1_bases* bases, size_t offset, const void* scalars_dev, size_t n, size_t stride,
                                     int batch, void* out_keys_dev, void* out_vals_dev, void* out_count_dev) {
  return msm_entries_debug_impl<Fq>(ctx, bases, offset, scalars_dev, n, stride, batch, out_keys_dev, out_vals_dev, out_count_dev);
}

// =============================================================================================== NTT
extern "C" int zkb_ntt_dev(zkb_ctx* ctx, const void* in_dev, void* out_dev, int log_n, int direction, int coset) {
  if (!ctx) return ZKB_ERR_INVALID_ARG;
  if (!in_dev || !out_dev) ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "ntt: null buffer");
  ZKB_ON_DEVICE(ctx);
  return ntt_dev_impl(ctx, static_cast<const Fr*>(in_dev), static_cast<Fr*>(out_dev), log_n, direction, coset);
}

extern "C" int zkb_ntt(zkb_ctx* ctx, const uint8_t* in_host, uint8_t* out_host, int log_n, int direction, int coset) {
  if (!ctx) return ZKB_ERR_INVALID_ARG;
  if (!in_host || !out_host) ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "ntt: null buffer");
  if (log_n < 0 || log_n > 28) ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "ntt: log_n %d outside [0, 28]", log_n);
  ZKB_ON_DEVICE(ctx);
  size_t bytes = (size_t(1) << log_n) * 32;
  CUDA_TRY(ctx, ctx->scal.reserve(bytes));
  CUDA_TRY(ctx, ctx->tmp2.reserve(bytes));
  CUDA_TRY(ctx, cudaMemcpyAsync(ctx->scal.p, in_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
  ZKB_TRY(ntt_dev_impl(ctx, ctx->scal.as<Fr>(), ctx->tmp2.as<Fr>(), log_n, direction, coset));
  CUDA_TRY(ctx, cudaMemcpyAsync(out_host, ctx->tmp2.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return ZKB_OK;
}

// =============================================================================================== L2-circuit Poseidon, batched
static int poseidon_params(zkb_ctx* ctx, std::vector<uint8_t>& buf) {
  buf.resize((64 * 3 + 9) * 32);
  if (zkb_l2_poseidon_params(buf.data(), buf.data() + 64 * 3 * 32) != ZKB_OK) ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "poseidon parameters unavailable");
  return ZKB_OK;
}

extern "C" int zkb_l2_poseidon_hash_batch_dev(zkb_ctx* ctx, int arity, const void* in_dev, size_t n, void* out_dev) {
  if (!ctx) return ZKB_ERR_INVALID_ARG;
  if (arity < 0 || arity > 3 || (n && (!out_dev || (arity && !in_dev)))) ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "zkb_l2_poseidon_hash_batch: arity %d outside [0, 3] or null buffer", arity);
  ZKB_ON_DEVICE(ctx);
  std::vector<uint8_t> prm;
  try {
    ZKB_TRY(poseidon_params(ctx, prm));
  } catch (const std::bad_alloc&) {
    ZKB_FAIL(ctx, ZKB_ERR_OOM, "zkb_l2_poseidon_hash_batch: host allocation failed");
  }
  return poseidon_hash_dev(ctx, prm.data(), arity, static_cast<const Fr*>(in_dev), n, static_cast<Fr*>(out_dev));
}

extern "C" int zkb_l2_poseidon_hash_batch(zkb_ctx* ctx, int arity, const uint8_t* in_host, size_t n, uint8_t* out_host) {
  if (!ctx) return ZKB_ERR_INVALID_ARG;
  if (arity < 0 || arity > 3 || (n && (!out_host || (arity && !in_host)))) ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "zkb_l2_poseidon_hash_batch: arity %d outside [0, 3] or null buffer", arity);
  if (!n) return ZKB_OK;
  ZKB_ON_DEVICE(ctx);
  CUDA_TRY(ctx, ctx->tmp0.reserve(n * size_t(arity ? arity : 1) * 32));
  CUDA_TRY(ctx, ctx->tmp2.reserve(n * 32));
  if (arity) CUDA_TRY(ctx, cudaMemcpyAsync(ctx->tmp0.p, in_host, n * size_t(arity) * 32, cudaMemcpyHostToDevice, ctx->stream));
  ZKB_TRY(zkb_l2_poseidon_hash_batch_dev(ctx, arity, ctx->tmp0.p, n, ctx->tmp2.p));
  CUDA_TRY(ctx, cudaMemcpyAsync(out_host, ctx->tmp2.p, n * 32, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return ZKB_OK;
}

// =============================================================================================== MiMC-7 / account Merkle tree
extern "C" int zkb_mimc_hash_dev(zkb_ctx* ctx, int arity, const void* in_dev, size_t n, void* out_dev) {
  if (!ctx) return ZKB_ERR_INVALID_ARG;
  if (arity < 1 || arity > 6 || (n && (!in_dev || !out_dev))) ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "zkb_mimc_hash: arity %d outside [1, 6] or null buffer", arity);
  ZKB_ON_DEVICE(ctx);
  ZKB_TRY(clear_flag(ctx));
  ZKB_TRY(mimc_hash_dev(ctx, arity, static_cast<const Fr*>(in_dev), n, static_cast<Fr*>(out_dev)));
  return ZKB_OK;   // asynchronous: a non-canonical input surfaces at the next call that checks the flag (the host variant does)
}

extern "C" int zkb_mimc_hash(zkb_ctx* ctx, int arity, const uint8_t* in_host, size_t n, uint8_t* out_host) {
  if (!ctx) return ZKB_ERR_INVALID_ARG;
  if (arity < 1 || arity > 6 || (n && (!in_host || !out_host))) ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "zkb_mimc_hash: arity %d outside [1, 6] or null buffer", arity);
  if (!n) return ZKB_OK;
  ZKB_ON_DEVICE(ctx);
  CUDA_TRY(ctx, ctx->tmp0.reserve(n * size_t(arity) * 32));
  CUDA_TRY(ctx, ctx->tmp2.reserve(n * 32));
  ZKB_TRY(clear_flag(ctx));
  CUDA_TRY(ctx, cudaMemcpyAsync(ctx->tmp0.p, in_host, n * size_t(arity) * 32, cudaMemcpyHostToDevice, ctx->stream));
  ZKB_TRY(mimc_hash_dev(ctx, arity, ctx->tmp0.as<Fr>(), n, ctx->tmp2.as<Fr>()));
  CUDA_TRY(ctx, cudaMemcpyAsync(out_host, ctx->tmp2.p, n * 32, cudaMemcpyDeviceToHost, ctx->stream));
  return check_flag(ctx, "zkb_mimc_hash");
}

extern "C" int zkb_mimc_merkle_roots_dev(zkb_ctx* ctx, const void* leaves_dev, const void* siblings_dev, const uint8_t* bits_dev, size_t n,
                                         int depth, void* out_dev) {
  if (!ctx) return ZKB_ERR_INVALID_ARG;
  if (depth < 0 || depth > 64 || (n && (!leaves_dev || !out_dev || (depth && (!siblings_dev || !bits_dev)))))
    ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "zkb_mimc_merkle_roots: depth %d outside [0, 64] or null buffer", depth);
  ZKB_ON_DEVICE(ctx);
  ZKB_TRY(clear_flag(ctx));
  return mimc_merkle_roots_dev(ctx, static_cast<const Fr*>(leaves_dev), static_cast<const Fr*>(siblings_dev), bits_dev, n, depth,
                               static_cast<Fr*>(out_dev));
}

extern "C" int zkb_mimc_merkle_roots(zkb_ctx* ctx, const uint8_t* leaves, const uint8_t* siblings, const uint8_t* bits, size_t n, int depth,
                                     uint8_t* out) {
  if (!ctx) return ZKB_ERR_INVALID_ARG;
  if (depth < 0 || depth > 64 || (n && (!leaves || !out || (depth && (!siblings || !bits)))))
    ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "zkb_mimc_merkle_roots: depth %d outside [0, 64] or null buffer", depth);
  if (!n) return ZKB_OK;
  ZKB_ON_DEVICE(ctx);
  const size_t nd = n * size_t(depth);
  CUDA_TRY(ctx, ctx->tmp0.reserve(n * 32));
  CUDA_TRY(ctx, ctx->tmp1.reserve(nd * 32 + 32));
  CUDA_TRY(ctx, ctx->tmp2.reserve(n * 32));
  CUDA_TRY(ctx, ctx->scal.reserve(nd + 32));
  ZKB_TRY(clear_flag(ctx));
  CUDA_TRY(ctx, cudaMemcpyAsync(ctx->tmp0.p, leaves, n * 32, cudaMemcpyHostToDevice, ctx->stream));
  if (nd) {
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->tmp1.p, siblings, nd * 32, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->scal.p, bits, nd, cudaMemcpyHostToDevice, ctx->stream));
  }
  ZKB_TRY(mimc_merkle_roots_dev(ctx, ctx->tmp0.as<Fr>(), ctx->tmp1.as<Fr>(), ctx->scal.as<uint8_t>(), n, depth, ctx->tmp2.as<Fr>()));
  CUDA_TRY(ctx, cudaMemcpyAsync(out, ctx->tmp2.p, n * 32, cudaMemcpyDeviceToHost, ctx->stream));
  return check_flag(ctx, "zkb_mimc_merkle_roots");
}

// =============================================================================================== R1CS + witness map
namespace {

void csr_free(CsrDev& c) {
  g_alloc_epoch.fetch_add(1, std::memory_order_relaxed);
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
    ZKB_TRY(fr_to_mont(ctx, ctx->tmp0.as<Fr>(), d.coeff, nnz));
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
  ZKB_ON_DEVICE(ctx);
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
  DeviceGuard dg(m->device);
  csr_free(m->a);
  csr_free(m->b);
  csr_free(m->c);
  delete m;
}

extern "C" int zkb_r1cs_log_domain(const zkb_r1cs* m) { return m ? m->log_domain : -1; }
extern "C" uint64_t zkb_r1cs_num_variables(const zkb_r1cs* m) { return m ? m->ni + m->nw : 0; }
extern "C" uint64_t zkb_r1cs_num_constraints(const zkb_r1cs* m) { return m ? m->nc : 0; }


static int reserve_prove_bufs(zkb_ctx* ctx, size_t n, size_t nv, size_t nw) {
  CUDA_TRY(ctx, ctx->pz.reserve(nv * 32));
  CUDA_TRY(ctx, ctx->pzm.reserve(nv * 32));
  CUDA_TRY(ctx, ctx->pwa.reserve(n * 32));
  CUDA_TRY(ctx, ctx->pwb.reserve(n * 32));
  CUDA_TRY(ctx, ctx->pwc.reserve(n * 32));
  CUDA_TRY(ctx, ctx->ph.reserve(n * 32));
  CUDA_TRY(ctx, ctx->pza.reserve((nv + 2) * 32));
  CUDA_TRY(ctx, ctx->pzb.reserve((nv + 3) * 32));
  CUDA_TRY(ctx, ctx->pzl.reserve((nw + 1) * 32));
  CUDA_TRY(ctx, ctx->prs.reserve(64));
  CUDA_TRY(ctx, ctx->ppts.reserve(4 * sizeof(XYZZ<Fq>) + sizeof(XYZZ<Fq2>) + 64 + 128 + 64 + sizeof(XYZZ<Fq>)));
  CUDA_TRY(ctx, ctx->pzsa.reserve((nv + 2) * 32));
  CUDA_TRY(ctx, ctx->pzrb.reserve((nv + 2) * 32));
  return ZKB_OK;
}

extern "C" int zkb_witness_map(zkb_ctx* ctx, const zkb_r1cs* m, const uint8_t* z_host, uint8_t* h_out_host) {
  if (!ctx) return ZKB_ERR_INVALID_ARG;
  if (!m || !z_host || !h_out_host) ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "zkb_witness_map: null argument");
  if (m->device != ctx->device) ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "zkb_witness_map: matrices on another device");
  ZKB_ON_DEVICE(ctx);
  const size_t n = size_t(1) << m->log_domain, nv = m->ni + m->nw;
  ZKB_TRY(reserve_prove_bufs(ctx, n, nv, m->nw));
  CUDA_TRY(ctx, cudaMemcpyAsync(ctx->pz.p, z_host, nv * 32, cudaMemcpyHostToDevice, ctx->stream));
  WitnessBufs w{ctx->pz.as<Fr>(), ctx->pzm.as<Fr>(), ctx->pwa.as<Fr>(), ctx->pwb.as<Fr>(), ctx->pwc.as<Fr>()};
  ZKB_TRY(witness_map_dev(ctx, m->a, m->b, m->c, m->nc, m->ni, m->nw, m->log_domain, w, ctx->ph.as<Fr>()));
  CUDA_TRY(ctx, cudaMemcpyAsync(h_out_host, ctx->ph.p, n * 32, cudaMemcpyDeviceToHost, ctx->stream));
  return check_flag(ctx, "witness assignment");
}

// =============================================================================================== proving key + prove
namespace {

template <class F>
int load_ext(zkb_ctx* ctx, const uint8_t* q, size_t qlen, bool rotate_first, const uint8_t* const* extra, int nextra,
             int validate, size_t lo, size_t cnt, typename GroupOf<F>::Bases** out) {
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
  return bases_load_impl<F>(ctx, buf.data() + lo * sz, cnt, validate, out);  // [lo, lo + cnt) of the extended vector
}

}  // namespace

extern "C" void zkb_pk_free(zkb_pk* pk) {
  if (!pk) return;
  zkb_g1_bases_free(pk->a_ext);
  zkb_g1_bases_free(pk->b1_ext);
  zkb_g1_bases_free(pk->l_ext);
  zkb_g1_bases_free(pk->h);
  zkb_g2_bases_free(pk->b2_ext);
  if (pk->fb_delta1) cudaFree(pk->fb_delta1);
  if (pk->fb_delta2) cudaFree(pk->fb_delta2);
  delete pk;
}

namespace {

// contiguous split with the remainder on the lowest shards (zelana_b200/multi.py shard_range)
void shard_span(size_t n, int shard, int world, size_t* lo, size_t* cnt) {
  size_t base = n / world, rem = n % world;
  *lo = size_t(shard) * base + (size_t(shard) < rem ? size_t(shard) : rem);
  *cnt = base + (size_t(shard) < rem ? 1 : 0);
}

int pk_load_common(zkb_ctx* ctx, const zkb_pk_desc* d, int validate, int shard, int world, zkb_pk** out) {
  if (!ctx) return ZKB_ERR_INVALID_ARG;
  if (!d || !out) ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "zkb_pk_load: null argument");
  *out = nullptr;
  if (world < 1 || shard < 0 || shard >= world) ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "zkb_pk_load: shard %d of %d", shard, world);
  if (!d->alpha_g1 || !d->beta_g1 || !d->beta_g2 || !d->delta_g1 || !d->delta_g2 || !d->a_query || !d->b_g1_query ||
      !d->b_g2_query || (!d->h_query && d->h_len) || (!d->l_query && d->l_len))
    ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "zkb_pk_load: null key component");
  if (d->a_len < 1 || d->b_g1_len != d->a_len || d->b_g2_len != d->a_len)
    ZKB_FAIL(ctx, ZKB_ERR_SHAPE, "zkb_pk_load: a/b_g1/b_g2 query lengths %zu/%zu/%zu disagree", d->a_len, d->b_g1_len, d->b_g2_len);
  if (d->l_len > d->a_len - 1) ZKB_FAIL(ctx, ZKB_ERR_SHAPE, "zkb_pk_load: l_query longer than the witness");
  ZKB_ON_DEVICE(ctx);
  zkb_pk* pk = new (std::nothrow) zkb_pk();
  if (!pk) ZKB_FAIL(ctx, ZKB_ERR_OOM, "zkb_pk_load: host allocation failed");
  pk->device = ctx->device;
  pk->nv = d->a_len;
  pk->nw = d->l_len;
  pk->nh = d->h_len;
  pk->shard = shard;
  pk->world = world;
  size_t cnt_a, cnt_l, cnt_h;
  shard_span(d->a_len + 2, shard, world, &pk->off_a, &cnt_a);
  shard_span(d->l_len + 1, shard, world, &pk->off_l, &cnt_l);
  shard_span(d->h_len, shard, world, &pk->off_h, &cnt_h);
  const uint8_t* ea[2] = {d->alpha_g1, d->delta_g1};
  const uint8_t* eb1[2] = {d->beta_g1, d->delta_g1};
  const uint8_t* eb2[2] = {d->beta_g2, d->delta_g2};
  const uint8_t* el[1] = {d->delta_g1};
  int s = load_ext<Fq>(ctx, d->a_query, d->a_len, true, ea, 2, validate, pk->off_a, cnt_a, &pk->a_ext);
  if (s == ZKB_OK) s = load_ext<Fq>(ctx, d->b_g1_query, d->b_g1_len, true, eb1, 2, validate, pk->off_a, cnt_a, &pk->b1_ext);
  if (s == ZKB_OK) s = load_ext<Fq2>(ctx, d->b_g2_query, d->b_g2_len, true, eb2, 2, validate, pk->off_a, cnt_a, &pk->b2_ext);
  if (s == ZKB_OK) s = load_ext<Fq>(ctx, d->l_query, d->l_len, false, el, 1, validate, pk->off_l, cnt_l, &pk->l_ext);
  if (s == ZKB_OK) s = load_ext<Fq>(ctx, d->h_query, d->h_len, false, nullptr, 0, validate, pk->off_h, cnt_h, &pk->h);
  if (s != ZKB_OK) {
    zkb_pk_free(pk);
    return s;
  }
  *out = pk;
  return ZKB_OK;
}

}  // namespace

extern "C" int zkb_pk_load(zkb_ctx* ctx, const zkb_pk_desc* d, int validate, zkb_pk** out) {
  return pk_load_common(ctx, d, validate, 0, 1, out);
}

// Shard `shard` of `world` of the same key: only the range shard_span(len) of every (extended) query vector is uploaded.
extern "C" int zkb_pk_load_shard(zkb_ctx* ctx, const zkb_pk_desc* d, int validate, int shard, int world, zkb_pk** out) {
  return pk_load_common(ctx, d, validate, shard, world, out);
}

// ProvingKey::<Bn254>::deserialize_compressed (Groth16Prover::from_bytes, prover.rs:263-277): ark-serialize layout
//   vk { alpha_g1, beta_g2, gamma_g2, delta_g2, gamma_abc_g1: Vec<G1> }, beta_g1, delta_g1, a_query, b_g1_query,
//   b_g2_query: Vec<G2>, h_query, l_query      (Vec<T> = u64 LE length || elements; G1 32 B, G2 64 B, flags in the top bits).
// Every point is decompressed (square root, sign choice) and validated (on curve; G2: prime-order subgroup) on the GPU.
extern "C" int zkb_pk_load_compressed(zkb_ctx* ctx, const uint8_t* bytes, size_t len, int validate, zkb_pk** out) {
  if (!ctx) return ZKB_ERR_INVALID_ARG;
  if (!bytes || !out) ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "zkb_pk_load_compressed: null argument");
  *out = nullptr;
  size_t off = 0;
  auto need = [&](size_t k) { return off + k <= len; };
  auto fixed = [&](size_t k, const uint8_t** p) {
    if (!need(k)) return false;
    *p = bytes + off;
    off += k;
    return true;
  };
  auto vec = [&](size_t elem, const uint8_t** p, size_t* count) {
    if (!need(8)) return false;
    uint64_t n64;
    memcpy(&n64, bytes + off, 8);
    off += 8;
    if (n64 > (len - off) / elem) return false;
    *p = bytes + off;
    *count = size_t(n64);
    off += size_t(n64) * elem;
    return true;
  };
  const uint8_t *alpha, *beta2, *gamma2, *delta2, *abc, *beta1, *delta1, *aq, *b1q, *b2q, *hq, *lq;
  size_t n_abc, n_a, n_b1, n_b2, n_h, n_l;
  bool ok = fixed(32, &alpha) && fixed(64, &beta2) && fixed(64, &gamma2) && fixed(64, &delta2) && vec(32, &abc, &n_abc) &&
            fixed(32, &beta1) && fixed(32, &delta1) && vec(32, &aq, &n_a) && vec(32, &b1q, &n_b1) && vec(64, &b2q, &n_b2) &&
            vec(32, &hq, &n_h) && vec(32, &lq, &n_l);
  if (!ok) ZKB_FAIL(ctx, ZKB_ERR_SHAPE, "zkb_pk_load_compressed: truncated proving key (%zu bytes)", len);
  if (off != len) ZKB_FAIL(ctx, ZKB_ERR_SHAPE, "zkb_pk_load_compressed: %zu trailing bytes after the proving key", len - off);
  if (n_a < 1 || n_b1 != n_a || n_b2 != n_a)
    ZKB_FAIL(ctx, ZKB_ERR_SHAPE, "zkb_pk_load_compressed: a/b_g1/b_g2 query lengths %zu/%zu/%zu disagree", n_a, n_b1, n_b2);
  if (n_l > n_a - 1) ZKB_FAIL(ctx, ZKB_ERR_SHAPE, "zkb_pk_load_compressed: l_query longer than the witness");
  ZKB_ON_DEVICE(ctx);
  zkb_pk* pk = new (std::nothrow) zkb_pk();
  if (!pk) ZKB_FAIL(ctx, ZKB_ERR_OOM, "zkb_pk_load_compressed: host allocation failed");
  pk->device = ctx->device;
  pk->nv = n_a;
  pk->nw = n_l;
  pk->nh = n_h;
  // q[1..] || q[0] || extras (or q || extras), in compressed form
  auto ext = [&](const uint8_t* q, size_t qlen, size_t elem, bool rotate, std::initializer_list<const uint8_t*> extra) {
    std::vector<uint8_t> buf((qlen + extra.size()) * elem);
    size_t o = 0;
    if (rotate && qlen) {
      memcpy(buf.data(), q + elem, (qlen - 1) * elem);
      o = (qlen - 1) * elem;
      memcpy(buf.data() + o, q, elem);
      o += elem;
    } else if (qlen) {
      memcpy(buf.data(), q, qlen * elem);
      o = qlen * elem;
    }
    for (const uint8_t* e : extra) {
      memcpy(buf.data() + o, e, elem);
      o += elem;
    }
    return buf;
  };
  int s = ZKB_OK;
  try {
    {
      auto b = ext(aq, n_a, 32, true, {alpha, delta1});
      s = bases_load_compressed_impl<Fq>(ctx, b.data(), n_a + 2, validate, &pk->a_ext);
    }
    if (s == ZKB_OK) {
      auto b = ext(b1q, n_b1, 32, true, {beta1, delta1});
      s = bases_load_compressed_impl<Fq>(ctx, b.data(), n_b1 + 2, validate, &pk->b1_ext);
    }
    if (s == ZKB_OK) {
      auto b = ext(b2q, n_b2, 64, true, {beta2, delta2});
      s = bases_load_compressed_impl<Fq2>(ctx, b.data(), n_b2 + 2, validate, &pk->b2_ext);
    }
    if (s == ZKB_OK) {
      auto b = ext(lq, n_l, 32, false, {delta1});
      s = bases_load_compressed_impl<Fq>(ctx, b.data(), n_l + 1, validate, &pk->l_ext);
    }
    if (s == ZKB_OK) s = bases_load_compressed_impl<Fq>(ctx, hq, n_h, validate, &pk->h);
    // the prover never reads gamma_g2 / gamma_abc_g1, but deserialize_compressed validates them: so do we
    if (s == ZKB_OK && validate) {
      zkb_g2_bases* g2 = nullptr;
      s = bases_load_compressed_impl<Fq2>(ctx, gamma2, 1, validate, &g2);
      bases_free_impl<Fq2>(g2);
      zkb_g1_bases* g1 = nullptr;
      if (s == ZKB_OK) s = bases_load_compressed_impl<Fq>(ctx, abc, n_abc, validate, &g1);
      bases_free_impl<Fq>(g1);
    }
  } catch (const std::bad_alloc&) {
    zkb_pk_free(pk);
    ZKB_FAIL(ctx, ZKB_ERR_OOM, "zkb_pk_load_compressed: host staging allocation failed");
  }
  if (s != ZKB_OK) {
    zkb_pk_free(pk);
    return s;
  }
  *out = pk;
  return ZKB_OK;
}

// Benchmark-only key: every query vector is [k_i] G for caller-supplied device scalars, so that a forge-sized key
//  can be fabricated in seconds on the GPU.  Proofs made with it do not verify -- the timing
// of zkb_prove does not depend on which curve points the key holds; correctness is covered by the real-key tests.
extern "C" int zkb_pk_synthetic_shard(zkb_ctx* ctx, size_t num_vars, size_t num_witness, size_t h_len, const void* k_dev,
                                      size_t k_len, int shard, int world, zkb_pk** out) {
  if (!ctx) return ZKB_ERR_INVALID_ARG;
  if (!out || !k_dev || num_vars < 1 || num_witness > num_vars - 1 || world < 1 || shard < 0 || shard >= world)
    ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "zkb_pk_synthetic: bad argument");
  size_t need = num_vars + 2 > h_len ? num_vars + 2 : h_len;
  if (k_len < need + 4) ZKB_FAIL(ctx, ZKB_ERR_SHAPE, "zkb_pk_synthetic: need %zu scalars, got %zu", need + 4, k_len);
  *out = nullptr;
  ZKB_ON_DEVICE(ctx);
  zkb_pk* pk = new (std::nothrow) zkb_pk();
  if (!pk) ZKB_FAIL(ctx, ZKB_ERR_OOM, "zkb_pk_synthetic: host allocation failed");
  pk->device = ctx->device;
  pk->nv = num_vars;
  pk->nw = num_witness;
  pk->nh = h_len;
  pk->shard = shard;
  pk->world = world;
  size_t cnt_a, cnt_l, cnt_h;
  shard_span(num_vars + 2, shard, world, &pk->off_a, &cnt_a);
  shard_span(num_witness + 1, shard, world, &pk->off_l, &cnt_l);
  shard_span(h_len, shard, world, &pk->off_h, &cnt_h);
  // point i of the vector v is [k[i + v]] G, whatever the sharding
  const char* k = static_cast<const char*>(k_dev);
  int s = bases_generate_impl<Fq>(ctx, k + 32 * pk->off_a, cnt_a, &pk->a_ext);
  if (s == ZKB_OK) s = bases_generate_impl<Fq>(ctx, k + 32 * (pk->off_a + 1), cnt_a, &pk->b1_ext);
  if (s == ZKB_OK) s = bases_generate_impl<Fq2>(ctx, k + 32 * (pk->off_a + 2), cnt_a, &pk->b2_ext);
  if (s == ZKB_OK) s = bases_generate_impl<Fq>(ctx, k + 32 * (pk->off_l + 3), cnt_l, &pk->l_ext);
  if (s == ZKB_OK) s = bases_generate_impl<Fq>(ctx, k + 32 * (pk->off_h + 4), cnt_h, &pk->h);
  if (s != ZKB_OK) {
    zkb_pk_free(pk);
    return s;
  }
  *out = pk;
  return ZKB_OK;
}

extern "C" int zkb_pk_synthetic(zkb_ctx* ctx, size_t num_vars, size_t num_witness, size_t h_len, const void* k_dev,
                                size_t k_len, zkb_pk** out) {
  return zkb_pk_synthetic_shard(ctx, num_vars, num_witness, h_len, k_dev, k_len, 0, 1, out);
}

// =============================================================================================== trusted setup
namespace {

// CSR (rows = constraints) -> CSC (columns = variables) on the host, coefficients kept as 32-byte canonical records
struct CscHost {
  std::vector<uint64_t> col_ptr;
  std::vector<uint32_t> row;
  std::vector<uint8_t> coeff;
};

int csr_to_csc(zkb_ctx* ctx, const zkb_csr& m, uint64_t nc, uint64_t nv, CscHost& out) {
  if (!m.row_ptr) ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "zkb_setup: null row_ptr");
  const uint64_t nnz = m.row_ptr[nc];
  if (nnz && (!m.col || !m.coeff)) ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "zkb_setup: null col/coeff");
  out.col_ptr.assign(nv + 1, 0);
  for (uint64_t k = 0; k < nnz; k++) {
    if (m.col[k] >= nv) ZKB_FAIL(ctx, ZKB_ERR_SHAPE, "zkb_setup: column %u >= %llu variables", m.col[k], (unsigned long long)nv);
    out.col_ptr[m.col[k] + 1]++;
  }
  for (uint64_t j = 0; j < nv; j++) out.col_ptr[j + 1] += out.col_ptr[j];
  out.row.resize(nnz ? nnz : 1);
  out.coeff.resize((nnz ? nnz : 1) * 32);
  std::vector<uint64_t> cursor(out.col_ptr.begin(), out.col_ptr.end() - 1);
  for (uint64_t i = 0; i < nc; i++) {
    if (m.row_ptr[i + 1] < m.row_ptr[i]) ZKB_FAIL(ctx, ZKB_ERR_SHAPE, "zkb_setup: row_ptr not monotone at row %llu", (unsigned long long)i);
    for (uint64_t k = m.row_ptr[i]; k < m.row_ptr[i + 1]; k++) {
      uint64_t dst = cursor[m.col[k]]++;
      out.row[dst] = uint32_t(i);
      memcpy(&out.coeff[dst * 32], m.coeff + k * 32, 32);
    }
  }
  return ZKB_OK;
}

}  // namespace

// Groth16 parameter generation (ark-groth16 generate_parameters_with_qap, reached from prover/src/bin/keygen.rs:87-91 through
// Groth16::circuit_specific_setup): Lagrange coefficients at tau, QAP column evaluations, then fixed-base batch multiplications
// of the (random) generators -- all on the GPU.  The randomness (alpha, beta, gamma, delta, tau, g1, g2) is the caller's:
// zelana_b200/keygen.py draws it from StdRng exactly as arkworks does.
extern "C" int zkb_setup(zkb_ctx* ctx, const zkb_r1cs_desc* d, const zkb_setup_params* prm, const zkb_setup_out* o) {
  if (!ctx) return ZKB_ERR_INVALID_ARG;
  if (!d || !prm || !o) ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "zkb_setup: null argument");
  if (d->num_instance < 1) ZKB_FAIL(ctx, ZKB_ERR_SHAPE, "zkb_setup: num_instance must include the constant ONE");
  if (!o->alpha_g1 || !o->beta_g1 || !o->delta_g1 || !o->beta_g2 || !o->gamma_g2 || !o->delta_g2 || !o->gamma_abc_g1 ||
      !o->a_query || !o->b_g1_query || !o->b_g2_query || !o->h_query || (!o->l_query && d->num_witness))
    ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "zkb_setup: null output buffer");
  const uint64_t nc = d->num_constraints, ni = d->num_instance, nw = d->num_witness, nv = ni + nw;
  int lg = 0;
  while ((uint64_t(1) << lg) < nc + ni) lg++;
  if (lg > 28) ZKB_FAIL(ctx, ZKB_ERR_SHAPE, "zkb_setup: domain 2^%d exceeds Fr two-adicity 28", lg);
  const size_t n = size_t(1) << lg;
  ZKB_ON_DEVICE(ctx);
  cudaStream_t st = ctx->stream;

  CscHost csc[3];
  try {
    ZKB_TRY(csr_to_csc(ctx, d->a, nc, nv, csc[0]));
    ZKB_TRY(csr_to_csc(ctx, d->b, nc, nv, csc[1]));
    ZKB_TRY(csr_to_csc(ctx, d->c, nc, nv, csc[2]));
  } catch (const std::bad_alloc&) {
    ZKB_FAIL(ctx, ZKB_ERR_OOM, "zkb_setup: host allocation failed");
  }
  // device buffers, freed on every exit path
  std::vector<void*> owned;
  auto dmalloc = [&](size_t bytes) -> void* {
    void* p = nullptr;
    if (cudaMalloc(&p, bytes ? bytes : 32) != cudaSuccess) {
      cudaGetLastError();
      return nullptr;
    }
    owned.push_back(p);
    return p;
  };
  auto run = [&]() -> int {
    const uint64_t* col_ptr[3];
    const uint32_t* row[3];
    const Fr* coeff[3];
    for (int t = 0; t < 3; t++) {
      size_t nnz = csc[t].col_ptr[nv];
      uint64_t* cp = static_cast<uint64_t*>(dmalloc((nv + 1) * 8));
      uint32_t* rw = static_cast<uint32_t*>(dmalloc(nnz * 4));
      Fr* raw = static_cast<Fr*>(dmalloc(nnz * 32));
      Fr* co = static_cast<Fr*>(dmalloc(nnz * 32));
      if (!cp || !rw || !raw || !co) ZKB_FAIL(ctx, ZKB_ERR_OOM, "zkb_setup: device allocation failed");
      CUDA_TRY(ctx, cudaMemcpyAsync(cp, csc[t].col_ptr.data(), (nv + 1) * 8, cudaMemcpyHostToDevice, st));
      if (nnz) {
        CUDA_TRY(ctx, cudaMemcpyAsync(rw, csc[t].row.data(), nnz * 4, cudaMemcpyHostToDevice, st));
        CUDA_TRY(ctx, cudaMemcpyAsync(raw, csc[t].coeff.data(), nnz * 32, cudaMemcpyHostToDevice, st));
        ZKB_TRY(clear_flag(ctx));
        ZKB_TRY(fr_to_mont(ctx, raw, co, nnz));
        ZKB_TRY(check_flag(ctx, "zkb_setup: matrix coefficient"));
      }
      col_ptr[t] = cp;
      row[t] = rw;
      coeff[t] = co;
    }
    Fr* in5 = static_cast<Fr*>(dmalloc(5 * 32));
    Fr* consts = static_cast<Fr*>(dmalloc(8 * 32));
    Fr* u = static_cast<Fr*>(dmalloc(n * 32));
    Fr* a_s = static_cast<Fr*>(dmalloc(nv * 32));
    Fr* b_s = static_cast<Fr*>(dmalloc(nv * 32));
    Fr* abc_s = static_cast<Fr*>(dmalloc(nv * 32));
    Fr* h_s = static_cast<Fr*>(dmalloc(n * 32));
    Fr* k3 = static_cast<Fr*>(dmalloc(6 * 32));
    if (!in5 || !consts || !u || !a_s || !b_s || !abc_s || !h_s || !k3) ZKB_FAIL(ctx, ZKB_ERR_OOM, "zkb_setup: device allocation failed");
    uint8_t host5[5 * 32];
    memcpy(host5, prm->tau, 32);
    memcpy(host5 + 32, prm->alpha, 32);
    memcpy(host5 + 64, prm->beta, 32);
    memcpy(host5 + 96, prm->gamma, 32);
    memcpy(host5 + 128, prm->delta, 32);
    CUDA_TRY(ctx, cudaMemcpyAsync(in5, host5, sizeof(host5), cudaMemcpyHostToDevice, st));
    ZKB_TRY(setup_scalars_dev(ctx, col_ptr, row, coeff, nc, ni, nw, lg, in5, consts, u, a_s, b_s, abc_s, h_s));
    {
      int h = 0;
      CUDA_TRY(ctx, cudaMemcpyAsync(&h, ctx->flag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
      CUDA_TRY(ctx, cudaStreamSynchronize(st));
      if (h == 1) ZKB_FAIL(ctx, ZKB_ERR_NOT_CANONICAL, "zkb_setup: a trapdoor value is >= the Fr modulus");
      if (h == 5) ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "zkb_setup: tau lies in the evaluation domain, or gamma / delta is zero");
    }
    // fixed constants: [alpha, beta, delta] on g1 ; [beta, gamma, delta] on g2
    uint8_t hostk[6 * 32];
    memcpy(hostk, prm->alpha, 32);
    memcpy(hostk + 32, prm->beta, 32);
    memcpy(hostk + 64, prm->delta, 32);
    memcpy(hostk + 96, prm->beta, 32);
    memcpy(hostk + 128, prm->gamma, 32);
    memcpy(hostk + 160, prm->delta, 32);
    CUDA_TRY(ctx, cudaMemcpyAsync(k3, hostk, sizeof(hostk), cudaMemcpyHostToDevice, st));
    uint8_t c1[3 * 64], c2[3 * 128];
    ZKB_TRY(fixed_base_batch<Fq>(ctx, prm->g1_generator, k3, 3, c1));
    ZKB_TRY(fixed_base_batch<Fq2>(ctx, prm->g2_generator, k3 + 3, 3, c2));
    memcpy(o->alpha_g1, c1, 64);
    memcpy(o->beta_g1, c1 + 64, 64);
    memcpy(o->delta_g1, c1 + 128, 64);
    memcpy(o->beta_g2, c2, 128);
    memcpy(o->gamma_g2, c2 + 128, 128);
    memcpy(o->delta_g2, c2 + 256, 128);
    ZKB_TRY(fixed_base_batch<Fq>(ctx, prm->g1_generator, a_s, nv, o->a_query));
    ZKB_TRY(fixed_base_batch<Fq>(ctx, prm->g1_generator, b_s, nv, o->b_g1_query));
    ZKB_TRY(fixed_base_batch<Fq2>(ctx, prm->g2_generator, b_s, nv, o->b_g2_query));
    ZKB_TRY(fixed_base_batch<Fq>(ctx, prm->g1_generator, abc_s, ni, o->gamma_abc_g1));
    ZKB_TRY(fixed_base_batch<Fq>(ctx, prm->g1_generator, abc_s + ni, nw, o->l_query));
    ZKB_TRY(fixed_base_batch<Fq>(ctx, prm->g1_generator, h_s, n - 1, o->h_query));
    return ZKB_OK;
  };
  int rc = run();
  cudaStreamSynchronize(st);
  for (void* p : owned) cudaFree(p);
  return rc;
}

namespace {

struct ProveOut {
  XYZZ<Fq>*pA, *pB1, *pL, *pH, *pSA;
  XYZZ<Fq2>* pB2;
  uint32_t *oA, *oB, *oC;
};

ProveOut prove_out(zkb_ctx* ctx) {
  char* pts = static_cast<char*>(ctx->ppts.p);
  ProveOut o;
  o.pA = reinterpret_cast<XYZZ<Fq>*>(pts);
  o.pB1 = o.pA + 1;
  o.pL = o.pA + 2;
  o.pH = o.pA + 3;
  o.pB2 = reinterpret_cast<XYZZ<Fq2>*>(pts + 4 * sizeof(XYZZ<Fq>));
  o.oA = reinterpret_cast<uint32_t*>(pts + 4 * sizeof(XYZZ<Fq>) + sizeof(XYZZ<Fq2>));
  o.oB = o.oA + 16;
  o.oC = o.oB + 32;
  o.pSA = reinterpret_cast<XYZZ<Fq>*>(o.oC + 16);
  return o;
}

// Everything of a prove that runs on the device, queued on the context's streams: z, r, s are already in pz / prs.
int prove_device_part(zkb_ctx* ctx, const zkb_pk* pk, const zkb_r1cs* m, bool partial, const char* who) {
  const size_t nv = m->ni + m->nw, nw = m->nw, ni = m->ni;
  const size_t n = size_t(1) << m->log_domain;
  cudaStream_t st = ctx->stream;
  Fr* rs = ctx->prs.as<Fr>();
  Fr* z = ctx->pz.as<Fr>();
  ProveOut o = prove_out(ctx);
  const size_t na = pk->a_ext->n;   // this shard's slice [off_a, off_a + na) of the nv + 2 extended scalars

  // Scalar vectors of the folded MSMs (they depend on z, r, s only):
  //   za = z[1..] || 1 || 1 || r   (A)        zb = z[1..] || 1 || 1 || s   (B1 and B2)        zl = aux || -(r s)   (L)
  Fr* za = ctx->pza.as<Fr>();
  Fr* zb = ctx->pzb.as<Fr>();
  Fr* zl = ctx->pzl.as<Fr>();
  if (nv > 1) {
    CUDA_TRY(ctx, cudaMemcpyAsync(za, z + 1, (nv - 1) * 32, cudaMemcpyDeviceToDevice, st));
    CUDA_TRY(ctx, cudaMemcpyAsync(zb, z + 1, (nv - 1) * 32, cudaMemcpyDeviceToDevice, st));
  }
  if (nw) CUDA_TRY(ctx, cudaMemcpyAsync(zl, z + ni, nw * 32, cudaMemcpyDeviceToDevice, st));
  ZKB_TRY(prove_tail_scalars(ctx, rs, rs + 1, za + (nv - 1), zl + nw));
  ZKB_TRY(prove_tail_scalars(ctx, rs + 1, rs + 1, zb + (nv - 1), zb + (nv + 2)));  // [1, 1, s]; the extra slot is scratch

  // A, B1, B2 and L need only those vectors, not h: each runs on its own lane (stream + MSM scratch) next to the witness map
  // and the H MSM of the main stream, so the latency-bound sort / reduction phases of one fill the SMs another leaves idle.
  if (!ctx->ev_inputs) {
    for (auto& lane : ctx->aux) {
      CUDA_TRY(ctx, cudaStreamCreateWithFlags(&lane.stream, cudaStreamNonBlocking));
      CUDA_TRY(ctx, cudaEventCreateWithFlags(&lane.done, cudaEventDisableTiming));
    }
    CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_inputs, cudaEventDisableTiming));
  }
  // Small circuits: fold s and r into the MSMs instead of multiplying their results -- s A = MSM(a_ext, s za) on a fifth lane,
  // r B1 = MSM(b1_ext, r zb) in place of B1 -- because the two 254-step double-and-add chains (1.9 ms on one thread each) are
  // half the latency of an L2-sized proof, while an extra 6 k-point MSM next to the others is free.  Large circuits keep
  // the chains: 2 ms is nothing next to a 2^21-point MSM, an extra one is not.
  const bool prescaled = !partial && nv <= (size_t(1) << 16);
  Fr* zsa = ctx->pzsa.as<Fr>();
  Fr* zrb = ctx->pzrb.as<Fr>();
  if (prescaled) {
    ZKB_TRY(fr_scale(ctx, za, rs + 1, zsa, nv + 2));
    ZKB_TRY(fr_scale(ctx, zb, rs, zrb, nv + 2));
  }
  CUDA_TRY(ctx, cudaEventRecord(ctx->ev_inputs, st));
  int lane_rc = ZKB_OK;
  bool lane_used[5] = {false, false, false, false, false};
  auto on_lane = [&](int idx, auto&& fn) {
    auto& lane = ctx->aux[idx];
    if (lane_rc != ZKB_OK) return;
    if (cudaStreamWaitEvent(lane.stream, ctx->ev_inputs, 0) != cudaSuccess) {
      lane_rc = ZKB_ERR_CUDA;
      return;
    }
    lane_used[idx] = true;
    std::swap(ctx->stream, lane.stream);
    std::swap(ctx->msm_ws, lane.ws);
    int rc = fn();
    if (rc == ZKB_OK && cudaEventRecord(lane.done, ctx->stream) != cudaSuccess) rc = ZKB_ERR_CUDA;
    std::swap(ctx->stream, lane.stream);
    std::swap(ctx->msm_ws, lane.ws);
    if (rc != ZKB_OK) lane_rc = rc;
  };
  on_lane(0, [&] { return msm_dev_impl<Fq2>(ctx, pk->b2_ext, 0, zb + pk->off_a, na, partial ? nullptr : o.oB, partial ? o.pB2 : nullptr); });
  on_lane(1, [&] { return msm_dev_impl<Fq>(ctx, pk->a_ext, 0, za + pk->off_a, na, partial ? nullptr : o.oA, prescaled ? nullptr : o.pA); });
  on_lane(2, [&] { return msm_dev_impl<Fq>(ctx, pk->b1_ext, 0, (prescaled ? zrb : zb) + pk->off_a, na, nullptr, o.pB1); });
  on_lane(3, [&] { return msm_dev_impl<Fq>(ctx, pk->l_ext, 0, zl + pk->off_l, pk->l_ext->n, nullptr, o.pL); });
  if (prescaled) on_lane(4, [&] { return msm_dev_impl<Fq>(ctx, pk->a_ext, 0, zsa + pk->off_a, na, nullptr, o.pSA); });

  // main stream: h = witness_map_from_matrices, then H = MSM(h_query, h) (msm_bigint truncates to the shorter of the two)
  auto main_part = [&]() -> int {
    WitnessBufs w{z, ctx->pzm.as<Fr>(), ctx->pwa.as<Fr>(), ctx->pwb.as<Fr>(), ctx->pwc.as<Fr>()};
    ZKB_TRY(witness_map_dev(ctx, m->a, m->b, m->c, m->nc, m->ni, m->nw, m->log_domain, w, ctx->ph.as<Fr>()));
    size_t hn = pk->nh < n ? pk->nh : n;
    size_t hcnt = hn > pk->off_h ? hn - pk->off_h : 0;
    if (hcnt > pk->h->n) hcnt = pk->h->n;
    ZKB_TRY((msm_dev_impl<Fq>(ctx, pk->h, 0, ctx->ph.as<Fr>() + pk->off_h, hcnt, nullptr, o.pH)));
    return ZKB_OK;
  };
  int mrc = lane_rc == ZKB_OK ? main_part() : ZKB_OK;
  // join every lane before anything else: their work reads buffers owned by this context
  bool joined = true;
  for (int i = 0; i < 5; i++)
    if (lane_used[i] && cudaStreamWaitEvent(st, ctx->aux[i].done, 0) != cudaSuccess) joined = false;
  if (lane_rc != ZKB_OK || mrc != ZKB_OK || !joined) {
    cudaGetLastError();
    for (auto& lane : ctx->aux) cudaStreamSynchronize(lane.stream);
    if (lane_rc == ZKB_ERR_CUDA || !joined) ctx->err = std::string(who) + ": CUDA error while queueing the MSM lanes";
    return lane_rc != ZKB_OK ? lane_rc : (mrc != ZKB_OK ? mrc : ZKB_ERR_CUDA);
  }
  // C = s A + r B1 + L + H
  if (prescaled) ZKB_TRY(prove_assemble_sum(ctx, o.pSA, o.pB1, o.pL, o.pH, o.oC));
  else if (!partial) ZKB_TRY(prove_assemble_c(ctx, o.pA, o.pB1, o.pL, o.pH, rs, rs + 1, o.oC));
  return ZKB_OK;
}


void prove_graphs_clear(zkb_ctx* ctx) {
  for (auto& g : ctx->graphs)
    if (g.exec) cudaGraphExecDestroy(g.exec);
  ctx->graphs.clear();
}

// Replays the captured device part when one is cached for (key, matrices) and no buffer has moved since; otherwise runs it
// directly (first call: the run that sizes every buffer, creates the lanes and the NTT tables) or captures it (second call).
// A capture that fails -- it is only ever an optimisation -- falls back to direct launches and is not retried for that key.
int prove_device_graphed(zkb_ctx* ctx, const zkb_pk* pk, const zkb_r1cs* m, bool partial, const char* who) {
  if (!ctx->graphs_on || ctx->prof.on) return prove_device_part(ctx, pk, m, partial, who);
  const unsigned long long epoch = g_alloc_epoch.load(std::memory_order_relaxed);
  zkb_ctx::ProveGraph* slot = nullptr;
  for (auto& g : ctx->graphs)
    if (g.pk == pk && g.m == m && g.partial == partial) slot = &g;
  if (slot && slot->epoch != epoch) {  // some buffer, key or table moved: whatever was captured is stale
    if (slot->exec) cudaGraphExecDestroy(slot->exec);
    slot->exec = nullptr;
    slot->epoch = epoch;
    return prove_device_part(ctx, pk, m, partial, who);  // warm again under the new layout
  }
  if (!slot) {
    if (ctx->graphs.size() >= 16) prove_graphs_clear(ctx);
    ctx->graphs.push_back({pk, m, partial, 0, nullptr, 0, 0});
    int rc = prove_device_part(ctx, pk, m, partial, who);
    ctx->graphs.back().epoch = g_alloc_epoch.load(std::memory_order_relaxed);  // after the allocations of this warm-up
    return rc;
  }
  if (slot->exec) {
    CUDA_TRY(ctx, cudaGraphLaunch(slot->exec, ctx->stream));
    ctx->launches += slot->kernels;
    ctx->graph_replays++;
    return ZKB_OK;
  }
  if (slot->failures) return prove_device_part(ctx, pk, m, partial, who);
  // capture
  const unsigned long long launches0 = ctx->launches;
  cudaGraph_t graph = nullptr;
  if (cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
    cudaGetLastError();
    slot->failures++;
    return prove_device_part(ctx, pk, m, partial, who);
  }
  int rc = prove_device_part(ctx, pk, m, partial, who);
  cudaError_t ce = cudaStreamEndCapture(ctx->stream, &graph);
  const unsigned long long kernels = ctx->launches - launches0;
  ctx->launches = launches0;  // nothing ran yet
  cudaGraphExec_t exec = nullptr;
  const unsigned long long epoch_now = g_alloc_epoch.load(std::memory_order_relaxed);
  const bool moved = rc == ZKB_OK && ce == cudaSuccess && epoch_now != epoch;  // another context allocated meanwhile: not a failure
  if (rc == ZKB_OK && ce == cudaSuccess && graph && !moved) ce = cudaGraphInstantiate(&exec, graph, 0);
  if (graph) cudaGraphDestroy(graph);
  if (moved || rc != ZKB_OK || ce != cudaSuccess || !exec) {
    cudaGetLastError();
    if (moved) slot->epoch = epoch_now; else slot->failures++;
    // leave the lanes in a defined state, then do the work directly
    for (auto& lane : ctx->aux)
      if (lane.stream) cudaStreamSynchronize(lane.stream);
    cudaGetLastError();
    return prove_device_part(ctx, pk, m, partial, who);
  }
  slot->exec = exec;
  slot->kernels = kernels;
  ctx->graph_captures++;
  CUDA_TRY(ctx, cudaGraphLaunch(exec, ctx->stream));
  ctx->launches += kernels;
  ctx->graph_replays++;
  return ZKB_OK;
}

// Queues the whole prover for this key (or key shard) on the context's streams.  partial: every MSM leaves its projective
// partial sum in ProveOut (A, B1, L, H, B2); otherwise A, B and C are finished to canonical affine bytes in oA, oB, oC.
int prove_enqueue(zkb_ctx* ctx, const zkb_pk* pk, const zkb_r1cs* m, const uint8_t* z_host, const uint8_t r[32],
                  const uint8_t s[32], bool partial, const char* who) {
  if (pk->device != ctx->device || m->device != ctx->device) ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "%s: key/matrices on another device", who);
  const size_t nv = m->ni + m->nw, nw = m->nw, ni = m->ni;
  const size_t n = size_t(1) << m->log_domain;
  if (pk->nv != nv) ZKB_FAIL(ctx, ZKB_ERR_SHAPE, "%s: key has %zu variables, circuit has %zu", who, pk->nv, nv);
  if (pk->nw != nw) ZKB_FAIL(ctx, ZKB_ERR_SHAPE, "%s: l_query has %zu points, circuit has %zu witness variables", who, pk->nw, nw);
  ZKB_ON_DEVICE(ctx);
  cudaStream_t st = ctx->stream;
  ZKB_TRY(reserve_prove_bufs(ctx, n, nv, nw));
  Fr* rs = ctx->prs.as<Fr>();
  Fr* z = ctx->pz.as<Fr>();
  CUDA_TRY(ctx, cudaMemcpyAsync(z, z_host, nv * 32, cudaMemcpyHostToDevice, st));
  CUDA_TRY(ctx, cudaMemcpyAsync(rs, r, 32, cudaMemcpyHostToDevice, st));
  CUDA_TRY(ctx, cudaMemcpyAsync(rs + 1, s, 32, cudaMemcpyHostToDevice, st));
  return prove_device_graphed(ctx, pk, m, partial, who);
}

}  // namespace

// =============================================================================================== batched prove
namespace {

struct BatchOut {
  XYZZ<Fq>*pA, *pB1, *pL, *pH;
  XYZZ<Fq2>* pB2;
};

BatchOut batch_out(zkb_ctx* ctx, size_t K) {
  BatchOut o;
  o.pA = ctx->bpart.as<XYZZ<Fq>>();
  o.pB1 = o.pA + K;
  o.pL = o.pA + 2 * K;
  o.pH = o.pA + 3 * K;
  o.pB2 = reinterpret_cast<XYZZ<Fq2>*>(o.pA + 4 * K);
  return o;
}

int ensure_lanes(zkb_ctx* ctx) {
  if (ctx->ev_inputs) return ZKB_OK;
  for (auto& lane : ctx->aux) {
    CUDA_TRY(ctx, cudaStreamCreateWithFlags(&lane.stream, cudaStreamNonBlocking));
    CUDA_TRY(ctx, cudaEventCreateWithFlags(&lane.done, cudaEventDisableTiming));
  }
  CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_inputs, cudaEventDisableTiming));
  return ZKB_OK;
}

}  // namespace

// K proofs of ONE circuit with ONE key, as one set of fat launches instead of K chains of small ones (ark-groth16
// create_proof_with_reduction_and_matrices x K; the data-parallel analogue of the forge coordinator's chunk-per-worker
// dispatch, forge/crates/prover-coordinator/src/dispatcher.rs:290-330): batched mat-vecs and NTTs (3 K polynomials per
// launch), five batched MSMs (K scalar vectors against one window table, per-proof bucket arrays: msm_run_batch) on five
// lanes, and one finishing kernel per group.  Asynchronous: _begin queues everything on ctx's streams including the copy of
// the K x 256 B results into pinned memory; _end waits and hands them out.  z_host and rs_host must stay valid until _end.
extern "C" int zkb_prove_batch_begin(zkb_ctx* ctx, const zkb_pk* pk_c, const zkb_r1cs* m, const uint8_t* z_host, const uint8_t* rs_host,
                                     size_t K) {
  return zkb_prove_batch_begin_ex(ctx, pk_c, m, z_host, rs_host, K, 0u);
}

// flags & ZKB_BATCH_Z_MONTGOMERY: the assignments are Montgomery limbs (x * 2^256 mod r, little-endian), the form a host
// synthesiser computes in -- the device needs both forms anyway (Montgomery for the A mat-vec, canonical for the MSM digits), so
// the conversion costs one kernel either way, and the host keeps the ~0.15 ms per L2-circuit proof of 6 000 from_mont products.
extern "C" int zkb_prove_batch_begin_ex(zkb_ctx* ctx, const zkb_pk* pk_c, const zkb_r1cs* m, const uint8_t* z_host, const uint8_t* rs_host,
                                        size_t K, unsigned flags) {
  if (!ctx) return ZKB_ERR_INVALID_ARG;
  if (flags & ~unsigned(ZKB_BATCH_Z_MONTGOMERY)) ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "zkb_prove_batch: unknown flag bits %#x", flags);
  const bool z_mont = (flags & ZKB_BATCH_Z_MONTGOMERY) != 0;
  if (!pk_c || !m || !z_host || !rs_host || K < 1 || K > 4096) ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "zkb_prove_batch: bad argument (1 <= K <= 4096)");
  zkb_pk* pk = const_cast<zkb_pk*>(pk_c);   // the delta tables are built on first use
  if (pk->world != 1) ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "zkb_prove_batch: sharded key");
  if (pk->device != ctx->device || m->device != ctx->device) ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "zkb_prove_batch: key/matrices on another device");
  const size_t nv = m->ni + m->nw, nw = m->nw, ni = m->ni;
  const size_t n = size_t(1) << m->log_domain;
  if (pk->nv != nv) ZKB_FAIL(ctx, ZKB_ERR_SHAPE, "zkb_prove_batch: key has %zu variables, circuit has %zu", pk->nv, nv);
  if (pk->nw != nw) ZKB_FAIL(ctx, ZKB_ERR_SHAPE, "zkb_prove_batch: l_query has %zu points, circuit has %zu witness variables", pk->nw, nw);
  ZKB_ON_DEVICE(ctx);
  cudaStream_t st = ctx->stream;
  ZKB_TRY(ensure_lanes(ctx));
  if (!pk->fb_delta1) {
    // delta_g1 / delta_g2 are the last base of a_ext / b2_ext.  Built once per key; the build runs on this context's stream and
    // is complete before any later work of this stream.  (Contexts sharing a key: call one batched prove before going wide.)
    void *t1 = nullptr, *t2 = nullptr;
    ZKB_TRY(fixed_table_for_base<Fq>(ctx, pk->a_ext, nv + 1, &t1));
    int rc = fixed_table_for_base<Fq2>(ctx, pk->b2_ext, nv + 1, &t2);
    if (rc != ZKB_OK) {
      cudaFree(t1);
      return rc;
    }
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    pk->fb_delta2 = t2;
    pk->fb_delta1 = t1;
  }
  if (pk->comb_state == 0) {
    // Small key, large HBM: tables of every digit multiple for all five query vectors (msm.cuh comb_build_kernel) turn each
    // MSM into a sum of gathered points: no buckets, no sort, no bucket reduction.  MEASURED AND NOT ADOPTED as the default
    // (profiles/r02_comb_tables_rejected.txt): even with the entries compacted and dealt out evenly the gathers from a
    // 75 + 35 GB table run the additions at 5.4 G/s (G1) and less than half the bucket kernel's rate (G2, 128-byte points);
    // the sub-batch takes 49.0 ms of kernels against 49.3 ms for the bucket method -- no gain for 110 GB of HBM.
    // Opt in with ZKB_COMB=1 (widest window <= 12 whose tables fit 60 % of the free memory; ZKB_COMB_MAX_GB caps it).
    int chosen = 0;
    const char* on = getenv("ZKB_COMB");
    if (on && on[0] == '1') {
      size_t free_b = 0, total_b = 0;
      CUDA_TRY(ctx, cudaMemGetInfo(&free_b, &total_b));
      double budget = double(free_b) * 0.6;
      if (const char* cap = getenv("ZKB_COMB_MAX_GB")) {
        double g = atof(cap) * 1e9;
        if (g >= 0 && g < budget) budget = g;
      }
      for (int c = 12; c >= 8 && !chosen; c--) {
        double tot = double(bases_comb_bytes<Fq>(pk->a_ext, c)) + double(bases_comb_bytes<Fq>(pk->b1_ext, c)) +
                     double(bases_comb_bytes<Fq>(pk->l_ext, c)) + double(bases_comb_bytes<Fq>(pk->h, c)) +
                     double(bases_comb_bytes<Fq2>(pk->b2_ext, c));
        if (tot <= budget) chosen = c;
      }
    }
    if (chosen) {
      int rc = bases_build_comb<Fq>(ctx, pk->a_ext, chosen);
      if (rc == ZKB_OK) rc = bases_build_comb<Fq>(ctx, pk->b1_ext, chosen);
      if (rc == ZKB_OK) rc = bases_build_comb<Fq>(ctx, pk->l_ext, chosen);
      if (rc == ZKB_OK) rc = bases_build_comb<Fq>(ctx, pk->h, chosen);
      if (rc == ZKB_OK) rc = bases_build_comb<Fq2>(ctx, pk->b2_ext, chosen);
      if (rc != ZKB_OK) {   // a failed build (out of memory) falls back to the bucket method and gives the memory back
        auto drop1 = [](zkb_g1_bases* b) {
          if (b && b->comb) cudaFree(b->comb);
          if (b) b->comb = nullptr, b->comb_c = 0;
        };
        drop1(pk->a_ext), drop1(pk->b1_ext), drop1(pk->l_ext), drop1(pk->h);
        if (pk->b2_ext && pk->b2_ext->comb) cudaFree(pk->b2_ext->comb);
        if (pk->b2_ext) pk->b2_ext->comb = nullptr, pk->b2_ext->comb_c = 0;
        cudaGetLastError();
      }
      pk->comb_state = rc == ZKB_OK ? 1 : -1;
    } else {
      pk->comb_state = -1;
    }
  }
  CUDA_TRY(ctx, ctx->bz.reserve(K * nv * 32));
  CUDA_TRY(ctx, ctx->bzm.reserve(K * nv * 32));
  CUDA_TRY(ctx, ctx->bw3.reserve(3 * K * n * 32));
  CUDA_TRY(ctx, ctx->bh.reserve(K * n * 32));
  CUDA_TRY(ctx, ctx->brs.reserve(K * 64));
  CUDA_TRY(ctx, ctx->bpart.reserve(K * (4 * sizeof(XYZZ<Fq>) + sizeof(XYZZ<Fq2>))));
  CUDA_TRY(ctx, ctx->bout.reserve(K * 256));
  if (ctx->bpinned_cap < K * 256) {
    if (ctx->bpinned) cudaFreeHost(ctx->bpinned);
    ctx->bpinned = nullptr;
    ctx->bpinned_cap = 0;
    CUDA_TRY(ctx, cudaHostAlloc(reinterpret_cast<void**>(&ctx->bpinned), K * 256, cudaHostAllocDefault));
    ctx->bpinned_cap = K * 256;
  }
  Fr* z = ctx->bz.as<Fr>();
  Fr* rs = ctx->brs.as<Fr>();
  ZKB_TRY(clear_flag(ctx));
  CUDA_TRY(ctx, cudaMemcpyAsync(z_mont ? ctx->bzm.p : ctx->bz.p, z_host, K * nv * 32, cudaMemcpyHostToDevice, st));
  CUDA_TRY(ctx, cudaMemcpyAsync(rs, rs_host, K * 64, cudaMemcpyHostToDevice, st));
  if (z_mont) ZKB_TRY(fr_from_mont(ctx, ctx->bzm.as<Fr>(), z, K * nv));   // the MSM lanes read canonical scalars
  CUDA_TRY(ctx, cudaEventRecord(ctx->ev_inputs, st));
  BatchOut o = batch_out(ctx, K);
  // A, B1, B2, L depend on z only: four lanes next to the witness maps + H on the main stream
  int lane_rc = ZKB_OK;
  bool lane_used[5] = {false, false, false, false, false};
  auto on_lane = [&](int idx, auto&& fn) {
    auto& lane = ctx->aux[idx];
    if (lane_rc != ZKB_OK) return;
    if (cudaStreamWaitEvent(lane.stream, ctx->ev_inputs, 0) != cudaSuccess) {
      lane_rc = ZKB_ERR_CUDA;
      return;
    }
    lane_used[idx] = true;
    std::swap(ctx->stream, lane.stream);
    std::swap(ctx->msm_ws, lane.ws);
    int rc = fn();
    if (rc == ZKB_OK && cudaEventRecord(lane.done, ctx->stream) != cudaSuccess) rc = ZKB_ERR_CUDA;
    std::swap(ctx->stream, lane.stream);
    std::swap(ctx->msm_ws, lane.ws);
    if (rc != ZKB_OK) lane_rc = rc;
  };
  const int Ki = int(K);
  const size_t hn = pk->nh < n ? pk->nh : n;
  const bool comb = pk->comb_state == 1;
  auto msm1 = [&](const zkb_g1_bases* b, const Fr* sc, size_t cnt, size_t stride, XYZZ<Fq>* out) {
    return comb ? msm_comb_dev_impl<Fq>(ctx, b, 0, sc, cnt, stride, Ki, out) : msm_batch_dev_impl<Fq>(ctx, b, 0, sc, cnt, stride, Ki, nullptr, out);
  };
  on_lane(0, [&] {
    return comb ? msm_comb_dev_impl<Fq2>(ctx, pk->b2_ext, 0, z + 1, nv - 1, nv, Ki, o.pB2)
                : msm_batch_dev_impl<Fq2>(ctx, pk->b2_ext, 0, z + 1, nv - 1, nv, Ki, nullptr, o.pB2);
  });
  on_lane(1, [&] { return msm1(pk->a_ext, z + 1, nv - 1, nv, o.pA); });
  on_lane(2, [&] { return msm1(pk->b1_ext, z + 1, nv - 1, nv, o.pB1); });
  on_lane(3, [&] { return msm1(pk->l_ext, z + ni, nw, nv, o.pL); });
  int mrc = ZKB_OK;
  if (lane_rc == ZKB_OK) {
    mrc = witness_map_batch_dev(ctx, m->a, m->b, m->c, m->nc, m->ni, m->nw, m->log_domain, Ki, z, ctx->bzm.as<Fr>(), z_mont, ctx->bw3.as<Fr>(),
                                ctx->bh.as<Fr>());
    if (mrc == ZKB_OK) mrc = msm1(pk->h, ctx->bh.as<Fr>(), hn, n, o.pH);
  }
  bool joined = true;
  for (int i = 0; i < 5; i++)
    if (lane_used[i] && cudaStreamWaitEvent(st, ctx->aux[i].done, 0) != cudaSuccess) joined = false;
  if (lane_rc != ZKB_OK || mrc != ZKB_OK || !joined) {
    cudaGetLastError();
    for (auto& lane : ctx->aux) cudaStreamSynchronize(lane.stream);
    cudaStreamSynchronize(st);
    if (lane_rc == ZKB_ERR_CUDA || !joined) ctx->err = "zkb_prove_batch: CUDA error while queueing the MSM lanes";
    return lane_rc != ZKB_OK ? lane_rc : (mrc != ZKB_OK ? mrc : ZKB_ERR_CUDA);
  }
  const Affine<Fq>* a_tail = pk->a_ext->p + (nv - 1);     // a_query[0], alpha_g1, delta_g1
  const Affine<Fq>* b1_tail = pk->b1_ext->p + (nv - 1);   // b_g1_query[0], beta_g1, delta_g1
  const Affine<Fq2>* b2_tail = pk->b2_ext->p + (nv - 1);  // b_g2_query[0], beta_g2, delta_g2
  ZKB_TRY(prove_batch_finish_g1(ctx, Ki, o.pA, o.pB1, o.pL, o.pH, rs, a_tail, b1_tail, pk->fb_delta1, ctx->bout.p));
  ZKB_TRY(prove_batch_finish_g2(ctx, Ki, o.pB2, rs, b2_tail, pk->fb_delta2, ctx->bout.p));
  CUDA_TRY(ctx, cudaMemcpyAsync(ctx->bpinned, ctx->bout.p, K * 256, cudaMemcpyDeviceToHost, st));
  return ZKB_OK;
}

// out: K x 256 B, proof i = A (64) | B (128) | C (64) canonical affine, A not negated -- three zkb_prove outputs back to back.
extern "C" int zkb_prove_batch_end(zkb_ctx* ctx, size_t K, uint8_t* out) {
  if (!ctx) return ZKB_ERR_INVALID_ARG;
  if (!out || K * 256 > ctx->bpinned_cap) ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "zkb_prove_batch_end: no batch of %zu proofs in flight", K);
  ZKB_ON_DEVICE(ctx);
  int rc = check_flag(ctx, "witness assignment");   // waits for the stream
  if (rc != ZKB_OK) return rc;
  memcpy(out, ctx->bpinned, K * 256);
  return ZKB_OK;
}

extern "C" int zkb_prove_batch(zkb_ctx* ctx, const zkb_pk* pk, const zkb_r1cs* m, const uint8_t* z_host, const uint8_t* rs_host, size_t K,
                               uint8_t* out) {
  int rc = zkb_prove_batch_begin(ctx, pk, m, z_host, rs_host, K);
  if (rc != ZKB_OK) return rc;
  return zkb_prove_batch_end(ctx, K, out);
}

extern "C" int zkb_prove(zkb_ctx* ctx, const zkb_pk* pk, const zkb_r1cs* m, const uint8_t* z_host, const uint8_t r[32],
                         const uint8_t s[32], uint8_t out_a[64], uint8_t out_b[128], uint8_t out_c[64]) {
  if (!ctx) return ZKB_ERR_INVALID_ARG;
  if (!pk || !m || !z_host || !r || !s || !out_a || !out_b || !out_c) ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "zkb_prove: null argument");
  if (pk->world != 1) ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "zkb_prove: this key is shard %d of %d (use zkb_prove_partial)", pk->shard, pk->world);
  ZKB_ON_DEVICE(ctx);   // also covers the copies and the wait after prove_enqueue (a host thread's current device is 0 by default)
  ZKB_TRY(prove_enqueue(ctx, pk, m, z_host, r, s, false, "zkb_prove"));
  ProveOut o = prove_out(ctx);
  cudaStream_t st = ctx->stream;
  // oA | oB | oC are contiguous (64 + 128 + 64 B): one copy into pinned memory, so that the wait below is the only wait
  ZKB_TRY(ensure_pinned(ctx));
  CUDA_TRY(ctx, cudaMemcpyAsync(ctx->pinned, o.oA, 256, cudaMemcpyDeviceToHost, st));
  ZKB_TRY(check_flag(ctx, "witness assignment"));
  memcpy(out_a, ctx->pinned, 64);
  memcpy(out_b, ctx->pinned + 64, 128);
  memcpy(out_c, ctx->pinned + 192, 64);
  return ZKB_OK;
}

// One rank's share of a proof whose key is sharded over `world` GPUs: every rank runs the witness map (no
// exchange) and the five MSMs over ITS range of the query vectors; out_partial_dev receives ZKB_PROVE_PARTIAL_BYTES.
extern "C" int zkb_prove_partial(zkb_ctx* ctx, const zkb_pk* pk, const zkb_r1cs* m, const uint8_t* z_host, const uint8_t r[32],
                                 const uint8_t s[32], void* out_partial_dev) {
  if (!ctx) return ZKB_ERR_INVALID_ARG;
  if (!pk || !m || !z_host || !r || !s || !out_partial_dev) ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "zkb_prove_partial: null argument");
  ZKB_ON_DEVICE(ctx);
  ZKB_TRY(prove_enqueue(ctx, pk, m, z_host, r, s, true, "zkb_prove_partial"));
  CUDA_TRY(ctx, cudaMemcpyAsync(out_partial_dev, ctx->ppts.p, ZKB_PROVE_PARTIAL_BYTES, cudaMemcpyDeviceToDevice, ctx->stream));
  return check_flag(ctx, "witness assignment");
}

// partials_dev: `world` records of ZKB_PROVE_PARTIAL_BYTES (all-gathered) -> the proof.
extern "C" int zkb_prove_combine(zkb_ctx* ctx, const void* partials_dev, int world, const uint8_t r[32], const uint8_t s[32],
                                 uint8_t out_a[64], uint8_t out_b[128], uint8_t out_c[64]) {
  if (!ctx) return ZKB_ERR_INVALID_ARG;
  if (!partials_dev || world < 1 || !r || !s || !out_a || !out_b || !out_c) ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "zkb_prove_combine: bad argument");
  ZKB_ON_DEVICE(ctx);
  cudaStream_t st = ctx->stream;
  CUDA_TRY(ctx, ctx->prs.reserve(64));
  CUDA_TRY(ctx, ctx->ppts.reserve(4 * sizeof(XYZZ<Fq>) + sizeof(XYZZ<Fq2>) + 64 + 128 + 64 + sizeof(XYZZ<Fq>)));
  Fr* rs = ctx->prs.as<Fr>();
  CUDA_TRY(ctx, cudaMemcpyAsync(rs, r, 32, cudaMemcpyHostToDevice, st));
  CUDA_TRY(ctx, cudaMemcpyAsync(rs + 1, s, 32, cudaMemcpyHostToDevice, st));
  ProveOut o = prove_out(ctx);
  ZKB_TRY(prove_combine_g1(ctx, partials_dev, world, ZKB_PROVE_PARTIAL_BYTES, rs, rs + 1, o.oA, o.oC));
  ZKB_TRY(prove_combine_g2(ctx, partials_dev, world, ZKB_PROVE_PARTIAL_BYTES, o.oB));
  CUDA_TRY(ctx, cudaMemcpyAsync(out_a, o.oA, 64, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(ctx, cudaMemcpyAsync(out_b, o.oB, 128, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(ctx, cudaMemcpyAsync(out_c, o.oC, 64, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(ctx, cudaStreamSynchronize(st));
  return ZKB_OK;
}

// ONE proof over the GPUs of one process: context i holds shard i of the key (zkb_pk_load_shard(.., i, n_gpus))
// and its own copy of the matrices; every GPU runs zkb_prove_partial from its own host thread, the 768-byte partial records are
// gathered through host memory and ctxs[0] finishes the proof.  No torch, no NCCL: what a Rust GpuGroth16Prover with one
// context per device calls from `BatchProver::prove` (core/src/sequencer/settlement/prover.rs:160-169, 350-425).
extern "C" int zkb_prove_multi(zkb_ctx* const* ctxs, const zkb_pk* const* pk_shards, const zkb_r1cs* const* ms, int n_gpus,
                               const uint8_t* z_host, const uint8_t r[32], const uint8_t s[32], uint8_t out_a[64], uint8_t out_b[128],
                               uint8_t out_c[64]) {
  if (!ctxs || n_gpus < 1 || !ctxs[0]) return ZKB_ERR_INVALID_ARG;
  zkb_ctx* c0 = ctxs[0];
  if (!pk_shards || !ms || !z_host || !r || !s || !out_a || !out_b || !out_c || n_gpus > 64)
    ZKB_FAIL(c0, ZKB_ERR_INVALID_ARG, "zkb_prove_multi: bad argument");
  for (int i = 0; i < n_gpus; i++) {
    if (!ctxs[i] || !pk_shards[i] || !ms[i]) ZKB_FAIL(c0, ZKB_ERR_INVALID_ARG, "zkb_prove_multi: null handle for GPU %d", i);
    if (pk_shards[i]->world != n_gpus || pk_shards[i]->shard != i)
      ZKB_FAIL(c0, ZKB_ERR_SHAPE, "zkb_prove_multi: key handle %d is shard %d of %d, expected %d of %d", i, pk_shards[i]->shard,
               pk_shards[i]->world, i, n_gpus);
  }
  std::vector<uint8_t> parts(size_t(n_gpus) * ZKB_PROVE_PARTIAL_BYTES);
  std::vector<int> rc(size_t(n_gpus), ZKB_OK);
  auto work = [&](int i) {
    zkb_ctx* ctx = ctxs[i];
    rc[size_t(i)] = [&]() -> int {
      ZKB_ON_DEVICE(ctx);
      CUDA_TRY(ctx, ctx->tmp2.reserve(ZKB_PROVE_PARTIAL_BYTES));
      ZKB_TRY(zkb_prove_partial(ctx, pk_shards[i], ms[i], z_host, r, s, ctx->tmp2.p));
      CUDA_TRY(ctx, cudaMemcpy(parts.data() + size_t(i) * ZKB_PROVE_PARTIAL_BYTES, ctx->tmp2.p, ZKB_PROVE_PARTIAL_BYTES, cudaMemcpyDeviceToHost));
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
    for (int i = 1; i < n_gpus; i++) jn.th.emplace_back(work, i);
    work(0);
  } catch (...) {
    ZKB_FAIL(c0, ZKB_ERR_OOM, "zkb_prove_multi: could not start the per-GPU host threads");
  }
  for (int i = 0; i < n_gpus; i++)
    if (rc[size_t(i)] != ZKB_OK) {
      if (i) c0->err = ctxs[i]->err;
      return rc[size_t(i)];
    }
  ZKB_ON_DEVICE(c0);
  CUDA_TRY(c0, c0->tmp1.reserve(parts.size()));
  CUDA_TRY(c0, cudaMemcpy(c0->tmp1.p, parts.data(), parts.size(), cudaMemcpyHostToDevice));
  return zkb_prove_combine(c0, c0->tmp1.p, n_gpus, r, s, out_a, out_b, out_c);
}
