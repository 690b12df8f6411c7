// INT32 multiply-pipe peak microbenchmark (the MSM roofline denominator P_mul32 is the MEASURED
// rate of dependency-free 32x32->64 multiply-accumulates on this GPU, in both instruction forms).
//   variant 0: mad.wide.u32 (one IMAD.WIDE per 32x32->64 product)
//   variant 1: mad.lo.cc.u32 + madc.hi.u32 pairs (two IMADs per product, the carry-chain form)
// Sixteen independent accumulator chains per thread hide the 4-cycle pipe latency; operands come from
// global memory so that the compiler cannot fold anything.
#include "internal.h"

namespace zkb {
namespace {

constexpr int CHAINS = 16;
constexpr int INNER = 64;

__global__ void __launch_bounds__(256) imad_wide_kernel(const uint32_t* __restrict__ seed, int iters, unsigned long long* __restrict__ sink) {
  uint32_t a = seed[threadIdx.x & 31] | 1u, b = seed[32 + (threadIdx.x & 31)] | 1u;
  unsigned long long acc[CHAINS];
#pragma unroll
  for (int c = 0; c < CHAINS; c++) acc[c] = seed[c] + threadIdx.x;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int k = 0; k < INNER; k++) {
#pragma unroll
      for (int c = 0; c < CHAINS; c++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[c]) : "r"(a + c), "r"(b));
    }
  }
  unsigned long long s = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; c++) s ^= acc[c];
  if (s == 0x1234567ull) sink[0] = s;  // never true in practice; keeps the chains live
}

__global__ void __launch_bounds__(256) imad_pair_kernel(const uint32_t* __restrict__ seed, int iters, unsigned long long* __restrict__ sink) {
  uint32_t a = seed[threadIdx.x & 31] | 1u, b = seed[32 + (threadIdx.x & 31)] | 1u;
  uint32_t lo[CHAINS], hi[CHAINS];
#pragma unroll
  for (int c = 0; c < CHAINS; c++) {
    lo[c] = seed[c] + threadIdx.x;
    hi[c] = seed[c + 16];
  }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int k = 0; k < INNER; k++) {
#pragma unroll
      for (int c = 0; c < CHAINS; c++)
        asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;" : "+r"(lo[c]), "+r"(hi[c]) : "r"(a + c), "r"(b));
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; c++) s ^= lo[c] ^ hi[c];
  if (s == 0x12345u) sink[0] = s;
}

}  // namespace
}  // namespace zkb

using namespace zkb;

// -> mul32_per_s: 32x32->64 multiply-accumulates per second over the whole GPU; sm_clock_mhz is not sampled here
// (bench.py samples nvidia-smi concurrently).
extern "C" int zkb_bench_int32_peak(zkb_ctx* ctx, int variant, int iters, double* mul32_per_s, double* elapsed_ms) {
  if (!ctx) return ZKB_ERR_INVALID_ARG;
  if (variant < 0 || variant > 1 || iters <= 0 || !mul32_per_s) ZKB_FAIL(ctx, ZKB_ERR_INVALID_ARG, "zkb_bench_int32_peak: bad argument");
  ZKB_ON_DEVICE(ctx);
  CUDA_TRY(ctx, ctx->tmp0.reserve(4096));
  std::vector<uint32_t> seed(64);
  for (int i = 0; i < 64; i++) seed[i] = 0x9e3779b9u * (i + 1);
  CUDA_TRY(ctx, cudaMemcpyAsync(ctx->tmp0.p, seed.data(), 256, cudaMemcpyHostToDevice, ctx->stream));
  const int threads = 256, blocks = ctx->sm_count * 8;
  unsigned long long* sink = reinterpret_cast<unsigned long long*>(static_cast<char*>(ctx->tmp0.p) + 1024);
  cudaEvent_t a, b;
  CUDA_TRY(ctx, cudaEventCreate(&a));
  CUDA_TRY(ctx, cudaEventCreate(&b));
  for (int rep = 0; rep < 2; rep++) {  // first repetition warms up
    cudaEventRecord(a, ctx->stream);
    if (variant == 0) imad_wide_kernel<<<blocks, threads, 0, ctx->stream>>>(ctx->tmp0.as<uint32_t>(), iters, sink);
    else imad_pair_kernel<<<blocks, threads, 0, ctx->stream>>>(ctx->tmp0.as<uint32_t>(), iters, sink);
    cudaEventRecord(b, ctx->stream);
    ctx->launches++;
  }
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  float ms = 0.f;
  if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, a, b);
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  CUDA_TRY(ctx, e);
  double muls = double(blocks) * threads * double(iters) * INNER * CHAINS;
  *mul32_per_s = muls / (double(ms) * 1e-3);
  if (elapsed_ms) *elapsed_ms = ms;
  return ZKB_OK;
}
