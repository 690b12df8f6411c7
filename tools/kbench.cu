// Kernel-level experiments for the MSM inner loop (not part of the product library).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o tools/_bin/kbench tools/kbench.cu
// Prints, for each variant, the fraction of the measured IMAD.WIDE peak that register-resident Montgomery
// multiplications / XYZZ mixed additions reach, as a function of warps per scheduler, ILP and register budget.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../zelana_b200/csrc/ec.cuh"
#include "fp29_proto.cuh"

using namespace zkb;

#define CK(x)                                                                        \
  do {                                                                               \
    cudaError_t e_ = (x);                                                            \
    if (e_ != cudaSuccess) {                                                         \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(1);                                                                       \
    }                                                                                \
  } while (0)

__device__ __forceinline__ Fq load_fq(const uint32_t* p) {
  Fq r;
  for (int i = 0; i < 8; i++) r.v[i] = p[i];
  return r;
}

// ---- E1: raw multiplication chains, ILP independent chains per thread
template <int ILP, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) mul_chain_kernel(const uint32_t* seed, int iters, uint32_t* sink) {
  Fq x[ILP], y = load_fq(seed + 8 * (threadIdx.x & 3));
#pragma unroll
  for (int k = 0; k < ILP; k++) x[k] = load_fq(seed + 8 * ((threadIdx.x + k) & 7));
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int k = 0; k < ILP; k++) x[k] = x[k] * y;
  }
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < ILP; k++) s ^= x[k].v[0] ^ x[k].v[7];
  if (s == 0xdeadbeefu) sink[0] = s;
}

// ---- E1b: 9 x 29-bit carry-free Montgomery product chains
template <int ILP, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) mul29_chain_kernel(const uint32_t* seed, int iters, uint32_t* sink) {
  Fe29 x[ILP], y;
#pragma unroll
  for (int i = 0; i < 9; i++) y.v[i] = seed[(threadIdx.x + i) & 31] >> 3;
#pragma unroll
  for (int k = 0; k < ILP; k++)
#pragma unroll
    for (int i = 0; i < 9; i++) x[k].v[i] = seed[(threadIdx.x + k + 2 * i) & 31] >> 3;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int k = 0; k < ILP; k++) x[k] = mul29<P29Fq>(x[k], y);
  }
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < ILP; k++) s ^= x[k].v[0] ^ x[k].v[8];
  if (s == 0xdeadbeefu) sink[0] = s;
}

// ---- E2: XYZZ mixed addition, accumulator in registers, NACC independent accumulators per thread
template <int NACC, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) madd_reg_kernel(const uint32_t* seed, int iters, uint32_t* sink) {
  XYZZ<Fq> acc[NACC];
  Affine<Fq> p0, p1;
  p0.x = load_fq(seed + 8 * (threadIdx.x & 3));
  p0.y = load_fq(seed + 8 * ((threadIdx.x + 1) & 7));
  p1.x = load_fq(seed + 8 * ((threadIdx.x + 2) & 7));
  p1.y = load_fq(seed + 8 * ((threadIdx.x + 3) & 7));
#pragma unroll
  for (int k = 0; k < NACC; k++) {
    acc[k].x = load_fq(seed + 8 * ((threadIdx.x + k) & 7));
    acc[k].y = load_fq(seed + 8 * ((threadIdx.x + k + 1) & 7));
    acc[k].zz = load_fq(seed + 8 * ((threadIdx.x + k + 2) & 7));
    acc[k].zzz = load_fq(seed + 8 * ((threadIdx.x + k + 3) & 7));
  }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int k = 0; k < NACC; k++) acc[k].madd((it & 1) ? p1 : p0);
  }
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < NACC; k++) s ^= acc[k].x.v[0] ^ acc[k].y.v[3] ^ acc[k].zz.v[1] ^ acc[k].zzz.v[2];
  if (s == 0xdeadbeefu) sink[0] = s;
}

// ---- E2b: G2 (Fq2) mixed addition at different register budgets
template <int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) madd_g2_kernel(const uint32_t* seed, int iters, uint32_t* sink) {
  XYZZ<Fq2> acc;
  Affine<Fq2> p0, p1;
  auto ld2 = [&](int o) { return Fq2{load_fq(seed + 8 * ((threadIdx.x + o) & 7)), load_fq(seed + 8 * ((threadIdx.x + o + 1) & 7))}; };
  p0.x = ld2(0); p0.y = ld2(1); p1.x = ld2(2); p1.y = ld2(3);
  acc.x = ld2(4); acc.y = ld2(5); acc.zz = ld2(6); acc.zzz = ld2(7);
  for (int it = 0; it < iters; it++) acc.madd((it & 1) ? p1 : p0);
  uint32_t s = acc.x.c0.v[0] ^ acc.y.c1.v[3] ^ acc.zz.c0.v[1] ^ acc.zzz.c1.v[2];
  if (s == 0xdeadbeefu) sink[0] = s;
}

// ---- E3: accumulator resident in shared memory (one 128-byte slot per thread, word-interleaved across threads
// so that lane l reads bank l: conflict-free), registers hold only the operands of the product in flight.
template <int THREADS>
struct SmemAcc {
  uint32_t* base;  // [32 words][THREADS]
  __device__ __forceinline__ Fq ld(int field) const {
    Fq r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = base[(field * 8 + i) * THREADS];
    return r;
  }
  __device__ __forceinline__ void st(int field, const Fq& v) const {
#pragma unroll
    for (int i = 0; i < 8; i++) base[(field * 8 + i) * THREADS] = v.v[i];
  }
};

template <int THREADS>
__device__ __forceinline__ void madd_smem(const SmemAcc<THREADS>& a, const Fq& qx, const Fq& qy) {
  // fields: 0 x, 1 y, 2 zz, 3 zzz.  (Special cases are not benchmarked here: random field elements never hit them.)
  Fq P = qx * a.ld(2) - a.ld(0);
  Fq R = qy * a.ld(3) - a.ld(1);
  Fq PP = P.sqr();
  Fq PPP = P * PP;
  a.st(2, a.ld(2) * PP);
  a.st(3, a.ld(3) * PPP);
  Fq Q = a.ld(0) * PP;
  Fq X3 = R.sqr() - PPP - Q.dbl();
  a.st(0, X3);
  a.st(1, R * (Q - X3) - a.ld(1) * PPP);
}

template <int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) madd_smem_kernel(const uint32_t* seed, int iters, uint32_t* sink) {
  extern __shared__ uint32_t sm[];
  SmemAcc<THREADS> a{sm + threadIdx.x};
  for (int f = 0; f < 4; f++) a.st(f, load_fq(seed + 8 * ((threadIdx.x + f) & 7)));
  Fq p0x = load_fq(seed + 8 * (threadIdx.x & 3)), p0y = load_fq(seed + 8 * ((threadIdx.x + 1) & 7));
  for (int it = 0; it < iters; it++) {
    madd_smem<THREADS>(a, p0x, p0y);
  }
  Fq x = a.ld(0);
  if (x.v[0] == 0xdeadbeefu) sink[0] = x.v[1];
}

static double g_peak = 18.13e12;
static int g_sms = 148;

template <class K>
void time_kernel(const char* name, K kernel, int threads, int blocks_per_sm, size_t smem, int iters, double mul_per_iter_thread,
                 const uint32_t* seed, uint32_t* sink) {
  cudaFuncAttributes fa;
  CK(cudaFuncGetAttributes(&fa, kernel));
  if (smem > 48 * 1024) CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, smem));
  int bps = blocks_per_sm < occ ? blocks_per_sm : occ;
  if (bps < 1) {
    printf("%-44s regs=%3d  cannot launch\n", name, fa.numRegs);
    return;
  }
  int blocks = g_sms * bps;
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  float best = 1e30f;
  for (int rep = 0; rep < 3; rep++) {
    CK(cudaEventRecord(a));
    kernel<<<blocks, threads, smem>>>(seed, iters, sink);
    CK(cudaEventRecord(b));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    if (ms < best) best = ms;
  }
  double muls = double(blocks) * threads * iters * mul_per_iter_thread;
  double rate = muls * 136.0 / (best * 1e-3);
  printf("%-44s regs=%3d spill=%4zuB warps/SMSP=%4.1f  %8.3f ms  %6.2f Tmul32/s  %5.1f%% of IMAD peak\n", name, fa.numRegs,
         (size_t)fa.localSizeBytes, bps * threads / 128.0, best, rate / 1e12, 100.0 * rate / g_peak);
  CK(cudaEventDestroy(a));
  CK(cudaEventDestroy(b));
}

int main(int argc, char** argv) {
  if (argc > 1) g_peak = atof(argv[1]) * 1e12;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  g_sms = prop.multiProcessorCount;
  printf("device %s, %d SMs, peak used for fractions: %.2f Tmul32/s\n", prop.name, g_sms, g_peak / 1e12);
  std::vector<uint32_t> h(64);
  for (int i = 0; i < 64; i++) h[i] = 0x9e3779b9u * (i + 1) >> ((i % 8 == 7) ? 3 : 0);
  uint32_t *seed, *sink;
  CK(cudaMalloc(&seed, 256));
  CK(cudaMalloc(&sink, 256));
  CK(cudaMemcpy(seed, h.data(), 256, cudaMemcpyHostToDevice));
  const int IT = 4000;
#define MULK(ILP, MINB) time_kernel("mul chain ILP=" #ILP " blocks/SM=" #MINB, mul_chain_kernel<ILP, 128, MINB>, 128, MINB, 0, IT, ILP, seed, sink)
  MULK(1, 4); MULK(1, 8); MULK(1, 12); MULK(1, 16);
  MULK(2, 4); MULK(2, 8); MULK(2, 12); MULK(2, 16);
  MULK(4, 4); MULK(4, 8);
#define MUL29K(ILP, MINB) time_kernel("mul29 chain ILP=" #ILP " blocks/SM=" #MINB, mul29_chain_kernel<ILP, 128, MINB>, 128, MINB, 0, IT, ILP, seed, sink)
  MUL29K(1, 4); MUL29K(1, 6); MUL29K(1, 8); MUL29K(1, 12); MUL29K(2, 4); MUL29K(2, 6); MUL29K(2, 8);
  const int IM = 600;
#define MADDK(NACC, MINB) time_kernel("madd regs NACC=" #NACC " blocks/SM=" #MINB, madd_reg_kernel<NACC, 128, MINB>, 128, MINB, 0, IM, 10 * NACC, seed, sink)
  MADDK(1, 3); MADDK(1, 4); MADDK(1, 5); MADDK(1, 6); MADDK(1, 8);
  MADDK(2, 2); MADDK(2, 3); MADDK(2, 4);
#define MADDG2(TH, MINB) time_kernel("madd G2 threads=" #TH " blocks/SM=" #MINB, madd_g2_kernel<TH, MINB>, TH, MINB, 0, 200, 28, seed, sink)
  MADDG2(64, 2); MADDG2(64, 3); MADDG2(64, 4); MADDG2(64, 5); MADDG2(64, 6); MADDG2(64, 8); MADDG2(128, 2); MADDG2(128, 3); MADDG2(128, 4);
#define MADDS(MINB) time_kernel("madd smem-acc blocks/SM=" #MINB, madd_smem_kernel<128, MINB>, 128, MINB, 128 * 128, IM, 10, seed, sink)
  MADDS(4); MADDS(5); MADDS(6); MADDS(8); MADDS(10); MADDS(12);
  return 0;
}
