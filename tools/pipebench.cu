// Which integer multiply forms share a pipe on B200?  (experiment, not part of the library)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/_bin/pipebench tools/pipebench.cu
// Every variant runs N independent dependency chains per thread; rates are warp-instructions per cycle per scheduler (SMSP).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA %s line %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

constexpr int CH = 8;     // chains per instruction class
constexpr int INNER = 32;

// MODE bits: 1 = IMAD.WIDE (64-bit acc, no carry), 2 = IMAD.WIDE with carry chain (.cc/.x pairs emulate via mad.lo.cc+madc.hi),
//            4 = plain IMAD lo (32-bit), 8 = IMAD.HI (mul.hi + add), 16 = IADD3, 32 = IMAD.WIDE.X chain (madc.wide not in PTX: use mad.lo.cc/madc.hi.cc)
template <int MODE>
__global__ void __launch_bounds__(256) k(const unsigned* seed, int iters, unsigned long long* sink) {
  unsigned a = seed[threadIdx.x & 31] | 1u, b = seed[32 + (threadIdx.x & 31)] | 1u;
  unsigned long long w[CH];
  unsigned lo[CH], hi[CH], ad[CH], cl[CH], chh[CH];
#pragma unroll
  for (int c = 0; c < CH; c++) { w[c] = seed[c] + threadIdx.x; lo[c] = seed[c + 8]; hi[c] = seed[c + 16]; ad[c] = seed[c + 24]; cl[c] = seed[c] ^ 5; chh[c] = seed[c] ^ 9; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < INNER; r++) {
#pragma unroll
      for (int c = 0; c < CH; c++) {
        if (MODE & 1) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[c]) : "r"(a + c), "r"(b));
        // nonlinear recurrences: ptxas folds x += a*b over a loop into a closed form
        if (MODE & 4) asm volatile("mad.lo.u32 %0, %0, %0, %1;" : "+r"(lo[c]) : "r"(a + c));
        if (MODE & 8) asm volatile("mad.hi.u32 %0, %0, %0, %1;" : "+r"(hi[c]) : "r"(a + c));
        if (MODE & 16) asm volatile("xor.b32 %0, %0, %1;\n\tadd.u32 %0, %0, %2;" : "+r"(ad[c]) : "r"(a + c), "r"(b));
        if (MODE & 32) asm volatile("mad.lo.cc.u32 %0, %1, %2, %0;\n\tmadc.hi.u32 %1, %0, %2, %1;" : "+r"(cl[c]), "+r"(chh[c]) : "r"(a + c));
      }
    }
  }
  unsigned long long s = 0;
#pragma unroll
  for (int c = 0; c < CH; c++) s ^= w[c] ^ lo[c] ^ hi[c] ^ ad[c] ^ cl[c] ^ chh[c];
  if ((unsigned)(s ^ (s >> 32)) == 0x12345678u) sink[0] = s;
}

template <int MODE>
void run(const char* name, int ninstr_classes, const unsigned* seed, unsigned long long* sink, int sms, double mhz) {
  const int iters = 2000, blocks = sms * 8, threads = 256;
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  float best = 1e30f;
  for (int rep = 0; rep < 3; rep++) {
    CK(cudaEventRecord(a));
    k<MODE><<<blocks, threads>>>(seed, iters, sink);
    CK(cudaEventRecord(b));
    CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    if (ms < best) best = ms;
  }
  double warp_instr = double(blocks) * (threads / 32) * iters * INNER * CH * ninstr_classes;
  double cycles = best * 1e-3 * mhz * 1e6;
  printf("%-46s %8.3f ms   %.3f warp-instr / cycle / SMSP\n", name, best, warp_instr / (cycles * sms * 4));
}

int main(int argc, char** argv) {
  double mhz = argc > 1 ? atof(argv[1]) : 1965.0;
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  unsigned h[64]; for (int i = 0; i < 64; i++) h[i] = 0x9e3779b9u * (i + 1);
  unsigned* seed; unsigned long long* sink;
  CK(cudaMalloc(&seed, 256)); CK(cudaMalloc(&sink, 64)); CK(cudaMemcpy(seed, h, 256, cudaMemcpyHostToDevice));
  int sms = p.multiProcessorCount;
  run<1>("IMAD.WIDE", 1, seed, sink, sms, mhz);
  run<4>("IMAD (lo)", 1, seed, sink, sms, mhz);
  run<8>("IMAD.HI", 1, seed, sink, sms, mhz);
  run<16>("XOR+IADD (2 ALU instr)", 2, seed, sink, sms, mhz);
  run<32>("mad.lo.cc + madc.hi (2 instr)", 2, seed, sink, sms, mhz);
  run<1 | 4>("IMAD.WIDE + IMAD lo", 2, seed, sink, sms, mhz);
  run<1 | 8>("IMAD.WIDE + IMAD.HI", 2, seed, sink, sms, mhz);
  run<1 | 16>("IMAD.WIDE + XOR+IADD", 3, seed, sink, sms, mhz);
  run<4 | 8>("IMAD lo + IMAD.HI", 2, seed, sink, sms, mhz);
  run<4 | 16>("IMAD lo + XOR+IADD", 3, seed, sink, sms, mhz);
  run<32 | 16>("carry pair + XOR+IADD", 4, seed, sink, sms, mhz);
  run<32 | 4>("carry pair + IMAD lo", 3, seed, sink, sms, mhz);
  run<1 | 4 | 16>("IMAD.WIDE + IMAD lo + XOR+IADD", 4, seed, sink, sms, mhz);
  return 0;
}
