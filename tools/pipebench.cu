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
  unsigned lo[CH], hi[CH], ad[CH], cl[CH], chh[CH], bank[CH];
  double fd[CH];
  unsigned x0[CH / 4 + 1], x1[CH / 4 + 1], x2[CH / 4 + 1], x3[CH / 4 + 1], x4[CH / 4 + 1], x5[CH / 4 + 1], x6[CH / 4 + 1], x7[CH / 4 + 1];
#pragma unroll
  for (int c = 0; c < CH; c++) { w[c] = seed[c] + threadIdx.x; lo[c] = seed[c + 8]; hi[c] = seed[c + 16]; ad[c] = seed[c + 24]; cl[c] = seed[c] ^ 5; chh[c] = seed[c] ^ 9; bank[c] = 0; fd[c] = 1.0 + 1e-9 * (seed[c] & 1023);
    x0[c / 4] = seed[c]; x1[c / 4] = seed[c + 1]; x2[c / 4] = seed[c + 2]; x3[c / 4] = seed[c + 3]; x4[c / 4] = seed[c + 4]; x5[c / 4] = seed[c + 5]; x6[c / 4] = seed[c + 6]; x7[c / 4] = seed[c + 7]; }
  const double fk = 1.0 + 1e-12 * (seed[3] & 7);
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
        // round 2: 64 = DFMA; 128 = 32x32->64 accumulate with carry-OUT only, the carry banked by an addc (the chain-head form);
        // 256 = one chain of four 32x32->64 accumulates linked by carries (what a row of the Montgomery product is), per 4 chains
        if (MODE & 64) asm volatile("fma.rn.f64 %0, %0, %1, %0;" : "+d"(fd[c]) : "d"(fk));
        if (MODE & 128) asm volatile("mad.lo.cc.u32 %0, %3, %4, %0;\n\tmadc.hi.cc.u32 %1, %3, %4, %1;\n\taddc.u32 %2, %2, 0;" : "+r"(cl[c]), "+r"(chh[c]), "+r"(bank[c]) : "r"(a + c), "r"(cl[c] | 1u));
        if ((MODE & 256) && (c & 3) == 0) asm volatile(
            "mad.lo.cc.u32 %0, %8, %9, %0;\n\tmadc.hi.cc.u32 %1, %8, %9, %1;\n\t"
            "madc.lo.cc.u32 %2, %8, %10, %2;\n\tmadc.hi.cc.u32 %3, %8, %10, %3;\n\t"
            "madc.lo.cc.u32 %4, %8, %11, %4;\n\tmadc.hi.cc.u32 %5, %8, %11, %5;\n\t"
            "madc.lo.cc.u32 %6, %8, %9, %6;\n\tmadc.hi.u32 %7, %8, %10, %7;"
            : "+r"(x0[c / 4]), "+r"(x1[c / 4]), "+r"(x2[c / 4]), "+r"(x3[c / 4]), "+r"(x4[c / 4]), "+r"(x5[c / 4]), "+r"(x6[c / 4]), "+r"(x7[c / 4])
            : "r"(x0[c / 4] | 1u), "r"(a + c), "r"(b), "r"(a ^ b));
        if (MODE & 32) asm volatile("mad.lo.cc.u32 %0, %1, %2, %0;\n\tmadc.hi.u32 %1, %0, %2, %1;" : "+r"(cl[c]), "+r"(chh[c]) : "r"(a + c));
      }
    }
  }
  unsigned long long s = 0;
#pragma unroll
  for (int c = 0; c < CH; c++) s ^= w[c] ^ lo[c] ^ hi[c] ^ ad[c] ^ cl[c] ^ chh[c] ^ bank[c] ^ (unsigned long long)__double_as_longlong(fd[c]);
#pragma unroll
  for (int c = 0; c < CH; c += 4) s ^= x0[c / 4] + 3 * x1[c / 4] + 5 * x2[c / 4] + 7 * x3[c / 4] + 11 * x4[c / 4] + 13 * x5[c / 4] + 17 * x6[c / 4] + 19 * x7[c / 4];
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
  // round 2 additions: the FP64 pipe and the carry forms of the wide multiply
  run<64>("DFMA", 1, seed, sink, sms, mhz);
  run<64 | 1>("DFMA + IMAD.WIDE", 2, seed, sink, sms, mhz);
  run<64 | 16>("DFMA + XOR+IADD", 3, seed, sink, sms, mhz);
  run<128>("wide MAD, carry OUT only + addc bank (3 PTX)", 3, seed, sink, sms, mhz);
  run<256>("chain of 4 wide MADs linked by carries (8 PTX) /4", 2, seed, sink, sms, mhz);   // 8 PTX per 4 chains = 2 per chain
  run<256 | 64>("carry chain of 4 + DFMA", 3, seed, sink, sms, mhz);
  return 0;
}
