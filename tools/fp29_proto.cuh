#pragma once
#include <cstdint>
#ifndef __CUDACC__
#define __host__
#define __device__
#define __forceinline__ inline
#endif
// prototype: 9 x 29-bit limbs, Montgomery radix 2^261, carry-free column accumulation
struct P29Fq {
  static __host__ __device__ constexpr uint32_t p(int i) {
    constexpr uint32_t t[9] = {0x187cfd47u, 0x10460b6u, 0x1c72a34fu, 0x2d522d0u, 0x1585d978u, 0x2db40c0u, 0xa6e141u, 0xe5c2634u, 0x30644eu};
    return t[i];
  }
  static constexpr uint32_t NINV = 0x4866389u;  // -p^-1 mod 2^29
};
struct Fe29 { uint32_t v[9]; };

template <class C>
__host__ __device__ __forceinline__ Fe29 mul29(const Fe29& a, const Fe29& b) {
  constexpr uint32_t MASK = (1u << 29) - 1;
  uint64_t t[18];
#pragma unroll
  for (int k = 0; k < 18; k++) t[k] = 0;
#pragma unroll
  for (int i = 0; i < 9; i++)
#pragma unroll
    for (int j = 0; j < 9; j++) t[i + j] += (uint64_t)a.v[i] * b.v[j];
#pragma unroll
  for (int i = 0; i < 9; i++) {
    uint32_t m = ((uint32_t)t[i] * C::NINV) & MASK;
#pragma unroll
    for (int j = 0; j < 9; j++) t[i + j] += (uint64_t)m * C::p(j);
    t[i + 1] += t[i] >> 29;
  }
  Fe29 r;
#pragma unroll
  for (int k = 9; k < 17; k++) {
    t[k + 1] += t[k] >> 29;
    r.v[k - 9] = (uint32_t)t[k] & MASK;
  }
  r.v[8] = (uint32_t)t[17];
  return r;
}
