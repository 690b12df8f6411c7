// G1 (Fq) instantiation of the group kernels: the four G1 MSMs of ark-groth16's create_proof_with_assignment
// (SURVEY.md 8a row a6), point import/export, fixed-base generation, Fq parity hook and the final assembly of C.
#include "group_impl.cuh"

namespace zkb {
ZKB_INSTANTIATE_GROUP(Fq)

namespace {

__global__ void fq_field_op_kernel(int op, const Fq* a, const Fq* b, size_t n, Fq* out, int* bad) {
  size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fq x = a[i];
  if (!fp_is_canonical(x)) { atomicExch(bad, 1); return; }
  Fq y = Fq::zero();
  if (op <= 2) {
    y = b[i];
    if (!fp_is_canonical(y)) { atomicExch(bad, 1); return; }
  }
  Fq r;
  switch (op) {
    case 0: r = x + y; break;
    case 1: r = x - y; break;
    case 2: r = (x.to_mont() * y.to_mont()).from_mont(); break;
    case 3: r = x.to_mont().inverse().from_mont(); break;
    default: r = x.neg(); break;
  }
  out[i] = r;
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
    store_affine_canonical<Fq>(acc.to_affine_vartime(), out_c);
  }
}

// C = SA + RB1 + L + H: the scalars s and r went into the MSMs (SA = MSM(a_ext, s za), RB1 = MSM(b1_ext, r zb)), so the
// two 254-step double-and-add chains of prove_assemble_c_kernel (1.9 ms on one thread) are not needed.
__global__ void prove_assemble_sum_kernel(const XYZZ<Fq>* SA, const XYZZ<Fq>* RB1, const XYZZ<Fq>* L, const XYZZ<Fq>* H,
                                          uint32_t* out_c) {
  if (blockIdx.x || threadIdx.x) return;
  XYZZ<Fq> acc = *SA;
  acc.add(*RB1);
  acc.add(*L);
  acc.add(*H);
  store_affine_canonical<Fq>(acc.to_affine_vartime(), out_c);
}

// parts: `world` records of `stride` bytes, each starting with the four G1 partial sums A, B1, L, H.
__global__ void prove_combine_g1_kernel(const char* parts, int world, size_t stride, const uint32_t* r, const uint32_t* s,
                                        uint32_t* out_a, uint32_t* out_c) {
  __shared__ XYZZ<Fq> sum[4];
  int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  if (lane == 0 && warp < 4) {
    XYZZ<Fq> acc = XYZZ<Fq>::inf();
    for (int k = 0; k < world; k++) acc.add(load_xyzz(reinterpret_cast<const XYZZ<Fq>*>(parts + size_t(k) * stride) + warp));
    if (warp < 2) {  // s * A, r * B1
      uint32_t e[8];
      for (int j = 0; j < 8; j++) e[j] = warp == 0 ? s[j] : r[j];
      if (warp == 0) store_affine_canonical<Fq>(acc.to_affine_vartime(), out_a);
      acc = acc.mul_words(e);
    }
    sum[warp] = acc;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    XYZZ<Fq> acc = sum[0];
    acc.add(sum[1]);
    acc.add(sum[2]);
    acc.add(sum[3]);
    store_affine_canonical<Fq>(acc.to_affine_vartime(), out_c);
  }
}

// Batched prove, G1 part.  Two lanes per proof, 16 proofs per warp: the even lane finishes A and multiplies it by s, the odd lane
// finishes B1 and multiplies it by r -- the same instruction stream on different data, so the lanes of a warp stay together;
// the even lane then adds L + H - r s delta and converts C to affine.  The two 254-bit multiplications are 4-bit fixed-window
// chains (14 additions for the table, then 252 doublings + <= 64 additions): ~3 400 products each on one thread.  This is a
// latency-bound tail; what matters in a pipeline of sub-batches is that it holds few registers while another sub-batch's
// accumulation fills the machine: 16 warps for 256 proofs (the first version ran a block of three warps per proof, one lane
// of each active: measured 3.2 ms of every 45 ms sub-batch NOT hidden behind the other slots' kernels).
__global__ void __launch_bounds__(32)
prove_batch_finish_g1_kernel(int K, const XYZZ<Fq>* __restrict__ PA, const XYZZ<Fq>* __restrict__ PB1, const XYZZ<Fq>* __restrict__ PL,
                             const XYZZ<Fq>* __restrict__ PH, const uint32_t* __restrict__ rs, const Affine<Fq>* __restrict__ a_tail,
                             const Affine<Fq>* __restrict__ b1_tail, const Affine<Fq>* __restrict__ fb_delta, uint32_t* __restrict__ out) {
  const int t = blockIdx.x * 32 + threadIdx.x, lane = threadIdx.x;
  const int role = t & 1;
  const bool live = (t >> 1) < K;
  const int p = live ? (t >> 1) : K - 1;   // idle lanes of the last warp repeat the last proof and store nothing
  uint32_t rw[8], sw[8];
  for (int j = 0; j < 8; j++) {
    rw[j] = rs[size_t(p) * 16 + j];
    sw[j] = rs[size_t(p) * 16 + 8 + j];
  }
  const uint32_t* mulw = role == 0 ? sw : rw;     // the chain's multiplier: s for A, r for B1
  const uint32_t* delw = role == 0 ? rw : sw;     // A carries r delta, B1 carries s delta
  const Affine<Fq>* tail = role == 0 ? a_tail : b1_tail;
  XYZZ<Fq> base = load_xyzz((role == 0 ? PA : PB1) + p);
  base.madd(tail[0]);
  base.madd(tail[1]);
  base.add(fixed_table_mul<Fq>(fb_delta, delw));
  if (role == 0 && live) store_affine_canonical<Fq>(base.to_affine_vartime(), out + size_t(p) * 64);
  XYZZ<Fq> tab[15];   // d * base, d = 1..15 (local memory)
  tab[0] = base;
  tab[1] = base.dbl();
#pragma unroll 1
  for (int d = 2; d < 15; d++) {
    XYZZ<Fq> v = tab[d - 1];
    v.add(base);
    tab[d] = v;
  }
  XYZZ<Fq> acc = XYZZ<Fq>::inf();
#pragma unroll 1
  for (int w = 63; w >= 0; w--) {
    if (w != 63) {
#pragma unroll 1
      for (int k = 0; k < 4; k++) acc = acc.dbl();
    }
    uint32_t word = mulw[0];
#pragma unroll
    for (int j = 1; j < 8; j++)
      if ((w >> 3) == j) word = mulw[j];
    const uint32_t d = (word >> (4 * (w & 7))) & 15u;
    if (d) acc.add(tab[d - 1]);
  }
  XYZZ<Fq> other = shfl_xyzz(acc, lane ^ 1);   // the even lane receives r * B1
  if (role != 0) return;
  acc.add(other);
  Fr rm, sm;
  for (int j = 0; j < 8; j++) {
    rm.v[j] = rw[j];
    sm.v[j] = sw[j];
  }
  Fr nrs = (rm.to_mont() * sm.to_mont()).from_mont().neg();   // -(r s) mod r, canonical
  XYZZ<Fq> l = load_xyzz(PL + p);
  l.add(load_xyzz(PH + p));
  l.add(fixed_table_mul<Fq>(fb_delta, nrs.v));
  acc.add(l);
  if (live) store_affine_canonical<Fq>(acc.to_affine_vartime(), out + size_t(p) * 64 + 48);
}

}  // namespace

int prove_batch_finish_g1(zkb_ctx* ctx, int K, const void* PA, const void* PB1, const void* PL, const void* PH, const void* rs,
                          const void* a_tail, const void* b1_tail, const void* fb_delta1, void* out) {
  ProfScope ps(ctx, PH_ASSEMBLE);
  prove_batch_finish_g1_kernel<<<unsigned((2 * K + 31) / 32), 32, 0, ctx->stream>>>(
      K, static_cast<const XYZZ<Fq>*>(PA), static_cast<const XYZZ<Fq>*>(PB1), static_cast<const XYZZ<Fq>*>(PL),
      static_cast<const XYZZ<Fq>*>(PH), static_cast<const uint32_t*>(rs), static_cast<const Affine<Fq>*>(a_tail),
      static_cast<const Affine<Fq>*>(b1_tail), static_cast<const Affine<Fq>*>(fb_delta1), static_cast<uint32_t*>(out));
  ctx->launches++;
  CUDA_TRY(ctx, cudaGetLastError());
  return ZKB_OK;
}

int prove_combine_g1(zkb_ctx* ctx, const void* parts, int world, size_t stride, const void* r_dev, const void* s_dev,
                     void* out_a_dev, void* out_c_dev) {
  ProfScope ps(ctx, PH_ASSEMBLE);
  prove_combine_g1_kernel<<<1, 128, 0, ctx->stream>>>(static_cast<const char*>(parts), world, stride,
                                                      static_cast<const uint32_t*>(r_dev), static_cast<const uint32_t*>(s_dev),
                                                      static_cast<uint32_t*>(out_a_dev), static_cast<uint32_t*>(out_c_dev));
  ctx->launches++;
  CUDA_TRY(ctx, cudaGetLastError());
  return ZKB_OK;
}

int fq_field_op(zkb_ctx* ctx, int op, const uint8_t* a, const uint8_t* b, size_t n, uint8_t* out) {
  size_t bytes = n * 32;
  CUDA_TRY(ctx, ctx->tmp0.reserve(bytes));
  CUDA_TRY(ctx, ctx->tmp1.reserve(bytes));
  CUDA_TRY(ctx, ctx->tmp2.reserve(bytes));
  ZKB_TRY(clear_flag(ctx));
  CUDA_TRY(ctx, cudaMemcpyAsync(ctx->tmp0.p, a, bytes, cudaMemcpyHostToDevice, ctx->stream));
  if (op <= 2) CUDA_TRY(ctx, cudaMemcpyAsync(ctx->tmp1.p, b, bytes, cudaMemcpyHostToDevice, ctx->stream));
  fq_field_op_kernel<<<blocks_for(n, 128), 128, 0, ctx->stream>>>(op, ctx->tmp0.as<Fq>(), ctx->tmp1.as<Fq>(), n, ctx->tmp2.as<Fq>(), ctx->flag.as<int>());
  ctx->launches++;
  CUDA_TRY(ctx, cudaGetLastError());
  CUDA_TRY(ctx, cudaMemcpyAsync(out, ctx->tmp2.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  return check_flag(ctx, "zkb_field_op");
}

int prove_assemble_c(zkb_ctx* ctx, const void* pA, const void* pB1, const void* pL, const void* pH, const void* r_dev,
                     const void* s_dev, void* out_c_dev) {
  ProfScope ps(ctx, PH_ASSEMBLE);
  prove_assemble_c_kernel<<<1, 64, 0, ctx->stream>>>(static_cast<const XYZZ<Fq>*>(pA), static_cast<const XYZZ<Fq>*>(pB1),
                                                     static_cast<const XYZZ<Fq>*>(pL), static_cast<const XYZZ<Fq>*>(pH),
                                                     static_cast<const uint32_t*>(r_dev), static_cast<const uint32_t*>(s_dev),
                                                     static_cast<uint32_t*>(out_c_dev));
  ctx->launches++;
  CUDA_TRY(ctx, cudaGetLastError());
  return ZKB_OK;
}

int prove_assemble_sum(zkb_ctx* ctx, const void* pSA, const void* pRB1, const void* pL, const void* pH, void* out_c_dev) {
  ProfScope ps(ctx, PH_ASSEMBLE);
  prove_assemble_sum_kernel<<<1, 32, 0, ctx->stream>>>(static_cast<const XYZZ<Fq>*>(pSA), static_cast<const XYZZ<Fq>*>(pRB1),
                                                       static_cast<const XYZZ<Fq>*>(pL), static_cast<const XYZZ<Fq>*>(pH),
                                                       static_cast<uint32_t*>(out_c_dev));
  ctx->launches++;
  CUDA_TRY(ctx, cudaGetLastError());
  return ZKB_OK;
}

}  // namespace zkb
