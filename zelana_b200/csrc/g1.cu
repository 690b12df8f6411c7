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

}  // namespace

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
