// G2 (Fq2) instantiation of the group kernels: b_g2_query MSM of ark-groth16's create_proof_with_assignment
// (SURVEY.md 8a row a7), point import/export, fixed-base generation.
#include "group_impl.cuh"

namespace zkb {
ZKB_INSTANTIATE_GROUP(Fq2)

namespace {
// parts: `world` records of `stride` bytes; the G2 partial sum B2 sits after the four G1 partial sums (4 x 128 B).
__global__ void prove_combine_g2_kernel(const char* parts, int world, size_t stride, uint32_t* out_b) {
  if (blockIdx.x || threadIdx.x) return;
  XYZZ<Fq2> acc = XYZZ<Fq2>::inf();
  for (int k = 0; k < world; k++) acc.add(load_xyzz(reinterpret_cast<const XYZZ<Fq2>*>(parts + size_t(k) * stride + 4 * sizeof(XYZZ<Fq>))));
  store_affine_canonical<Fq2>(acc.to_affine_vartime(), out_b);
}
// Batched prove, G2 part: B = PB2 + b_g2_query[0] + beta_g2 + s delta_g2, one thread per proof.
__global__ void prove_batch_finish_g2_kernel(int K, const XYZZ<Fq2>* __restrict__ PB2, const uint32_t* __restrict__ rs,
                                             const Affine<Fq2>* __restrict__ b2_tail, const Affine<Fq2>* __restrict__ fb_delta,
                                             uint32_t* __restrict__ out) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= K) return;
  uint32_t sw[8];
  for (int j = 0; j < 8; j++) sw[j] = rs[size_t(p) * 16 + 8 + j];
  XYZZ<Fq2> b = load_xyzz(PB2 + p);
  b.madd(b2_tail[0]);
  b.madd(b2_tail[1]);
  b.add(fixed_table_mul<Fq2>(fb_delta, sw));
  store_affine_canonical<Fq2>(b.to_affine_vartime(), out + size_t(p) * 64 + 16);
}
}  // namespace

int prove_batch_finish_g2(zkb_ctx* ctx, int K, const void* PB2, const void* rs, const void* b2_tail, const void* fb_delta2, void* out) {
  prove_batch_finish_g2_kernel<<<unsigned((K + 31) / 32), 32, 0, ctx->stream>>>(
      K, static_cast<const XYZZ<Fq2>*>(PB2), static_cast<const uint32_t*>(rs), static_cast<const Affine<Fq2>*>(b2_tail),
      static_cast<const Affine<Fq2>*>(fb_delta2), static_cast<uint32_t*>(out));
  ctx->launches++;
  CUDA_TRY(ctx, cudaGetLastError());
  return ZKB_OK;
}

int prove_combine_g2(zkb_ctx* ctx, const void* parts, int world, size_t stride, void* out_b_dev) {
  prove_combine_g2_kernel<<<1, 32, 0, ctx->stream>>>(static_cast<const char*>(parts), world, stride, static_cast<uint32_t*>(out_b_dev));
  ctx->launches++;
  CUDA_TRY(ctx, cudaGetLastError());
  return ZKB_OK;
}
}  // namespace zkb
