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
}  // namespace

int prove_combine_g2(zkb_ctx* ctx, const void* parts, int world, size_t stride, void* out_b_dev) {
  prove_combine_g2_kernel<<<1, 32, 0, ctx->stream>>>(static_cast<const char*>(parts), world, stride, static_cast<uint32_t*>(out_b_dev));
  ctx->launches++;
  CUDA_TRY(ctx, cudaGetLastError());
  return ZKB_OK;
}
}  // namespace zkb
