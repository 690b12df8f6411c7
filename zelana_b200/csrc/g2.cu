// G2 (Fq2) instantiation of the group kernels: b_g2_query MSM of ark-groth16's create_proof_with_assignment
// (SURVEY.md 8a row a7), point import/export, fixed-base generation.
#include "group_impl.cuh"

namespace zkb {
ZKB_INSTANTIATE_GROUP(Fq2)
}  // namespace zkb
