#include "group_impl.cuh"
namespace zkb { const GroupOps* group_ops_bls_g1() { return GroupImpl<Fp<BlsFq>, BlsFr>::ops(); } }
