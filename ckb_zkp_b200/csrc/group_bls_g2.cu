#include "group_impl.cuh"
namespace zkb { const GroupOps* group_ops_bls_g2() { return GroupImpl<Fp2<BlsFq>, BlsFr>::ops(); } }
