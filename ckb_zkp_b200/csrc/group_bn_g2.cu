#include "group_impl.cuh"
namespace zkb { const GroupOps* group_ops_bn_g2() { return GroupImpl<Fp2<BnFq>, BnFr>::ops(); } }
