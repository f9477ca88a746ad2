#include "group_impl.cuh"
namespace zkb { const GroupOps* group_ops_bn_g1() { return GroupImpl<Fp<BnFq>, BnFr>::ops(); } }
