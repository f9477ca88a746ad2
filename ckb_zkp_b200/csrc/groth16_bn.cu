#include "groth16_impl.cuh"
namespace zkb { const Groth16Ops* groth16_ops_bn() { return Groth16Impl<BnFr, BnFq, ZKB_BN254>::ops(); } }
