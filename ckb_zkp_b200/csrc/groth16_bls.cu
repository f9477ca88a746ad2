#include "groth16_impl.cuh"
namespace zkb { const Groth16Ops* groth16_ops_bls() { return Groth16Impl<BlsFr, BlsFq, ZKB_BLS12_381>::ops(); } }
