// Small device helpers shared by the kernels: 128-bit vector load/store of POD structs.
#pragma once
#include <cstdint>

namespace zkb {

// ------------------------------------------------------------------------------------------
// vector load/store of POD structs whose size is a multiple of 16 bytes
// ------------------------------------------------------------------------------------------
template <class T>
__device__ __forceinline__ T ld_vec(const T* p) {
  static_assert(sizeof(T) % 16 == 0, "16-byte multiple expected");
  T r;
  const uint4* s = reinterpret_cast<const uint4*>(p);
  uint4* d = reinterpret_cast<uint4*>(&r);
#pragma unroll
  for (int i = 0; i < (int)(sizeof(T) / 16); i++) d[i] = __ldg(s + i);
  return r;
}
template <class T>
__device__ __forceinline__ T ld_vec_rw(const T* p) {   // data written earlier by this grid's predecessors
  T r;
  const uint4* s = reinterpret_cast<const uint4*>(p);
  uint4* d = reinterpret_cast<uint4*>(&r);
#pragma unroll
  for (int i = 0; i < (int)(sizeof(T) / 16); i++) d[i] = s[i];
  return r;
}
template <class T>
__device__ __forceinline__ void st_vec(T* p, const T& v) {
  uint4* d = reinterpret_cast<uint4*>(p);
  const uint4* s = reinterpret_cast<const uint4*>(&v);
#pragma unroll
  for (int i = 0; i < (int)(sizeof(T) / 16); i++) d[i] = s[i];
}

}  // namespace zkb
