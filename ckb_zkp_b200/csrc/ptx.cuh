// Carry-chain integer primitives: one PTX instruction each on the device, and a
// bit-faithful emulation (explicit carry flag) when the same headers are compiled
// by a host compiler.  The host emulation exists so that the exact instruction
// sequences of field.cuh can be unit-tested without a GPU (tests/test_host_emu.py);
// it is never a product path -- the C-ABI library contains device code only.
#pragma once
#include <cstdint>

#ifdef __CUDACC__
#define ZKB_HD __host__ __device__ __forceinline__
#define ZKB_D __device__ __forceinline__
#define ZKB_NOINLINE __host__ __device__ __noinline__
#else
#define ZKB_HD inline
#define ZKB_D inline
#define ZKB_NOINLINE inline
#endif

namespace zkb {
namespace ptx {

#ifdef __CUDA_ARCH__

ZKB_D uint32_t add_cc(uint32_t a, uint32_t b) {
  uint32_t r; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
ZKB_D uint32_t addc_cc(uint32_t a, uint32_t b) {
  uint32_t r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
ZKB_D uint32_t addc(uint32_t a, uint32_t b) {
  uint32_t r; asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
ZKB_D uint32_t sub_cc(uint32_t a, uint32_t b) {
  uint32_t r; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
ZKB_D uint32_t subc_cc(uint32_t a, uint32_t b) {
  uint32_t r; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
ZKB_D uint32_t subc(uint32_t a, uint32_t b) {
  uint32_t r; asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
ZKB_D uint32_t mul_lo(uint32_t a, uint32_t b) {
  uint32_t r; asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
ZKB_D uint32_t mul_hi(uint32_t a, uint32_t b) {
  uint32_t r; asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
ZKB_D uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r; asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
}
ZKB_D uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r; asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
}
ZKB_D uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r; asm volatile("mad.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
}
ZKB_D uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r; asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
}
ZKB_D uint32_t madc_lo(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r; asm volatile("madc.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
}
ZKB_D uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r; asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
}

// (hi:lo) += a * b as ONE asm statement per pair so ptxas fuses it into IMAD.WIDE.U32[.X]
ZKB_D void mad_wide_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {      // no carry in, carry out
  asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(a), "r"(b));
}
ZKB_D void madc_wide_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {     // carry in, carry out
  asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(a), "r"(b));
}
ZKB_D void madc_wide(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {        // carry in, no carry out
  asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(a), "r"(b));
}
// (d_hi:d_lo) = a * b + (c_hi:c_lo) + carry, carry out  (destination pair != addend pair)
ZKB_D void madc_wide_cc_3(uint32_t& d_lo, uint32_t& d_hi, uint32_t a, uint32_t b, uint32_t c_lo, uint32_t c_hi) {
  asm volatile("madc.lo.cc.u32 %0, %2, %3, %4; madc.hi.cc.u32 %1, %2, %3, %5;"
               : "=r"(d_lo), "=r"(d_hi) : "r"(a), "r"(b), "r"(c_lo), "r"(c_hi));
}
ZKB_D void mul_wide(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {         // (hi:lo) = a * b
  asm volatile("mul.lo.u32 %0, %2, %3; mul.hi.u32 %1, %2, %3;" : "=r"(lo), "=r"(hi) : "r"(a), "r"(b));
}

#else  // ---- host emulation ---------------------------------------------------

inline uint32_t& cc() { static thread_local uint32_t flag = 0; return flag; }

inline uint32_t add_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a + b; cc() = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t addc_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a + b + cc(); cc() = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t addc(uint32_t a, uint32_t b) { return a + b + cc(); }
inline uint32_t sub_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a - b; cc() = (uint32_t)((t >> 32) & 1); return (uint32_t)t; }
inline uint32_t subc_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a - b - cc(); cc() = (uint32_t)((t >> 32) & 1); return (uint32_t)t; }
inline uint32_t subc(uint32_t a, uint32_t b) { return a - b - cc(); }
inline uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }
inline uint32_t mul_hi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
inline uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return add_cc(a * b, c); }
inline uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return addc_cc(a * b, c); }
inline uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) { return add_cc(mul_hi(a, b), c); }
inline uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { return addc_cc(mul_hi(a, b), c); }
inline uint32_t madc_lo(uint32_t a, uint32_t b, uint32_t c) { return addc(a * b, c); }
inline uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { return addc(mul_hi(a, b), c); }
inline void mad_wide_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) { lo = mad_lo_cc(a, b, lo); hi = madc_hi_cc(a, b, hi); }
inline void madc_wide_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) { lo = madc_lo_cc(a, b, lo); hi = madc_hi_cc(a, b, hi); }
inline void madc_wide(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) { lo = madc_lo_cc(a, b, lo); hi = madc_hi(a, b, hi); }
inline void mul_wide(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) { lo = a * b; hi = mul_hi(a, b); }
inline void madc_wide_cc_3(uint32_t& d_lo, uint32_t& d_hi, uint32_t a, uint32_t b, uint32_t c_lo, uint32_t c_hi) {
  uint32_t l = madc_lo_cc(a, b, c_lo); uint32_t h = madc_hi_cc(a, b, c_hi); d_lo = l; d_hi = h;
}

#endif

}  // namespace ptx
}  // namespace zkb
