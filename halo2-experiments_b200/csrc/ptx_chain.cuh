// Carry-chain integer primitives for 254-bit Montgomery arithmetic on the sm_100a
// integer pipe.  On the device every primitive is exactly one PTX instruction
// (add.cc / addc / sub.cc / subc / mad.{lo,hi}.cc / madc.{lo,hi}); ptxas fuses a
// mad.lo.cc + madc.hi pair on the same operands into one IMAD.WIDE.U32(.X).
// When this header is compiled by a plain host compiler (g++) the same primitives are
// emulated with an explicit carry flag so the limb algorithms in field.cuh can
// be unit-tested on a CPU-only box (tests/test_limb_emulation.py).
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define ZK_D __device__ __forceinline__
#else
#define ZK_D inline
#endif

namespace b200zk {
namespace ptx {

#if defined(__CUDACC__)

ZK_D uint32_t add_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ZK_D uint32_t addc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ZK_D uint32_t addc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ZK_D uint32_t sub_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ZK_D uint32_t subc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ZK_D uint32_t subc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ZK_D uint32_t mul_lo(uint32_t a, uint32_t b) { uint32_t r; asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ZK_D uint32_t mul_hi(uint32_t a, uint32_t b) { uint32_t r; asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ZK_D uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
ZK_D uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("mad.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
ZK_D uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
ZK_D uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
ZK_D uint32_t madc_lo(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
ZK_D uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
ZK_D uint32_t mad_lo(uint32_t a, uint32_t b, uint32_t c) { return a * b + c; }

#else  // host emulation of the PTX condition-code register

static thread_local uint32_t g_cf = 0;
inline uint32_t emu_add(uint32_t a, uint32_t b, uint32_t cin, bool set) {
    uint64_t s = (uint64_t)a + b + cin; if (set) g_cf = (uint32_t)(s >> 32); return (uint32_t)s;
}
inline uint32_t emu_sub(uint32_t a, uint32_t b, uint32_t bin, bool set) {
    uint64_t d = (uint64_t)a - b - bin; if (set) g_cf = (uint32_t)(d >> 63); return (uint32_t)d;
}
inline uint32_t add_cc(uint32_t a, uint32_t b) { return emu_add(a, b, 0, true); }
inline uint32_t addc_cc(uint32_t a, uint32_t b) { return emu_add(a, b, g_cf, true); }
inline uint32_t addc(uint32_t a, uint32_t b) { return emu_add(a, b, g_cf, false); }
inline uint32_t sub_cc(uint32_t a, uint32_t b) { return emu_sub(a, b, 0, true); }
inline uint32_t subc_cc(uint32_t a, uint32_t b) { return emu_sub(a, b, g_cf, true); }
inline uint32_t subc(uint32_t a, uint32_t b) { return emu_sub(a, b, g_cf, false); }
inline uint32_t mul_lo(uint32_t a, uint32_t b) { return (uint32_t)((uint64_t)a * b); }
inline uint32_t mul_hi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
inline uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return emu_add(mul_lo(a, b), c, 0, true); }
inline uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) { return emu_add(mul_hi(a, b), c, 0, true); }
inline uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return emu_add(mul_lo(a, b), c, g_cf, true); }
inline uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { return emu_add(mul_hi(a, b), c, g_cf, true); }
inline uint32_t madc_lo(uint32_t a, uint32_t b, uint32_t c) { return emu_add(mul_lo(a, b), c, g_cf, false); }
inline uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { return emu_add(mul_hi(a, b), c, g_cf, false); }
inline uint32_t mad_lo(uint32_t a, uint32_t b, uint32_t c) { return a * b + c; }

#endif

}  // namespace ptx
}  // namespace b200zk
