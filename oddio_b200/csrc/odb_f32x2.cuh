// Packed FP32x2 arithmetic (sm_100: FADD2 / FMUL2 / FFMA2) on a 64-bit register pair: both ears of a frame ride in one
// instruction. Every operation rounds each half exactly like the scalar operation of the same name.
#pragma once
#include <stdint.h>

#include "odb_async.cuh"

namespace odbk {

__device__ __forceinline__ u64 pk2(float lo, float hi) {
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk2(u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ void upk2u(u64 v, uint32_t& lo, uint32_t& hi) { asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v)); }
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
    u64 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64 add2_rm(u64 a, u64 b) {  // round toward -inf: floor(x) + 2^23 for 0 <= x < 2^23
    u64 r;
    asm("add.rm.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64 sub2(u64 a, u64 b) {
    u64 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
    u64 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
// ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even under -fmad=false (the scalar forms are
// left alone), and also sees through fma(a, b, -0.0). The strict kernel therefore multiplies with an FFMA2
// whose addend is a (-0.0, -0.0) pair that arrives as a kernel argument: RN(a*b + -0) == RN(a*b) for every
// input including signed zeros, and the compiler cannot fold what it cannot see.
__device__ __forceinline__ u64 mulx(u64 a, u64 b, u64 neg_zero2) {
    u64 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(neg_zero2));
    return r;
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
    u64 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

__device__ __forceinline__ uint2 lds_u64x(uint32_t addr) { return *reinterpret_cast<const uint2*>(__cvta_shared_to_generic(addr)); }
__device__ __forceinline__ uint4 lds_u128(uint32_t addr) { return *reinterpret_cast<const uint4*>(__cvta_shared_to_generic(addr)); }

}  // namespace odbk
