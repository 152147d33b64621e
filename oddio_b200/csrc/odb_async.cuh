// Shared-memory, mbarrier and bulk-async-copy (TMA) helpers used by the staged kernels.
#pragma once
#include <stdint.h>

namespace odbk {

typedef unsigned long long u64;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// Shared-memory loads are plain C++ loads through the shared window (not `asm volatile`, which would pin
// them in program order and serialise the eight unrolled frames of a chunk): the compiler is free to hoist
// and interleave them between the barriers, which is where the kernel's instruction-level parallelism comes from.
__device__ __forceinline__ float lds_f32(uint32_t addr) { return *reinterpret_cast<const float*>(__cvta_shared_to_generic(addr)); }
__device__ __forceinline__ float lds_f32_4(uint32_t addr) { return *reinterpret_cast<const float*>(__cvta_shared_to_generic(addr + 4u)); }
__device__ __forceinline__ u64 lds_u64(uint32_t addr) { return *reinterpret_cast<const u64*>(__cvta_shared_to_generic(addr)); }
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) { return *reinterpret_cast<const uint32_t*>(__cvta_shared_to_generic(addr)); }
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!ok);
}

}  // namespace odbk
