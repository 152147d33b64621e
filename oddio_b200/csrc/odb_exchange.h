// Multi-GPU exchange of the mixed tile over NVLink peer memory: the inbox geometry and host-side state shared by
// odb_exchange.cu (stand-alone push / pull kernels) and the one-launch callback kernel (odb_scene_mix.cu), which
// pushes - and optionally pulls - from its reduce phase. Protocol: see odb_exchange.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define ODB_KIND_EXCHANGE 0x58434847u
#define ODB_MAX_RANKS 16
#define ODB_MAX_DEPTH 8

namespace odbk {

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

struct ExchangePeers {
    char* inbox[ODB_MAX_RANKS];  // inbox of every rank as mapped into this process ([rank] = the local one)
};

#define ODB_EXCHANGE_SLICE 2048  // floats per CTA

struct ExchangeGeom {
    int rank, world, max_slices, depth;
    uint32_t cap;        // floats per inbox slot
    size_t flags_off;    // byte offset of the pushed flags [depth][world][max_slices]
    size_t acks_off;     // byte offset of the pulled acknowledgements [world][max_slices]
};

}  // namespace odbk

struct odb_ctx;
struct odb_exchange {
    uint32_t kind = ODB_KIND_EXCHANGE;
    odb_ctx* ctx = nullptr;
    int rank = 0, world = 1;
    uint32_t cap = 0;          // floats per slot
    int max_slices = 1, depth = 2;
    size_t flags_off = 0, acks_off = 0, bytes = 0;
    char* local = nullptr;
    odbk::ExchangePeers peers;
    bool connected = false;
    uint32_t seq = 0;          // exchanges pushed
    uint32_t pulled = 0;       // exchanges pulled
    uint32_t pushed_floats[ODB_MAX_DEPTH] = {0};
    odbk::ExchangeGeom geom() const { return odbk::ExchangeGeom{rank, world, max_slices, depth, cap, flags_off, acks_off}; }
};
