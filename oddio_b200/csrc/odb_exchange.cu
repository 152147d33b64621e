// Multi-GPU exchange of the mixed tile over NVLink peer memory (sources are sharded over the GPUs
// of one box, the only exchange is the additive 8 KiB output tile). The reference is single-process and has no
// counterpart; `Tanh`/`Reinhard` (tanh.rs:22-29, reinhard.rs:28-35) are applied here because they act on the SUM.
//
// One process per GPU. Every rank owns an inbox in its own HBM: [2 parities][world][cap] floats plus
// [2][world][slices] sequence flags, exported to the other ranks of the box as a CUDA IPC handle and mapped by them.
// One kernel per exchange and rank, one CTA per 2048-float slice of the tile (a 1024-frame stereo callback is one
// slice; offline rendering exchanges several callbacks at once), no NCCL on the data path. Per slice:
//   1. push: the rank stores its partial tile into slot `rank` of every rank's inbox (plain stores over NVLink;
//      its own inbox included), fences to system scope and then publishes the callback's sequence number in the
//      same slot's flag with a release store;
//   2. pull: it waits until its own inbox holds this callback's flag from every rank (acquire loads of local
//      memory), sums the `world` tiles in rank order - every rank adds the same numbers in the same order, so all
//      ranks end up with bit-identical tiles - and applies the epilogue.
// Inbox slots alternate with the parity of the sequence number: a rank can only reach callback k+2 after every
// peer has pushed k+1, i.e. after every peer has finished pulling k (kernels of one rank run in stream order), so
// two parities suffice and nobody overwrites a tile that is still being read.
#include <cuda_runtime.h>

#include "odb_host.h"

#define ODB_KIND_EXCHANGE 0x58434847u
#define ODB_MAX_RANKS 16

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

__global__ void __launch_bounds__(512) k_exchange_tiles(float* __restrict__ tile_all, int n_floats_all, ExchangePeers peers,
                                                         int rank, int world, uint32_t cap, size_t flags_off, int max_slices,
                                                         uint32_t seq, int epilogue) {
    const uint32_t par = seq & 1u;
    const int tid = threadIdx.x, nth = blockDim.x;
    const int first = blockIdx.x * ODB_EXCHANGE_SLICE;
    const int n_floats = min(ODB_EXCHANGE_SLICE, n_floats_all - first);
    float* tile = tile_all + first;
    // 1. push (16-byte stores; the tile and the slots are 16-byte aligned)
    const size_t slot = ((size_t)par * world + rank) * cap + first;
    const int n4 = n_floats >> 2;
    for (int g = 0; g < world; g++) {
        float* dst = reinterpret_cast<float*>(peers.inbox[g]) + slot;
        for (int i = tid; i < n4; i += nth) reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(tile)[i];
        for (int i = 4 * n4 + tid; i < n_floats; i += nth) dst[i] = tile[i];
    }
    __threadfence_system();
    __syncthreads();
    const size_t flag_idx = ((size_t)par * world) * max_slices + blockIdx.x;  // + rank * max_slices
    if (tid < world)
        st_release_sys(reinterpret_cast<uint32_t*>(peers.inbox[tid] + flags_off) + flag_idx + (size_t)rank * max_slices, seq);
    // 2. pull
    const char* mine = peers.inbox[rank];
    if (tid < world) {
        const uint32_t* flag = reinterpret_cast<const uint32_t*>(mine + flags_off) + flag_idx + (size_t)tid * max_slices;
        while ((int)(ld_acquire_sys(flag) - seq) < 0) __nanosleep(20);
    }
    __syncthreads();
    const float* in = reinterpret_cast<const float*>(mine) + (size_t)par * world * cap + first;
    for (int i = tid; i < n_floats; i += nth) {
        float sum = 0.0f;
        for (int g = 0; g < world; g++) sum = sum + __ldcv(in + (size_t)g * cap + i);  // rank order: same sum on every rank
        if (epilogue == 1) sum = tanhf(sum);
        else if (epilogue == 2) sum = sum / (1.0f + fabsf(sum));
        tile[i] = sum;
    }
}

}  // namespace odbk

struct odb_exchange {
    uint32_t kind = ODB_KIND_EXCHANGE;
    odb_ctx* ctx = nullptr;
    int rank = 0, world = 1;
    uint32_t cap = 0;          // floats per slot
    int max_slices = 1;
    size_t flags_off = 0, bytes = 0;
    char* local = nullptr;
    odbk::ExchangePeers peers;
    bool connected = false;
    uint32_t seq = 0;
};

static int exchange_check(odb_exchange* ex) {
    if (!ex || ex->kind != ODB_KIND_EXCHANGE) return odb_fail(ODB_E_INVALID, "not an exchange handle");
    return ODB_OK;
}

extern "C" int odb_exchange_create(odb_ctx* ctx, int rank, int world, uint32_t max_floats, odb_exchange** out) {
    if (!ctx || !out) return odb_fail(ODB_E_INVALID, "NULL argument");
    if (world < 1 || world > ODB_MAX_RANKS || rank < 0 || rank >= world)
        return odb_fail(ODB_E_INVALID, "rank %d of %d: at most %d ranks (the GPUs of one box)", rank, world, ODB_MAX_RANKS);
    if (max_floats == 0) return odb_fail(ODB_E_INVALID, "max_floats is 0");
    ODB_CUDA(cudaSetDevice(ctx->device));
    odb_exchange* ex = new odb_exchange();
    ex->ctx = ctx;
    ex->rank = rank;
    ex->world = world;
    ex->cap = (max_floats + 31u) & ~31u;  // slots stay 128-byte aligned
    ex->max_slices = (int)((ex->cap + ODB_EXCHANGE_SLICE - 1) / ODB_EXCHANGE_SLICE);
    ex->flags_off = (size_t)2 * world * ex->cap * sizeof(float);
    ex->bytes = ex->flags_off + (size_t)2 * world * ex->max_slices * sizeof(uint32_t);
    for (int g = 0; g < ODB_MAX_RANKS; g++) ex->peers.inbox[g] = nullptr;
    cudaError_t e = cudaMalloc((void**)&ex->local, ex->bytes);
    if (e != cudaSuccess) {
        delete ex;
        return odb_fail(ODB_E_NOMEM, "cudaMalloc of the %zu-byte inbox failed: %s", ex->bytes, cudaGetErrorString(e));
    }
    ODB_CUDA(cudaMemset(ex->local, 0, ex->bytes));
    ODB_CUDA(cudaDeviceSynchronize());  // zeroed flags are in place before the handle can reach a peer
    ex->peers.inbox[rank] = ex->local;
    ex->connected = world == 1;
    *out = ex;
    return ODB_OK;
}

extern "C" int odb_exchange_handle_size(void) { return (int)sizeof(cudaIpcMemHandle_t); }

extern "C" int odb_exchange_export(odb_exchange* ex, void* handle_out) {
    ODB_TRY(exchange_check(ex));
    if (!handle_out) return odb_fail(ODB_E_INVALID, "NULL argument");
    ODB_CUDA(cudaSetDevice(ex->ctx->device));
    cudaIpcMemHandle_t h;
    ODB_CUDA(cudaIpcGetMemHandle(&h, ex->local));
    memcpy(handle_out, &h, sizeof h);
    return ODB_OK;
}

extern "C" int odb_exchange_connect(odb_exchange* ex, const void* handles) {
    ODB_TRY(exchange_check(ex));
    if (!handles) return odb_fail(ODB_E_INVALID, "NULL argument");
    if (ex->connected) return odb_fail(ODB_E_INVALID, "exchange is already connected");
    ODB_CUDA(cudaSetDevice(ex->ctx->device));
    const char* hs = (const char*)handles;
    for (int g = 0; g < ex->world; g++) {
        if (g == ex->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, hs + (size_t)g * sizeof h, sizeof h);
        void* p = nullptr;
        ODB_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        ex->peers.inbox[g] = (char*)p;
    }
    ex->connected = true;
    return ODB_OK;
}

extern "C" int odb_exchange_allreduce(odb_exchange* ex, void* dev_tile, uint32_t n_floats, int epilogue, void* cuda_stream) {
    ODB_TRY(exchange_check(ex));
    if (!dev_tile && n_floats) return odb_fail(ODB_E_INVALID, "dev_tile is NULL");
    if (!ex->connected) return odb_fail(ODB_E_INVALID, "exchange is not connected to its peers yet");
    if (n_floats > ex->cap) return odb_fail(ODB_E_INVALID, "%u floats exceed the exchange's capacity of %u", n_floats, ex->cap);
    if (epilogue < 0 || epilogue > 2) return odb_fail(ODB_E_INVALID, "unknown epilogue %d", epilogue);
    if (((uintptr_t)dev_tile & 15u) != 0) return odb_fail(ODB_E_INVALID, "dev_tile must be 16-byte aligned");
    ODB_CUDA(cudaSetDevice(ex->ctx->device));
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : ex->ctx->stream;
    ex->seq++;  // flags are compared as signed differences, so wrap-around is harmless
    if (n_floats == 0) return ODB_OK;
    const int n_slices = (int)((n_floats + ODB_EXCHANGE_SLICE - 1) / ODB_EXCHANGE_SLICE);
    odbk::k_exchange_tiles<<<n_slices, 512, 0, st>>>((float*)dev_tile, (int)n_floats, ex->peers, ex->rank, ex->world, ex->cap,
                                                     ex->flags_off, ex->max_slices, ex->seq, epilogue);
    ODB_CUDA(cudaGetLastError());
    return ODB_OK;
}

extern "C" int odb_exchange_destroy(odb_exchange* ex) {
    if (!ex) return ODB_OK;
    ODB_TRY(exchange_check(ex));
    cudaSetDevice(ex->ctx->device);
    cudaDeviceSynchronize();
    for (int g = 0; g < ex->world; g++)
        if (g != ex->rank && ex->peers.inbox[g]) cudaIpcCloseMemHandle(ex->peers.inbox[g]);
    if (ex->local) cudaFree(ex->local);
    ex->kind = 0;
    delete ex;
    return ODB_OK;
}
