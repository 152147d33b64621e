// Multi-GPU exchange of the mixed tile over NVLink peer memory (sources are sharded over the GPUs
// of one box, the only exchange is the additive 8 KiB output tile). The reference is single-process and has no
// counterpart; `Tanh`/`Reinhard` (tanh.rs:22-29, reinhard.rs:28-35) are applied here because they act on the SUM.
//
// One process per GPU. Every rank owns an inbox in its own HBM: [depth][world][cap] floats,
// [depth][world][slices] "pushed" flags and [world][slices] "pulled" acknowledgements, exported to the other ranks of the
// box as a CUDA IPC handle and mapped by them. No NCCL on the data path; two small kernels per exchange and rank, one
// CTA per 2048-float slice of the tile (a 1024-frame stereo callback is one slice; offline rendering exchanges several
// callbacks at once):
//   push: the rank stores its partial tile into slot `rank` of every rank's inbox (plain stores over NVLink; its own
//      inbox included), fences to system scope and then publishes the exchange's sequence number in the same slot's
//      flag with a release store. It never waits for a peer's data, so it can be queued right behind the mix and does
//      not hold an SM while other ranks are still mixing;
//   pull: waits until its own inbox holds this exchange's flag from every rank (acquire loads of local memory), sums
//      the `world` tiles in rank order - every rank adds the same numbers in the same order, so all ranks end up with
//      bit-identical tiles - applies the epilogue, and acknowledges the sequence number in every peer's inbox.
// Inbox slots rotate with the sequence number modulo `depth` (2..8 exchanges in flight); before a push overwrites a
// peer's slot of `depth` exchanges ago it checks that peer's acknowledgement of that exchange (practically never a wait). A caller that
// wants the sum at once queues pull right behind push (odb_exchange_allreduce); a pipelined renderer queues the pull
// one group of callbacks later, when every peer has long pushed, and nobody spins (bench.py).
#include <cuda_runtime.h>

#include "odb_host.h"
#include "odb_exchange.h"

namespace odbk {

__global__ void __launch_bounds__(512) k_exchange_push(const float* __restrict__ tile_all, int n_floats_all, ExchangePeers peers,
                                                        ExchangeGeom g, uint32_t seq) {
    const uint32_t par = seq % (uint32_t)g.depth;
    const int tid = threadIdx.x, nth = blockDim.x;
    const int first = blockIdx.x * ODB_EXCHANGE_SLICE;
    const int n_floats = min(ODB_EXCHANGE_SLICE, n_floats_all - first);
    const float* tile = tile_all + first;
    // the slot about to be overwritten held exchange seq - depth: every peer must have pulled that one
    if (tid < g.world && seq > (uint32_t)g.depth) {
        const uint32_t* ack = reinterpret_cast<const uint32_t*>(peers.inbox[g.rank] + g.acks_off) + (size_t)tid * g.max_slices + blockIdx.x;
        while ((int)(ld_acquire_sys(ack) - (seq - (uint32_t)g.depth)) < 0) __nanosleep(20);
    }
    __syncthreads();
    // 16-byte stores; the tile and the slots are 16-byte aligned
    const size_t slot = ((size_t)par * g.world + g.rank) * g.cap + first;
    const int n4 = n_floats >> 2;
    for (int p = 0; p < g.world; p++) {
        float* dst = reinterpret_cast<float*>(peers.inbox[p]) + slot;
        for (int i = tid; i < n4; i += nth) reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(tile)[i];
        for (int i = 4 * n4 + tid; i < n_floats; i += nth) dst[i] = tile[i];
    }
    __threadfence_system();
    __syncthreads();
    if (tid < g.world)
        st_release_sys(reinterpret_cast<uint32_t*>(peers.inbox[tid] + g.flags_off) + ((size_t)par * g.world + g.rank) * g.max_slices + blockIdx.x, seq);
}

__global__ void __launch_bounds__(512) k_exchange_pull(float* __restrict__ tile_all, int n_floats_all, ExchangePeers peers,
                                                        ExchangeGeom g, uint32_t seq, int epilogue) {
    const uint32_t par = seq % (uint32_t)g.depth;
    const int tid = threadIdx.x, nth = blockDim.x;
    const int first = blockIdx.x * ODB_EXCHANGE_SLICE;
    const int n_floats = min(ODB_EXCHANGE_SLICE, n_floats_all - first);
    float* tile = tile_all + first;
    const char* mine = peers.inbox[g.rank];
    if (tid < g.world) {
        const uint32_t* flag = reinterpret_cast<const uint32_t*>(mine + g.flags_off) + ((size_t)par * g.world + tid) * g.max_slices + blockIdx.x;
        while ((int)(ld_acquire_sys(flag) - seq) < 0) __nanosleep(20);
    }
    __syncthreads();
    const float* in = reinterpret_cast<const float*>(mine) + (size_t)par * g.world * g.cap + first;
    for (int i = tid; i < n_floats; i += nth) {
        float sum = 0.0f;
        for (int p = 0; p < g.world; p++) sum = sum + __ldcv(in + (size_t)p * g.cap + i);  // rank order: same sum on every rank
        if (epilogue == 1) sum = tanhf(sum);
        else if (epilogue == 2) sum = sum / (1.0f + fabsf(sum));
        tile[i] = sum;
    }
    __syncthreads();  // every thread has read its share of the inbox: the peers may overwrite this slot `depth` exchanges on
    if (tid < g.world) {
        uint32_t* ack = reinterpret_cast<uint32_t*>(peers.inbox[tid] + g.acks_off) + (size_t)g.rank * g.max_slices;
        st_release_sys(ack + blockIdx.x, seq);
        // slices this exchange did not use are free as well (a later, larger exchange checks them)
        if (blockIdx.x == 0)
            for (int sl = (int)gridDim.x; sl < g.max_slices; sl++) st_release_sys(ack + sl, seq);
    }
}

}  // namespace odbk

static int exchange_check(odb_exchange* ex) {
    if (!ex || ex->kind != ODB_KIND_EXCHANGE) return odb_fail(ODB_E_INVALID, "not an exchange handle");
    return ODB_OK;
}

extern "C" int odb_exchange_create(odb_ctx* ctx, int rank, int world, uint32_t max_floats, int depth, odb_exchange** out) {
    if (!ctx || !out) return odb_fail(ODB_E_INVALID, "NULL argument");
    if (depth < 2 || depth > ODB_MAX_DEPTH) return odb_fail(ODB_E_INVALID, "depth %d: 2..%d exchanges in flight", depth, ODB_MAX_DEPTH);
    if (world < 1 || world > ODB_MAX_RANKS || rank < 0 || rank >= world)
        return odb_fail(ODB_E_INVALID, "rank %d of %d: at most %d ranks (the GPUs of one box)", rank, world, ODB_MAX_RANKS);
    if (max_floats == 0) return odb_fail(ODB_E_INVALID, "max_floats is 0");
    ODB_CUDA(cudaSetDevice(ctx->device));
    odb_exchange* ex = new odb_exchange();
    ex->ctx = ctx;
    ex->rank = rank;
    ex->world = world;
    ex->depth = depth;
    ex->cap = (max_floats + 31u) & ~31u;  // slots stay 128-byte aligned
    ex->max_slices = (int)((ex->cap + ODB_EXCHANGE_SLICE - 1) / ODB_EXCHANGE_SLICE);
    ex->flags_off = (size_t)depth * world * ex->cap * sizeof(float);
    ex->acks_off = ex->flags_off + (size_t)depth * world * ex->max_slices * sizeof(uint32_t);
    ex->bytes = ex->acks_off + (size_t)world * ex->max_slices * sizeof(uint32_t);
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

static int exchange_args(odb_exchange* ex, const void* dev_tile, uint32_t n_floats) {
    ODB_TRY(exchange_check(ex));
    if (!dev_tile && n_floats) return odb_fail(ODB_E_INVALID, "dev_tile is NULL");
    if (!ex->connected) return odb_fail(ODB_E_INVALID, "exchange is not connected to its peers yet");
    if (n_floats == 0 || n_floats > ex->cap) return odb_fail(ODB_E_INVALID, "%u floats: between 1 and the exchange's capacity of %u", n_floats, ex->cap);
    if (((uintptr_t)dev_tile & 15u) != 0) return odb_fail(ODB_E_INVALID, "dev_tile must be 16-byte aligned");
    return ODB_OK;
}

extern "C" int odb_exchange_push(odb_exchange* ex, const void* dev_tile, uint32_t n_floats, void* cuda_stream) {
    ODB_TRY(exchange_args(ex, dev_tile, n_floats));
    if (ex->seq - ex->pulled >= (uint32_t)ex->depth)
        return odb_fail(ODB_E_INVALID, "%d exchanges are already pushed and not pulled (the depth given at creation)", ex->depth);
    ODB_CUDA(cudaSetDevice(ex->ctx->device));
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : ex->ctx->stream;
    ex->seq++;  // flags are compared as signed differences, so wrap-around is harmless
    ex->pushed_floats[ex->seq % (uint32_t)ex->depth] = n_floats;
    const int n_slices = (int)((n_floats + ODB_EXCHANGE_SLICE - 1) / ODB_EXCHANGE_SLICE);
    odbk::k_exchange_push<<<n_slices, 512, 0, st>>>((const float*)dev_tile, (int)n_floats, ex->peers, ex->geom(), ex->seq);
    ODB_CUDA(cudaGetLastError());
    return ODB_OK;
}

extern "C" int odb_exchange_pull(odb_exchange* ex, void* dev_tile, uint32_t n_floats, int epilogue, void* cuda_stream) {
    ODB_TRY(exchange_args(ex, dev_tile, n_floats));
    if (epilogue < 0 || epilogue > 2) return odb_fail(ODB_E_INVALID, "unknown epilogue %d", epilogue);
    if (ex->pulled == ex->seq) return odb_fail(ODB_E_INVALID, "nothing pushed that has not been pulled");
    const uint32_t seq = ex->pulled + 1;
    if (ex->pushed_floats[seq % (uint32_t)ex->depth] != n_floats)
        return odb_fail(ODB_E_INVALID, "pull of %u floats does not match the push of %u", n_floats, ex->pushed_floats[seq % (uint32_t)ex->depth]);
    ODB_CUDA(cudaSetDevice(ex->ctx->device));
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : ex->ctx->stream;
    ex->pulled = seq;
    const int n_slices = (int)((n_floats + ODB_EXCHANGE_SLICE - 1) / ODB_EXCHANGE_SLICE);
    odbk::k_exchange_pull<<<n_slices, 512, 0, st>>>((float*)dev_tile, (int)n_floats, ex->peers, ex->geom(), seq, epilogue);
    ODB_CUDA(cudaGetLastError());
    return ODB_OK;
}

extern "C" int odb_exchange_allreduce(odb_exchange* ex, void* dev_tile, uint32_t n_floats, int epilogue, void* cuda_stream) {
    ODB_TRY(exchange_args(ex, dev_tile, n_floats));
    if (ex->seq != ex->pulled) return odb_fail(ODB_E_INVALID, "a pushed exchange is still waiting for its pull");
    ODB_TRY(odb_exchange_push(ex, dev_tile, n_floats, cuda_stream));
    return odb_exchange_pull(ex, dev_tile, n_floats, epilogue, cuda_stream);
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
