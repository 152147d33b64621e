// Host-side runtime shared by the scene and mixer implementations of the C ABI.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <atomic>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/oddio_b200.h"
#include "odb_kernels.h"

std::string& odb_err();
int odb_fail(int code, const char* fmt, ...);

#define ODB_CUDA(call)                                                                              \
    do {                                                                                            \
        cudaError_t e__ = (call);                                                                   \
        if (e__ != cudaSuccess)                                                                     \
            return odb_fail(ODB_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)
#define ODB_TRY(call)            \
    do {                         \
        int r__ = (call);        \
        if (r__ != ODB_OK) return r__; \
    } while (0)

// Growable device array; growth happens on the control/apply path only, never mid-kernel.
template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;
    int ensure(size_t n, cudaStream_t st, bool keep) {
        if (n <= cap) return ODB_OK;
        size_t ncap = cap ? cap : 256;
        while (ncap < n) ncap *= 2;
        T* np = nullptr;
        ODB_CUDA(cudaMalloc((void**)&np, ncap * sizeof(T)));
        if (keep && p && cap) ODB_CUDA(cudaMemcpyAsync(np, p, cap * sizeof(T), cudaMemcpyDeviceToDevice, st));
        if (p) {
            ODB_CUDA(cudaStreamSynchronize(st));
            ODB_CUDA(cudaFree(p));
        }
        p = np;
        cap = ncap;
        return ODB_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};
template <class T>
struct PinBuf {
    T* p = nullptr;
    size_t cap = 0;
    int ensure(size_t n) {
        if (n <= cap) return ODB_OK;
        size_t ncap = cap ? cap : 256;
        while (ncap < n) ncap *= 2;
        if (p) cudaFreeHost(p);
        p = nullptr;
        ODB_CUDA(cudaMallocHost((void**)&p, ncap * sizeof(T)));
        cap = ncap;
        return ODB_OK;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
};

struct FramesRec {  // Arc<Frames<T>> (frames.rs:16-22)
    float* dev = nullptr;  // first sample; ODB_PCM_PAD zero floats on both sides
    uint64_t n_frames = 0;
    int channels = 1;
    uint32_t rate = 0;
    int refs = 0;
    int block = -1;
};
struct ArenaBlock {
    char* base = nullptr;
    size_t size = 0, used = 0;
    int live = 0;
};

static inline void odb_cpu_pause() {
#if defined(__x86_64__)
    __builtin_ia32_pause();
#endif
}
// The audio thread never blocks on the control plane (signal.rs:11-13): it announces itself - batched control calls
// check the flag between chunks of 256 messages and stand back - and then TRIES the control-plane mutex a bounded
// number of times. If a control call still holds it (a thread preempted inside a control call, say), the callback
// goes ahead with the state it already has: what was queued is applied one callback later, which the reference's
// swap / spsc transports allow as well (a message written while `refresh` runs is seen by the next one,
// swap.rs:57-64). `held()` says whether the control-plane queues may be touched.
struct AudioLock {
    std::mutex& mu;
    bool got = false;
    AudioLock(std::mutex& m, std::atomic<int>& wants, int tries = 512) : mu(m) {
        wants.fetch_add(1, std::memory_order_acq_rel);
        for (int i = 0; i < tries && !(got = mu.try_lock()); i++) odb_cpu_pause();
        wants.fetch_sub(1, std::memory_order_acq_rel);
    }
    bool held() const { return got; }
    ~AudioLock() { if (got) mu.unlock(); }
    AudioLock(const AudioLock&) = delete;
    AudioLock& operator=(const AudioLock&) = delete;
};
static inline void odb_yield_to_audio(const std::atomic<int>& wants) {
    while (wants.load(std::memory_order_acquire) > 0) odb_cpu_pause();
}

struct odb_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = true;
    std::mutex mu;
    std::unordered_map<uint64_t, FramesRec> frames;
    uint64_t next_frames_id = 1;
    std::vector<ArenaBlock> blocks;
    std::vector<char*> dead_blocks;  // blocks whose last allocation died on the audio thread: freed by the next control call
    // Delay rings of finished buffered sources, by byte size: a steady stream of play_buffered sounds reuses them
    // instead of bump-allocating new arena space (a ring lives in a block shared with long-lived PCM, which the
    // block-level reference count would keep pinned).
    struct FreeRing { float* p; int block; };
    std::unordered_map<size_t, std::vector<FreeRing>> free_rings;
    int arena_alloc(size_t bytes, float** out, int* block);
    // `defer`: called from the audio thread - never cudaFree there; the block goes to dead_blocks
    void arena_unref(int block, bool defer = false);
    void arena_collect();            // control side: free what the audio thread left in dead_blocks
    int ring_alloc(size_t bytes, float** out, int* block);
    void ring_release(float* p, size_t bytes, int block);
    int frames_ref(odb_frames id, FramesRec* out);
    void frames_unref(odb_frames id);
};

// Builds the device record of FramesSignal::new(frames, start) under the chain's wrappers; takes a reference on the Frames.
int odb_make_source(odb_ctx* ctx, const odb_chain* chain, int want_channels, OdbSource* out, FramesRec* rec);

#define ODB_KIND_SCENE 0x5343454Eu
#define ODB_KIND_MIXER 0x4D495852u

struct SlotHost {
    odb_frames frames = 0;
    uint32_t gen = 1;
    bool in_use = false;
    bool stopped = false;   // Spatial::is_finished / Mixed::is_stopped as seen by the control side
    bool stop_requested = false;  // Mixed::stop() was called (mixer.rs:34-36); visible to is_stopped at once
    uint32_t chain_flags = 0;
    int ring_block = -1;    // arena block of a buffered source's delay ring, -1 if none
    float* ring_ptr = nullptr;  // ... the ring itself and its size, so that it can be reused when the source is gone
    size_t ring_bytes = 0;
    uint64_t n_frames = 0;  // FramesSignalControl::samples
    double rate = 0.0;
    // latest-wins de-duplication of queued control messages (swap.rs semantics): index into the
    // pending vectors, -1 if nothing is queued for this slot since the last apply()
    int motion_idx = -1, speed_idx = -1, gain_idx = -1;
    uint32_t motion_gen = 0;  // motion_idx is valid only while this equals SourceSet::mot_gen
};

// A set of playing sources living in HBM plus the control-plane queues feeding it; the common
// part of odb_scene (seek set, buffered set) and odb_mixer.
struct SourceSet {
    std::vector<SlotHost> slots;
    std::vector<uint32_t> free_slots;
    std::vector<uint32_t> order;        // the reference's Vec order (set.rs:206): slot ids
    bool order_dirty = false;
    // queued by the control side, applied at the next sample()
    std::vector<OdbSource> ins_src;
    std::vector<uint32_t> ins_slot;
    // set_motion messages are written by the control side straight into one of two pinned buffers (at most one
    // message per slot: latest value wins); apply() sends the filled one to the device as it lies and hands the
    // control side the other. `motions` only takes what does not fit (sources played since the last apply()).
    PinBuf<OdbMotionMsg> h_mot[2];
    cudaEvent_t ev_mot[2] = {nullptr, nullptr};  // behind the H2D copy out of h_mot[b]
    bool ev_mot_pending[2] = {false, false};
    int mot_buf = 0;
    size_t mot_n = 0;
    uint32_t mot_gen = 1;
    std::vector<OdbMotionMsg> motions;
    std::vector<OdbParamMsg> params;
    // device
    DevBuf<OdbSource> d_src;
    DevBuf<uint32_t> d_order;
    DevBuf<OdbSource> d_stage_src;
    DevBuf<uint32_t> d_stage_slot;
    DevBuf<OdbMotionMsg> d_motions;
    DevBuf<OdbParamMsg> d_params;
    DevBuf<uint32_t> d_removed;         // [0] = monotonic count of removals the walk kernels reported,
                                        // [1 + (k & (removed_cap-1))] = slot of the k-th report
    PinBuf<OdbSource> h_stage_src;
    PinBuf<uint32_t> h_stage_slot;
    PinBuf<OdbMotionMsg> h_motions;
    PinBuf<OdbParamMsg> h_params;
    PinBuf<uint32_t> h_order;
    PinBuf<uint32_t> h_removed;         // report entries, fetched only when the count moved
    PinBuf<uint32_t> h_removed_count;   // landing slot of the asynchronous count read-back
    uint32_t removed_cap = 0;           // ring capacity: power of two >= live sources
    uint32_t removed_consumed = 0;      // reports already folded into `order`
    cudaEvent_t ev_removed = nullptr;   // recorded behind the count read-back
    bool count_in_flight = false;
    std::vector<int> pos_of_slot;       // position of a slot in `order`, -1 if not a member
    std::mutex grow_mu;                 // held while the source table is reallocated (rare: the set outgrew it) and while a
                                        // control-side read-back copies a record out of it

    uint32_t alloc_slot();
    odb_source handle_of(uint32_t slot, uint32_t tag) const;
    // ODB_OK and *slot if `h` names a live source; *stale = true (and ODB_OK) if it names a source that
    // has since been removed (the reference's handles outlive their signal); error otherwise.
    int lookup(odb_source h, uint32_t tag, uint32_t* slot, bool* stale) const;
    void queue_motion(uint32_t slot, const float* pos, const float* vel, int disc);
    void queue_param(uint32_t slot, uint32_t what, float value);
    // audio side: push queued control messages to the device; returns kernels launched
    int apply(odb_ctx* ctx, cudaStream_t st, uint32_t* launches);
    // audio side, after a callback's kernels are queued: start the asynchronous read-back of the report count
    int post_callback(odb_ctx* ctx, cudaStream_t st);
    // audio side: fold the removals the kernels reported into `order` (set.rs:183-188 swap_remove). With
    // wait = false only reports whose read-back has already completed are folded (no host-device sync).
    // `mu` (may be NULL when the caller already holds it) is the owner's control-plane mutex: it is taken only
    // when there is something to fold, so a callback without removals never contends with the control thread.
    int fold_removed(odb_ctx* ctx, cudaStream_t st, bool wait, std::mutex* mu);
    // the same with the report count already in hand (the callback kernel stored it into h_removed_count)
    int fold_count(odb_ctx* ctx, cudaStream_t st, uint32_t count, std::mutex* mu);
    void release_all(odb_ctx* ctx);
};
