// odb_scene_*: SpatialSceneControl + SpatialScene over the device-resident source sets.
// Reference: src/spatial.rs (cited per function), src/lib.rs:90-93.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "odb_host.h"
#include "odb_exchange.h"

#define ODB_TAG_SEEK 1u
#define ODB_TAG_BUFFERED 2u
#define ODB_MAX_FRAMES (4 * ODB_TILE_FRAMES)

struct odb_scene {
    uint32_t kind = ODB_KIND_SCENE;
    odb_ctx* ctx = nullptr;
    std::mutex mu;  // guards the control-plane queues of both sets and the rotation
    std::atomic<int> audio_wants{0};  // the audio thread is waiting for `mu` (see AudioLock)
    SourceSet seek, buffered;
    // swap::Receiver<Quaternion> (spatial.rs:161): received + pending, stored already inverted (:346)
    OdbQuat rot_received = {0.0f, 0.0f, 0.0f, 1.0f};
    OdbQuat rot_pending = {0.0f, 0.0f, 0.0f, 1.0f};
    bool rot_fresh = false;
    int epilogue = ODB_EPILOGUE_NONE;
    int variant = 2;  // 0: strict (unfused value arithmetic), 1: literal kernels only, 2: value multiply-adds contracted to FMA
    uint32_t last_launches = 0;
    // Optional two-stage pipeline (odb_set_kernel_variant bit 8): the per-source set-up of callback k+1
    // (control-plane scatter + walk kernels, on the scene's own `wst` stream) overlaps the mix kernels of
    // callback k (on the context's stream). Measured on C3 it hides only ~4 us of the 15 us walk kernel (the mix
    // kernel leaves it no registers or issue slots), so the default is one stream with programmatic dependent
    // launches between the kernels of a callback. What the walk kernels hand to the mix kernels is
    // double-buffered by callback parity either way.
    bool pipelined = false;
    cudaStream_t wst = nullptr;
    cudaEvent_t ev_walk[2] = {nullptr, nullptr}, ev_mix[2] = {nullptr, nullptr};
    uint64_t callback_no = 0;
    DevBuf<OdbJob> d_jobs[2];
    DevBuf<uint32_t> d_counters[2];
    DevBuf<OdbRingJob> d_ring_jobs[2];
    DevBuf<OdbRingWrite> d_ring_writes[2];
    DevBuf<uint32_t> d_ring_list[2];   // (tile, source) jobs the literal ring kernel takes, written by k_walk_buffered
    DevBuf<float> d_partials;
    DevBuf<float> d_partials_fast;
    DevBuf<float> d_partials_ring;
    // One-launch callback (odb_scene_mix.cu), the default for scenes without buffered sources: partial tiles by
    // callback parity (the next callback's CTAs may start while this one's reducers still read), the grid's
    // arrive / done counters and the running totals the kernel compares them with.
    DevBuf<float> d_partials_fused[2];
    DevBuf<float> d_xtile[2];            // this rank's raw sum, handed to the grid's last CTA (exchange / host tile)
    DevBuf<unsigned long long> d_sync;   // [0..1] arrivals by callback parity, [2..3] finished CTAs by parity, [4] last completed launch
    unsigned long long arrive_total[2] = {0, 0}, done_total[2] = {0, 0};
    unsigned long long fused_seq = 0;    // launches of the one-launch kernel so far
    unsigned long long walked_total = 0; // walk blocks launched for it so far (d_sync[5] counts the finished ones)
    DevBuf<uint32_t> d_counters_ring;    // four sets of job counters: callback k uses set k % 4 (see odb_walk.cuh)
    uint32_t* last_counters = nullptr;   // the set the last callback counted into
    DevBuf<OdbJob> d_jobs_ring[3];       // the one-launch callback's job records: launch k uses set k % 3
    bool legacy = false;                 // odb_set_kernel_variant bit 9: the multi-kernel path of round 1
    bool flag_armed = false;             // the callback just queued publishes flag_seq to h_flag when its tile is stored
    bool count_by_kernel = false;        // ... and the seek set's removal-report count to seek.h_removed_count
    const void* pinned_probe = nullptr;  // the caller buffer odb_scene_sample last looked up, and what it found
    float* pinned_dev = nullptr;
    PinBuf<unsigned long long> h_flag;   // sequence number of the last callback whose tile has landed in h_out
    unsigned long long flag_seq = 0;
    DevBuf<float> d_out;
    bool profiling = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    PinBuf<float> h_out;
    // ODB_TRACE=1: host-side time of the phases of odb_scene_sample, printed when the scene is destroyed
    double tr_enqueue = 0.0, tr_wait = 0.0, tr_tail = 0.0, tr_lock = 0.0, tr_apply = 0.0, tr_enqueue_max = 0.0;
    uint64_t tr_calls = 0;
    double tr_seg[8] = {0, 0, 0, 0, 0, 0, 0, 0};
};
static inline double now_us() {
    return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
static const bool g_trace = getenv("ODB_TRACE") != nullptr;

static int scene_check(odb_scene* s) {
    if (!s || s->kind != ODB_KIND_SCENE) return odb_fail(ODB_E_INVALID, "not a scene handle");
    return ODB_OK;
}

extern "C" int odb_scene_create(odb_ctx* ctx, odb_scene** out) {
    if (!ctx || !out) return odb_fail(ODB_E_INVALID, "NULL argument");
    ODB_CUDA(cudaSetDevice(ctx->device));
    odb_scene* s = new odb_scene();
    s->ctx = ctx;
    ODB_CUDA(cudaStreamCreateWithFlags(&s->wst, cudaStreamNonBlocking));
    for (int p = 0; p < 2; p++) {
        ODB_CUDA(cudaEventCreateWithFlags(&s->ev_walk[p], cudaEventDisableTiming));
        ODB_CUDA(cudaEventCreateWithFlags(&s->ev_mix[p], cudaEventDisableTiming));
    }
    *out = s;
    return ODB_OK;
}
extern "C" int odb_scene_destroy(odb_scene* scene) {
    if (!scene) return ODB_OK;
    ODB_TRY(scene_check(scene));
    cudaSetDevice(scene->ctx->device);
    cudaStreamSynchronize(scene->wst);
    cudaStreamSynchronize(scene->ctx->stream);
    if (g_trace && scene->tr_calls)
        fprintf(stderr, "[odb trace] odb_scene_sample x%llu: enqueue %.1f us (max %.1f; of which lock wait %.1f, control-plane apply %.1f), wait for the tile %.1f us, tail %.1f us\n",
                (unsigned long long)scene->tr_calls, scene->tr_enqueue / scene->tr_calls, scene->tr_enqueue_max,
                scene->tr_lock / scene->tr_calls, scene->tr_apply / scene->tr_calls, scene->tr_wait / scene->tr_calls,
                scene->tr_tail / scene->tr_calls);
    if (g_trace && scene->tr_calls)
        fprintf(stderr, "[odb trace]   enqueue segments: walk %.1f, staged mix %.1f, rest of the kernels %.1f, removal read-back %.1f us\n",
                scene->tr_seg[0] / scene->tr_calls, scene->tr_seg[1] / scene->tr_calls, scene->tr_seg[2] / scene->tr_calls,
                scene->tr_seg[3] / scene->tr_calls);
    scene->seek.release_all(scene->ctx);
    scene->buffered.release_all(scene->ctx);
    for (int p = 0; p < 2; p++) {
        scene->d_jobs[p].release(); scene->d_counters[p].release();
        scene->d_ring_jobs[p].release(); scene->d_ring_writes[p].release(); scene->d_ring_list[p].release();
        cudaEventDestroy(scene->ev_walk[p]); cudaEventDestroy(scene->ev_mix[p]);
    }
    cudaStreamDestroy(scene->wst);
    scene->d_partials.release(); scene->d_partials_fast.release(); scene->d_partials_ring.release();
    scene->d_partials_fused[0].release(); scene->d_partials_fused[1].release(); scene->d_sync.release(); scene->h_flag.release();
    scene->d_xtile[0].release(); scene->d_xtile[1].release(); scene->d_counters_ring.release();
    for (int q = 0; q < 3; q++) scene->d_jobs_ring[q].release();
    scene->d_out.release();
    scene->h_out.release();
    if (scene->ev0) { cudaEventDestroy(scene->ev0); cudaEventDestroy(scene->ev1); }
    scene->kind = 0;
    delete scene;
    return ODB_OK;
}
extern "C" int odb_scene_set_epilogue(odb_scene* scene, int epilogue) {
    ODB_TRY(scene_check(scene));
    if (epilogue < 0 || epilogue > 2) return odb_fail(ODB_E_INVALID, "unknown epilogue %d", epilogue);
    scene->epilogue = epilogue;
    return ODB_OK;
}

// Builds the device record of FramesSignal::new(frames, start) under the chain's wrappers.
int odb_make_source(odb_ctx* ctx, const odb_chain* chain, int want_channels, OdbSource* out, FramesRec* rec) {
    if (!chain) return odb_fail(ODB_E_INVALID, "chain is NULL");
    ODB_TRY(ctx->frames_ref(chain->frames, rec));
    if (rec->channels != want_channels) {
        ctx->frames_unref(chain->frames);
        return odb_fail(ODB_E_UNSUPPORTED, "signal has %d channel(s), %d required", rec->channels, want_channels);
    }
    OdbSource s;
    memset(&s, 0, sizeof s);
    s.pcm = rec->dev;
    s.rate = (double)rec->rate;                                   // frames.rs:40
    s.t = chain->start_seconds;                                   // frames.rs:159
    s.t_end = (double)(rec->n_frames - 1) / s.rate;               // frames.rs:205, evaluated once
    s.sample_t = (long long)(chain->start_seconds * s.rate);      // frames.rs:160
    s.len = (int)rec->n_frames;
    s.channels = rec->channels;
    s.speed = (chain->flags & ODB_CHAIN_SPEED) ? chain->speed : 1.0f;
    s.fixed_gain = (chain->flags & ODB_CHAIN_FIXED_GAIN) ? powf(10.0f, chain->fixed_gain_db / 20.0f) : 1.0f;  // gain.rs:20
    float g = (chain->flags & ODB_CHAIN_GAIN) ? chain->gain_ratio : 1.0f;
    s.gain_shared = g; s.gain_prev = g; s.gain_next = g; s.gain_progress = 1.0f;                              // gain.rs:90-93
    if (chain->flags & ODB_CHAIN_SPEED) s.flags |= ODB_SF_SPEED;
    if (chain->flags & ODB_CHAIN_FIXED_GAIN) s.flags |= ODB_SF_FIXED_GAIN;
    if (chain->flags & ODB_CHAIN_GAIN) s.flags |= ODB_SF_GAIN;
    if (chain->flags & ODB_CHAIN_CYCLE) {                         // Cycle::new (cycle.rs:15-20) + earlier seeks
        if (!(chain->start_seconds >= 0.0 && chain->start_seconds < (double)rec->n_frames)) {
            ctx->frames_unref(chain->frames);
            return odb_fail(ODB_E_INVALID, "Cycle cursor %g outside [0, %llu)", chain->start_seconds, (unsigned long long)rec->n_frames);
        }
        s.flags |= ODB_SF_CYCLE;
        s.t_end = 1.0 / 0.0;                                      // Signal::is_finished default: never (signal.rs:25-27)
        s.sample_t = 0;
    }
    *out = s;
    return ODB_OK;
}

// SpatialSceneControl::play, spatial.rs:289-302 (+ SpatialSignal::new :66-81, Common::new :94-116)
extern "C" int odb_scene_play(odb_scene* scene, const odb_chain* chain, const float position[3], const float velocity[3],
                              float radius, odb_source* out) {
    ODB_TRY(scene_check(scene));
    if (!chain || !position || !velocity || !out) return odb_fail(ODB_E_INVALID, "NULL argument");
    if (chain->flags & (ODB_CHAIN_SPEED | ODB_CHAIN_GAIN))
        return odb_fail(ODB_E_UNSUPPORTED, "SpatialSceneControl::play requires Seek; Speed and Gain do not implement it (use play_buffered)");
    OdbSource s;
    FramesRec rec;
    ODB_TRY(odb_make_source(scene->ctx, chain, 1, &s, &rec));
    s.radius = radius;
    for (int k = 0; k < 3; k++) {
        s.pos[k] = position[k]; s.vel[k] = velocity[k];          // Motion, discontinuity: false (:100-104)
        s.ppos[k] = position[k]; s.pvel[k] = velocity[k];
        s.prev_position[k] = position[k];                         // State::new (:494-499)
    }
    s.state_dt = 0.0f;
    std::lock_guard<std::mutex> lk(scene->mu);
    uint32_t slot = scene->seek.alloc_slot();
    SlotHost& sh = scene->seek.slots[slot];
    sh.frames = chain->frames; sh.chain_flags = chain->flags; sh.n_frames = rec.n_frames; sh.rate = (double)rec.rate;
    scene->seek.ins_src.push_back(s);
    scene->seek.ins_slot.push_back(slot);
    *out = scene->seek.handle_of(slot, ODB_TAG_SEEK);
    return ODB_OK;
}

// SpatialSceneControl::play_buffered (spatial.rs:314-340) + SpatialSignalBuffered::new (:30-56)
extern "C" int odb_scene_play_buffered(odb_scene* scene, const odb_chain* chain, const float position[3],
                                       const float velocity[3], float radius, float max_distance, uint32_t rate,
                                       float buffer_duration, odb_source* out) {
    ODB_TRY(scene_check(scene));
    if (!chain || !position || !velocity || !out) return odb_fail(ODB_E_INVALID, "NULL argument");
    if (rate == 0) return odb_fail(ODB_E_INVALID, "rate must be nonzero");
    if (chain->flags & ODB_CHAIN_CYCLE) return odb_fail(ODB_E_UNSUPPORTED, "device path: Cycle plays under a Mixer only");
    odb_ctx* ctx = scene->ctx;
    OdbSource s;
    FramesRec rec;
    ODB_TRY(odb_make_source(ctx, chain, 1, &s, &rec));
    s.radius = radius;
    for (int k = 0; k < 3; k++) {
        s.pos[k] = position[k]; s.vel[k] = velocity[k];
        s.ppos[k] = position[k]; s.pvel[k] = velocity[k];
        s.prev_position[k] = position[k];
    }
    s.state_dt = 0.0f;
    const float max_delay = max_distance / 343.0f + buffer_duration;             // spatial.rs:330
    const float capf = ceilf(max_delay * (float)rate);                           // spatial.rs:39
    if (!(capf >= 1.0f) || capf > 8388607.0f) {  // below 2^23: ring offsets stay exactly representable cursors (ring.rs:6)
        ctx->frames_unref(chain->frames);
        return odb_fail(ODB_E_INVALID, "delay ring of %g samples is out of range", (double)capf);
    }
    const int cap = (int)capf + 1;
    // Ring::new (ring.rs:10-15): zeroed buffer in HBM
    float* ring = nullptr;
    int block = -1;
    ODB_CUDA(cudaSetDevice(ctx->device));
    int rc;
    ctx->arena_collect();
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        rc = ctx->ring_alloc((size_t)cap * sizeof(float), &ring, &block);
    }
    if (rc != ODB_OK) {
        ctx->frames_unref(chain->frames);
        return rc;
    }
    ODB_CUDA(cudaMemsetAsync(ring, 0, (size_t)cap * sizeof(float), ctx->stream));
    // queue.delay(rate, (norm(position) / SPEED_OF_SOUND).min(max_delay))   spatial.rs:40-43, ring.rs:45-47
    float acc = 0.0f;
    acc = acc + position[0] * position[0];
    acc = acc + position[1] * position[1];
    acc = acc + position[2] * position[2];
    const float dist = sqrtf(acc);                                               // math/mod.rs:33-35
    const float dt0 = fminf(dist / 343.0f, max_delay);
    s.ring = ring;
    s.ring_cap = cap;
    s.ring_write = fmodf(0.0f + (float)rate * dt0, (float)cap);                  // ring.rs:46
    s.max_delay = max_delay;
    s.ring_rate = rate;
    std::lock_guard<std::mutex> lk(scene->mu);
    uint32_t slot = scene->buffered.alloc_slot();
    SlotHost& sh = scene->buffered.slots[slot];
    sh.frames = chain->frames; sh.chain_flags = chain->flags; sh.n_frames = rec.n_frames; sh.rate = (double)rec.rate;
    sh.ring_block = block;
    sh.ring_ptr = ring;
    sh.ring_bytes = (size_t)cap * sizeof(float);
    scene->buffered.ins_src.push_back(s);
    scene->buffered.ins_slot.push_back(slot);
    *out = scene->buffered.handle_of(slot, ODB_TAG_BUFFERED);
    return ODB_OK;
}

// SpatialSceneControl::set_listener_rotation, spatial.rs:345-349 (stores the conjugate, math/mod.rs:62-67)
extern "C" int odb_scene_set_listener_rotation(odb_scene* scene, const float q[4]) {
    ODB_TRY(scene_check(scene));
    if (!q) return odb_fail(ODB_E_INVALID, "NULL argument");
    std::lock_guard<std::mutex> lk(scene->mu);
    scene->rot_pending = OdbQuat{-q[0], -q[1], -q[2], q[3]};
    scene->rot_fresh = true;
    return ODB_OK;
}

static SourceSet* set_of(odb_scene* scene, odb_source src, uint32_t* tag) {
    *tag = (uint32_t)((src >> 32) & 0xFF);
    return *tag == ODB_TAG_SEEK ? &scene->seek : (*tag == ODB_TAG_BUFFERED ? &scene->buffered : nullptr);
}

// Spatial::set_motion, spatial.rs:137-149
extern "C" int odb_spatial_set_motion(odb_scene* scene, odb_source src, const float position[3], const float velocity[3],
                                      int discontinuity) {
    ODB_TRY(scene_check(scene));
    if (!position || !velocity) return odb_fail(ODB_E_INVALID, "NULL argument");
    std::lock_guard<std::mutex> lk(scene->mu);
    uint32_t tag, slot; bool stale;
    SourceSet* set = set_of(scene, src, &tag);
    if (!set) return odb_fail(ODB_E_INVALID, "not a spatial source handle");
    ODB_TRY(set->lookup(src, tag, &slot, &stale));
    if (stale) return ODB_OK;  // the signal is gone; the reference's Sender writes into a slot nobody reads
    set->queue_motion(slot, position, velocity, discontinuity);
    return ODB_OK;
}
extern "C" int odb_spatial_set_motion_many(odb_scene* scene, uint32_t n, const odb_source* srcs, const float* positions,
                                           const float* velocities, const uint8_t* discontinuity) {
    ODB_TRY(scene_check(scene));
    if (n && (!srcs || !positions || !velocities)) return odb_fail(ODB_E_INVALID, "NULL argument");
    // Every update is an independent latest-value-wins write (swap.rs:36-47), so the batch need not be atomic: the
    // lock is taken per chunk, and a callback that starts meanwhile picks up what has been queued so far instead of
    // waiting for the whole batch (the rest takes effect one callback later, as it would with the reference's
    // per-source swap cells).
    const uint32_t CHUNK = 256;
    for (uint32_t i0 = 0; i0 < n; i0 += CHUNK) {
        odb_yield_to_audio(scene->audio_wants);
        std::lock_guard<std::mutex> lk(scene->mu);
        const uint32_t i1 = i0 + CHUNK < n ? i0 + CHUNK : n;
        for (uint32_t i = i0; i < i1; i++) {
            uint32_t tag, slot; bool stale;
            SourceSet* set = set_of(scene, srcs[i], &tag);
            if (!set) return odb_fail(ODB_E_INVALID, "srcs[%u] is not a spatial source handle", i);
            ODB_TRY(set->lookup(srcs[i], tag, &slot, &stale));
            if (stale) continue;
            set->queue_motion(slot, positions + 3 * (size_t)i, velocities + 3 * (size_t)i, discontinuity ? discontinuity[i] : 0);
        }
    }
    return ODB_OK;
}
// Spatial::is_finished, spatial.rs:154-156
extern "C" int odb_spatial_is_finished(odb_scene* scene, odb_source src, int* out) {
    ODB_TRY(scene_check(scene));
    if (!out) return odb_fail(ODB_E_INVALID, "NULL argument");
    std::lock_guard<std::mutex> lk(scene->mu);
    uint32_t tag, slot; bool stale;
    SourceSet* set = set_of(scene, src, &tag);
    if (!set) return odb_fail(ODB_E_INVALID, "not a spatial source handle");
    ODB_TRY(set->lookup(src, tag, &slot, &stale));
    *out = stale ? 1 : 0;
    return ODB_OK;
}
extern "C" int odb_scene_len(odb_scene* scene, int buffered, uint64_t* out) {
    ODB_TRY(scene_check(scene));
    if (!out) return odb_fail(ODB_E_INVALID, "NULL argument");
    std::lock_guard<std::mutex> lk(scene->mu);
    *out = buffered ? scene->buffered.order.size() : scene->seek.order.size();
    return ODB_OK;
}

// <SpatialScene as Signal>::sample, spatial.rs:376-471. Leaves the mixed tile in `dev_out` (device memory, or
// `dev_out` when given) on the context's stream.
// Grows a device buffer that kernels on either stream may still be using: both streams are drained first.
template <class T>
static int ensure_idle(odb_scene* scene, DevBuf<T>& buf, size_t n) {
    if (n <= buf.cap) return ODB_OK;
    ODB_CUDA(cudaStreamSynchronize(scene->wst));
    ODB_CUDA(cudaStreamSynchronize(scene->ctx->stream));
    return buf.ensure(n, scene->ctx->stream, false);
}

// What odb_scene_sample_exchange asks of a callback: push the tile as the exchange's next sequence number and, once
// more than `lag` pushes are outstanding, pull the oldest one into dev_out with `epilogue`.
struct SceneExchange {
    odb_exchange* ex;
    int lag, epilogue;
    int written;  // out: dev_out received a summed tile
};

static int scene_sample_impl(odb_scene* scene, float interval, float* dev_out, uint32_t n_frames, bool as_i16 = false,
                             bool host_flag = false, SceneExchange* xr = nullptr) {
    odb_ctx* ctx = scene->ctx;
    cudaStream_t st = ctx->stream, wst = scene->pipelined ? scene->wst : ctx->stream;
    scene->flag_armed = false;
    scene->count_by_kernel = false;
    ODB_CUDA(cudaSetDevice(ctx->device));
    uint32_t launches = 0;
    const int p = (int)(scene->callback_no & 1);
    scene->callback_no++;
    // ---- stage 1 on `wst`: everything that is per source and O(1) per chunk -------------------------------------
    // this parity's job / counter buffers are free once the mix of two callbacks ago has finished
    if (scene->pipelined) ODB_CUDA(cudaStreamWaitEvent(wst, scene->ev_mix[p], 0));
    OdbCallback cb;
    {
        const double tl0 = g_trace ? now_us() : 0.0;
        AudioLock lk(scene->mu, scene->audio_wants);
        const double tl1 = g_trace ? now_us() : 0.0;
        cb.prev_rot = scene->rot_received;                                 // spatial.rs:382-386
        if (lk.held()) {  // otherwise: a control call is in progress; what it queues is applied by the next callback
            // removals reported by earlier callbacks whose read-back has landed (never waits for the device)
            ODB_TRY(scene->seek.fold_removed(ctx, wst, false, nullptr));
            ODB_TRY(scene->buffered.fold_removed(ctx, wst, false, nullptr));
            ODB_TRY(scene->buffered.apply(ctx, wst, &launches));            // set.update(), spatial.rs:379
            ODB_TRY(scene->seek.apply(ctx, wst, &launches));                // set.update(), spatial.rs:437
            if (scene->rot_fresh) { scene->rot_received = scene->rot_pending; scene->rot_fresh = false; }
        }
        if (g_trace && scene->callback_no > 4) { scene->tr_lock += tl1 - tl0; scene->tr_apply += now_us() - tl1; }
        cb.rot = scene->rot_received;
    }
    double ts = g_trace ? now_us() : 0.0;
    auto seg = [&](int i) { if (g_trace && scene->callback_no > 4) { const double t = now_us(); scene->tr_seg[i] += t - ts; ts = t; } };
    cb.interval = interval;
    cb.n_frames = (int)n_frames;
    cb.elapsed = interval * (float)n_frames;                               // spatial.rs:394
    cb.n_tiles = (int)((n_frames + ODB_TILE_FRAMES - 1) / ODB_TILE_FRAMES);
    cb.n_sources = (int)scene->seek.order.size();
    cb.force_general = scene->variant == 1;
    const int ns = cb.n_sources, nt = cb.n_tiles;
    const int nb = (int)scene->buffered.order.size();
    // The one-launch kernel and the seek set's walk loop over any number of 1024-frame tiles (spatial.rs:456 takes any
    // out.len()); the buffered set's walk kernel and round 1's path keep per-tile state for four tiles.
    if (n_frames > ODB_MAX_FRAMES && (nb > 0 || scene->legacy))
        return odb_fail(ODB_E_UNSUPPORTED, "n_frames %u exceeds the %d frames one callback may render with buffered sources", n_frames,
                        ODB_MAX_FRAMES);

    for (int q = 0; q < 2; q++) {  // job counters: zeroed at creation, afterwards by k_reduce_tiles of the previous callback
        if (!scene->d_counters[q].p) {
            ODB_TRY(ensure_idle(scene, scene->d_counters[q], ODB_CNT_WORDS));
            ODB_CUDA(cudaMemsetAsync(scene->d_counters[q].p, 0, ODB_CNT_WORDS * sizeof(uint32_t), wst));
        }
    }
    uint32_t* counters = scene->d_counters[p].p;
    if (scene->pipelined) {  // the next walk may overlap this callback's reduce: reset on the walk stream instead
        ODB_CUDA(cudaMemsetAsync(counters, 0, ODB_CNT_WORDS * sizeof(uint32_t), wst));
    } else if (nt == 0) {    // no reduce kernel will run for this callback
        ODB_CUDA(cudaMemsetAsync(scene->d_counters[0].p, 0, ODB_CNT_WORDS * sizeof(uint32_t), wst));
        ODB_CUDA(cudaMemsetAsync(scene->d_counters[1].p, 0, ODB_CNT_WORDS * sizeof(uint32_t), wst));
    }
    // One OdbJob array serves both sets: per tile the seek sources' records first, then the buffered sources'
    // (whose "PCM" is their delay ring); the staged kernel mixes both in one launch.
    const int ntot = ns + nb;
    cb.job_stride = ntot;
    cb.job_offset = 0;
    if (ntot > 0) ODB_TRY(ensure_idle(scene, scene->d_jobs[p], (size_t)ntot * (nt > 0 ? nt : 1)));
    if (nb > 0) {  // buffered set first (spatial.rs:395-433)
        OdbCallback cbb = cb;
        cbb.n_sources = nb;
        cbb.job_offset = ns;
        ODB_TRY(ensure_idle(scene, scene->d_ring_jobs[p], (size_t)nb * (nt > 0 ? nt : 1)));
        ODB_TRY(ensure_idle(scene, scene->d_ring_writes[p], (size_t)nb));
        ODB_TRY(ensure_idle(scene, scene->d_ring_list[p], (size_t)nb * (nt > 0 ? nt : 1)));
        odb_launch_walk_buffered(scene->buffered.d_src.p, scene->buffered.d_order.p, scene->d_ring_jobs[p].p, scene->d_jobs[p].p,
                                 scene->d_ring_writes[p].p, scene->buffered.d_removed.p, (int)scene->buffered.removed_cap,
                                 counters, scene->d_ring_list[p].p, cbb, wst);
        launches++;
    }
    const bool fused = nb == 0 && ns > 0 && nt > 0 && !scene->legacy;
    if (fused && !scene->d_counters_ring.p) {
        ODB_TRY(ensure_idle(scene, scene->d_counters_ring, 4 * ODB_CNT_WORDS));
        ODB_CUDA(cudaMemsetAsync(scene->d_counters_ring.p, 0, 4 * ODB_CNT_WORDS * sizeof(uint32_t), wst));
    }
    OdbJob* fused_jobs = nullptr;
    if (fused) {
        // the one-launch callback: counters from a ring of four sets, job records from a ring of three, no wait in
        // the walk (odb_walk.cuh, odb_scene_mix.cu)
        const unsigned long long k = scene->fused_seq;
        counters = scene->d_counters_ring.p + (k & 3) * ODB_CNT_WORDS;
        ODB_TRY(ensure_idle(scene, scene->d_jobs_ring[k % 3], (size_t)ns * nt));
        fused_jobs = scene->d_jobs_ring[k % 3].p;
        if (!scene->d_sync.p) {
            ODB_TRY(ensure_idle(scene, scene->d_sync, 8));
            ODB_CUDA(cudaMemsetAsync(scene->d_sync.p, 0, 8 * sizeof(unsigned long long), wst));
        }
        scene->walked_total += (unsigned long long)odb_launch_walk_seek(
            scene->seek.d_src.p, scene->seek.d_order.p, fused_jobs, scene->seek.d_removed.p, (int)scene->seek.removed_cap, counters,
            scene->d_counters_ring.p + ((k + 2) & 3) * ODB_CNT_WORDS, scene->d_sync.p + 5, cb, wst);
        launches++;
    } else if (ns > 0) {
        odb_launch_walk_seek(scene->seek.d_src.p, scene->seek.d_order.p, scene->d_jobs[p].p, scene->seek.d_removed.p,
                             (int)scene->seek.removed_cap, counters, nullptr, nullptr, cb, wst);
        launches++;
    }
    scene->last_counters = counters;
    seg(0);
    if (scene->pipelined) {
        ODB_CUDA(cudaEventRecord(scene->ev_walk[p], wst));
        // ---- stage 2 on the context's stream: O(sources x frames) -----------------------------------------------------
        ODB_CUDA(cudaStreamWaitEvent(st, scene->ev_walk[p], 0));
    }
    const bool use_fast = scene->variant != 1;
    // ---- the one-launch callback: staged mix + literal tail + grid reduce + epilogue (seek set only) ----------------
    if (fused) {
        int n_ctas = 1, batch = 8;
        odb_scene_mix_shape(ns, ctx->sm_count, &batch, &n_ctas);
        ODB_TRY(ensure_idle(scene, scene->d_partials_fused[scene->fused_seq & 1], (size_t)nt * n_ctas * 2 * ODB_TILE_FRAMES));
        const int fp = (int)(scene->fused_seq & 1);  // parity of this launch among the one-launch kernels
        OdbSceneMixArgs a;
        memset(&a, 0, sizeof a);
        a.jobs = fused_jobs;
        a.n_sources = ns; a.n_tiles = nt; a.n_frames = (int)n_frames;
        a.batch = batch;
        a.epilogue = scene->epilogue | (as_i16 ? ODB_EPILOGUE_I16_BIT : 0);
        a.partials = scene->d_partials_fused[fp].p;
        a.out = dev_out;
        a.arrive = scene->d_sync.p + fp;
        a.arrive_base = scene->arrive_total[fp];
        scene->arrive_total[fp] += (unsigned long long)nt * (unsigned long long)n_ctas;
        a.done = scene->d_sync.p + 2 + fp;
        a.done_base = scene->done_total[fp];
        scene->done_total[fp] += (unsigned long long)n_ctas;
        a.walked = scene->d_sync.p + 5;
        a.walked_target = scene->walked_total;
        a.completed = scene->d_sync.p + 4;
        a.my_seq = ++scene->fused_seq;
        if (xr) {  // multi-GPU: the reduce phase pushes (and pulls) over NVLink peer memory
            odb_exchange* ex = xr->ex;
            a.peers = ex->peers;
            a.xg = ex->geom();
            a.epilogue = xr->epilogue | (as_i16 ? ODB_EPILOGUE_I16_BIT : 0);
            a.push_seq = ++ex->seq;
            ex->pushed_floats[ex->seq % (uint32_t)ex->depth] = n_frames * 2;
            if (ex->seq - ex->pulled > (uint32_t)xr->lag) {
                a.pull_seq = ++ex->pulled;
                xr->written = 1;
            }
        }
        if (host_flag || xr) {  // the grid's last CTA finishes the callback from the raw sum in xtile
            ODB_TRY(ensure_idle(scene, scene->d_xtile[fp], (size_t)nt * 2 * ODB_TILE_FRAMES));
            a.xtile = scene->d_xtile[fp].p;
        }
        bool count_by_kernel = false;
        if (host_flag) {
            if (!scene->h_flag.p) {
                ODB_TRY(scene->h_flag.ensure(1));
                scene->h_flag.p[0] = 0ull;
            }
            a.host_flag = scene->h_flag.p;
            a.seq = ++scene->flag_seq;
            scene->flag_armed = true;
            if (scene->seek.d_removed.p && scene->seek.h_removed_count.p && !scene->seek.count_in_flight) {
                // the grid's last CTA also hands the host the removal-report count: no copy-engine operation behind the kernel
                a.removed_count = scene->seek.d_removed.p;
                a.removed_count_host = scene->seek.h_removed_count.p;
                count_by_kernel = true;
            }
        }
        scene->count_by_kernel = count_by_kernel;
        if (scene->profiling) ODB_CUDA(cudaEventRecord(scene->ev0, st));
        cudaError_t e = odb_launch_scene_mix(a, n_ctas, /*mode=*/scene->variant == 2 ? 1 : 0, st);
        if (e != cudaSuccess) return odb_fail(ODB_E_CUDA, "scene_mix launch failed: %s", cudaGetErrorString(e));
        if (scene->profiling) ODB_CUDA(cudaEventRecord(scene->ev1, st));
        launches++;
        seg(1);
        seg(2);
        if (scene->pipelined) ODB_CUDA(cudaEventRecord(scene->ev_mix[p], st));
        if (!count_by_kernel) ODB_TRY(scene->seek.post_callback(ctx, wst));
        seg(3);
        scene->last_launches = launches;
        ODB_CUDA(cudaGetLastError());
        return ODB_OK;
    }
    int n_ring = 0;
    if (nb > 0) {  // extend the delay rings (Ring::write), before anything reads them
        odb_launch_ring_write(scene->d_ring_writes[p].p, nb, st);
        launches++;
    }
    if (nt > 0) {
        int n_fast = 0, n_gen = 0;
        if (ntot > 0) {
            if (use_fast) {  // staged kernel for everything the walk kernels did not flag
                n_fast = odb_mix_fast_ctas(ntot, ctx->sm_count);
                ODB_TRY(ensure_idle(scene, scene->d_partials_fast, (size_t)nt * n_fast * 2 * ODB_TILE_FRAMES));
                if (scene->profiling) ODB_CUDA(cudaEventRecord(scene->ev0, st));
                cudaError_t e = odb_launch_mix_fast(scene->d_jobs[p].p, ntot, nt, scene->d_partials_fast.p, n_fast,
                                                    /*mode=*/scene->variant == 2 ? 1 : 0, st);
                if (e != cudaSuccess) return odb_fail(ODB_E_CUDA, "mix_fast launch failed: %s", cudaGetErrorString(e));
                if (scene->profiling) ODB_CUDA(cudaEventRecord(scene->ev1, st));
                launches++;
                seg(1);
            }
            if (ns > 0) {
                // literal kernel for the flagged rest of the seek set (exits at once when the walk kernel flagged nothing)
                n_gen = odb_mix_general_ctas(use_fast ? (ns < 2048 ? ns : 2048) : ns, ctx->sm_count);
                ODB_TRY(ensure_idle(scene, scene->d_partials, (size_t)nt * n_gen * 2 * ODB_TILE_FRAMES));
                cudaError_t e = odb_launch_mix_general(scene->d_jobs[p].p, ntot, nt, scene->d_partials.p, n_gen,
                                                       /*only_flagged=*/use_fast ? 1 : 0, counters, st);
                if (e != cudaSuccess) return odb_fail(ODB_E_CUDA, "mix_general launch failed: %s", cudaGetErrorString(e));
                launches++;
            }
            if (nb > 0) {
                // literal ring kernel for the buffered sources whose reads wrap around their ring this callback
                n_ring = odb_mix_ring_ctas(use_fast ? (nb < 2048 ? nb : 2048) : nb, ctx->sm_count);
                ODB_TRY(ensure_idle(scene, scene->d_partials_ring, (size_t)nt * n_ring * 2 * ODB_TILE_FRAMES));
                cudaError_t e = odb_launch_mix_ring(scene->d_ring_jobs[p].p, nb, nt, scene->d_partials_ring.p, n_ring,
                                                    /*only_flagged=*/use_fast ? 1 : 0, counters, scene->d_ring_list[p].p, st);
                if (e != cudaSuccess) return odb_fail(ODB_E_CUDA, "mix_ring launch failed: %s", cudaGetErrorString(e));
                launches++;
            }
        }
        odb_launch_reduce(scene->d_partials_fast.p, n_fast, scene->d_partials.p, n_gen, scene->d_partials_ring.p, n_ring,
                          counters, /*b_is_general=*/n_fast > 0 ? 1 : 0,
                          /*c_counter=*/(n_fast > 0 && n_ring > 0) ? ODB_CNT_RING_GENERAL : -1,
                          /*zero_counters=*/scene->pipelined ? nullptr : scene->d_counters[p ^ 1].p, dev_out,
                          (int)n_frames, nt, 2, scene->epilogue | (as_i16 ? ODB_EPILOGUE_I16_BIT : 0), st);
        launches++;
    }
    seg(2);
    if (xr) {  // the multi-kernel path rendered this rank's tile into dev_out: exchange it with the stand-alone kernels
        ODB_TRY(odb_exchange_push(xr->ex, dev_out, n_frames * 2, st));
        if (xr->ex->seq - xr->ex->pulled > (uint32_t)xr->lag) {
            ODB_TRY(odb_exchange_pull(xr->ex, dev_out, n_frames * 2, xr->epilogue, st));
            xr->written = 1;
        }
        launches += 2;
    }
    if (scene->pipelined) ODB_CUDA(cudaEventRecord(scene->ev_mix[p], st));
    // start the read-back of what walk_set removed; folded in by a later call without waiting (touches only
    // audio-side state: no control-plane lock)
    ODB_TRY(scene->seek.post_callback(ctx, wst));
    ODB_TRY(scene->buffered.post_callback(ctx, wst));
    seg(3);
    scene->last_launches = launches;
    ODB_CUDA(cudaGetLastError());
    return ODB_OK;
}

// Waits for the callback just queued: spins on the pinned word the grid's last CTA writes after the tile is stored
// (no driver call on the way back), falling back to a stream synchronisation when the callback took a path that
// does not publish the flag.
static int scene_wait(odb_scene* scene) {
    if (!scene->flag_armed) {
        ODB_CUDA(cudaStreamSynchronize(scene->ctx->stream));
        return ODB_OK;
    }
    volatile unsigned long long* flag = scene->h_flag.p;
    for (unsigned long long spins = 1; *flag < scene->flag_seq; spins++) {
        odb_cpu_pause();
        if ((spins & 0xFFFFF) == 0) {  // every ~million polls: a failed launch would never raise the flag
            cudaError_t q = cudaStreamQuery(scene->ctx->stream);
            if (q != cudaSuccess && q != cudaErrorNotReady) return odb_fail(ODB_E_CUDA, "callback failed: %s", cudaGetErrorString(q));
            if (q == cudaSuccess && *flag < scene->flag_seq) return odb_fail(ODB_E_CUDA, "callback finished without publishing its flag");
        }
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    return ODB_OK;
}
// Removals the callback reported (visible when sample returns, like the reference's).
static int scene_fold_after(odb_scene* scene) {
    odb_ctx* ctx = scene->ctx;
    cudaStream_t ws = scene->pipelined ? scene->wst : ctx->stream;
    // a set without members cannot have reported a removal: no read-back (a 4-byte copy + event wait is ~11 us)
    if (!scene->buffered.order.empty()) ODB_TRY(scene->buffered.fold_removed(ctx, ws, true, &scene->mu));
    if (scene->seek.order.empty()) return ODB_OK;
    if (scene->count_by_kernel) return scene->seek.fold_count(ctx, ws, scene->seek.h_removed_count.p[0], &scene->mu);
    return scene->seek.fold_removed(ctx, ws, true, &scene->mu);
}

extern "C" int odb_scene_sample(odb_scene* scene, float interval, float* out, uint32_t n_frames) {
    ODB_TRY(scene_check(scene));
    if (!out && n_frames) return odb_fail(ODB_E_INVALID, "out is NULL");
    odb_ctx* ctx = scene->ctx;
    ODB_CUDA(cudaSetDevice(ctx->device));
    size_t n = (size_t)n_frames * 2;
    const double t0 = g_trace ? now_us() : 0.0;
    // A caller buffer that is pinned or registered host memory is rendered into directly; anything else goes through
    // the scene's own pinned tile and one memcpy. Either way the kernel's reduce phase stores the tile straight
    // into host memory (unified addressing: no device-side staging tile, no copy-engine operation).
    // (a negative answer is looked up again every 256 callbacks: the caller may pin the buffer later, odb_pin_buffer)
    if (out != scene->pinned_probe || (!scene->pinned_dev && (scene->callback_no & 255) == 0)) {
        scene->pinned_probe = out;
        scene->pinned_dev = nullptr;
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, out) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer)
            scene->pinned_dev = (float*)at.devicePointer;
        else
            cudaGetLastError();
    }
    float* target = scene->pinned_dev;
    if (!target) {
        ODB_TRY(scene->h_out.ensure(n ? n : 2));
        target = scene->h_out.p;
    }
    ODB_TRY(scene_sample_impl(scene, interval, target, n_frames, false, /*host_flag=*/true));
    const double t1 = g_trace ? now_us() : 0.0;
    ODB_TRY(scene_wait(scene));
    const double t2 = g_trace ? now_us() : 0.0;
    if (n && target != out) {
        // the tile was just written by the GPU: every line misses to DRAM; asking for all of them at once before the
        // copy lets the misses overlap (a plain memcpy of these 8 KiB took 11 us)
        for (size_t b = 0; b < n * sizeof(float); b += 64) __builtin_prefetch((const char*)target + b, 0, 0);
        memcpy(out, target, n * sizeof(float));
    }
    int rc = scene_fold_after(scene);
    if (g_trace && scene->callback_no > 4) {
        scene->tr_enqueue += t1 - t0; if (t1 - t0 > scene->tr_enqueue_max) scene->tr_enqueue_max = t1 - t0;
        scene->tr_wait += t2 - t1; scene->tr_tail += now_us() - t2;
        scene->tr_calls++;
    }
    return rc;
}
// Offline render (examples/offline.rs:33-43): one callback quantised to 16-bit PCM on the device,
// `(sample * i16::MAX as f32) as i16`; half the bytes of the f32 tile cross to the host.
extern "C" int odb_scene_sample_i16(odb_scene* scene, float interval, int16_t* out, uint32_t n_frames) {
    ODB_TRY(scene_check(scene));
    if (!out && n_frames) return odb_fail(ODB_E_INVALID, "out is NULL");
    odb_ctx* ctx = scene->ctx;
    ODB_CUDA(cudaSetDevice(ctx->device));
    size_t n = (size_t)n_frames * 2;
    ODB_TRY(scene->h_out.ensure(n ? n : 2));
    ODB_TRY(scene_sample_impl(scene, interval, scene->h_out.p, n_frames, /*as_i16=*/true, /*host_flag=*/true));
    ODB_TRY(scene_wait(scene));
    if (n) memcpy(out, scene->h_out.p, n * sizeof(int16_t));
    return scene_fold_after(scene);
}
// oddio::run, lib.rs:90-93
extern "C" int odb_scene_run(odb_scene* scene, uint32_t sample_rate, float* out, uint32_t n_frames) {
    float interval = 1.0f / (float)sample_rate;
    return odb_scene_sample(scene, interval, out, n_frames);
}
extern "C" int odb_scene_sample_device(odb_scene* scene, float interval, void* dev_out, uint32_t n_frames) {
    ODB_TRY(scene_check(scene));
    if (!dev_out && n_frames) return odb_fail(ODB_E_INVALID, "dev_out is NULL");
    return scene_sample_impl(scene, interval, (float*)dev_out, n_frames);
}
// One callback of a source-sharded scene (one rank of several): mixes this rank's shard and exchanges the tile with
// the other ranks from inside the callback kernel - its reduce phase stores the rank's sum into every rank's inbox
// over NVLink and, once more than `lag` exchanges are outstanding, sums the oldest one over the ranks (rank order,
// bit-identical everywhere), applies `epilogue` and leaves it in dev_out. lag = 0: this callback's own sum (live
// playback; every rank waits for the slowest one inside the kernel); lag >= 1: a pipelined renderer receives callback
// k - lag while callback k is mixed, and collects the last `lag` tiles with odb_exchange_pull.
extern "C" int odb_scene_sample_exchange(odb_scene* scene, odb_exchange* ex, float interval, void* dev_out, uint32_t n_frames,
                                         int lag, int epilogue, int* out_written) {
    ODB_TRY(scene_check(scene));
    if (!ex || ex->kind != ODB_KIND_EXCHANGE) return odb_fail(ODB_E_INVALID, "not an exchange handle");
    if (!dev_out || !n_frames) return odb_fail(ODB_E_INVALID, "dev_out is NULL or n_frames is 0");
    if (!ex->connected) return odb_fail(ODB_E_INVALID, "exchange is not connected to its peers yet");
    if (ex->ctx != scene->ctx) return odb_fail(ODB_E_INVALID, "scene and exchange belong to different contexts");
    if (epilogue < 0 || epilogue > 2) return odb_fail(ODB_E_INVALID, "unknown epilogue %d", epilogue);
    if (lag < 0 || lag >= ex->depth) return odb_fail(ODB_E_INVALID, "lag %d: 0..depth-1 (%d) exchanges may await their pull", lag, ex->depth - 1);
    if ((size_t)n_frames * 2 > ex->cap) return odb_fail(ODB_E_INVALID, "%u frames exceed the exchange's capacity of %u floats", n_frames, ex->cap);
    if (((uintptr_t)dev_out & 15u) != 0) return odb_fail(ODB_E_INVALID, "dev_out must be 16-byte aligned");
    if (scene->epilogue != ODB_EPILOGUE_NONE) return odb_fail(ODB_E_INVALID, "a sharded scene runs with ODB_EPILOGUE_NONE; the epilogue acts on the sum");
    if (ex->seq - ex->pulled >= (uint32_t)ex->depth)
        return odb_fail(ODB_E_INVALID, "%d exchanges are already pushed and not pulled (the depth given at creation)", ex->depth);
    SceneExchange xr{ex, lag, epilogue, 0};
    int rc = scene_sample_impl(scene, interval, (float*)dev_out, n_frames, false, false, &xr);
    if (out_written) *out_written = xr.written;
    return rc;
}

// ---- per-source controls / read-backs ---------------------------------------------------------------
static int owner_set(void* owner, odb_source src, odb_ctx** ctx, std::mutex** mu, SourceSet** set, uint32_t* tag);

static int read_source(void* owner, odb_source src, OdbSource* out, bool* stale, SlotHost* shost) {
    odb_ctx* ctx; std::mutex* mu; SourceSet* set; uint32_t tag, slot;
    ODB_TRY(owner_set(owner, src, &ctx, &mu, &set, &tag));
    const OdbSource* dev = nullptr;
    cudaStream_t rs = ctx->stream;
    {   // resolve the handle under the control-plane lock ...
        std::lock_guard<std::mutex> lk(*mu);
        ODB_TRY(set->lookup(src, tag, &slot, stale));
        if (shost) *shost = set->slots[slot];
        if (*stale) return ODB_OK;
        for (size_t i = 0; i < set->ins_slot.size(); i++)  // queued but not yet applied: answer from the queue
            if (set->ins_slot[i] == slot) { *out = set->ins_src[i]; return ODB_OK; }
        dev = set->d_src.p + slot;
        // the source table is written on the scene's walk stream (on the context's stream for a mixer)
        if (*(uint32_t*)owner == ODB_KIND_SCENE && ((odb_scene*)owner)->pipelined) rs = ((odb_scene*)owner)->wst;
    }
    // ... and read the record without it: a poll from the game thread must not hold the lock the audio thread's
    // callback wants for the length of an in-flight callback (the reference's FramesSignalControl reads are atomics).
    // The table is reallocated only when the set outgrows it, under grow_mu.
    ODB_CUDA(cudaSetDevice(ctx->device));
    std::lock_guard<std::mutex> gl(set->grow_mu);
    dev = set->d_src.p + slot;
    ODB_CUDA(cudaMemcpyAsync(out, dev, sizeof(OdbSource), cudaMemcpyDeviceToHost, rs));
    ODB_CUDA(cudaStreamSynchronize(rs));
    return ODB_OK;
}

extern "C" int odb_source_playback_position(void* owner, odb_source src, double* out_seconds) {
    if (!out_seconds) return odb_fail(ODB_E_INVALID, "NULL argument");
    OdbSource s; bool stale; SlotHost sh;
    ODB_TRY(read_source(owner, src, &s, &stale, &sh));
    if (stale) return odb_fail(ODB_E_INVALID, "source no longer exists");
    *out_seconds = (double)s.sample_t / s.rate;  // frames.rs:238-240
    return ODB_OK;
}
extern "C" int odb_source_frames_is_finished(void* owner, odb_source src, int* out) {
    if (!out) return odb_fail(ODB_E_INVALID, "NULL argument");
    OdbSource s; bool stale; SlotHost sh;
    ODB_TRY(read_source(owner, src, &s, &stale, &sh));
    if (stale) { *out = 1; return ODB_OK; }
    *out = (s.sample_t >= 0 && (unsigned long long)s.sample_t >= (unsigned long long)s.len) ? 1 : 0;  // frames.rs:244-247
    return ODB_OK;
}
extern "C" int odb_source_cursor(void* owner, odb_source src, double* out_t, float* out_ring_write) {
    OdbSource s; bool stale; SlotHost sh;
    ODB_TRY(read_source(owner, src, &s, &stale, &sh));
    if (stale) return odb_fail(ODB_E_INVALID, "source no longer exists");
    if (out_t) *out_t = s.t;
    if (out_ring_write) *out_ring_write = s.ring_write;
    return ODB_OK;
}

// mixer side lives in odb_mixer.cu
int odb_mixer_owner_set(void* owner, odb_source src, odb_ctx** ctx, std::mutex** mu, SourceSet** set, uint32_t* tag);
int odb_mixer_last_launches(void* owner, uint32_t* out);
int odb_mixer_set_variant(void* owner, int variant);

static int owner_set(void* owner, odb_source src, odb_ctx** ctx, std::mutex** mu, SourceSet** set, uint32_t* tag) {
    if (!owner) return odb_fail(ODB_E_INVALID, "owner is NULL");
    uint32_t kind = *(uint32_t*)owner;
    if (kind == ODB_KIND_SCENE) {
        odb_scene* sc = (odb_scene*)owner;
        *ctx = sc->ctx; *mu = &sc->mu;
        *set = set_of(sc, src, tag);
        if (!*set) return odb_fail(ODB_E_INVALID, "not a source of this scene");
        return ODB_OK;
    }
    if (kind == ODB_KIND_MIXER) return odb_mixer_owner_set(owner, src, ctx, mu, set, tag);
    return odb_fail(ODB_E_INVALID, "owner is neither a scene nor a mixer");
}

static int queue_param(void* owner, odb_source src, uint32_t what, float value, uint32_t need_flag, const char* name) {
    odb_ctx* ctx; std::mutex* mu; SourceSet* set; uint32_t tag, slot; bool stale;
    ODB_TRY(owner_set(owner, src, &ctx, &mu, &set, &tag));
    std::lock_guard<std::mutex> lk(*mu);
    ODB_TRY(set->lookup(src, tag, &slot, &stale));
    if (stale) return ODB_OK;
    if (need_flag && !(set->slots[slot].chain_flags & need_flag))
        return odb_fail(ODB_E_INVALID, "source has no %s in its chain", name);
    set->queue_param(slot, what, value);
    return ODB_OK;
}
extern "C" int odb_source_set_speed(void* owner, odb_source src, float factor) {
    return queue_param(owner, src, ODB_PARAM_SPEED, factor, ODB_CHAIN_SPEED, "Speed");
}
extern "C" int odb_source_set_amplitude_ratio(void* owner, odb_source src, float factor) {
    return queue_param(owner, src, ODB_PARAM_GAIN, factor, ODB_CHAIN_GAIN, "Gain");
}
extern "C" int odb_source_set_gain_db(void* owner, odb_source src, float db) {
    return queue_param(owner, src, ODB_PARAM_GAIN, powf(10.0f, db / 20.0f), ODB_CHAIN_GAIN, "Gain");  // gain.rs:143-145
}

extern "C" int odb_last_launch_count(void* owner, uint32_t* out) {
    if (!owner || !out) return odb_fail(ODB_E_INVALID, "NULL argument");
    uint32_t kind = *(uint32_t*)owner;
    if (kind == ODB_KIND_SCENE) { *out = ((odb_scene*)owner)->last_launches; return ODB_OK; }
    if (kind == ODB_KIND_MIXER) return odb_mixer_last_launches(owner, out);
    return odb_fail(ODB_E_INVALID, "owner is neither a scene nor a mixer");
}
extern "C" int odb_set_profiling(void* owner, int enabled) {
    if (!owner) return odb_fail(ODB_E_INVALID, "NULL argument");
    if (*(uint32_t*)owner != ODB_KIND_SCENE) return odb_fail(ODB_E_UNSUPPORTED, "profiling events: scene only");
    odb_scene* sc = (odb_scene*)owner;
    ODB_CUDA(cudaSetDevice(sc->ctx->device));
    if (enabled && !sc->ev0) {
        ODB_CUDA(cudaEventCreate(&sc->ev0));
        ODB_CUDA(cudaEventCreate(&sc->ev1));
    }
    sc->profiling = enabled != 0;
    return ODB_OK;
}
extern "C" int odb_last_mix_kernel_ms(void* owner, float* out_ms) {
    if (!owner || !out_ms) return odb_fail(ODB_E_INVALID, "NULL argument");
    if (*(uint32_t*)owner != ODB_KIND_SCENE) return odb_fail(ODB_E_UNSUPPORTED, "profiling events: scene only");
    odb_scene* sc = (odb_scene*)owner;
    if (!sc->profiling || !sc->ev0) return odb_fail(ODB_E_INVALID, "profiling is not enabled");
    ODB_CUDA(cudaSetDevice(sc->ctx->device));
    ODB_CUDA(cudaEventSynchronize(sc->ev1));
    ODB_CUDA(cudaEventElapsedTime(out_ms, sc->ev0, sc->ev1));
    return ODB_OK;
}
int odb_mixer_job_counters(void* owner, uint32_t out[4]);
extern "C" int odb_last_job_counters(void* owner, uint32_t out[4]) {
    if (!owner || !out) return odb_fail(ODB_E_INVALID, "NULL argument");
    uint32_t kind = *(uint32_t*)owner;
    if (kind == ODB_KIND_MIXER) return odb_mixer_job_counters(owner, out);
    if (kind != ODB_KIND_SCENE) return odb_fail(ODB_E_INVALID, "owner is neither a scene nor a mixer");
    odb_scene* sc = (odb_scene*)owner;
    for (int i = 0; i < 4; i++) out[i] = 0;
    if (sc->callback_no == 0 || !sc->last_counters) return ODB_OK;
    ODB_CUDA(cudaSetDevice(sc->ctx->device));
    ODB_CUDA(cudaStreamSynchronize(sc->ctx->stream));
    ODB_CUDA(cudaMemcpyAsync(out, sc->last_counters, ODB_CNT_WORDS * sizeof(uint32_t), cudaMemcpyDeviceToHost, sc->ctx->stream));
    ODB_CUDA(cudaStreamSynchronize(sc->ctx->stream));
    return ODB_OK;
}
extern "C" int odb_set_kernel_variant(void* owner, int variant) {
    if (!owner) return odb_fail(ODB_E_INVALID, "NULL argument");
    uint32_t kind = *(uint32_t*)owner;
    if (kind == ODB_KIND_SCENE) {
        odb_scene* sc = (odb_scene*)owner;
        cudaStreamSynchronize(sc->wst);
        cudaStreamSynchronize(sc->ctx->stream);
        sc->variant = variant & 0xFF;
        sc->pipelined = (variant & 0x100) != 0;
        sc->legacy = (variant & 0x200) != 0;
        return ODB_OK;
    }
    if (kind == ODB_KIND_MIXER) return odb_mixer_set_variant(owner, variant);
    return odb_fail(ODB_E_INVALID, "owner is neither a scene nor a mixer");
}
