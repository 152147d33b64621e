// Context, PCM arena, Frames handles and the SourceSet control plane (host side of the C ABI).
#include "odb_host.h"

#include <algorithm>

static thread_local std::string g_err;
std::string& odb_err() { return g_err; }
int odb_fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

extern "C" const char* odb_last_error(void) { return g_err.c_str(); }
extern "C" uint32_t odb_abi_version(void) { return 1; }

// ---- context -----------------------------------------------------------------------------------
static int ctx_create(int cuda_device, bool own_stream, cudaStream_t stream, odb_ctx** out) {
    if (!out) return odb_fail(ODB_E_INVALID, "odb_ctx_create: out is NULL");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return odb_fail(ODB_E_CUDA, "no CUDA device available (%s); oddio_b200 has no CPU fallback",
                        e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    if (cuda_device < 0 || cuda_device >= count) return odb_fail(ODB_E_INVALID, "cuda_device %d out of range", cuda_device);
    ODB_CUDA(cudaSetDevice(cuda_device));
    odb_ctx* c = new odb_ctx();
    c->device = cuda_device;
    cudaDeviceProp prop;
    ODB_CUDA(cudaGetDeviceProperties(&prop, cuda_device));
    c->sm_count = prop.multiProcessorCount;
    c->own_stream = own_stream;
    if (own_stream) ODB_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    else c->stream = stream;
    *out = c;
    return ODB_OK;
}
extern "C" int odb_ctx_create(int cuda_device, odb_ctx** out) { return ctx_create(cuda_device, true, nullptr, out); }
extern "C" int odb_ctx_create_on_stream(int cuda_device, void* cuda_stream, odb_ctx** out) {
    return ctx_create(cuda_device, false, (cudaStream_t)cuda_stream, out);
}
extern "C" int odb_ctx_destroy(odb_ctx* ctx) {
    if (!ctx) return ODB_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (char* p : ctx->dead_blocks) cudaFree(p);
    for (auto& b : ctx->blocks)
        if (b.base) cudaFree(b.base);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return ODB_OK;
}
extern "C" int odb_ctx_synchronize(odb_ctx* ctx) {
    if (!ctx) return odb_fail(ODB_E_INVALID, "ctx is NULL");
    ODB_CUDA(cudaSetDevice(ctx->device));
    ODB_CUDA(cudaStreamSynchronize(ctx->stream));
    return ODB_OK;
}
extern "C" int odb_pin_buffer(odb_ctx* ctx, void* host_ptr, uint64_t bytes) {
    if (!ctx || !host_ptr || !bytes) return odb_fail(ODB_E_INVALID, "NULL argument");
    ODB_CUDA(cudaSetDevice(ctx->device));
    ODB_CUDA(cudaHostRegister(host_ptr, (size_t)bytes, cudaHostRegisterPortable | cudaHostRegisterMapped));
    return ODB_OK;
}
extern "C" int odb_unpin_buffer(odb_ctx* ctx, void* host_ptr) {
    if (!ctx || !host_ptr) return odb_fail(ODB_E_INVALID, "NULL argument");
    ODB_CUDA(cudaSetDevice(ctx->device));
    ODB_CUDA(cudaHostUnregister(host_ptr));
    return ODB_OK;
}
extern "C" int odb_ctx_stream(odb_ctx* ctx, void** out_stream) {
    if (!ctx || !out_stream) return odb_fail(ODB_E_INVALID, "NULL argument");
    *out_stream = (void*)ctx->stream;
    return ODB_OK;
}

// ---- arena: large blocks, bump allocation, freed when every Frames in a block is gone -----------
int odb_ctx::arena_alloc(size_t bytes, float** out, int* block) {
    bytes = (bytes + 255) & ~(size_t)255;
    int bi = -1;
    if (!blocks.empty()) {
        ArenaBlock& b = blocks.back();
        if (b.base && b.size - b.used >= bytes) bi = (int)blocks.size() - 1;
    }
    if (bi < 0) {
        ArenaBlock nb;
        nb.size = std::max(bytes, (size_t)256 << 20);
        ODB_CUDA(cudaMalloc((void**)&nb.base, nb.size));
        blocks.push_back(nb);
        bi = (int)blocks.size() - 1;
    }
    ArenaBlock& b = blocks[bi];
    *out = (float*)(b.base + b.used);
    b.used += bytes;
    b.live++;
    *block = bi;
    return ODB_OK;
}
void odb_ctx::arena_unref(int block, bool defer) {
    if (block < 0 || block >= (int)blocks.size()) return;
    ArenaBlock& b = blocks[block];
    if (--b.live == 0 && block != (int)blocks.size() - 1) {
        if (defer) {
            dead_blocks.push_back(b.base);
        } else {
            cudaStreamSynchronize(stream);
            cudaFree(b.base);
        }
        b.base = nullptr;
    } else if (b.live == 0) {
        b.used = 0;  // the open block can be reused from the start
    }
}
void odb_ctx::arena_collect() {  // takes `mu` only to detach the list: the synchronisation and the frees run without it
    std::vector<char*> dead;
    {
        std::lock_guard<std::mutex> lk(mu);
        dead.swap(dead_blocks);
    }
    if (dead.empty()) return;
    cudaStreamSynchronize(stream);
    for (char* p : dead) cudaFree(p);
}
int odb_ctx::ring_alloc(size_t bytes, float** out, int* block) {
    bytes = (bytes + 255) & ~(size_t)255;
    auto it = free_rings.find(bytes);
    if (it != free_rings.end() && !it->second.empty()) {
        *out = it->second.back().p;
        *block = it->second.back().block;
        it->second.pop_back();
        return ODB_OK;
    }
    return arena_alloc(bytes, out, block);
}
void odb_ctx::ring_release(float* p, size_t bytes, int block) {
    bytes = (bytes + 255) & ~(size_t)255;
    free_rings[bytes].push_back(FreeRing{p, block});  // keeps its arena reference: the block stays alive for the next tenant
}
int odb_ctx::frames_ref(odb_frames id, FramesRec* out) {
    std::lock_guard<std::mutex> lk(mu);
    auto it = frames.find(id);
    if (it == frames.end()) return odb_fail(ODB_E_INVALID, "unknown frames handle %llu", (unsigned long long)id);
    it->second.refs++;
    *out = it->second;
    return ODB_OK;
}
void odb_ctx::frames_unref(odb_frames id) {
    std::lock_guard<std::mutex> lk(mu);
    auto it = frames.find(id);
    if (it == frames.end()) return;
    if (--it->second.refs == 0) {
        arena_unref(it->second.block, /*defer=*/true);  // may run on the audio thread (a finished source's last reference)
        frames.erase(it);
    }
}

static int frames_new(odb_ctx* ctx, uint32_t rate, int channels, const void* samples, uint64_t n_frames, bool from_device,
                      odb_frames* out) {
    if (!ctx || !out) return odb_fail(ODB_E_INVALID, "NULL argument");
    if (channels != 1 && channels != 2) return odb_fail(ODB_E_UNSUPPORTED, "channels must be 1 or 2, got %d", channels);
    if (n_frames == 0 || n_frames > (1ull << 29)) return odb_fail(ODB_E_INVALID, "n_frames %llu out of range", (unsigned long long)n_frames);
    if (rate == 0) return odb_fail(ODB_E_INVALID, "rate must be nonzero");
    ODB_CUDA(cudaSetDevice(ctx->device));
    FramesRec rec;
    size_t elems = (size_t)n_frames * channels;
    float* base = nullptr;
    {
        ctx->arena_collect();  // blocks the audio thread found dead are freed here, on the control side
        std::lock_guard<std::mutex> lk(ctx->mu);
        ODB_TRY(ctx->arena_alloc((elems + 2 * ODB_PCM_PAD) * sizeof(float), &base, &rec.block));
    }
    ODB_CUDA(cudaMemsetAsync(base, 0, ODB_PCM_PAD * sizeof(float), ctx->stream));
    ODB_CUDA(cudaMemsetAsync(base + ODB_PCM_PAD + elems, 0, ODB_PCM_PAD * sizeof(float), ctx->stream));
    if (samples)
        ODB_CUDA(cudaMemcpyAsync(base + ODB_PCM_PAD, samples, elems * sizeof(float),
                                 from_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, ctx->stream));
    ODB_CUDA(cudaStreamSynchronize(ctx->stream));
    rec.dev = base + ODB_PCM_PAD;
    rec.n_frames = n_frames;
    rec.channels = channels;
    rec.rate = rate;
    rec.refs = 1;
    std::lock_guard<std::mutex> lk(ctx->mu);
    odb_frames id = ctx->next_frames_id++;
    ctx->frames[id] = rec;
    *out = id;
    return ODB_OK;
}
// examples/wav.rs:30-46: integer PCM is uploaded as it is (half the bytes of f32) and scaled to f32 on the device
extern "C" int odb_frames_from_i16(odb_ctx* ctx, uint32_t rate, int channels, const int16_t* samples, uint64_t n_frames,
                                   int bits_per_sample, odb_frames* out) {
    if (!samples) return odb_fail(ODB_E_INVALID, "samples is NULL");
    if (bits_per_sample < 2 || bits_per_sample > 16)
        return odb_fail(ODB_E_UNSUPPORTED, "bits_per_sample %d: 2..16 fit the 16-bit container", bits_per_sample);
    ODB_TRY(frames_new(ctx, rate, channels, nullptr, n_frames, false, out));
    FramesRec rec;
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        rec = ctx->frames[*out];
    }
    const size_t elems = (size_t)n_frames * channels;
    short* tmp = nullptr;
    cudaError_t e = cudaMalloc((void**)&tmp, elems * sizeof(short));
    if (e != cudaSuccess) {
        ctx->frames_unref(*out);
        return odb_fail(ODB_E_NOMEM, "cudaMalloc of the %zu-byte upload buffer failed: %s", elems * sizeof(short), cudaGetErrorString(e));
    }
    e = cudaMemcpyAsync(tmp, samples, elems * sizeof(short), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
        odb_launch_convert_i16(tmp, rec.dev, elems, (float)((1u << (bits_per_sample - 1)) - 1u), ctx->stream);
        e = cudaStreamSynchronize(ctx->stream);
    }
    cudaFree(tmp);
    if (e != cudaSuccess) {
        ctx->frames_unref(*out);
        return odb_fail(ODB_E_CUDA, "integer PCM upload failed: %s", cudaGetErrorString(e));
    }
    return ODB_OK;
}
extern "C" int odb_frames_from_slice(odb_ctx* ctx, uint32_t rate, int channels, const float* samples, uint64_t n_frames,
                                     odb_frames* out) {
    if (!samples) return odb_fail(ODB_E_INVALID, "samples is NULL");
    return frames_new(ctx, rate, channels, samples, n_frames, false, out);
}
extern "C" int odb_frames_from_device(odb_ctx* ctx, uint32_t rate, int channels, const void* dev_samples, uint64_t n_frames,
                                      odb_frames* out) {
    if (!dev_samples) return odb_fail(ODB_E_INVALID, "dev_samples is NULL");
    return frames_new(ctx, rate, channels, dev_samples, n_frames, true, out);
}
extern "C" int odb_frames_release(odb_ctx* ctx, odb_frames frames) {
    if (!ctx) return odb_fail(ODB_E_INVALID, "ctx is NULL");
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        if (ctx->frames.find(frames) == ctx->frames.end())
            return odb_fail(ODB_E_INVALID, "unknown frames handle %llu", (unsigned long long)frames);
    }
    ctx->frames_unref(frames);
    return ODB_OK;
}

// ---- SourceSet -----------------------------------------------------------------------------------
uint32_t SourceSet::alloc_slot() {
    uint32_t s;
    if (!free_slots.empty()) {
        s = free_slots.back();
        free_slots.pop_back();
    } else {
        s = (uint32_t)slots.size();
        slots.emplace_back();
    }
    slots[s].in_use = true;
    slots[s].stopped = false;
    return s;
}
odb_source SourceSet::handle_of(uint32_t slot, uint32_t tag) const {
    return ((uint64_t)slots[slot].gen << 40) | ((uint64_t)(tag & 0xFF) << 32) | (uint64_t)slot;
}
int SourceSet::lookup(odb_source h, uint32_t tag, uint32_t* slot, bool* stale) const {
    uint32_t s = (uint32_t)(h & 0xFFFFFFFFu), t = (uint32_t)((h >> 32) & 0xFF), g = (uint32_t)(h >> 40);
    if (t != tag || s >= slots.size() || g == 0 || g > slots[s].gen)
        return odb_fail(ODB_E_INVALID, "foreign source handle %llx", (unsigned long long)h);
    *slot = s;
    *stale = (g != slots[s].gen) || !slots[s].in_use;
    return ODB_OK;
}
void SourceSet::queue_motion(uint32_t slot, const float* pos, const float* vel, int disc) {
    OdbMotionMsg m;
    m.slot = slot;
    for (int k = 0; k < 3; k++) { m.pos[k] = pos[k]; m.vel[k] = vel[k]; }
    m.discontinuity = disc ? 1u : 0u;
    SlotHost& sh = slots[slot];
    PinBuf<OdbMotionMsg>& hb = h_mot[mot_buf];
    if (sh.motion_gen == mot_gen) {  // latest value wins (swap.rs:36-47)
        if (sh.motion_idx >= 0) hb.p[sh.motion_idx] = m;
        else motions[(size_t)(-sh.motion_idx - 2)] = m;  // spilled message: index encoded as -(i + 2)
        return;
    }
    sh.motion_gen = mot_gen;
    if (mot_n < hb.cap) { sh.motion_idx = (int)mot_n; hb.p[mot_n++] = m; }
    else { sh.motion_idx = -(int)motions.size() - 2; motions.push_back(m); }
}
void SourceSet::queue_param(uint32_t slot, uint32_t what, float value) {
    OdbParamMsg m = {slot, what, value, 0u};
    SlotHost& sh = slots[slot];
    int* idx = what == ODB_PARAM_SPEED ? &sh.speed_idx : (what == ODB_PARAM_GAIN ? &sh.gain_idx : nullptr);
    if (idx && *idx >= 0) { params[*idx] = m; return; }
    if (idx) *idx = (int)params.size();
    params.push_back(m);
}

int SourceSet::apply(odb_ctx* ctx, cudaStream_t st, uint32_t* launches) {
    // inserts: set.rs:159-165 — appended to the table in send order
    size_t ni = ins_slot.size();
    if (ni) {
        if (slots.size() > d_src.cap) {
            std::lock_guard<std::mutex> gl(grow_mu);
            ODB_TRY(d_src.ensure(slots.size(), st, true));
        }
        ODB_TRY(h_stage_src.ensure(ni));
        ODB_TRY(h_stage_slot.ensure(ni));
        ODB_TRY(d_stage_src.ensure(ni, st, false));
        ODB_TRY(d_stage_slot.ensure(ni, st, false));
        memcpy(h_stage_src.p, ins_src.data(), ni * sizeof(OdbSource));
        memcpy(h_stage_slot.p, ins_slot.data(), ni * sizeof(uint32_t));
        ODB_CUDA(cudaMemcpyAsync(d_stage_src.p, h_stage_src.p, ni * sizeof(OdbSource), cudaMemcpyHostToDevice, st));
        ODB_CUDA(cudaMemcpyAsync(d_stage_slot.p, h_stage_slot.p, ni * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        odb_launch_scatter_sources(d_src.p, d_stage_src.p, d_stage_slot.p, (int)ni, st);
        (*launches)++;
        for (uint32_t s : ins_slot) {
            if (pos_of_slot.size() <= s) pos_of_slot.resize(slots.size(), -1);
            pos_of_slot[s] = (int)order.size();
            order.push_back(s);
        }
        order_dirty = true;
        ins_src.clear();
        ins_slot.clear();
    }
    size_t nm = mot_n + motions.size();
    if (nm) {
        ODB_TRY(d_motions.ensure(nm, st, false));
        const int b = mot_buf;
        if (mot_n && motions.empty()) {
            // The common case: the scatter kernel reads the pinned buffer the control side filled directly over PCIe
            // (unified addressing) - no copy-engine operation in front of the callback's kernels. The event behind
            // the kernel tells the control side when the buffer may be refilled.
            if (!ev_mot[b]) ODB_CUDA(cudaEventCreateWithFlags(&ev_mot[b], cudaEventDisableTiming));
            odb_launch_scatter_motion(d_src.p, h_mot[b].p, (int)mot_n, st);
            ODB_CUDA(cudaEventRecord(ev_mot[b], st));
            ev_mot_pending[b] = true;
        } else {
            if (mot_n) {  // the pinned buffer the control side filled goes to the device as it lies
                if (!ev_mot[b]) ODB_CUDA(cudaEventCreateWithFlags(&ev_mot[b], cudaEventDisableTiming));
                ODB_CUDA(cudaMemcpyAsync(d_motions.p, h_mot[b].p, mot_n * sizeof(OdbMotionMsg), cudaMemcpyHostToDevice, st));
                ODB_CUDA(cudaEventRecord(ev_mot[b], st));
                ev_mot_pending[b] = true;
            }
            if (!motions.empty()) {  // what did not fit (rare: sources played since the buffers were sized)
                ODB_TRY(h_motions.ensure(motions.size()));
                memcpy(h_motions.p, motions.data(), motions.size() * sizeof(OdbMotionMsg));
                ODB_CUDA(cudaMemcpyAsync(d_motions.p + mot_n, h_motions.p, motions.size() * sizeof(OdbMotionMsg), cudaMemcpyHostToDevice, st));
                motions.clear();
            }
            odb_launch_scatter_motion(d_src.p, d_motions.p, (int)nm, st);
        }
        (*launches)++;
        mot_n = 0;
        mot_gen++;  // forgets every slot's queued-message index at once
        if (b == mot_buf && ev_mot_pending[b]) mot_buf = b ^ 1;
    }
    {   // the buffer the control side fills next: its last copy must have left, and it holds one message per slot
        const int b = mot_buf;
        if (ev_mot_pending[b]) { ODB_CUDA(cudaEventSynchronize(ev_mot[b])); ev_mot_pending[b] = false; }
        if (mot_n == 0 && h_mot[b].cap < slots.size()) ODB_TRY(h_mot[b].ensure(slots.size() + slots.size() / 2 + 256));
    }
    size_t np = params.size();
    if (np) {
        ODB_TRY(h_params.ensure(np));
        ODB_TRY(d_params.ensure(np, st, false));
        memcpy(h_params.p, params.data(), np * sizeof(OdbParamMsg));
        ODB_CUDA(cudaMemcpyAsync(d_params.p, h_params.p, np * sizeof(OdbParamMsg), cudaMemcpyHostToDevice, st));
        odb_launch_scatter_params(d_src.p, d_params.p, (int)np, st);
        (*launches)++;
        for (auto& m : params)
            if (m.slot != 0xFFFFFFFFu) { slots[m.slot].speed_idx = -1; slots[m.slot].gain_idx = -1; }
        params.clear();
    }
    // removal report ring: every live source reports at most once, so a ring of >= order.size() never overflows
    if (removed_cap < order.size() + ins_slot.size() || !d_removed.p) {
        if (d_removed.p) ODB_TRY(fold_removed(ctx, st, true, nullptr));
        uint32_t ncap = 1024;
        while (ncap < 2 * (order.size() + ins_slot.size())) ncap *= 2;
        d_removed.release();
        ODB_TRY(d_removed.ensure((size_t)ncap + 1, st, false));
        ODB_CUDA(cudaMemsetAsync(d_removed.p, 0, sizeof(uint32_t), st));
        ODB_TRY(h_removed.ensure(ncap));
        ODB_TRY(h_removed_count.ensure(1));
        if (!ev_removed) ODB_CUDA(cudaEventCreateWithFlags(&ev_removed, cudaEventDisableTiming));
        removed_cap = ncap;
        removed_consumed = 0;
        count_in_flight = false;
    }
    if (order_dirty) {
        size_t n = order.size();
        if (n) {
            ODB_TRY(h_order.ensure(n));
            ODB_TRY(d_order.ensure(n, st, false));
            // the previous H2D copy out of h_order has completed: every sample() either
            // synchronises the stream or is followed by one before the next apply()
            memcpy(h_order.p, order.data(), n * sizeof(uint32_t));
            ODB_CUDA(cudaMemcpyAsync(d_order.p, h_order.p, n * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        }
        order_dirty = false;
    }
    return ODB_OK;
}

int SourceSet::post_callback(odb_ctx* ctx, cudaStream_t st) {
    (void)ctx;
    if (count_in_flight || !d_removed.p) return ODB_OK;  // the previous read-back has not been looked at yet
    ODB_CUDA(cudaMemcpyAsync(h_removed_count.p, d_removed.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    ODB_CUDA(cudaEventRecord(ev_removed, st));
    count_in_flight = true;
    return ODB_OK;
}

int SourceSet::fold_removed(odb_ctx* ctx, cudaStream_t st, bool wait, std::mutex* mu) {
    if (!d_removed.p) return ODB_OK;
    if (!count_in_flight) {
        if (!wait) return ODB_OK;
        ODB_TRY(post_callback(ctx, st));
    }
    if (wait) ODB_CUDA(cudaEventSynchronize(ev_removed));
    else {
        cudaError_t q = cudaEventQuery(ev_removed);
        if (q == cudaErrorNotReady) return ODB_OK;
        ODB_CUDA(q);
    }
    count_in_flight = false;
    return fold_count(ctx, st, h_removed_count.p[0], mu);
}

int SourceSet::fold_count(odb_ctx* ctx, cudaStream_t st, uint32_t count, std::mutex* mu) {
    uint32_t n_new = count - removed_consumed;
    if (n_new == 0) return ODB_OK;
    if (n_new > removed_cap) return odb_fail(ODB_E_INVALID, "internal: removal ring overflow (%u reports)", n_new);
    // from here on the membership (order, slots) changes: that is shared with the control side. The audio thread
    // only tries the lock (bounded): if a control call holds it, these removals are folded by the next callback.
    std::unique_lock<std::mutex> lk;
    if (mu) {
        lk = std::unique_lock<std::mutex>(*mu, std::defer_lock);
        bool got = false;
        for (int i = 0; i < 512 && !(got = lk.try_lock()); i++) odb_cpu_pause();
        if (!got) return ODB_OK;
    }
    // fetch the report entries (only happens on callbacks where sources actually finished)
    const uint32_t mask = removed_cap - 1, first = removed_consumed & mask;
    const uint32_t span1 = n_new < removed_cap - first ? n_new : removed_cap - first;
    ODB_CUDA(cudaMemcpyAsync(h_removed.p, d_removed.p + 1 + first, span1 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    if (n_new > span1)
        ODB_CUDA(cudaMemcpyAsync(h_removed.p + span1, d_removed.p + 1, (n_new - span1) * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    ODB_CUDA(cudaStreamSynchronize(st));
    removed_consumed = count;
    // The reference walks i from len-1 down to 0 and swap_removes as it goes (spatial.rs:204,259;
    // mixer.rs:100,104; set.rs:183-188): replaying the positions in descending order gives the same Vec.
    std::vector<int> pos;
    pos.reserve(n_new);
    for (uint32_t i = 0; i < n_new; i++) {
        uint32_t slot = h_removed.p[i];
        if (slot < pos_of_slot.size() && pos_of_slot[slot] >= 0) pos.push_back(pos_of_slot[slot]);
    }
    std::sort(pos.begin(), pos.end(), [](int a, int b) { return a > b; });
    for (int i : pos) {
        uint32_t slot = order[(size_t)i];
        uint32_t moved = order.back();
        order[(size_t)i] = moved;
        order.pop_back();
        pos_of_slot[moved] = i;
        pos_of_slot[slot] = -1;
        SlotHost& sh = slots[slot];
        sh.stopped = true;
        sh.in_use = false;
        sh.gen++;  // older handles now read as "finished" (see lookup)
        if (sh.motion_gen == mot_gen) {  // a queued set_motion for a source that is gone must not reach the slot's next tenant
            if (sh.motion_idx >= 0) h_mot[mot_buf].p[sh.motion_idx].slot = 0xFFFFFFFFu;
            else if (sh.motion_idx <= -2) motions[(size_t)(-sh.motion_idx - 2)].slot = 0xFFFFFFFFu;
            sh.motion_gen = 0;
        }
        if (sh.speed_idx >= 0 || sh.gain_idx >= 0 || sh.stop_requested) {  // ... nor may a queued set_speed / set_gain / stop
            for (auto& m : params)
                if (m.slot == slot) m.slot = 0xFFFFFFFFu;
        }
        sh.motion_idx = sh.speed_idx = sh.gain_idx = -1;
        if (sh.frames) ctx->frames_unref(sh.frames);
        sh.frames = 0;
        if (sh.ring_block >= 0) {  // the delay ring goes to the context's free list for the next play_buffered of that size
            std::lock_guard<std::mutex> lk(ctx->mu);
            ctx->ring_release(sh.ring_ptr, sh.ring_bytes, sh.ring_block);
            sh.ring_block = -1;
            sh.ring_ptr = nullptr;
        }
        free_slots.push_back(slot);
    }
    order_dirty = true;
    return ODB_OK;
}

void SourceSet::release_all(odb_ctx* ctx) {
    for (auto& sh : slots) {
        if (sh.in_use && sh.frames) ctx->frames_unref(sh.frames);
        if (sh.in_use && sh.ring_block >= 0) {
            std::lock_guard<std::mutex> lk(ctx->mu);
            ctx->ring_release(sh.ring_ptr, sh.ring_bytes, sh.ring_block);
        }
    }
    slots.clear();
    d_src.release(); d_order.release(); d_stage_src.release(); d_stage_slot.release();
    d_motions.release(); d_params.release(); d_removed.release();
    h_stage_src.release(); h_stage_slot.release(); h_motions.release(); h_params.release();
    h_order.release(); h_removed.release(); h_removed_count.release();
    for (int b = 0; b < 2; b++) {
        h_mot[b].release();
        if (ev_mot[b]) cudaEventDestroy(ev_mot[b]);
        ev_mot[b] = nullptr;
        ev_mot_pending[b] = false;
    }
    mot_n = 0;
    if (ev_removed) cudaEventDestroy(ev_removed);
    ev_removed = nullptr;
    removed_cap = 0;
}
