// odb_mixer_*: MixerControl<T> + Mixer<T> over a device-resident source set.
// Reference: src/mixer.rs (cited per function), src/speed.rs, src/gain.rs, src/tanh.rs.
#include "odb_host.h"

#define ODB_TAG_MIXED 3u
#define ODB_MIXER_MAX_TILES 1024

struct odb_mixer {
    uint32_t kind = ODB_KIND_MIXER;
    odb_ctx* ctx = nullptr;
    int channels = 2;
    std::mutex mu;
    std::atomic<int> audio_wants{0};  // the audio thread is waiting for `mu` (see AudioLock)
    SourceSet set;
    int epilogue = ODB_EPILOGUE_NONE;
    int variant = 0;
    uint32_t last_launches = 0;
    DevBuf<OdbMixJob> d_jobs;
    DevBuf<float> d_partials_unit, d_partials_gen, d_partials_res;
    DevBuf<uint32_t> d_counters;
    DevBuf<float> d_out;
    PinBuf<float> h_out;
};

static int mixer_check(odb_mixer* m) {
    if (!m || m->kind != ODB_KIND_MIXER) return odb_fail(ODB_E_INVALID, "not a mixer handle");
    return ODB_OK;
}

// Mixer::new, mixer.rs:70-81
extern "C" int odb_mixer_create(odb_ctx* ctx, int channels, odb_mixer** out) {
    if (!ctx || !out) return odb_fail(ODB_E_INVALID, "NULL argument");
    if (channels != 1 && channels != 2) return odb_fail(ODB_E_UNSUPPORTED, "channels must be 1 or 2");
    odb_mixer* m = new odb_mixer();
    m->ctx = ctx;
    m->channels = channels;
    *out = m;
    return ODB_OK;
}
extern "C" int odb_mixer_destroy(odb_mixer* mixer) {
    if (!mixer) return ODB_OK;
    ODB_TRY(mixer_check(mixer));
    cudaSetDevice(mixer->ctx->device);
    cudaStreamSynchronize(mixer->ctx->stream);
    mixer->set.release_all(mixer->ctx);
    mixer->d_jobs.release(); mixer->d_partials_unit.release(); mixer->d_partials_gen.release(); mixer->d_partials_res.release();
    mixer->d_counters.release(); mixer->d_out.release(); mixer->h_out.release();
    mixer->kind = 0;
    delete mixer;
    return ODB_OK;
}
extern "C" int odb_mixer_set_epilogue(odb_mixer* mixer, int epilogue) {
    ODB_TRY(mixer_check(mixer));
    if (epilogue < 0 || epilogue > 2) return odb_fail(ODB_E_INVALID, "unknown epilogue %d", epilogue);
    mixer->epilogue = epilogue;
    return ODB_OK;
}

// MixerControl::play, mixer.rs:18-26
extern "C" int odb_mixer_play(odb_mixer* mixer, const odb_chain* chain, odb_source* out) {
    ODB_TRY(mixer_check(mixer));
    if (!chain || !out) return odb_fail(ODB_E_INVALID, "NULL argument");
    OdbSource s;
    FramesRec rec;
    ODB_TRY(odb_make_source(mixer->ctx, chain, mixer->channels, &s, &rec));
    std::lock_guard<std::mutex> lk(mixer->mu);
    uint32_t slot = mixer->set.alloc_slot();
    SlotHost& sh = mixer->set.slots[slot];
    sh.frames = chain->frames; sh.chain_flags = chain->flags; sh.n_frames = rec.n_frames; sh.rate = (double)rec.rate;
    sh.stop_requested = false;
    mixer->set.ins_src.push_back(s);
    mixer->set.ins_slot.push_back(slot);
    *out = mixer->set.handle_of(slot, ODB_TAG_MIXED);
    return ODB_OK;
}

// Mixed::stop, mixer.rs:34-36
extern "C" int odb_mixed_stop(odb_mixer* mixer, odb_source src) {
    ODB_TRY(mixer_check(mixer));
    std::lock_guard<std::mutex> lk(mixer->mu);
    uint32_t slot; bool stale;
    ODB_TRY(mixer->set.lookup(src, ODB_TAG_MIXED, &slot, &stale));
    if (stale) return ODB_OK;
    mixer->set.slots[slot].stop_requested = true;
    mixer->set.queue_param(slot, ODB_PARAM_STOP, 0.0f);
    return ODB_OK;
}
// Mixed::is_stopped, mixer.rs:41-43: true once stop() was called or the mixer dropped the finished signal
extern "C" int odb_mixed_is_stopped(odb_mixer* mixer, odb_source src, int* out) {
    ODB_TRY(mixer_check(mixer));
    if (!out) return odb_fail(ODB_E_INVALID, "NULL argument");
    std::lock_guard<std::mutex> lk(mixer->mu);
    uint32_t slot; bool stale;
    ODB_TRY(mixer->set.lookup(src, ODB_TAG_MIXED, &slot, &stale));
    *out = (stale || mixer->set.slots[slot].stop_requested) ? 1 : 0;
    return ODB_OK;
}

// <Mixer<T> as Signal>::sample, mixer.rs:92-119
static int mixer_sample_impl(odb_mixer* mixer, float interval, float* dev_out, uint32_t n_frames, bool as_i16 = false) {
    odb_ctx* ctx = mixer->ctx;
    cudaStream_t st = ctx->stream;
    if (n_frames > ODB_MIXER_MAX_TILES * ODB_MIXER_CHUNK)
        return odb_fail(ODB_E_UNSUPPORTED, "n_frames %u exceeds the %d frames one callback may render", n_frames,
                        ODB_MIXER_MAX_TILES * ODB_MIXER_CHUNK);
    ODB_CUDA(cudaSetDevice(ctx->device));
    uint32_t launches = 0;
    {
        AudioLock lk(mixer->mu, mixer->audio_wants);
        if (lk.held()) {  // otherwise: a control call is in progress; what it queues is applied by the next callback
            ODB_TRY(mixer->set.fold_removed(ctx, st, false, nullptr));
            ODB_TRY(mixer->set.apply(ctx, st, &launches));  // set.update(), mixer.rs:94
        }
    }
    OdbCallback cb;
    memset(&cb, 0, sizeof cb);
    cb.interval = interval;
    cb.n_frames = (int)n_frames;
    cb.elapsed = interval * (float)n_frames;
    cb.n_tiles = (int)((n_frames + ODB_MIXER_CHUNK - 1) / ODB_MIXER_CHUNK);
    cb.n_sources = (int)mixer->set.order.size();
    cb.force_general = mixer->variant == 1;
    const int ns = cb.n_sources, nt = cb.n_tiles, ch = mixer->channels;
    const size_t tile_floats = (size_t)ODB_MIXER_CHUNK * ch;

    ODB_TRY(mixer->d_counters.ensure(ODB_CNT_WORDS, st, false));
    ODB_CUDA(cudaMemsetAsync(mixer->d_counters.p, 0, ODB_CNT_WORDS * sizeof(uint32_t), st));
    if (ns > 0) {
        ODB_TRY(mixer->d_jobs.ensure((size_t)ns * (nt > 0 ? nt : 1), st, false));
        odb_launch_walk_mixer(mixer->set.d_src.p, mixer->set.d_order.p, mixer->d_jobs.p, mixer->set.d_removed.p,
                              (int)mixer->set.removed_cap, mixer->d_counters.p, cb, st);
        launches++;
    }
    if (nt > 0) {
        int n_unit = 0, n_gen = 0, n_res = 0;
        if (ns > 0) {
            const bool use_unit = mixer->variant != 1;
            cudaError_t e;
            if (use_unit) {  // staged resampling kernel (leaves at once when the walk kernel counted no such job)
                n_res = odb_mixer_resample_ctas(ns, ctx->sm_count);
                ODB_TRY(mixer->d_partials_res.ensure((size_t)nt * n_res * tile_floats, st, false));
                e = odb_launch_mixer_resample(mixer->d_jobs.p, ns, nt, ch, mixer->d_partials_res.p, n_res, mixer->d_counters.p, st);
                if (e != cudaSuccess) return odb_fail(ODB_E_CUDA, "mixer_resample launch failed: %s", cudaGetErrorString(e));
                launches++;
            }
            if (use_unit) {
                n_unit = odb_mixer_ctas(ns, ctx->sm_count, ch == 1 ? 3 : 1);
                ODB_TRY(mixer->d_partials_unit.ensure((size_t)nt * n_unit * tile_floats, st, false));
                e = odb_launch_mixer_unit(mixer->d_jobs.p, ns, nt, ch, mixer->d_partials_unit.p, n_unit, st);
                if (e != cudaSuccess) return odb_fail(ODB_E_CUDA, "mixer_unit launch failed: %s", cudaGetErrorString(e));
                launches++;
            }
            n_gen = odb_mixer_ctas(ns, ctx->sm_count, ch == 1 ? 4 : 2);
            ODB_TRY(mixer->d_partials_gen.ensure((size_t)nt * n_gen * tile_floats, st, false));
            e = odb_launch_mixer_general(mixer->d_jobs.p, ns, nt, ch, mixer->d_partials_gen.p, n_gen, use_unit ? 1 : 0,
                                         mixer->d_counters.p, st);
            if (e != cudaSuccess) return odb_fail(ODB_E_CUDA, "mixer_general launch failed: %s", cudaGetErrorString(e));
            launches++;
        }
        odb_launch_reduce(mixer->d_partials_unit.p, n_unit, mixer->d_partials_gen.p, n_gen, mixer->d_partials_res.p, n_res,
                          mixer->d_counters.p, n_unit > 0 ? 1 : 0, n_res > 0 ? ODB_CNT_RESAMPLE : -1, nullptr, dev_out, (int)n_frames, nt, ch, mixer->epilogue | (as_i16 ? ODB_EPILOGUE_I16_BIT : 0), st);
        launches++;
    }
    ODB_TRY(mixer->set.post_callback(ctx, st));  // audio-side state only: no control-plane lock
    mixer->last_launches = launches;
    ODB_CUDA(cudaGetLastError());
    return ODB_OK;
}

extern "C" int odb_mixer_sample(odb_mixer* mixer, float interval, float* out, uint32_t n_frames) {
    ODB_TRY(mixer_check(mixer));
    if (!out && n_frames) return odb_fail(ODB_E_INVALID, "out is NULL");
    odb_ctx* ctx = mixer->ctx;
    ODB_CUDA(cudaSetDevice(ctx->device));
    size_t n = (size_t)n_frames * mixer->channels;
    ODB_TRY(mixer->h_out.ensure(n ? n : 2));
    // the reduce kernel stores the tile straight into the pinned host buffer (unified addressing)
    ODB_TRY(mixer_sample_impl(mixer, interval, mixer->h_out.p, n_frames));
    ODB_CUDA(cudaStreamSynchronize(ctx->stream));
    if (n) memcpy(out, mixer->h_out.p, n * sizeof(float));
    return mixer->set.fold_removed(ctx, ctx->stream, true, &mixer->mu);
}
// Offline render (examples/offline.rs:33-43): one callback quantised to 16-bit PCM on the device
extern "C" int odb_mixer_sample_i16(odb_mixer* mixer, float interval, int16_t* out, uint32_t n_frames) {
    ODB_TRY(mixer_check(mixer));
    if (!out && n_frames) return odb_fail(ODB_E_INVALID, "out is NULL");
    odb_ctx* ctx = mixer->ctx;
    ODB_CUDA(cudaSetDevice(ctx->device));
    size_t n = (size_t)n_frames * mixer->channels;
    ODB_TRY(mixer->h_out.ensure(n ? n : 2));
    ODB_TRY(mixer_sample_impl(mixer, interval, mixer->h_out.p, n_frames, /*as_i16=*/true));
    ODB_CUDA(cudaStreamSynchronize(ctx->stream));
    if (n) memcpy(out, mixer->h_out.p, n * sizeof(int16_t));
    return mixer->set.fold_removed(ctx, ctx->stream, true, &mixer->mu);
}
// oddio::run, lib.rs:90-93
extern "C" int odb_mixer_run(odb_mixer* mixer, uint32_t sample_rate, float* out, uint32_t n_frames) {
    float interval = 1.0f / (float)sample_rate;  // lib.rs:91
    return odb_mixer_sample(mixer, interval, out, n_frames);
}
extern "C" int odb_mixer_sample_device(odb_mixer* mixer, float interval, void* dev_out, uint32_t n_frames) {
    ODB_TRY(mixer_check(mixer));
    if (!dev_out && n_frames) return odb_fail(ODB_E_INVALID, "dev_out is NULL");
    return mixer_sample_impl(mixer, interval, (float*)dev_out, n_frames);
}
extern "C" int odb_mixer_len(odb_mixer* mixer, uint64_t* out) {
    ODB_TRY(mixer_check(mixer));
    if (!out) return odb_fail(ODB_E_INVALID, "NULL argument");
    std::lock_guard<std::mutex> lk(mixer->mu);
    *out = mixer->set.order.size();
    return ODB_OK;
}

int odb_mixer_owner_set(void* owner, odb_source src, odb_ctx** ctx, std::mutex** mu, SourceSet** set, uint32_t* tag) {
    odb_mixer* m = (odb_mixer*)owner;
    *tag = (uint32_t)((src >> 32) & 0xFF);
    if (*tag != ODB_TAG_MIXED) return odb_fail(ODB_E_INVALID, "not a source of this mixer");
    *ctx = m->ctx; *mu = &m->mu; *set = &m->set;
    return ODB_OK;
}
int odb_mixer_last_launches(void* owner, uint32_t* out) { *out = ((odb_mixer*)owner)->last_launches; return ODB_OK; }
int odb_mixer_set_variant(void* owner, int variant) { ((odb_mixer*)owner)->variant = variant; return ODB_OK; }
int odb_mixer_job_counters(void* owner, uint32_t out[4]) {
    odb_mixer* m = (odb_mixer*)owner;
    for (int i = 0; i < 4; i++) out[i] = 0;
    if (!m->d_counters.p) return ODB_OK;
    ODB_CUDA(cudaSetDevice(m->ctx->device));
    ODB_CUDA(cudaMemcpyAsync(out, m->d_counters.p, ODB_CNT_WORDS * sizeof(uint32_t), cudaMemcpyDeviceToHost, m->ctx->stream));
    ODB_CUDA(cudaStreamSynchronize(m->ctx->stream));
    return ODB_OK;
}
