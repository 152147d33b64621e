// odb_mixer_*: MixerControl<T> + Mixer<T> over a device-resident source set.
// Reference: src/mixer.rs (cited per function), src/speed.rs, src/gain.rs, src/tanh.rs.
#include "odb_host.h"

#define ODB_TAG_MIXED 3u

struct odb_mixer {
    uint32_t kind = ODB_KIND_MIXER;
    odb_ctx* ctx = nullptr;
    int channels = 2;
    std::mutex mu;
    SourceSet set;
    int epilogue = ODB_EPILOGUE_NONE;
    int variant = 0;
    uint32_t last_launches = 0;
};

static int mixer_check(odb_mixer* m) {
    if (!m || m->kind != ODB_KIND_MIXER) return odb_fail(ODB_E_INVALID, "not a mixer handle");
    return ODB_OK;
}

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
extern "C" int odb_mixer_play(odb_mixer* mixer, const odb_chain* chain, odb_source* out) {
    ODB_TRY(mixer_check(mixer));
    (void)chain; (void)out;
    return odb_fail(ODB_E_UNSUPPORTED, "mixer device path not built yet");
}
extern "C" int odb_mixed_stop(odb_mixer* mixer, odb_source src) {
    ODB_TRY(mixer_check(mixer));
    (void)src;
    return odb_fail(ODB_E_UNSUPPORTED, "mixer device path not built yet");
}
extern "C" int odb_mixed_is_stopped(odb_mixer* mixer, odb_source src, int* out) {
    ODB_TRY(mixer_check(mixer));
    (void)src; (void)out;
    return odb_fail(ODB_E_UNSUPPORTED, "mixer device path not built yet");
}
extern "C" int odb_mixer_sample(odb_mixer* mixer, float interval, float* out, uint32_t n_frames) {
    ODB_TRY(mixer_check(mixer));
    (void)interval; (void)out; (void)n_frames;
    return odb_fail(ODB_E_UNSUPPORTED, "mixer device path not built yet");
}
extern "C" int odb_mixer_run(odb_mixer* mixer, uint32_t sample_rate, float* out, uint32_t n_frames) {
    float interval = 1.0f / (float)sample_rate;  // lib.rs:91
    return odb_mixer_sample(mixer, interval, out, n_frames);
}
extern "C" int odb_mixer_sample_device(odb_mixer* mixer, float interval, void* dev_out, uint32_t n_frames) {
    ODB_TRY(mixer_check(mixer));
    (void)interval; (void)dev_out; (void)n_frames;
    return odb_fail(ODB_E_UNSUPPORTED, "mixer device path not built yet");
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
int odb_mixer_job_counters(void* owner, uint32_t out[4]) { (void)owner; for (int i = 0; i < 4; i++) out[i] = 0; return ODB_OK; }
