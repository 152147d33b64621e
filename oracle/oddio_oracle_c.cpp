// C API over oddio_oracle.hpp so tests/bench can drive the oracle through ctypes.
// TEST INFRASTRUCTURE ONLY — see the header of oddio_oracle.hpp.
#include "oddio_oracle.hpp"

#include <chrono>
#include <condition_variable>
#include <mutex>
#include <thread>

using namespace orc;

namespace {
enum Kind { K_FRAMES = 1, K_SIGNAL, K_MIXED, K_SPATIAL, K_RING, K_SMOOTHED };
struct Obj {
    Kind kind;
    FramesP frames;
    SignalP signal;
    std::shared_ptr<bool> flag;
    SpatialHandle spatial;
    std::unique_ptr<Ring> ring;
    Smoothed smoothed;
};
Obj* mk_signal(SignalP s) { Obj* o = new Obj(); o->kind = K_SIGNAL; o->signal = s; return o; }
template <class T> T* as(void* h) { return dynamic_cast<T*>(((Obj*)h)->signal.get()); }

// Test fixtures of the reference's own unit tests, needed to replay them verbatim.
struct TimeSignal : Signal {  // ring.rs:86-97
    float t; explicit TimeSignal(float t0) : t(t0) {}
    int channels() const override { return 1; }
    void sample(float interval, float* out, size_t n) override { for (size_t i = 0; i < n; i++) { out[i] = t; t = t + interval; } }
};
struct CountingSignal : Signal {  // signal.rs:97-108
    uint32_t c; explicit CountingSignal(uint32_t c0) : c(c0) {}
    int channels() const override { return 1; }
    void sample(float, float* out, size_t n) override { for (size_t i = 0; i < n; i++) { out[i] = (float)c; c = c + 1; } }
};
struct FinishedSignal : Signal {  // spatial.rs:611-627
    int channels() const override { return 1; }
    void sample(float, float* out, size_t n) override { for (size_t i = 0; i < n; i++) out[i] = 0.0f; }
    bool is_finished() const override { return true; }
    bool can_seek() const override { return true; }
};
}  // namespace

extern "C" {

void orc_release(void* h) { delete (Obj*)h; }

// ---- sources --------------------------------------------------------------------------
void* orc_frames_new(uint32_t rate, int channels, const float* samples, size_t n_frames) {
    Obj* o = new Obj(); o->kind = K_FRAMES;
    o->frames = std::make_shared<Frames>();
    o->frames->rate = (double)rate; o->frames->ch = channels;
    o->frames->samples.assign(samples, samples + n_frames * (size_t)channels);
    return o;
}
void* orc_frames_signal_new(void* frames, double start) { return mk_signal(std::make_shared<FramesSignal>(((Obj*)frames)->frames, start)); }
void* orc_cycle_new(void* frames) { return mk_signal(std::make_shared<Cycle>(((Obj*)frames)->frames)); }
void* orc_constant_new(int ch, const float* v) { return mk_signal(std::make_shared<Constant>(ch, v)); }
void* orc_sine_new(float phase, float hz) { return mk_signal(std::make_shared<Sine>(phase, hz)); }
void* orc_time_signal_new(float t0) { return mk_signal(std::make_shared<TimeSignal>(t0)); }
void* orc_counting_signal_new(uint32_t c0) { return mk_signal(std::make_shared<CountingSignal>(c0)); }
void* orc_finished_signal_new() { return mk_signal(std::make_shared<FinishedSignal>()); }

// ---- filters --------------------------------------------------------------------------
void* orc_mono_to_stereo_new(void* s) { return mk_signal(std::make_shared<MonoToStereo>(((Obj*)s)->signal)); }
void* orc_speed_new(void* s) { return mk_signal(std::make_shared<Speed>(((Obj*)s)->signal)); }
void orc_speed_set(void* h, float f) { as<Speed>(h)->speed = f; }
float orc_speed_get(void* h) { return as<Speed>(h)->speed; }
void* orc_gain_new(void* s) { return mk_signal(std::make_shared<Gain>(((Obj*)s)->signal)); }
void orc_gain_set_initial_ratio(void* h, float f) { as<Gain>(h)->set_initial_amplitude_ratio(f); }
void orc_gain_set_initial_db(void* h, float db) { as<Gain>(h)->set_initial_gain_db(db); }
void orc_gain_control_set_ratio(void* h, float f) { as<Gain>(h)->control_set_amplitude_ratio(f); }
void orc_gain_control_set_db(void* h, float db) { as<Gain>(h)->control_set_gain_db(db); }
float orc_gain_control_ratio(void* h) { return as<Gain>(h)->shared; }
float orc_gain_control_db(void* h) { return as<Gain>(h)->control_gain_db(); }
void* orc_fixed_gain_new(void* s, float db) { return mk_signal(std::make_shared<FixedGain>(((Obj*)s)->signal, db)); }
float orc_fixed_gain_value(void* h) { return as<FixedGain>(h)->gain; }
void* orc_tanh_new(void* s) { return mk_signal(std::make_shared<Tanh>(((Obj*)s)->signal)); }
void* orc_reinhard_new(void* s) { return mk_signal(std::make_shared<Reinhard>(((Obj*)s)->signal)); }

// ---- Signal / Seek / run -----------------------------------------------------------------
int orc_signal_channels(void* h) { return ((Obj*)h)->signal->channels(); }
void orc_signal_sample(void* h, float interval, float* out, size_t n) { ((Obj*)h)->signal->sample(interval, out, n); }
void orc_run(void* h, uint32_t rate, float* out, size_t n) { run(*((Obj*)h)->signal, rate, out, n); }
int orc_signal_is_finished(void* h) { return ((Obj*)h)->signal->is_finished() ? 1 : 0; }
void orc_signal_seek(void* h, float s) { ((Obj*)h)->signal->seek(s); }
// Times `reps` calls of run() and returns the best wall time in seconds (steady_clock); used
// by bench.py's cpu_baseline so that the timed region contains no Python.
double orc_time_run(void* h, uint32_t rate, float* out, size_t n, int warmup, int reps) {
    Signal& s = *((Obj*)h)->signal;
    for (int i = 0; i < warmup; i++) run(s, rate, out, n);
    double best = 1e300;
    for (int i = 0; i < reps; i++) {
        auto t0 = std::chrono::steady_clock::now();
        run(s, rate, out, n);
        auto t1 = std::chrono::steady_clock::now();
        double dt = std::chrono::duration<double>(t1 - t0).count();
        if (dt < best) best = dt;
    }
    return best;
}

// Source-sharded CPU run: `n_sig` independent aggregators (each holding a shard of the sources) are run
// on one thread each into private tiles which are then summed. NOT reference behaviour (the reference is
// single-threaded per signal graph, signal.rs:19); it is the "all host threads" arm of bench.py.
// Returns total wall seconds of `reps` rounds (after `warmup` untimed rounds); out = last summed tile.
double orc_time_run_sharded(void** sigs, int n_sig, uint32_t rate, float* out, size_t n, int channels, int warmup, int reps) {
    std::vector<std::vector<float>> tiles((size_t)n_sig, std::vector<float>(n * (size_t)channels, 0.0f));
    // Persistent workers, released round by round (no thread is created or joined inside the timed region).
    std::mutex mu;
    std::condition_variable cv_go, cv_done;
    int round_no = 0, done = 0;
    bool quit = false;
    std::vector<std::thread> th;
    for (int i = 0; i < n_sig; i++)
        th.emplace_back([&, i]() {
            int seen = 0;
            for (;;) {
                {
                    std::unique_lock<std::mutex> lk(mu);
                    cv_go.wait(lk, [&] { return quit || round_no != seen; });
                    if (quit) return;
                    seen = round_no;
                }
                run(*((Obj*)sigs[i])->signal, rate, tiles[(size_t)i].data(), n);
                {
                    std::lock_guard<std::mutex> lk(mu);
                    if (++done == n_sig) cv_done.notify_one();
                }
            }
        });
    auto round = [&]() {
        {
            std::lock_guard<std::mutex> lk(mu);
            done = 0;
            round_no++;
        }
        cv_go.notify_all();
        {
            std::unique_lock<std::mutex> lk(mu);
            cv_done.wait(lk, [&] { return done == n_sig; });
        }
        for (size_t k = 0; k < n * (size_t)channels; k++) {
            float acc = 0.0f;
            for (int i = 0; i < n_sig; i++) acc = acc + tiles[(size_t)i][k];
            out[k] = acc;
        }
    };
    for (int i = 0; i < warmup; i++) round();
    auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < reps; i++) round();
    auto t1 = std::chrono::steady_clock::now();
    {
        std::lock_guard<std::mutex> lk(mu);
        quit = true;
    }
    cv_go.notify_all();
    for (auto& t : th) t.join();
    return std::chrono::duration<double>(t1 - t0).count();
}

// ---- FramesSignal read-backs ---------------------------------------------------------------
double orc_frames_signal_t(void* h) { return as<FramesSignal>(h)->t; }
long orc_frames_signal_sample_t(void* h) { return as<FramesSignal>(h)->sample_t; }
double orc_frames_signal_playback_position(void* h) { return as<FramesSignal>(h)->playback_position(); }
int orc_frames_signal_control_is_finished(void* h) { return as<FramesSignal>(h)->control_is_finished() ? 1 : 0; }
double orc_cycle_cursor(void* h) { return as<Cycle>(h)->cursor; }

// ---- Mixer -----------------------------------------------------------------------------
void* orc_mixer_new(int channels) { return mk_signal(std::make_shared<Mixer>(channels)); }
void* orc_mixer_play(void* mixer, void* sig) {
    Obj* o = new Obj(); o->kind = K_MIXED;
    o->flag = as<Mixer>(mixer)->play(((Obj*)sig)->signal);
    return o;
}
void orc_mixed_stop(void* h) { *((Obj*)h)->flag = true; }
int orc_mixed_is_stopped(void* h) { return *((Obj*)h)->flag ? 1 : 0; }
size_t orc_mixer_len(void* mixer) { return as<Mixer>(mixer)->set.len(); }

// ---- SpatialScene ------------------------------------------------------------------------
void* orc_scene_new() { return mk_signal(std::make_shared<SpatialScene>()); }
void* orc_scene_play(void* scene, void* sig, const float* pos, const float* vel, float radius) {
    Obj* o = new Obj(); o->kind = K_SPATIAL;
    o->spatial = as<SpatialScene>(scene)->play(((Obj*)sig)->signal, Vec3{pos[0], pos[1], pos[2]}, Vec3{vel[0], vel[1], vel[2]}, radius);
    return o;
}
void* orc_scene_play_buffered(void* scene, void* sig, const float* pos, const float* vel, float radius,
                              float max_distance, uint32_t rate, float buffer_duration) {
    Obj* o = new Obj(); o->kind = K_SPATIAL;
    o->spatial = as<SpatialScene>(scene)->play_buffered(((Obj*)sig)->signal, Vec3{pos[0], pos[1], pos[2]},
                                                        Vec3{vel[0], vel[1], vel[2]}, radius, max_distance, rate, buffer_duration);
    return o;
}
void orc_scene_set_listener_rotation(void* scene, const float* q_xyzs) {  // mint layout {v:{x,y,z}, s}
    as<SpatialScene>(scene)->set_listener_rotation(Quat{q_xyzs[3], {q_xyzs[0], q_xyzs[1], q_xyzs[2]}});
}
size_t orc_scene_len(void* scene, int buffered) {
    SpatialScene* s = as<SpatialScene>(scene);
    return buffered ? s->recv_buffered.len() : s->recv.len();
}
void orc_spatial_set_motion(void* h, const float* pos, const float* vel, int disc) {
    ((Obj*)h)->spatial.set_motion(Vec3{pos[0], pos[1], pos[2]}, Vec3{vel[0], vel[1], vel[2]}, disc != 0);
}
int orc_spatial_is_finished(void* h) { return ((Obj*)h)->spatial.is_finished() ? 1 : 0; }
// out[0..2]=state.prev_position, [3]=state.dt, [4]=finished_for, [5]=has_finished_for, [6]=stopped
void orc_spatial_state(void* h, float* out) {
    Common& c = *((Obj*)h)->spatial.common;
    out[0] = c.state.prev_position.x; out[1] = c.state.prev_position.y; out[2] = c.state.prev_position.z;
    out[3] = c.state.dt; out[4] = c.finished_for; out[5] = c.has_finished_for ? 1.0f : 0.0f; out[6] = *c.stopped ? 1.0f : 0.0f;
}
// f64-accumulated mix of the last sample() call of a Mixer or SpatialScene.
size_t orc_out64(void* h, double* out, size_t cap) {
    const std::vector<double>* v = nullptr;
    if (Mixer* m = as<Mixer>(h)) v = &m->out64;
    else if (SpatialScene* s = as<SpatialScene>(h)) v = &s->out64;
    if (!v) return 0;
    size_t n = v->size() < cap ? v->size() : cap;
    memcpy(out, v->data(), n * sizeof(double));
    return n;
}

// ---- small pieces exposed for known-answer tests --------------------------------------------
void* orc_ring_new(size_t cap) { Obj* o = new Obj(); o->kind = K_RING; o->ring.reset(new Ring(cap)); return o; }
void orc_ring_write(void* ring, void* sig, uint32_t rate, float dt) { ((Obj*)ring)->ring->write_from(*((Obj*)sig)->signal, rate, dt); }
void orc_ring_delay(void* ring, uint32_t rate, float dt) { ((Obj*)ring)->ring->delay(rate, dt); }
void orc_ring_sample(void* ring, uint32_t rate, float t, float interval, float* out, size_t n) { ((Obj*)ring)->ring->sample(rate, t, interval, out, n); }
float orc_ring_write_cursor(void* ring) { return ((Obj*)ring)->ring->write; }
size_t orc_ring_buffer(void* ring, float* out, size_t cap) {
    Ring& r = *((Obj*)ring)->ring; size_t n = r.buffer.size() < cap ? r.buffer.size() : cap;
    memcpy(out, r.buffer.data(), n * sizeof(float)); return r.buffer.size();
}
void* orc_smoothed_new(float x) { Obj* o = new Obj(); o->kind = K_SMOOTHED; o->smoothed = Smoothed(x); return o; }
void orc_smoothed_set(void* h, float v) { ((Obj*)h)->smoothed.set(v); }
void orc_smoothed_advance(void* h, float p) { ((Obj*)h)->smoothed.advance(p); }
float orc_smoothed_get(void* h) { return ((Obj*)h)->smoothed.get(); }
float orc_smoothed_progress(void* h) { return ((Obj*)h)->smoothed.progress; }

void orc_rotate(const float* q_xyzs, const float* p, float* out) {
    Vec3 r = rotate(Quat{q_xyzs[3], {q_xyzs[0], q_xyzs[1], q_xyzs[2]}}, Vec3{p[0], p[1], p[2]});
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
void orc_ear_state(const float* p, int ear, float radius, float* out) {
    EarState e = ear_state(Vec3{p[0], p[1], p[2]}, (Ear)ear, radius);
    out[0] = e.offset; out[1] = e.gain;
}
// frames.rs:94-102 Frames::interpolate (mono)
float orc_frames_interpolate(void* frames, double s) {
    Frames& f = *((Obj*)frames)->frames;
    long x0 = (long)s; float fract = (float)(s - (double)x0); float a[2], b[2];
    f.get_pair(x0, a, b);
    return lerp1(a[0], b[0], fract);
}

}  // extern "C"
