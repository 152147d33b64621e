// oddio_oracle.hpp — CPU restatement of Ralith/oddio 0.7.4's spatial/mixer hot path.
//
// TEST INFRASTRUCTURE ONLY. This is the parity checker for the CUDA path; the only
// callers allowed are tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// `--impl reference` legs. Nothing under oddio_b200/ may include, link or call it.
//
// Provenance: the reference is Rust and cannot be compiled in this image (no rustc/cargo),
// so this is a "port" oracle: every function restates the cited reference lines with the
// same operation order, compiled with -ffp-contract=off (Rust never contracts to FMA) on
// x86-64 SSE2 (no excess precision). It is PINNED against every known-answer vector the
// reference's own unit tests hold for this path (tests/test_oracle_kat.py), see SURVEY.md §4.
// What the reference itself leaves unpinned (the numeric output of SpatialScene::sample) is
// unpinned here too and is stated so in DESIGN.md; for that part a second restatement, written
// separately from the reference in numpy (tests/independent.py), must agree with this one bit
// for bit (tests/test_oracle_independent.py, tests/test_golden.py).
//
// libm: the reference's std build calls the platform libm (sinf/tanhf/powf/log10f); on Linux
// that is glibc, which is what this file calls.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <deque>
#include <memory>
#include <vector>

namespace orc {

typedef float Sample;  // lib.rs:85

// ---------------------------------------------------------------------------------------
// frame.rs:39-51 — per-channel frame algebra, never fused.
static inline float lerp1(float a, float b, float t) { return a + t * (b - a); }  // frame.rs:39-41

// ---------------------------------------------------------------------------------------
// math/mod.rs:33-94 — 3-vector and quaternion helpers (mint types are plain structs).
struct Vec3 { float x, y, z; };
struct Quat { float s; Vec3 v; };  // mint::Quaternion {v, s}

static inline float norm(Vec3 a) {  // math/mod.rs:33-35: map(powi(2)).sum().sqrt(), sum from 0 left-to-right
    float acc = 0.0f;
    acc = acc + a.x * a.x;
    acc = acc + a.y * a.y;
    acc = acc + a.z * a.z;
    return sqrtf(acc);
}
static inline float dot(Vec3 a, Vec3 b) {  // math/mod.rs:37-43
    float acc = 0.0f;
    acc = acc + a.x * b.x;
    acc = acc + a.y * b.y;
    acc = acc + a.z * b.z;
    return acc;
}
static inline Vec3 scale(Vec3 v, float f) { return {v.x * f, v.y * f, v.z * f}; }            // :45-47
static inline Vec3 sub(Vec3 a, Vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }         // :49-51
static inline Vec3 add(Vec3 a, Vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }         // :53-55
static inline Vec3 mix(Vec3 a, Vec3 b, float r) {                                             // :57-60
    float ir = 1.0f - r;
    return {ir * a.x + r * b.x, ir * a.y + r * b.y, ir * a.z + r * b.z};
}
static inline Quat invert_quat(const Quat& q) { return {q.s, {-q.v.x, -q.v.y, -q.v.z}}; }   // :62-67
static inline Quat quat_mul(const Quat& q, const Quat& r) {                                  // :69-79
    Quat o;
    o.s = q.s * r.s - q.v.x * r.v.x - q.v.y * r.v.y - q.v.z * r.v.z;
    o.v.x = q.s * r.v.x + q.v.x * r.s + q.v.y * r.v.z - q.v.z * r.v.y;
    o.v.y = q.s * r.v.y - q.v.x * r.v.z + q.v.y * r.s + q.v.z * r.v.x;
    o.v.z = q.s * r.v.z + q.v.x * r.v.y - q.v.y * r.v.x + q.v.z * r.s;
    return o;
}
static inline Vec3 rotate(const Quat& rot, const Vec3& p) {                                  // :81-94
    Quat pq = {0.0f, p};
    return quat_mul(rot, quat_mul(pq, invert_quat(rot))).v;
}
static inline float rem_euclid(float a, float b) {  // f32::rem_euclid (core): r = a % b; if r < 0 { r + |b| }
    float r = fmodf(a, b);
    return r < 0.0f ? r + fabsf(b) : r;
}
static inline double rem_euclid64(double a, double b) {
    double r = fmod(a, b);
    return r < 0.0 ? r + fabs(b) : r;
}

// ---------------------------------------------------------------------------------------
// signal.rs:14-58 — the plugin traits. `channels` stands in for the associated Frame type
// (1 = Sample, 2 = [Sample; 2]); buffers are interleaved.
struct Signal {
    virtual ~Signal() {}
    virtual int channels() const = 0;
    virtual void sample(float interval, float* out, size_t n) = 0;  // signal.rs:19
    virtual bool is_finished() const { return false; }             // signal.rs:24-27
    virtual bool can_seek() const { return false; }
    virtual void seek(float /*seconds*/) {}                          // signal.rs:48-51
};
typedef std::shared_ptr<Signal> SignalP;

// ---------------------------------------------------------------------------------------
// frames.rs:19-123 — static PCM.
struct Frames {
    double rate;  // frames.rs:20 (u32 -> f64, :40)
    int ch;
    std::vector<float> samples;  // interleaved, len*ch
    size_t len() const { return samples.size() / (size_t)ch; }
    // frames.rs:105-123 get_pair; writes a[ch], b[ch]
    inline void get_pair(long sample, float* a, float* b) const {
        const long n = (long)len();
        const float* s = samples.data();
        if (sample >= 0) {
            if (sample < n - 1) {
                for (int c = 0; c < ch; c++) { a[c] = s[sample * ch + c]; b[c] = s[(sample + 1) * ch + c]; }
            } else if (sample < n) {
                for (int c = 0; c < ch; c++) { a[c] = s[sample * ch + c]; b[c] = 0.0f; }
            } else {
                for (int c = 0; c < ch; c++) { a[c] = 0.0f; b[c] = 0.0f; }
            }
        } else if (sample < -1) {
            for (int c = 0; c < ch; c++) { a[c] = 0.0f; b[c] = 0.0f; }
        } else {
            for (int c = 0; c < ch; c++) { a[c] = 0.0f; b[c] = s[c]; }
        }
    }
};
typedef std::shared_ptr<Frames> FramesP;

// frames.rs:141-220 — FramesSignal (+ the FramesSignalControl read-backs, :229-248)
struct FramesSignal : Signal {
    FramesP data;
    double t;        // frames.rs:145
    long sample_t;   // frames.rs:149 (AtomicIsize)
    FramesSignal(FramesP d, double start) : data(d), t(start), sample_t((long)(start * d->rate)) {}  // :156-161
    int channels() const override { return data->ch; }
    void sample(float interval, float* out, size_t n) override {  // frames.rs:176-201
        const int ch = data->ch;
        double s0 = t * data->rate;                       // :177
        float ds = interval * (float)data->rate;          // :178
        long base = (long)s0;                             // :179  (as isize: trunc toward zero)
        float a[2], b[2];
        if (fabsf(ds - 1.0f) <= 1.1920929e-7f) {          // :180  f32::EPSILON
            float fract = (float)(s0 - (double)base);     // :183
            for (size_t i = 0; i < n; i++) {              // :184-187
                data->get_pair(base + (long)i, a, b);
                for (int c = 0; c < ch; c++) out[i * ch + c] = lerp1(a[c], b[c], fract);
            }
        } else {
            float offset = (float)(s0 - (double)base);    // :189
            for (size_t i = 0; i < n; i++) {              // :190-196
                long trunc = (long)offset;                // to_int_unchecked::<isize>
                data->get_pair(base + trunc, a, b);
                float fract = offset - (float)trunc;
                for (int c = 0; c < ch; c++) out[i * ch + c] = lerp1(a[c], b[c], fract);
                offset += ds;
            }
        }
        t += (double)interval * (double)n;                // :198
        sample_t = (long)(t * data->rate);                // :199-200
    }
    bool is_finished() const override {                   // :204-206
        return t >= (double)(data->len() - 1) / data->rate;
    }
    bool can_seek() const override { return true; }
    void seek(float seconds) override { t += (double)seconds; }  // :211-213
    // FramesSignalControl
    double playback_position() const { return (double)sample_t / data->rate; }  // :238-240
    bool control_is_finished() const {                                           // :244-247
        return sample_t >= 0 && (size_t)sample_t >= data->len();
    }
};

// cycle.rs:6-61 — looping PCM ("next" row of SURVEY §8f)
struct Cycle : Signal {
    double cursor;  // in samples
    FramesP frames;
    explicit Cycle(FramesP f) : cursor(0.0), frames(f) {}
    int channels() const override { return frames->ch; }
    inline void pair_wrap(size_t x, float* a, float* b) const {
        const int ch = frames->ch;
        const size_t n = frames->len();
        const float* s = frames->samples.data();
        if (x < n - 1) { for (int c = 0; c < ch; c++) { a[c] = s[x * ch + c]; b[c] = s[(x + 1) * ch + c]; } }
        else { for (int c = 0; c < ch; c++) { a[c] = s[x * ch + c]; b[c] = s[c]; } }
    }
    void sample(float interval, float* out, size_t n) override {  // cycle.rs:26-53
        const int ch = frames->ch;
        const size_t len = frames->len();
        float ds = interval * (float)(uint32_t)frames->rate;   // :27 rate() as f32
        size_t base = (size_t)cursor;                          // :28
        float offset = (float)(cursor - (double)base);         // :29
        float a[2], b[2];
        for (size_t i = 0; i < n; i++) {
            size_t trunc = (size_t)offset;                     // :31
            float fract = offset - (float)trunc;               // :32
            size_t x = base + trunc;                           // :33
            if (x < len - 1 || x < len) {                      // :34-37
                pair_wrap(x, a, b);
            } else {                                           // :38-47
                base = 0;
                offset = (float)(x % len) + fract;
                size_t x2 = (size_t)offset;
                pair_wrap(x2, a, b);
            }
            for (int c = 0; c < ch; c++) out[i * ch + c] = lerp1(a[c], b[c], fract);
            offset += ds;                                      // :50
        }
        cursor = (double)base + (double)offset;                // :52
    }
    bool can_seek() const override { return true; }
    void seek(float seconds) override {                        // :57-60
        cursor = rem_euclid64(cursor + (double)seconds * (double)(uint32_t)frames->rate, (double)frames->len());
    }
};

// constant.rs:4-23
struct Constant : Signal {
    int ch; float v[2];
    Constant(int c, const float* f) : ch(c) { v[0] = f[0]; v[1] = c > 1 ? f[1] : 0.0f; }
    int channels() const override { return ch; }
    void sample(float, float* out, size_t n) override { for (size_t i = 0; i < n; i++) for (int c = 0; c < ch; c++) out[i * ch + c] = v[c]; }
    bool can_seek() const override { return true; }
};

// sine.rs:6-47
struct Sine : Signal {
    float phase, frequency;
    Sine(float ph, float hz) : phase(ph), frequency(hz * 6.28318530717958647692f) {}  // :21 TAU
    int channels() const override { return 1; }
    void seek_to(float t) { phase = fmodf(phase + t * frequency, 6.28318530717958647692f); }  // :25-28
    void sample(float interval, float* out, size_t n) override {                                // :34-40
        for (size_t i = 0; i < n; i++) {
            float t = interval * (float)i;
            out[i] = sinf(t * frequency + phase);
        }
        seek_to(interval * (float)n);
    }
    bool can_seek() const override { return true; }
    void seek(float s) override { seek_to(s); }
};

// signal.rs:61-91
struct MonoToStereo : Signal {
    SignalP inner;
    explicit MonoToStereo(SignalP s) : inner(s) {}
    int channels() const override { return 2; }
    void sample(float interval, float* out, size_t n) override {  // :73-80
        inner->sample(interval, out, n);
        for (size_t i = 2 * n; i-- > 0;) out[i] = out[i / 2];
    }
    bool is_finished() const override { return inner->is_finished(); }
    bool can_seek() const override { return inner->can_seek(); }
    void seek(float s) override { inner->seek(s); }
};

// speed.rs:9-55
struct Speed : Signal {
    float speed;  // Arc<AtomicU32> bits
    SignalP inner;
    explicit Speed(SignalP s) : speed(1.0f), inner(s) {}
    int channels() const override { return inner->channels(); }
    void sample(float interval, float* out, size_t n) override { inner->sample(interval * speed, out, n); }  // :32-35
    bool is_finished() const override { return inner->is_finished(); }
};

// smooth.rs:26-91
struct Smoothed {
    float prev, next, progress;
    explicit Smoothed(float x = 0.0f) : prev(x), next(x), progress(1.0f) {}     // :34-43
    void advance(float p) { progress = fminf(progress + p, 1.0f); }            // :47-49
    float get() const { float diff = next - prev; return prev + progress * diff; }  // :67-72, :86-91
    void set(float v) { prev = get(); next = v; progress = 0.0f; }             // :57-64
};

// gain.rs:9-51
struct FixedGain : Signal {
    float gain; SignalP inner;
    FixedGain(SignalP s, float db) : gain(powf(10.0f, db / 20.0f)), inner(s) {}  // :18-23
    int channels() const override { return inner->channels(); }
    void sample(float interval, float* out, size_t n) override {                  // :32-37
        inner->sample(interval, out, n);
        size_t m = n * (size_t)channels();
        for (size_t i = 0; i < m; i++) out[i] = out[i] * gain;
    }
    bool is_finished() const override { return inner->is_finished(); }
    bool can_seek() const override { return inner->can_seek(); }
    void seek(float s) override { inner->seek(s); }
};

// gain.rs:58-160
static const float GAIN_SMOOTHING_PERIOD = 0.1f;  // gain.rs:163
struct Gain : Signal {
    float shared;  // Arc<AtomicU32> bits
    Smoothed gain;
    SignalP inner;
    explicit Gain(SignalP s) : shared(1.0f), gain(1.0f), inner(s) {}
    void set_initial_amplitude_ratio(float f) { shared = f; gain = Smoothed(f); }     // :90-93 (Gain::set_amplitude_ratio)
    void set_initial_gain_db(float db) { set_initial_amplitude_ratio(powf(10.0f, db / 20.0f)); }  // :81-83
    void control_set_amplitude_ratio(float f) { shared = f; }                          // :157-159 (GainControl)
    void control_set_gain_db(float db) { shared = powf(10.0f, db / 20.0f); }           // :143-145
    float control_gain_db() const { return 20.0f * log10f(shared); }                   // :134-136
    int channels() const override { return inner->channels(); }
    void sample(float interval, float* out, size_t n) override {                       // :103-122
        inner->sample(interval, out, n);
        const size_t ch = (size_t)channels();
        if (gain.next != shared) gain.set(shared);                                      // :106-108
        if (gain.progress == 1.0f) {                                                    // :109-117
            float g = gain.get();
            if (g != 1.0f) for (size_t i = 0; i < n * ch; i++) out[i] = out[i] * g;
            return;
        }
        for (size_t i = 0; i < n; i++) {                                                // :118-121
            float g = gain.get();
            for (size_t c = 0; c < ch; c++) out[i * ch + c] = out[i * ch + c] * g;
            gain.advance(interval / GAIN_SMOOTHING_PERIOD);
        }
    }
    bool is_finished() const override { return inner->is_finished(); }
};

// tanh.rs:7-44 / reinhard.rs:13-50
struct Tanh : Signal {
    SignalP inner;
    explicit Tanh(SignalP s) : inner(s) {}
    int channels() const override { return inner->channels(); }
    void sample(float interval, float* out, size_t n) override {
        inner->sample(interval, out, n);
        size_t m = n * (size_t)channels();
        for (size_t i = 0; i < m; i++) out[i] = tanhf(out[i]);  // tanh.rs:24-28
    }
    bool is_finished() const override { return inner->is_finished(); }
    bool can_seek() const override { return inner->can_seek(); }
    void seek(float s) override { inner->seek(s); }
};
struct Reinhard : Signal {
    SignalP inner;
    explicit Reinhard(SignalP s) : inner(s) {}
    int channels() const override { return inner->channels(); }
    void sample(float interval, float* out, size_t n) override {
        inner->sample(interval, out, n);
        size_t m = n * (size_t)channels();
        for (size_t i = 0; i < m; i++) out[i] = out[i] / (1.0f + fabsf(out[i]));  // reinhard.rs:30-34
    }
    bool is_finished() const override { return inner->is_finished(); }
    bool can_seek() const override { return inner->can_seek(); }
    void seek(float s) override { inner->seek(s); }
};

// ---------------------------------------------------------------------------------------
// set.rs:141-204, audio side only: membership + order. Inserts arrive in send order at
// `update`; `remove` is Vec::swap_remove; iteration by the owners is (0..len).rev().
template <class T>
struct Set {
    std::vector<T> signals;
    std::deque<T> inbox;  // stands in for the spsc channel (set.rs:55-66)
    void insert(T x) { inbox.push_back(std::move(x)); }
    void update() { while (!inbox.empty()) { signals.push_back(std::move(inbox.front())); inbox.pop_front(); } }  // :141-178
    void remove(size_t i) { signals[i] = std::move(signals.back()); signals.pop_back(); }                          // :183-188
    size_t len() const { return signals.size(); }
};

// ---------------------------------------------------------------------------------------
// mixer.rs:46-120
struct MixedSignal { std::shared_ptr<bool> stop; SignalP inner; };
struct Mixer : Signal {
    int ch;
    Set<MixedSignal> set;
    std::vector<float> buffer;       // mixer.rs:77: 1024 frames
    std::vector<double> out64;       // checker aid: f64-accumulated mix of the last call
    explicit Mixer(int c) : ch(c), buffer(1024 * (size_t)c, 0.0f) {}
    int channels() const override { return ch; }
    std::shared_ptr<bool> play(SignalP s) {                         // mixer.rs:18-26
        MixedSignal m{std::make_shared<bool>(false), s};
        auto h = m.stop;
        set.insert(std::move(m));
        return h;
    }
    void sample(float interval, float* out, size_t n) override {   // mixer.rs:92-119
        set.update();
        const size_t C = (size_t)ch;
        for (size_t i = 0; i < n * C; i++) out[i] = 0.0f;
        out64.assign(n * C, 0.0);
        for (size_t i = set.len(); i-- > 0;) {
            MixedSignal& sig = set.signals[i];
            if (*sig.stop || sig.inner->is_finished()) {            // :102-106
                *sig.stop = true;
                set.remove(i);
                continue;
            }
            size_t done = 0;
            while (done < n) {                                      // :109-117
                size_t m = n - done; if (m > 1024) m = 1024;
                sig.inner->sample(interval, buffer.data(), m);
                for (size_t k = 0; k < m * C; k++) {
                    out[done * C + k] = out[done * C + k] + buffer[k];   // frame::mix, frame.rs:44-46
                    out64[done * C + k] += (double)buffer[k];
                }
                done += m;
            }
        }
    }
};

// ---------------------------------------------------------------------------------------
// ring.rs:4-80
struct Ring {
    std::vector<float> buffer;
    float write;
    explicit Ring(size_t cap) : buffer(cap, 0.0f), write(0.0f) {}
    void write_from(Signal& signal, uint32_t rate, float dt) {                         // ring.rs:18-41
        float end = fmodf(write + dt * (float)rate, (float)buffer.size());             // :28
        size_t start_idx = (size_t)ceilf(write);                                       // :30
        size_t end_idx = (size_t)ceilf(end);                                           // :31
        float interval = 1.0f / (float)rate;                                           // :32
        if (end_idx > start_idx) {
            signal.sample(interval, buffer.data() + start_idx, end_idx - start_idx);   // :34
        } else {
            signal.sample(interval, buffer.data() + start_idx, buffer.size() - start_idx);  // :36
            signal.sample(interval, buffer.data(), end_idx);                                // :37
        }
        write = end;                                                                   // :40
    }
    void delay(uint32_t rate, float dt) { write = fmodf(write + (float)rate * dt, (float)buffer.size()); }  // :45-47
    void sample(uint32_t rate, float t, float interval, float* out, size_t n) const {  // ring.rs:51-79
        const size_t len = buffer.size();
        float offset = rem_euclid(write + t * (float)rate, (float)len);                // :57
        float ds = interval * (float)rate;                                             // :58
        for (size_t i = 0; i < n; i++) {
            size_t trunc = (size_t)offset;                                             // :60
            float fract = offset - (float)trunc;                                       // :61
            size_t x = trunc;
            float a, b;
            if (x < len - 1) { a = buffer[x]; b = buffer[x + 1]; }                     // :63-64
            else if (x < len) { a = buffer[x]; b = buffer[0]; }                        // :65-66
            else {                                                                      // :67-75
                x = x % len;
                offset = (float)x + fract;
                if (x < len - 1) { a = buffer[x]; b = buffer[x + 1]; }
                else { a = buffer[x]; b = buffer[0]; }
            }
            out[i] = lerp1(a, b, fract);                                               // :76
            offset += ds;                                                              // :77
        }
    }
};

// ---------------------------------------------------------------------------------------
// spatial.rs
static const float POSITION_SMOOTHING_PERIOD = 0.5f;  // spatial.rs:520
static const float SPEED_OF_SOUND = 343.0f;           // spatial.rs:602
static const float HEAD_RADIUS = 0.1075f;             // spatial.rs:605

struct Motion { Vec3 position; Vec3 velocity; bool discontinuity; };  // spatial.rs:480-484

// swap.rs semantics seen from the consumer: latest flushed value wins, refresh() reports
// whether something new arrived since the last refresh (swap.rs:36-68).
template <class T>
struct Swap {
    T received_; T pending_; bool fresh;
    explicit Swap(const T& init) : received_(init), pending_(init), fresh(false) {}
    void send(const T& v) { pending_ = v; fresh = true; }      // pending() + flush()
    bool refresh() { if (!fresh) return false; received_ = pending_; fresh = false; return true; }
    const T& received() const { return received_; }
};

struct State {                                         // spatial.rs:486-512
    Vec3 prev_position; float dt;
    Vec3 smoothed_position(float d, const Motion& next) const {   // :501-511
        float dt2 = dt + d;
        Vec3 position_change = scale(next.velocity, dt2);
        Vec3 naive_position = add(prev_position, position_change);
        Vec3 intended_position = add(next.position, position_change);
        return mix(naive_position, intended_position, fminf(dt2 / POSITION_SMOOTHING_PERIOD, 1.0f));
    }
};

enum Ear { Left = 0, Right = 1 };
static inline Vec3 ear_pos(Ear e) { return {e == Left ? -HEAD_RADIUS : HEAD_RADIUS, 0.0f, 0.0f}; }  // :573-583
static inline Vec3 ear_dir(Ear e) {                                                                  // :586-598
    return {(e == Left ? -1.0f : 1.0f) * 4.0f / sqrtf(17.0f), 0.0f, -1.0f / sqrtf(17.0f)};
}
struct EarState { float offset, gain; };
static inline EarState ear_state(Vec3 p, Ear ear, float radius) {  // spatial.rs:531-549
    float distance = norm(sub(p, ear_pos(ear)));
    float offset = distance * (-1.0f / SPEED_OF_SOUND);
    float distance_gain = radius / fmaxf(distance, radius);
    float stereo_gain = 0.5f + (distance < 1e-3f ? 0.5f : dot(ear_dir(ear), scale(p, 0.5f / distance)));
    return {offset, stereo_gain * distance_gain};
}

struct Common {                                        // spatial.rs:84-117
    float radius;
    Swap<Motion> motion;
    State state;
    bool has_finished_for; float finished_for;
    std::shared_ptr<bool> stopped;
    Common(float r, Vec3 pos, Vec3 vel)
        : radius(r), motion(Motion{pos, vel, false}), state{pos, 0.0f}, has_finished_for(false), finished_for(0.0f),
          stopped(std::make_shared<bool>(false)) {}
};

struct SpatialHandle {                                 // spatial.rs:120-157 (`Spatial`)
    std::shared_ptr<Common> common;  // control side only touches motion.send and *stopped
    void set_motion(Vec3 p, Vec3 v, bool disc) { common->motion.send(Motion{p, v, disc}); }  // :137-149
    bool is_finished() const { return *common->stopped; }                                    // :154-156
};

struct SpatialSignal { std::shared_ptr<Common> common; SignalP inner; };                     // :60-63
struct SpatialSignalBuffered {                                                               // :18-28
    uint32_t rate; float max_delay; std::shared_ptr<Common> common; Ring queue; SignalP inner;
};

// One per-source record of what the last SpatialScene::sample call did, for cursor parity.
struct SpatialTrace {
    Vec3 prev_position, next_position;
    EarState prev_ear[2], next_ear[2];
    float dt[2], d_gain[2];
};

struct SpatialScene : Signal {                         // spatial.rs:160-189
    Swap<Quat> rot;
    Set<SpatialSignalBuffered> recv_buffered;
    Set<SpatialSignal> recv;
    std::vector<double> out64;  // checker aid, see Mixer::out64
    SpatialScene() : rot(Quat{1.0f, {0.0f, 0.0f, 0.0f}}) {}
    int channels() const override { return 2; }

    SpatialHandle play(SignalP s, Vec3 pos, Vec3 vel, float radius) {                        // :289-302
        auto c = std::make_shared<Common>(radius, pos, vel);
        recv.insert(SpatialSignal{c, s});
        return SpatialHandle{c};
    }
    SpatialHandle play_buffered(SignalP s, Vec3 pos, Vec3 vel, float radius, float max_distance, uint32_t rate,
                                float buffer_duration) {                                      // :314-340, :30-56
        float max_delay = max_distance / SPEED_OF_SOUND + buffer_duration;                   // :330
        Ring q((size_t)ceilf(max_delay * (float)rate) + 1);                                  // :39
        q.delay(rate, fminf(norm(pos) / SPEED_OF_SOUND, max_delay));                         // :40-43
        auto c = std::make_shared<Common>(radius, pos, vel);
        recv_buffered.insert(SpatialSignalBuffered{rate, max_delay, c, std::move(q), s});
        return SpatialHandle{c};
    }
    void set_listener_rotation(const Quat& q) { rot.send(invert_quat(q)); }                  // :345-349

    // spatial.rs:191-265. Returns false if the source was removed (no mixing this call).
    template <class T>
    bool walk_one(Set<T>& set, size_t i, const Quat& prev_rot, const Quat& rotq, float elapsed, Vec3& prev_position,
                  Vec3& next_position) {
        T& signal = set.signals[i];
        Common& common = *signal.common;
        State& state = common.state;
        Motion orig_next = common.motion.received();                                         // :216
        if (common.motion.refresh()) {                                                       // :217-223
            state.prev_position = common.motion.received().discontinuity ? common.motion.received().position
                                                                         : state.smoothed_position(0.0f, orig_next);
            state.dt = 0.0f;
        }
        prev_position = rotate(prev_rot, state.smoothed_position(0.0f, common.motion.received()));     // :228-231
        next_position = rotate(rotq, state.smoothed_position(elapsed, common.motion.received()));      // :232-235
        state.dt += elapsed;                                                                 // :238
        float distance = norm(prev_position);                                                // :243
        if (common.has_finished_for) {                                                       // :244-251
            if (common.finished_for > distance / SPEED_OF_SOUND) *common.stopped = true;
            else common.finished_for = common.finished_for + elapsed;
        } else if (signal.inner->is_finished()) {                                            // :252-256
            common.has_finished_for = true; common.finished_for = elapsed;
        }
        if (*common.stopped) { set.remove(i); return false; }                               // :258-261
        return true;
    }

    void sample(float interval, float* out, size_t n) override {                             // spatial.rs:376-471
        recv_buffered.update();                                                              // :379
        Quat prev_rot = rot.received(); rot.refresh(); Quat rotq = rot.received();           // :382-386
        for (size_t i = 0; i < 2 * n; i++) out[i] = 0.0f;                                    // :389-391
        out64.assign(2 * n, 0.0);
        float buf[256];                                                                      // :393
        float elapsed = interval * (float)n;                                                 // :394
        const float nf = (float)n;

        // buffered set, :395-433 (walk_set calls set.update() again, :203 — a no-op here)
        recv_buffered.update();
        for (size_t si = recv_buffered.len(); si-- > 0;) {
            Vec3 prev_position, next_position;
            if (!walk_one(recv_buffered, si, prev_rot, rotq, elapsed, prev_position, next_position)) continue;
            SpatialSignalBuffered& signal = recv_buffered.signals[si];
            signal.queue.write_from(*signal.inner, signal.rate, elapsed);                    // :406
            for (int e = 0; e < 2; e++) {                                                    // :409
                EarState prev_state = ear_state(prev_position, (Ear)e, signal.common->radius);
                EarState next_state = ear_state(next_position, (Ear)e, signal.common->radius);
                float prev_offset = fmaxf(prev_state.offset - elapsed, -signal.max_delay);   // :414
                float next_offset = fmaxf(next_state.offset, -signal.max_delay);             // :415
                float dt = (next_offset - prev_offset) / nf;                                 // :417
                float d_gain = (next_state.gain - prev_state.gain) / nf;                     // :418
                size_t i = 0;
                for (size_t c0 = 0; c0 < n; c0 += 256) {                                     // :422
                    size_t m = n - c0; if (m > 256) m = 256;
                    float t = prev_offset + (float)i * dt;                                   // :423
                    signal.queue.sample(signal.rate, t, dt, buf, m);                         // :424
                    for (size_t k = 0; k < m; k++) {                                         // :425-429
                        float gain = prev_state.gain + (float)i * d_gain;
                        float contrib = buf[k] * gain;
                        out[(c0 + k) * 2 + e] = out[(c0 + k) * 2 + e] + contrib;
                        out64[(c0 + k) * 2 + e] += (double)contrib;
                        i++;
                    }
                }
            }
        }

        // seek set, :435-470
        recv.update();                                                                       // :437
        for (size_t si = recv.len(); si-- > 0;) {
            Vec3 prev_position, next_position;
            if (!walk_one(recv, si, prev_rot, rotq, elapsed, prev_position, next_position)) continue;
            SpatialSignal& signal = recv.signals[si];
            for (int e = 0; e < 2; e++) {                                                    // :446
                EarState prev_state = ear_state(prev_position, (Ear)e, signal.common->radius);
                EarState next_state = ear_state(next_position, (Ear)e, signal.common->radius);
                signal.inner->seek(prev_state.offset);                                       // :449
                float effective_elapsed = (elapsed + next_state.offset) - prev_state.offset; // :451
                float dt = effective_elapsed / nf;                                           // :452
                float d_gain = (next_state.gain - prev_state.gain) / nf;                     // :453
                size_t i = 0;
                for (size_t c0 = 0; c0 < n; c0 += 256) {                                     // :456
                    size_t m = n - c0; if (m > 256) m = 256;
                    signal.inner->sample(dt, buf, m);                                        // :457
                    for (size_t k = 0; k < m; k++) {                                         // :458-462
                        float gain = prev_state.gain + (float)i * d_gain;
                        float contrib = buf[k] * gain;
                        out[(c0 + k) * 2 + e] = out[(c0 + k) * 2 + e] + contrib;
                        out64[(c0 + k) * 2 + e] += (double)contrib;
                        i++;
                    }
                }
                signal.inner->seek(-effective_elapsed - prev_state.offset);                  // :465
            }
            signal.inner->seek(elapsed);                                                     // :468
        }
    }
};

// lib.rs:90-93
static inline void run(Signal& signal, uint32_t sample_rate, float* out, size_t n) {
    float interval = 1.0f / (float)sample_rate;
    signal.sample(interval, out, n);
}

}  // namespace orc
