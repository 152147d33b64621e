// Device/host shared data layout of the oddio_b200 hot path. See DESIGN.md "Data layout in HBM".
#pragma once
#include <stdint.h>

#define ODB_SPATIAL_CHUNK 256   // spatial.rs:393 `let mut buf = [0.0; 256]`
#define ODB_MIXER_CHUNK 1024    // mixer.rs:77 staging buffer
#define ODB_TILE_FRAMES 1024    // output frames one mix-kernel pass accumulates in registers
#define ODB_TILE_CHUNKS (ODB_TILE_FRAMES / ODB_SPATIAL_CHUNK)
#define ODB_PCM_PAD 64          // zero floats kept before and after every Frames block in the arena

// source flags
#define ODB_SF_STOPPED 0x1u          // Common::stopped (spatial.rs:90) / MixedSignal::stop (mixer.rs:47)
#define ODB_SF_HAS_FINISHED_FOR 0x2u // Common::finished_for is Some (spatial.rs:89)
#define ODB_SF_MOTION_FRESH 0x4u     // swap::Receiver::refresh() would return true (swap.rs:57-64)
#define ODB_SF_PENDING_DISC 0x8u     // pending Motion::discontinuity
#define ODB_SF_MOTION_DISC 0x10u     // received Motion::discontinuity
#define ODB_SF_FIXED_GAIN 0x20u      // chain has FixedGain
#define ODB_SF_SPEED 0x40u           // chain has Speed
#define ODB_SF_GAIN 0x80u            // chain has Gain
#define ODB_SF_STOP_REQ 0x100u       // Mixed::stop() requested from the control side
#define ODB_SF_CYCLE 0x200u          // the innermost signal is Cycle (cycle.rs:6-61): `t` is its cursor in SAMPLES

// One playing source: SpatialSignal<FramesSignal chain> (spatial.rs:60-63) or MixedSignal (mixer.rs:46-49).
// 160 bytes, 16-byte aligned; an array of these lives in HBM, indexed by slot.
struct __attribute__((aligned(16))) OdbSource {
    const float* pcm;      // first sample of the Frames block (ODB_PCM_PAD zeros on both sides)
    double rate;           // Frames::rate (frames.rs:20)
    double t;              // FramesSignal::t (frames.rs:145)
    double t_end;          // (len - 1) as f64 / rate, the right-hand side of FramesSignal::is_finished (frames.rs:205)
    long long sample_t;    // FramesSignal::sample_t (frames.rs:149)
    int len;               // frames in the Frames block
    int channels;          // 1 or 2
    uint32_t flags;        // ODB_SF_*
    float radius;          // Common::radius
    float finished_for;    // Common::finished_for payload
    float state_dt;        // State::dt (spatial.rs:490)
    float prev_position[3];// State::prev_position (spatial.rs:488)
    float pos[3];          // received Motion (spatial.rs:480-484)
    float vel[3];
    float ppos[3];         // pending Motion (written by set_motion, consumed at the next sample)
    float pvel[3];
    float speed;           // Speed::speed (speed.rs:10)
    float fixed_gain;      // FixedGain::gain = 10^(db/20) (gain.rs:20)
    float gain_shared;     // Gain::shared (gain.rs:59)
    float gain_prev;       // Smoothed::prev / next / progress (smooth.rs:27-31)
    float gain_next;
    float gain_progress;
    // buffered (play_buffered) sources only
    float* ring;           // Ring::buffer (ring.rs:5)
    int ring_cap;
    float ring_write;      // Ring::write (ring.rs:6)
    float max_delay;       // SpatialSignalBuffered::max_delay
    uint32_t ring_rate;    // SpatialSignalBuffered::rate
};

// What one (source, 1024-frame tile) pass of the spatial mix kernels needs; written by the walk
// kernel each callback. Exactly one 128-byte line: a warp reads it as 32 words, lane l = word l.
struct __attribute__((aligned(16))) OdbJob {
    const float* pcm;               // words 0-1
    int len;                        // word 2
    uint32_t flags;                 // word 3: ODB_JF_*
    float ds[2];                    // words 4-5: per ear dt * rate as f32 (frames.rs:178)
    float pg[2];                    // words 6-7: prev_state.gain (spatial.rs:459)
    float dg[2];                    // words 8-9: d_gain (spatial.rs:453)
    float fixed_gain;               // word 10
    int n_frames;                   // word 11: frames of this tile (<= ODB_TILE_FRAMES)
    int base[2][ODB_TILE_CHUNKS];   // words 12-19: per ear, per 256-chunk `base` (frames.rs:179), saturated to int32
    float off0[2][ODB_TILE_CHUNKS]; // words 20-27: initial `offset` / constant `fract` (frames.rs:183,189)
    int window[2][2];               // words 28-31: per 512-frame half {first PCM index it can touch (multiple of 4),
                                    // floats from there covering every index of both ears (multiple of 4; 0 = unused)}
};
#define ODB_JW_PCM_LO 0
#define ODB_JW_PCM_HI 1
#define ODB_JW_LEN 2
#define ODB_JW_FLAGS 3
#define ODB_JW_DS 4
#define ODB_JW_PG 6
#define ODB_JW_DG 8
#define ODB_JW_FIXED_GAIN 10
#define ODB_JW_N_FRAMES 11
#define ODB_JW_BASE 12
#define ODB_JW_OFF0 20
#define ODB_JW_WINDOW 28
// What one (source, 1024-frame chunk) pass of the mixer kernels needs. 64 bytes.
struct __attribute__((aligned(16))) OdbMixJob {
    const float* pcm;
    int len;          // frames
    uint32_t flags;   // ODB_JF_SKIP / ODB_JF_FAST_L (ds ~= 1 path) / ODB_JF_GENERAL / ODB_JF_RAMP
    int base;         // frames.rs:179, saturated
    float off0;       // constant fract (ds ~= 1) or initial offset
    float ds;         // (interval * speed) * rate as f32
    float fixed_gain; // FixedGain::gain, 1.0 if absent
    float g;          // Gain at rest: Smoothed::get() with progress == 1; 1.0 if absent
    float gprev, gnext, gprog, gstep;  // Gain mid-transition (ODB_JF_RAMP): Smoothed state at the chunk start
    int n_frames;
    uint32_t pad[2];
};
static_assert(sizeof(OdbMixJob) == 64, "OdbMixJob is half a 128-byte line");
// Buffered (play_buffered) sources: what one (source, 1024-frame tile) pass of k_mix_ring needs. 128 bytes.
struct __attribute__((aligned(16))) OdbRingJob {
    const float* ring;              // Ring::buffer (ring.rs:5)
    int cap;                        // buffer.len()
    uint32_t flags;                 // ODB_JF_SKIP
    float ds[2];                    // per ear: dt * rate as f32 (ring.rs:58)
    float pg[2];                    // prev_state.gain
    float dg[2];                    // d_gain (spatial.rs:418)
    float off0[2][ODB_TILE_CHUNKS]; // per ear, per 256-chunk: (write + t * rate).rem_euclid(len) (ring.rs:57)
    int n_frames;
    uint32_t pad[13];
};
static_assert(sizeof(OdbRingJob) == 128, "OdbRingJob is one 128-byte line");
// What Ring::write (ring.rs:18-41) hands to inner.sample() for one buffered source and callback: one or two spans.
struct __attribute__((aligned(16))) OdbRingWrite {
    float* ring;
    const float* pcm;
    int cap, len;
    uint32_t flags;       // ODB_JF_SKIP / ODB_JF_FAST_L (ds ~= 1 path) / ODB_JF_RAMP (span 0) / ODB_JF_RAMP1 (span 1)
    float ds;             // (interval * speed) * pcm rate as f32, interval = 1 / ring rate
    float fixed_gain, gstep;
    int start[2], n[2], base[2];
    float off0[2];
    float g[2];
    float gprev[2], gnext[2], gprog[2];
};
#define ODB_JF_SKIP 0x1u        // source removed/stopped this callback: contributes nothing
#define ODB_JF_FAST_L 0x2u      // |ds-1| <= EPSILON for the left ear (frames.rs:180)
#define ODB_JF_FAST_R 0x4u
#define ODB_JF_FIXED_GAIN 0x8u
#define ODB_JF_RAMP 0x20u       // mixer: Gain is mid-transition during this chunk (gain.rs:118-121)
#define ODB_JF_RAMP1 0x40u      // ring write: Gain is mid-transition during the second span
#define ODB_JF_RESAMPLE 0x80u   // mixer: a resampling chain the staged resampling kernel takes (never together with GENERAL)
#define ODB_JF_RING 0x100u      // the job belongs to a buffered source: pcm is its delay ring, flagged ones go to k_mix_ring
#define ODB_JF_CYCLE 0x200u      // mixer: the innermost signal is Cycle; base/off0 are its cursor split at the chunk start (always GENERAL)
#define ODB_JF_GENERAL 0x10u    // must take the general kernel (window too large, ds <= 0, negative offset, ...)

// Device counters written by the walk kernels each callback (uint32 each).
#define ODB_CNT_GENERAL 0       // jobs flagged ODB_JF_GENERAL
#define ODB_CNT_FAST 1          // jobs the fast mix kernel takes
#define ODB_CNT_RESAMPLE 2      // mixer jobs the staged resampling kernel takes
#define ODB_CNT_RING_GENERAL 3   // buffered-source jobs the literal ring kernel takes (the read wraps around the ring)
#define ODB_CNT_WORDS 4

struct OdbQuat { float x, y, z, s; };

// Parameters of one scene callback, passed by value to the kernels.
struct OdbCallback {
    OdbQuat prev_rot, rot;   // (prev_rot, rot) of spatial.rs:382-386, already inverted (spatial.rs:346)
    float interval;
    float elapsed;           // interval * n_frames as f32 (spatial.rs:394)
    int n_frames;
    int n_tiles;
    int n_sources;           // entries of the active list
    int force_general;       // kernel variant 1: every job takes the general (literal) kernel
    int job_stride;          // OdbJob records per tile: seek sources first, then buffered sources
    int job_offset;          // index of this set's first source inside a tile's records
};

static_assert(sizeof(OdbSource) % 16 == 0, "OdbSource is moved in 16-byte words");
static_assert(sizeof(OdbJob) == 128, "OdbJob is one 128-byte line");
