/* oddio_b200.h — C ABI of the B200-native oddio hot path.
 *
 * This is the drop-in boundary for ONE path of Ralith/oddio 0.7.4: what `oddio::run` drives
 * through `SpatialScene` / `Mixer` (`Signal::sample` over the active `Set`). The reference has
 * no FFI of its own (it is pure Rust); the entry points below are what a Rust shim keeping the
 * reference's `Signal`/`Seek`/`Frame` traits and `SpatialScene::new`/`play`/`set_motion`/`run`
 * API binds with `extern "C"` (stub in INTEGRATION.md). Every function cites the reference item
 * (file:line under the reference's src/) it stands in for.
 *
 * Conventions
 *  - plain pointers and sizes only; opaque handles; no C++/torch types.
 *  - every function returns an `int` status: ODB_OK (0) or a negative ODB_E_* code;
 *    `odb_last_error()` returns a thread-local message for the last failure. The reference has
 *    no error returns on this path (misuse panics); codes exist because FFI cannot panic.
 *  - threading mirrors the reference: one "audio" thread calls *_sample / *_run on a scene or
 *    mixer; one control thread calls play / set_* / stop. Control calls take effect at the next
 *    *_sample boundary, latest value wins (swap.rs:36-68, set.rs:141-178).
 *  - `out` buffers of *_sample / *_run are HOST memory, interleaved frames, caller-owned.
 *    *_sample_device leaves the result in device memory on the given CUDA stream instead.
 *  - there is NO CPU fallback: without a CUDA device every call fails with ODB_E_CUDA.
 */
#ifndef ODDIO_B200_H
#define ODDIO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ODB_OK 0
#define ODB_E_INVALID (-1)     /* bad argument / unknown handle */
#define ODB_E_CUDA (-2)        /* CUDA runtime error (message has the cudaError string) */
#define ODB_E_UNSUPPORTED (-3) /* signal chain outside the closed set the device path accepts */
#define ODB_E_NOMEM (-4)

typedef struct odb_ctx odb_ctx;     /* one CUDA device + stream + PCM arena; one per process/GPU */
typedef struct odb_scene odb_scene; /* SpatialSceneControl + SpatialScene pair (spatial.rs:160-189) */
typedef struct odb_mixer odb_mixer; /* MixerControl<T> + Mixer<T> pair (mixer.rs:61-87) */
typedef struct odb_exchange odb_exchange; /* multi-GPU sum of the mixed tile over NVLink peer memory (no reference counterpart) */
typedef uint64_t odb_frames;        /* Arc<Frames<T>> (frames.rs:16-22); 0 is never valid */
typedef uint64_t odb_source;        /* a playing signal: Spatial / Mixed + its inner controls */

const char* odb_last_error(void);
/* ABI version of this header; bumped on any incompatible change. */
uint32_t odb_abi_version(void);

/* ---- context ------------------------------------------------------------------------------- */
int odb_ctx_create(int cuda_device, odb_ctx** out);
/* Same, but all work is queued on the caller's CUDA stream (a cudaStream_t; NULL = the legacy
 * default stream) instead of a private one, so it orders with the caller's own kernels and
 * collectives (e.g. the NCCL reduce of per-GPU tiles) without extra events. */
int odb_ctx_create_on_stream(int cuda_device, void* cuda_stream, odb_ctx** out);
int odb_ctx_destroy(odb_ctx* ctx);
/* Blocks until all work queued on the context's stream has finished. */
int odb_ctx_synchronize(odb_ctx* ctx);
/* The context's CUDA stream (a cudaStream_t) so callers can order their own work after it. */
int odb_ctx_stream(odb_ctx* ctx, void** out_stream);

/* Page-locks (cudaHostRegister) / releases a host buffer of the caller, so that a host without the CUDA runtime can make
 * its output tile directly writable by the device: `odb_scene_sample` renders straight into an `out` that is pinned
 * or registered (no staging tile, no memcpy on the way back); any other `out` goes through the scene's own pinned tile.
 * Optional; the buffer must stay valid until it is unpinned. No reference counterpart (the reference mixes into
 * ordinary memory, lib.rs:90). */
int odb_pin_buffer(odb_ctx* ctx, void* host_ptr, uint64_t bytes);
int odb_unpin_buffer(odb_ctx* ctx, void* host_ptr);

/* ---- Frames (frames.rs:19-77) ------------------------------------------------------------------ */
/* Frames::from_slice (frames.rs:26-47): copies `n_frames` interleaved frames of `channels`
 * (1 = Sample, 2 = [Sample; 2]) from HOST memory into HBM. */
int odb_frames_from_slice(odb_ctx* ctx, uint32_t rate, int channels, const float* samples, uint64_t n_frames,
                          odb_frames* out);
/* Same, but `dev_samples` already is DEVICE memory on ctx's device (copied device-to-device into
 * the arena, so the caller may free it afterwards). For PCM decoded or synthesised on the GPU. */
int odb_frames_from_device(odb_ctx* ctx, uint32_t rate, int channels, const void* dev_samples, uint64_t n_frames,
                           odb_frames* out);
/* PCM ingest of integer WAV data (examples/wav.rs:30-46): `n_frames` interleaved frames of 16-bit containers holding
 * `bits_per_sample`-bit signed samples (2..16) are uploaded as they are and scaled on the device exactly as the
 * example does on the CPU, `sample as f32 / (2^(bits-1) - 1) as f32`, into an f32 Frames block. */
int odb_frames_from_i16(odb_ctx* ctx, uint32_t rate, int channels, const int16_t* samples, uint64_t n_frames,
                        int bits_per_sample, odb_frames* out);
/* Drops one reference (Arc drop). Storage is freed once no playing source uses it. */
int odb_frames_release(odb_ctx* ctx, odb_frames frames);

/* ---- the closed set of signal chains the device path accepts ------------------ */
/* A chain is  [Gain]( [FixedGain]( [Speed]( FramesSignal ) ) )  with each bracket optional:
 *   FramesSignal::new(frames, start_seconds)            frames.rs:156-169
 *   Speed::new(..)   + SpeedControl::set_speed          speed.rs:16-23, :52-54
 *   FixedGain::new(.., db)                              gain.rs:18-23
 *   Gain::new(..)    + Gain::set_amplitude_ratio        gain.rs:66-93
 * SpatialSceneControl::play requires `Seek`, which Speed and Gain do not implement
 * (speed.rs:26-40, gain.rs:95-127), so odb_scene_play rejects those flags with ODB_E_UNSUPPORTED;
 * play_buffered and the mixer accept all of them. */
#define ODB_CHAIN_SPEED 0x1u
#define ODB_CHAIN_FIXED_GAIN 0x2u
#define ODB_CHAIN_GAIN 0x4u
#define ODB_CHAIN_CYCLE 0x8u /* the innermost signal is Cycle<T> (cycle.rs:6-61) instead of FramesSignal<T>: `start_seconds`
                              * then holds the initial cursor in SAMPLES (0 after Cycle::new; whatever Seek::seek calls made
                              * before play left, cycle.rs:57-60). Plays under odb_mixer_play and odb_scene_play (Cycle is Seek); odb_scene_play_buffered
                              * answers ODB_E_UNSUPPORTED. Cycle sources take the literal kernels. */
typedef struct odb_chain {
    odb_frames frames;     /* the Arc<Frames<T>> played */
    double start_seconds;  /* FramesSignal::new start_seconds, may be negative */
    uint32_t flags;        /* ODB_CHAIN_* */
    float speed;           /* initial SpeedControl value (Speed::new starts at 1.0) */
    float fixed_gain_db;   /* FixedGain::new db */
    float gain_ratio;      /* Gain::set_amplitude_ratio initial factor (Gain::new starts at 1.0) */
} odb_chain;

/* Post-mix wrappers around the whole aggregator: Tanh<T> (tanh.rs:22-29), Reinhard<T> (reinhard.rs:28-35) */
#define ODB_EPILOGUE_NONE 0
#define ODB_EPILOGUE_TANH 1
#define ODB_EPILOGUE_REINHARD 2

/* ---- SpatialScene (spatial.rs) -------------------------------------------------------------------- */
/* SpatialScene::new (spatial.rs:170-188) */
int odb_scene_create(odb_ctx* ctx, odb_scene** out);
int odb_scene_destroy(odb_scene* scene);
/* Wrap the scene in Tanh / Reinhard: `Tanh::new(scene)`. */
int odb_scene_set_epilogue(odb_scene* scene, int epilogue);
/* SpatialSceneControl::play (spatial.rs:289-302) with SpatialOptions {position, velocity, radius} (:354-371) */
int odb_scene_play(odb_scene* scene, const odb_chain* chain, const float position[3], const float velocity[3],
                   float radius, odb_source* out);
/* SpatialSceneControl::play_buffered (spatial.rs:314-340) */
int odb_scene_play_buffered(odb_scene* scene, const odb_chain* chain, const float position[3],
                            const float velocity[3], float radius, float max_distance, uint32_t rate,
                            float buffer_duration, odb_source* out);
/* SpatialSceneControl::set_listener_rotation (spatial.rs:345-349); q = mint::Quaternion as {x, y, z, s} */
int odb_scene_set_listener_rotation(odb_scene* scene, const float q_xyzs[4]);
/* Spatial::set_motion (spatial.rs:137-149) */
int odb_spatial_set_motion(odb_scene* scene, odb_source src, const float position[3], const float velocity[3],
                           int discontinuity);
/* Spatial::set_motion for `n` sources in one FFI call (same semantics, applied in array order; one foreign
 * call per moving source per callback is what the reference's cheap Rust method call would become).
 * positions / velocities are n x 3 floats, discontinuity n bytes (may be NULL = all false). */
int odb_spatial_set_motion_many(odb_scene* scene, uint32_t n, const odb_source* srcs, const float* positions,
                                const float* velocities, const uint8_t* discontinuity);
/* Spatial::is_finished (spatial.rs:154-156) */
int odb_spatial_is_finished(odb_scene* scene, odb_source src, int* out);
/* <SpatialScene as Signal>::sample (spatial.rs:376-471): n_frames stereo frames, interleaved L,R */
int odb_scene_sample(odb_scene* scene, float interval, float* out, uint32_t n_frames);
/* Offline render (examples/offline.rs:33-43): as odb_scene_sample, with the tile quantised on the device as the
 * example does before it writes the WAV file, `(sample * i16::MAX as f32) as i16` (toward zero, saturating). */
int odb_scene_sample_i16(odb_scene* scene, float interval, int16_t* out, uint32_t n_frames);
/* oddio::run(&mut scene, sample_rate, out) (lib.rs:90-93): interval = 1.0 / sample_rate as f32 */
int odb_scene_run(odb_scene* scene, uint32_t sample_rate, float* out, uint32_t n_frames);
/* As odb_scene_sample, but the mixed tile is left in DEVICE memory `dev_out` (2*n_frames f32),
 * ordered on ctx's stream, with no host synchronisation; for multi-GPU reduction of per-shard
 * tiles and for device-resident benchmarking. */
int odb_scene_sample_device(odb_scene* scene, float interval, void* dev_out, uint32_t n_frames);
/* Number of sources currently in the seek set / buffered set (set.rs:191-204 Deref len). */
int odb_scene_len(odb_scene* scene, int buffered, uint64_t* out);

/* ---- Mixer (mixer.rs) ----------------------------------------------------------------------------- */
/* Mixer::<T>::new (mixer.rs:70-81); channels 1 => Mixer<Sample>, 2 => Mixer<[Sample; 2]> */
int odb_mixer_create(odb_ctx* ctx, int channels, odb_mixer** out);
int odb_mixer_destroy(odb_mixer* mixer);
int odb_mixer_set_epilogue(odb_mixer* mixer, int epilogue);
/* MixerControl::play (mixer.rs:18-26); the chain's Frames must have the mixer's channel count */
int odb_mixer_play(odb_mixer* mixer, const odb_chain* chain, odb_source* out);
/* Mixed::stop / Mixed::is_stopped (mixer.rs:34-43) */
int odb_mixed_stop(odb_mixer* mixer, odb_source src);
int odb_mixed_is_stopped(odb_mixer* mixer, odb_source src, int* out);
/* <Mixer<T> as Signal>::sample (mixer.rs:92-119) and oddio::run over it */
int odb_mixer_sample(odb_mixer* mixer, float interval, float* out, uint32_t n_frames);
int odb_mixer_run(odb_mixer* mixer, uint32_t sample_rate, float* out, uint32_t n_frames);
/* Offline render: as odb_scene_sample_i16 (examples/offline.rs:39) */
int odb_mixer_sample_i16(odb_mixer* mixer, float interval, int16_t* out, uint32_t n_frames);
int odb_mixer_sample_device(odb_mixer* mixer, float interval, void* dev_out, uint32_t n_frames);
int odb_mixer_len(odb_mixer* mixer, uint64_t* out);

/* ---- per-source controls; `owner` is the odb_scene* or odb_mixer* that returned `src` ------------------ */
/* SpeedControl::set_speed / speed (speed.rs:47-54) */
int odb_source_set_speed(void* owner, odb_source src, float factor);
/* GainControl::set_amplitude_ratio / set_gain (gain.rs:143-159) */
int odb_source_set_amplitude_ratio(void* owner, odb_source src, float factor);
int odb_source_set_gain_db(void* owner, odb_source src, float db);
/* FramesSignalControl::playback_position / is_finished (frames.rs:238-247) */
int odb_source_playback_position(void* owner, odb_source src, double* out_seconds);
int odb_source_frames_is_finished(void* owner, odb_source src, int* out);
/* Parity aid: the FramesSignal's f64 time cursor `t` (frames.rs:145) and, for buffered sources,
 * the Ring's f32 write cursor (ring.rs:6). Bit-exact against the reference by contract. */
int odb_source_cursor(void* owner, odb_source src, double* out_t, float* out_ring_write);

/* ---- introspection for tests / profiling ------------------------------------------------------------------ */
/* Number of kernels launched by the last *_sample* call on this owner. */
int odb_last_launch_count(void* owner, uint32_t* out);
/* Per-callback job counters of the last *_sample* call on this owner, out[0] = (source, tile) jobs that
 * took the literal general kernel, out[1] = jobs that took the staged / streaming kernel, out[2] = jobs that took
 * the staged resampling kernel (mixer), out[3] = buffered-source jobs that took the literal ring kernel.
 * Synchronises the context's stream. Benchmarks assert out[0] == 0. */
int odb_last_job_counters(void* owner, uint32_t out[4]);
/* Kernel timing for roofline reports: when enabled, *_sample* brackets its mix kernel with CUDA events on
 * the context's stream; odb_last_mix_kernel_ms then synchronises and returns the device time of the
 * dominant (staged mix) kernel of the last call. Off by default (events cost a few microseconds). */
int odb_set_profiling(void* owner, int enabled);
int odb_last_mix_kernel_ms(void* owner, float* out_ms);
/* Selects the mix-kernel variant: 2 = default: the three value multiply-adds (lerp, gain ramp, accumulate) contracted
 * to FMA, cursors and indices bit-exact; 0 = strict arithmetic: every value operation unfused in the reference's
 * order, so a source's contribution is bit-identical to the reference's; 1 = force the literal path for every source
 * (slow, used to cross-check). Adding 0x100 runs a scene's per-source set-up kernels on a second stream; adding 0x200
 * selects the multi-kernel callback (walk, staged mix, literal mix, reduce as separate launches) instead of the
 * one-launch callback kernel. */
int odb_set_kernel_variant(void* owner, int variant);

/* ---- multi-GPU: sum of the per-GPU tiles over NVLink peer memory ---------------------------------------
 * The reference is single-process (one `Signal` graph, signal.rs:19); when its sources are sharded over the
 * GPUs of one box (one process per GPU, each with its own scene/mixer over its shard), the only exchange is
 * the additive output tile that `SpatialScene::sample` (spatial.rs:376-471) / `Mixer::sample` (mixer.rs:92-119)
 * accumulate into. These entry points do that exchange without a collective library: every rank pushes its
 * tile into every rank's inbox with stores over NVLink and sums the inbox in rank order (bit-identical result
 * on all ranks), in one kernel per callback. Set-up: create on every rank, export the 64-byte handle, gather
 * the handles of all ranks by any host-side means (rank order), connect. */
/* `depth` (2..8): how many pushed exchanges may await their pull (inbox slots per rank). */
int odb_exchange_create(odb_ctx* ctx, int rank, int world, uint32_t max_floats, int depth, odb_exchange** out);
int odb_exchange_destroy(odb_exchange* ex);
/* Bytes of one exported handle (a cudaIpcMemHandle_t). */
int odb_exchange_handle_size(void);
int odb_exchange_export(odb_exchange* ex, void* handle_out);
/* `handles`: world x odb_exchange_handle_size() bytes in rank order (the own entry is ignored). */
int odb_exchange_connect(odb_exchange* ex, const void* handles);
/* In-place sum over the ranks of `dev_tile` (n_floats f32 in device memory, 16-byte aligned), then the
 * epilogue (ODB_EPILOGUE_*: Tanh / Reinhard act on the sum, tanh.rs:22-29, reinhard.rs:28-35), queued on
 * `cuda_stream` (NULL = the context's stream). Every rank must call it once per callback, in the same order.
 * The shards' own scenes/mixers run with ODB_EPILOGUE_NONE. */
int odb_exchange_allreduce(odb_exchange* ex, void* dev_tile, uint32_t n_floats, int epilogue, void* cuda_stream);
/* The two halves of odb_exchange_allreduce for a pipelined renderer: push sends this rank's tile to every rank's
 * inbox and never waits for a peer's data; pull (same n_floats, in push order, at most `depth` pushes outstanding)
 * leaves the sum over the ranks in `dev_tile` (which need not be the pushed buffer). Queuing the pull one callback
 * group later keeps every rank's GPU busy with the next mixes instead of waiting for the slowest rank. */
int odb_exchange_push(odb_exchange* ex, const void* dev_tile, uint32_t n_floats, void* cuda_stream);
int odb_exchange_pull(odb_exchange* ex, void* dev_tile, uint32_t n_floats, int epilogue, void* cuda_stream);
/* `SpatialScene::sample` (spatial.rs:376-471) of one rank's shard with the exchange folded into the callback kernel:
 * its reduce phase stores the rank's sum into every rank's inbox over NVLink (no separate push launch, no local
 * tile) and, once more than `lag` exchanges are outstanding, leaves the oldest one - summed over the ranks in rank
 * order, `epilogue` applied - in `dev_out` (*out_written = 1). lag = 0: this callback's own sum (live playback);
 * lag >= 1 (< depth): callback k - lag while callback k is mixed; the last `lag` tiles are collected with
 * odb_exchange_pull. Scenes with buffered sources take the multi-kernel path and the stand-alone exchange kernels,
 * with the same result. */
int odb_scene_sample_exchange(odb_scene* scene, odb_exchange* ex, float interval, void* dev_out, uint32_t n_frames,
                              int lag, int epilogue, int* out_written);

#ifdef __cplusplus
}
#endif
#endif /* ODDIO_B200_H */
