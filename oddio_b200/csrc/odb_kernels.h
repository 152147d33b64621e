// Host-callable launchers of the oddio_b200 kernels (implemented in the .cu files).
#pragma once
#include <cuda_runtime.h>

#include "odb_types.h"
#include "odb_exchange.h"

// Largest per-frame source advance (samples per output frame) the fast mix kernel stages in
// shared memory; sources outside (0, ODB_FAST_DS_MAX] take the general kernel.
#define ODB_FAST_DS_MAX 2.0f
// Floats of PCM the staged kernel can hold per source and 512-frame half tile (two such buffers per warp): a
// full half fits up to ds = 1.17 (512 * 1.17 + ear skew + slack), shorter callbacks up to ODB_FAST_DS_MAX.
#define ODB_FAST_PCM_CAP 640
#define ODB_FAST_HALF_CHUNKS 2
// Floats of PCM the staged mixer resampling kernel holds per (source, 1024-frame chunk): mono up to ds = 2.0,
// stereo up to ds = 1.0 (two such buffers per warp).
#define ODB_MIXER_RESAMPLE_CAP 2112
#define ODB_MAGIC 8388608.0f          // 2^23: ulp 1, so x +rd 2^23 = 2^23 + floor(x)
#define ODB_MAGIC_BITS 0x4B000000u
// Launches `kernel` so that it may overlap the tail of the previous kernel in `st` (programmatic dependent
// launch); the kernel must call odbk::pdl_wait() before it touches anything the previous kernel wrote.
template <class... KArgs, class... Args>
static inline cudaError_t odb_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

struct OdbMotionMsg {  // Spatial::set_motion payload (spatial.rs:137-149)
    uint32_t slot;
    float pos[3];
    float vel[3];
    uint32_t discontinuity;
};
#define ODB_PARAM_SPEED 1u
#define ODB_PARAM_GAIN 2u
#define ODB_PARAM_STOP 3u
struct OdbParamMsg {  // SpeedControl::set_speed / GainControl::set_amplitude_ratio / Mixed::stop
    uint32_t slot;
    uint32_t what;
    float value;
    uint32_t pad;
};

// Internal bit of the reduce kernel's epilogue argument: store the tile as 16-bit PCM (offline render) instead of f32.
#define ODB_EPILOGUE_I16_BIT 0x100
void odb_launch_convert_i16(const short* in, float* out, size_t n, float max_value, cudaStream_t st);
void odb_launch_scatter_sources(OdbSource* src, const OdbSource* staged, const uint32_t* slots, int n, cudaStream_t st);
void odb_launch_scatter_motion(OdbSource* src, const OdbMotionMsg* msgs, int n, cudaStream_t st);
void odb_launch_scatter_params(OdbSource* src, const OdbParamMsg* msgs, int n, cudaStream_t st);
// `walked` != NULL: the one-launch callback's walk (no waits; every block counts itself into `walked` when it is done).
// Returns the number of blocks launched.
int odb_launch_walk_seek(OdbSource* src, const uint32_t* order, OdbJob* jobs, uint32_t* removed, int removed_cap,
                         uint32_t* counters, uint32_t* zero_counters, unsigned long long* walked, const OdbCallback& cb,
                         cudaStream_t st);
int odb_mix_general_ctas(int n_sources, int sm_count);
cudaError_t odb_launch_mix_general(const OdbJob* jobs, int n_sources, int n_tiles, float* partials, int n_ctas,
                                   int only_flagged, const uint32_t* counters, cudaStream_t st);
// Arguments of the one-launch scene callback kernel (odb_scene_mix.cu).
struct OdbSceneMixArgs {
    const OdbJob* jobs;               // [n_tiles][n_sources], written by the walk kernel
    int n_sources, n_tiles, n_frames;
    int batch;                        // sources per batch (odb_scene_mix_shape)
    int epilogue;                     // 0 none, 1 Tanh, 2 Reinhard, | ODB_EPILOGUE_I16_BIT
    float* partials;                  // [n_tiles][grid][2 * ODB_TILE_FRAMES]
    float* out;                       // interleaved stereo, device or pinned host memory (f32, or int16 with the I16 bit)
    unsigned long long* arrive;       // monotonic count of (CTA, tile) arrivals of this callback parity
    unsigned long long arrive_base;   // its value before this launch
    unsigned long long* done;         // monotonic count of finished CTAs of this callback parity
    unsigned long long done_base;
    const unsigned long long* walked; // walk blocks finished so far (all callbacks); this launch's job records are complete ...
    unsigned long long walked_target; // ... when it reaches this
    unsigned long long* completed;    // sequence number of the last callback kernel that has finished entirely
    unsigned long long my_seq;        // this launch's sequence number; it starts once `completed` >= my_seq - 2
    unsigned long long* host_flag;    // pinned host word that receives `seq` when the whole grid has stored its output
    unsigned long long seq;
    const uint32_t* removed_count;    // optional: the walk kernel's removal-report count ...
    uint32_t* removed_count_host;     // ... copied to this pinned host word before host_flag is raised
    unsigned long long nz;            // (-0.0, -0.0), see odb_f32x2.cuh mulx
    // Multi-GPU (odb_exchange.h; world <= 1: none of this is used). push_seq != 0: the grid's sum goes, without the
    // epilogue, into slot `rank` of every rank's inbox as exchange push_seq instead of `out`. pull_seq != 0: the
    // reduce phase also sums exchange pull_seq (this callback's, or an earlier one's) over the ranks in rank order,
    // applies the epilogue and stores that into `out`.
    odbk::ExchangePeers peers;
    odbk::ExchangeGeom xg;
    uint32_t push_seq, pull_seq;
    float* xtile;                     // [n_tiles][2 * ODB_TILE_FRAMES] raw sum of this rank, for the grid's last CTA (exchange / host tile)
};
void odb_scene_mix_shape(int n_sources, int sm_count, int* batch, int* ctas);
cudaError_t odb_launch_scene_mix(const OdbSceneMixArgs& args, int n_ctas, int mode, cudaStream_t st);
int odb_mix_fast_ctas(int n_sources, int sm_count);
cudaError_t odb_launch_mix_fast(const OdbJob* jobs, int n_sources, int n_tiles, float* partials, int n_ctas, int mode,
                                cudaStream_t st);
void odb_launch_walk_mixer(OdbSource* src, const uint32_t* order, OdbMixJob* jobs, uint32_t* removed, int removed_cap,
                           uint32_t* counters, const OdbCallback& cb, cudaStream_t st);
int odb_mixer_ctas(int n_sources, int sm_count, int per_sm);
cudaError_t odb_launch_mixer_unit(const OdbMixJob* jobs, int n_sources, int n_tiles, int channels, float* partials,
                                  int n_ctas, cudaStream_t st);
int odb_mixer_resample_ctas(int n_sources, int sm_count);
cudaError_t odb_launch_mixer_resample(const OdbMixJob* jobs, int n_sources, int n_tiles, int channels, float* partials,
                                      int n_ctas, const uint32_t* counters, cudaStream_t st);
cudaError_t odb_launch_mixer_general(const OdbMixJob* jobs, int n_sources, int n_tiles, int channels, float* partials,
                                     int n_ctas, int only_flagged, const uint32_t* counters, cudaStream_t st);
void odb_launch_walk_buffered(OdbSource* src, const uint32_t* order, OdbRingJob* jobs, OdbJob* fjobs, OdbRingWrite* writes,
                              uint32_t* removed, int removed_cap, uint32_t* counters, uint32_t* literal_list,
                              const OdbCallback& cb, cudaStream_t st);
void odb_launch_ring_write(const OdbRingWrite* writes, int n_sources, cudaStream_t st);
int odb_mix_ring_ctas(int n_sources, int sm_count);
cudaError_t odb_launch_mix_ring(const OdbRingJob* jobs, int n_sources, int n_tiles, float* partials, int n_ctas,
                                int only_flagged, const uint32_t* counters, const uint32_t* literal_list, cudaStream_t st);
// Sums partial tiles of `tile_floats` floats each (1024 frames x channels) into the interleaved output.
// Up to three partial sets: a (staged / streaming kernel), b (general kernel; skipped when the walk kernel
// counted no general job and b_is_general is set), c (ring kernel).
// `c_counter` >= 0: set c is skipped when that job counter is zero. `zero_counters` (may be NULL): ODB_CNT_WORDS
// counters the kernel resets for the next callback.
void odb_launch_reduce(const float* pa, int na, const float* pb, int nb, const float* pc, int nc, const uint32_t* counters,
                       int b_is_general, int c_counter, uint32_t* zero_counters, float* out, int n_frames, int n_tiles,
                       int channels, int epilogue, cudaStream_t st);
