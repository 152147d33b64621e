// Spatial-scene kernels: walk (per-source set-up), general mix (exact, any parameters), reduce.
// Reference: src/spatial.rs, src/frames.rs, src/frame.rs, src/math/mod.rs (cited per function).
#include <cuda_runtime.h>

#include "odb_kernels.h"
#include "odb_math.cuh"
#include "odb_walk.cuh"

namespace odbk {

// ------------------------------------------------------------------------------------------
// Control-plane scatter kernels: apply what SetHandle::insert / Spatial::set_motion queued since
// the previous callback (set.rs:141-178 drain_msgs, swap.rs:57-64 refresh).
__global__ void k_scatter_sources(OdbSource* __restrict__ src, const OdbSource* __restrict__ staged,
                                  const uint32_t* __restrict__ slots, int n) {
    // one 16-byte word per thread: sizeof(OdbSource)/16 words per source
    const int words = (int)(sizeof(OdbSource) / 16);
    int gid = blockIdx.x * blockDim.x + threadIdx.x;
    int i = gid / words, w = gid % words;
    if (i >= n) return;
    const uint4* s = reinterpret_cast<const uint4*>(staged + i);
    uint4* d = reinterpret_cast<uint4*>(src + slots[i]);
    d[w] = s[w];
}

__global__ void k_scatter_motion(OdbSource* __restrict__ src, const OdbMotionMsg* __restrict__ msgs, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // two 16-byte loads per message: the buffer may be pinned host memory read over PCIe (odb_host.cu apply())
    static_assert(sizeof(OdbMotionMsg) == 32, "a message is two 16-byte words");
    OdbMotionMsg m;
    reinterpret_cast<uint4*>(&m)[0] = __ldcs(reinterpret_cast<const uint4*>(msgs + i));
    reinterpret_cast<uint4*>(&m)[1] = __ldcs(reinterpret_cast<const uint4*>(msgs + i) + 1);
    if (m.slot == 0xFFFFFFFFu) return;  // withdrawn: its source was removed after the message was queued
    OdbSource* s = src + m.slot;
    s->ppos[0] = m.pos[0]; s->ppos[1] = m.pos[1]; s->ppos[2] = m.pos[2];
    s->pvel[0] = m.vel[0]; s->pvel[1] = m.vel[1]; s->pvel[2] = m.vel[2];
    uint32_t f = s->flags | ODB_SF_MOTION_FRESH;
    f = m.discontinuity ? (f | ODB_SF_PENDING_DISC) : (f & ~ODB_SF_PENDING_DISC);
    s->flags = f;
}

__global__ void k_scatter_params(OdbSource* __restrict__ src, const OdbParamMsg* __restrict__ msgs, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    OdbParamMsg m = msgs[i];
    if (m.slot == 0xFFFFFFFFu) return;  // withdrawn: its source was removed after the message was queued
    OdbSource* s = src + m.slot;
    if (m.what == ODB_PARAM_SPEED) s->speed = m.value;               // SpeedControl::set_speed speed.rs:52-54
    else if (m.what == ODB_PARAM_GAIN) s->gain_shared = m.value;     // GainControl::set_amplitude_ratio gain.rs:157-159
    else if (m.what == ODB_PARAM_STOP) s->flags |= ODB_SF_STOP_REQ;  // Mixed::stop mixer.rs:34-36
}

// ------------------------------------------------------------------------------------------
// walk_set for the seek set (spatial.rs:191-265) and the per-chunk set-up of the mix closure: see odb_walk.cuh.
constexpr int WALK_THREADS = 128;
// LATE: round 1's multi-kernel callback (and the buffered set's companion): waits for the previous callback's kernels
// before it touches the job counters. !LATE: the one-launch callback - no wait at all; ordering against earlier
// callbacks comes from the launch dependencies and the callback kernel's own completion counter (odb_scene_mix.cu).
template <bool LATE>
__global__ void __launch_bounds__(WALK_THREADS) k_walk_seek(OdbSource* __restrict__ src, const uint32_t* __restrict__ order,
                                                             OdbJob* __restrict__ jobs, uint32_t* __restrict__ removed,
                                                             int removed_cap, uint32_t* __restrict__ counters,
                                                             uint32_t* __restrict__ zero_counters,
                                                             unsigned long long* __restrict__ walked, OdbCallback cb) {
    __shared__ __align__(16) unsigned char smem[WalkSmem<WALK_THREADS>::BYTES];
    // Launched behind the previous callback's mix kernel, which lets its dependents start once ITS walk is complete:
    // this grid's blocks run as SMs come free, under that kernel. Control-plane scatter kernels are ordinary launches
    // and therefore complete before this grid starts.
    if (LATE) pdl_launch_dependents();  // round 1's mix kernel may be set up now; it waits for this grid (griddepcontrol.wait)
    walk_seek_block<WALK_THREADS, LATE>(src, order, jobs, removed, removed_cap, counters, zero_counters, cb,
                                        blockIdx.x * (WALK_THREADS / 2), smem, threadIdx.x);
    if (!LATE) {
        // The one-launch callback kernel does not use griddepcontrol.wait (a dependent grid's completion is ordered
        // behind its primary's, which would put the previous callback's exchange tail back on the critical path): it
        // starts when every block of this grid has got here, and reads `walked` with acquire semantics.
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            atomicAdd(walked, 1ull);
        }
        pdl_launch_dependents();
    }
}

// ------------------------------------------------------------------------------------------
// General mix kernel: exact for every parameter combination (any ds incl. <= 0, negative offsets,
// windows that leave the PCM block or exceed the staged kernel's buffers, FixedGain). One warp per
// (tile, source):
//   1. lanes 0..7 each walk one (ear, chunk) cursor chain literally, as FramesSignal::sample does
//      (frames.rs:189-196), and park all 256 cursor values of their chain in warp-private shared memory;
//   2. all 32 lanes (lane l owns frames l, l+32, ...) turn cursors into indices (`as isize` truncation)
//      and fractions, read the sample pair from HBM with get_pair's bounds rules (frames.rs:105-123),
//      lerp, apply FixedGain (gain.rs:35) and the per-frame gain (spatial.rs:459-460) and accumulate in
//      registers.
// Every operation is a single unfused IEEE op in the reference's order, so a source's contribution
// is bit-identical to the reference's.
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) k_mix_general(const OdbJob* __restrict__ jobs, int n_sources,
                                                            float* __restrict__ partials, int only_flagged,
                                                            const uint32_t* __restrict__ counters) {
    extern __shared__ float smem[];
    pdl_launch_dependents();
    pdl_wait();  // jobs and counters come from the walk kernel (two launches back in the stream)
    // nothing flagged for this kernel: leave at once; k_reduce_tiles reads the same counter and skips our tiles
    if (only_flagged && counters[ODB_CNT_GENERAL] == 0) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* tile = smem + warp * (2 * ODB_TILE_FRAMES);  // cursors [ear][1024] while mixing, the partial tile at the end
    const int tl = blockIdx.y;
    const int gw = blockIdx.x * WARPS + warp, GW = gridDim.x * WARPS;
    float2 acc[ODB_TILE_FRAMES / 32];
#pragma unroll
    for (int j = 0; j < ODB_TILE_FRAMES / 32; j++) acc[j] = make_float2(0.0f, 0.0f);

    for (int sidx = gw; sidx < n_sources; sidx += GW) {
        const OdbJob* job = jobs + (size_t)tl * n_sources + sidx;
        const uint32_t jf = job->flags;
        if (jf & (ODB_JF_SKIP | ODB_JF_RING)) continue;  // flagged buffered-source jobs belong to k_mix_ring
        if (only_flagged && !(jf & ODB_JF_GENERAL)) continue;
        const int nfr = job->n_frames;
        const bool cycle = (jf & ODB_JF_CYCLE) != 0;
        if (cycle && lane < 2 * ODB_TILE_CHUNKS) {  // Cycle::sample, cycle.rs:26-53: the chain lanes also take the (wrapping) taps
            const int e = lane & 1, c = lane >> 1;
            const int m = min(ODB_SPATIAL_CHUNK, nfr - c * ODB_SPATIAL_CHUNK);
            if (m > 0) {
                const float* __restrict__ x0 = job->pcm;
                const unsigned long long ulen = (unsigned long long)job->len;
                const float ds = job->ds[e];
                unsigned long long cbase = (unsigned long long)job->base[e][c];
                float offset = job->off0[e][c];
                float* dst = tile + e * ODB_TILE_FRAMES + c * ODB_SPATIAL_CHUNK;
                for (int k = 0; k < m; k++) {
                    const unsigned long long tr = (unsigned long long)offset;              // :31
                    const float fract = offset - (float)tr;                                // :32 (kept across a wrap)
                    unsigned long long x = cbase + tr;                                     // :33
                    if (x >= ulen) {                                                       // :38-41
                        cbase = 0;
                        offset = (float)(x % ulen) + fract;
                        x = (unsigned long long)offset;
                    }
                    const float a = x0[x];
                    const float b = x < ulen - 1 ? x0[x + 1] : x0[0];                      // :34-37 / :42-46
                    dst[k] = a + fract * (b - a);                                          // frame.rs:39-41
                    offset = offset + ds;                                                  // :50
                }
            }
        } else if (lane < 2 * ODB_TILE_CHUNKS) {
            const int e = lane & 1, c = lane >> 1;
            const bool unit = (jf & (e == 0 ? ODB_JF_FAST_L : ODB_JF_FAST_R)) != 0;
            if (!unit) {
                const float ds = job->ds[e];
                float offset = job->off0[e][c];
                float* dst = tile + e * ODB_TILE_FRAMES + c * ODB_SPATIAL_CHUNK;
                for (int k = 0; k < ODB_SPATIAL_CHUNK; k++) {
                    dst[k] = offset;
                    offset = offset + ds;                                  // frames.rs:195
                }
            }
        }
        __syncwarp();
        const float* __restrict__ pcm = job->pcm;
        const int len = job->len;
        const float fg = job->fixed_gain;
        const bool has_fg = (jf & ODB_JF_FIXED_GAIN) != 0;
#pragma unroll
        for (int e = 0; e < 2; e++) {
            const bool unit = (jf & (e == 0 ? ODB_JF_FAST_L : ODB_JF_FAST_R)) != 0;
            const float pg = job->pg[e], dg = job->dg[e];
#pragma unroll
            for (int j = 0; j < ODB_TILE_FRAMES / 32; j++) {
                const int c = j >> 3;
                const int k = 32 * (j & 7) + lane, i = 32 * j + lane;
                if (i < nfr) {
                    const long long base = job->base[e][c];
                    float smp;
                    if (cycle) {
                        smp = tile[e * ODB_TILE_FRAMES + i];               // lerped by the chain lane above
                    } else {
                        float a, b, fract;
                        if (unit) {                                        // frames.rs:183-187
                            get_pair_mono(pcm, len, base + k, a, b);
                            fract = job->off0[e][c];
                        } else {                                           // frames.rs:191-193
                            const float offset = tile[e * ODB_TILE_FRAMES + i];
                            const long long tr = (long long)offset;
                            get_pair_mono(pcm, len, base + tr, a, b);
                            fract = offset - (float)tr;
                        }
                        smp = a + fract * (b - a);                         // frame.rs:39-41
                    }
                    if (has_fg) smp = smp * fg;                            // gain.rs:35
                    const float gain = pg + (float)(tl * ODB_TILE_FRAMES + i) * dg;  // spatial.rs:459
                    const float contrib = smp * gain;                      // spatial.rs:460
                    if (e == 0) acc[j].x = acc[j].x + contrib;
                    else acc[j].y = acc[j].y + contrib;
                }
            }
        }
        __syncwarp();
    }
    // CTA-level fold in fixed warp order, then one partial tile per CTA.
#pragma unroll
    for (int j = 0; j < ODB_TILE_FRAMES / 32; j++) *reinterpret_cast<float2*>(tile + (32 * j + lane) * 2) = acc[j];
    __syncthreads();
    float* dst = partials + ((size_t)tl * gridDim.x + blockIdx.x) * (2 * ODB_TILE_FRAMES);
    for (int f = threadIdx.x; f < 2 * ODB_TILE_FRAMES; f += WARPS * 32) {
        float sum = 0.0f;
#pragma unroll
        for (int w = 0; w < WARPS; w++) sum = sum + smem[w * (2 * ODB_TILE_FRAMES) + f];
        dst[f] = sum;
    }
}

// ------------------------------------------------------------------------------------------
// Second-stage reduce: sums the per-CTA partial tiles of up to two partial sets (a: staged kernel,
// b: general kernel) in a fixed order (deterministic), applies the optional Tanh / Reinhard wrapper
// (tanh.rs:24-28, reinhard.rs:30-34) and writes the interleaved stereo output. Each block owns 32
// consecutive output floats; its 8 warps take every 8th partial tile, then fold through shared memory.
#define RED_GROUPS 16

__global__ void __launch_bounds__(32 * RED_GROUPS) k_reduce_tiles(const float* __restrict__ pa, int na,
                                                                 const float* __restrict__ pb, int nb,
                                                                 const float* __restrict__ pc, int nc,
                                                                 const uint32_t* __restrict__ counters, int b_is_general,
                                                                 int c_counter, uint32_t* __restrict__ zero_counters,
                                                                 float* __restrict__ out, int n_frames, int channels,
                                                                 int epilogue) {
    __shared__ float fold[RED_GROUPS][32];
    pdl_wait();  // partial tiles come from the mix kernels launched just before
    // reset the other parity's job counters for the next callback (keeps the stream free of memset nodes
    // between the kernels, which programmatic dependent launch needs)
    if (zero_counters && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x < ODB_CNT_WORDS) zero_counters[threadIdx.x] = 0u;
    const int tl = blockIdx.y;
    const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
    const int f = blockIdx.x * 32 + lane;  // float index inside the tile
    if (b_is_general && counters[ODB_CNT_GENERAL] == 0) nb = 0;
    if (c_counter >= 0 && counters[c_counter] == 0) nc = 0;  // that kernel left at once and wrote no tiles
    float sum = 0.0f;
    const size_t tile_floats = (size_t)ODB_TILE_FRAMES * channels;
    const float* p = pa + (size_t)tl * na * tile_floats + f;
    {   // the dominant set: loads issued four at a time, summed in index order
        int i = grp;
        for (; i + 3 * RED_GROUPS < na; i += 4 * RED_GROUPS) {
            const float v0 = p[(size_t)i * tile_floats], v1 = p[(size_t)(i + RED_GROUPS) * tile_floats];
            const float v2 = p[(size_t)(i + 2 * RED_GROUPS) * tile_floats], v3 = p[(size_t)(i + 3 * RED_GROUPS) * tile_floats];
            sum = sum + v0; sum = sum + v1; sum = sum + v2; sum = sum + v3;
        }
        for (; i < na; i += RED_GROUPS) sum = sum + p[(size_t)i * tile_floats];
    }
    const float* q = pb + (size_t)tl * nb * tile_floats + f;
    for (int i = grp; i < nb; i += RED_GROUPS) sum = sum + q[(size_t)i * tile_floats];
    const float* r = pc + (size_t)tl * nc * tile_floats + f;
    for (int i = grp; i < nc; i += RED_GROUPS) sum = sum + r[(size_t)i * tile_floats];
    fold[grp][lane] = sum;
    __syncthreads();
    if (grp != 0) return;
    sum = 0.0f;
#pragma unroll
    for (int g = 0; g < RED_GROUPS; g++) sum = sum + fold[g][lane];
    const int frame = tl * ODB_TILE_FRAMES + f / channels;
    if (frame >= n_frames) return;
    if ((epilogue & 0xFF) == 1) sum = tanhf(sum);
    else if ((epilogue & 0xFF) == 2) sum = sum / (1.0f + fabsf(sum));
    if (epilogue & ODB_EPILOGUE_I16_BIT) {
        // offline render, examples/offline.rs:39 `(sample * i16::MAX as f32) as i16`: one f32 multiply, then Rust's
        // float -> int `as`: toward zero, saturating, NaN -> 0 (cvt.rzi.s32.f32 does the same for the i32 range)
        int v = __float2int_rz(sum * 32767.0f);
        v = max(-32768, min(32767, v));
        reinterpret_cast<short*>(out)[(size_t)tl * tile_floats + f] = (short)v;
    } else {
        out[(size_t)tl * tile_floats + f] = sum;
    }
}

// PCM ingest, examples/wav.rs:30-37: integer samples -> `sample as f32 / max_value as f32`, max_value = 2^(bits-1) - 1
__global__ void k_convert_i16(const short* __restrict__ in, float* __restrict__ out, size_t n, float max_value) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (float)in[i] / max_value;
}

}  // namespace odbk

// ------------------------------------------------------------------------------------------
// Launchers (host side, called from odb_api.cu)
using namespace odbk;

void odb_launch_convert_i16(const short* in, float* out, size_t n, float max_value, cudaStream_t st) {
    if (n == 0) return;
    k_convert_i16<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(in, out, n, max_value);
}
void odb_launch_scatter_sources(OdbSource* src, const OdbSource* staged, const uint32_t* slots, int n, cudaStream_t st) {
    if (n <= 0) return;
    long long total = (long long)n * (long long)(sizeof(OdbSource) / 16);
    k_scatter_sources<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(src, staged, slots, n);
}
void odb_launch_scatter_motion(OdbSource* src, const OdbMotionMsg* msgs, int n, cudaStream_t st) {
    if (n <= 0) return;
    k_scatter_motion<<<(n + 127) / 128, 128, 0, st>>>(src, msgs, n);
}
void odb_launch_scatter_params(OdbSource* src, const OdbParamMsg* msgs, int n, cudaStream_t st) {
    if (n <= 0) return;
    k_scatter_params<<<(n + 127) / 128, 128, 0, st>>>(src, msgs, n);
}
int odb_launch_walk_seek(OdbSource* src, const uint32_t* order, OdbJob* jobs, uint32_t* removed, int removed_cap,
                         uint32_t* counters, uint32_t* zero_counters, unsigned long long* walked, const OdbCallback& cb,
                         cudaStream_t st) {
    if (cb.n_sources <= 0) return 0;
    const int per_block = WALK_THREADS / 2;
    const dim3 grid((cb.n_sources + per_block - 1) / per_block);
    if (!walked)
        odb_launch_pdl(k_walk_seek<true>, grid, dim3(WALK_THREADS), 0, st, src, order, jobs, removed, removed_cap, counters, zero_counters,
                       walked, cb);
    else
        odb_launch_pdl(k_walk_seek<false>, grid, dim3(WALK_THREADS), 0, st, src, order, jobs, removed, removed_cap, counters, zero_counters,
                       walked, cb);
    return (int)grid.x;
}

static const int GEN_WARPS = 8;
int odb_mix_general_ctas(int n_sources, int sm_count) {
    int want = (n_sources + GEN_WARPS - 1) / GEN_WARPS;
    int cap = sm_count * 3;
    return want < 1 ? 1 : (want > cap ? cap : want);
}
cudaError_t odb_launch_mix_general(const OdbJob* jobs, int n_sources, int n_tiles, float* partials, int n_ctas,
                                   int only_flagged, const uint32_t* counters, cudaStream_t st) {
    static bool attr_set = false;
    const int smem = GEN_WARPS * 2 * ODB_TILE_FRAMES * (int)sizeof(float);
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k_mix_general<GEN_WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    dim3 grid(n_ctas, n_tiles);
    return odb_launch_pdl(k_mix_general<GEN_WARPS>, grid, dim3(GEN_WARPS * 32), (size_t)smem, st, jobs, n_sources, partials, only_flagged,
                          counters);
}
void odb_launch_reduce(const float* pa, int na, const float* pb, int nb, const float* pc, int nc, const uint32_t* counters,
                       int b_is_general, int c_counter, uint32_t* zero_counters, float* out, int n_frames, int n_tiles,
                       int channels, int epilogue, cudaStream_t st) {
    if (n_frames <= 0) return;
    dim3 grid(channels * ODB_TILE_FRAMES / 32, n_tiles);
    odb_launch_pdl(k_reduce_tiles, grid, dim3(32 * RED_GROUPS), 0, st, pa, na, pb, nb, pc, nc, counters, b_is_general, c_counter, zero_counters, out, n_frames,
                   channels, epilogue);
}
