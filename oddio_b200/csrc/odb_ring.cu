// Buffered spatial sources (`SpatialSceneControl::play_buffered`, spatial.rs:314-340): a per-source delay
// ring in HBM (ring.rs) written from the inner chain Gain(FixedGain(Speed(FramesSignal))) each callback
// and read back with a fractional, wrapping cursor per ear.
//   k_walk_buffered : walk_set for the buffered set (spatial.rs:191-265) + the O(1) part of the mix closure
//                     (spatial.rs:402-431) and of Ring::write (ring.rs:18-41)
//   k_ring_write    : Ring::write's inner.sample() spans (frames.rs:176-201, gain.rs:32-37, :103-122)
//   k_mix_ring      : Ring::sample (ring.rs:51-79) + gain ramp + accumulate (spatial.rs:422-429)
#include <cuda_runtime.h>

#include "odb_kernels.h"
#include "odb_math.cuh"

namespace odbk {

// f32::rem_euclid (core): r = a % b; if r < 0 { r + |b| }
__device__ __forceinline__ float rem_euclidf(float a, float b) {
    float r = fmodf(a, b);
    return r < 0.0f ? r + fabsf(b) : r;
}

// One thread per buffered source.
__global__ void __launch_bounds__(128) k_walk_buffered(OdbSource* __restrict__ src, const uint32_t* __restrict__ order,
                                                       OdbRingJob* __restrict__ jobs, OdbJob* __restrict__ fjobs,
                                                       OdbRingWrite* __restrict__ writes, uint32_t* __restrict__ removed,
                                                       int removed_cap, uint32_t* __restrict__ counters,
                                                       uint32_t* __restrict__ literal_list, OdbCallback cb) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= cb.n_sources) return;
    const uint32_t slot = order[idx];
    OdbSource* sp = src + slot;
    OdbSource s;
    load_source(s, sp);
    const int n = cb.n_frames, nt = cb.n_tiles, ns = cb.n_sources;
    const float elapsed = cb.elapsed;
    V3 prev_position, next_position;
    uint32_t flags;
    if (!walk_common(sp, s, cb, slot, removed, removed_cap, prev_position, next_position, flags)) {
        for (int tl = 0; tl < nt; tl++) {
            jobs[(size_t)tl * ns + idx].flags = ODB_JF_SKIP;
            fjobs[(size_t)tl * cb.job_stride + cb.job_offset + idx].flags = ODB_JF_SKIP | ODB_JF_RING;
        }
        writes[idx].flags = ODB_JF_SKIP;
        return;
    }

    // --- queue.write(&mut inner, rate, elapsed), ring.rs:18-41
    const float ratef = (float)s.ring_rate;
    const float capf = (float)s.ring_cap;
    const float end = fmodf(s.ring_write + elapsed * ratef, capf);               // ring.rs:28
    const int start_idx = (int)ceilf(s.ring_write), end_idx = (int)ceilf(end);   // ring.rs:30-31
    const float w_interval = 1.0f / ratef;                                        // ring.rs:32
    OdbRingWrite w;
    w.ring = s.ring; w.pcm = s.pcm; w.cap = s.ring_cap; w.len = s.len;
    w.fixed_gain = s.fixed_gain;
    w.gstep = w_interval / ODB_GAIN_SMOOTHING;                                    // gain.rs:120
    int n_spans;
    if (end_idx > start_idx) {                                                    // ring.rs:33-34
        n_spans = 1;
        w.start[0] = start_idx; w.n[0] = end_idx - start_idx;
        w.start[1] = 0; w.n[1] = 0;
    } else {                                                                      // ring.rs:36-37 (end_idx == start_idx: the whole ring)
        n_spans = 2;
        w.start[0] = start_idx; w.n[0] = s.ring_cap - start_idx;
        w.start[1] = 0; w.n[1] = end_idx;
    }
    const double rate = s.rate;
    const float iv = (flags & ODB_SF_SPEED) ? w_interval * s.speed : w_interval;  // speed.rs:34
    const float ds = iv * (float)rate;                                            // frames.rs:178
    const bool unit = fabsf(ds - 1.0f) <= ODB_F32_EPSILON;                        // frames.rs:180
    w.ds = ds;
    uint32_t wf = unit ? ODB_JF_FAST_L : 0u;
    double t = s.t;
    long long sample_t = s.sample_t;
    float gprev = s.gain_prev, gnext = s.gain_next, gprog = s.gain_progress;
    for (int k = 0; k < 2; k++) {
        w.base[k] = 0; w.off0[k] = 0.0f; w.g[k] = 1.0f; w.gprev[k] = 0.0f; w.gnext[k] = 0.0f; w.gprog[k] = 0.0f;
        if (k >= n_spans) continue;
        const int m = w.n[k];
        const double s0 = t * rate;                                               // frames.rs:177
        const long long base = (long long)s0;                                     // frames.rs:179
        w.base[k] = sat_i32(base);
        w.off0[k] = (float)(s0 - (double)base);                                   // frames.rs:183 / :189
        if (base > (1ll << 29) || base < -(1ll << 29)) w.n[k] = 0;               // beyond any PCM: zeros, ring is already zero there
        t = t + (double)iv * (double)m;                                           // frames.rs:198
        sample_t = (long long)(t * rate);
        if (flags & ODB_SF_GAIN) {                                                // gain.rs:104-121, once per inner.sample call
            if (gnext != s.gain_shared) {
                gprev = gprev + gprog * (gnext - gprev);
                gnext = s.gain_shared;
                gprog = 0.0f;
            }
            if (gprog == 1.0f) {
                w.g[k] = gprev + gprog * (gnext - gprev);
            } else {
                wf |= (k == 0 ? ODB_JF_RAMP : ODB_JF_RAMP1);
                w.gprev[k] = gprev; w.gnext[k] = gnext; w.gprog[k] = gprog;
                for (int i = 0; i < m; i++) gprog = fminf(gprog + w.gstep, 1.0f);
            }
        }
    }
    w.flags = wf;
    writes[idx] = w;
    sp->t = t;
    sp->sample_t = sample_t;
    sp->gain_prev = gprev; sp->gain_next = gnext; sp->gain_progress = gprog;
    sp->ring_write = end;                                                         // ring.rs:40

    // --- per ear: clamp into the ring, cursor and gain set-up, spatial.rs:409-431 + ring.rs:57-58
    // Every (tile, source) gets two records: an OdbRingJob for the literal ring kernel and an OdbJob for the staged
    // mix kernel, which reads the delay ring exactly like a Frames block whenever the tile's reads do not wrap
    // around the ring (then Ring::sample's cursor is a plain `offset += ds` chain, ring.rs:59-77 without :67-75).
    const float nf = (float)n;
    const int n_chunks = (n + ODB_SPATIAL_CHUNK - 1) / ODB_SPATIAL_CHUNK;
    uint32_t wrap_tiles = 0;      // bit tl: some read of tile tl may wrap, or leaves the staged kernel's envelope
    int wlo[4][2], whi[4][2];     // per tile and 512-frame half: ring index range both ears can touch
    for (int tl = 0; tl < 4; tl++) { wlo[tl][0] = wlo[tl][1] = 0x7fffffff; whi[tl][0] = whi[tl][1] = -0x7fffffff; }
    for (int e = 0; e < 2; e++) {
        EarSt ps = ear_state(prev_position, e, s.radius);
        EarSt nx = ear_state(next_position, e, s.radius);
        const float prev_offset = fmaxf(ps.offset - elapsed, -s.max_delay);      // :414
        const float next_offset = fmaxf(nx.offset, -s.max_delay);                // :415
        const float dt = (next_offset - prev_offset) / nf;                        // :417
        const float d_gain = (nx.gain - ps.gain) / nf;                            // :418
        const float rds = dt * ratef;                                             // ring.rs:58
        const bool ds_ok = rds > 0.0f && rds <= ODB_FAST_DS_MAX;
        for (int cg = 0; cg < n_chunks; cg++) {
            const int tl = cg / ODB_TILE_CHUNKS, c = cg % ODB_TILE_CHUNKS;
            const int m = min(ODB_SPATIAL_CHUNK, n - cg * ODB_SPATIAL_CHUNK);
            const float tt = prev_offset + (float)(cg * ODB_SPATIAL_CHUNK) * dt;  // :423
            const float off0 = rem_euclidf(end + tt * ratef, capf);               // ring.rs:57 (write is already `end`)
            jobs[(size_t)tl * ns + idx].off0[e][c] = off0;
            OdbJob* fj = fjobs + (size_t)tl * cb.job_stride + cb.job_offset + idx;
            fj->base[e][c] = 0;
            fj->off0[e][c] = off0;
            const int lo = (int)off0;
            // The staged kernel's window and its index split assume the f32 chain `offset += ds` stays within a few
            // samples of off0 + k * rds. off0 is an absolute ring offset here: from 2^17 samples on, ulp(off0) >= 1/64
            // and 255 roundings may drift by more than the margin - those reads take the literal ring kernel, which
            // walks the chain itself (a ring that long is max_distance > 900 m at 48 kHz).
            const float last_est = __fmaf_rn((float)(m - 1), rds, off0);
            const bool far = !(last_est < 131072.0f);
            const int hi = (ds_ok && !far) ? (int)last_est + 4 : 0x7ffffff0;
            if (!ds_ok || far || lo < 0 || hi > s.ring_cap - 2) wrap_tiles |= 1u << tl;  // x + 1 must stay below buffer.len()
            const int h = c / ODB_FAST_HALF_CHUNKS;
            wlo[tl][h] = min(wlo[tl][h], lo);
            whi[tl][h] = max(whi[tl][h], hi);
        }
        for (int tl = 0; tl < nt; tl++) {
            OdbRingJob* j = jobs + (size_t)tl * ns + idx;
            j->ds[e] = rds; j->pg[e] = ps.gain; j->dg[e] = d_gain;
            OdbJob* fj = fjobs + (size_t)tl * cb.job_stride + cb.job_offset + idx;
            fj->ds[e] = rds; fj->pg[e] = ps.gain; fj->dg[e] = d_gain;
        }
    }
    uint32_t n_staged = 0;
    for (int tl = 0; tl < nt; tl++) {
        uint32_t f = ODB_JF_RING;
        if (((wrap_tiles >> tl) & 1u) || cb.force_general) f |= ODB_JF_GENERAL;
        int ws[2] = {0, 0}, wl[2] = {0, 0};
        for (int h = 0; h < 2; h++) {
            if (whi[tl][h] >= wlo[tl][h] && !(f & ODB_JF_GENERAL)) {
                ws[h] = wlo[tl][h] & ~3;
                wl[h] = ((whi[tl][h] - ws[h] + 1) + 3) & ~3;
                if (wl[h] > ODB_FAST_PCM_CAP || ws[h] + wl[h] > s.ring_cap) f |= ODB_JF_GENERAL;
            }
        }
        OdbJob* fj = fjobs + (size_t)tl * cb.job_stride + cb.job_offset + idx;
        fj->pcm = s.ring; fj->len = s.ring_cap; fj->fixed_gain = 1.0f;
        fj->n_frames = min(ODB_TILE_FRAMES, n - tl * ODB_TILE_FRAMES);
        for (int h = 0; h < 2; h++) {
            fj->window[h][0] = ws[h];
            fj->window[h][1] = (f & ODB_JF_GENERAL) ? 0 : wl[h];
        }
        fj->flags = f;
        OdbRingJob* j = jobs + (size_t)tl * ns + idx;
        j->ring = s.ring; j->cap = s.ring_cap; j->flags = (f & ODB_JF_GENERAL) ? ODB_JF_GENERAL : 0u;
        j->n_frames = min(ODB_TILE_FRAMES, n - tl * ODB_TILE_FRAMES);
        if (f & ODB_JF_GENERAL) {  // compact list of the jobs k_mix_ring has to take: one warp each, no scanning
            const uint32_t k = atomicAdd(counters + ODB_CNT_RING_GENERAL, 1u);
            literal_list[k] = (uint32_t)(tl * ns + idx);
        } else {
            n_staged++;
        }
    }
    if (n_staged) atomicAdd(counters + ODB_CNT_FAST, n_staged);
}

// One warp per buffered source: fills the span(s) Ring::write hands to inner.sample().
__global__ void __launch_bounds__(256) k_ring_write(const OdbRingWrite* __restrict__ writes, int n_sources) {
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (gw >= n_sources) return;
    const OdbRingWrite w = writes[gw];
    if (w.flags & ODB_JF_SKIP) return;
    const bool unit = (w.flags & ODB_JF_FAST_L) != 0;
    for (int k = 0; k < 2; k++) {
        const int m = w.n[k];
        if (m <= 0) continue;
        const bool ramp = (w.flags & (k == 0 ? ODB_JF_RAMP : ODB_JF_RAMP1)) != 0;
        float* __restrict__ dst = w.ring + w.start[k];
        const long long base = w.base[k];
        if (unit && !ramp) {  // frames.rs:183-187: every frame independent
            const float fract = w.off0[k], fg = w.fixed_gain, g = w.g[k];
            constexpr int U = 8;  // frames per lane whose loads are issued together
            for (int i0 = lane; i0 < m; i0 += 32 * U) {
                float a[U], b[U];
#pragma unroll
                for (int u = 0; u < U; u++) {
                    a[u] = 0.0f; b[u] = 0.0f;
                    if (i0 + 32 * u < m) get_pair_mono(w.pcm, w.len, base + i0 + 32 * u, a[u], b[u]);
                }
#pragma unroll
                for (int u = 0; u < U; u++) {
                    float v = a[u] + fract * (b[u] - a[u]);   // frame.rs:39-41
                    v = v * fg;                               // gain.rs:35 (x * 1.0 == x)
                    v = v * g;                                // gain.rs:112-114
                    if (i0 + 32 * u < m) dst[i0 + 32 * u] = v;
                }
            }
        } else if (lane == 0) {  // serial cursor and/or serial gain ramp: literal
            float offset = w.off0[k], gprog = w.gprog[k];
            const float ds = w.ds, fg = w.fixed_gain, g = w.g[k], gprev = w.gprev[k], gnext = w.gnext[k], gstep = w.gstep;
            for (int i = 0; i < m; i++) {
                long long idx;
                float fract;
                if (unit) { idx = base + i; fract = offset; }
                else {                                                  // frames.rs:191-195
                    const long long tr = (long long)offset;
                    idx = base + tr;
                    fract = offset - (float)tr;
                    offset = offset + ds;
                }
                float a, b;
                get_pair_mono(w.pcm, w.len, idx, a, b);
                float v = a + fract * (b - a);
                v = v * fg;
                if (ramp) {                                             // gain.rs:118-121
                    v = v * (gprev + gprog * (gnext - gprev));
                    gprog = fminf(gprog + gstep, 1.0f);
                } else {
                    v = v * g;
                }
                dst[i] = v;
            }
        }
    }
}

// One warp per (tile, buffered source). Lanes 0..7 walk one (ear, chunk) cursor chain of Ring::sample
// literally (ring.rs:59-77, including the wrap re-basing `offset = x as f32 + fract`) and park index and
// fraction of every frame in warp-private shared memory; then all lanes gather from the ring in HBM, lerp,
// apply the gain ramp and accumulate in registers (lane l owns frames l, l+32, ...).
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) k_mix_ring(const OdbRingJob* __restrict__ jobs, int n_sources,
                                                         float* __restrict__ partials, int only_flagged,
                                                         const uint32_t* __restrict__ counters,
                                                         const uint32_t* __restrict__ literal_list) {
    extern __shared__ float smem[];
    pdl_launch_dependents();
    pdl_wait();
    // nothing wraps this callback: leave at once; k_reduce_tiles reads the same counter and skips our tiles
    if (only_flagged && counters[ODB_CNT_RING_GENERAL] == 0) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* fr_s = smem + warp * (4 * ODB_TILE_FRAMES);                             // [ear][1024] fractions
    int* ix_s = reinterpret_cast<int*>(fr_s + 2 * ODB_TILE_FRAMES);                // [ear][1024] ring indices
    const int tl = blockIdx.y;
    const int gw = blockIdx.x * WARPS + warp, GW = gridDim.x * WARPS;
    float2 acc[ODB_TILE_FRAMES / 32];
#pragma unroll
    for (int j = 0; j < ODB_TILE_FRAMES / 32; j++) acc[j] = make_float2(0.0f, 0.0f);

    // only_flagged: the walk kernel's compact list of (tile, source) jobs whose reads wrap; else every job of the tile
    const int n_items = only_flagged ? (int)counters[ODB_CNT_RING_GENERAL] : n_sources;
    for (int it = gw; it < n_items; it += GW) {
        int sidx = it;
        if (only_flagged) {
            const uint32_t lin = literal_list[it];
            if ((int)(lin / (uint32_t)n_sources) != tl) continue;
            sidx = (int)(lin % (uint32_t)n_sources);
        }
        const OdbRingJob* job = jobs + (size_t)tl * n_sources + sidx;
        if (job->flags & ODB_JF_SKIP) continue;
        const int nfr = job->n_frames;
        const unsigned len = (unsigned)job->cap;
        if (lane < 2 * ODB_TILE_CHUNKS) {
            const int e = lane & 1, c = lane >> 1;
            const float ds = job->ds[e];
            float offset = job->off0[e][c];
            float* fdst = fr_s + e * ODB_TILE_FRAMES + c * ODB_SPATIAL_CHUNK;
            int* xdst = ix_s + e * ODB_TILE_FRAMES + c * ODB_SPATIAL_CHUNK;
            for (int k = 0; k < ODB_SPATIAL_CHUNK; k++) {
                unsigned x = (unsigned)offset;                           // ring.rs:60 to_int_unchecked::<usize>
                const float fract = offset - (float)x;                   // :61
                if (x >= len) {                                          // :67-69
                    x = x % len;
                    offset = (float)x + fract;
                }
                fdst[k] = fract;
                xdst[k] = (int)x;
                offset = offset + ds;                                    // :77
            }
        }
        __syncwarp();
        const float* __restrict__ ring = job->ring;
#pragma unroll
        for (int e = 0; e < 2; e++) {
            const float pg = job->pg[e], dg = job->dg[e];
#pragma unroll
            for (int j = 0; j < ODB_TILE_FRAMES / 32; j++) {
                const int i = 32 * j + lane;
                if (i < nfr) {
                    const unsigned x = (unsigned)ix_s[e * ODB_TILE_FRAMES + i];
                    const float fract = fr_s[e * ODB_TILE_FRAMES + i];
                    const float a = ring[x];
                    const float b = ring[x + 1 < len ? x + 1 : 0];       // :63-66, :70-74
                    const float smp = a + fract * (b - a);               // :76
                    const float gain = pg + (float)(tl * ODB_TILE_FRAMES + i) * dg;  // spatial.rs:426
                    const float contrib = smp * gain;                    // spatial.rs:427
                    if (e == 0) acc[j].x = acc[j].x + contrib;
                    else acc[j].y = acc[j].y + contrib;
                }
            }
        }
        __syncwarp();
    }
    float* tile = smem + warp * (4 * ODB_TILE_FRAMES);
#pragma unroll
    for (int j = 0; j < ODB_TILE_FRAMES / 32; j++) *reinterpret_cast<float2*>(tile + (32 * j + lane) * 2) = acc[j];
    __syncthreads();
    float* dst = partials + ((size_t)tl * gridDim.x + blockIdx.x) * (2 * ODB_TILE_FRAMES);
    for (int f = threadIdx.x; f < 2 * ODB_TILE_FRAMES; f += WARPS * 32) {
        float sum = 0.0f;
#pragma unroll
        for (int w = 0; w < WARPS; w++) sum = sum + smem[w * (4 * ODB_TILE_FRAMES) + f];
        dst[f] = sum;
    }
}

}  // namespace odbk

using namespace odbk;

static const int RING_WARPS = 8;

void odb_launch_walk_buffered(OdbSource* src, const uint32_t* order, OdbRingJob* jobs, OdbJob* fjobs, OdbRingWrite* writes,
                              uint32_t* removed, int removed_cap, uint32_t* counters, uint32_t* literal_list,
                              const OdbCallback& cb, cudaStream_t st) {
    if (cb.n_sources <= 0) return;
    k_walk_buffered<<<(cb.n_sources + 127) / 128, 128, 0, st>>>(src, order, jobs, fjobs, writes, removed, removed_cap, counters,
                                                               literal_list, cb);
}
void odb_launch_ring_write(const OdbRingWrite* writes, int n_sources, cudaStream_t st) {
    if (n_sources <= 0) return;
    k_ring_write<<<(n_sources + 7) / 8, 256, 0, st>>>(writes, n_sources);
}
int odb_mix_ring_ctas(int n_sources, int sm_count) {
    int want = (n_sources + RING_WARPS - 1) / RING_WARPS;
    return want < 1 ? 1 : (want > sm_count ? sm_count : want);
}
cudaError_t odb_launch_mix_ring(const OdbRingJob* jobs, int n_sources, int n_tiles, float* partials, int n_ctas,
                                int only_flagged, const uint32_t* counters, const uint32_t* literal_list, cudaStream_t st) {
    const int smem = RING_WARPS * 4 * ODB_TILE_FRAMES * (int)sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(k_mix_ring<RING_WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    dim3 grid(n_ctas, n_tiles);
    return odb_launch_pdl(k_mix_ring<RING_WARPS>, grid, dim3(RING_WARPS * 32), (size_t)smem, st, jobs, n_sources, partials, only_flagged,
                          counters, literal_list);
}
