// Mixer kernels: walk (per-source set-up of Mixer::sample, mixer.rs:92-119, over the closed chain
// Gain(FixedGain(Speed(FramesSignal)))), the streaming kernel for FramesSignal's ds ~= 1 path
// (frames.rs:180-187) and the literal kernel for everything else (resampling chains, gain ramps).
// A mixer tile is one staging chunk of the reference: 1024 frames (mixer.rs:77, :110-111).
#include <cuda_runtime.h>

#include "odb_kernels.h"
#include "odb_math.cuh"
#include "odb_async.cuh"

namespace odbk {

// ------------------------------------------------------------------------------------------
// One thread per source. mixer.rs:100-107 (stop / finished -> remove), then per <=1024-frame chunk
// the O(1) part of the chain: Speed (speed.rs:32-35), FramesSignal's cursor set-up (frames.rs:177-183)
// and f64 advance (:198-200), Gain's Smoothed state machine (gain.rs:104-121, smooth.rs:47-72).
__global__ void __launch_bounds__(128) k_walk_mixer(OdbSource* __restrict__ src, const uint32_t* __restrict__ order,
                                                    OdbMixJob* __restrict__ jobs, uint32_t* __restrict__ removed,
                                                    int removed_cap, uint32_t* __restrict__ counters, OdbCallback cb) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= cb.n_sources) return;
    OdbSource* sp = src + order[idx];
    OdbSource s;
    load_source(s, sp);
    const int ns = cb.n_sources, nt = cb.n_tiles;
    uint32_t flags = s.flags;
    const bool was_stopped = (flags & ODB_SF_STOPPED) != 0;
    double t = s.t;
    const double rate = s.rate;
    // mixer.rs:102: `signal.stop.load() || signal.inner.is_finished()`; is_finished forwards through
    // Gain / FixedGain / Speed (gain.rs:39-41,:124-126, speed.rs:37-39) to frames.rs:204-206
    if (was_stopped || (flags & ODB_SF_STOP_REQ) || t >= s.t_end) {
        if (!was_stopped) {
            uint32_t k = atomicAdd(removed, 1u);
            removed[1 + (k & (uint32_t)(removed_cap - 1))] = order[idx];
            sp->flags = flags | ODB_SF_STOPPED;
        }
        for (int tl = 0; tl < nt; tl++) jobs[(size_t)tl * ns + idx].flags = ODB_JF_SKIP;
        return;
    }
    const float ratef = (float)rate;
    const float iv = (flags & ODB_SF_SPEED) ? cb.interval * s.speed : cb.interval;  // speed.rs:34
    const float ds = iv * ratef;                                                    // frames.rs:178
    const bool unit = fabsf(ds - 1.0f) <= ODB_F32_EPSILON;                          // frames.rs:180
    const int ch = s.channels;
    float gprev = s.gain_prev, gnext = s.gain_next, gprog = s.gain_progress;
    const float gstep = cb.interval / ODB_GAIN_SMOOTHING;                           // gain.rs:120
    long long sample_t = s.sample_t;
    uint32_t n_general = 0, n_fast = 0, n_resample = 0;
    const bool cycle = (flags & ODB_SF_CYCLE) != 0;
    for (int tl = 0; tl < nt; tl++) {
        const int n = min(ODB_MIXER_CHUNK, cb.n_frames - tl * ODB_MIXER_CHUNK);
        OdbMixJob j;
        j.pcm = s.pcm; j.len = s.len; j.n_frames = n; j.ds = ds;
        uint32_t jf = unit ? ODB_JF_FAST_L : 0u;
        long long base;
        float off0;
        if (cycle) {
            // Cycle::sample (cycle.rs:26-53): `t` is the cursor in samples. Where the next chunk starts depends on the f32
            // chain of this one (and on where it wrapped), so the chain is walked here once without the taps; the literal
            // kernel walks it again from (base, off0) with them.
            jf = ODB_JF_CYCLE;
            const unsigned long long len = (unsigned long long)s.len;
            unsigned long long cbase = (unsigned long long)t;                       // :28 `self.cursor as usize`
            float offset = (float)(t - (double)cbase);                              // :29
            base = (long long)cbase; off0 = offset;
            for (int i = 0; i < n; i++) {
                const unsigned long long tr = (unsigned long long)offset;           // :31
                const float fract = offset - (float)tr;                             // :32
                const unsigned long long x = cbase + tr;                            // :33
                if (x >= len) { cbase = 0; offset = (float)(x % len) + fract; }     // :38-40
                offset = offset + ds;                                               // :50
            }
            t = (double)cbase + (double)offset;                                     // :52
        } else {
            const double s0 = t * rate;                                             // frames.rs:177
            base = (long long)s0;                                                   // frames.rs:179
            off0 = (float)(s0 - (double)base);                                      // frames.rs:183 / :189
            t = t + (double)iv * (double)n;                                         // frames.rs:198
            sample_t = (long long)(t * rate);                                       // frames.rs:199-200
        }
        j.base = sat_i32(base); j.off0 = off0;
        j.fixed_gain = s.fixed_gain;                                                // 1.0 when the chain has no FixedGain
        j.g = 1.0f; j.gprev = 0.0f; j.gnext = 0.0f; j.gprog = 0.0f; j.gstep = gstep;
        if (flags & ODB_SF_GAIN) {                                                  // gain.rs:104-121
            if (gnext != s.gain_shared) {                                           // Smoothed::set, smooth.rs:57-64
                gprev = gprev + gprog * (gnext - gprev);
                gnext = s.gain_shared;
                gprog = 0.0f;
            }
            if (gprog == 1.0f) {
                j.g = gprev + gprog * (gnext - gprev);                              // Smoothed::get at progress 1 (smooth.rs:86-91)
            } else {
                jf |= ODB_JF_RAMP | ODB_JF_GENERAL;
                j.gprev = gprev; j.gnext = gnext; j.gprog = gprog;
                for (int i = 0; i < n; i++) gprog = fminf(gprog + gstep, 1.0f);     // Smoothed::advance per sample (gain.rs:120)
            }
        }
        // what the streaming kernel may touch: frames [base, base + n] of the zero-padded block
        const long long pad_frames = ODB_PCM_PAD / ch;
        // (4 frames of slack on both sides: the bulk copy rounds its window to 16-byte boundaries)
        // frames the kernels may touch: unit path [base, base + n], resampling path [base, base + trunc(offset_{n-1}) + 1]
        // (the f32 cursor stays within 0.1 of off0 + (n-1)*ds over 1024 steps; +4 is a safe bound)
        const long long reach = unit ? n : (long long)((double)off0 + (double)(n - 1) * (double)ds) + 4;
        const bool in_block = base >= -(pad_frames - 4) && base + reach + 1 <= (long long)s.len + pad_frames - 4 &&
                              base < (1ll << 29) && base > -(1ll << 29);
        if (cb.force_general || off0 < 0.0f || !in_block || (jf & (ODB_JF_RAMP | ODB_JF_CYCLE))) jf |= ODB_JF_GENERAL;
        else if (!unit) {
            if (ds > 0.0f && (reach + 2) * ch + 4 <= ODB_MIXER_RESAMPLE_CAP) jf |= ODB_JF_RESAMPLE;
            else jf |= ODB_JF_GENERAL;
        }
        j.flags = jf;
        jobs[(size_t)tl * ns + idx] = j;
        if (jf & ODB_JF_GENERAL) n_general++; else if (jf & ODB_JF_RESAMPLE) n_resample++; else n_fast++;
    }
    sp->t = t;
    sp->sample_t = sample_t;
    sp->gain_prev = gprev; sp->gain_next = gnext; sp->gain_progress = gprog;
    if (n_general) atomicAdd(counters + ODB_CNT_GENERAL, n_general);
    if (n_fast) atomicAdd(counters + ODB_CNT_FAST, n_fast);
    if (n_resample) atomicAdd(counters + ODB_CNT_RESAMPLE, n_resample);
}

// ------------------------------------------------------------------------------------------
// Streaming kernel, ds ~= 1 path: out[i] += ((a + fract * (b - a)) * fixed_gain) * g with a = x[base+i],
// b = x[base+i+1]. One warp per (source, chunk); the chunk's PCM - frames [base, base+n], one contiguous
// block of (n+1)*CH floats - is brought into the warp's shared-memory buffer by one bulk async copy (TMA)
// while the warp is still consuming the previous source from its other buffer, so HBM sees long contiguous
// requests with two of them in flight per warp. Lane l owns frames l, l+32, ... (conflict-free LDS) and keeps
// 1024*CH/32 register accumulators. Multiplying by a gain of exactly 1.0 is the identity, so the reference's
// `if g != 1.0` (gain.rs:111) and the absence of a FixedGain need no branches.
template <int CH>
struct MixerStream {
    static constexpr int WARPS = CH == 2 ? 12 : 8;                             // stereo: one 12-warp CTA per SM; mono: three 8-warp CTAs
    static constexpr int BUF_FLOATS = (ODB_MIXER_CHUNK + 1) * CH + 8;          // window + alignment slack, multiple of 4
    static constexpr int BUF_BYTES = ((BUF_FLOATS * 4 + 127) / 128) * 128;
    static constexpr int WARP_BYTES = 2 * BUF_BYTES;                            // >= one partial tile (1024*CH*4)
    static constexpr int SMEM_BYTES = WARPS * WARP_BYTES + WARPS * 16;
};

template <int CH>
__global__ void __launch_bounds__(MixerStream<CH>::WARPS * 32) k_mixer_unit(const OdbMixJob* __restrict__ jobs, int n_sources,
                                                                            float* __restrict__ partials) {
    typedef MixerStream<CH> C;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tl = blockIdx.y;
    const int gw = blockIdx.x * C::WARPS + warp, GW = gridDim.x * C::WARPS;
    const uint32_t buf_sa = smem_u32(smem_raw) + (uint32_t)(warp * C::WARP_BYTES);
    const uint32_t bar_sa = smem_u32(smem_raw) + (uint32_t)(C::WARPS * C::WARP_BYTES + warp * 16);
    if (lane == 0) {
        mbar_init(bar_sa, 1);
        mbar_init(bar_sa + 8, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    constexpr int NACC = ODB_MIXER_CHUNK / 32;
    float acc[NACC][CH];
#pragma unroll
    for (int j = 0; j < NACC; j++)
#pragma unroll
        for (int c = 0; c < CH; c++) acc[j][c] = 0.0f;
    const OdbMixJob* tile_jobs = jobs + (size_t)tl * n_sources;

    // next job at or after `from` (stride GW) that this kernel mixes; n_sources if none. Warp-uniform.
    auto next_job = [&](int from) {
        for (; from < n_sources; from += GW)
            if (!(tile_jobs[from].flags & (ODB_JF_SKIP | ODB_JF_GENERAL | ODB_JF_RESAMPLE))) break;
        return from < n_sources ? from : n_sources;
    };
    // lane 0: start the copy of job `j`'s window into buffer `b`. The source address is rounded down to 16 bytes.
    auto start_copy = [&](int j, uint32_t b) {
        if (lane == 0) {
            const OdbMixJob* job = tile_jobs + j;
            const long long first = (long long)job->base * CH;                 // float index of frame `base`
            const long long start = first & ~3ll;
            const uint32_t floats = (uint32_t)((first - start) + (job->n_frames + 1) * CH + 3) & ~3u;
            mbar_expect_tx(bar_sa + b * 8, floats * 4u);
            bulk_g2s(buf_sa + b * C::BUF_BYTES, job->pcm + start, floats * 4u, bar_sa + b * 8);
        }
    };

    uint32_t parity = 0, buf = 0;
    int cur = next_job(gw);
    if (cur < n_sources) start_copy(cur, buf);
    while (cur < n_sources) {
        const int nxt = next_job(cur + GW);
        if (nxt < n_sources) start_copy(nxt, buf ^ 1u);
        const OdbMixJob* job = tile_jobs + cur;
        const float fract = job->off0, fg = job->fixed_gain, g = job->g;
        const int n = job->n_frames;
        const uint32_t skew = (uint32_t)(((long long)job->base * CH) & 3ll);   // floats the copy started early
        const uint32_t x_sa = buf_sa + buf * C::BUF_BYTES + skew * 4u + (uint32_t)(lane * CH * 4);
        mbar_wait(bar_sa + buf * 8, (parity >> buf) & 1u);
        parity ^= 1u << buf;
#pragma unroll
        for (int j = 0; j < NACC; j++) {
            if (32 * j >= n) break;  // warp-uniform
            // a frame past the end of a short callback reads stale but in-bounds shared memory and is masked
            const uint32_t a_sa = x_sa + (uint32_t)(32 * j * CH * 4);
            const bool valid = 32 * j + lane < n;
#pragma unroll
            for (int c = 0; c < CH; c++) {
                const float a = lds_f32(a_sa + (uint32_t)(c * 4)), b = lds_f32(a_sa + (uint32_t)((CH + c) * 4));
                float v = a + fract * (b - a);      // frame::lerp (frame.rs:39-41)
                v = v * fg;                         // FixedGain (gain.rs:35)
                v = v * g;                          // Gain, steady state (gain.rs:112-114)
                if (!valid) v = 0.0f;
                acc[j][c] = acc[j][c] + v;          // frame::mix (frame.rs:44-46, mixer.rs:115)
            }
        }
        __syncwarp();  // every lane is done with this buffer before it is refilled
        buf ^= 1u;
        cur = nxt;
    }
    // fold: warp -> CTA (fixed order) -> one partial tile per CTA
    float* tile = reinterpret_cast<float*>(smem_raw + warp * C::WARP_BYTES);
#pragma unroll
    for (int j = 0; j < NACC; j++)
#pragma unroll
        for (int c = 0; c < CH; c++) tile[(32 * j + lane) * CH + c] = acc[j][c];
    __syncthreads();
    float* dst = partials + ((size_t)tl * gridDim.x + blockIdx.x) * (ODB_MIXER_CHUNK * CH);
    for (int f = threadIdx.x; f < ODB_MIXER_CHUNK * CH; f += C::WARPS * 32) {
        float sum = 0.0f;
#pragma unroll
        for (int w = 0; w < C::WARPS; w++) sum = sum + *reinterpret_cast<const float*>(smem_raw + w * C::WARP_BYTES + f * 4);
        dst[f] = sum;
    }
}

// ------------------------------------------------------------------------------------------
// Staged resampling kernel: FramesSignal's serial-cursor path (frames.rs:189-196) for chains with ds != 1
// (Speed, PCM at another rate) and a gain at rest. Same construction as k_mix_fast: a warp takes 8
// (source, chunk) jobs; lanes 0..7 walk the 1024 literal `offset += ds` steps of one job each and store
// every 4th cursor value; then the jobs are consumed one by one from double-buffered TMA windows, lane l
// owning frames l, l+32, ...: checkpoint + <= 3 literal additions, round-down magic add for index and
// fraction, gather, lerp, FixedGain, Gain, accumulate. Unfused arithmetic: bit-identical per source.
template <int CH>
struct MixerResample {
    static constexpr int WARPS = 8;
    static constexpr int BATCH = 8;
    static constexpr int POINTS = ODB_MIXER_CHUNK / 4;
    static constexpr int ROW_BYTES = POINTS * 4 + 4;                            // +4 skews the banks between the 8 chain lanes
    static constexpr int BUF_BYTES = ODB_MIXER_RESAMPLE_CAP * 4;                // 8448
    static constexpr int WARP_BYTES = 2 * BUF_BYTES + ((BATCH * ROW_BYTES + 15) / 16) * 16;  // 25120
    static constexpr int SMEM_BYTES = WARPS * WARP_BYTES + WARPS * 16;
    static_assert(WARP_BYTES >= ODB_MIXER_CHUNK * CH * 4, "the warp region doubles as its partial tile");
};

template <int CH>
__global__ void __launch_bounds__(MixerResample<CH>::WARPS * 32) k_mixer_resample(const OdbMixJob* __restrict__ jobs, int n_sources,
                                                                                 float* __restrict__ partials,
                                                                                 const uint32_t* __restrict__ counters) {
    typedef MixerResample<CH> C;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    if (counters[ODB_CNT_RESAMPLE] == 0) return;  // k_reduce_tiles reads the same counter and skips our tiles
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tl = blockIdx.y;
    const uint32_t buf_sa = smem_u32(smem_raw) + (uint32_t)(warp * C::WARP_BYTES);
    const uint32_t offs_sa = buf_sa + 2 * C::BUF_BYTES;
    const uint32_t bar_sa = smem_u32(smem_raw) + (uint32_t)(C::WARPS * C::WARP_BYTES + warp * 16);
    if (lane == 0) {
        mbar_init(bar_sa, 1);
        mbar_init(bar_sa + 8, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    constexpr int NACC = ODB_MIXER_CHUNK / 32;
    float acc[NACC][CH];
#pragma unroll
    for (int j = 0; j < NACC; j++)
#pragma unroll
        for (int c = 0; c < CH; c++) acc[j][c] = 0.0f;
    const OdbMixJob* tile_jobs = jobs + (size_t)tl * n_sources;
    const int gw = blockIdx.x * C::WARPS + warp, GW = gridDim.x * C::WARPS;
    const int n_batches = (n_sources + C::BATCH - 1) / C::BATCH;
    const int r = lane & 3;
    uint32_t parity = 0, buf = 0;

    auto start_copy = [&](int sidx, uint32_t b) {  // lane 0: window of job sidx -> buffer b (16-byte aligned start)
        if (lane == 0) {
            const OdbMixJob* job = tile_jobs + sidx;
            const long long first = (long long)job->base * CH, start = first & ~3ll;
            const int reach = (int)((double)job->off0 + (double)(job->n_frames - 1) * (double)job->ds) + 4;
            const uint32_t floats = (uint32_t)((first - start) + (reach + 2) * CH + 3) & ~3u;
            mbar_expect_tx(bar_sa + b * 8, floats * 4u);
            bulk_g2s(buf_sa + b * C::BUF_BYTES, job->pcm + start, floats * 4u, bar_sa + b * 8);
        }
    };

    for (int bi = gw; bi < n_batches; bi += GW) {
        const int s0 = bi * C::BATCH;
        uint32_t act = 0;
#pragma unroll
        for (int q = 0; q < C::BATCH; q++)
            if (s0 + q < n_sources && (tile_jobs[s0 + q].flags & (ODB_JF_SKIP | ODB_JF_GENERAL | ODB_JF_RESAMPLE)) == ODB_JF_RESAMPLE)
                act |= 1u << q;
        if (!act) continue;
        start_copy(s0 + __ffs(act) - 1, buf);
        if (lane < C::BATCH && ((act >> lane) & 1u)) {  // literal cursor chains, one lane per job
            const OdbMixJob* job = tile_jobs + s0 + lane;
            float o = job->off0;
            const float ds = job->ds;
            const uint32_t dst = offs_sa + (uint32_t)(lane * C::ROW_BYTES);
#pragma unroll 8
            for (int m = 0; m < C::POINTS; m++) {
                sts_f32(dst + (uint32_t)(m * 4), o);
                o = __fadd_rn(o, ds); o = __fadd_rn(o, ds); o = __fadd_rn(o, ds); o = __fadd_rn(o, ds);
            }
        }
        __syncwarp();
        for (int q = 0; q < C::BATCH; q++) {
            if (!((act >> q) & 1u)) continue;
            const uint32_t rest = act >> (q + 1);
            if (rest) start_copy(s0 + q + __ffs(rest), buf ^ 1u);
            const OdbMixJob* job = tile_jobs + s0 + q;
            const float ds = job->ds, fg = job->fixed_gain, g = job->g;
            const int n = job->n_frames;
            const uint32_t skew = (uint32_t)(((long long)job->base * CH) & 3ll);
            // shared address of frame `base`, minus the magic bits the index arrives with
            const uint32_t K = buf_sa + buf * C::BUF_BYTES + skew * 4u - (ODB_MAGIC_BITS * (uint32_t)(CH * 4));
            const uint32_t row_sa = offs_sa + (uint32_t)(q * C::ROW_BYTES + (lane >> 2) * 4);
            const float d1 = r >= 1 ? ds : 0.0f, d2 = r >= 2 ? ds : 0.0f, d3 = r >= 3 ? ds : 0.0f;
            mbar_wait(bar_sa + buf * 8, (parity >> buf) & 1u);
            parity ^= 1u << buf;
            constexpr int ILP = 4;
#pragma unroll
            for (int j0 = 0; j0 < NACC; j0 += ILP) {
                if (32 * j0 >= n) break;  // warp-uniform
                float o[ILP], fr[ILP];
                uint32_t ad[ILP];
#pragma unroll
                for (int u = 0; u < ILP; u++) o[u] = lds_f32(row_sa + (uint32_t)(32 * (j0 + u)));  // checkpoint (32j + lane) & ~3
#pragma unroll
                for (int u = 0; u < ILP; u++) o[u] = __fadd_rn(__fadd_rn(__fadd_rn(o[u], d1), d2), d3);
#pragma unroll
                for (int u = 0; u < ILP; u++) {
                    const float t = __fadd_rd(o[u], ODB_MAGIC);         // 2^23 + trunc(offset), offset >= 0
                    fr[u] = __fadd_rn(o[u], -__fadd_rn(t, -ODB_MAGIC)); // offset - trunc as f32 (frames.rs:193)
                    ad[u] = K + __float_as_uint(t) * (uint32_t)(CH * 4);
                }
#pragma unroll
                for (int u = 0; u < ILP; u++) {
                    const bool valid = 32 * (j0 + u) + lane < n;
#pragma unroll
                    for (int c = 0; c < CH; c++) {
                        const float a = lds_f32(ad[u] + (uint32_t)(c * 4)), b = lds_f32(ad[u] + (uint32_t)((CH + c) * 4));
                        float v = a + fr[u] * (b - a);  // frame::lerp (frame.rs:39-41)
                        v = v * fg;                     // FixedGain (gain.rs:35)
                        v = v * g;                      // Gain at rest (gain.rs:112-114)
                        if (!valid) v = 0.0f;
                        acc[j0 + u][c] = acc[j0 + u][c] + v;
                    }
                }
            }
            __syncwarp();
            buf ^= 1u;
        }
        __syncwarp();
    }
    float* tile = reinterpret_cast<float*>(smem_raw + warp * C::WARP_BYTES);
#pragma unroll
    for (int j = 0; j < NACC; j++)
#pragma unroll
        for (int c = 0; c < CH; c++) tile[(32 * j + lane) * CH + c] = acc[j][c];
    __syncthreads();
    float* dst = partials + ((size_t)tl * gridDim.x + blockIdx.x) * (ODB_MIXER_CHUNK * CH);
    for (int f = threadIdx.x; f < ODB_MIXER_CHUNK * CH; f += C::WARPS * 32) {
        float sum = 0.0f;
#pragma unroll
        for (int w = 0; w < C::WARPS; w++) sum = sum + *reinterpret_cast<const float*>(smem_raw + w * C::WARP_BYTES + f * 4);
        dst[f] = sum;
    }
}

// ------------------------------------------------------------------------------------------
// Literal kernel: exact for every parameter combination (any ds, negative offsets, out-of-range
// indices, gain ramps). One warp per (source, tile); lane c < CH walks channel c's chain exactly as
// FramesSignal::sample / FixedGain::sample / Gain::sample do and parks the samples in a warp-private
// shared tile, which all lanes then fold into their register accumulators.
template <int CH, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) k_mixer_general(const OdbMixJob* __restrict__ jobs, int n_sources,
                                                              float* __restrict__ partials, int only_flagged,
                                                              const uint32_t* __restrict__ counters) {
    extern __shared__ float smem[];
    if (only_flagged && counters[ODB_CNT_GENERAL] == 0) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tl = blockIdx.y;
    const int gw = blockIdx.x * WARPS + warp, GW = gridDim.x * WARPS;
    constexpr int TILE = ODB_MIXER_CHUNK * CH;
    float* tile = smem + warp * TILE;
    float acc[TILE / 32];
#pragma unroll
    for (int j = 0; j < TILE / 32; j++) acc[j] = 0.0f;

    for (int sidx = gw; sidx < n_sources; sidx += GW) {
        const OdbMixJob* job = jobs + (size_t)tl * n_sources + sidx;
        const uint32_t jf = job->flags;
        if (jf & ODB_JF_SKIP) continue;
        if (only_flagged && !(jf & ODB_JF_GENERAL)) continue;
        const int n = job->n_frames;
        if (lane < CH) {
            const float* __restrict__ pcm = job->pcm;
            const long long len = job->len, base = job->base, pad_frames = ODB_PCM_PAD / CH;
            const float ds = job->ds, fg = job->fixed_gain;
            const bool unit = (jf & ODB_JF_FAST_L) != 0, ramp = (jf & ODB_JF_RAMP) != 0, cycle = (jf & ODB_JF_CYCLE) != 0;
            const float g = job->g, gprev = job->gprev, gnext = job->gnext, gstep = job->gstep;
            float gprog = job->gprog;
            float offset = job->off0;
            unsigned long long cbase = (unsigned long long)(base < 0 ? 0 : base);  // Cycle: `base`, reset to 0 by a wrap
            for (int i = 0; i < ODB_MIXER_CHUNK; i++) {
                float v = 0.0f;
                if (i < n && cycle) {                              // cycle.rs:30-51
                    const unsigned long long ulen = (unsigned long long)len;
                    const unsigned long long tr = (unsigned long long)offset;
                    const float fract = offset - (float)tr;
                    unsigned long long x = cbase + tr;
                    if (x >= ulen) {                               // :38-47 (fract is the one computed before the wrap)
                        cbase = 0;
                        offset = (float)(x % ulen) + fract;
                        x = (unsigned long long)offset;
                    }
                    const float a = pcm[x * CH + lane];
                    const float b = x < ulen - 1 ? pcm[(x + 1) * CH + lane] : pcm[lane];   // :34-37 / :42-46 wrap to frames[0]
                    v = a + fract * (b - a);                       // frame.rs:39-41
                    offset = offset + ds;                          // :50
                    v = v * fg;                                    // gain.rs:35
                    if (ramp) {                                    // gain.rs:118-121
                        v = v * (gprev + gprog * (gnext - gprev));
                        gprog = fminf(gprog + gstep, 1.0f);
                    } else {
                        v = v * g;
                    }
                } else if (i < n) {
                    long long k;
                    float fract;
                    if (unit) {                                    // frames.rs:183-187
                        k = base + i;
                        fract = offset;
                    } else {                                       // frames.rs:189-196
                        const long long tr = (long long)offset;
                        k = base + tr;
                        fract = offset - (float)tr;
                        offset = offset + ds;
                    }
                    float a = 0.0f, b = 0.0f;                      // get_pair, frames.rs:105-123 (zeros from the padding)
                    if (k >= -(pad_frames - 1) && k < len + pad_frames - 2) {
                        a = pcm[k * CH + lane];
                        b = pcm[(k + 1) * CH + lane];
                    }
                    v = a + fract * (b - a);                       // frame.rs:39-41
                    v = v * fg;                                    // gain.rs:35
                    if (ramp) {                                    // gain.rs:118-121
                        v = v * (gprev + gprog * (gnext - gprev));
                        gprog = fminf(gprog + gstep, 1.0f);
                    } else {
                        v = v * g;                                 // gain.rs:112-114 (x * 1.0 == x)
                    }
                }
                tile[i * CH + lane] = v;
            }
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < TILE / 32; j++) acc[j] = acc[j] + tile[32 * j + lane];
        __syncwarp();
    }
#pragma unroll
    for (int j = 0; j < TILE / 32; j++) tile[32 * j + lane] = acc[j];
    __syncthreads();
    float* dst = partials + ((size_t)tl * gridDim.x + blockIdx.x) * TILE;
    for (int f = threadIdx.x; f < TILE; f += WARPS * 32) {
        float sum = 0.0f;
#pragma unroll
        for (int w = 0; w < WARPS; w++) sum = sum + smem[w * TILE + f];
        dst[f] = sum;
    }
}

}  // namespace odbk

using namespace odbk;

static const int MIX_WARPS = 8;

void odb_launch_walk_mixer(OdbSource* src, const uint32_t* order, OdbMixJob* jobs, uint32_t* removed, int removed_cap,
                           uint32_t* counters, const OdbCallback& cb, cudaStream_t st) {
    if (cb.n_sources <= 0) return;
    k_walk_mixer<<<(cb.n_sources + 127) / 128, 128, 0, st>>>(src, order, jobs, removed, removed_cap, counters, cb);
}

int odb_mixer_ctas(int n_sources, int sm_count, int per_sm) {
    int want = (n_sources + MIX_WARPS - 1) / MIX_WARPS;
    int cap = sm_count * per_sm;
    return want < 1 ? 1 : (want > cap ? cap : want);
}

template <class K>
static cudaError_t set_smem(K kernel, int bytes) {
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}

cudaError_t odb_launch_mixer_unit(const OdbMixJob* jobs, int n_sources, int n_tiles, int channels, float* partials,
                                  int n_ctas, cudaStream_t st) {
    dim3 grid(n_ctas, n_tiles);
    cudaError_t e;
    if (channels == 1) {
        if ((e = set_smem(k_mixer_unit<1>, MixerStream<1>::SMEM_BYTES)) != cudaSuccess) return e;
        k_mixer_unit<1><<<grid, MixerStream<1>::WARPS * 32, MixerStream<1>::SMEM_BYTES, st>>>(jobs, n_sources, partials);
    } else {
        if ((e = set_smem(k_mixer_unit<2>, MixerStream<2>::SMEM_BYTES)) != cudaSuccess) return e;
        k_mixer_unit<2><<<grid, MixerStream<2>::WARPS * 32, MixerStream<2>::SMEM_BYTES, st>>>(jobs, n_sources, partials);
    }
    return cudaGetLastError();
}

int odb_mixer_resample_ctas(int n_sources, int sm_count) {
    int want = (n_sources + 63) / 64;
    return want < 1 ? 1 : (want > sm_count ? sm_count : want);
}
cudaError_t odb_launch_mixer_resample(const OdbMixJob* jobs, int n_sources, int n_tiles, int channels, float* partials,
                                      int n_ctas, const uint32_t* counters, cudaStream_t st) {
    dim3 grid(n_ctas, n_tiles);
    cudaError_t e;
    if (channels == 1) {
        if ((e = set_smem(k_mixer_resample<1>, MixerResample<1>::SMEM_BYTES)) != cudaSuccess) return e;
        k_mixer_resample<1><<<grid, MixerResample<1>::WARPS * 32, MixerResample<1>::SMEM_BYTES, st>>>(jobs, n_sources, partials, counters);
    } else {
        if ((e = set_smem(k_mixer_resample<2>, MixerResample<2>::SMEM_BYTES)) != cudaSuccess) return e;
        k_mixer_resample<2><<<grid, MixerResample<2>::WARPS * 32, MixerResample<2>::SMEM_BYTES, st>>>(jobs, n_sources, partials, counters);
    }
    return cudaGetLastError();
}

cudaError_t odb_launch_mixer_general(const OdbMixJob* jobs, int n_sources, int n_tiles, int channels, float* partials,
                                     int n_ctas, int only_flagged, const uint32_t* counters, cudaStream_t st) {
    dim3 grid(n_ctas, n_tiles);
    const int smem = MIX_WARPS * ODB_MIXER_CHUNK * channels * (int)sizeof(float);
    cudaError_t e;
    if (channels == 1) {
        if ((e = set_smem(k_mixer_general<1, MIX_WARPS>, smem)) != cudaSuccess) return e;
        k_mixer_general<1, MIX_WARPS><<<grid, MIX_WARPS * 32, smem, st>>>(jobs, n_sources, partials, only_flagged, counters);
    } else {
        if ((e = set_smem(k_mixer_general<2, MIX_WARPS>, smem)) != cudaSuccess) return e;
        k_mixer_general<2, MIX_WARPS><<<grid, MIX_WARPS * 32, smem, st>>>(jobs, n_sources, partials, only_flagged, counters);
    }
    return cudaGetLastError();
}
