// walk_set for the seek set (spatial.rs:191-265) plus everything of the mix closure (spatial.rs:445-469) that is
// O(1) per source and chunk: ear states, dt, d_gain and the f64 cursor bookkeeping of FramesSignal::seek/sample
// (frames.rs:176-213) - as a block-level device function, used by k_walk_seek (odb_spatial.cu) and by the first
// phase of the one-launch callback kernel (odb_scene_mix.cu).
//
// Two threads per source, one per ear: the work of one source is a long chain of dependent high-latency operations
// (IEEE divides and square roots, f64 conversions), and the two ears are independent once the source's motion is
// smoothed. Both threads evaluate the shared part (motion smoothing, listener rotation) redundantly and
// bit-identically; the left ear's thread writes the source's state. A thread reproduces the f64 cursor at the start
// of each of its chunks with the reference's own sequence of additions (seek(prev.offset), then
// `t += f64(dt) * f64(m)` per earlier chunk, the rewinding seek between the ears), so every (base, offset) pair is
// the one the serial code computes.
//
// Memory traffic is staged through shared memory: the block first gathers the 160-byte records of its sources with
// coalesced 16-byte loads, every thread works on the shared copy, and records and job lines go back to HBM as whole
// lines (one 128-byte OdbJob per 32 lanes). Round 1's kernel had every thread load its own record and store its
// job fields one 4-byte word at a time: ~25 store instructions per thread, each touching 16-32 different lines - the
// kernel was bound by those wavefronts (15 us for 65 536 sources), not by its arithmetic.
#pragma once
#include "odb_kernels.h"
#include "odb_math.cuh"

namespace odbk {

constexpr int WALK_REC_WORDS = (int)(sizeof(OdbSource) / 16);  // 16-byte words per source record

// Shared memory a block of THREADS threads needs: records + one job line per source + slot ids.
template <int THREADS>
struct WalkSmem {
    static constexpr int SOURCES = THREADS / 2;
    static constexpr int REC_BYTES = SOURCES * (int)sizeof(OdbSource);
    static constexpr int JOB_STRIDE = (int)sizeof(OdbJob) + 16;             // padded: the threads of a warp write the same field of
                                                                            // 16 different lines (stride 36 words: 4-way instead of 16-way conflicts)
    static constexpr int JOB_BYTES = SOURCES * JOB_STRIDE;
    static constexpr int BYTES = REC_BYTES + JOB_BYTES + SOURCES * 4 + 16;  // + the block's two job counters
};

// Walks sources [first, first + THREADS / 2) of the active list (indices beyond cb.n_sources are idle lanes).
// Must be called by all THREADS threads of the block (it synchronises the block); `tid` = thread index within them.
template <int THREADS, bool LATE_WAIT>
__device__ __forceinline__ void walk_seek_block(OdbSource* __restrict__ src, const uint32_t* __restrict__ order,
                                                OdbJob* __restrict__ jobs, uint32_t* __restrict__ removed, const int removed_cap,
                                                uint32_t* __restrict__ counters, uint32_t* __restrict__ zero_counters,
                                                const OdbCallback& cb, const int first, unsigned char* __restrict__ smem, const int tid) {
    typedef WalkSmem<THREADS> SM;
    OdbSource* rec_sm = reinterpret_cast<OdbSource*>(smem);
    OdbJob* job_sm = reinterpret_cast<OdbJob*>(smem + SM::REC_BYTES);
    uint32_t* slot_sm = reinterpret_cast<uint32_t*>(smem + SM::REC_BYTES + SM::JOB_BYTES);
    uint32_t* cnt_sm = slot_sm + SM::SOURCES;
    const int n_here = min(SM::SOURCES, cb.n_sources - first);  // live sources of this block (may be <= 0)
    if (tid < 2) cnt_sm[tid] = 0u;

    // ---- gather the records -----------------------------------------------------------------------------------
    for (int i = tid; i < SM::SOURCES; i += THREADS) slot_sm[i] = i < n_here ? order[first + i] : 0xFFFFFFFFu;
    __syncthreads();
    for (int w = tid; w < SM::SOURCES * WALK_REC_WORDS; w += THREADS) {
        const int i = w / WALK_REC_WORDS, k = w - i * WALK_REC_WORDS;
        if (i < n_here)
            reinterpret_cast<uint4*>(rec_sm + i)[k] = __ldcg(reinterpret_cast<const uint4*>(src + slot_sm[i]) + k);
    }
    __syncthreads();

    const int il = tid >> 1, e = tid & 1;    // source of this block, ear
    const bool live = il < n_here;
    const int idx = first + il;
    const uint32_t slot = live ? slot_sm[il] : 0u;
    OdbSource* sp = rec_sm + (live ? il : 0);  // state updates go to the shared copy (the left ear's thread writes)
    OdbSource s;
    {
        uint4* d = reinterpret_cast<uint4*>(&s);
        const uint4* g = reinterpret_cast<const uint4*>(sp);
#pragma unroll
        for (int i = 0; i < WALK_REC_WORDS; i++) d[i] = live ? g[i] : make_uint4(0u, 0u, 0u, 0u);  // idle lanes compute on zeros
    }
    __syncthreads();  // every thread holds its copy before the left ears start updating the shared records
    OdbJob* jm = reinterpret_cast<OdbJob*>(reinterpret_cast<unsigned char*>(job_sm) + il * SM::JOB_STRIDE);
    const int n = cb.n_frames;
    const float elapsed = cb.elapsed;
    const int nt = cb.n_tiles, ns = cb.job_stride;
    const bool leader = live && e == 0;
    V3 prev_position, next_position;
    uint32_t flags;
    const bool mixing = walk_common(sp, s, cb, slot, removed, removed_cap, prev_position, next_position, flags, leader);
    const bool emit = mixing && live;
    const double rate = s.rate;

    // --- mix closure set-up, spatial.rs:446-468: this thread's ear
    const float nf = (float)n;
    const float ratef = (float)rate;  // `self.data.rate as f32` frames.rs:178
    const int n_chunks = (n + ODB_SPATIAL_CHUNK - 1) / ODB_SPATIAL_CHUNK;
    const EarSt ps = ear_state(prev_position, e, s.radius);
    const EarSt nx = ear_state(next_position, e, s.radius);
    const float eff = (elapsed + nx.offset) - ps.offset;                // :451
    const float dt = eff / nf;                                          // :452
    const float d_gain = (nx.gain - ps.gain) / nf;                      // :453
    const float ds = dt * ratef;                                        // frames.rs:178
    const bool cycle = (s.flags & ODB_SF_CYCLE) != 0;                   // the inner signal is Cycle, not FramesSignal
    const bool fast = !cycle && fabsf(ds - 1.0f) <= ODB_F32_EPSILON;    // frames.rs:180
    const bool general = !fast && !(ds > 0.0f && ds <= ODB_FAST_DS_MAX);
    // the left ear's scalars, needed by the right ear's thread to replay the cursor up to its own start
    const unsigned full = 0xffffffffu;
    const int lane = tid & 31, src_lane0 = lane - e;
    const float ps_off_l = __shfl_sync(full, ps.offset, src_lane0);
    const float eff_l = __shfl_sync(full, eff, src_lane0);
    const float dt_l = __shfl_sync(full, dt, src_lane0);

    // Cycle (cycle.rs:26-61): `t` is the cursor in samples, and where a chunk starts depends on the f32 chain of the
    // one before it (and on where it wrapped), so the chains are walked here once without the taps; the literal mix
    // path walks each chunk again from the recorded (base, offset) with them. The right ear's thread replays the
    // left ear's pass first - the reference runs the ears one after the other on the same cursor (spatial.rs:446-466).
    const double dlen = (double)s.len;
    const unsigned long long ulen = (unsigned long long)(s.len > 0 ? s.len : 1);
    auto cyc_seek = [&](double cur, float seconds) {                    // cycle.rs:57-60
        const double r = fmod(cur + (double)seconds * rate, dlen);
        return r < 0.0 ? r + dlen : r;                                  // f64::rem_euclid
    };
    auto cyc_chunk = [&](double cur, float ds_e, int m, unsigned long long& base0, float& off0) {
        unsigned long long cbase = (unsigned long long)cur;            // :28
        float offset = (float)(cur - (double)cbase);                    // :29
        base0 = cbase; off0 = offset;
        for (int i = 0; i < m; i++) {
            const unsigned long long tr = (unsigned long long)offset;              // :31
            const float fract = offset - (float)tr;                                // :32
            const unsigned long long x = cbase + tr;                               // :33
            if (x >= ulen) { cbase = 0; offset = (float)(x % ulen) + fract; }      // :38-40
            offset = offset + ds_e;                                                // :50
        }
        return (double)cbase + (double)offset;                                     // :52
    };
    double cyc_cur = s.t;
    if (cycle && mixing) {
        if (e == 1) {
            cyc_cur = cyc_seek(cyc_cur, ps_off_l);                      // spatial.rs:449 (left ear)
            const float ds_l = dt_l * ratef;                            // cycle.rs:27
            for (int cg = 0; cg < n_chunks; cg++) {
                unsigned long long b0; float o0;
                cyc_cur = cyc_chunk(cyc_cur, ds_l, min(ODB_SPATIAL_CHUNK, n - cg * ODB_SPATIAL_CHUNK), b0, o0);
            }
            cyc_cur = cyc_seek(cyc_cur, -eff_l - ps_off_l);             // :465
        }
        cyc_cur = cyc_seek(cyc_cur, ps.offset);                         // :449
    }
    // cursor at the start of this thread's ear
    double t = s.t;
    if (e == 1) {  // left ear first: seek(prev.offset), all chunks, seek(-eff - prev.offset)  (:449-465)
        t = t + (double)ps_off_l;
        for (int cg = 0; cg < n_chunks; cg++) {
            const int m = min(ODB_SPATIAL_CHUNK, n - cg * ODB_SPATIAL_CHUNK);
            t = t + (double)dt_l * (double)m;                           // frames.rs:198
        }
        t = t + (double)(-eff_l - ps_off_l);
    }
    t = t + (double)ps.offset;                                          // :449 seek(prev.offset)

    // ---- tile by tile: this ear's chunks, the source's window per 512-frame half, one job line --------------------
    double tc = t;  // cursor at the start of the next chunk
    uint32_t n_general = 0, n_fast = 0;
    for (int tl = 0; tl < nt; tl++) {
        int wlo[2] = {0x7fffffff, 0x7fffffff}, whi[2] = {-0x7fffffff, -0x7fffffff};
        bool gen = false;
#pragma unroll
        for (int c = 0; c < ODB_TILE_CHUNKS; c++) {
            const int cg = tl * ODB_TILE_CHUNKS + c;
            if (cg < n_chunks) {
                const int m = min(ODB_SPATIAL_CHUNK, n - cg * ODB_SPATIAL_CHUNK);
                int base;
                float off0;
                bool g;
                if (cycle) {
                    unsigned long long b0 = 0; float o0 = 0.0f;
                    if (mixing) cyc_cur = cyc_chunk(cyc_cur, ds, m, b0, o0);
                    base = (int)b0; off0 = o0;
                    g = true;                                           // Cycle sources always take the literal path
                } else {
                    const double s0 = tc * rate;                        // frames.rs:177
                    // frames.rs:179 `s0 as isize`: 32-bit conversion (saturating); any |s0| >= 2^29 is far outside every
                    // Frames block and only ever yields zeros, which the literal path produces
                    base = __double2int_rz(s0);
                    off0 = (float)(s0 - (double)base);                  // frames.rs:183 / :189
                    g = general || off0 < 0.0f;                         // negative-fract quirk
                    if (base > (1 << 29) || base < -(1 << 29)) { g = true; base = base > 0 ? (1 << 30) : -(1 << 30); }
                    tc = tc + (double)dt * (double)m;                   // frames.rs:198
                }
                if (emit) {
                    jm->base[e][c] = base;
                    jm->off0[e][c] = off0;
                }
                // PCM indices this chain can read: [base, base + trunc(offset_{m-1}) + 1]; the f32 chain stays within
                // 1e-2 of off0 + (m-1)*ds for m <= 256, so +4 on the f32 estimate is a safe upper bound.
                const float span = g ? 0.0f : __fmaf_rn((float)(m - 1), ds, off0);
                const int last = fast ? base + m : base + __float2int_rz(span) + 4;
                const int h = c / ODB_FAST_HALF_CHUNKS;
                wlo[h] = min(wlo[h], base);
                whi[h] = max(whi[h], last);
                gen = gen || g;
            }
        }
        uint32_t f = gen ? ODB_JF_GENERAL : 0u;
        int ws[2], wl[2];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            int lo = min(wlo[h], __shfl_xor_sync(full, wlo[h], 1));     // over the two ears of this source
            int hi = max(whi[h], __shfl_xor_sync(full, whi[h], 1));
            ws[h] = 0; wl[h] = 0;
            if (hi >= lo) {  // the half has frames
                ws[h] = lo & ~3;
                wl[h] = ((hi - ws[h] + 1) + 3) & ~3;
                if (wl[h] > ODB_FAST_PCM_CAP || ws[h] < -ODB_PCM_PAD || ws[h] + wl[h] > s.len + ODB_PCM_PAD) f |= ODB_JF_GENERAL;
            }
        }
        f |= fast ? (e == 0 ? ODB_JF_FAST_L : ODB_JF_FAST_R) : 0u;
        f |= __shfl_xor_sync(full, f, 1);
        if (flags & ODB_SF_FIXED_GAIN) f |= ODB_JF_FIXED_GAIN | ODB_JF_GENERAL;
        if (cycle) f |= ODB_JF_CYCLE | ODB_JF_GENERAL;
        if (cb.force_general) f |= ODB_JF_GENERAL;
        if (emit) {
            jm->ds[e] = ds; jm->pg[e] = ps.gain; jm->dg[e] = d_gain;
            if (e == 0) {
                jm->window[0][0] = ws[0]; jm->window[0][1] = (f & ODB_JF_GENERAL) ? 0 : wl[0];
                jm->window[1][0] = ws[1]; jm->window[1][1] = (f & ODB_JF_GENERAL) ? 0 : wl[1];
                jm->pcm = s.pcm; jm->len = s.len;
                jm->fixed_gain = s.fixed_gain;
                jm->n_frames = min(ODB_TILE_FRAMES, n - tl * ODB_TILE_FRAMES);
                jm->flags = f;
                if (f & ODB_JF_GENERAL) n_general++; else n_fast++;
            }
        } else if (leader) {
            jm->flags = ODB_JF_SKIP;
        }
        __syncthreads();
        // the tile's job lines of this block: lane = word, one whole line per 32 lanes
        {
            OdbJob* dst = jobs + (size_t)tl * ns + cb.job_offset + first;
            for (int w = tid; w < n_here * 32; w += THREADS)
                __stcg(reinterpret_cast<uint32_t*>(dst) + w,
                       reinterpret_cast<const uint32_t*>(reinterpret_cast<const unsigned char*>(job_sm) + (w >> 5) * SM::JOB_STRIDE)[w & 31]);
        }
        __syncthreads();
    }
    if (emit && e == 1 && cycle) {
        cyc_cur = cyc_seek(cyc_cur, -eff - ps.offset);                  // :465
        cyc_cur = cyc_seek(cyc_cur, elapsed);                           // :468
        sp->t = cyc_cur;
    } else if (emit && e == 1) {  // right ear: finish the cursor, :465-468 (tc has advanced over every chunk)
        // frames.rs:199-200 stores (t * rate) as isize at the end of every sample() call; only the last store
        // (right ear, last chunk) is observable, and the seeks that follow do not touch sample_t
        if (n_chunks > 0) sp->sample_t = (long long)(tc * rate);
        double te = tc + (double)(-eff - ps.offset);                    // :465
        te = te + (double)elapsed;                                      // :468
        sp->t = te;
    }
    __syncthreads();
    // ---- scatter the updated records back -------------------------------------------------------------------------
    for (int w = tid; w < SM::SOURCES * WALK_REC_WORDS; w += THREADS) {
        const int i = w / WALK_REC_WORDS, k = w - i * WALK_REC_WORDS;
        if (i < n_here) reinterpret_cast<uint4*>(src + slot_sm[i])[k] = reinterpret_cast<const uint4*>(rec_sm + i)[k];
    }
    // job counters: one atomic per warp into shared memory, one per block into HBM. Everything above only touches
    // what no earlier kernel still uses (this callback's job records, the source table). LATE_WAIT (round 1's
    // multi-kernel callback): the counters are reset by the previous callback's reduce kernel, so this is where the grid
    // waits for it. Otherwise the counters live in a ring of four sets: this grid resets the set of the callback after
    // next (`zero_counters`), which that callback's walk cannot reach before this grid has completed (the launch
    // dependencies walk k -> mix k -> walk k + 1 -> mix k + 1 -> walk k + 2 are each "complete before start").
    if (LATE_WAIT) pdl_wait();
    if (zero_counters && first == 0 && tid < ODB_CNT_WORDS) zero_counters[tid] = 0u;
    n_general = __reduce_add_sync(full, n_general);
    n_fast = __reduce_add_sync(full, n_fast);
    if (lane == 0 && n_general) atomicAdd(cnt_sm + 0, n_general);
    if (lane == 0 && n_fast) atomicAdd(cnt_sm + 1, n_fast);
    __syncthreads();
    if (tid == 0 && cnt_sm[0]) atomicAdd(counters + ODB_CNT_GENERAL, cnt_sm[0]);
    if (tid == 1 && cnt_sm[1]) atomicAdd(counters + ODB_CNT_FAST, cnt_sm[1]);
}

}  // namespace odbk
