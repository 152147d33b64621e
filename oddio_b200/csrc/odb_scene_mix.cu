// One-launch scene callback for the seek path: SpatialScene's mix closure (spatial.rs:445-469) over
// FramesSignal::sample (frames.rs:176-201) for every source of the scene, the literal path for the jobs the
// walk kernel flagged, the sum of the per-CTA partial tiles, the Tanh / Reinhard epilogue and - for a source-sharded
// scene - the exchange of the tile with the other GPUs: what used to be k_mix_fast + k_mix_general + k_reduce_tiles
// (+ k_exchange_push / _pull), three to five launches and as many grid-wide dependencies, in one persistent grid.
//
// A persistent grid of one 16-warp CTA per SM; the two warps of a team mix the two 512-frame halves of the same
// sources. Per 1024-frame tile of the callback (the tile loop is inside the kernel, so a callback may have any length):
//   1. batches of up to 8 sources per warp, as in k_mix_fast (odb_mix_fast.cu): job records staged in bank-swizzled
//      shared memory, the PCM window of each source fetched by one elected lane with a bulk async copy (TMA, UBLKCP)
//      into one of two buffers, the reference's serial cursor `offset += ds` (frames.rs:195) walked literally on 32
//      lanes (source x ear x chunk) with every 4th value parked in shared memory;
//   2. consume (shipped shape, SmxCfg<2, 16, 0, 4>): lane l owns frames l + 32 j of a 256-frame chunk; it reloads its
//      checkpoint, applies <= 3 more literal additions, splits the cursor into index and fraction with a round-down
//      magic add (no F2I / I2F), gathers the sample pair from shared memory, lerps (frame.rs:39-41), applies the
//      per-frame gain ramp (spatial.rs:459) and accumulates into packed (L, R) register accumulators (FADD2 / FFMA2).
//      Alternative shapes measured on C3 and kept selectable for experiments: frame pairs per
//      lane (3 cursor steps per 2 frames, but 2-way bank conflicts on the taps), one warp per whole tile (12 warps);
//   3. tail: jobs flagged ODB_JF_GENERAL (any ds, windows outside the block, FixedGain, Cycle) are mixed literally by
//      the warp that owns their batch, straight into its parked accumulators (general_part, the code of
//      k_mix_general for one part of the tile);
//   4. fold warp -> CTA in fixed order, one partial tile per CTA to HBM, then a grid-wide arrive counter: when every
//      CTA of the grid has arrived, CTA b sums 16-float slices b, b + grid, ... of all partial tiles in index order
//      (deterministic; no float atomics) and either applies the epilogue and stores the output, or leaves the raw sum
//      for the grid's last CTA, which finishes the callback while all other CTAs have left their SMs: the tile into
//      pinned host memory + the host's completion flag, or the push / pull of the multi-GPU exchange over NVLink
//      peer memory (odb_exchange.h). All CTAs are resident (grid <= SM count, one CTA per SM), so the arrive wait
//      cannot deadlock.
// Ordering against the kernels around it is device-side (`walked`, `completed`): see the comments at the top of the
// kernel and DESIGN.md section 4.
#include <cuda_runtime.h>

#include <cstdlib>

#include "odb_kernels.h"
#include "odb_math.cuh"
#include "odb_async.cuh"
#include "odb_f32x2.cuh"

namespace odbk {

// Compile-time shape of the kernel. SPLIT: a 1024-frame tile is mixed as SPLIT parts, one warp each (2: warp pairs,
// 16 packed accumulators per lane; 1: one warp per tile, 32 accumulators, fewer and longer visits per source).
// LOOP: 0 = lane l owns frames l + 32 j (conflict-light taps, three literal steps per frame),
//       1 = lane l owns the frame pairs 64 j + 2 l, + 1 (three literal steps per two frames, 2-way tap conflicts).
// ILP: frames of a chunk a lane has in flight.
template <int SPLIT_, int WARPS_, int LOOP_, int ILP_>
struct SmxCfg {
    static constexpr int SPLIT = SPLIT_, WARPS = WARPS_, LOOP = LOOP_, ILP = ILP_;
    static constexpr bool WS = false;                                         // every warp does everything (see SmxWsCfg)
    static constexpr int CWARPS = WARPS_;                                     // warps that mix
    static constexpr int HCHUNKS = ODB_TILE_CHUNKS / SPLIT;                   // 256-frame chunks per part
    static constexpr int NACC = HCHUNKS * 8;                                  // packed (L, R) accumulators per lane
    static constexpr int BATCH = 16 / HCHUNKS;                                // sources per batch: BATCH x 2 ears x HCHUNKS = 32 chains
    static constexpr int PCM_FLOATS = ODB_FAST_PCM_CAP * HCHUNKS / 2;         // 640 per 512-frame half
    static constexpr int PCM_BYTES = PCM_FLOATS * 4;
    static constexpr int POINTS = ODB_SPATIAL_CHUNK / 4;                      // every 4th cursor value of a chunk
    static constexpr int ROW_BYTES = POINTS * 8 + 8;                          // (L, R) cursors of one (source, chunk); +8 skews the banks
    static constexpr int OFFS_BYTES = BATCH * HCHUNKS * ROW_BYTES;            // 8320
    static constexpr int WARP_BYTES = 2 * PCM_BYTES + OFFS_BYTES;
    static constexpr int PART_FRAMES = ODB_TILE_FRAMES / SPLIT;
    static constexpr int REC_BYTES = 128;
    static constexpr int JOBS_OFF = WARPS * WARP_BYTES;
    static constexpr int BARS_OFF = JOBS_OFF + WARPS * BATCH * REC_BYTES;
    static constexpr int SMEM_BYTES = BARS_OFF + WARPS * 16;
    static_assert(WARP_BYTES >= PART_FRAMES * 8 + 2 * PART_FRAMES * 4, "parked part of the tile + literal-path scratch");
    static_assert(WARP_BYTES % 16 == 0, "TMA destinations are 16-byte aligned");
    static_assert(SMEM_BYTES <= 232448, "fits the 227 KB a CTA may use");
    static_assert(WARPS % SPLIT == 0 && (WARPS * 32) % 16 == 0, "warps come in groups of SPLIT");
};
// Warp-specialised shape: CWARPS_ consumer warps (teams of two, one 512-frame half each, as SmxCfg<2, ..>) never leave
// the consume loop; CWARPS_ / 4 producer warps do what is serial or once-per-source for them - fetch the job records
// (bulk async copies into a private staging buffer), derive each half's window / tap offsets, walk the literal cursor
// chains (one producer pass = 4 sources of each of its 2 teams: 32 lanes x both ears packed = 64 chains) and hand the
// result over through a ring of S_ slots per consumer warp (record + cursor rows; mbarrier pairs full / empty per
// team and slot; tests/test_ws_handover_model.py checks the arithmetic of that hand-over on the CPU). The consumer
// still issues the bulk copy of its PCM windows itself (double-buffered, one source ahead). REGC_ / REGP_ != 0: the
// register file is re-divided with setmaxnreg (consumers REGC_, producers REGP_ registers per thread, out of the
// 96 x 640 the launch allocates: 16 x 32 x 112 + 4 x 32 x 32).
// Measured slower than SmxCfg<2, 16, 0, 4> on C3: an experiment behind ODB_SMX_CFG, not the default.
template <int CWARPS_, int S_, int REGC_, int REGP_>
struct SmxWsCfg {
    static constexpr bool WS = true;
    static constexpr int SPLIT = 2, LOOP = 0, ILP = 4;
    static constexpr int CWARPS = CWARPS_, PWARPS = CWARPS_ / 4, WARPS = CWARPS + PWARPS, S = S_, REGC = REGC_, REGP = REGP_;
    static constexpr int HCHUNKS = ODB_TILE_CHUNKS / SPLIT, NACC = HCHUNKS * 8, BATCH = 16 / HCHUNKS;
    static constexpr int PCM_FLOATS = ODB_FAST_PCM_CAP * HCHUNKS / 2, PCM_BYTES = PCM_FLOATS * 4;
    static constexpr int POINTS = ODB_SPATIAL_CHUNK / 4, ROW_BYTES = POINTS * 8 + 8;
    static constexpr int SLOT_ROWS = HCHUNKS * ROW_BYTES;                     // cursor rows of one (source, half)
    static constexpr int REC_BYTES = 128;
    static constexpr int ROWS_OFF = 2 * PCM_BYTES, RECS_OFF = ROWS_OFF + S * SLOT_ROWS;
    static constexpr int WARP_BYTES = RECS_OFF + S * REC_BYTES;               // per consumer warp
    static constexpr int PART_FRAMES = ODB_TILE_FRAMES / SPLIT;
    static constexpr int STAGE_OFF = CWARPS * WARP_BYTES;                     // per producer warp: 2 x (2 PASS) raw job records
    static constexpr int PASS = 4;                                            // sources per team and producer pass
    static constexpr int STAGE_BYTES = 2 * 2 * PASS * REC_BYTES;
    static constexpr int BARS_OFF = STAGE_OFF + PWARPS * STAGE_BYTES;         // per consumer warp 2 (windows), per team 2 S, per producer 2
    static constexpr int TEAM_BARS_OFF = BARS_OFF + CWARPS * 16;
    static constexpr int PROD_BARS_OFF = TEAM_BARS_OFF + (CWARPS / 2) * 16 * S;
    static constexpr int SMEM_BYTES = PROD_BARS_OFF + PWARPS * 16;
    static_assert(CWARPS % 4 == 0, "a producer warp serves two teams of two consumer warps");
    static_assert(S >= PASS + 2, "a producer pass fills PASS slots per team while the consumer holds one (PASS + 1 is the minimum) - one to spare");
    static_assert(WARP_BYTES >= PART_FRAMES * 8 + 2 * PART_FRAMES * 4, "parked part of the tile + literal-path scratch");
    static_assert(WARP_BYTES % 16 == 0 && RECS_OFF % 16 == 0 && SLOT_ROWS % 8 == 0, "alignment of windows, records, rows");
    static_assert(SMEM_BYTES <= 232448, "fits the 227 KB a CTA may use");
    // setmaxnreg moves registers inside what the launch gave the CTA (96 x 32 x WARPS), between whole warpgroups
    static_assert(REGC == 0 || (CWARPS * 32 * REGC + PWARPS * 32 * REGP <= 96 * 32 * WARPS && PWARPS == 4 && WARPS == 20), "register pool of the CTA");
};
#define SMX_WS_SKIP 0x100u      // code bits of a slot record: nothing to mix for this half
#define SMX_WS_FLAGGED 0x200u   // the job takes the literal path (tail)
constexpr int SMX_SLICE = 16;                                              // output floats one reducing CTA sums at a time
constexpr int SMX_SLICES = 2 * ODB_TILE_FRAMES / SMX_SLICE;                // 128

// Words of a staged job record after the warp's prologue (see k_mix_fast).
#define SJ_SRC 0
#define SJ_BYTES 2
#define SJ_CODE 10
#define SJ_K 12

// One 256-frame chunk of one source. LOOP 0: lane l owns frames l + 32 j (j = 0..7); LOOP 1: the frame pairs
// (64 j + 2 l, + 1), j = 0..3 - either way 8 frames and 8 packed accumulators per lane and chunk; frame_of(u) is the
// chunk-relative frame of the lane's u-th accumulator minus the lane's own offset (lane or 2 lane).
// UL / UR: that ear is on FramesSignal's ds ~= 1 path (frames.rs:180-187). FULL: every frame of the chunk is inside
// the tile. KL / KR: doppler ear = shared address of PCM index `base` minus the magic bits; unit ear = shared address
// of PCM index base + the lane's offset. d1..d3: the literal steps between the lane's checkpoint and its frame
// (LOOP 0: ds or +0.0 by lane & 3; LOOP 1: d1 = d2 = (ds or +0.0 by lane & 1), d3 = ds for the second frame of a pair).
template <class CFG, bool STRICT, bool FULL, bool UL, bool UR>
__device__ __forceinline__ void consume_chunk(u64* __restrict__ acc, const int cc, const int c, const int lane, const float fbase,
                                              const uint32_t row_sa, const uint32_t KL, const uint32_t KR, const u64 d1,
                                              const u64 d2, const u64 d3, const u64 fr_unit, const u64 pgp, const u64 dgp,
                                              const int nfr, const u64 nz) {
    constexpr int ILP = CFG::ILP;
    constexpr int LSTEP = CFG::LOOP == 0 ? 1 : 2;                                       // frames per lane step
    auto frame_of = [](int u) { return CFG::LOOP == 0 ? 32 * u : 64 * (u >> 1) + (u & 1); };
    const u64 magic = pk2(ODB_MAGIC, ODB_MAGIC);
    const u64 g0 = STRICT ? 0ull : fma2(pk2(fbase, fbase), dgp, pgp);
#pragma unroll
    for (int j0 = 0; j0 < 8; j0 += ILP) {
        if (!FULL && c * ODB_SPATIAL_CHUNK + frame_of(j0) >= nfr) break;  // warp-uniform
        uint32_t aL[ILP], aR[ILP];
        u64 fr[ILP], o[ILP];
        if (!(UL && UR)) {
            if (CFG::LOOP == 0) {
                // cursor of frame k = 32 j + l: checkpoint k & ~3, then k & 3 literal `offset += ds` steps (frames.rs:195);
                // adding +0.0 is exact
#pragma unroll
                for (int u = 0; u < ILP; u++) o[u] = lds_u64(row_sa + (uint32_t)(64 * (j0 + u)));
#pragma unroll
                for (int u = 0; u < ILP; u++) o[u] = add2(o[u], d1);
#pragma unroll
                for (int u = 0; u < ILP; u++) o[u] = add2(o[u], d2);
#pragma unroll
                for (int u = 0; u < ILP; u++) o[u] = add2(o[u], d3);
            } else {
                // cursor of frame k = 64 j + 2 l: checkpoint k & ~3, then (k & 3) in {0, 2} literal steps; frame k + 1
                // is one more step
#pragma unroll
                for (int u = 0; u < ILP; u += 2) o[u] = lds_u64(row_sa + (uint32_t)(128 * ((j0 + u) >> 1)));
#pragma unroll
                for (int u = 0; u < ILP; u += 2) o[u] = add2(o[u], d1);
#pragma unroll
                for (int u = 0; u < ILP; u += 2) o[u] = add2(o[u], d2);
#pragma unroll
                for (int u = 0; u < ILP; u += 2) o[u + 1] = add2(o[u], d3);
            }
            // trunc = offset as isize; fract = offset - trunc as f32 (frames.rs:191-193): a round-down add of 2^23
            // leaves trunc(offset) in the low mantissa bits (offset >= 0 here). (An F2I.TRUNC / I2FP split was
            // measured 8 % slower: the conversions are not full rate on sm_100a.)
            u64 t[ILP];
#pragma unroll
            for (int u = 0; u < ILP; u++) t[u] = add2_rm(o[u], magic);
#pragma unroll
            for (int u = 0; u < ILP; u++) {
                uint32_t tL, tR;
                upk2u(t[u], tL, tR);
                aL[u] = KL + (tL << 2);
                aR[u] = KR + (tR << 2);
            }
#pragma unroll
            for (int u = 0; u < ILP; u++) fr[u] = sub2(o[u], sub2(t[u], magic));
        }
        if (UL || UR) {  // frames.rs:183-187
            float u0, u1;
            upk2(fr_unit, u0, u1);
#pragma unroll
            for (int u = 0; u < ILP; u++) {
                float f0, f1;
                if (UL && UR) { f0 = u0; f1 = u1; }
                else { upk2(fr[u], f0, f1); if (UL) f0 = u0; else f1 = u1; }
                fr[u] = pk2(f0, f1);
                if (UL) aL[u] = KL + (uint32_t)(4 * frame_of(j0 + u));
                if (UR) aR[u] = KR + (uint32_t)(4 * frame_of(j0 + u));
            }
        }
        u64 a[ILP], b[ILP];
#pragma unroll
        for (int u = 0; u < ILP; u++) {  // get_pair (frames.rs:105-123); zeros come from the arena padding
            a[u] = pk2(lds_f32(aL[u]), lds_f32(aR[u]));
            b[u] = pk2(lds_f32_4(aL[u]), lds_f32_4(aR[u]));
        }
        u64 g[ILP], s[ILP];
#pragma unroll
        for (int u = 0; u < ILP; u++) {
            if (STRICT) {
                const float fi = fbase + (float)frame_of(j0 + u);  // `i as f32` (spatial.rs:459); exact small integer
                g[u] = mulx(pk2(fi, fi), dgp, nz);
            } else {
                // prev_gain + i * d_gain as (prev_gain + i0 * d_gain) + (i - i0) * d_gain, i0 = the lane's first frame of the
                // chunk: the second term's factor is a compile-time constant, so no per-frame `i as f32` is needed
                const float cj = (float)frame_of(j0 + u);
                g[u] = frame_of(j0 + u) == 0 ? g0 : fma2(pk2(cj, cj), dgp, g0);
            }
        }
        if (STRICT) {
#pragma unroll
            for (int u = 0; u < ILP; u++) g[u] = add2(pgp, g[u]);  // prev_state.gain + i as f32 * d_gain
        }
#pragma unroll
        for (int u = 0; u < ILP; u++) b[u] = sub2(b[u], a[u]);     // frame::lerp = a + t * (b - a) (frame.rs:39-41)
        if (STRICT) {
#pragma unroll
            for (int u = 0; u < ILP; u++) s[u] = mulx(fr[u], b[u], nz);
#pragma unroll
            for (int u = 0; u < ILP; u++) s[u] = add2(a[u], s[u]);
        } else {
#pragma unroll
            for (int u = 0; u < ILP; u++) s[u] = fma2(fr[u], b[u], a[u]);
        }
        if (!FULL) {
#pragma unroll
            for (int u = 0; u < ILP; u++)
                if (c * ODB_SPATIAL_CHUNK + frame_of(j0 + u) + LSTEP * lane >= nfr) s[u] = 0ull;  // beyond the tile: +0
        }
#pragma unroll
        for (int u = 0; u < ILP; u++) {
            if (STRICT) acc[cc * 8 + j0 + u] = add2(acc[cc * 8 + j0 + u], mulx(s[u], g[u], nz));  // o[ear] += s * gain (spatial.rs:460)
            else acc[cc * 8 + j0 + u] = fma2(s[u], g[u], acc[cc * 8 + j0 + u]);
        }
    }
}

template <class CFG, bool STRICT, bool FULL, bool UL, bool UR>
__device__ __forceinline__ void consume_source(u64* __restrict__ acc, const int lane, const float lanef, const int part,
                                               const uint32_t rec_sa, const uint32_t rec_x, const uint32_t pcm_b,
                                               const uint32_t rows_sa, const uint32_t* __restrict__ K, const int nfr, const u64 d1,
                                               const u64 d2, const u64 d3, const u64 pgp, const u64 dgp, const u64 nz) {
    const uint32_t lane_b = (uint32_t)(lane * (CFG::LOOP == 0 ? 4 : 8));
#pragma unroll
    for (int cc = 0; cc < CFG::HCHUNKS; cc++) {
        const int c = part * CFG::HCHUNKS + cc;
        if (!FULL && c * ODB_SPATIAL_CHUNK >= nfr) break;
        const uint32_t KL = pcm_b + K[2 * cc] + (UL ? lane_b : 0u);
        const uint32_t KR = pcm_b + K[2 * cc + 1] + (UR ? lane_b : 0u);
        u64 fr_unit = 0ull;
        if (UL || UR)
            fr_unit = pk2(__uint_as_float(lds_u32(rec_sa + ((uint32_t)((ODB_JW_OFF0 + c) * 4) ^ rec_x))),
                          __uint_as_float(lds_u32(rec_sa + ((uint32_t)((ODB_JW_OFF0 + ODB_TILE_CHUNKS + c) * 4) ^ rec_x))));
        consume_chunk<CFG, STRICT, FULL, UL, UR>(acc, cc, c, lane, lanef + (float)(c * ODB_SPATIAL_CHUNK),
                                                 rows_sa + (uint32_t)(cc * CFG::ROW_BYTES), KL, KR, d1, d2, d3, fr_unit, pgp, dgp,
                                                 nfr, nz);
    }
}

// The literal path for one flagged job and one part of the tile (PART_FRAMES frames from `first`), added into the
// warp's parked accumulators `tile` (float2 per frame of the part). The arithmetic is k_mix_general's
// (odb_spatial.cu): every operation a single unfused IEEE operation in the reference's order.
// `scratch`: 2 x part_frames floats (cursors, or Cycle's lerped samples).
__device__ __noinline__ void general_part(const OdbJob* __restrict__ job, const int first, const int part_frames, const int tl,
                                          const int lane, float2* __restrict__ tile, float* __restrict__ scratch) {
    const uint32_t jf = job->flags;
    const int nfr = job->n_frames;
    if (nfr <= first) return;
    const bool cycle = (jf & ODB_JF_CYCLE) != 0;
    const int part_chunks = part_frames / ODB_SPATIAL_CHUNK;
    if (lane < 2 * part_chunks) {
        const int e = lane & 1, cc = lane >> 1, c = first / ODB_SPATIAL_CHUNK + cc;
        const int m = min(ODB_SPATIAL_CHUNK, nfr - c * ODB_SPATIAL_CHUNK);
        float* dst = scratch + e * part_frames + cc * ODB_SPATIAL_CHUNK;
        if (m > 0 && cycle) {  // Cycle::sample, cycle.rs:26-53: the chain lane also takes the (wrapping) taps
            const float* __restrict__ x0 = job->pcm;
            const unsigned long long ulen = (unsigned long long)job->len;
            const float ds = job->ds[e];
            unsigned long long cbase = (unsigned long long)job->base[e][c];
            float offset = job->off0[e][c];
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
        } else if (m > 0 && !(jf & (e == 0 ? ODB_JF_FAST_L : ODB_JF_FAST_R))) {
            const float ds = job->ds[e];
            float offset = job->off0[e][c];
            for (int k = 0; k < ODB_SPATIAL_CHUNK; k++) {
                dst[k] = offset;
                offset = offset + ds;                                                  // frames.rs:195
            }
        }
    }
    __syncwarp();
    const float* __restrict__ pcm = job->pcm;
    const int len = job->len;
    const float fg = job->fixed_gain;
    const bool has_fg = (jf & ODB_JF_FIXED_GAIN) != 0;
    for (int e = 0; e < 2; e++) {
        const bool unit = (jf & (e == 0 ? ODB_JF_FAST_L : ODB_JF_FAST_R)) != 0;
        const float pg = job->pg[e], dg = job->dg[e];
        for (int j = 0; j < part_frames / 32; j++) {
            const int f = 32 * j + lane, i = first + f;  // frame of the part / of the tile
            if (i >= nfr) continue;
            const int c = i >> 8, k = i & (ODB_SPATIAL_CHUNK - 1);
            float smp;
            if (cycle) {
                smp = scratch[e * part_frames + f];
            } else {
                const long long base = job->base[e][c];
                float a, b, fract;
                if (unit) {                                                            // frames.rs:183-187
                    get_pair_mono(pcm, len, base + k, a, b);
                    fract = job->off0[e][c];
                } else {                                                               // frames.rs:191-193
                    const float offset = scratch[e * part_frames + f];
                    const long long tr = (long long)offset;
                    get_pair_mono(pcm, len, base + tr, a, b);
                    fract = offset - (float)tr;
                }
                smp = a + fract * (b - a);                                             // frame.rs:39-41
            }
            if (has_fg) smp = smp * fg;                                                // gain.rs:35
            const float gain = pg + (float)(tl * ODB_TILE_FRAMES + i) * dg;            // spatial.rs:459
            const float contrib = smp * gain;                                          // spatial.rs:460
            float* o = reinterpret_cast<float*>(tile + f) + e;
            *o = *o + contrib;
        }
    }
    __syncwarp();
}

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// Epilogue (tanh.rs:24-28 / reinhard.rs:30-34 act on the finished sum) and store of one output float of tile `tl`.
__device__ __forceinline__ void store_output(const OdbSceneMixArgs& A, const int tl, const int f, float total) {
    const int frame = tl * ODB_TILE_FRAMES + f / 2;
    if (frame >= A.n_frames) return;
    if ((A.epilogue & 0xFF) == 1) total = tanhf(total);
    else if ((A.epilogue & 0xFF) == 2) total = total / (1.0f + fabsf(total));
    const size_t oi = (size_t)tl * (2 * ODB_TILE_FRAMES) + f;
    if (A.epilogue & ODB_EPILOGUE_I16_BIT) {  // examples/offline.rs:39 `(sample * i16::MAX as f32) as i16`
        int v = __float2int_rz(total * 32767.0f);
        v = max(-32768, min(32767, v));
        reinterpret_cast<short*>(A.out)[oi] = (short)v;
    } else {
        A.out[oi] = total;
    }
}

// ---- warp-specialised mix phase (SmxWsCfg) ---------------------------------------------------------------------------
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
template <int R>
__device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(R)); }
template <int R>
__device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(R)); }

// Ring positions that carry over from tile to tile (the barriers' phases do). A consumer warp: the slot of its next
// source and that slot's `full` parity, PCM windows issued / consumed. A producer warp: per team the slot of the next
// source and how often the ring has wrapped (`empty` parity), staging buffers issued / consumed.
// (One set of registers for both roles: a warp is a consumer or a producer for the whole kernel.)
struct WsState {
    uint32_t a = 0, b = 0, c = 0, d = 0;  // consumer: slot, parity, windows, -; producer: slot A, wraps A, slot B, wraps B
    uint32_t n = 0;                       // producer: passes staged so far
};

template <class CFG>
__device__ __forceinline__ void ws_init_barriers(const uint32_t smem_sa, const int warp, const int lane) {
    if (lane != 0) return;
    if (warp < CFG::CWARPS) {
        const uint32_t tma_sa = smem_sa + (uint32_t)(CFG::BARS_OFF + warp * 16);
        mbar_init(tma_sa, 1);
        mbar_init(tma_sa + 8, 1);
        if ((warp & 1) == 0) {  // the team's slot barriers: full (one arrival: the producer), empty (both consumer warps)
            const uint32_t tb = smem_sa + (uint32_t)(CFG::TEAM_BARS_OFF + (warp >> 1) * 16 * CFG::S);
            for (int s = 0; s < CFG::S; s++) {
                mbar_init(tb + 8 * s, 1);
                mbar_init(tb + 8 * CFG::S + 8 * s, 2);
            }
        }
    } else {
        const uint32_t pb = smem_sa + (uint32_t)(CFG::PROD_BARS_OFF + (warp - CFG::CWARPS) * 16);
        mbar_init(pb, 1);
        mbar_init(pb + 8, 1);
    }
}

// One 1024-frame tile of the callback, warp-specialised. Returns whether one of this (consumer) warp's jobs needs the
// literal path. The team -> source assignment is the shipped kernel's (batches of `bsz` consecutive sources, batch
// gp + r GP in round r), and so is the order in which a warp accumulates them: the output is bit-identical.
template <class CFG, bool STRICT, bool VARBATCH>
__device__ __forceinline__ bool ws_mix_tile(WsState& ws, const uint32_t smem_sa, const int warp, const int lane,
                                            const OdbJob* __restrict__ tile_jobs, const int n_sources, const int bsz, const int G,
                                            const int tl, const u64 nz) {
    constexpr int S = CFG::S, HCHUNKS = CFG::HCHUNKS, NACC = CFG::NACC, TEAMS = CFG::CWARPS / 2;
    const int GP = G * TEAMS;
    bool saw_flagged = false;
    if (warp < CFG::CWARPS) {
        // ================= consumer: never leaves the consume loop =================
        if (CFG::REGC) reg_inc<CFG::REGC ? CFG::REGC : 96>();
        const int part = warp & 1, team = warp >> 1;
        const int gp = blockIdx.x * TEAMS + team;
        const uint32_t win_sa = smem_sa + (uint32_t)(warp * CFG::WARP_BYTES);
        const uint32_t rows_base = win_sa + CFG::ROWS_OFF, recs_base = win_sa + CFG::RECS_OFF;
        const uint32_t tma_sa = smem_sa + (uint32_t)(CFG::BARS_OFF + warp * 16);
        const uint32_t full_sa = smem_sa + (uint32_t)(CFG::TEAM_BARS_OFF + team * 16 * S), empty_sa = full_sa + 8 * S;
        const float lanef = (float)(tl * ODB_TILE_FRAMES + lane);
        u64 acc[NACC];
#pragma unroll
        for (int j = 0; j < NACC; j++) acc[j] = 0ull;
        uint32_t slot = ws.a, par = ws.b, wi = ws.c, wc = ws.c;  // (every window issued in a tile is consumed in it)
        auto issue = [&](const uint4& d) {  // the PCM window of a source into window buffer wi & 1 (one elected lane)
            const u64 p = ((u64)d.y << 32) | (u64)d.x;
            const uint32_t b = wi & 1u;
            asm volatile(
                "{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\t"
                "@p mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t"
                "@p cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%2], [%3], %1, [%0];\n\t}" ::"r"(tma_sa + b * 8),
                "r"(d.z), "r"(win_sa + b * CFG::PCM_BYTES), "l"(p)
                : "memory");
            wi++;
        };
        int r = 0, q = 0;  // position in the team's sequence: source q of its batch of round r
        if ((gp + r * GP) * bsz + q < n_sources) {
            mbar_wait(full_sa + 8 * slot, par);
            uint4 hdr = lds_u128(recs_base + slot * CFG::REC_BYTES);  // window address (2 words), bytes, code
            if (!(hdr.w & SMX_WS_SKIP)) issue(hdr);
            for (;;) {
                int q1 = q + 1, r1 = r;
                if (q1 == bsz) { q1 = 0; r1++; }
                const bool more = (gp + r1 * GP) * bsz + q1 < n_sources;
                uint32_t nslot = slot + 1u, npar = par;
                if (nslot == (uint32_t)S) { nslot = 0u; npar ^= 1u; }
                uint4 hdr1 = make_uint4(0u, 0u, 0u, SMX_WS_SKIP);
                if (more) {  // the next source's window streams in while this one is consumed
                    mbar_wait(full_sa + 8 * nslot, npar);
                    hdr1 = lds_u128(recs_base + nslot * CFG::REC_BYTES);
                    if (!(hdr1.w & SMX_WS_SKIP)) issue(hdr1);
                }
                if (!(hdr.w & SMX_WS_SKIP)) {
                    const uint32_t rec_sa = recs_base + slot * CFG::REC_BYTES;
                    const uint4 P = lds_u128(rec_sa + ODB_JW_DS * 4);   // ds (L, R), prev gain (L, R)
                    const uint4 B = lds_u128(rec_sa + ODB_JW_DG * 4);   // d_gain (L, R), -, n_frames
                    uint32_t K[2 * HCHUNKS];
#pragma unroll
                    for (int i = 0; i < 2 * HCHUNKS; i += 4) {
                        const uint4 k4 = lds_u128(rec_sa + (SJ_K + i) * 4);
                        K[i] = k4.x; K[i + 1] = k4.y; K[i + 2] = k4.z; K[i + 3] = k4.w;
                    }
                    const u64 dsp = ((u64)P.y << 32) | P.x, pgp = ((u64)P.w << 32) | P.z, dgp = ((u64)B.y << 32) | B.x;
                    const int nfr = (int)B.w;
                    const int rr = lane & 3;
                    const u64 d1 = rr >= 1 ? dsp : 0ull, d2 = rr >= 2 ? dsp : 0ull, d3 = rr >= 3 ? dsp : 0ull;
                    const uint32_t rows_sa = rows_base + slot * CFG::SLOT_ROWS + (uint32_t)((lane >> 2) * 8);
                    const uint32_t b = wc & 1u;
                    const uint32_t pcm_b = win_sa + b * CFG::PCM_BYTES;
                    mbar_wait(tma_sa + b * 8, (wc >> 1) & 1u);
                    wc++;
                    const uint32_t code = hdr.w & 7u;
#define ODB_CONSUME(F, L, R) consume_source<CFG, STRICT, F, L, R>(acc, lane, lanef, part, rec_sa, 0u, pcm_b, rows_sa, K, nfr, d1, d2, d3, pgp, dgp, nz)
                    if (code == 4u) ODB_CONSUME(true, false, false);  // the common case: full tile, both ears on the doppler path
                    else if (code == 7u) ODB_CONSUME(true, true, true);
                    else if (code == 0u) ODB_CONSUME(false, false, false);
                    else if (code == 3u) ODB_CONSUME(false, true, true);
                    else if (code == 6u) ODB_CONSUME(true, true, false);
                    else if (code == 5u) ODB_CONSUME(true, false, true);
                    else if (code == 2u) ODB_CONSUME(false, true, false);
                    else ODB_CONSUME(false, false, true);
#undef ODB_CONSUME
                }
                saw_flagged = saw_flagged || (hdr.w & SMX_WS_FLAGGED) != 0u;
                __syncwarp();  // every lane is done with the slot (and the window) before it is handed back
                if (lane == 0) mbar_arrive(empty_sa + 8 * slot);
                slot = nslot; par = npar;
                if (!more) break;
                hdr = hdr1; q = q1; r = r1;
            }
        }
        ws.a = slot; ws.b = par; ws.c = wc;
        // park the accumulators: float2 per frame of this part at the start of the warp's region (nothing of the ring is
        // in use any more: the producer has delivered, and this warp consumed, all of the tile's sources)
#pragma unroll
        for (int j = 0; j < NACC; j++)
            asm volatile("st.shared.b64 [%0], %1;" ::"r"(win_sa + (uint32_t)((32 * j + lane) * 8)), "l"(acc[j]) : "memory");
        if (CFG::REGC) reg_dec<96>();
    } else {
        // ================= producer: records, windows' addresses, cursor chains for two teams =================
        if (CFG::REGC) reg_dec<CFG::REGP ? CFG::REGP : 96>();
        constexpr int E = CFG::PASS;                            // entries (sources) per team and pass: 2 teams x E = 8 sources
        const int pw = warp - CFG::CWARPS;
        const int gA = blockIdx.x * TEAMS + 2 * pw;             // team A; team B = gA + 1
        const uint32_t stage_sa = smem_sa + (uint32_t)(CFG::STAGE_OFF + pw * CFG::STAGE_BYTES);
        const uint32_t pbar_sa = smem_sa + (uint32_t)(CFG::PROD_BARS_OFF + pw * 16);
        // this lane's pair of chains (both ears packed): pass source k (team k / E, entry k % E), chunk c of the tile
        const int k = lane >> 2, c = lane & 3;
        const int kt = k / E, kj = k % E;
        const uint32_t my_bars = smem_sa + (uint32_t)(CFG::TEAM_BARS_OFF + (2 * pw + kt) * 16 * S);
        uint32_t slotA = ws.a, useA = ws.b, slotB = ws.c, useB = ws.d, staged = ws.n, used = ws.n;  // (every pass staged in a tile is used in it)
        auto region = [&](int t, int half) {  // region of the consumer warp that mixes `half` for team t of the pair
            return smem_sa + (uint32_t)((2 * (2 * pw + t) + half) * CFG::WARP_BYTES);
        };
        auto src_of = [&](int t, int r, int q, int j) {  // source of entry j after position (r, q) in team t's sequence
            int qq = q + j, rr = r;
            while (qq >= bsz) { qq -= bsz; rr++; }
            return (gA + t + rr * GP) * bsz + qq;
        };
        // raw job records of the pass that starts at (r, q) into staging buffer staged & 1: lane kk < 2 E fetches one
        auto stage = [&](int r, int q) {
            const bool v = lane < 2 * E && src_of(lane / E, r, q, lane % E) < n_sources;
            const uint32_t m = __ballot_sync(0xffffffffu, v);
            if (m == 0u) return;
            const uint32_t b = staged & 1u;
            if (lane == 0) mbar_expect_tx(pbar_sa + b * 8, (uint32_t)(__popc(m) * CFG::REC_BYTES));
            __syncwarp();
            if (v) bulk_g2s(stage_sa + (uint32_t)((b * 2 * E + lane) * CFG::REC_BYTES), tile_jobs + src_of(lane / E, r, q, lane % E), CFG::REC_BYTES, pbar_sa + b * 8);
            staged++;
        };
        int r = 0, q = 0;
        stage(r, q);
        for (;;) {
            // valid entries of this pass per team (a team's sources end at n_sources; team B's lie behind team A's)
            int nA = 0, nB = 0;
#pragma unroll
            for (int j = 0; j < E; j++) {
                nA += src_of(0, r, q, j) < n_sources ? 1 : 0;
                nB += src_of(1, r, q, j) < n_sources ? 1 : 0;
            }
            if (nA == 0) break;
            int r2 = r, q2 = q + E;
            while (q2 >= bsz) { q2 -= bsz; r2++; }
            // this lane's slot: entry kj of team kt
            const bool my_valid = kj < (kt ? nB : nA);
            uint32_t my_slot = (kt ? slotB : slotA) + (uint32_t)kj, my_use = kt ? useB : useA;
            if (my_slot >= (uint32_t)S) { my_slot -= (uint32_t)S; my_use++; }
            // the pass's raw records have arrived
            const uint32_t b = used & 1u;
            mbar_wait(pbar_sa + b * 8, (used >> 1) & 1u);
            used++;
            // both consumers of the team have handed the slot back (its previous tenant was S sources ago)
            if (my_valid && my_use > 0u) mbar_wait(my_bars + 8 * S + 8 * my_slot, (my_use - 1u) & 1u);
            __syncwarp();
            // record -> both halves' slots (lane l moves word l)
#pragma unroll
            for (int kk = 0; kk < 2 * E; kk++) {
                const int t = kk / E, j = kk % E;
                if (j < (t ? nB : nA)) {
                    uint32_t sl = (t ? slotB : slotA) + (uint32_t)j;
                    if (sl >= (uint32_t)S) sl -= (uint32_t)S;
                    const uint32_t w = lds_u32(stage_sa + (uint32_t)((b * 2 * E + kk) * CFG::REC_BYTES + lane * 4));
                    sts_u32(region(t, 0) + CFG::RECS_OFF + sl * CFG::REC_BYTES + lane * 4, w);
                    sts_u32(region(t, 1) + CFG::RECS_OFF + sl * CFG::REC_BYTES + lane * 4, w);
                }
            }
            __syncwarp();
            stage(r2, q2);  // the next pass's records are requested at once (its staging buffer was read a pass ago)
            // lane (kk, half) < 4 E derives what that half's consumer needs: window address and bytes, dispatch code, per
            // chunk the shared-memory offset of PCM index `base` minus the magic bits (k_mix_fast's prologue)
            {
                const int kk = lane >> 1, half = lane & 1, t = kk / E, j = kk % E;
                if (lane < 4 * E && j < (t ? nB : nA)) {
                    uint32_t sl = (t ? slotB : slotA) + (uint32_t)j;
                    if (sl >= (uint32_t)S) sl -= (uint32_t)S;
                    const uint32_t rec_sa = region(t, half) + CFG::RECS_OFF + sl * CFG::REC_BYTES;
                    const uint4 h = lds_u128(rec_sa);                                           // pcm lo/hi, len, flags
                    const int nfr = (int)lds_u32(rec_sa + ODB_JW_N_FRAMES * 4);
                    const bool mine = !(h.w & (ODB_JF_SKIP | ODB_JF_GENERAL)) && nfr > half * CFG::PART_FRAMES;
                    const bool flagged = (h.w & ODB_JF_GENERAL) && !(h.w & (ODB_JF_SKIP | ODB_JF_RING));
                    if (mine) {
                        const uint2 win = lds_u64x(rec_sa + (ODB_JW_WINDOW + 2 * half) * 4);
                        const int w_start = (int)win.x, w_len = (int)win.y;
                        const u64 p = (((u64)h.y << 32) | (u64)h.x) + (u64)((long long)w_start * 4);
                        const uint32_t mL = (h.w & ODB_JF_FAST_L) ? 0u : (ODB_MAGIC_BITS << 2);
                        const uint32_t mR = (h.w & ODB_JF_FAST_R) ? 0u : (ODB_MAGIC_BITS << 2);
                        int bL[HCHUNKS], bR[HCHUNKS];
#pragma unroll
                        for (int cc = 0; cc < HCHUNKS; cc++) {
                            bL[cc] = (int)lds_u32(rec_sa + (ODB_JW_BASE + half * HCHUNKS + cc) * 4);
                            bR[cc] = (int)lds_u32(rec_sa + (ODB_JW_BASE + ODB_TILE_CHUNKS + half * HCHUNKS + cc) * 4);
                        }
                        const uint32_t code = (nfr == ODB_TILE_FRAMES ? 4u : 0u) | ((h.w & ODB_JF_FAST_L) ? 2u : 0u) | ((h.w & ODB_JF_FAST_R) ? 1u : 0u);
                        sts_u32(rec_sa, (uint32_t)p);
                        sts_u32(rec_sa + 4, (uint32_t)(p >> 32));
                        sts_u32(rec_sa + 8, (uint32_t)w_len * 4u);
                        sts_u32(rec_sa + 12, code);
#pragma unroll
                        for (int cc = 0; cc < HCHUNKS; cc++) {  // SJ_K + 2 cc (+ 1): overwrites `base` words, read above
                            sts_u32(rec_sa + (SJ_K + 2 * cc) * 4, (uint32_t)((bL[cc] - w_start) * 4) - mL);
                            sts_u32(rec_sa + (SJ_K + 2 * cc + 1) * 4, (uint32_t)((bR[cc] - w_start) * 4) - mR);
                        }
                    } else {
                        sts_u32(rec_sa + 12, SMX_WS_SKIP | (flagged ? SMX_WS_FLAGGED : 0u));
                    }
                }
            }
            __syncwarp();
            // the literal cursor chains `offset += ds` (frames.rs:195) of both ears in one packed add per step, every 4th
            // value into the consumer's rows (an ear on the ds ~= 1 path has a row nobody reads)
            if (my_valid) {
                const int half = c / HCHUNKS, cc = c % HCHUNKS;
                const uint32_t reg_sa = region(kt, half);
                const uint32_t rec_sa = reg_sa + CFG::RECS_OFF + my_slot * CFG::REC_BYTES;
                if (!(lds_u32(rec_sa + 12) & SMX_WS_SKIP)) {
                    u64 o = pk2(__uint_as_float(lds_u32(rec_sa + (ODB_JW_OFF0 + c) * 4)),
                                __uint_as_float(lds_u32(rec_sa + (ODB_JW_OFF0 + ODB_TILE_CHUNKS + c) * 4)));
                    const u64 ds = lds_u64(rec_sa + ODB_JW_DS * 4);
                    const uint32_t dst = reg_sa + CFG::ROWS_OFF + my_slot * CFG::SLOT_ROWS + (uint32_t)(cc * CFG::ROW_BYTES);
#pragma unroll 8
                    for (int m = 0; m < CFG::POINTS; m++) {
                        asm volatile("st.shared.b64 [%0], %1;" ::"r"(dst + (uint32_t)(m * 8)), "l"(o) : "memory");  // checkpoint m = cursors of frame 4 m
                        o = add2(o, ds); o = add2(o, ds); o = add2(o, ds); o = add2(o, ds);
                    }
                }
            }
            __syncwarp();  // the release below is cumulative over the warp's stores through this barrier
            if (c == 0 && my_valid) mbar_arrive(my_bars + 8 * my_slot);
            // next pass
            slotA += (uint32_t)nA;
            if (slotA >= (uint32_t)S) { slotA -= (uint32_t)S; useA++; }
            slotB += (uint32_t)nB;
            if (slotB >= (uint32_t)S) { slotB -= (uint32_t)S; useB++; }
            r = r2; q = q2;
        }
        ws.a = slotA; ws.b = useA; ws.c = slotB; ws.d = useB; ws.n = used;
        if (CFG::REGC) reg_inc<96>();
    }
    return saw_flagged;
}

// VARBATCH: the batch size is a run-time argument (small scenes, sharded scenes); otherwise it is the compile-time
// CFG::BATCH - measurably faster on the full-size scene (90.2 vs 93.5 us on C3), where the chooser picks 8 anyway.
template <class CFG, bool STRICT, bool VARBATCH>
__global__ void __launch_bounds__(CFG::WARPS * 32, 1) k_scene_mix(const OdbSceneMixArgs A) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int WARPS = CFG::WARPS, CWARPS = CFG::CWARPS, SPLIT = CFG::SPLIT, HCHUNKS = CFG::HCHUNKS, BATCH = CFG::BATCH, NACC = CFG::NACC;
    // RGROUPS: groups of 16 threads that share the slice reduce. The warp-specialised shapes use the 512-thread shapes' 32
    // (their extra warps sit the reduce out), so that their sums round exactly like the shipped kernel's.
    constexpr int THREADS = WARPS * 32, RGROUPS = CFG::WS && THREADS / SMX_SLICE > 32 ? 32 : THREADS / SMX_SLICE;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int part = warp % SPLIT, team = warp / SPLIT;  // the SPLIT warps of a team mix the parts of the same sources
    const uint32_t smem_sa = smem_u32(smem_raw);
    const uint32_t pcm_sa = smem_sa + (uint32_t)(warp * CFG::WARP_BYTES);  // a mixing warp's region: PCM windows first
    if constexpr (!CFG::WS) {
        const uint32_t bar_sa = smem_sa + (uint32_t)(CFG::BARS_OFF + warp * 16);
        if (lane == 0) {
            mbar_init(bar_sa, 1);
            mbar_init(bar_sa + 8, 1);
        }
    } else {
        ws_init_barriers<CFG>(smem_sa, warp, lane);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    // The job records come from the walk kernel launched just before: this grid starts when every walk block has
    // counted itself into `walked` (its trigger comes after that), and acquires the count - no griddepcontrol.wait,
    // which would also wait for everything the walk was launched behind (the previous callback's exchange tail).
    // Neither this kernel nor the walk waits for the PREVIOUS callback's kernel to finish (its last CTA may still be
    // exchanging tiles with the other GPUs while this grid mixes). What consecutive callbacks share is double-buffered
    // by callback parity - partial tiles, xtile, the arrive / done counters - and the job records live in a ring of
    // three, so the other thing to wait for is the callback before the previous one (almost never an actual wait).
    // Two threads poll the two words side by side (one L2 round trip instead of two).
    if (threadIdx.x == 0)
        while (ld_acquire_u64(A.walked) < A.walked_target) __nanosleep(20);
    if (threadIdx.x == 32 && A.my_seq > 2ull)
        while (ld_acquire_u64(A.completed) < A.my_seq - 2ull) __nanosleep(40);
    __syncthreads();
    // Only now may the next callback's walk start: it finds this callback's walk complete (the wait above) and callback
    // my_seq - 2 finished entirely, whose job records it overwrites.
    pdl_launch_dependents();
    bool saw_flagged = false;  // one of this warp's jobs needs the literal path
    uint32_t parity = 0;  // bit b = phase parity of PCM buffer b's barrier
    uint32_t buf = 0;
    WsState ws;           // warp-specialised shape: ring positions that carry over from tile to tile

    const int n_sources = A.n_sources;
    const u64 nz = A.nz;
    const int G = (int)gridDim.x;
    const int gp = blockIdx.x * (CWARPS / SPLIT) + team, GP = G * (CWARPS / SPLIT);
    // Sources per batch: at most BATCH (the 32 chain lanes), fewer when that evens out the rounds - a warp's work is
    // whole batches, so e.g. 8192 sources are 1024 batches of 8 for 1184 teams (160 teams idle, the others 8 sources
    // each) but 1171 batches of 7 (every team busy, 7 sources each). Chosen by the host (odb_scene_mix_batch).
    const int bsz = VARBATCH ? A.batch : BATCH;
    const int n_batches = (n_sources + bsz - 1) / bsz;
    const int first_frame = part * CFG::PART_FRAMES;

    for (int tl = 0; tl < A.n_tiles; tl++) {
        saw_flagged = false;
        const OdbJob* tile_jobs = A.jobs + (size_t)tl * n_sources;
      if constexpr (CFG::WS) {
        saw_flagged = ws_mix_tile<CFG, STRICT, VARBATCH>(ws, smem_sa, warp, lane, tile_jobs, n_sources, bsz, G, tl, nz);
      } else {
        const uint32_t offs_sa = pcm_sa + 2 * CFG::PCM_BYTES;
        const uint32_t jobs_sa = smem_sa + (uint32_t)(CFG::JOBS_OFF + warp * BATCH * CFG::REC_BYTES);
        const uint32_t bar_sa = smem_sa + (uint32_t)(CFG::BARS_OFF + warp * 16);
        const int c0 = part * HCHUNKS;
        auto rec = [&](int q, int w) { return jobs_sa + (uint32_t)(q * CFG::REC_BYTES) + (uint32_t)((w * 4) ^ (q * 16)); };
        auto start_copy = [&](int q, uint32_t b) {
            const uint4 d = lds_u128(rec(q, SJ_SRC));
            const u64 p = ((u64)d.y << 32) | (u64)d.x;
            asm volatile(
                "{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\t"
                "@p mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t"
                "@p cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%2], [%3], %1, [%0];\n\t}" ::"r"(bar_sa + b * 8),
                "r"(d.z), "r"(pcm_sa + b * CFG::PCM_BYTES), "l"(p)
                : "memory");
        };
        u64 acc[NACC];
#pragma unroll
        for (int j = 0; j < NACC; j++) acc[j] = 0ull;
        const float lanef = (float)(tl * ODB_TILE_FRAMES + (CFG::LOOP == 0 ? lane : 2 * lane));

        for (int bi = gp; bi < n_batches; bi += GP) {
            const int s0 = bi * bsz;
            // 1. stage the batch's job records (one 128-byte line each: lane l moves word l)
#pragma unroll
            for (int q = 0; q < BATCH; q++) {
                uint32_t w = lane == ODB_JW_FLAGS ? ODB_JF_SKIP : 0u;
                if (q < bsz && s0 + q < n_sources) w = __ldcg(reinterpret_cast<const uint32_t*>(tile_jobs + s0 + q) + lane);
                sts_u32(rec(q, lane), w);
            }
            __syncwarp();
            // lane q < BATCH derives what this part needs of source q and rewrites its private record (SJ_* words)
            bool mine = false, flagged = false;
            if (lane < BATCH) {
                const uint4 h = lds_u128(rec(lane, 0));                                     // pcm lo/hi, len, flags
                const int nfr = (int)lds_u32(rec(lane, ODB_JW_N_FRAMES));
                mine = !(h.w & (ODB_JF_SKIP | ODB_JF_GENERAL)) && nfr > first_frame;
                flagged = (h.w & ODB_JF_GENERAL) && !(h.w & (ODB_JF_SKIP | ODB_JF_RING));
                if (mine) {
                    int w_start, w_len;
                    if (SPLIT == 2) {
                        const uint2 win = lds_u64x(rec(lane, ODB_JW_WINDOW + 2 * part));
                        w_start = (int)win.x; w_len = (int)win.y;
                    } else {  // the whole tile: both halves' windows abut or overlap (the cursor only moves forward)
                        const uint4 win = lds_u128(rec(lane, ODB_JW_WINDOW));
                        w_start = (int)win.x;
                        w_len = win.w ? (int)win.z + (int)win.w - w_start : (int)win.y;
                    }
                    const u64 p = (((u64)h.y << 32) | (u64)h.x) + (u64)((long long)w_start * 4);
                    const uint32_t mL = (h.w & ODB_JF_FAST_L) ? 0u : (ODB_MAGIC_BITS << 2);
                    const uint32_t mR = (h.w & ODB_JF_FAST_R) ? 0u : (ODB_MAGIC_BITS << 2);
                    int bL[HCHUNKS], bR[HCHUNKS];
#pragma unroll
                    for (int cc = 0; cc < HCHUNKS; cc++) {
                        bL[cc] = (int)lds_u32(rec(lane, ODB_JW_BASE + c0 + cc));
                        bR[cc] = (int)lds_u32(rec(lane, ODB_JW_BASE + ODB_TILE_CHUNKS + c0 + cc));
                    }
                    const uint32_t code = (nfr == ODB_TILE_FRAMES ? 4u : 0u) | ((h.w & ODB_JF_FAST_L) ? 2u : 0u) | ((h.w & ODB_JF_FAST_R) ? 1u : 0u);
                    sts_u32(rec(lane, SJ_SRC), (uint32_t)p);
                    sts_u32(rec(lane, SJ_SRC + 1), (uint32_t)(p >> 32));
                    sts_u32(rec(lane, SJ_BYTES), (uint32_t)w_len * 4u);
                    sts_u32(rec(lane, SJ_CODE), code);
#pragma unroll
                    for (int cc = 0; cc < HCHUNKS; cc++) {  // SJ_K + 2 cc (+ 1): overwrites the `base` words (12..19), read above
                        sts_u32(rec(lane, SJ_K + 2 * cc), (uint32_t)((bL[cc] - w_start) * 4) - mL);
                        sts_u32(rec(lane, SJ_K + 2 * cc + 1), (uint32_t)((bR[cc] - w_start) * 4) - mR);
                    }
                }
            }
            __syncwarp();
            uint32_t act = __ballot_sync(0xffffffffu, mine);
            saw_flagged = saw_flagged || __any_sync(0xffffffffu, flagged);
            if (act) {
                start_copy(__ffs(act) - 1, buf);
                {   // 2. literal cursor chains: lane = (source q, ear e, chunk cc of this part)
                    const int q = lane / (2 * HCHUNKS), e = (lane / HCHUNKS) & 1, cc = lane % HCHUNKS;
                    const uint32_t jf = lds_u32(rec(q, ODB_JW_FLAGS));
                    if (((act >> q) & 1u) && !(jf & (e ? ODB_JF_FAST_R : ODB_JF_FAST_L))) {
                        float o = __uint_as_float(lds_u32(rec(q, ODB_JW_OFF0 + ODB_TILE_CHUNKS * e + c0 + cc)));
                        const float ds = __uint_as_float(lds_u32(rec(q, ODB_JW_DS + e)));
                        const uint32_t dst = offs_sa + (uint32_t)((q * HCHUNKS + cc) * CFG::ROW_BYTES + e * 4);
#pragma unroll 8
                        for (int m = 0; m < CFG::POINTS; m++) {
                            sts_f32(dst + (uint32_t)(m * 8), o);  // checkpoint m = cursor of frame 4 m
                            o = __fadd_rn(o, ds); o = __fadd_rn(o, ds); o = __fadd_rn(o, ds); o = __fadd_rn(o, ds);
                        }
                    }
                }
                __syncwarp();
                // 3. consume the batch source by source
                while (act) {
                    const int q = __ffs(act) - 1;
                    act &= act - 1u;
                    if (act) start_copy(__ffs(act) - 1, buf ^ 1u);
                    const uint4 P = lds_u128(rec(q, ODB_JW_DS));   // ds (L, R), prev gain (L, R)
                    const uint4 B = lds_u128(rec(q, ODB_JW_DG));   // d_gain (L, R), code, n_frames
                    uint32_t K[2 * HCHUNKS];
#pragma unroll
                    for (int i = 0; i < 2 * HCHUNKS; i += 4) {
                        const uint4 k4 = lds_u128(rec(q, SJ_K + i));
                        K[i] = k4.x; K[i + 1] = k4.y; K[i + 2] = k4.z; K[i + 3] = k4.w;
                    }
                    const u64 dsp = ((u64)P.y << 32) | P.x, pgp = ((u64)P.w << 32) | P.z, dgp = ((u64)B.y << 32) | B.x;
                    const int nfr = (int)B.w;
                    u64 d1, d2, d3;
                    uint32_t rows_sa = offs_sa + (uint32_t)(q * HCHUNKS * CFG::ROW_BYTES);
                    if (CFG::LOOP == 0) {
                        const int r = lane & 3;
                        d1 = r >= 1 ? dsp : 0ull; d2 = r >= 2 ? dsp : 0ull; d3 = r >= 3 ? dsp : 0ull;
                        rows_sa += (uint32_t)((lane >> 2) * 8);
                    } else {
                        d1 = (lane & 1) ? dsp : 0ull; d2 = d1; d3 = dsp;
                        rows_sa += (uint32_t)((lane >> 1) * 8);
                    }
                    const uint32_t pcm_b = pcm_sa + buf * CFG::PCM_BYTES;
                    const uint32_t off0_sa = jobs_sa + (uint32_t)(q * CFG::REC_BYTES), off0_x = (uint32_t)(q * 16);
                    mbar_wait(bar_sa + buf * 8, (parity >> buf) & 1u);
                    parity ^= 1u << buf;
#define ODB_CONSUME(F, L, R) consume_source<CFG, STRICT, F, L, R>(acc, lane, lanef, part, off0_sa, off0_x, pcm_b, rows_sa, K, nfr, d1, d2, d3, pgp, dgp, nz)
                    // The common case gets a direct branch: the comparison is hidden from the compiler, which would otherwise
                    // fold it into the jump table of the switch below (a constant-bank load and an indirect branch in front
                    // of every source's consume loop).
                    uint32_t common;
                    asm("set.eq.u32.u32 %0, %1, 4;" : "=r"(common) : "r"(B.z));
                    if (common) ODB_CONSUME(true, false, false);  // full tile, both ears on the doppler path
                    else switch (B.z) {
                        case 7: ODB_CONSUME(true, true, true); break;
                        case 0: ODB_CONSUME(false, false, false); break;
                        case 3: ODB_CONSUME(false, true, true); break;
                        case 6: ODB_CONSUME(true, true, false); break;
                        case 5: ODB_CONSUME(true, false, true); break;
                        case 2: ODB_CONSUME(false, true, false); break;
                        default: ODB_CONSUME(false, false, true); break;
                    }
#undef ODB_CONSUME
                    __syncwarp();  // every lane is done with this PCM buffer before it is refilled
                    buf ^= 1u;
                }
            }
            __syncwarp();  // ... and with the staged job records and cursor rows
        }

        // park the accumulators: float2 per frame of this part at the start of the warp's region
        if (CFG::LOOP == 0) {
#pragma unroll
            for (int j = 0; j < NACC; j++)
                asm volatile("st.shared.b64 [%0], %1;" ::"r"(pcm_sa + (uint32_t)((32 * j + lane) * 8)), "l"(acc[j]) : "memory");
        } else {
#pragma unroll
            for (int j = 0; j < NACC / 2; j++) {
                const int f = (j >> 2) * ODB_SPATIAL_CHUNK + 64 * (j & 3) + 2 * lane;  // accumulators 2 j, 2 j + 1 = frames f, f + 1
                asm volatile("st.shared.v2.b64 [%0], {%1, %2};" ::"r"(pcm_sa + (uint32_t)(f * 8)), "l"(acc[2 * j]), "l"(acc[2 * j + 1]) : "memory");
            }
        }
      }  // !CFG::WS
        __syncwarp();
        // 3'. the flagged jobs of this warp's batches, literally (rare: none on C3); the batch prologues noticed them
        if (saw_flagged) {
            float2* tile = reinterpret_cast<float2*>(smem_raw + warp * CFG::WARP_BYTES);
            float* scratch = reinterpret_cast<float*>(smem_raw + warp * CFG::WARP_BYTES + CFG::PART_FRAMES * 8);
            for (int bi = gp; bi < n_batches; bi += GP) {
                const int s0 = bi * bsz;
                uint32_t f = ODB_JF_SKIP;
                if (lane < bsz && s0 + lane < n_sources) f = __ldcg(&tile_jobs[s0 + lane].flags);
                uint32_t m = __ballot_sync(0xffffffffu, (f & ODB_JF_GENERAL) && !(f & (ODB_JF_SKIP | ODB_JF_RING)));
                while (m) {
                    const int q = __ffs(m) - 1;
                    m &= m - 1u;
                    general_part(tile_jobs + s0 + q, first_frame, CFG::PART_FRAMES, tl, lane, tile, scratch);
                }
            }
        }
        __syncthreads();
        // 4. fold warp -> CTA in fixed order: one partial tile per CTA
        float* pdst = A.partials + ((size_t)tl * G + blockIdx.x) * (2 * ODB_TILE_FRAMES);
        constexpr int PART_FLOATS = 2 * CFG::PART_FRAMES;
        for (int f = threadIdx.x; f < 2 * ODB_TILE_FRAMES; f += THREADS) {
            const int h = f / PART_FLOATS, fh = f - h * PART_FLOATS;
            float sum = 0.0f;
#pragma unroll
            for (int w = 0; w < CWARPS / SPLIT; w++)
                sum = sum + *reinterpret_cast<const float*>(smem_raw + (w * SPLIT + h) * CFG::WARP_BYTES + fh * 4);
            __stcg(pdst + f, sum);
        }
        __syncthreads();
        if (threadIdx.x == 0) {  // one device-scope fence per CTA, cumulative over the CTA's stores through the barrier
            __threadfence();
            atomicAdd(A.arrive, 1ull);
        }
        // 5. when every CTA has arrived: sum slices of the partial tiles in index order. A plain device-resident
        //    callback applies the epilogue and stores the output here; one that ends in an exchange or in host memory
        //    leaves the raw sum in `xtile` for the grid's last CTA (below)
        const unsigned long long target = A.arrive_base + (unsigned long long)(tl + 1) * (unsigned long long)G;
        float* red = reinterpret_cast<float*>(smem_raw);  // [RGROUPS][SMX_SLICE]
        bool waited = false;
        for (int sl = blockIdx.x; sl < SMX_SLICES; sl += G) {
            if (!waited) {
                if (threadIdx.x == 0)
                    while (ld_acquire_u64(A.arrive) < target) __nanosleep(40);
                __syncthreads();
                waited = true;
            }
            const int fl = threadIdx.x & (SMX_SLICE - 1), grp = threadIdx.x / SMX_SLICE;
            const int f = sl * SMX_SLICE + fl;
            const float* p = A.partials + (size_t)tl * G * (2 * ODB_TILE_FRAMES) + f;
            if (!CFG::WS || grp < RGROUPS) {
                float sum = 0.0f;
                for (int i = grp; i < G; i += RGROUPS) sum = sum + __ldcg(p + (size_t)i * (2 * ODB_TILE_FRAMES));
                red[grp * SMX_SLICE + fl] = sum;
            }
            __syncthreads();
            if (threadIdx.x < SMX_SLICE) {
                float total = 0.0f;
#pragma unroll
                for (int g = 0; g < RGROUPS; g++) total = total + red[g * SMX_SLICE + fl];
                if (A.xtile) __stcg(A.xtile + (size_t)tl * (2 * ODB_TILE_FRAMES) + f, total);
                else store_output(A, tl, f, total);
            }
            __syncthreads();
        }
        __syncthreads();  // the warp regions are reused by the next tile
    }
    // ---- the grid's last CTA finishes the callback: the exchange over NVLink and / or the hand-over to the host, and
    // the completion sequence number. Every other CTA leaves here, so its SM is free for the next callback's kernels
    // while the last one waits on fences and peers.
    __shared__ int last_done;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();  // one device-scope fence per CTA, cumulative over its stores through the barrier
        const unsigned long long d = atomicAdd(A.done, 1ull) + 1ull;
        last_done = d == A.done_base + (unsigned long long)G;
        if (last_done) __threadfence();
    }
    __syncthreads();
    if (!last_done) return;
    const int n_floats = 2 * A.n_frames;        // what the callback renders ...
    const int n_float4 = (n_floats + 3) / 4;    // ... moved in 16-byte words (inbox slots and xtile are padded to a multiple of 32 floats)
    if (A.push_seq != 0u) {
        // push: this rank's sum (no epilogue) into slot `rank` of every rank's inbox, then the sequence number in the
        // slot's flags - the stand-alone kernels' protocol (odb_exchange.cu). The slots about to be overwritten held
        // exchange push_seq - depth: every peer must have pulled that one (practically never a wait).
        const uint32_t par = A.push_seq % (uint32_t)A.xg.depth;
        if (threadIdx.x < A.xg.world && A.push_seq > (uint32_t)A.xg.depth) {
            const uint32_t* ack = reinterpret_cast<const uint32_t*>(A.peers.inbox[A.xg.rank] + A.xg.acks_off) + (size_t)threadIdx.x * A.xg.max_slices;
            for (int tl = 0; tl < A.n_tiles; tl++)
                while ((int)(ld_acquire_sys(ack + tl) - (A.push_seq - (uint32_t)A.xg.depth)) < 0) __nanosleep(20);
        }
        __syncthreads();
        const size_t slot = ((size_t)par * A.xg.world + A.xg.rank) * A.xg.cap;
        for (int p = 0; p < A.xg.world; p++) {
            float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(A.peers.inbox[p]) + slot);
            for (int i = threadIdx.x; i < n_float4; i += blockDim.x) dst[i] = __ldcg(reinterpret_cast<const float4*>(A.xtile) + i);
        }
        __syncthreads();  // the release stores below are cumulative over the CTA's data stores through this barrier
        if (threadIdx.x < A.xg.world) {
            uint32_t* flag = reinterpret_cast<uint32_t*>(A.peers.inbox[threadIdx.x] + A.xg.flags_off) +
                             ((size_t)par * A.xg.world + A.xg.rank) * A.xg.max_slices;
            for (int tl = 0; tl < A.n_tiles; tl++) st_release_sys(flag + tl, A.push_seq);
        }
    }
    if (A.pull_seq != 0u) {
        // pull: the sum of exchange pull_seq over the ranks, in rank order (bit-identical on every rank), epilogue, store.
        // The push above is out before this waits, so no rank can wait for a flag that depends on its own progress.
        const uint32_t par = A.pull_seq % (uint32_t)A.xg.depth;
        const char* mine = A.peers.inbox[A.xg.rank];
        if (threadIdx.x < A.xg.world) {
            const uint32_t* flag = reinterpret_cast<const uint32_t*>(mine + A.xg.flags_off) + ((size_t)par * A.xg.world + threadIdx.x) * A.xg.max_slices;
            for (int tl = 0; tl < A.n_tiles; tl++)
                while ((int)(ld_acquire_sys(flag + tl) - A.pull_seq) < 0) __nanosleep(20);
        }
        __syncthreads();
        const float* in = reinterpret_cast<const float*>(mine) + (size_t)par * A.xg.world * A.xg.cap;
        for (int i = threadIdx.x; i < n_floats; i += blockDim.x) {
            float sum = 0.0f;
            for (int p = 0; p < A.xg.world; p++) sum = sum + __ldcv(in + (size_t)p * A.xg.cap + i);
            store_output(A, i / (2 * ODB_TILE_FRAMES), i % (2 * ODB_TILE_FRAMES), sum);
        }
    } else if (A.push_seq == 0u && A.xtile) {
        // one GPU, host tile: epilogue and store into (pinned) host memory, 512 coalesced lanes
        for (int i = threadIdx.x; i < n_floats; i += blockDim.x)
            store_output(A, i / (2 * ODB_TILE_FRAMES), i % (2 * ODB_TILE_FRAMES), __ldcg(A.xtile + i));
    }
    __syncthreads();
    if (A.host_flag) {
        if (threadIdx.x == 0) __threadfence_system();  // cumulative over the CTA's output stores (host memory)
        __syncthreads();
    }
    if (A.pull_seq != 0u && threadIdx.x < A.xg.world) {  // the peers may overwrite this slot `depth` exchanges on
        uint32_t* ack = reinterpret_cast<uint32_t*>(A.peers.inbox[threadIdx.x] + A.xg.acks_off) + (size_t)A.xg.rank * A.xg.max_slices;
        for (int sl = 0; sl < A.xg.max_slices; sl++) st_release_sys(ack + sl, A.pull_seq);
    }
    if (A.host_flag && threadIdx.x == 0) {
        if (A.removed_count_host) {
            *reinterpret_cast<volatile uint32_t*>(A.removed_count_host) = __ldcg(A.removed_count);
            __threadfence_system();
        }
        *reinterpret_cast<volatile unsigned long long*>(A.host_flag) = A.seq;
    }
    __syncthreads();
    if (threadIdx.x == 0) {  // everything of this callback is done: its parity's buffers may be reused
        __threadfence();
        atomicMax(A.completed, A.my_seq);
    }
}

}  // namespace odbk

using namespace odbk;

// Shapes built into the library (experiments: ODB_SMX_CFG = 1 or 3 in the environment selects the alternatives measured
// in DESIGN.md section 7; the 8-frames-in-flight variants measured there are not kept in the build).
typedef SmxCfg<2, 16, 0, 4> SmxDefault;
typedef SmxCfg<2, 16, 1, 4> SmxPairs;
typedef SmxCfg<1, 12, 0, 4> SmxWhole;
typedef SmxWsCfg<16, 7, 112, 32> SmxWs;        // 16 consumer + 4 producer warps, register file re-divided (setmaxnreg)
typedef SmxWsCfg<12, 8, 0, 0> SmxWs12;         // 12 consumer + 3 producer warps at the launch's 136 registers

static int g_smx_cfg = -1;
static int smx_cfg() {
    if (g_smx_cfg < 0) {
        const char* e = getenv("ODB_SMX_CFG");
        g_smx_cfg = e ? atoi(e) : 0;
        if (g_smx_cfg != 1 && g_smx_cfg != 3 && g_smx_cfg != 4 && g_smx_cfg != 5) g_smx_cfg = 0;
    }
    return g_smx_cfg;
}

// Sources per batch and CTAs for `n_sources`. Scenes that give every team at least half a full batch keep the
// compile-time batch of 8 (the faster instantiation; and at 8 GPUs the 128-CTA grid of an 8192-source shard leaves 20
// SMs on which the next callback's walk runs underneath the mix - batches of 7 would fill 147 SMs and expose the walk:
// measured 29.7 us per callback instead of 23.6). Smaller scenes take the batch size (1..BATCH) that minimises a
// team's work, rounds x (batch + the per-batch overhead of prologue and cursor chains, about 0.6 of a source's consume).
template <class CFG>
static void smx_shape(int n_sources, int sm_count, int* batch, int* ctas) {
    const int teams_per_cta = CFG::CWARPS / CFG::SPLIT;
    const long long teams = (long long)teams_per_cta * sm_count;
    int best = CFG::BATCH;
    if ((long long)n_sources < teams * (CFG::BATCH / 2)) {
        double best_cost = 1e300;
        for (int b = CFG::BATCH; b >= 1; b--) {
            const long long batches = ((long long)n_sources + b - 1) / b;
            const long long rounds = (batches + teams - 1) / teams;
            const double cost = (double)rounds * ((double)b + 0.6);
            if (cost < best_cost - 1e-9) { best_cost = cost; best = b; }
        }
    }
    // The grid: the FEWEST CTAs that still finish in the same number of rounds as the whole chip would. A team's work
    // is whole batches, so e.g. the 4096 batches of a 32 768-source shard take four rounds on 148 SMs (3.46 per team)
    // and exactly four on 128 - and the 20 SMs left over are where the next callback's walk kernel runs underneath
    // this grid instead of after it.
    const long long batches = ((long long)n_sources + best - 1) / best;
    const long long rounds = (batches + teams - 1) / teams;
    long long want = (batches + teams_per_cta * rounds - 1) / (teams_per_cta * rounds);
    *batch = best;
    *ctas = (int)(want < 1 ? 1 : (want > sm_count ? sm_count : want));
}
void odb_scene_mix_shape(int n_sources, int sm_count, int* batch, int* ctas) {
    switch (smx_cfg()) {
        case 1: return smx_shape<SmxPairs>(n_sources, sm_count, batch, ctas);
        case 3: return smx_shape<SmxWhole>(n_sources, sm_count, batch, ctas);
        case 4: return smx_shape<SmxWs>(n_sources, sm_count, batch, ctas);
        case 5: return smx_shape<SmxWs12>(n_sources, sm_count, batch, ctas);
        default: return smx_shape<SmxDefault>(n_sources, sm_count, batch, ctas);
    }
}

template <class CFG, bool STRICT, bool VARBATCH>
static cudaError_t launch_scene_mix(const OdbSceneMixArgs& a, int n_ctas, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(k_scene_mix<CFG, STRICT, VARBATCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, CFG::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    return odb_launch_pdl(k_scene_mix<CFG, STRICT, VARBATCH>, dim3(n_ctas), dim3(CFG::WARPS * 32), (size_t)CFG::SMEM_BYTES, st, a);
}
template <class CFG>
static cudaError_t launch_scene_mix_mode(const OdbSceneMixArgs& a, int n_ctas, int mode, cudaStream_t st) {
    if (a.batch == CFG::BATCH)
        return (mode & 1) ? launch_scene_mix<CFG, false, false>(a, n_ctas, st) : launch_scene_mix<CFG, true, false>(a, n_ctas, st);
    return (mode & 1) ? launch_scene_mix<CFG, false, true>(a, n_ctas, st) : launch_scene_mix<CFG, true, true>(a, n_ctas, st);
}

// mode bit 0: value multiply-adds contracted to FMA
cudaError_t odb_launch_scene_mix(const OdbSceneMixArgs& args, int n_ctas, int mode, cudaStream_t st) {
    OdbSceneMixArgs a = args;
    a.nz = 0x8000000080000000ull;
    switch (smx_cfg()) {
        case 1: return launch_scene_mix_mode<SmxPairs>(a, n_ctas, mode, st);
        case 3: return launch_scene_mix_mode<SmxWhole>(a, n_ctas, mode, st);
        case 4: return launch_scene_mix_mode<SmxWs>(a, n_ctas, mode, st);
        case 5: return launch_scene_mix_mode<SmxWs12>(a, n_ctas, mode, st);
        default: return launch_scene_mix_mode<SmxDefault>(a, n_ctas, mode, st);
    }
}
