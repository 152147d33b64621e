// Staged ("fast") spatial mix kernel for the seek path: SpatialScene's mix closure (spatial.rs:445-469)
// over FramesSignal::sample (frames.rs:176-201) for every source whose PCM window of one 1024-frame
// tile fits in shared memory.
//
// A persistent grid of one 16-warp CTA per SM; the two warps of a pair mix the two 512-frame halves of the same
// sources. Every warp works on its own batches of 8 consecutive sources and never synchronises with the other
// warps until the final fold:
//   1. the warp stages the batch's eight 128-byte job records in shared memory (bank-swizzled), lanes 0-7 each
//      derive what this half needs of one source (window address and size, per-chunk tap offsets, dispatch code)
//      and rewrite their record, so that the per-source set-up later is three broadcast 16-byte loads; one elected
//      lane starts a bulk async copy (TMA, cp.async.bulk -> UBLKCP) of the first source's PCM window HBM -> one
//      of the warp's two PCM buffers, completion on a per-buffer mbarrier;
//   2. while the copy is in flight all 32 lanes - one per (source, ear, 256-frame chunk of the half) - walk the
//      reference's serial cursor `offset += ds` (frames.rs:195) literally and store every 4th cursor value
//      to shared memory. The chain is the one part of the path that is not associative: it is evaluated
//      with exactly the reference's sequence of f32 additions, so frame indices are bit-exact
//      (SURVEY.md §7 H1);
//   3. per source, all 32 lanes consume (the next source's window is already being copied into the other
//      buffer): lane l owns frames l, l+32, ... of the tile, re-derives its cursor from the stored
//      checkpoint with <= 3 more literal additions, splits it into index and fraction with a round-down
//      magic add (no F2I/I2F), gathers the sample pair from shared memory, lerps (frame.rs:39-41),
//      applies the per-frame gain ramp (spatial.rs:459) and accumulates into 32 packed register accumulators
//      (16 frames x 2 ears, ears packed as FP32x2: FADD2/FFMA2). An ear on FramesSignal's ds ~= 1 path
//      (frames.rs:180-187) skips the cursor and uses index base + i with a constant fraction;
//   4. after its last batch a warp parks the accumulators in shared memory, the CTA folds its warps in a
//      fixed order and writes one partial tile; k_reduce_tiles sums the partial tiles.
//
// STRICT = true keeps every value operation unfused in the reference's order, so a source's
// contribution is bit-identical to the reference's; STRICT = false contracts the three value
// multiply-adds (lerp, gain ramp, accumulate) into FMAs. The cursor/index arithmetic is identical
// (and exact) in both.
#include <cuda_runtime.h>

#include "odb_kernels.h"
#include "odb_math.cuh"
#include "odb_async.cuh"
#include "odb_f32x2.cuh"

namespace odbk {

constexpr int FAST_WARPS = 16;                                              // one CTA per SM: 8 warp pairs, one warp per half tile
constexpr int FAST_ILP = 4;                                                 // frames of a chunk a lane processes interleaved
constexpr int FAST_SPLIT = 2;                                               // a 1024-frame tile is mixed as 2 halves of 512 frames
constexpr int FAST_HCHUNKS = ODB_TILE_CHUNKS / FAST_SPLIT;                  // 256-frame chunks per half: 2
constexpr int FAST_NACC = FAST_HCHUNKS * 8;                                 // packed (L, R) accumulators per lane: 512 frames / 32
constexpr int FAST_BATCH = 8;                                               // sources per warp batch: 8 x 2 ears x 2 chunks = 32 chains
constexpr int FAST_PCM_BYTES = ODB_FAST_PCM_CAP * 4;                        // 2560, two of them per warp
constexpr int FAST_POINTS = ODB_SPATIAL_CHUNK / 4;                          // every 4th cursor value of a chunk
constexpr int FAST_ROW_BYTES = FAST_POINTS * 8 + 8;                         // one (source, chunk) row of (L, R) cursors; +8 skews the banks
constexpr int FAST_OFFS_BYTES = FAST_BATCH * FAST_HCHUNKS * FAST_ROW_BYTES; // 8320
constexpr int FAST_WARP_BYTES = 2 * FAST_PCM_BYTES + FAST_OFFS_BYTES;       // 13440
static_assert(FAST_WARP_BYTES >= 2 * ODB_TILE_FRAMES * 4 / FAST_SPLIT, "the warp region doubles as its partial half tile");
static_assert(FAST_WARP_BYTES % 16 == 0, "TMA destinations are 16-byte aligned");
constexpr int FAST_REC_BYTES = 128;                                         // staged job record stride
constexpr int FAST_JOBS_OFF = FAST_WARPS * FAST_WARP_BYTES;                 // staged job records: 8 per warp
constexpr int FAST_BARS_OFF = FAST_JOBS_OFF + FAST_WARPS * FAST_BATCH * FAST_REC_BYTES;
constexpr int FAST_SMEM_BYTES = FAST_BARS_OFF + FAST_WARPS * 16;
static_assert(FAST_SMEM_BYTES <= 232448, "fits the 227 KB a CTA may use");

// One 256-frame chunk of one source. UL / UR: that ear is on the ds ~= 1 path (constant fraction, index
// base + i); otherwise its cursor comes from the chain checkpoints. FULL: every frame of the chunk is
// inside the tile. KL / KR: doppler ear = shared address of PCM index `base` minus the magic bits,
// unit ear = shared address of PCM index base + lane.
template <bool STRICT, bool FULL, bool UL, bool UR>
__device__ __forceinline__ void consume_chunk(u64* __restrict__ acc, const int cc, const int c, const int lane, const float fbase,
                                              const uint32_t row_sa, const uint32_t KL, const uint32_t KR, const u64 d1,
                                              const u64 d2, const u64 d3, const u64 fr_unit, const u64 pgp,
                                              const u64 dgp, const int nfr, const u64 nz) {
    const u64 magic = pk2(ODB_MAGIC, ODB_MAGIC);
    // The 8 frames a lane owns in this chunk are independent; they are written stage by stage over groups of
    // FAST_ILP frames so that the loads and the 4-cycle dependent FP32x2 steps of different frames interleave.
#pragma unroll
    for (int j0 = 0; j0 < 8; j0 += FAST_ILP) {
        if (!FULL && c * ODB_SPATIAL_CHUNK + 32 * j0 >= nfr) break;  // warp-uniform
        uint32_t aL[FAST_ILP], aR[FAST_ILP];
        u64 fr[FAST_ILP], o[FAST_ILP];
        if (!(UL && UR)) {
            // cursor of frame k = 32j + lane: checkpoint k & ~3, then (k & 3) literal `offset += ds` steps (frames.rs:195)
#pragma unroll
            for (int u = 0; u < FAST_ILP; u++) o[u] = lds_u64(row_sa + (uint32_t)(64 * (j0 + u)));
#pragma unroll
            for (int u = 0; u < FAST_ILP; u++) o[u] = add2(o[u], d1);
#pragma unroll
            for (int u = 0; u < FAST_ILP; u++) o[u] = add2(o[u], d2);
#pragma unroll
            for (int u = 0; u < FAST_ILP; u++) o[u] = add2(o[u], d3);
            // trunc = offset as isize; fract = offset - trunc as f32 (frames.rs:191-193), offset >= 0 here
            // all-FP32 index split: a round-down add of 2^23 leaves trunc(offset) in the low mantissa bits. (An
            // F2I.TRUNC / I2FP split was measured 8 % slower: the conversions are not full rate on sm_100a.)
            u64 t[FAST_ILP];
#pragma unroll
            for (int u = 0; u < FAST_ILP; u++) t[u] = add2_rm(o[u], magic);
#pragma unroll
            for (int u = 0; u < FAST_ILP; u++) {
                uint32_t tL, tR;
                upk2u(t[u], tL, tR);
                aL[u] = KL + (tL << 2);
                aR[u] = KR + (tR << 2);
            }
#pragma unroll
            for (int u = 0; u < FAST_ILP; u++) fr[u] = sub2(o[u], sub2(t[u], magic));

        }
        if (UL || UR) {  // frames.rs:183-187
            float u0, u1;
            upk2(fr_unit, u0, u1);
#pragma unroll
            for (int u = 0; u < FAST_ILP; u++) {
                float f0, f1;
                if (UL && UR) { f0 = u0; f1 = u1; }
                else { upk2(fr[u], f0, f1); if (UL) f0 = u0; else f1 = u1; }
                fr[u] = pk2(f0, f1);
                if (UL) aL[u] = KL + (uint32_t)(128 * (j0 + u));
                if (UR) aR[u] = KR + (uint32_t)(128 * (j0 + u));
            }
        }
        u64 a[FAST_ILP], b[FAST_ILP];
#pragma unroll
        for (int u = 0; u < FAST_ILP; u++) {   // get_pair (frames.rs:105-123); zeros come from the arena padding
            a[u] = pk2(lds_f32(aL[u]), lds_f32(aR[u]));
            b[u] = pk2(lds_f32_4(aL[u]), lds_f32_4(aR[u]));
        }
        u64 g[FAST_ILP], s[FAST_ILP];
#pragma unroll
        for (int u = 0; u < FAST_ILP; u++) {
            const float fi = fbase + (float)(32 * (j0 + u));  // `i as f32` (spatial.rs:459); exact small integer
            const u64 fi2 = pk2(fi, fi);
            if (STRICT) g[u] = mulx(fi2, dgp, nz);
            else g[u] = fma2(fi2, dgp, pgp);
        }
        if (STRICT) {
#pragma unroll
            for (int u = 0; u < FAST_ILP; u++) g[u] = add2(pgp, g[u]);  // prev_state.gain + i as f32 * d_gain
        }
#pragma unroll
        for (int u = 0; u < FAST_ILP; u++) b[u] = sub2(b[u], a[u]);     // frame::lerp = a + t * (b - a) (frame.rs:39-41)
        if (STRICT) {
#pragma unroll
            for (int u = 0; u < FAST_ILP; u++) s[u] = mulx(fr[u], b[u], nz);
#pragma unroll
            for (int u = 0; u < FAST_ILP; u++) s[u] = add2(a[u], s[u]);
        } else {
#pragma unroll
            for (int u = 0; u < FAST_ILP; u++) s[u] = fma2(fr[u], b[u], a[u]);
        }
        if (!FULL) {
#pragma unroll
            for (int u = 0; u < FAST_ILP; u++)
                if (c * ODB_SPATIAL_CHUNK + 32 * (j0 + u) + lane >= nfr) s[u] = 0ull;  // frame beyond the tile: contributes +0
        }
        if (STRICT) {
#pragma unroll
            for (int u = 0; u < FAST_ILP; u++) s[u] = mulx(s[u], g[u], nz);
#pragma unroll
            for (int u = 0; u < FAST_ILP; u++) acc[cc * 8 + j0 + u] = add2(acc[cc * 8 + j0 + u], s[u]);  // o[ear] += s * gain (spatial.rs:460)
        } else {
#pragma unroll
            for (int u = 0; u < FAST_ILP; u++) acc[cc * 8 + j0 + u] = fma2(s[u], g[u], acc[cc * 8 + j0 + u]);
        }
    }
}

// Words of a staged job record after the warp's prologue: lane q rewrites source q's record (a private copy per
// warp) with what this warp's half needs, so that the per-source set-up is four broadcast 16-byte loads.
#define FJ_SRC 0     // words 0-1: global address of the first float of this half's PCM window
#define FJ_BYTES 2   // bytes of the window
#define FJ_CODE 10   // bit 2: full tile, bit 1: left ear on the ds ~= 1 path, bit 0: right ear
#define FJ_K 12      // words 12-15: per chunk of this half (KL, KR): byte offset of PCM index `base` inside the window,
                     // minus the magic bits for a doppler ear


template <bool STRICT, bool FULL, bool UL, bool UR>
__device__ __forceinline__ void consume_source(u64* __restrict__ acc, const int lane, const float lanef, const int half,
                                               const uint32_t rec_sa, const uint32_t rec_x, const uint32_t pcm_b,
                                               const uint32_t rows_sa, const uint4 K, const int nfr, const u64 d1, const u64 d2, const u64 d3,
                                               const u64 pgp, const u64 dgp, const u64 nz) {
    const uint32_t lane4 = (uint32_t)(lane * 4);
#pragma unroll
    for (int cc = 0; cc < FAST_HCHUNKS; cc++) {
        const int c = half * FAST_HCHUNKS + cc;  // chunk of the tile
        if (!FULL && c * ODB_SPATIAL_CHUNK >= nfr) break;
        const uint32_t KL = pcm_b + (cc ? K.z : K.x) + (UL ? lane4 : 0u);
        const uint32_t KR = pcm_b + (cc ? K.w : K.y) + (UR ? lane4 : 0u);
        u64 fr_unit = 0ull;
        if (UL || UR)
            fr_unit = pk2(__uint_as_float(lds_u32(rec_sa + ((uint32_t)((ODB_JW_OFF0 + c) * 4) ^ rec_x))),
                          __uint_as_float(lds_u32(rec_sa + ((uint32_t)((ODB_JW_OFF0 + ODB_TILE_CHUNKS + c) * 4) ^ rec_x))));
        consume_chunk<STRICT, FULL, UL, UR>(acc, cc, c, lane, lanef + (float)(c * ODB_SPATIAL_CHUNK),
                                            rows_sa + (uint32_t)(cc * FAST_ROW_BYTES), KL, KR, d1, d2, d3, fr_unit, pgp, dgp,
                                            nfr, nz);
    }
}

template <bool STRICT>
__global__ void __launch_bounds__(FAST_WARPS * 32, 1) k_mix_fast(const OdbJob* __restrict__ jobs, int n_sources,
                                                                  float* __restrict__ partials, const u64 nz) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int half = warp & 1, pair = warp >> 1;  // the two warps of a pair mix the two halves of the same sources
    const int tl = blockIdx.y;
    const uint32_t smem_sa = smem_u32(smem_raw);
    const uint32_t pcm_sa = smem_sa + (uint32_t)(warp * FAST_WARP_BYTES);
    const uint32_t offs_sa = pcm_sa + 2 * FAST_PCM_BYTES;
    const uint32_t jobs_sa = smem_sa + (uint32_t)(FAST_JOBS_OFF + warp * FAST_BATCH * FAST_REC_BYTES);
    const uint32_t bar_sa = smem_sa + (uint32_t)(FAST_BARS_OFF + warp * 16);
    if (lane == 0) {
        mbar_init(bar_sa, 1);
        mbar_init(bar_sa + 8, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    pdl_launch_dependents();
    pdl_wait();           // the job records come from the walk kernel launched just before
    uint32_t parity = 0;  // bit b = phase parity of PCM buffer b's barrier
    uint32_t buf = 0;     // PCM buffer the next source to consume lands in

    u64 acc[FAST_NACC];
#pragma unroll
    for (int j = 0; j < FAST_NACC; j++) acc[j] = 0ull;

    const int gp = blockIdx.x * (FAST_WARPS / FAST_SPLIT) + pair, GP = gridDim.x * (FAST_WARPS / FAST_SPLIT);
    const float lanef = (float)(tl * ODB_TILE_FRAMES + lane);
    const int r = lane & 3;
    const OdbJob* tile_jobs = jobs + (size_t)tl * n_sources;
    const int n_batches = (n_sources + FAST_BATCH - 1) / FAST_BATCH;
    const int first_frame = half * (ODB_TILE_FRAMES / FAST_SPLIT);
    const int c0 = half * FAST_HCHUNKS;

    // Staged job records: record q's word w lives at position w ^ 4q of its 128-byte line, so that the eight lanes
    // of the prologue (one per source, same field) and the chain lanes hit different banks; aligned 8- and 16-byte
    // groups stay aligned groups. rec(q, w) = shared address of word w of record q.
    auto rec = [&](int q, int w) { return jobs_sa + (uint32_t)(q * FAST_REC_BYTES) + (uint32_t)((w * 4) ^ (q * 16)); };

    // lane 0: start the bulk copy of source q's PCM window of this half into PCM buffer `b`
    // (every lane loads the same descriptor; one elected lane arms the barrier and issues the copy)
    auto start_copy = [&](int q, uint32_t b) {
        const uint4 d = lds_u128(rec(q, FJ_SRC));
        const u64 p = ((u64)d.y << 32) | (u64)d.x;
        asm volatile(
            "{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\t"
            "@p mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t"
            "@p cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%2], [%3], %1, [%0];\n\t}" ::"r"(bar_sa + b * 8),
            "r"(d.z), "r"(pcm_sa + b * FAST_PCM_BYTES), "l"(p)
            : "memory");
    };

    for (int bi = gp; bi < n_batches; bi += GP) {
        const int s0 = bi * FAST_BATCH;
        // 1. stage the batch's job records (one 128-byte line each: lane l moves word l)
#pragma unroll
        for (int q = 0; q < FAST_BATCH; q++) {
            uint32_t w = lane == ODB_JW_FLAGS ? ODB_JF_SKIP : 0u;
            if (s0 + q < n_sources) w = __ldg(reinterpret_cast<const uint32_t*>(tile_jobs + s0 + q) + lane);
            sts_u32(rec(q, lane), w);
        }
        __syncwarp();
        // lane q < 8 derives what this half needs of source q and rewrites its private record (FJ_* words)
        bool mine = false;
        if (lane < FAST_BATCH) {
            const uint4 h = lds_u128(rec(lane, 0));                                     // pcm lo/hi, len, flags
            const int nfr = (int)lds_u32(rec(lane, ODB_JW_N_FRAMES));
            mine = !(h.w & (ODB_JF_SKIP | ODB_JF_GENERAL)) && nfr > first_frame;       // staged, and with frames in this half
            if (mine) {
                const uint2 win = lds_u64x(rec(lane, ODB_JW_WINDOW + 2 * half));        // first PCM index, floats
                const int w_start = (int)win.x;
                const u64 p = (((u64)h.y << 32) | (u64)h.x) + (u64)((long long)w_start * 4);
                const uint32_t mL = (h.w & ODB_JF_FAST_L) ? 0u : (ODB_MAGIC_BITS << 2);
                const uint32_t mR = (h.w & ODB_JF_FAST_R) ? 0u : (ODB_MAGIC_BITS << 2);
                const uint2 bL = lds_u64x(rec(lane, ODB_JW_BASE + c0));                   // `base` of this half's two chunks
                const uint2 bR = lds_u64x(rec(lane, ODB_JW_BASE + ODB_TILE_CHUNKS + c0));
                const uint32_t code = (nfr == ODB_TILE_FRAMES ? 4u : 0u) | ((h.w & ODB_JF_FAST_L) ? 2u : 0u) | ((h.w & ODB_JF_FAST_R) ? 1u : 0u);
                sts_u32(rec(lane, FJ_SRC), (uint32_t)p);
                sts_u32(rec(lane, FJ_SRC + 1), (uint32_t)(p >> 32));
                sts_u32(rec(lane, FJ_BYTES), win.y * 4u);
                sts_u32(rec(lane, FJ_CODE), code);
                sts_u32(rec(lane, FJ_K + 0), (uint32_t)(((int)bL.x - w_start) * 4) - mL);
                sts_u32(rec(lane, FJ_K + 1), (uint32_t)(((int)bR.x - w_start) * 4) - mR);
                sts_u32(rec(lane, FJ_K + 2), (uint32_t)(((int)bL.y - w_start) * 4) - mL);
                sts_u32(rec(lane, FJ_K + 3), (uint32_t)(((int)bR.y - w_start) * 4) - mR);
            }
        }
        __syncwarp();
        uint32_t act = __ballot_sync(0xffffffffu, mine);  // sources of the batch this warp mixes
        if (act) {
            start_copy(__ffs(act) - 1, buf);
            {   // 2. literal cursor chains: lane = (source q, ear e, chunk cc of this half)
                const int q = lane >> 2, e = (lane >> 1) & 1, cc = lane & 1;
                const uint32_t jf = lds_u32(rec(q, ODB_JW_FLAGS));
                if (((act >> q) & 1u) && !(jf & (e ? ODB_JF_FAST_R : ODB_JF_FAST_L))) {
                    float o = __uint_as_float(lds_u32(rec(q, ODB_JW_OFF0 + ODB_TILE_CHUNKS * e + c0 + cc)));
                    const float ds = __uint_as_float(lds_u32(rec(q, ODB_JW_DS + e)));
                    const uint32_t dst = offs_sa + (uint32_t)((q * FAST_HCHUNKS + cc) * FAST_ROW_BYTES + e * 4);
#pragma unroll 8
                    for (int m = 0; m < FAST_POINTS; m++) {
                        sts_f32(dst + (uint32_t)(m * 8), o);
                        o = __fadd_rn(o, ds); o = __fadd_rn(o, ds); o = __fadd_rn(o, ds); o = __fadd_rn(o, ds);
                    }
                }
            }
            __syncwarp();
            // 3. consume the batch source by source
            while (act) {
                const int q = __ffs(act) - 1;
                act &= act - 1u;
                if (act) start_copy(__ffs(act) - 1, buf ^ 1u);  // next window -> the other buffer
                const uint4 A = lds_u128(rec(q, ODB_JW_DS));   // ds (L, R), prev gain (L, R)
                const uint4 B = lds_u128(rec(q, ODB_JW_DG));   // d_gain (L, R), code, n_frames
                const uint4 K = lds_u128(rec(q, FJ_K));
                const u64 dsp = ((u64)A.y << 32) | A.x, pgp = ((u64)A.w << 32) | A.z, dgp = ((u64)B.y << 32) | B.x;
                const int nfr = (int)B.w;
                // lane's frames are k = lane + 32 j: k & 3 = r literal steps after the checkpoint, the rest add +0.0 (exact)
                const u64 d1 = r >= 1 ? dsp : 0ull, d2 = r >= 2 ? dsp : 0ull, d3 = r >= 3 ? dsp : 0ull;
                const uint32_t pcm_b = pcm_sa + buf * FAST_PCM_BYTES;
                const uint32_t rows_sa = offs_sa + (uint32_t)(q * FAST_HCHUNKS * FAST_ROW_BYTES + (lane >> 2) * 8);
                const uint32_t off0_sa = jobs_sa + (uint32_t)(q * FAST_REC_BYTES), off0_x = (uint32_t)(q * 16);
                mbar_wait(bar_sa + buf * 8, (parity >> buf) & 1u);
                parity ^= 1u << buf;
#define ODB_CONSUME(F, L, R) consume_source<STRICT, F, L, R>(acc, lane, lanef, half, off0_sa, off0_x, pcm_b, rows_sa, K, nfr, d1, d2, d3, pgp, dgp, nz)
                if (B.z == 4u) ODB_CONSUME(true, false, false);  // the common case: full tile, both ears on the doppler path
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

    // 4. fold: warp -> CTA (fixed warp order) -> one partial tile per CTA
#pragma unroll
    for (int j = 0; j < FAST_NACC; j++)
        asm volatile("st.shared.b64 [%0], %1;" ::"r"(pcm_sa + (uint32_t)((32 * j + lane) * 8)), "l"(acc[j]) : "memory");
    __syncthreads();
    float* dst = partials + ((size_t)tl * gridDim.x + blockIdx.x) * (2 * ODB_TILE_FRAMES);
    constexpr int HALF_FLOATS = 2 * ODB_TILE_FRAMES / FAST_SPLIT;
    for (int f = threadIdx.x; f < 2 * ODB_TILE_FRAMES; f += FAST_WARPS * 32) {
        const int h = f / HALF_FLOATS, fh = f - h * HALF_FLOATS;
        float sum = 0.0f;
#pragma unroll
        for (int w = 0; w < FAST_WARPS / FAST_SPLIT; w++)
            sum = sum + *reinterpret_cast<const float*>(smem_raw + (w * FAST_SPLIT + h) * FAST_WARP_BYTES + fh * 4);
        dst[f] = sum;
    }
}

}  // namespace odbk

using namespace odbk;

int odb_mix_fast_ctas(int n_sources, int sm_count) {
    const int per_cta = (FAST_WARPS / FAST_SPLIT) * FAST_BATCH;
    int want = (n_sources + per_cta - 1) / per_cta;
    return want < 1 ? 1 : (want > sm_count ? sm_count : want);
}

template <bool STRICT>
static cudaError_t launch_fast(const OdbJob* jobs, int n_sources, int n_tiles, float* partials, int n_ctas, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(k_mix_fast<STRICT>, cudaFuncAttributeMaxDynamicSharedMemorySize, FAST_SMEM_BYTES);
    if (e != cudaSuccess) return e;
    dim3 grid(n_ctas, n_tiles);
    return odb_launch_pdl(k_mix_fast<STRICT>, grid, dim3(FAST_WARPS * 32), (size_t)FAST_SMEM_BYTES, st, jobs, n_sources, partials,
                          0x8000000080000000ull);
}

// mode bit 0: value multiply-adds contracted to FMA
cudaError_t odb_launch_mix_fast(const OdbJob* jobs, int n_sources, int n_tiles, float* partials, int n_ctas, int mode,
                                cudaStream_t st) {
    return (mode & 1) ? launch_fast<false>(jobs, n_sources, n_tiles, partials, n_ctas, st)
                      : launch_fast<true>(jobs, n_sources, n_tiles, partials, n_ctas, st);
}
