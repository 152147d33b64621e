// Staged ("fast") spatial mix kernel for the seek path: SpatialScene's mix closure (spatial.rs:445-469)
// over FramesSignal::sample (frames.rs:176-201) for every source whose PCM window of one 1024-frame
// tile fits in shared memory.
//
// One warp per source, sources strided over a persistent grid:
//   1. lane 0 starts a bulk async copy (TMA, cp.async.bulk -> UBLKCP) of the source's PCM window
//      HBM -> this warp's shared-memory buffer, completion on a per-warp mbarrier;
//   2. while the copy is in flight, lanes 0..3 (one per 256-frame chunk, both ears packed in an
//      f32x2) walk the reference's serial cursor `offset += ds` (frames.rs:195) literally and store
//      every 4th cursor value to shared memory. The chain is the one part of the path that is not
//      associative: it is evaluated with exactly the reference's sequence of f32 additions, so frame
//      indices are bit-exact;
//   3. all 32 lanes consume: lane l owns frames l, l+32, ... of the tile, re-derives its cursor from
//      the stored checkpoint with <= 3 more literal additions, splits it into index and fraction with
//      a round-down magic add (no F2I/I2F), gathers the sample pair from shared memory, lerps
//      (frame.rs:39-41), applies the per-frame gain ramp (spatial.rs:459) and accumulates into 64
//      register accumulators (32 frames x 2 ears, ears packed as FP32x2: FADD2/FMUL2/FFMA2);
//   4. after its last source a warp parks the accumulators in shared memory, the CTA folds its warps
//      in a fixed order and writes one partial tile; k_reduce_tiles sums the partial tiles.
//
// STRICT = true keeps every value operation unfused in the reference's order, so a source's
// contribution is bit-identical to the reference's; STRICT = false contracts the three value
// multiply-adds (lerp, gain ramp, accumulate) into FMAs. The cursor/index arithmetic is identical
// (and exact) in both.
#include <cuda_runtime.h>

#include "odb_kernels.h"
#include "odb_math.cuh"

namespace odbk {

typedef unsigned long long u64;

__device__ __forceinline__ u64 pk2(float lo, float hi) {
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk2(u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ void upk2u(u64 v, uint32_t& lo, uint32_t& hi) { asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v)); }
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
    u64 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64 add2_rm(u64 a, u64 b) {  // round toward -inf: floor(x) + 2^23 for 0 <= x < 2^23
    u64 r;
    asm("add.rm.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64 sub2(u64 a, u64 b) {
    u64 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
    u64 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
// ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even under -fmad=false (the scalar forms are
// left alone), and also sees through fma(a, b, -0.0). The strict kernel therefore multiplies with an FFMA2
// whose addend is a (-0.0, -0.0) pair that arrives as a kernel argument: RN(a*b + -0) == RN(a*b) for every
// input including signed zeros, and the compiler cannot fold what it cannot see.
__device__ __forceinline__ u64 mulx(u64 a, u64 b, u64 neg_zero2) {
    u64 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(neg_zero2));
    return r;
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
    u64 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ float lds_f32_4(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1+4];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ u64 lds_u64(uint32_t addr) {
    u64 v;
    asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_v2u64(uint32_t addr, u64 a, u64 b) {
    asm volatile("st.shared.v2.b64 [%0], {%1, %2};" ::"r"(addr), "l"(a), "l"(b) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!ok);
}

constexpr int FAST_WARPS = 8;
constexpr int FAST_PCM_BYTES = ODB_FAST_PCM_CAP * 4;                        // 6144
constexpr int FAST_POINTS = ODB_SPATIAL_CHUNK / 4;                          // every 4th cursor value of a chunk
constexpr int FAST_OFFS_BYTES = ODB_TILE_CHUNKS * FAST_POINTS * 8;          // 2048: [chunk][point] (L, R)
constexpr int FAST_WARP_BYTES = FAST_PCM_BYTES + FAST_OFFS_BYTES;           // 8192 = one stereo tile, reused for the fold
static_assert(FAST_WARP_BYTES == 2 * ODB_TILE_FRAMES * 4, "the warp region doubles as its partial tile");
constexpr int FAST_SMEM_BYTES = FAST_WARPS * FAST_WARP_BYTES + FAST_WARPS * 8;
#define ODB_MAGIC 8388608.0f          // 2^23: ulp 1, so x +rd 2^23 = 2^23 + floor(x)
#define ODB_MAGIC_BITS 0x4B000000u

// One 256-frame chunk of one source, doppler (serial-cursor) path. FULL: every frame of the chunk is inside the tile.
template <bool STRICT, bool FULL>
__device__ __forceinline__ void consume_chunk_doppler(u64* __restrict__ acc, const int c, const int lane, const float fbase,
                                                      const uint32_t offs_sa, const uint32_t KL, const uint32_t KR,
                                                      const u64 d1, const u64 d2, const u64 d3, const u64 pgp,
                                                      const u64 dgp, const int nfr, const u64 nz) {
    const u64 magic = pk2(ODB_MAGIC, ODB_MAGIC);
#pragma unroll
    for (int j = 0; j < 8; j++) {
        if (!FULL && c * ODB_SPATIAL_CHUNK + 32 * j >= nfr) break;  // warp-uniform
        // cursor of frame k = 32j + lane: checkpoint k & ~3, then (k & 3) literal `offset += ds` steps (frames.rs:195)
        u64 o = lds_u64(offs_sa + (uint32_t)(((c * FAST_POINTS) + 8 * j) * 8) + (uint32_t)((lane >> 2) * 8));
        o = add2(o, d1);
        o = add2(o, d2);
        o = add2(o, d3);
        // trunc = offset as isize; fract = offset - trunc as f32 (frames.rs:191-193), offset >= 0 here
        const u64 t = add2_rm(o, magic);
        const u64 fl = sub2(t, magic);
        const u64 fr = sub2(o, fl);
        uint32_t tL, tR;
        upk2u(t, tL, tR);
        const uint32_t aL = KL + (tL << 2), aR = KR + (tR << 2);
        const u64 a = pk2(lds_f32(aL), lds_f32(aR));      // get_pair (frames.rs:105-123); zeros come from the arena padding
        const u64 b = pk2(lds_f32_4(aL), lds_f32_4(aR));
        const u64 d = sub2(b, a);                          // frame::lerp = a + t * (b - a) (frame.rs:39-41)
        const float fi = fbase + (float)(32 * j);          // `i as f32` (spatial.rs:459); exact small integer
        const u64 fi2 = pk2(fi, fi);
        u64 s, g;
        if (STRICT) {
            s = add2(a, mulx(fr, d, nz));
            g = add2(pgp, mulx(fi2, dgp, nz));             // prev_state.gain + i as f32 * d_gain
        } else {
            s = fma2(fr, d, a);
            g = fma2(fi2, dgp, pgp);
        }
        if (!FULL && c * ODB_SPATIAL_CHUNK + 32 * j + lane >= nfr) s = 0ull;  // frame beyond the tile: contributes +0
        if (STRICT) acc[c * 8 + j] = add2(acc[c * 8 + j], mulx(s, g, nz));     // o[ear] += s * gain (spatial.rs:460)
        else acc[c * 8 + j] = fma2(s, g, acc[c * 8 + j]);
    }
}

// Same for the ds ~= 1 path (frames.rs:180-187): constant fract, index base + i.
template <bool STRICT, bool FULL>
__device__ __forceinline__ void consume_chunk_unit(u64* __restrict__ acc, const int c, const int lane, const float fbase,
                                                   const uint32_t AL, const uint32_t AR, const u64 fr, const u64 pgp,
                                                   const u64 dgp, const int nfr, const u64 nz) {
#pragma unroll
    for (int j = 0; j < 8; j++) {
        if (!FULL && c * ODB_SPATIAL_CHUNK + 32 * j >= nfr) break;
        const uint32_t aL = AL + (uint32_t)(128 * j), aR = AR + (uint32_t)(128 * j);
        const u64 a = pk2(lds_f32(aL), lds_f32(aR));
        const u64 b = pk2(lds_f32_4(aL), lds_f32_4(aR));
        const u64 d = sub2(b, a);
        const float fi = fbase + (float)(32 * j);
        const u64 fi2 = pk2(fi, fi);
        u64 s, g;
        if (STRICT) {
            s = add2(a, mulx(fr, d, nz));
            g = add2(pgp, mulx(fi2, dgp, nz));
        } else {
            s = fma2(fr, d, a);
            g = fma2(fi2, dgp, pgp);
        }
        if (!FULL && c * ODB_SPATIAL_CHUNK + 32 * j + lane >= nfr) s = 0ull;
        if (STRICT) acc[c * 8 + j] = add2(acc[c * 8 + j], mulx(s, g, nz));
        else acc[c * 8 + j] = fma2(s, g, acc[c * 8 + j]);
    }
}

template <bool STRICT>
__global__ void __launch_bounds__(FAST_WARPS * 32, 2) k_mix_fast(const OdbJob* __restrict__ jobs, int n_sources,
                                                                  float* __restrict__ partials, const u64 nz) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tl = blockIdx.y;
    const uint32_t pcm_sa = smem_u32(smem_raw + warp * FAST_WARP_BYTES);
    const uint32_t offs_sa = pcm_sa + FAST_PCM_BYTES;
    const uint32_t bar_sa = smem_u32(smem_raw + FAST_WARPS * FAST_WARP_BYTES + warp * 8);
    if (lane == 0) mbar_init(bar_sa, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    uint32_t parity = 0;

    u64 acc[ODB_TILE_FRAMES / 32];
#pragma unroll
    for (int j = 0; j < ODB_TILE_FRAMES / 32; j++) acc[j] = 0ull;

    const int gw = blockIdx.x * FAST_WARPS + warp, GW = gridDim.x * FAST_WARPS;
    const float lanef = (float)(tl * ODB_TILE_FRAMES + lane);
    const int r = lane & 3;
    const OdbJob* tile_jobs = jobs + (size_t)tl * n_sources;

    for (int sidx = gw; sidx < n_sources; sidx += GW) {
        // the job is one 128-byte line: lane l holds word l
        const uint32_t jw = reinterpret_cast<const uint32_t*>(tile_jobs + sidx)[lane];
        const uint32_t jf = __shfl_sync(0xffffffffu, jw, ODB_JW_FLAGS);
        if (jf & (ODB_JF_SKIP | ODB_JF_GENERAL)) continue;
        const int w_start = (int)__shfl_sync(0xffffffffu, jw, ODB_JW_W_START);
        const uint32_t w_bytes = __shfl_sync(0xffffffffu, jw, ODB_JW_W_LEN) * 4u;
        const uint32_t pcm_hi = __shfl_sync(0xffffffffu, jw, ODB_JW_PCM_HI);
        if (lane == 0) {  // 1. PCM window HBM -> shared, asynchronously
            const float* pcm = reinterpret_cast<const float*>(((u64)pcm_hi << 32) | (u64)jw);
            mbar_expect_tx(bar_sa, w_bytes);
            bulk_g2s(pcm_sa, pcm + w_start, w_bytes, bar_sa);
        }
        const int nfr = (int)__shfl_sync(0xffffffffu, jw, ODB_JW_N_FRAMES);
        const bool unit = (jf & ODB_JF_FAST_L) != 0;  // both ears or neither (walk kernel guarantees)
        const float dsL = __uint_as_float(__shfl_sync(0xffffffffu, jw, ODB_JW_DS));
        const float dsR = __uint_as_float(__shfl_sync(0xffffffffu, jw, ODB_JW_DS + 1));
        const u64 dsp = pk2(dsL, dsR);
        const float o0L = __uint_as_float(__shfl_sync(0xffffffffu, jw, ODB_JW_OFF0 + r));
        const float o0R = __uint_as_float(__shfl_sync(0xffffffffu, jw, ODB_JW_OFF0 + ODB_TILE_CHUNKS + r));
        if (!unit) {
            if (lane < ODB_TILE_CHUNKS) {  // 2. literal cursor chains, one lane per chunk, ears packed
                u64 o = pk2(o0L, o0R);
                const uint32_t dst = offs_sa + (uint32_t)(lane * FAST_POINTS * 8);
#pragma unroll 4
                for (int m = 0; m < FAST_POINTS; m += 2) {
                    const u64 p0 = o;
                    o = add2(o, dsp); o = add2(o, dsp); o = add2(o, dsp); o = add2(o, dsp);
                    const u64 p1 = o;
                    o = add2(o, dsp); o = add2(o, dsp); o = add2(o, dsp); o = add2(o, dsp);
                    sts_v2u64(dst + (uint32_t)(m * 8), p0, p1);
                }
            }
            __syncwarp();
        }
        const u64 pgp = pk2(__uint_as_float(__shfl_sync(0xffffffffu, jw, ODB_JW_PG)),
                            __uint_as_float(__shfl_sync(0xffffffffu, jw, ODB_JW_PG + 1)));
        const u64 dgp = pk2(__uint_as_float(__shfl_sync(0xffffffffu, jw, ODB_JW_DG)),
                            __uint_as_float(__shfl_sync(0xffffffffu, jw, ODB_JW_DG + 1)));
        const u64 d1 = r >= 1 ? dsp : 0ull, d2 = r >= 2 ? dsp : 0ull, d3 = r >= 3 ? dsp : 0ull;
        mbar_wait(bar_sa, parity);
        parity ^= 1u;
        // 3. consume
        const bool full = nfr == ODB_TILE_FRAMES;
#pragma unroll
        for (int c = 0; c < ODB_TILE_CHUNKS; c++) {
            if (c * ODB_SPATIAL_CHUNK >= nfr) break;
            const int baseL = (int)__shfl_sync(0xffffffffu, jw, ODB_JW_BASE + c);
            const int baseR = (int)__shfl_sync(0xffffffffu, jw, ODB_JW_BASE + ODB_TILE_CHUNKS + c);
            const float fbase = lanef + (float)(c * ODB_SPATIAL_CHUNK);
            if (!unit) {
                const uint32_t KL = pcm_sa + (uint32_t)((baseL - w_start) * 4) - (ODB_MAGIC_BITS << 2);
                const uint32_t KR = pcm_sa + (uint32_t)((baseR - w_start) * 4) - (ODB_MAGIC_BITS << 2);
                if (full) consume_chunk_doppler<STRICT, true>(acc, c, lane, fbase, offs_sa, KL, KR, d1, d2, d3, pgp, dgp, nfr, nz);
                else consume_chunk_doppler<STRICT, false>(acc, c, lane, fbase, offs_sa, KL, KR, d1, d2, d3, pgp, dgp, nfr, nz);
            } else {
                const uint32_t AL = pcm_sa + (uint32_t)((baseL - w_start + lane) * 4);
                const uint32_t AR = pcm_sa + (uint32_t)((baseR - w_start + lane) * 4);
                const u64 fr = pk2(__uint_as_float(__shfl_sync(0xffffffffu, jw, ODB_JW_OFF0 + c)),
                                   __uint_as_float(__shfl_sync(0xffffffffu, jw, ODB_JW_OFF0 + ODB_TILE_CHUNKS + c)));
                if (full) consume_chunk_unit<STRICT, true>(acc, c, lane, fbase, AL, AR, fr, pgp, dgp, nfr, nz);
                else consume_chunk_unit<STRICT, false>(acc, c, lane, fbase, AL, AR, fr, pgp, dgp, nfr, nz);
            }
        }
        __syncwarp();  // every lane is done with the PCM and cursor buffers before they are refilled
    }

    // 4. fold: warp -> CTA (fixed warp order) -> one partial tile per CTA
    {
        const uint32_t tile_sa = pcm_sa;
#pragma unroll
        for (int j = 0; j < ODB_TILE_FRAMES / 32; j++)
            asm volatile("st.shared.b64 [%0], %1;" ::"r"(tile_sa + (uint32_t)((32 * j + lane) * 8)), "l"(acc[j]) : "memory");
    }
    __syncthreads();
    const float* all = reinterpret_cast<const float*>(smem_raw);
    float* dst = partials + ((size_t)tl * gridDim.x + blockIdx.x) * (2 * ODB_TILE_FRAMES);
    for (int f = threadIdx.x; f < 2 * ODB_TILE_FRAMES; f += FAST_WARPS * 32) {
        float sum = 0.0f;
#pragma unroll
        for (int w = 0; w < FAST_WARPS; w++) sum = sum + all[w * (2 * ODB_TILE_FRAMES) + f];
        dst[f] = sum;
    }
}

}  // namespace odbk

using namespace odbk;

int odb_mix_fast_ctas(int n_sources, int sm_count) {
    int want = (n_sources + FAST_WARPS - 1) / FAST_WARPS;
    int cap = sm_count * 2;
    return want < 1 ? 1 : (want > cap ? cap : want);
}

cudaError_t odb_launch_mix_fast(const OdbJob* jobs, int n_sources, int n_tiles, float* partials, int n_ctas, int strict,
                                cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k_mix_fast<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FAST_SMEM_BYTES);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(k_mix_fast<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FAST_SMEM_BYTES);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    dim3 grid(n_ctas, n_tiles);
    if (strict) k_mix_fast<true><<<grid, FAST_WARPS * 32, FAST_SMEM_BYTES, st>>>(jobs, n_sources, partials, 0x8000000080000000ull);
    else k_mix_fast<false><<<grid, FAST_WARPS * 32, FAST_SMEM_BYTES, st>>>(jobs, n_sources, partials, 0x8000000080000000ull);
    return cudaGetLastError();
}
