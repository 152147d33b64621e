// Scalar/vector math of the hot path with the reference's exact operation order.
// Compiled with -fmad=false: the reference (Rust) never contracts a*b+c, so neither may we in
// anything that feeds a cursor, an index or a per-source gain. IEEE div/sqrt (nvcc defaults
// -prec-div=true -prec-sqrt=true, -ftz=false) match the reference's `/` and `sqrt()`.
#pragma once
#include "odb_types.h"

namespace odbk {

// Programmatic dependent launch (sm_90+): the kernels of one callback are launched back to back with the
// programmatic-stream-serialization attribute. `pdl_launch_dependents` lets the next kernel of the stream be
// set up while this one still runs; `pdl_wait` blocks until everything the previous kernel wrote is visible.
// Both are no-ops for a kernel launched the ordinary way.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

struct V3 { float x, y, z; };

// math/mod.rs:33-35  norm = sqrt(((0 + x*x) + y*y) + z*z)
__device__ __forceinline__ float v_norm(V3 a) {
    float acc = 0.0f;
    acc = acc + a.x * a.x;
    acc = acc + a.y * a.y;
    acc = acc + a.z * a.z;
    return sqrtf(acc);
}
// math/mod.rs:37-43
__device__ __forceinline__ float v_dot(V3 a, V3 b) {
    float acc = 0.0f;
    acc = acc + a.x * b.x;
    acc = acc + a.y * b.y;
    acc = acc + a.z * b.z;
    return acc;
}
__device__ __forceinline__ V3 v_scale(V3 v, float f) { return {v.x * f, v.y * f, v.z * f}; }          // :45-47
__device__ __forceinline__ V3 v_sub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }       // :49-51
__device__ __forceinline__ V3 v_add(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }       // :53-55
__device__ __forceinline__ V3 v_mix(V3 a, V3 b, float r) {                                           // :57-60
    float ir = 1.0f - r;
    return {ir * a.x + r * b.x, ir * a.y + r * b.y, ir * a.z + r * b.z};
}
struct Q4 { float s, x, y, z; };
__device__ __forceinline__ Q4 q_mul(Q4 q, Q4 r) {                                                     // :69-79
    Q4 o;
    o.s = q.s * r.s - q.x * r.x - q.y * r.y - q.z * r.z;
    o.x = q.s * r.x + q.x * r.s + q.y * r.z - q.z * r.y;
    o.y = q.s * r.y - q.x * r.z + q.y * r.s + q.z * r.x;
    o.z = q.s * r.z + q.x * r.y - q.y * r.x + q.z * r.s;
    return o;
}
__device__ __forceinline__ V3 q_rotate(OdbQuat rq, V3 p) {                                           // :81-94
    Q4 rot = {rq.s, rq.x, rq.y, rq.z};
    Q4 inv = {rot.s, -rot.x, -rot.y, -rot.z};                                                        // :62-67
    Q4 pq = {0.0f, p.x, p.y, p.z};
    Q4 r = q_mul(rot, q_mul(pq, inv));
    return {r.x, r.y, r.z};
}

#define ODB_SPEED_OF_SOUND 343.0f   // spatial.rs:602
#define ODB_HEAD_RADIUS 0.1075f     // spatial.rs:605
#define ODB_POS_SMOOTHING 0.5f      // spatial.rs:520
#define ODB_GAIN_SMOOTHING 0.1f     // gain.rs:163
#define ODB_F32_EPSILON 1.1920929e-7f

// State::smoothed_position, spatial.rs:501-511
__device__ __forceinline__ V3 smoothed_position(V3 prev_position, float state_dt, float d, V3 position, V3 velocity) {
    float dt = state_dt + d;
    V3 position_change = v_scale(velocity, dt);
    V3 naive = v_add(prev_position, position_change);
    V3 intended = v_add(position, position_change);
    return v_mix(naive, intended, fminf(dt / ODB_POS_SMOOTHING, 1.0f));
}

struct EarSt { float offset, gain; };
// EarState::new, spatial.rs:531-549; ear 0 = Left, 1 = Right (Ear::pos :573-583, Ear::dir :586-598)
__device__ __forceinline__ EarSt ear_state(V3 p, int ear, float radius) {
    V3 epos = {ear == 0 ? -ODB_HEAD_RADIUS : ODB_HEAD_RADIUS, 0.0f, 0.0f};
    float distance = v_norm(v_sub(p, epos));
    float offset = distance * (-1.0f / ODB_SPEED_OF_SOUND);
    float distance_gain = radius / fmaxf(distance, radius);
    float sq17 = sqrtf(17.0f);
    V3 dir = {(ear == 0 ? -1.0f : 1.0f) * 4.0f / sq17, 0.0f, -1.0f / sq17};
    float stereo = 0.5f + (distance < 1e-3f ? 0.5f : v_dot(dir, v_scale(p, 0.5f / distance)));
    return {offset, stereo * distance_gain};
}

// frames.rs:105-123 get_pair for a mono Frames block. The arena keeps ODB_PCM_PAD zeros on each
// side of every block, so indices within the pad may be read directly and yield the zeros the
// reference substitutes; anything further out is forced to zero without touching memory.
__device__ __forceinline__ void get_pair_mono(const float* __restrict__ pcm, int len, long long k, float& a, float& b) {
    if (k >= -(long long)(ODB_PCM_PAD - 1) && k < (long long)len + (ODB_PCM_PAD - 2)) {
        a = pcm[k];
        b = pcm[k + 1];
    } else {
        a = 0.0f;
        b = 0.0f;
    }
}

__device__ __forceinline__ int sat_i32(long long v) {
    const long long lim = 1ll << 30;
    return (int)(v > lim ? lim : (v < -lim ? -lim : v));
}

// Loads a whole source record with back-to-back 16-byte loads. Written with `asm volatile` so that the loads
// are issued together at the top of the walk kernels: left to itself the compiler fetches each field right
// before its first use, which strings a dozen dependent HBM round trips along one thread (measured: 15 us
// for a kernel with ~1200 instructions per thread).
__device__ __forceinline__ void load_source(OdbSource& dst, const OdbSource* __restrict__ p) {
    static_assert(sizeof(OdbSource) % 16 == 0, "16-byte words");
    uint4* d = reinterpret_cast<uint4*>(&dst);
    const uint4* g = reinterpret_cast<const uint4*>(p);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(OdbSource) / 16); i++)
        asm volatile("ld.global.v4.u32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(d[i].x), "=r"(d[i].y), "=r"(d[i].z), "=r"(d[i].w)
                     : "l"(g + i));
}

// The part of walk_set that is common to both sets (spatial.rs:204-261) for one source: motion refresh,
// smoothed start / end positions in the listener's frame, State::dt advance, finished_for / stopped
// bookkeeping and the removal report. Returns false if the source is (now) stopped: it does not mix.
__device__ __forceinline__ bool walk_common(OdbSource* sp, const OdbSource& s, const OdbCallback& cb, uint32_t slot,
                                            uint32_t* __restrict__ removed, int removed_cap, V3& prev_position,
                                            V3& next_position, uint32_t& flags_out, const bool write = true) {
    const float elapsed = cb.elapsed;
    // --- motion refresh, spatial.rs:216-226
    V3 pos = {s.pos[0], s.pos[1], s.pos[2]}, vel = {s.vel[0], s.vel[1], s.vel[2]};
    V3 statep = {s.prev_position[0], s.prev_position[1], s.prev_position[2]};
    float state_dt = s.state_dt;
    uint32_t flags = s.flags;
    if (flags & ODB_SF_MOTION_FRESH) {
        V3 npos = {s.ppos[0], s.ppos[1], s.ppos[2]}, nvel = {s.pvel[0], s.pvel[1], s.pvel[2]};
        statep = (flags & ODB_SF_PENDING_DISC) ? npos : smoothed_position(statep, state_dt, 0.0f, pos, vel);
        state_dt = 0.0f;
        pos = npos; vel = nvel;
        flags &= ~ODB_SF_MOTION_FRESH;
        if (write) {
            sp->pos[0] = pos.x; sp->pos[1] = pos.y; sp->pos[2] = pos.z;
            sp->vel[0] = vel.x; sp->vel[1] = vel.y; sp->vel[2] = vel.z;
        }
    }
    // --- smoothed start/end positions in the listener's frame, spatial.rs:228-235
    prev_position = q_rotate(cb.prev_rot, smoothed_position(statep, state_dt, 0.0f, pos, vel));
    next_position = q_rotate(cb.rot, smoothed_position(statep, state_dt, elapsed, pos, vel));
    state_dt = state_dt + elapsed;  // :238
    if (write) {
        sp->prev_position[0] = statep.x; sp->prev_position[1] = statep.y; sp->prev_position[2] = statep.z;
        sp->state_dt = state_dt;
    }

    // --- finished / stopped bookkeeping, spatial.rs:243-261
    const bool was_stopped = (flags & ODB_SF_STOPPED) != 0;
    if (!was_stopped) {
        float distance = v_norm(prev_position);
        if (flags & ODB_SF_HAS_FINISHED_FOR) {
            if (s.finished_for > distance / ODB_SPEED_OF_SOUND) flags |= ODB_SF_STOPPED;
            else if (write) sp->finished_for = s.finished_for + elapsed;
        } else if (s.t >= s.t_end) {  // inner.is_finished(): frames.rs:204-206 through the wrappers
            flags |= ODB_SF_HAS_FINISHED_FOR;
            if (write) sp->finished_for = elapsed;
        }
    }
    if (write) sp->flags = flags;
    flags_out = flags;
    if (flags & ODB_SF_STOPPED) {
        if (!was_stopped && write) {  // set.remove(i): report the slot so the host can swap_remove it from its Vec
            uint32_t k = atomicAdd(removed, 1u);
            removed[1 + (k & (uint32_t)(removed_cap - 1))] = slot;
        }
        return false;
    }
    return true;
}

}  // namespace odbk
