/* End-to-end callback time of BASELINE.json's C3 through the C ABI from compiled code (no Python): what a Rust/C
 * host sees. An audio thread calls odb_scene_run with a host tile; a control thread queues set_motion for 1/16 of
 * the sources during every callback (paced by a semaphore, as in bench.py's e2e pass). Developer tool for the GPU box:
 *
 *   gcc -std=gnu11 -O2 -Iinclude tools/e2e_native.c -Loddio_b200 -loddio_b200 -lm -lpthread -o /tmp/e2e_native
 *   LD_LIBRARY_PATH=oddio_b200 /tmp/e2e_native [sources=65536] [callbacks=32] [frames=1024]
 * (bench.py runs the in-tree build of this file, tools/e2e_native, for its `e2e` figure at N = 1.)
 */
#include <math.h>
#include <pthread.h>
#include <semaphore.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "oddio_b200.h"

#define RATE 48000u
static uint32_t FRAMES = 1024u;
#define WARMUP 3

#define CHECK(call)                                                           \
    do {                                                                      \
        int rc__ = (call);                                                    \
        if (rc__ != ODB_OK) {                                                 \
            fprintf(stderr, "%s -> %d: %s\n", #call, rc__, odb_last_error()); \
            exit(1);                                                          \
        }                                                                     \
    } while (0)

static double now_us(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3;
}
static uint32_t rng_state = 0x0DD10u;
static float urand(void) {  /* xorshift32 -> [0, 1) */
    rng_state ^= rng_state << 13; rng_state ^= rng_state >> 17; rng_state ^= rng_state << 5;
    return (float)(rng_state >> 8) * (1.0f / 16777216.0f);
}

typedef struct {
    odb_scene* scene;
    odb_source* srcs;
    float *pos, *vel;
    uint32_t n_src, n_upd, n_callbacks;
    sem_t go;
} Control;

static void* control_thread(void* arg) {
    Control* c = (Control*)arg;
    odb_source* ids = (odb_source*)malloc(c->n_upd * sizeof(odb_source));
    float* p = (float*)malloc(c->n_upd * 3 * sizeof(float));
    float* v = (float*)malloc(c->n_upd * 3 * sizeof(float));
    for (uint32_t k = 0; k < c->n_callbacks; k++) {
        sem_wait(&c->go);
        const float t = (float)((k + 1) * FRAMES) / (float)RATE;
        for (uint32_t i = 0; i < c->n_upd; i++) {  /* the game thread nudges sources along their announced trajectory */
            const uint32_t s = (uint32_t)(urand() * (float)c->n_src) % c->n_src;
            ids[i] = c->srcs[s];
            for (int d = 0; d < 3; d++) {
                v[3 * i + d] = c->vel[3 * s + d];
                p[3 * i + d] = c->pos[3 * s + d] + v[3 * i + d] * t;
            }
        }
        CHECK(odb_spatial_set_motion_many(c->scene, c->n_upd, ids, p, v, NULL));
    }
    free(ids); free(p); free(v);
    return NULL;
}

int main(int argc, char** argv) {
    const uint32_t n_src = argc > 1 ? (uint32_t)atoi(argv[1]) : 65536u;
    const uint32_t n_cb = argc > 2 ? (uint32_t)atoi(argv[2]) : 32u;
    if (argc > 3) FRAMES = (uint32_t)atoi(argv[3]);
    const uint32_t total = n_cb + WARMUP;
    const uint32_t L = RATE + (uint32_t)(1.16f * FRAMES * (float)(total + 2)) + 2048u;  /* start 1 s in, ds <= 1.16 */
    odb_ctx* ctx = NULL;
    odb_scene* scene = NULL;
    CHECK(odb_ctx_create(0, &ctx));
    CHECK(odb_scene_create(ctx, &scene));
    if (getenv("ODB_VARIANT")) CHECK(odb_set_kernel_variant(scene, (int)strtol(getenv("ODB_VARIANT"), NULL, 0)));  /* default: the library's */
    Control c;
    memset(&c, 0, sizeof c);
    c.scene = scene; c.n_src = n_src; c.n_upd = n_src / 16u ? n_src / 16u : 1u; c.n_callbacks = total;
    c.srcs = (odb_source*)malloc(n_src * sizeof(odb_source));
    c.pos = (float*)malloc(n_src * 3 * sizeof(float));
    c.vel = (float*)malloc(n_src * 3 * sizeof(float));
    float* pcm = (float*)malloc(L * sizeof(float));
    const double t_setup = now_us();
    for (uint32_t s = 0; s < n_src; s++) {  /* one private PCM block per source: every callback reads fresh HBM */
        const float w = 2.0f * 3.14159265f * (100.0f + 3900.0f * urand()) / (float)RATE, ph = 6.2831853f * urand();
        /* cheap stand-in for bench.py's sine + noise (a recurrence instead of 72k sinf calls per source) */
        float y0 = 0.5f * sinf(ph), y1 = 0.5f * sinf(ph + w);
        const float k2 = 2.0f * cosf(w);
        for (uint32_t i = 0; i < L; i++) {
            pcm[i] = y0 + 0.05f * (2.0f * urand() - 1.0f);
            const float y2 = k2 * y1 - y0;
            y0 = y1; y1 = y2;
        }
        odb_frames fr = 0;
        CHECK(odb_frames_from_slice(ctx, RATE, 1, pcm, L, &fr));
        float dir[3], n2 = 0.0f;
        for (int d = 0; d < 3; d++) { dir[d] = 2.0f * urand() - 1.0f; n2 += dir[d] * dir[d]; }
        const float r = (2.0f + 298.0f * urand()) / sqrtf(n2 > 1e-6f ? n2 : 1.0f);
        for (int d = 0; d < 3; d++) { c.pos[3 * s + d] = dir[d] * r; c.vel[3 * s + d] = 60.0f * urand() - 30.0f; }
        odb_chain chain;
        memset(&chain, 0, sizeof chain);
        chain.frames = fr; chain.start_seconds = 1.0; chain.speed = 1.0f; chain.gain_ratio = 1.0f;
        CHECK(odb_scene_play(scene, &chain, c.pos + 3 * s, c.vel + 3 * s, 0.1f, &c.srcs[s]));
        CHECK(odb_frames_release(ctx, fr));  /* the playing source keeps its own reference (Arc semantics) */
    }
    fprintf(stderr, "set-up: %u sources, %.1f GB of PCM, %.1f s\n", n_src, (double)n_src * L * 4e-9, (now_us() - t_setup) * 1e-6);
    sem_init(&c.go, 0, 0);
    pthread_t th;
    pthread_create(&th, NULL, control_thread, &c);
    float* out = (float*)malloc(FRAMES * 2 * sizeof(float));
    /* The output tile is the caller's own buffer; page-locking it through the ABI lets the callback kernel store the
     * tile straight into it (E2E_NO_PIN=1: leave it pageable - the library then stages through its own pinned tile). */
    const int pinned = getenv("E2E_NO_PIN") ? 0 : (odb_pin_buffer(ctx, out, FRAMES * 2 * sizeof(float)) == ODB_OK);
    double t0 = 0.0;
    for (uint32_t k = 0; k < total; k++) {
        if (k == WARMUP) t0 = now_us();
        sem_post(&c.go);
        CHECK(odb_scene_run(scene, RATE, out, FRAMES));
    }
    const double us = (now_us() - t0) / (double)n_cb;
    pthread_join(th, NULL);
    double acc = 0.0;
    for (uint32_t i = 0; i < FRAMES * 2; i++) acc += fabs((double)out[i]);
    printf("{\"sources\": %u, \"frames\": %u, \"callbacks\": %u, \"us_per_callback\": %.1f, \"source_frames_per_s\": %.4e, \"checksum\": %.6f, \"out_pinned\": %d}\n",
           n_src, FRAMES, n_cb, us, (double)n_src * FRAMES / (us * 1e-6), acc, pinned);
    if (pinned) CHECK(odb_unpin_buffer(ctx, out));
    CHECK(odb_scene_destroy(scene));
    CHECK(odb_ctx_destroy(ctx));
    return 0;
}
