/* examples/offline.rs through the C ABI from plain C: the compiled-code view of the drop-in boundary
 * (include/oddio_b200.h), with no Python in between. A 500 Hz boop flies past the listener at 50 m/s; every
 * 512-frame block is mixed on the GPU, quantised to 16-bit PCM there (`(sample * i16::MAX as f32) as i16`,
 * examples/offline.rs:39) and appended to a WAV file.
 *
 *   gcc -std=c11 -O2 -Iinclude examples/offline.c -Loddio_b200 -loddio_b200 -lm -o offline
 *   LD_LIBRARY_PATH=oddio_b200 ./offline [out.wav]
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "oddio_b200.h"

#define DURATION_SECS 3u  /* examples/offline.rs:1-4 */
#define RATE 44100u
#define BLOCK_SIZE 512u
#define SPEED 50.0f

#define CHECK(call)                                                              \
    do {                                                                         \
        int rc__ = (call);                                                       \
        if (rc__ != ODB_OK) {                                                    \
            fprintf(stderr, "%s -> %d: %s\n", #call, rc__, odb_last_error());    \
            return 1;                                                            \
        }                                                                        \
    } while (0)

static void put_u32(FILE* f, uint32_t v) { fputc(v & 255, f); fputc((v >> 8) & 255, f); fputc((v >> 16) & 255, f); fputc(v >> 24, f); }
static void put_u16(FILE* f, uint16_t v) { fputc(v & 255, f); fputc(v >> 8, f); }

int main(int argc, char** argv) {
    const char* path = argc > 1 ? argv[1] : "offline.wav";
    const uint32_t n_pcm = RATE * DURATION_SECS, n_blocks = RATE * DURATION_SECS / BLOCK_SIZE;

    /* oddio::Frames::from_iter(RATE, (0..RATE * DURATION_SECS).map(|i| (t * 500 * 2 * PI).sin() * 80)), offline.rs:7-15 */
    float* boop = (float*)malloc(n_pcm * sizeof(float));
    if (!boop) return 1;
    for (uint32_t i = 0; i < n_pcm; i++) {
        const float t = (float)i / (float)RATE;
        boop[i] = sinf(t * 500.0f * 2.0f * 3.14159265358979323846f) * 80.0f;
    }

    odb_ctx* ctx = NULL;
    odb_scene* scene = NULL;
    odb_frames frames = 0;
    odb_source spatial = 0;
    CHECK(odb_ctx_create(0, &ctx));
    CHECK(odb_frames_from_slice(ctx, RATE, 1, boop, n_pcm, &frames));
    CHECK(odb_scene_create(ctx, &scene));                                   /* SpatialScene::new, offline.rs:16 */
    odb_chain chain;
    memset(&chain, 0, sizeof chain);
    chain.frames = frames;                                                  /* FramesSignal::from(boop), offline.rs:18 */
    chain.start_seconds = 0.0;
    chain.speed = 1.0f;
    chain.gain_ratio = 1.0f;
    const float position[3] = {-SPEED, 10.0f, 0.0f}, velocity[3] = {SPEED, 0.0f, 0.0f};
    CHECK(odb_scene_play(scene, &chain, position, velocity, 0.1f, &spatial)); /* offline.rs:17-24 */

    FILE* f = fopen(path, "wb");
    if (!f) { perror(path); return 1; }
    const uint32_t data_bytes = n_blocks * BLOCK_SIZE * 2u * 2u;            /* hound::WavSpec { 2 ch, RATE, 16 bit, Int } */
    fwrite("RIFF", 1, 4, f); put_u32(f, 36u + data_bytes); fwrite("WAVEfmt ", 1, 8, f);
    put_u32(f, 16u); put_u16(f, 1u); put_u16(f, 2u); put_u32(f, RATE); put_u32(f, RATE * 4u); put_u16(f, 4u); put_u16(f, 16u);
    fwrite("data", 1, 4, f); put_u32(f, data_bytes);

    int16_t block[BLOCK_SIZE * 2];
    const float interval = 1.0f / (float)RATE;                              /* oddio::run, lib.rs:91 */
    for (uint32_t b = 0; b < n_blocks; b++) {
        CHECK(odb_scene_sample_i16(scene, interval, block, BLOCK_SIZE));    /* oddio::run + the `as i16` cast, offline.rs:35-41 */
        for (uint32_t i = 0; i < BLOCK_SIZE * 2u; i++) put_u16(f, (uint16_t)block[i]);
    }
    fclose(f);
    int finished = 0;
    CHECK(odb_spatial_is_finished(scene, spatial, &finished));
    printf("wrote %u frames to %s (source finished: %d)\n", n_blocks * BLOCK_SIZE, path, finished);
    CHECK(odb_scene_destroy(scene));
    CHECK(odb_frames_release(ctx, frames));
    CHECK(odb_ctx_destroy(ctx));
    free(boop);
    return 0;
}
