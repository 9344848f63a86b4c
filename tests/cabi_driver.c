/* Plain-C consumer of libpawsome_cuda.so: what a foreign-language binding (Julia ccall, cgo, JNI…) does.
 * Builds a synthetic 240x320 u8 video with a dark disk, runs Tracker-create / fill / step on host frames
 * and the batched chained path, and prints the positions so the pytest wrapper can compare them with
 * the oracle.  Usage: cabi_driver <n_frames>.  Exit code 0 = all library calls succeeded. */
#include "pawsome.h"

#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define H 240
#define W 320
#define CHECK(call)                                                        \
    do {                                                                   \
        int rc__ = (call);                                                 \
        if (rc__ != PT_OK) {                                               \
            fprintf(stderr, "%s -> %d: %s\n", #call, rc__, pt_last_error()); \
            return 2;                                                      \
        }                                                                  \
    } while (0)

static void render(uint8_t *f, int cy, int cx)
{
    memset(f, 128, (size_t)H * W);
    for (int y = cy - 12; y <= cy + 12; ++y)
        for (int x = cx - 12; x <= cx + 12; ++x)
            if (y >= 0 && y < H && x >= 0 && x < W && (y - cy) * (y - cy) + (x - cx) * (x - cx) <= 144) f[(size_t)y * W + x] = 0;
}

int main(int argc, char **argv)
{
    const int T = argc > 1 ? atoi(argv[1]) : 12;
    if (pt_version() != PT_VERSION) { fprintf(stderr, "version mismatch\n"); return 3; }
    if (pt_device_count() < 1) { fprintf(stderr, "no device: %s\n", pt_last_error()); return 4; }
    printf("sigma %.9f l %d window %d\n", pt_sigma(25.0), pt_kernel_len(25.0), pt_default_window(25.0));

    uint8_t *frames = (uint8_t *)malloc((size_t)T * H * W);
    for (int t = 0; t < T; ++t) render(frames + (size_t)t * H * W, 100 + 3 * t - 1, 150 + 5 * t - 1); /* 1-based centre (100+3t, 150+5t) */

    /* Tracker(img, 25, (45,45), true); fill = mode(img); ij[t] = trckr(ij[t-1]) on host frames */
    pt_tracker *trk = NULL;
    CHECK(pt_tracker_create(H, W, 25.0, 45, 45, 1, PT_PIX_U8, 0, &trk));
    CHECK(pt_tracker_set_frame(trk, frames, W));
    int fill = -1;
    CHECK(pt_tracker_compute_fill(trk, &fill));
    printf("fill %d\n", fill);
    int gi = 97, gj = 154;
    for (int t = 0; t < T; ++t) {
        int oi, oj; float r;
        CHECK(pt_tracker_step_host(trk, frames + (size_t)t * H * W, W, gi, gj, &oi, &oj, &r));
        printf("single %d %d %d %.9g\n", t, oi, oj, r);
        gi = oi; gj = oj;
    }
    pt_tracker_destroy(trk);

    /* the same video twice as a batch of 2, chained on the device after one upload per step */
    pt_batch *b = NULL;
    CHECK(pt_batch_create(2, H, W, 25.0, 45, 45, 1, PT_PIX_U8, 0, &b));
    const void **ptrs = (const void **)malloc(sizeof(void *) * 2 * (size_t)T);
    for (int t = 0; t < T; ++t) { ptrs[2 * t] = frames + (size_t)t * H * W; ptrs[2 * t + 1] = frames + (size_t)t * H * W; }
    CHECK(pt_batch_set_frames(b, ptrs, W));
    int fills[2];
    CHECK(pt_batch_compute_fill(b, fills));
    int32_t g0[4] = {97, 154, 103, 146};
    CHECK(pt_batch_set_guess(b, g0));
    int32_t *ij = (int32_t *)malloc(sizeof(int32_t) * 4 * (size_t)T);
    float *resp = (float *)malloc(sizeof(float) * 2 * (size_t)T);
    CHECK(pt_batch_track_host(b, ptrs, T, W, 0, ij, resp));
    for (int t = 0; t < T; ++t) printf("batch %d %d %d %d %d\n", t, ij[4 * t], ij[4 * t + 1], ij[4 * t + 2], ij[4 * t + 3]);
    printf("kernel %s launches %lld\n", pt_batch_kernel_name(b), pt_batch_launch_count(b));

    /* error behaviour through the C ABI */
    if (pt_batch_step(b, NULL, NULL, NULL, NULL) != PT_OK) { fprintf(stderr, "chained step failed: %s\n", pt_last_error()); return 5; }
    if (pt_batch_set_window(b, 0, 3) != PT_ERR_ARG) return 6;
    pt_batch_destroy(b);
    free(frames); free(ptrs); free(ij); free(resp);
    printf("ok\n");
    return 0;
}
