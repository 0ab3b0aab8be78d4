/* Plain-C user of libssdk (include/ssdk.h): no C++, no torch, no Python.
 *
 *   gcc -std=c99 -I include -I /usr/local/cuda/include examples/c_abi_example.c \
 *       -L single-shot-detector_b200/lib -lssdk -L /usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/single-shot-detector_b200/lib -o /tmp/ssdk_example
 *
 * It generates the anchors of a 640x896 image with the reference's default generator (detector/anchor_generator.py:13-16
 * with three scale multipliers), assigns targets for one ground-truth box (detector/training_target_creation.py:5-45) and
 * prints the number of matched anchors.  Without a CUDA device it prints the library's error message and exits 0 (the
 * library has no CPU fallback; ssdk_num_anchors is host-only arithmetic and always works). */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <cuda_runtime_api.h>

#include "ssdk.h"

#define CHECK(call)                                                                  \
    do {                                                                             \
        int s_ = (call);                                                             \
        if (s_ != SSDK_OK) {                                                         \
            fprintf(stderr, "%s failed (%d): %s\n", #call, s_, ssdk_last_error());   \
            return 1;                                                                \
        }                                                                            \
    } while (0)

int main(void) {
    const int strides[5] = {8, 16, 32, 64, 128};
    const float base_scales[5] = {32, 64, 128, 256, 512};
    const double multipliers[3] = {1.0, 1.2599210498948732, 1.5874010519681994};
    const float aspect[3] = {1.0f, 2.0f, 0.5f};
    float scales[5 * 9], ratios[9];
    int32_t per_level[5];
    int64_t A = 0;
    int l, m, r;
    for (m = 0; m < 3; ++m)
        for (r = 0; r < 3; ++r) ratios[m * 3 + r] = aspect[r];                        /* itertools.product order (:70-71) */
    for (l = 0; l < 5; ++l)
        for (m = 0; m < 3; ++m)
            for (r = 0; r < 3; ++r) scales[l * 9 + m * 3 + r] = (float)(multipliers[m] * base_scales[l]);   /* :75 */

    printf("libssdk version %d\n", ssdk_version());
    CHECK(ssdk_num_anchors(640, 896, strides, 5, 9, &A, per_level));
    printf("anchors: %lld (per level %d %d %d %d %d)\n", (long long)A, per_level[0], per_level[1], per_level[2], per_level[3], per_level[4]);
    if (A != 107415) return 2;

    ssdk_ctx* ctx = NULL;
    if (ssdk_ctx_create(0, NULL, &ctx) != SSDK_OK) {
        printf("no GPU work done: %s\n", ssdk_last_error());
        return 0;
    }
    float *d_anchors = NULL, *d_gt = NULL, *d_reg = NULL;
    int32_t *d_labels = NULL, *d_cls = NULL, *d_matches = NULL;
    double* d_count = NULL;
    const float gt[4] = {0.30f, 0.25f, 0.62f, 0.55f};
    const int32_t label = 17;
    double count = -1.0;
    float first[4];
    if (cudaMalloc((void**)&d_anchors, (size_t)A * 16) != cudaSuccess || cudaMalloc((void**)&d_gt, 16) != cudaSuccess ||
        cudaMalloc((void**)&d_labels, 4) != cudaSuccess || cudaMalloc((void**)&d_reg, (size_t)A * 16) != cudaSuccess ||
        cudaMalloc((void**)&d_cls, (size_t)A * 4) != cudaSuccess || cudaMalloc((void**)&d_matches, (size_t)A * 4) != cudaSuccess ||
        cudaMalloc((void**)&d_count, 8) != cudaSuccess) {
        fprintf(stderr, "cudaMalloc failed\n");
        return 1;
    }
    cudaMemcpy(d_gt, gt, 16, cudaMemcpyHostToDevice);
    cudaMemcpy(d_labels, &label, 4, cudaMemcpyHostToDevice);
    CHECK(ssdk_anchors(ctx, 640, 896, strides, scales, ratios, 5, 9, d_anchors, NULL));
    CHECK(ssdk_training_targets_count(ctx, d_anchors, A, d_gt, d_labels, NULL, 1, 1, 0.5, 0.4, d_reg, d_cls, d_matches, d_count));
    CHECK(ssdk_ctx_synchronize(ctx));
    cudaMemcpy(&count, d_count, 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(first, d_anchors, 16, cudaMemcpyDeviceToHost);
    printf("first anchor [%.6f %.6f %.6f %.6f]\n", first[0], first[1], first[2], first[3]);
    printf("matched anchors for one %.0fx%.0f px box: %.0f\n", (gt[2] - gt[0]) * 640, (gt[3] - gt[1]) * 896, count);
    if (!(count >= 1.0)) return 3;                     /* forced matching guarantees at least one */
    /* bad arguments are reported, not crashed on (the reference's assert at training_target_creation.py:86) */
    if (ssdk_training_targets_count(ctx, d_anchors, A, d_gt, d_labels, NULL, 1, 1, 0.3, 0.4, d_reg, d_cls, d_matches, d_count) != SSDK_ERR_ARG) return 4;
    printf("expected error: %s\n", ssdk_last_error());
    cudaFree(d_anchors); cudaFree(d_gt); cudaFree(d_labels); cudaFree(d_reg); cudaFree(d_cls); cudaFree(d_matches); cudaFree(d_count);
    CHECK(ssdk_ctx_destroy(ctx));
    printf("ok\n");
    return 0;
}
