/*
 * ORACLE (test infrastructure, not product) -- plain C restatement of
 * TensorFlow 1.12's NonMaxSuppressionV3 CPU kernel, the one piece of the
 * reference's hot path whose algorithm lives outside /root/reference
 * (called at reference detector/utils/nms.py:33, pinned only by
 * "tensorflow 1.12" in README.md:22).  Pinned to TensorFlow's own unit-test
 * vectors (tests/golden/tf_nms_vectors.py <- non_max_suppression_op_test.cc);
 * semantics restated from the published kernel
 * (tensorflow/core/kernels/non_max_suppression_op.cc):
 *   - candidates are the boxes with score > score_threshold (strict);
 *   - they are visited in descending score order (equal scores: lower index
 *     first -- TF 1.12's heap order is unspecified, TF>=2 uses this rule);
 *   - a candidate is dropped iff IoU > iou_threshold (strict) with an already
 *     selected box, scanning selected boxes newest first;
 *   - IoU = inter / (area_i + area_j - inter) in float32, corners min/max
 *     normalised, 0 when either area <= 0, no epsilon;
 *   - stop at max_output_size selections.
 * Built by oracle/Makefile into oracle/_build/liboracle.so.
 */
#include <stdlib.h>
#include <string.h>

typedef struct { float score; int index; } cand_t;

static int cand_cmp(const void* a, const void* b) {
    const cand_t* x = (const cand_t*)a; const cand_t* y = (const cand_t*)b;
    if (x->score > y->score) return -1;
    if (x->score < y->score) return 1;
    return (x->index > y->index) - (x->index < y->index);
}

static inline float fminf_(float a, float b) { return a < b ? a : b; }
static inline float fmaxf_(float a, float b) { return a > b ? a : b; }

static int iou_gt(const float* bi, const float* bj, float thr) {
    const float ymin_i = fminf_(bi[0], bi[2]), xmin_i = fminf_(bi[1], bi[3]);
    const float ymax_i = fmaxf_(bi[0], bi[2]), xmax_i = fmaxf_(bi[1], bi[3]);
    const float ymin_j = fminf_(bj[0], bj[2]), xmin_j = fminf_(bj[1], bj[3]);
    const float ymax_j = fmaxf_(bj[0], bj[2]), xmax_j = fmaxf_(bj[1], bj[3]);
    volatile float hi = ymax_i - ymin_i, wi = xmax_i - xmin_i;
    volatile float hj = ymax_j - ymin_j, wj = xmax_j - xmin_j;
    volatile float area_i = hi * wi, area_j = hj * wj;
    if (area_i <= 0.0f || area_j <= 0.0f) return 0;
    volatile float ih = fminf_(ymax_i, ymax_j) - fmaxf_(ymin_i, ymin_j);
    volatile float iw = fminf_(xmax_i, xmax_j) - fmaxf_(xmin_i, xmin_j);
    volatile float inter = fmaxf_(ih, 0.0f) * fmaxf_(iw, 0.0f);
    volatile float sum = area_i + area_j;
    volatile float uni = sum - inter;
    volatile float iou = inter / uni;
    return iou > thr;
}

/* boxes [n,4], scores [n] (stride score_stride floats, so one column of an [n,C]
 * array can be passed without a copy).  Writes selected indices (descending
 * score) to out[<=max_out]; returns how many. */
int oracle_nms_v3(const float* boxes, const float* scores, int score_stride, int n,
                  int max_out, float iou_thr, float score_thr, int* out) {
    cand_t* c = (cand_t*)malloc(sizeof(cand_t) * (size_t)(n > 0 ? n : 1));
    int m = 0, k = 0;
    for (int i = 0; i < n; ++i) {
        float s = scores[(size_t)i * score_stride];
        if (s > score_thr) { c[m].score = s; c[m].index = i; ++m; }
    }
    qsort(c, (size_t)m, sizeof(cand_t), cand_cmp);
    for (int t = 0; t < m && k < max_out; ++t) {
        int keep = 1;
        for (int j = k - 1; j >= 0; --j)
            if (iou_gt(boxes + 4 * (size_t)c[t].index, boxes + 4 * (size_t)out[j], iou_thr)) { keep = 0; break; }
        if (keep) out[k++] = c[t].index;
    }
    free(c);
    return k;
}

/* All classes of one image (reference detector/utils/nms.py:31-44): boxes [n,4],
 * scores [n,C] row-major; out_idx [C*max_out], out_cnt [C]. */
void oracle_multiclass_nms(const float* boxes, const float* scores, int n, int C,
                           int max_out, float iou_thr, float score_thr,
                           int* out_idx, int* out_cnt) {
    for (int c = 0; c < C; ++c)
        out_cnt[c] = oracle_nms_v3(boxes, scores + c, C, n, max_out, iou_thr, score_thr,
                                   out_idx + (size_t)c * max_out);
}
