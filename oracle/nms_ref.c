/* TEST INFRASTRUCTURE ONLY -- plain-C restatement of the greedy box NMS that the
 * reference reaches through `torchvision.ops.nms(boxes, scores, iou_thres)`
 * (call site: /root/reference/ultralytics/utils/ops.py:296; torchvision is a
 * third-party dependency pinned to 0.17.2 in the reference's requirements.txt:2 and
 * not vendored).  Published algorithm, as pinned empirically in SURVEY.md section 8c:
 *   1. order = stable argsort of scores, descending (ties: lower index first);
 *   2. sweep in that order; a box is suppressed iff an earlier KEPT box has
 *      inter / (area_a + area_b - inter) >  thr   (strict; no epsilon; IoU in fp32,
 *      compared against the threshold as a double -- the CPU kernel's signature is
 *      `double iou_threshold`, so 0.3 is NOT rounded to float32 first);
 *   3. return the kept original indices in sweep order.
 * Build with -ffp-contract=off so no FMA is formed (the CPU kernel in torchvision
 * is compiled without FMA contraction); checked against the installed
 * torchvision CPU op in tests/test_oracle_golden.py.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float s; int32_t i; } el_key;

static int el_cmp(const void *pa, const void *pb) {
    const el_key *a = (const el_key *)pa, *b = (const el_key *)pb;
    if (a->s > b->s) return -1;
    if (a->s < b->s) return 1;
    return (a->i > b->i) - (a->i < b->i); /* stable: ties by ascending index */
}

int el_oracle_nms(const float *boxes, const float *scores, int n, double thr, int64_t *keep) {
    if (n <= 0) return 0;
    el_key *order = (el_key *)malloc(sizeof(el_key) * (size_t)n);
    float *area = (float *)malloc(sizeof(float) * (size_t)n);
    uint8_t *dead = (uint8_t *)calloc((size_t)n, 1);
    for (int i = 0; i < n; ++i) {
        order[i].s = scores[i];
        order[i].i = i;
        const float *b = boxes + 4 * (size_t)i;
        area[i] = (b[2] - b[0]) * (b[3] - b[1]);
    }
    qsort(order, (size_t)n, sizeof(el_key), el_cmp);
    int k = 0;
    for (int a = 0; a < n; ++a) {
        int i = order[a].i;
        if (dead[i]) continue;
        keep[k++] = i;
        const float *bi = boxes + 4 * (size_t)i;
        for (int c = a + 1; c < n; ++c) {
            int j = order[c].i;
            if (dead[j]) continue;
            const float *bj = boxes + 4 * (size_t)j;
            float xx1 = bi[0] > bj[0] ? bi[0] : bj[0];
            float yy1 = bi[1] > bj[1] ? bi[1] : bj[1];
            float xx2 = bi[2] < bj[2] ? bi[2] : bj[2];
            float yy2 = bi[3] < bj[3] ? bi[3] : bj[3];
            float w = xx2 - xx1; if (!(w > 0.0f)) w = 0.0f;
            float h = yy2 - yy1; if (!(h > 0.0f)) h = 0.0f;
            float inter = w * h;
            float uni = area[i] + area[j];
            uni = uni - inter;
            float ovr = inter / uni;
            if ((double)ovr > thr) dead[j] = 1; /* torchvision compares the fp32 IoU with a double threshold */
        }
    }
    free(order); free(area); free(dead);
    return k;
}
