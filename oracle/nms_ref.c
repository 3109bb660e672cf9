/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle, never linked into the product library.
 *
 * Greedy non-maximum suppression, restated from the published algorithm of
 * torchvision's CPU kernel (torchvision/csrc/ops/cpu/nms_kernel.cpp, reached from the
 * reference at python/src/utils.py:110 and :115 through torchvision.ops).  The
 * torchvision source is NOT vendored anywhere in this image (only its compiled _C.so),
 * and the reference pins no torchvision version; tests/test_oracle_nms.py pins this
 * restatement against the installed torch.ops.torchvision.nms (0.26.0, CPU).
 *
 * Semantics restated:
 *   - areas a_k = (x2-x1)*(y2-y1) in fp32;
 *   - visiting order = stable descending sort of the scores (NaN first, ties by index);
 *   - a box is kept unless a previously *kept* box suppresses it;
 *   - i suppresses j iff  inter/(a_i + a_j - inter) > threshold, the fp32 ratio promoted
 *     to double for the comparison (threshold is a double), inter = max(0,w)*max(0,h)
 *     with max/min evaluated as (a<b)?b:a / (b<a)?b:a  (NaN behaviour of std::max/min);
 *   - output = kept original indices in visiting order.
 *
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC -o libnms_ref.so nms_ref.c
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline float maxf_std(float a, float b) { return (a < b) ? b : a; }
static inline float minf_std(float a, float b) { return (b < a) ? b : a; }

/* descending, NaN greatest, -0 == +0, all NaNs equal: returns 1 if a must come before b */
static inline int before(float a, float b) {
    int an = isnan(a), bn = isnan(b);
    if (an || bn) return an && !bn;
    return a > b;
}

/* stable merge sort of indices by descending score */
static void sort_desc_stable(const float *s, int64_t *idx, int64_t *tmp, int64_t n) {
    for (int64_t w = 1; w < n; w *= 2) {
        for (int64_t lo = 0; lo < n; lo += 2 * w) {
            int64_t mid = lo + w < n ? lo + w : n, hi = lo + 2 * w < n ? lo + 2 * w : n;
            int64_t a = lo, b = mid, o = lo;
            while (a < mid && b < hi) {
                /* take from the right run only if it is strictly before the left head */
                if (before(s[idx[b]], s[idx[a]])) tmp[o++] = idx[b++];
                else tmp[o++] = idx[a++];
            }
            while (a < mid) tmp[o++] = idx[a++];
            while (b < hi) tmp[o++] = idx[b++];
        }
        memcpy(idx, tmp, (size_t)n * sizeof(int64_t));
    }
}

/* boxes: n x 4 fp32 (x1,y1,x2,y2); returns number kept, indices written to keep[0..k) */
int64_t oracle_nms_f32(const float *boxes, const float *scores, int64_t n, double thr, int64_t *keep) {
    if (n <= 0) return 0;
    int64_t *order = (int64_t *)malloc((size_t)n * sizeof(int64_t));
    int64_t *tmp = (int64_t *)malloc((size_t)n * sizeof(int64_t));
    float *area = (float *)malloc((size_t)n * sizeof(float));
    unsigned char *dead = (unsigned char *)calloc((size_t)n, 1);
    for (int64_t i = 0; i < n; ++i) {
        order[i] = i;
        area[i] = (boxes[4 * i + 2] - boxes[4 * i + 0]) * (boxes[4 * i + 3] - boxes[4 * i + 1]);
    }
    sort_desc_stable(scores, order, tmp, n);
    int64_t k = 0;
    for (int64_t p = 0; p < n; ++p) {
        int64_t i = order[p];
        if (dead[i]) continue;
        keep[k++] = i;
        const float ix1 = boxes[4 * i], iy1 = boxes[4 * i + 1], ix2 = boxes[4 * i + 2], iy2 = boxes[4 * i + 3];
        const float ia = area[i];
        for (int64_t q = p + 1; q < n; ++q) {
            int64_t j = order[q];
            if (dead[j]) continue;
            float xx1 = maxf_std(ix1, boxes[4 * j]), yy1 = maxf_std(iy1, boxes[4 * j + 1]);
            float xx2 = minf_std(ix2, boxes[4 * j + 2]), yy2 = minf_std(iy2, boxes[4 * j + 3]);
            float w = maxf_std(0.0f, xx2 - xx1), h = maxf_std(0.0f, yy2 - yy1);
            float inter = w * h;
            float ovr = inter / (ia + area[j] - inter);
            if ((double)ovr > thr) dead[j] = 1;
        }
    }
    free(order); free(tmp); free(area); free(dead);
    return k;
}
