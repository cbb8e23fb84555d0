// Validator metrics on the device (the step after NMS in `val`):
//   el_box_iou            ultralytics/utils/metrics.py:55-71   box_iou(box1 (N,4), box2 (M,4), eps=1e-7) -> (N,M), xyxy boxes
//   el_match_predictions  ultralytics/engine/validator.py:222-262, the non-scipy branch of DetectionValidator.match_predictions:
//                         per IoU threshold, pairs (label, detection) of equal class with IoU >= threshold are sorted by IoU
//                         (descending), every detection keeps its best label, every label then keeps the FIRST of its detections in
//                         detection order; correct[d][t] marks the survivors.
// The reference does this on the host in numpy after a D2H copy of the IoU matrix, once per image and threshold.  Closed form used
// here (bit-identical decisions): best[d] = argmax over labels of the class-masked IoU (it does not depend on the threshold), and for
// threshold t label l keeps min{ d : best[d] = l and iou[best[d]][d] >= thr_t }.  Exactly tied IoUs of one detection go to the larger
// label index (what numpy's stable ascending argsort, reversed, yields); the reference's default quicksort leaves that case open.
// IoU arithmetic is written with explicit fp32 intrinsics (no FMA contraction): the >= threshold decisions must equal numpy's.
#include <climits>

#include "el_internal.h"

namespace el {

__device__ __forceinline__ float iou_xyxy(const float4 a, const float4 b, float eps) {
    const float w = fmaxf(__fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)), 0.f);
    const float h = fmaxf(__fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y)), 0.f);
    const float inter = __fmul_rn(w, h);
    const float area_a = __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y));
    const float area_b = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
    // inter / (area1[:, None] + area2 - inter + eps), evaluated left to right
    return __fdiv_rn(inter, __fadd_rn(__fsub_rn(__fadd_rn(area_a, area_b), inter), eps));
}

__global__ void __launch_bounds__(256) box_iou_kernel(const float* __restrict__ a, int64_t a_stride, const float* __restrict__ b, int64_t b_stride,
                                                      float* __restrict__ out, int N, int M, float eps) {
    const int64_t total = (int64_t)N * M;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int n = (int)(i / M), m = (int)(i - (int64_t)n * M);
        const float* pa = a + n * a_stride;
        const float* pb = b + m * b_stride;
        out[i] = iou_xyxy(make_float4(pa[0], pa[1], pa[2], pa[3]), make_float4(pb[0], pb[1], pb[2], pb[3]), eps);
    }
}

constexpr int kMaxThr = 16;

// one CTA per image; iou (L, D) row-major, pred_cls (D), true_cls (L); correct (D, T) bytes
__global__ void __launch_bounds__(256) match_predictions_kernel(const float* __restrict__ iou, const float* __restrict__ pred_cls,
                                                                const float* __restrict__ true_cls, const float* __restrict__ iouv, int L, int D, int T,
                                                                uint8_t* __restrict__ correct) {
    extern __shared__ int s_first[];  // [T][L]: smallest detection index that keeps label l at threshold t
    for (int i = threadIdx.x; i < T * L; i += blockDim.x) s_first[i] = INT_MAX;
    __syncthreads();
    for (int d0 = 0; d0 < D; d0 += blockDim.x) {  // uniform trip count: the barriers below are reached by every thread
        const int d = d0 + threadIdx.x;
        float best = 0.f;
        int bl = -1;
        if (d < D) {
            const float pc = pred_cls[d];
            for (int l = 0; l < L; ++l) {
                const float v = true_cls[l] == pc ? iou[(int64_t)l * D + d] : 0.f;  // iou * correct_class
                if (v >= best && v > 0.f) { best = v; bl = l; }                     // ties -> larger label index
            }
            if (bl >= 0)
                for (int t = 0; t < T; ++t)
                    if (best >= iouv[t]) atomicMin(&s_first[t * L + bl], d);
        }
        __syncthreads();
        if (d < D)
            for (int t = 0; t < T; ++t) correct[(int64_t)d * T + t] = 0;
    }
    __syncthreads();
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        // recompute this detection's best label (cheap) and test whether it is the label's first detection at each threshold
        float best = 0.f;
        int bl = -1;
        const float pc = pred_cls[d];
        for (int l = 0; l < L; ++l) {
            const float v = true_cls[l] == pc ? iou[(int64_t)l * D + d] : 0.f;
            if (v >= best && v > 0.f) { best = v; bl = l; }
        }
        if (bl >= 0)
            for (int t = 0; t < T; ++t)
                if (best >= iouv[t] && s_first[t * L + bl] == d) correct[(int64_t)d * T + t] = 1;
    }
}

}  // namespace el

using namespace el;

extern "C" int el_box_iou(const float* box1, int64_t stride1, const float* box2, int64_t stride2, float* out, int N, int M, float eps, void* stream) {
    if (!out || N < 0 || M < 0 || ((!box1 || !box2) && N > 0 && M > 0) || stride1 < 4 || stride2 < 4) return EL_ERR_ARG;
    if (N == 0 || M == 0) return EL_OK;
    const int64_t total = (int64_t)N * M;
    const int grid = (int)(ceil_div(total, 256) < 148 * 8 ? ceil_div(total, 256) : 148 * 8);
    box_iou_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(box1, stride1, box2, stride2, out, N, M, eps);
    note_launches(1);
    return check_launch();
}

extern "C" int el_match_predictions(const float* iou, const float* pred_cls, const float* true_cls, const float* iouv, int L, int D, int T,
                                    uint8_t* correct, void* stream) {
    if (L < 0 || D < 0 || T <= 0 || T > kMaxThr || !iouv || (D > 0 && (!correct || !pred_cls)) || (L > 0 && D > 0 && (!iou || !true_cls))) return EL_ERR_ARG;
    if (D == 0) return EL_OK;
    const size_t smem = (size_t)T * (L > 0 ? L : 1) * sizeof(int);
    if (smem > 200 * 1024) return EL_ERR_UNSUPPORTED;
    if (smem > 48 * 1024) cudaFuncSetAttribute(match_predictions_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    match_predictions_kernel<<<1, 256, smem, (cudaStream_t)stream>>>(iou, pred_cls, true_cls, iouv, L, D, T, correct);
    note_launches(1);
    return check_launch();
}
