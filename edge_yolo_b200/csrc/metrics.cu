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


// =====================================================================================================================
// ap_per_class / compute_ap (ultralytics/utils/metrics.py:505-623) and scale_boxes / clip_boxes (utils/ops.py:92-127, 319-338)
// =====================================================================================================================
// The reference runs ap_per_class in numpy on the host after copying every detection of the validation set back: argsort by
// confidence, then per class boolean masks, cumsums, two 1000-point interpolations and, per IoU threshold, a precision envelope,
// a 101-point interpolation and a trapezoid sum -- all in float64.  Here:
//   ap_keys      order-preserving 64-bit keys (confidence bits : ~index), sorted by the library's bitonic sort -> np.argsort(-conf, stable)
//   ap_gather    confidence / class index / TP flags in sorted order (TP flags threshold-major), predictions per class (atomics)
//   ap_offsets   exclusive scan of the per-class counts
//   ap_scan      one CTA per class: ballot-compaction of the class's detections in confidence order with running TP counts
//                -> conf_c [n_p], tpc [T][n_p] (fpc = position + 1 - tpc)
//   ap_curves    one CTA per class: for every threshold the precision envelope (block-wide reverse running maximum, float64) and
//                numpy's interp semantics (largest j with xp[j] <= x, exact hit returns fp[j]) on the 101-point grid, trapezoid
//                sum; for threshold 0 also the 1000-point precision-at-recall, recall and precision curves.
// Everything after the curves (F1, smoothing, arg-max, rounding of tp / fp) is O(nc x 1000) host work on the returned arrays.
constexpr int kApGrid = 1000, kApPts = 101;

__device__ __forceinline__ uint32_t float_order_key(float f) {  // larger float -> larger key (NaN sorts high)
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__global__ void __launch_bounds__(256) ap_keys(const float* __restrict__ conf, int64_t N, unsigned long long* __restrict__ keys, int* __restrict__ count,
                                               int* __restrict__ n_pred, int nc) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) keys[i] = ((unsigned long long)float_order_key(conf[i]) << 32) | (uint32_t)(~(uint32_t)i);
    if (i == 0) *count = (int)N;
    if (i < nc) n_pred[i] = 0;
}

__global__ void __launch_bounds__(256) ap_gather(const unsigned long long* __restrict__ keys, const uint8_t* __restrict__ tp, const float* __restrict__ conf,
                                                 const float* __restrict__ pred_cls, int64_t N, int T, const float* __restrict__ classes, int nc,
                                                 float* __restrict__ conf_s, int* __restrict__ ci_s, uint8_t* __restrict__ tp_s, int* __restrict__ n_pred) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= N) return;
    const uint32_t idx = ~(uint32_t)keys[k];
    conf_s[k] = conf[idx];
    const float c = pred_cls[idx];
    int lo = 0, hi = nc - 1, ci = -1;  // classes ascending (np.unique)
    while (lo <= hi) {
        const int mid = (lo + hi) >> 1;
        const float v = classes[mid];
        if (v == c) { ci = mid; break; }
        if (v < c) lo = mid + 1; else hi = mid - 1;
    }
    ci_s[k] = ci;
    for (int t = 0; t < T; ++t) tp_s[(int64_t)t * N + k] = tp[(int64_t)idx * T + t];
    if (ci >= 0) atomicAdd(&n_pred[ci], 1);
}

__global__ void __launch_bounds__(1024) ap_offsets(const int* __restrict__ n_pred, int nc, int64_t* __restrict__ off) {
    __shared__ int64_t s_w[32];
    __shared__ int64_t s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < nc; base += 1024) {
        const int i = base + threadIdx.x;
        const int64_t v = i < nc ? n_pred[i] : 0;
        int64_t incl = v;
        for (int o = 1; o < 32; o <<= 1) { const int64_t t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) s_w[warp] = incl;
        __syncthreads();
        int64_t pre = s_carry;
        for (int w = 0; w < warp; ++w) pre += s_w[w];
        if (i < nc) off[i] = pre + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = pre + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) off[nc] = s_carry;
}

// one CTA per class: compaction in confidence order + running TP counts per threshold (ballots: a warp handles 32 consecutive
// sorted detections, the CTA's warps take consecutive groups; the carries chain through shared memory)
__global__ void __launch_bounds__(256) ap_scan(const int* __restrict__ ci_s, const float* __restrict__ conf_s, const uint8_t* __restrict__ tp_s, int64_t N, int T,
                                               const int64_t* __restrict__ off, float* __restrict__ conf_c, int* __restrict__ tpc) {
    __shared__ int s_cnt[8][kMaxThr + 1];
    __shared__ int s_base[kMaxThr + 1];
    const int ci = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t o0 = off[ci], np_all = off[gridDim.x];
    if (off[ci + 1] == o0) return;
    if (threadIdx.x <= kMaxThr) s_base[threadIdx.x] = 0;
    __syncthreads();
    for (int64_t k0 = 0; k0 < N; k0 += 256) {
        const int64_t k = k0 + threadIdx.x;
        const bool m = k < N && ci_s[k] == ci;
        const unsigned bm = __ballot_sync(0xffffffffu, m);
        unsigned bt[kMaxThr];
        for (int t = 0; t < T; ++t) bt[t] = __ballot_sync(0xffffffffu, m && tp_s[(int64_t)t * N + k] != 0);
        if (lane == 0) {
            s_cnt[warp][0] = __popc(bm);
            for (int t = 0; t < T; ++t) s_cnt[warp][1 + t] = __popc(bt[t]);
        }
        __syncthreads();
        if (m) {
            const unsigned below = (1u << lane) - 1u;
            int pos = s_base[0] + __popc(bm & below);
            for (int w = 0; w < warp; ++w) pos += s_cnt[w][0];
            conf_c[o0 + pos] = conf_s[k];
            for (int t = 0; t < T; ++t) {
                int c = s_base[1 + t] + __popc(bt[t] & (below | (1u << lane)));
                for (int w = 0; w < warp; ++w) c += s_cnt[w][1 + t];
                tpc[(int64_t)t * np_all + o0 + pos] = c;
            }
        }
        __syncthreads();
        if (threadIdx.x <= T) {
            int add = 0;
            for (int w = 0; w < 8; ++w) add += s_cnt[w][threadIdx.x];
            s_base[threadIdx.x] += add;
        }
        __syncthreads();
    }
}

// numpy's arr_interp for one query on xp (non-decreasing, length n >= 1) given as a functor; returns the interval index j
// (-1: left of xp[0]; n: right of xp[n-1]; else the largest j with xp[j] <= x)
template <typename XP>
__device__ __forceinline__ int64_t interp_index(double x, int64_t n, XP xp) {
    if (x > xp(n - 1)) return n;
    if (x < xp(0)) return -1;
    int64_t lo = 0, hi = n - 1;  // invariant: xp(lo) <= x, answer in [lo, hi]
    while (lo < hi) {
        const int64_t mid = (lo + hi + 1) >> 1;
        if (xp(mid) <= x) lo = mid; else hi = mid - 1;
    }
    return lo;
}
template <typename XP, typename FP>
__device__ __forceinline__ double interp_eval(double x, int64_t n, XP xp, FP fp, double left, double right) {
    const int64_t j = interp_index(x, n, xp);
    if (j == -1) return left;
    if (j == n) return right;
    if (j == n - 1) return fp(j);
    const double xj = xp(j);
    if (xj == x) return fp(j);
    const double slope = (fp(j + 1) - fp(j)) / (xp(j + 1) - xj);
    return slope * (x - xj) + fp(j);
}

__global__ void __launch_bounds__(256) ap_curves(const float* __restrict__ conf_c, const int* __restrict__ tpc, const int64_t* __restrict__ off,
                                                 const int64_t* __restrict__ n_labels, int T, double eps, double* __restrict__ env_ws,
                                                 double* __restrict__ ap, double* __restrict__ p_curve, double* __restrict__ r_curve,
                                                 double* __restrict__ prec_values) {
    __shared__ double s_w[8];
    __shared__ double s_carry;
    __shared__ double s_y[kApPts];
    const int ci = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t o0 = off[ci], n_p = off[ci + 1] - o0, np_all = off[gridDim.x];
    const int64_t n_l = n_labels[ci];
    if (n_p == 0 || n_l == 0) return;  // outputs were zeroed by the caller (metrics.py:571-572: `continue`)
    const double denom = (double)n_l + eps;
    const int64_t M = n_p + 2;                      // mrec / mpre with their sentinels
    double* env = env_ws + o0 + 2 * (int64_t)ci;    // this class's envelope, reused across thresholds
    const double step1000 = 1.0 / (double)(kApGrid - 1), step101 = 1.0 / (double)(kApPts - 1);
    for (int t = 0; t < T; ++t) {
        const int* tc = tpc + (int64_t)t * np_all + o0;
        auto mrec = [&](int64_t j) -> double { return j == 0 ? 0.0 : (j == M - 1 ? 1.0 : (double)tc[j - 1] / denom); };
        auto mpre = [&](int64_t j) -> double { return j == 0 ? 1.0 : (j == M - 1 ? 0.0 : (double)tc[j - 1] / (double)j); };  // tpc / (tpc + fpc) = tpc / position
        // ---- precision envelope: np.flip(np.maximum.accumulate(np.flip(mpre))), tiles of 256 from the end
        if (tid == 0) s_carry = -1.0;
        __syncthreads();
        for (int64_t hi = M; hi > 0; hi -= 256) {
            const int64_t j = hi - 1 - tid;
            double v = j >= 0 ? mpre(j) : -1.0;
            for (int o = 1; o < 32; o <<= 1) { const double u = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v = fmax(v, u); }
            if (lane == 31) s_w[warp] = v;
            __syncthreads();
            double pre = s_carry;
            for (int w = 0; w < warp; ++w) pre = fmax(pre, s_w[w]);
            v = fmax(v, pre);
            if (j >= 0) env[j] = v;
            __syncthreads();
            if (tid == 255) s_carry = v;
            __syncthreads();
        }
        __threadfence_block();
        auto envf = [&](int64_t j) -> double { return env[j]; };
        // ---- AP: trapezoid over the 101-point interpolation of the envelope
        if (tid < kApPts) {
            const double x = tid == kApPts - 1 ? 1.0 : (double)tid * step101;
            s_y[tid] = interp_eval(x, M, mrec, envf, envf(0), envf(M - 1));
        }
        __syncthreads();
        if (tid == 0) {
            double acc = 0.0;
            for (int i = 0; i + 1 < kApPts; ++i) {
                const double x0 = (double)i * step101, x1 = i + 1 == kApPts - 1 ? 1.0 : (double)(i + 1) * step101;
                acc += (x1 - x0) * (s_y[i + 1] + s_y[i]) / 2.0;
            }
            ap[(int64_t)ci * T + t] = acc;
        }
        if (t == 0) {
            auto xpc = [&](int64_t j) -> double { return -(double)conf_c[o0 + j]; };
            auto rec = [&](int64_t j) -> double { return (double)tc[j] / denom; };
            auto pre = [&](int64_t j) -> double { return (double)tc[j] / (double)(j + 1); };
            for (int q = tid; q < kApGrid; q += 256) {
                const double x = q == kApGrid - 1 ? 1.0 : (double)q * step1000;
                prec_values[(int64_t)ci * kApGrid + q] = interp_eval(x, M, mrec, envf, envf(0), envf(M - 1));
                r_curve[(int64_t)ci * kApGrid + q] = interp_eval(-x, n_p, xpc, rec, 0.0, rec(n_p - 1));   // left=0
                p_curve[(int64_t)ci * kApGrid + q] = interp_eval(-x, n_p, xpc, pre, 1.0, pre(n_p - 1));   // left=1
            }
        }
        __syncthreads();  // the envelope is rewritten by the next threshold
    }
}

// scale_boxes + clip_boxes for rows of >= 4 fp32 values, in place: (x - pad) / gain, clamp to [0, w] / [0, h]
__global__ void __launch_bounds__(256) scale_boxes_kernel(float* __restrict__ boxes, int64_t n, int64_t stride, float pad_x, float pad_y, float gain,
                                                          int padding, int xywh, float clip_w, float clip_h) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float* p = boxes + i * stride;
    float x1 = p[0], y1 = p[1], x2 = p[2], y2 = p[3];
    if (padding) {
        x1 = __fsub_rn(x1, pad_x); y1 = __fsub_rn(y1, pad_y);
        if (!xywh) { x2 = __fsub_rn(x2, pad_x); y2 = __fsub_rn(y2, pad_y); }
    }
    x1 = __fdiv_rn(x1, gain); y1 = __fdiv_rn(y1, gain); x2 = __fdiv_rn(x2, gain); y2 = __fdiv_rn(y2, gain);
    p[0] = fminf(fmaxf(x1, 0.f), clip_w); p[1] = fminf(fmaxf(y1, 0.f), clip_h);
    p[2] = fminf(fmaxf(x2, 0.f), clip_w); p[3] = fminf(fmaxf(y2, 0.f), clip_h);
}

struct ApLayout { size_t keys, count, conf_s, ci_s, tp_s, off, conf_c, tpc, env, total; int64_t key_stride; };
static ApLayout ap_layout(int64_t N, int T, int nc) {
    ApLayout L{};
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const uint32_t p2 = pow2ceil((uint32_t)(N > 1 ? N : 1));
    L.key_stride = p2 <= 16384u ? p2 : (int64_t)ceil_div(p2, 16384) * 16384;
    size_t o = 0;
    L.keys = o; o = al(o + 8 * (size_t)L.key_stride);
    L.count = o; o = al(o + 4);
    L.conf_s = o; o = al(o + 4 * (size_t)N);
    L.ci_s = o; o = al(o + 4 * (size_t)N);
    L.tp_s = o; o = al(o + (size_t)N * T);
    L.off = o; o = al(o + 8 * (size_t)(nc + 1));
    L.conf_c = o; o = al(o + 4 * (size_t)N);
    L.tpc = o; o = al(o + 4 * (size_t)N * T);
    L.env = o; o = al(o + 8 * ((size_t)N + 2 * (size_t)nc));
    L.total = o;
    return L;
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

extern "C" int el_ap_per_class_workspace_bytes(int64_t N, int T, int nc, size_t* bytes) {
    if (!bytes || N < 0 || N >= ((int64_t)1 << 31) || T <= 0 || T > kMaxThr || nc < 0) return EL_ERR_ARG;
    *bytes = ap_layout(N, T, nc).total;
    return EL_OK;
}

extern "C" int el_ap_per_class(const uint8_t* tp, const float* conf, const float* pred_cls, int64_t N, int T, const float* classes,
                               const int64_t* n_labels, int nc, double eps, void* workspace, size_t workspace_bytes, double* ap, double* p_curve,
                               double* r_curve, double* prec_values, int32_t* n_pred, void* stream) {
    if (N < 0 || N >= ((int64_t)1 << 31) || T <= 0 || T > kMaxThr || nc < 0) return EL_ERR_ARG;
    if (nc == 0) return EL_OK;
    if (!classes || !n_labels || !ap || !p_curve || !r_curve || !prec_values || !n_pred || !workspace) return EL_ERR_ARG;
    if (N > 0 && (!tp || !conf || !pred_cls)) return EL_ERR_ARG;
    const ApLayout L = ap_layout(N, T, nc);
    if (workspace_bytes < L.total) return EL_ERR_WORKSPACE;
    cudaStream_t s = (cudaStream_t)stream;
    char* ws = (char*)workspace;
    cudaMemsetAsync(ap, 0, sizeof(double) * (size_t)nc * T, s);
    cudaMemsetAsync(p_curve, 0, sizeof(double) * (size_t)nc * kApGrid, s);
    cudaMemsetAsync(r_curve, 0, sizeof(double) * (size_t)nc * kApGrid, s);
    cudaMemsetAsync(prec_values, 0, sizeof(double) * (size_t)nc * kApGrid, s);
    cudaMemsetAsync(n_pred, 0, sizeof(int32_t) * (size_t)nc, s);
    if (N == 0) return check_launch();
    unsigned long long* keys = (unsigned long long*)(ws + L.keys);
    int* count = (int*)(ws + L.count);
    const int blocks = (int)ceil_div(N > nc ? N : nc, 256);
    ap_keys<<<blocks, 256, 0, s>>>(conf, N, keys, count, n_pred, nc);
    sort_keys_desc(keys, L.key_stride, count, (int)N, 1, s);
    ap_gather<<<(int)ceil_div(N, 256), 256, 0, s>>>(keys, tp, conf, pred_cls, N, T, classes, nc, (float*)(ws + L.conf_s), (int*)(ws + L.ci_s),
                                                   (uint8_t*)(ws + L.tp_s), n_pred);
    ap_offsets<<<1, 1024, 0, s>>>(n_pred, nc, (int64_t*)(ws + L.off));
    ap_scan<<<nc, 256, 0, s>>>((const int*)(ws + L.ci_s), (const float*)(ws + L.conf_s), (const uint8_t*)(ws + L.tp_s), N, T, (const int64_t*)(ws + L.off),
                               (float*)(ws + L.conf_c), (int*)(ws + L.tpc));
    ap_curves<<<nc, 256, 0, s>>>((const float*)(ws + L.conf_c), (const int*)(ws + L.tpc), (const int64_t*)(ws + L.off), n_labels, T, eps, (double*)(ws + L.env), ap,
                                 p_curve, r_curve, prec_values);
    note_launches(6);
    return check_launch();
}

extern "C" int el_scale_boxes(float* boxes, int64_t n, int64_t row_stride, float pad_x, float pad_y, float gain, int padding, int xywh, float clip_w,
                              float clip_h, void* stream) {
    if (n < 0 || row_stride < 4 || (n > 0 && !boxes) || !(gain > 0.f)) return EL_ERR_ARG;
    if (n == 0) return EL_OK;
    scale_boxes_kernel<<<(int)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(boxes, n, row_stride, pad_x, pad_y, gain, padding, xywh, clip_w, clip_h);
    note_launches(1);
    return check_launch();
}
