// Task-aligned target assignment on the device (the step between decode and the DFL / QFL loss kernels of
// every training step):
//   el_tal_assign   ultralytics/utils/tal.py:14-295  TaskAlignedAssigner.forward
//                     get_pos_mask            :120-130   candidates = anchor centre strictly inside a valid ground truth (:241-262)
//                     get_box_metrics         :132-155   metric = score[gt class]^alpha * CIoU(gt, pred).clamp(0)^beta   (CIoU: utils/metrics.py:74-134)
//                     select_topk_candidates  :157-190   the topk anchors of every ground truth by metric
//                     select_highest_overlaps :265-295   an anchor claimed by several ground truths goes to the one it overlaps most
//                     get_targets             :192-238   labels / boxes / one-hot scores of the assigned ground truth
//                     normalisation           :110-116   scores *= max_gt( metric * best_iou_of_gt / (best_metric_of_gt + eps) )
// The reference materialises ~20 dense (B, n_gt, A) tensors and a 10-iteration scatter_add_ loop; here it is three kernels over one
// (B, n_gt, A) workspace of metric / overlap / flag planes:
//   1. tal_metric_topk   one CTA per (image, ground truth): candidates, CIoU, metric, then `topk` rounds of a block-wide arg-max
//                        (ties -> lower anchor index; torch.topk leaves that order open, the loss does not depend on it: a tie can only
//                        involve zero metrics, whose target scores are zero)
//   2. tal_resolve       one thread per (image, anchor): multi-claim resolution, labels / boxes / fg / gt index, per-ground-truth
//                        maxima of metric and overlap over the final positives (integer atomicMax on non-negative floats: exact and
//                        order-independent)
//   3. tal_emit          128-row tiles of the dense (B*A, nc) target-score tensor: the normalised soft one-hot rows, 16-byte stores
// Arithmetic is written with explicit fp32 intrinsics in the reference's evaluation order (no FMA contraction), so the metrics, and
// with them the selected anchors, are those of the reference's own eager CUDA ops (IEEE division / sqrt, libdevice atanf / powf).
#include "el_internal.h"

namespace el {

constexpr int kTalThreads = 256;
constexpr uint8_t kTalPos = 1, kTalTaken = 2, kTalCand = 4;

// x.pow(p) as torch evaluates it for a Python-scalar exponent (ATen pow_tensor_scalar: 0.5 -> sqrt, 1 -> x, 2 -> x*x, 3 -> x*x*x, else pow)
__device__ __forceinline__ float pow_scalar(float x, float p) {
    if (p == 1.f) return x;
    if (p == 0.5f) return sqrtf(x);
    if (p == 2.f) return __fmul_rn(x, x);
    if (p == 3.f) return __fmul_rn(__fmul_rn(x, x), x);
    return powf(x, p);
}

// CIoU of xyxy boxes, utils/metrics.py:74-134 with xywh=False, CIoU=True, eps=1e-7; a = ground truth, b = prediction (iou_calculation, tal.py:153-155)
__device__ __forceinline__ float ciou_xyxy(const float4 a, const float4 b) {
    const float eps = 1e-7f;
    const float aw = __fsub_rn(a.z, a.x), ah = __fadd_rn(__fsub_rn(a.w, a.y), eps);
    const float bw = __fsub_rn(b.z, b.x), bh = __fadd_rn(__fsub_rn(b.w, b.y), eps);
    const float iw = fmaxf(__fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)), 0.f);
    const float ih = fmaxf(__fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y)), 0.f);
    const float inter = __fmul_rn(iw, ih);
    const float uni = __fadd_rn(__fsub_rn(__fadd_rn(__fmul_rn(aw, ah), __fmul_rn(bw, bh)), inter), eps);
    const float iou = __fdiv_rn(inter, uni);
    const float cw = __fsub_rn(fmaxf(a.z, b.z), fminf(a.x, b.x));
    const float ch = __fsub_rn(fmaxf(a.w, b.w), fminf(a.y, b.y));
    const float c2 = __fadd_rn(__fadd_rn(__fmul_rn(cw, cw), __fmul_rn(ch, ch)), eps);
    const float dx = __fsub_rn(__fsub_rn(__fadd_rn(b.x, b.z), a.x), a.z);
    const float dy = __fsub_rn(__fsub_rn(__fadd_rn(b.y, b.w), a.y), a.w);
    const float rho2 = __fdiv_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), 4.f);
    const float da = __fsub_rn(atanf(__fdiv_rn(bw, bh)), atanf(__fdiv_rn(aw, ah)));
    const float v = __fmul_rn(0.40528473456935109f /* 4 / pi^2 */, __fmul_rn(da, da));
    const float alpha = __fdiv_rn(v, __fadd_rn(__fsub_rn(v, iou), 1.0000001f /* 1 + eps */));
    return __fsub_rn(iou, __fadd_rn(__fdiv_rn(rho2, c2), __fmul_rn(v, alpha)));
}

__device__ __forceinline__ int clamp_label(float v, int nc) {
    const int l = (int)v;  // .long(): truncation
    return l < 0 ? 0 : (l >= nc ? nc - 1 : l);
}

// grid (n_gt, B).  metric / overlap / flags: (B, n_gt, A) planes of the workspace; best: (B, n_gt, 2) zeroed here for tal_resolve.
// The planes are written and re-read by this CTA between barriers: no __restrict__ / non-coherent loads on them.
// kSmem: the selection rounds scan a shared-memory copy of the CTA's metric row (A floats, taken entries set to -1) instead of the
// global planes -- ten dependent L2 round trips per thread and round were ~50 of the first version's 57 us.  Rows that do not fit
// (A > ~50 000 anchors) keep the global scan.
template <bool kSmem>
__global__ void __launch_bounds__(kTalThreads) tal_metric_topk_kernel(const float* __restrict__ scores, const float* __restrict__ boxes,
                                                                      const float* __restrict__ anchors, const float* __restrict__ gt_labels,
                                                                      const float* __restrict__ gt_boxes, const uint8_t* __restrict__ gt_valid,
                                                                      int A, int nc, int M, int topk, float alpha, float beta, float* metric,
                                                                      float* overlap, uint8_t* flags, int* best) {
    extern __shared__ float s_met[];
    const int m = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    const int64_t g = (int64_t)b * M + m;
    float* met = metric + g * A;
    float* ovl = overlap + g * A;
    uint8_t* flg = flags + g * A;
    if (tid == 0) best[2 * g] = best[2 * g + 1] = 0;
    const bool valid = gt_valid[g] != 0;
    const float4 gt = make_float4(gt_boxes[4 * g], gt_boxes[4 * g + 1], gt_boxes[4 * g + 2], gt_boxes[4 * g + 3]);
    const int lab = clamp_label(gt_labels[g], nc);
    for (int a = tid; a < A; a += kTalThreads) {
        const float ax = anchors[2 * a], ay = anchors[2 * a + 1];
        // min(anchor - x1y1, x2y2 - anchor) > 1e-9 (tal.py:258-262)
        const float d = fminf(fminf(__fsub_rn(ax, gt.x), __fsub_rn(ay, gt.y)), fminf(__fsub_rn(gt.z, ax), __fsub_rn(gt.w, ay)));
        const bool cand = valid && d > 1e-9f;
        float v = 0.f, u = 0.f;
        if (cand) {
            const float* pb = boxes + ((int64_t)b * A + a) * 4;
            u = fmaxf(ciou_xyxy(gt, make_float4(pb[0], pb[1], pb[2], pb[3])), 0.f);
            const float s = scores[((int64_t)b * A + a) * nc + lab];
            v = __fmul_rn(pow_scalar(s, alpha), pow_scalar(u, beta));
        }
        met[a] = v;
        ovl[a] = u;
        flg[a] = cand ? kTalCand : 0;
        if (kSmem) s_met[a] = v;
    }
    if (!valid) return;  // a padded ground truth selects nothing (uniform over the CTA)
    __shared__ unsigned long long s_key[kTalThreads / 32];
    __syncthreads();
    // key = metric bits (non-negative floats order like integers) : ~anchor  ->  max key = largest metric, lowest anchor index
    for (int r = 0; r < topk; ++r) {
        unsigned long long key = 0;  // real keys have non-zero low words (A < 2^31)
        for (int a = tid; a < A; a += kTalThreads) {
            float v;
            if (kSmem) {
                v = s_met[a];
                if (v < 0.f) continue;
            } else {
                if (flg[a] & kTalTaken) continue;
                v = met[a];
            }
            const unsigned long long k = ((unsigned long long)__float_as_uint(v) << 32) | (0xffffffffu - (unsigned)a);
            key = k > key ? k : key;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
            key = other > key ? other : key;
        }
        if ((tid & 31) == 0) s_key[tid >> 5] = key;
        __syncthreads();
        if (tid == 0) {
#pragma unroll
            for (int w = 1; w < kTalThreads / 32; ++w) key = s_key[w] > key ? s_key[w] : key;
            if (key) {
                const unsigned a = 0xffffffffu - (unsigned)(key & 0xffffffffu);
                const uint8_t f = flg[a];
                flg[a] = f | kTalTaken | ((f & kTalCand) ? kTalPos : 0);  // in the top k AND inside a valid ground truth (tal.py:126-128)
                if (kSmem) s_met[a] = -1.f;
            }
        }
        __syncthreads();
    }
}

// one thread per (image, anchor)
__global__ void __launch_bounds__(256) tal_resolve_kernel(const float* __restrict__ metric, const float* __restrict__ overlap,
                                                          const uint8_t* __restrict__ flags, const float* __restrict__ gt_labels,
                                                          const float* __restrict__ gt_boxes, int B, int A, int nc, int M, int* best,
                                                          int64_t* __restrict__ labels, float* __restrict__ tboxes, uint8_t* __restrict__ fg,
                                                          int64_t* __restrict__ gt_idx) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)B * A) return;
    const int b = (int)(i / A), a = (int)(i - (int64_t)b * A);
    const int64_t base = (int64_t)b * M * A + a;
    int claims = 0, first = 0, bm = 0;
    float bi = overlap[base];
    for (int m = 0; m < M; ++m) {
        const int64_t o = base + (int64_t)m * A;
        if (flags[o] & kTalPos) {
            if (claims == 0) first = m;
            ++claims;
        }
        const float u = overlap[o];
        if (u > bi) { bi = u; bm = m; }  // argmax over ground truths: first maximum (tal.py:286)
    }
    const bool is_fg = claims > 0;
    const int gi = claims > 1 ? bm : first;  // `first` is 0 for background anchors, like mask_pos.argmax(-2) (tal.py:294)
    const int64_t g = (int64_t)b * M + gi;
    labels[i] = clamp_label(gt_labels[g], nc);
    reinterpret_cast<float4*>(tboxes)[i] = make_float4(gt_boxes[4 * g], gt_boxes[4 * g + 1], gt_boxes[4 * g + 2], gt_boxes[4 * g + 3]);
    fg[i] = is_fg ? 1 : 0;
    gt_idx[i] = gi;
    if (is_fg) {  // pos_align_metrics / pos_overlaps: maxima over the final positives of this ground truth (tal.py:112-113)
        atomicMax(best + 2 * g, __float_as_int(metric[g * A + a]));
        atomicMax(best + 2 * g + 1, __float_as_int(overlap[g * A + a]));
    }
}

// target_scores = one_hot(label) * fg * metric * best_overlap / (best_metric + eps), dense (B*A, nc) rows.  A CTA takes kTalRows rows at
// a time: the first kTalRows threads fetch (label, value) of their row into shared memory, then all threads write the tile with V-wide
// stores (V = 4 when nc % 4 == 0), walking (row, column) incrementally -- the first version spent a 64-bit division per element and
// reached 2.0 TB/s on a pure 172 MB write.
constexpr int kTalRows = 128;
template <int V>
__global__ void __launch_bounds__(256) tal_emit_kernel(const float* __restrict__ metric, const int* __restrict__ best, const int64_t* __restrict__ labels,
                                                       const uint8_t* __restrict__ fg, const int64_t* __restrict__ gt_idx, int64_t rows, int A, int nc,
                                                       int M, float eps, float* __restrict__ tscores) {
    __shared__ int s_lab[kTalRows];
    __shared__ float s_val[kTalRows];
    const int tid = threadIdx.x;
    const int nv = nc / V;                       // vectors per row
    const int dr = 256 / nv, dc = 256 % nv;      // (row, column) step of a thread between its consecutive vectors
    for (int64_t row0 = (int64_t)blockIdx.x * kTalRows; row0 < rows; row0 += (int64_t)gridDim.x * kTalRows) {
        const int n_rows = rows - row0 < kTalRows ? (int)(rows - row0) : kTalRows;
        if (tid < n_rows) {
            const int64_t ba = row0 + tid;
            int lab = -1;
            float val = 0.f;
            if (fg[ba]) {
                const int b = (int)(ba / A), a = (int)(ba - (int64_t)b * A);
                const int64_t g = (int64_t)b * M + gt_idx[ba];
                lab = (int)labels[ba];
                val = __fdiv_rn(__fmul_rn(metric[g * A + a], __int_as_float(best[2 * g + 1])), __fadd_rn(__int_as_float(best[2 * g]), eps));
            }
            s_lab[tid] = lab;
            s_val[tid] = val;
        }
        __syncthreads();
        float* dst = tscores + row0 * nc;
        int r = tid / nv, c = tid - r * nv;
        while (r < n_rows) {
            const int rel = s_lab[r] - c * V;    // position of the hot class inside this vector, if any
            if (V == 4) {
                float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
                const float v = s_val[r];
                if (rel == 0) o.x = v;
                if (rel == 1) o.y = v;
                if (rel == 2) o.z = v;
                if (rel == 3) o.w = v;
                reinterpret_cast<float4*>(dst + (int64_t)r * nc)[c] = o;
            } else {
                dst[(int64_t)r * nc + c] = rel == 0 ? s_val[r] : 0.f;
            }
            r += dr;
            c += dc;
            while (c >= nv) { c -= nv; ++r; }  // rows wider than the CTA (nc > 256 V) wrap more than once
        }
        __syncthreads();
    }
}

struct TalLayout { size_t metric, overlap, flags, best, total; };
static TalLayout tal_layout(int B, int M, int A) {
    const size_t n = (size_t)B * M * A;
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    TalLayout L;
    L.metric = 0;
    L.overlap = up(L.metric + n * 4);
    L.flags = up(L.overlap + n * 4);
    L.best = up(L.flags + n);
    L.total = up(L.best + (size_t)B * M * 2 * 4);
    return L;
}

}  // namespace el

using namespace el;

extern "C" int el_tal_workspace_bytes(int B, int M, int A, size_t* bytes) {
    if (!bytes || B <= 0 || M <= 0 || A <= 0) return EL_ERR_ARG;
    *bytes = tal_layout(B, M, A).total;
    return EL_OK;
}

extern "C" int el_tal_assign(const float* scores, const float* boxes, const float* anchors, const float* gt_labels, const float* gt_boxes,
                             const uint8_t* gt_valid, int B, int A, int nc, int M, int topk, float alpha, float beta, float eps, void* workspace,
                             size_t workspace_bytes, int64_t* labels, float* tboxes, float* tscores, uint8_t* fg, int64_t* gt_idx, void* stream) {
    if (!scores || !boxes || !anchors || !gt_labels || !gt_boxes || !gt_valid || !workspace || !labels || !tboxes || !tscores || !fg || !gt_idx)
        return EL_ERR_ARG;
    if (B <= 0 || A <= 0 || nc <= 0 || M <= 0 || topk <= 0) return EL_ERR_ARG;
    if (B > 65535 || nc > (1 << 20) || !aligned16(tboxes)) return EL_ERR_UNSUPPORTED;
    const TalLayout L = tal_layout(B, M, A);
    if (workspace_bytes < L.total) return EL_ERR_WORKSPACE;
    if (!aligned16(workspace)) return EL_ERR_ARG;
    char* ws = static_cast<char*>(workspace);
    float* metric = reinterpret_cast<float*>(ws + L.metric);
    float* overlap = reinterpret_cast<float*>(ws + L.overlap);
    uint8_t* flags = reinterpret_cast<uint8_t*>(ws + L.flags);
    int* best = reinterpret_cast<int*>(ws + L.best);
    cudaStream_t st = (cudaStream_t)stream;
    const int k = topk < A ? topk : A;
    const size_t smem = (size_t)A * sizeof(float);
    if (smem <= 200 * 1024) {
        if (smem > 48 * 1024) cudaFuncSetAttribute(tal_metric_topk_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        tal_metric_topk_kernel<true><<<dim3(M, B), kTalThreads, smem, st>>>(scores, boxes, anchors, gt_labels, gt_boxes, gt_valid, A, nc, M, k, alpha, beta,
                                                                            metric, overlap, flags, best);
    } else {
        tal_metric_topk_kernel<false><<<dim3(M, B), kTalThreads, 0, st>>>(scores, boxes, anchors, gt_labels, gt_boxes, gt_valid, A, nc, M, k, alpha, beta,
                                                                          metric, overlap, flags, best);
    }
    const int64_t n_ba = (int64_t)B * A;
    tal_resolve_kernel<<<(unsigned)ceil_div(n_ba, 256), 256, 0, st>>>(metric, overlap, flags, gt_labels, gt_boxes, B, A, nc, M, best, labels, tboxes, fg,
                                                                      gt_idx);
    const int64_t tiles = ceil_div(n_ba, kTalRows);
    const unsigned grid = (unsigned)(tiles < kSMs * 8 ? tiles : kSMs * 8);
    if (nc % 4 == 0 && aligned16(tscores))
        tal_emit_kernel<4><<<grid, 256, 0, st>>>(metric, best, labels, fg, gt_idx, n_ba, A, nc, M, eps, tscores);
    else
        tal_emit_kernel<1><<<grid, 256, 0, st>>>(metric, best, labels, fg, gt_idx, n_ba, A, nc, M, eps, tscores);
    note_launches(3);
    return check_launch();
}
