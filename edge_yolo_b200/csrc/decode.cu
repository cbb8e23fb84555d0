// Fused GFLv2 x UniHead decode: one kernel per forward over all levels.
// Replaces, for the eval path of GFLHeadv2_uniH (nn/modules/head.py:880-908):
//   GF2Detect._compute_quality_from_logits  head.py:227-243  (softmax16, top-4, mean, 20->64->1 MLP)
//   GF2Detect._inference_with_quality       head.py:301-345  (cat, split, sigmoid, clamp(q), cat)
//   DFL.forward                             block.py:87-90   (softmax16 . arange)
//   make_anchors / dist2bbox                tal.py:333-357
// The 16-bin softmax is computed once and shared by the DFL integral and the DGQP statistics
// (the reference computes it twice).  HBM-bound: reads (64+nc) logits, writes (4+nc) fp32 per anchor.
#include "el_common.cuh"

namespace el {

constexpr int kMaxLevels = 4;
constexpr int kTile = 64;        // anchors per CTA
constexpr int kRegMax = 16;
constexpr int kStat = 20;        // 4 sides x (top-4 + mean)
constexpr int kHidden = 64;

struct DecodeLevel {
    const void* box; Strides4 bs;
    const void* cls; Strides4 cs;
    const float *w1, *b1, *w2, *b2;
    int H, W, a_off, tile_off;
    float stride;
};
struct DecodeParams {
    DecodeLevel lv[kMaxLevels];
    int nl, nc, A;
};

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

template <typename T, bool BOX_CH_FAST, bool CLS_STAGE>
__global__ void __launch_bounds__(256) gfl_decode_kernel(const __grid_constant__ DecodeParams P, float* __restrict__ y, float* __restrict__ q_out) {
    extern __shared__ float s_cls[];  // [kTile][nc+1] when CLS_STAGE
    __shared__ float s_w1[kHidden * kStat], s_b1[kHidden], s_w2[kHidden], s_b2;
    __shared__ float s_stat[kTile][kStat + 1];
    __shared__ float s_dist[4][kTile];
    __shared__ float s_part[4][kTile];
    __shared__ float s_q[kTile];

    const int tid = threadIdx.x, b = blockIdx.y;
    int l = 0;
#pragma unroll
    for (int i = 1; i < kMaxLevels; ++i)
        if (i < P.nl && (int)blockIdx.x >= P.lv[i].tile_off) l = i;
    const DecodeLevel& L = P.lv[l];
    const int HW = L.H * L.W;
    const int pix0 = ((int)blockIdx.x - L.tile_off) * kTile;
    const int nc = P.nc;

    for (int i = tid; i < kHidden * kStat; i += 256) s_w1[i] = __ldg(L.w1 + i);
    if (tid < kHidden) { s_b1[tid] = __ldg(L.b1 + tid); s_w2[tid] = __ldg(L.w2 + tid); }
    if (tid == 0) s_b2 = __ldg(L.b2);

    // ---- phase 1: one thread per (anchor, side): softmax over 16 bins, integral, top-4, mean
    {
        const int a = tid & (kTile - 1), side = tid >> 6;
        const int pix = pix0 + a;
        float lg[kRegMax];
        if (pix < HW) {
            const int py = pix / L.W, px = pix - py * L.W;
            const T* p = reinterpret_cast<const T*>(L.box) + (int64_t)b * L.bs.n + (int64_t)py * L.bs.h + (int64_t)px * L.bs.w;
            if constexpr (BOX_CH_FAST && sizeof(T) == 2) {
                const T* pv = p + side * kRegMax;  // 16 contiguous 2-byte logits = 2 x 16 B
                float f0[8], f1[8];
                unpack<T>(ldg_stream(pv), f0);
                unpack<T>(ldg_stream(pv + 8), f1);
#pragma unroll
                for (int k = 0; k < 8; ++k) { lg[k] = f0[k]; lg[8 + k] = f1[k]; }
            } else {
#pragma unroll
                for (int k = 0; k < kRegMax; ++k) lg[k] = to_f(p[(int64_t)(side * kRegMax + k) * L.bs.c]);
            }
        } else {
#pragma unroll
            for (int k = 0; k < kRegMax; ++k) lg[k] = 0.f;
        }
        float m = lg[0];
#pragma unroll
        for (int k = 1; k < kRegMax; ++k) m = fmaxf(m, lg[k]);
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < kRegMax; ++k) { lg[k] = expf(lg[k] - m); s += lg[k]; }
        float dist = 0.f, psum = 0.f;
        float t0 = -1.f, t1 = -1.f, t2 = -1.f, t3 = -1.f;  // running top-4, descending
#pragma unroll
        for (int k = 0; k < kRegMax; ++k) {
            float p = lg[k] / s;
            dist += (float)k * p;
            psum += p;
            // insert p into (t0>=t1>=t2>=t3)
            float v = p;
            float n0 = fmaxf(t0, v); v = fminf(t0, v); t0 = n0;
            float n1 = fmaxf(t1, v); v = fminf(t1, v); t1 = n1;
            float n2 = fmaxf(t2, v); v = fminf(t2, v); t2 = n2;
            t3 = fmaxf(t3, v);
        }
        s_dist[side][a] = dist;
        float* st = &s_stat[a][side * 5];
        st[0] = t0; st[1] = t1; st[2] = t2; st[3] = t3;
        st[4] = psum * (1.f / kRegMax);  // prob.mean(dim=2): carries no information (== 1/16) but is reproduced
    }
    // ---- stage the class tile (channel-contiguous inputs): coalesced read, transposed use
    if (CLS_STAGE) {
        const int n_el = kTile * nc;
        for (int i = tid; i < n_el; i += 256) {
            int a = i / nc, c = i - a * nc;
            int pix = pix0 + a;
            float v = 0.f;
            if (pix < HW) {
                int py = pix / L.W, px = pix - py * L.W;
                v = to_f(reinterpret_cast<const T*>(L.cls)[(int64_t)b * L.cs.n + (int64_t)py * L.cs.h + (int64_t)px * L.cs.w + c]);
            }
            s_cls[a * (nc + 1) + c] = v;
        }
    }
    __syncthreads();

    // ---- phase 2: DGQP MLP 20 -> 64 (ReLU) -> 1 (sigmoid); 4 threads per anchor, 16 hidden units each
    {
        const int a = tid & (kTile - 1), og = tid >> 6;
        float st[kStat];
#pragma unroll
        for (int c = 0; c < kStat; ++c) st[c] = s_stat[a][c];
        float z = 0.f;
#pragma unroll 4
        for (int o = og * 16; o < og * 16 + 16; ++o) {
            float acc = s_b1[o];
#pragma unroll
            for (int c = 0; c < kStat; ++c) acc += s_w1[o * kStat + c] * st[c];  // w1 read is a warp broadcast
            z += s_w2[o] * fmaxf(acc, 0.f);
        }
        s_part[og][a] = z;
    }
    __syncthreads();
    if (tid < kTile) {
        const int a = tid, pix = pix0 + a;
        float z = s_b2 + ((s_part[0][a] + s_part[1][a]) + (s_part[2][a] + s_part[3][a]));
        float q = sigmoidf_(z);
        q = fminf(fmaxf(q, 1e-6f), 1.f - 1e-6f);
        s_q[a] = q;
        if (pix < HW) {
            const int py = pix / L.W, px = pix - py * L.W;
            const float ax = (float)px + 0.5f, ay = (float)py + 0.5f;
            const float x1 = ax - s_dist[0][a], y1 = ay - s_dist[1][a], x2 = ax + s_dist[2][a], y2 = ay + s_dist[3][a];
            float* o = y + (int64_t)b * (4 + nc) * P.A + L.a_off + pix;
            o[0] = ((x1 + x2) / 2.f) * L.stride;
            o[P.A] = ((y1 + y2) / 2.f) * L.stride;
            o[2 * (int64_t)P.A] = (x2 - x1) * L.stride;
            o[3 * (int64_t)P.A] = (y2 - y1) * L.stride;
            if (q_out) q_out[(int64_t)b * P.A + L.a_off + pix] = q;
        }
    }
    __syncthreads();

    // ---- phase 3: class scores, written channel-major (coalesced along anchors)
    {
        float* o = y + ((int64_t)b * (4 + nc) + 4) * P.A + L.a_off + pix0;
        const int n_el = kTile * nc;
        for (int i = tid; i < n_el; i += 256) {
            int c = i >> 6, a = i & (kTile - 1);
            int pix = pix0 + a;
            if (pix >= HW) continue;
            float v;
            if (CLS_STAGE) {
                v = s_cls[a * (nc + 1) + c];
            } else {
                int py = pix / L.W, px = pix - py * L.W;
                v = to_f(reinterpret_cast<const T*>(L.cls)[(int64_t)b * L.cs.n + (int64_t)c * L.cs.c + (int64_t)py * L.cs.h + (int64_t)px * L.cs.w]);
            }
            o[(int64_t)c * P.A + a] = sigmoidf_(v) * s_q[a];
        }
    }
}

}  // namespace el

using namespace el;

extern "C" int el_gfl_decode_fwd(int nl, const void* const* box, const int64_t* box_s, const void* const* cls, const int64_t* cls_s,
                                 const int32_t* hw, const float* stride, const float* const* w1, const float* const* b1, const float* const* w2,
                                 const float* const* b2, float* y, float* q_out, int B, int nc, int dtype, void* stream) {
    if (nl <= 0 || nl > kMaxLevels || !box || !cls || !box_s || !cls_s || !hw || !stride || !w1 || !b1 || !w2 || !b2 || !y || B <= 0 || nc <= 0)
        return EL_ERR_ARG;
    DecodeParams P;
    P.nl = nl; P.nc = nc;
    int a_off = 0, tile_off = 0;
    bool box_fast = true, cls_fast = true;
    for (int l = 0; l < nl; ++l) {
        DecodeLevel& L = P.lv[l];
        if (!box[l] || !cls[l] || !w1[l] || !b1[l] || !w2[l] || !b2[l] || hw[2 * l] <= 0 || hw[2 * l + 1] <= 0) return EL_ERR_ARG;
        L.box = box[l]; L.bs = s4(box_s + 4 * l);
        L.cls = cls[l]; L.cs = s4(cls_s + 4 * l);
        L.w1 = w1[l]; L.b1 = b1[l]; L.w2 = w2[l]; L.b2 = b2[l];
        L.H = hw[2 * l]; L.W = hw[2 * l + 1];
        L.stride = stride[l];
        L.a_off = a_off; L.tile_off = tile_off;
        a_off += L.H * L.W;
        tile_off += (int)ceil_div((int64_t)L.H * L.W, kTile);
        // 16 B vector reads of the 16 bins of a side need channel-contiguous, 16 B aligned views
        box_fast = box_fast && L.bs.c == 1 && aligned16(L.box) && L.bs.n % 8 == 0 && L.bs.h % 8 == 0 && L.bs.w % 8 == 0;
        cls_fast = cls_fast && L.cs.c == 1;
    }
    P.A = a_off;
    const size_t cls_smem = (size_t)kTile * (nc + 1) * sizeof(float);
    if (cls_smem > 160 * 1024) cls_fast = false;  // very wide heads: direct (uncoalesced) reads
    dim3 grid(tile_off, B);
    cudaStream_t st = (cudaStream_t)stream;
#define EL_LAUNCH_DECODE(BF, CS)                                                                                                  \
    do {                                                                                                                          \
        auto kern = gfl_decode_kernel<T, BF, CS>;                                                                                 \
        size_t sm = CS ? cls_smem : 0;                                                                                            \
        if (sm > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);                     \
        kern<<<grid, 256, sm, st>>>(P, y, q_out);                                                                                 \
    } while (0)
    EL_DISPATCH_DTYPE(dtype, {
        const bool bf = box_fast && sizeof(T) == 2;
        if (bf && cls_fast) EL_LAUNCH_DECODE(true, true);
        else if (bf) EL_LAUNCH_DECODE(true, false);
        else if (cls_fast) EL_LAUNCH_DECODE(false, true);
        else EL_LAUNCH_DECODE(false, false);
    });
#undef EL_LAUNCH_DECODE
    note_launches(1);
    return check_launch();
}
