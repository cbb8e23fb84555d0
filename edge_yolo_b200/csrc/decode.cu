// Fused GFLv2 x UniHead decode.  Replaces, for the eval path of GFLHeadv2_uniH (nn/modules/head.py:880-908):
//   GF2Detect._compute_quality_from_logits  head.py:227-243  (softmax16, top-4, mean, 20->64->1 MLP)
//   GF2Detect._inference_with_quality       head.py:301-345  (cat, split, sigmoid, clamp(q), cat)
//   DFL.forward                             block.py:87-90   (softmax16 . arange)
//   make_anchors / dist2bbox                tal.py:333-357
// The 16-bin softmax is computed once and shared by the DFL integral and the DGQP statistics
// (the reference computes it twice).
//
// Two kernels share every arithmetic helper (so they produce bit-identical boxes and scores):
//   gfl_decode_kernel        dense API path: writes y (B, 4+nc, A) fp32; any strides / layouts
//   gfl_decode_emit_kernel   engine path: persistent CTAs, 64-anchor tiles of the NHWC head outputs staged
//                            with 1-D bulk TMA copies (cp.async.bulk + mbarrier, double buffered), decode in
//                            shared memory, and instead of the dense score matrix only the xywh boxes
//                            (B, A, 4) and the NMS candidate keys are written (warp-ballot compaction)
//                            -- 45 % of the dense kernel's HBM traffic.
#include <climits>
#include <cstdlib>

#include "el_internal.h"

namespace el {

constexpr int kMaxLevels = 4;
constexpr int kTile = 64;        // anchors per tile
constexpr int kRegMax = 16;
constexpr int kStat = 20;        // 4 sides x (top-4 + mean)
constexpr int kHidden = 64;
// The DGQP hidden layer (64 anchors x 20 statistics -> 64 units per tile) has two implementations, selected by the activation type:
//   fp32 maps   : CUDA-core FMAs in IEEE fp32 (1e-5 contract against the fp32 reference)
//   16-bit maps : mma.sync m16n8k8 TF32 tensor-core tiles (K = 20 padded to 24), fp32 accumulation; the TF32 rounding of the statistics
//                 and weights (2^-11) sits two octaves below the 2^-9 rounding the 16-bit box logits already carry.  The FMA version
//                 was 1/3 of the fused kernel's issue slots (ncu: issue-bound at 4 warps per scheduler).
// Per level the weights live in shared memory as  w1 | b1 | w2 | b2 (+pad):
//   FMA: w1 row-major (64 x 20);  MMA: w1 as TF32 B fragments [n-tile 8][k-step 3][lane 32][2] (one 8-byte load per lane and MMA).
template <bool MMA> struct WL {
    static constexpr int w1n = MMA ? 8 * 3 * 32 * 2 : kHidden * kStat;
    static constexpr int b1 = w1n, w2 = b1 + kHidden, b2 = w2 + kHidden, total = b2 + 4;  // multiple of 4 floats: 16 B aligned levels
    static constexpr int stat_ld = MMA ? 28 : kStat + 1;  // 28 = 4 x odd: conflict-free A-fragment loads; 21: conflict-free row reads
};

struct DecodeLevel {
    const void* box; Strides4 bs;
    const void* cls; Strides4 cs;
    const float *w1, *b1, *w2, *b2;
    const float *bb, *cb;  // optional biases of the final 1x1 convs of the box / class towers (folded in here), may be NULL
    int H, W, a_off, tile_off;
    float stride;
};
struct DecodeParams {
    DecodeLevel lv[kMaxLevels];
    int nl, nc, A, tiles_per_image;
};

// Transcendentals: fp32 maps keep IEEE expf / division (1e-5 contract against the fp32 reference); 16-bit maps use the SFU
// intrinsics (relative error ~1e-6, two orders below the bf16 / fp16 input rounding).  The policy is a template parameter of
// every helper so the dense and the fused kernel stay bit-identical for a given dtype.
// FAST: exp = FMUL + MUFU.EX2 (ex2.approx.ftz), reciprocal = MUFU.RCP (rcp.approx.ftz): a sigmoid is 4 instructions instead of the ~13 of
// __expf (denormal-range fix-up) + __frcp_rn (correctly rounded).
__device__ __forceinline__ float ex2_fast(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float rcp_fast(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
template <bool FAST> __device__ __forceinline__ float exp_(float x) { return FAST ? ex2_fast(x * 1.4426950408889634f) : expf(x); }
template <bool FAST> __device__ __forceinline__ float rcp_(float x) { return FAST ? rcp_fast(x) : 1.f / x; }
template <bool FAST> __device__ __forceinline__ float sigmoidf_(float x) { return rcp_<FAST>(1.f + exp_<FAST>(-x)); }

// ---- shared arithmetic ------------------------------------------------------------------------
// softmax over the 16 bins of one side: DFL integral, sorted top-4 probabilities and their mean
__device__ __forceinline__ float to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__device__ __forceinline__ float ex2_approx(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// Sorted (descending) top-4 of 16 values by a merge network: four 4-sorters (5 compare-exchanges each), then three "top half of a
// bitonic merge" steps (c_i = max(a_i, b_{3-i}) keeps the four largest of two sorted quads as a bitonic sequence, two more exchange
// levels sort it): 76 min / max instead of the 112 of a running insertion -- the insertion was 14 % of the fused kernel's instructions.
// min / max are exact, so the result is the same multiset the insertion produced, whatever the input order.
__device__ __forceinline__ void cex(float& a, float& b) {  // a >= b afterwards
    const float hi = fmaxf(a, b), lo = fminf(a, b);
    a = hi; b = lo;
}
__device__ __forceinline__ void sort4_desc(float& a, float& b, float& c, float& d) {
    cex(a, b); cex(c, d); cex(a, c); cex(b, d); cex(b, c);
}
__device__ __forceinline__ void merge_top4(float& a0, float& a1, float& a2, float& a3, float b0, float b1, float b2, float b3) {
    a0 = fmaxf(a0, b3); a1 = fmaxf(a1, b2); a2 = fmaxf(a2, b1); a3 = fmaxf(a3, b0);
    cex(a0, a2); cex(a1, a3); cex(a0, a1); cex(a2, a3);
}
__device__ __forceinline__ void top4_of_16(float (&v)[kRegMax], float& t0, float& t1, float& t2, float& t3) {
    sort4_desc(v[0], v[1], v[2], v[3]);
    sort4_desc(v[4], v[5], v[6], v[7]);
    sort4_desc(v[8], v[9], v[10], v[11]);
    sort4_desc(v[12], v[13], v[14], v[15]);
    merge_top4(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]);
    merge_top4(v[8], v[9], v[10], v[11], v[12], v[13], v[14], v[15]);
    merge_top4(v[0], v[1], v[2], v[3], v[8], v[9], v[10], v[11]);
    t0 = v[0]; t1 = v[1]; t2 = v[2]; t3 = v[3];
}

// `swapped` (16-bit maps of the fused kernel): lg[0..7] holds bins 8..15 and lg[8..15] bins 0..7 -- the two 16-byte halves of a side are
// read in alternating order to keep the shared-memory reads conflict-free.  The sums run per half in bin order and are combined as
// (bins 0-7) + (bins 8-15), so the result does not depend on which half came first: the dense kernel (swapped = false) and the fused
// kernel stay bit-identical.
template <bool FAST>
__device__ __forceinline__ void side_stats(float (&lg)[kRegMax], float& dist, float* __restrict__ stat5, bool swapped = false) {
    float m = lg[0];
#pragma unroll
    for (int k = 1; k < kRegMax; ++k) m = fmaxf(m, lg[k]);
    float t0, t1, t2, t3;  // top-4, descending
    if constexpr (FAST) {
        // 16-bit maps: exp as one FFMA + MUFU.EX2 per bin; the integral, the top-4 and the mean are taken on the unnormalised
        // exponentials and scaled once (same values up to fp32 rounding, ~45 fewer instructions per side)
        constexpr float kLog2e = 1.4426950408889634f;
        const float ml = m * kLog2e;
        float sa = 0.f, sb = 0.f, da = 0.f, db = 0.f;  // per loaded half: sum and sum of (bin mod 8) * e
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            lg[k] = ex2_approx(fmaf(lg[k], kLog2e, -ml));
            lg[8 + k] = ex2_approx(fmaf(lg[8 + k], kLog2e, -ml));
            sa += lg[k]; sb += lg[8 + k];
            da = fmaf((float)k, lg[k], da); db = fmaf((float)k, lg[8 + k], db);
        }
        const float s_lo = swapped ? sb : sa, s_hi = swapped ? sa : sb;
        const float d_lo = swapped ? db : da, d_hi = swapped ? da : db;
        const float s = s_lo + s_hi;
        const float d = fmaf(8.f, s_hi, d_lo + d_hi);
        top4_of_16(lg, t0, t1, t2, t3);
        const float inv = rcp_fast(s);
        dist = d * inv;
        stat5[0] = to_tf32(t0 * inv); stat5[1] = to_tf32(t1 * inv); stat5[2] = to_tf32(t2 * inv); stat5[3] = to_tf32(t3 * inv);
        stat5[4] = to_tf32((s * inv) * (1.f / kRegMax));
    } else {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < kRegMax; ++k) { lg[k] = exp_<FAST>(lg[k] - m); s += lg[k]; }
        const float inv = rcp_<FAST>(s);
        float d = 0.f, psum = 0.f;
#pragma unroll
        for (int k = 0; k < kRegMax; ++k) {
            lg[k] = lg[k] * inv;
            d += (float)k * lg[k];
            psum += lg[k];
        }
        top4_of_16(lg, t0, t1, t2, t3);
        dist = d;
        stat5[0] = t0; stat5[1] = t1; stat5[2] = t2; stat5[3] = t3;
        stat5[4] = psum * (1.f / kRegMax);  // prob.mean(dim=2): == 1/16, carries no information but is reproduced
    }
}

// DGQP hidden layer for one 64-anchor tile with 256 threads, FMA version: thread = (anchor pair ap, ap+32 ; 8 hidden units).
// partial z sums go to s_part[8][kTile] (the fused kernel's shared-memory map reserves 8 rows for this version, 2 for the MMA version).
__device__ __forceinline__ void dgqp_hidden_fma(const float* __restrict__ w, const float* __restrict__ s_stat, float (*s_part)[kTile]) {
    constexpr int LD = WL<false>::stat_ld;
    const int ap = threadIdx.x & 31, hg = threadIdx.x >> 5;
    float st0[kStat], st1[kStat];
#pragma unroll
    for (int c = 0; c < kStat; ++c) { st0[c] = s_stat[ap * LD + c]; st1[c] = s_stat[(ap + 32) * LD + c]; }
    const float* b1 = w + WL<false>::b1;
    const float* w2 = w + WL<false>::w2;
    float z0 = 0.f, z1 = 0.f;
#pragma unroll 2
    for (int o = hg * 8; o < hg * 8 + 8; ++o) {
        float a0 = b1[o], a1 = a0;
        const float4* wr = reinterpret_cast<const float4*>(w + o * kStat);  // warp-uniform address: broadcast
#pragma unroll
        for (int c4 = 0; c4 < kStat / 4; ++c4) {
            const float4 wv = wr[c4];
            a0 += wv.x * st0[4 * c4]; a1 += wv.x * st1[4 * c4];
            a0 += wv.y * st0[4 * c4 + 1]; a1 += wv.y * st1[4 * c4 + 1];
            a0 += wv.z * st0[4 * c4 + 2]; a1 += wv.z * st1[4 * c4 + 2];
            a0 += wv.w * st0[4 * c4 + 3]; a1 += wv.w * st1[4 * c4 + 3];
        }
        const float wo = w2[o];
        z0 += wo * fmaxf(a0, 0.f);
        z1 += wo * fmaxf(a1, 0.f);
    }
    s_part[hg][ap] = z0;
    s_part[hg][ap + 32] = z1;
}

// D (16x8, fp32) += A (16x8, tf32, row) * B (8x8, tf32, col)
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// Tensor-core version: warp = (16-anchor m-tile mt, half nh of the hidden units = 4 n-tiles of 8); 12 MMAs per warp.
// Fragment coordinates (g = lane / 4, t = lane % 4): A a0 (g, t) a1 (g+8, t) a2 (g, t+4) a3 (g+8, t+4); B b0 (k = t, n = g) b1 (k = t+4, n = g);
// D d0 (g, 2t) d1 (g, 2t+1) d2 (g+8, 2t) d3 (g+8, 2t+1).  Epilogue in registers: z += w2[o] * relu(h[o] + b1[o]), reduced over the
// four lanes of a group; the two halves meet in s_part[0..1][anchor].
__device__ __forceinline__ void dgqp_hidden_mma(const float* __restrict__ w, const float* __restrict__ s_stat, float (*s_part)[kTile]) {
    constexpr int LD = WL<true>::stat_ld;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3, mt = warp & 3, nh = warp >> 2;
    const uint32_t* S = reinterpret_cast<const uint32_t*>(s_stat) + (16 * mt + g) * LD + t;
    uint32_t a[3][4];
#pragma unroll
    for (int ks = 0; ks < 3; ++ks) {
        a[ks][0] = S[8 * ks]; a[ks][1] = S[8 * LD + 8 * ks]; a[ks][2] = S[8 * ks + 4]; a[ks][3] = S[8 * LD + 8 * ks + 4];
    }
    const uint2* Wf = reinterpret_cast<const uint2*>(w) + lane;
    const float2* b1 = reinterpret_cast<const float2*>(w + WL<true>::b1);
    const float2* w2 = reinterpret_cast<const float2*>(w + WL<true>::w2);
    float z0 = 0.f, z1 = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int nt = 4 * nh + j;
        float d[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int ks = 0; ks < 3; ++ks) {
            const uint2 b = Wf[(nt * 3 + ks) * 32];
            mma_tf32(d, a[ks], b.x, b.y);
        }
        const float2 bb = b1[4 * nt + t], ww = w2[4 * nt + t];  // hidden units 8 nt + 2 t, + 1
        z0 = fmaf(ww.x, fmaxf(d[0] + bb.x, 0.f), z0); z0 = fmaf(ww.y, fmaxf(d[1] + bb.y, 0.f), z0);
        z1 = fmaf(ww.x, fmaxf(d[2] + bb.x, 0.f), z1); z1 = fmaf(ww.y, fmaxf(d[3] + bb.y, 0.f), z1);
    }
    z0 += __shfl_xor_sync(0xffffffffu, z0, 1); z0 += __shfl_xor_sync(0xffffffffu, z0, 2);
    z1 += __shfl_xor_sync(0xffffffffu, z1, 1); z1 += __shfl_xor_sync(0xffffffffu, z1, 2);
    if (t == 0) { s_part[nh][16 * mt + g] = z0; s_part[nh][16 * mt + g + 8] = z1; }
}

template <bool MMA>
__device__ __forceinline__ void dgqp_hidden(const float* __restrict__ w, const float* __restrict__ s_stat, float (*s_part)[kTile]) {
    if constexpr (MMA) dgqp_hidden_mma(w, s_stat, s_part);
    else dgqp_hidden_fma(w, s_stat, s_part);
}

template <bool FAST>
__device__ __forceinline__ float dgqp_quality(const float* __restrict__ w, const float (*s_part)[kTile], int a) {
    float z = w[WL<FAST>::b2];
    if constexpr (FAST) z += s_part[0][a] + s_part[1][a];
    else z += ((s_part[0][a] + s_part[1][a]) + (s_part[2][a] + s_part[3][a])) + ((s_part[4][a] + s_part[5][a]) + (s_part[6][a] + s_part[7][a]));
    const float q = sigmoidf_<FAST>(z);
    return fminf(fmaxf(q, 1e-6f), 1.f - 1e-6f);  // clamp(1e-6, 1 - 1e-6), head.py:343
}

// dist2bbox (xywh) * stride, tal.py:348-357 + head.py:341
__device__ __forceinline__ float4 decode_box(int px, int py, float l, float t, float r, float b, float stride) {
    const float ax = (float)px + 0.5f, ay = (float)py + 0.5f;
    const float x1 = ax - l, y1 = ay - t, x2 = ax + r, y2 = ay + b;
    return make_float4(((x1 + x2) / 2.f) * stride, ((y1 + y2) / 2.f) * stride, (x2 - x1) * stride, (y2 - y1) * stride);
}

template <bool MMA>
__device__ __forceinline__ void load_level_weights(float* __restrict__ dst, const DecodeLevel& L) {
    if constexpr (MMA) {
        // fragment slot of w1[n][k]: ((nt * 3 + ks) * 32 + lane) * 2 + j with nt = n / 8, ks = k / 8, lane = (n % 8) * 4 + k % 4, j = (k % 8) / 4.
        // One task = (hidden unit n, four consecutive k): a 16-byte row read (k = 20 .. 23: the zero padding) and four slots two floats apart --
        // the per-element version (a division and a scalar __ldg per slot) was ~4 % of the fused kernel's instructions and most of the dense kernel's prologue.
        const bool vec = (reinterpret_cast<uintptr_t>(L.w1) & 15) == 0;
        for (int task = threadIdx.x; task < kHidden * 6; task += blockDim.x) {
            const int n = task / 6, q = task - 6 * n;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (q < 5) {
                const float* src = L.w1 + n * kStat + 4 * q;
                if (vec) v = __ldg(reinterpret_cast<const float4*>(src));
                else v = make_float4(__ldg(src), __ldg(src + 1), __ldg(src + 2), __ldg(src + 3));
            }
            const int nt = n >> 3, gq = n & 7, ks = q >> 1, j = q & 1;
            float* d = dst + ((nt * 3 + ks) * 32 + gq * 4) * 2 + j;
            d[0] = to_tf32(v.x); d[2] = to_tf32(v.y); d[4] = to_tf32(v.z); d[6] = to_tf32(v.w);
        }
    } else {
        for (int i = threadIdx.x; i < kHidden * kStat; i += blockDim.x) dst[i] = __ldg(L.w1 + i);
    }
    if (threadIdx.x < kHidden) {
        dst[WL<MMA>::b1 + threadIdx.x] = __ldg(L.b1 + threadIdx.x);
        dst[WL<MMA>::w2 + threadIdx.x] = __ldg(L.w2 + threadIdx.x);
    }
    if (threadIdx.x == 0) dst[WL<MMA>::b2] = __ldg(L.b2);
}

// ---------------------------------------------------------------------------------- dense kernel
template <typename T, bool BOX_CH_FAST, bool CLS_STAGE>
__global__ void __launch_bounds__(256) gfl_decode_kernel(const __grid_constant__ DecodeParams P, float* __restrict__ y, float* __restrict__ q_out) {
    constexpr bool kFast = sizeof(T) == 2;  // 16-bit maps: SFU transcendentals + tensor-core DGQP layer
    constexpr int LD = WL<kFast>::stat_ld;
    extern __shared__ float s_cls[];  // [kTile][nc+1] when CLS_STAGE
    __shared__ __align__(16) float s_w[WL<kFast>::total];
    __shared__ __align__(16) float s_stat[kTile * LD];
    __shared__ float s_dist[4][kTile];
    __shared__ float s_part[8][kTile];
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
    load_level_weights<kFast>(s_w, L);

    {   // phase 1: one thread per (anchor, side)
        const int side = tid & 3, a = tid >> 2;
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
            if (L.bb) {
#pragma unroll
                for (int k = 0; k < kRegMax; ++k) lg[k] += __ldg(L.bb + side * kRegMax + k);
            }
        } else {
#pragma unroll
            for (int k = 0; k < kRegMax; ++k) lg[k] = 0.f;
        }
        float dist;
        side_stats<kFast>(lg, dist, &s_stat[a * LD + side * 5]);
        s_dist[side][a] = dist;
        if (kFast && side == 0) *reinterpret_cast<float4*>(&s_stat[a * LD + kStat]) = make_float4(0.f, 0.f, 0.f, 0.f);  // K padding 20 -> 24
    }
    if (CLS_STAGE) {  // channel-contiguous class maps: coalesced read now, transposed use in phase 3
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
    dgqp_hidden<kFast>(s_w, s_stat, s_part);
    __syncthreads();
    if (tid < kTile) {
        const int a = tid, pix = pix0 + a;
        const float q = dgqp_quality<kFast>(s_w, s_part, a);
        s_q[a] = q;
        if (pix < HW) {
            const int py = pix / L.W, px = pix - py * L.W;
            const float4 bx = decode_box(px, py, s_dist[0][a], s_dist[1][a], s_dist[2][a], s_dist[3][a], L.stride);
            float* o = y + (int64_t)b * (4 + nc) * P.A + L.a_off + pix;
            o[0] = bx.x; o[P.A] = bx.y; o[2 * (int64_t)P.A] = bx.z; o[3 * (int64_t)P.A] = bx.w;
            if (q_out) q_out[(int64_t)b * P.A + L.a_off + pix] = q;
        }
    }
    __syncthreads();
    {   // phase 3: class scores, written channel-major (coalesced along anchors)
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
            if (L.cb) v += __ldg(L.cb + c);
            o[(int64_t)c * P.A + a] = sigmoidf_<sizeof(T) == 2>(v) * s_q[a];
        }
    }
}

// ------------------------------------------------------------------------------- fused emit kernel
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 1-D bulk TMA copy global -> shared, completion counted in bytes on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
                 "r"(smem_u32(bar))
                 : "memory");
}

struct EmitArgs {
    float conf;
    const int32_t* class_keep;
    unsigned long long* keys; int64_t key_stride;
    int* counts;
    float4* boxes;  // (B, A) xywh * stride
    int B;
};

struct TileInfo { int b, l, pix0, nvalid; };
__device__ __forceinline__ TileInfo tile_info(const DecodeParams& P, int t) {
    TileInfo ti;
    ti.b = t / P.tiles_per_image;
    const int r = t - ti.b * P.tiles_per_image;
    ti.l = 0;
#pragma unroll
    for (int i = 1; i < kMaxLevels; ++i)
        if (i < P.nl && r >= P.lv[i].tile_off) ti.l = i;
    ti.pix0 = (r - P.lv[ti.l].tile_off) * kTile;
    const int HW = P.lv[ti.l].H * P.lv[ti.l].W;
    ti.nvalid = min(kTile, HW - ti.pix0);
    return ti;
}

constexpr int kStageCap = 512;   // staged candidate keys per tile (4 KiB per stage; a single-label tile emits <= 64); overflow goes straight to global memory

// shared-memory map of the fused kernel (byte offsets), one definition for the host (size) and the device (carve-up)
struct EmitLayout {
    uint32_t box[2], cls[2], w, stat, dist, part, bias, bar, stage[2], ctl, total;
    int bias_ld, w_stride;
};
template <bool MMA>
__host__ __device__ inline EmitLayout emit_layout(int nc, uint32_t esz, int nl) {
    EmitLayout L;
    const uint32_t box_bytes = kTile * 4 * kRegMax * esz;
    const uint32_t cls_bytes = (uint32_t)((kTile * nc * esz + 127) & ~127u);
    uint32_t cur = 0;
    L.box[0] = cur; cur += box_bytes; L.box[1] = cur; cur += box_bytes;
    L.cls[0] = cur; cur += cls_bytes; L.cls[1] = cur; cur += cls_bytes;
    L.w_stride = WL<MMA>::total;
    L.w = cur; cur += (uint32_t)nl * L.w_stride * 4;
    L.stat = cur; cur += kTile * WL<MMA>::stat_ld * 4;
    L.dist = cur; cur += 4 * kTile * 4;
    L.part = cur; cur += (MMA ? 2 : 8) * kTile * 4;  // partial z sums: two hidden halves (MMA) or eight hidden groups (FMA)
    L.bias_ld = (4 * kRegMax + nc + 3) & ~3;
    L.bias = cur; cur += (uint32_t)nl * L.bias_ld * 4;  // per level: box bias (64) | class bias (nc)
    L.bar = cur; cur += 16;
    L.stage[0] = cur; cur += kStageCap * 8; L.stage[1] = cur; cur += kStageCap * 8;
    L.ctl = cur; cur += 96;  // ten ints of staging state, then (at +48) the TileInfo of the tile staged in each buffer
    L.total = cur;
    return L;
}

template <typename T> __device__ __forceinline__ void unpack2(uint32_t w, float& lo, float& hi);
template <> __device__ __forceinline__ void unpack2<__nv_bfloat16>(uint32_t w, float& lo, float& hi) {
    lo = __uint_as_float(w << 16); hi = __uint_as_float(w & 0xffff0000u);
}
template <> __device__ __forceinline__ void unpack2<__half>(uint32_t w, float& lo, float& hi) {
    const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&w));
    lo = t.x; hi = t.y;
}
template <> __device__ __forceinline__ void unpack2<float>(uint32_t w, float& lo, float& hi) { lo = hi = __uint_as_float(w); }  // never used

template <typename T, bool MULTI>
__global__ void __launch_bounds__(256, sizeof(T) == 2 ? 3 : 2) gfl_decode_emit_kernel(const __grid_constant__ DecodeParams P, const __grid_constant__ EmitArgs E) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr bool kFast = sizeof(T) == 2;
    constexpr int LD = WL<kFast>::stat_ld;
    const int nc = P.nc, tid = threadIdx.x, lane = tid & 31;
    const EmitLayout ML = emit_layout<kFast>(nc, sizeof(T), P.nl);
    // stage-indexed buffers as arithmetic on the stage bit (pointer arrays indexed at run time would live in local memory)
    auto s_box = [&](int st) { return (T*)(smem_raw + ML.box[0] + st * (ML.box[1] - ML.box[0])); };
    auto s_clsT = [&](int st) { return (T*)(smem_raw + ML.cls[0] + st * (ML.cls[1] - ML.cls[0])); };
    float* s_w = (float*)(smem_raw + ML.w);
    float* s_stat = (float*)(smem_raw + ML.stat);
    float (*s_dist)[kTile] = (float (*)[kTile])(smem_raw + ML.dist);
    float (*s_part)[kTile] = (float (*)[kTile])(smem_raw + ML.part);
    float* s_bias = (float*)(smem_raw + ML.bias);
    const int bias_ld = ML.bias_ld;
    uint64_t* bar = (uint64_t*)(smem_raw + ML.bar);
    // candidate keys are staged per tile in shared memory and flushed to the image's key list one tile later, so the global
    // atomic that reserves the slots (one per tile instead of one per warp) has a whole tile of work to hide behind
    auto s_stage = [&](int st) { return (unsigned long long*)(smem_raw + ML.stage[0] + st * (kStageCap * 8)); };
    int* s_scnt = (int*)(smem_raw + ML.ctl);  // [2] staged count (may overshoot the capacity)
    int* s_slimit = s_scnt + 2;   // [2] first position that did not fit (INT_MAX if none)
    int* s_fcnt = s_scnt + 4;     // [2] number of staged keys to flush
    int* s_fbase = s_scnt + 6;    // [2] reserved base slot in the image's key list
    int* s_fimg = s_scnt + 8;     // [2] image of the staged tile
    // tile coordinates of the tile in flight in each stage, written by the issuing thread before it arms the barrier (its arrive releases,
    // the waiters' try_wait acquires): the other 255 threads read 16 bytes instead of redoing the division and the level search per tile
    TileInfo* s_tinfo = reinterpret_cast<TileInfo*>(smem_raw + ML.ctl + 48);
    if (tid < 2) { s_scnt[tid] = 0; s_slimit[tid] = INT_MAX; s_fcnt[tid] = 0; }

    for (int l = 0; l < P.nl; ++l) {
        load_level_weights<kFast>(s_w + l * ML.w_stride, P.lv[l]);
        for (int i = tid; i < 4 * kRegMax + nc; i += 256)
            s_bias[l * bias_ld + i] = i < 4 * kRegMax ? (P.lv[l].bb ? __ldg(P.lv[l].bb + i) : 0.f) : (P.lv[l].cb ? __ldg(P.lv[l].cb + i - 4 * kRegMax) : 0.f);
    }
    if (kFast && tid < kTile) *reinterpret_cast<float4*>(&s_stat[tid * LD + kStat]) = make_float4(0.f, 0.f, 0.f, 0.f);  // K padding 20 -> 24, written once
    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int total = E.B * P.tiles_per_image;
    auto issue = [&](int t, int stage) {  // one thread: arm the barrier with the byte count, then two bulk copies
        const TileInfo ti = tile_info(P, t);
        s_tinfo[stage] = ti;
        const DecodeLevel& L = P.lv[ti.l];
        const uint32_t nb = (uint32_t)ti.nvalid * 4 * kRegMax * sizeof(T), ncb = (uint32_t)ti.nvalid * nc * sizeof(T);
        const T* gb = reinterpret_cast<const T*>(L.box) + (int64_t)ti.b * L.bs.n + (int64_t)ti.pix0 * (4 * kRegMax);
        const T* gc = reinterpret_cast<const T*>(L.cls) + (int64_t)ti.b * L.cs.n + (int64_t)ti.pix0 * nc;
        mbar_expect_tx(&bar[stage], nb + ncb);
        bulk_g2s(s_box(stage), gb, nb, &bar[stage]);
        bulk_g2s(s_clsT(stage), gc, ncb, &bar[stage]);
    };
    if (tid == 0) {
        if ((int)blockIdx.x < total) issue(blockIdx.x, 0);
        if ((int)(blockIdx.x + gridDim.x) < total) issue(blockIdx.x + gridDim.x, 1);
    }

    const int cq = (nc + 3) >> 2;  // classes per quarter-thread in phase 3
    int it = 0;
    int pending_base = 0;  // thread 0: slot reserved (atomicAdd issued at the end of the previous tile, consumed one tile later)
    // append one warp's candidates (ballot-compacted) to the tile's staging buffer; spill to global memory when it is full
    auto append = [&](bool pass, unsigned long long key, int stage, int img) {
        const unsigned m = __ballot_sync(0xffffffffu, pass);
        if (!m) return;
        const int nn = __popc(m), rank = __popc(m & ((1u << lane) - 1));
        int pos = 0;
        if (lane == 0) pos = atomicAdd(&s_scnt[stage], nn);
        pos = __shfl_sync(0xffffffffu, pos, 0);
        if (pos + nn <= kStageCap) {
            if (pass) s_stage(stage)[pos + rank] = key;
        } else {
            int gb = 0;
            if (lane == 0) { atomicMin(&s_slimit[stage], pos); gb = atomicAdd(E.counts + img, nn); }
            gb = __shfl_sync(0xffffffffu, gb, 0);
            if (pass) E.keys[(int64_t)img * E.key_stride + gb + rank] = key;
        }
    };
    for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
        const int stage = it & 1;
        mbar_wait(&bar[stage], (it >> 1) & 1);
        const TileInfo ti = s_tinfo[stage];
        const DecodeLevel& L = P.lv[ti.l];
        const float* w = s_w + ti.l * ML.w_stride;
        const float* bias_b = s_bias + ti.l * bias_ld;
        const float* bias_c = bias_b + 4 * kRegMax;

        {   // phase 1: thread = (anchor a, side): 16 logits from the staged tile
            const int side = tid & 3, a = tid >> 2;
            const int hswap = sizeof(T) == 2 ? (lane >> 2) & 1 : 0;  // 1: this thread holds bins 8..15 first
            float lg[kRegMax];
            const T* pv = s_box(stage) + a * (4 * kRegMax) + side * kRegMax;
            if constexpr (sizeof(T) == 2) {
                // alternate the half read first: conflict-free 16 B shared loads; lg keeps the LOADED order (side_stats un-swaps the sums)
                float f0[8], f1[8];
                unpack<T>(*reinterpret_cast<const uint4*>(pv + 8 * hswap), f0);
                unpack<T>(*reinterpret_cast<const uint4*>(pv + 8 * (hswap ^ 1)), f1);
#pragma unroll
                for (int k = 0; k < 8; ++k) { lg[k] = f0[k]; lg[8 + k] = f1[k]; }
            } else {
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    float f[4];
                    unpack<T>(*reinterpret_cast<const uint4*>(pv + 4 * v), f);
#pragma unroll
                    for (int k = 0; k < 4; ++k) lg[4 * v + k] = f[k];
                }
            }
            // rows past nvalid hold a previous tile's (finite) logits: their results are never stored and MMA rows are independent
            const float4* bb4 = reinterpret_cast<const float4*>(bias_b + side * kRegMax);
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                const float4 bv = bb4[v ^ (2 * hswap)];  // the bias of the bins this half really holds
                lg[4 * v] += bv.x; lg[4 * v + 1] += bv.y; lg[4 * v + 2] += bv.z; lg[4 * v + 3] += bv.w;
            }
            if (a >= ti.nvalid) {
#pragma unroll
                for (int k = 0; k < kRegMax; ++k) lg[k] = 0.f;
            }
            float dist;
            side_stats<kFast>(lg, dist, &s_stat[a * LD + side * 5], hswap != 0);
            s_dist[side][a] = dist;
        }
        __syncthreads();
        if (tid == 0 && it > 0) s_fbase[stage ^ 1] = pending_base;  // the previous tile's atomic has had a whole phase to return
        dgqp_hidden<kFast>(w, s_stat, s_part);
        __syncthreads();
        if (it > 0) {  // flush the previous tile's staged keys (coalesced 8-byte stores)
            const int ps = stage ^ 1, fc = s_fcnt[ps];
            unsigned long long* dst = E.keys + (int64_t)s_fimg[ps] * E.key_stride + s_fbase[ps];
            for (int i = tid; i < fc; i += 256) dst[i] = s_stage(ps)[i];
        }
        {   // phase 3: thread = (anchor a, quarter qd of the classes); scores never leave the SM unless they are candidates
            const int a = tid >> 2, qd = tid & 3;
            const bool av = a < ti.nvalid;
            const float q = dgqp_quality<kFast>(w, s_part, a);  // 4 lanes per anchor evaluate it redundantly: no extra barrier
            if (av && qd == 0) {
                const int pix = ti.pix0 + a, py = pix / L.W, px = pix - py * L.W;
                E.boxes[(int64_t)ti.b * P.A + L.a_off + pix] = decode_box(px, py, s_dist[0][a], s_dist[1][a], s_dist[2][a], s_dist[3][a], L.stride);
            }
            const T* pc = s_clsT(stage) + a * nc;
            const uint32_t anchor = (uint32_t)(L.a_off + ti.pix0 + a);
            if (MULTI) {
                for (int j = 0; j < cq; ++j) {
                    const int c = qd * cq + j;
                    float sc = 0.f;
                    bool pass = false;
                    if (av && c < nc) {
                        sc = sigmoidf_<kFast>(to_f(pc[c]) + bias_c[c]) * q;
                        pass = sc > E.conf && (!E.class_keep || E.class_keep[c]);
                    }
                    const uint32_t idx = anchor * (uint32_t)nc + (uint32_t)c;
                    append(pass, ((unsigned long long)__float_as_uint(sc) << 32) | (uint32_t)(~idx), stage, ti.b);
                }
            } else {
                float best = -INFINITY;
                int bc = 0;
                if (kFast && nc == 80) {
                    // COCO-sized head, 16-bit maps: sigma(x) * q is monotone in the logit, so the class maximum is found on the logits
                    // (one max per class instead of a sigmoid) and only logits within 1/64 of the maximum are scored -- the window keeps the
                    // result identical to scoring every class even if the SFU exp is not monotone to the last ulp (a 1/64 step moves the
                    // score by >= 87 ulp for logits <= 8; above 8 every class is scored).  Anchors whose best possible score is clearly
                    // below conf skip the scoring altogether.
                    constexpr int CQ = 20;
                    float xs[CQ];
                    const uint2* pcv = reinterpret_cast<const uint2*>(pc + qd * CQ);       // (160 a + 40 qd) bytes: 8 B aligned
                    const float4* bv4 = reinterpret_cast<const float4*>(bias_c + qd * CQ);  // 16 B aligned
#pragma unroll
                    for (int v = 0; v < CQ / 4; ++v) {
                        const uint2 r = pcv[v];
                        const float4 bb = bv4[v];
                        unpack2<T>(r.x, xs[4 * v], xs[4 * v + 1]);
                        unpack2<T>(r.y, xs[4 * v + 2], xs[4 * v + 3]);
                        xs[4 * v] += bb.x; xs[4 * v + 1] += bb.y; xs[4 * v + 2] += bb.z; xs[4 * v + 3] += bb.w;
                    }
                    float mx = xs[0];
#pragma unroll
                    for (int j = 1; j < CQ; ++j) mx = fmaxf(mx, xs[j]);
                    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
                    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
                    const float bound = sigmoidf_<kFast>(mx) * q;
                    if (av && bound * 1.0001f > E.conf) {
                        const float thr = mx > 8.f ? -INFINITY : mx - 0.015625f;
#pragma unroll
                        for (int j = 0; j < CQ; ++j) {
                            if (xs[j] >= thr) {
                                const float sc = sigmoidf_<kFast>(xs[j]) * q;
                                if (sc > best) { best = sc; bc = qd * CQ + j; }
                            }
                        }
                    }
                } else if (av) {
                    const int c1 = min(nc, (qd + 1) * cq);
                    for (int c = qd * cq; c < c1; ++c) {  // first maximum inside the quarter
                        const float sc = sigmoidf_<kFast>(to_f(pc[c]) + bias_c[c]) * q;
                        if (sc > best) { best = sc; bc = c; }
                    }
                }
#pragma unroll
                for (int o = 1; o <= 2; o <<= 1) {  // combine the 4 quarters; ties go to the lower class (cls.max(1) first-max rule)
                    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
                    const int oc = __shfl_xor_sync(0xffffffffu, bc, o);
                    if (ob > best || (ob == best && oc < bc)) { best = ob; bc = oc; }
                }
                const bool pass = av && qd == 0 && best > E.conf && (!E.class_keep || E.class_keep[bc]);
                const uint32_t idx = anchor * (uint32_t)nc + (uint32_t)bc;
                append(pass, ((unsigned long long)__float_as_uint(best) << 32) | (uint32_t)(~idx), stage, ti.b);
            }
        }
        __syncthreads();  // every read of this stage is done: refill it with the tile two rounds ahead
        if (tid == 0) {
            const int tn = t + 2 * gridDim.x;
            if (tn < total) issue(tn, stage);
            // close this tile's staging buffer and reserve its slots; the result is only needed during the next tile
            const int cnt = min(s_scnt[stage], s_slimit[stage]);
            s_fcnt[stage] = cnt; s_fimg[stage] = ti.b;
            s_scnt[stage] = 0; s_slimit[stage] = INT_MAX;
            pending_base = cnt ? atomicAdd(E.counts + ti.b, cnt) : 0;
        }
    }
    if (it > 0) {  // flush the last tile
        const int ps = (it - 1) & 1;
        if (tid == 0) s_fbase[ps] = pending_base;
        __syncthreads();
        const int fc = s_fcnt[ps];
        unsigned long long* dst = E.keys + (int64_t)s_fimg[ps] * E.key_stride + s_fbase[ps];
        for (int i = tid; i < fc; i += 256) dst[i] = s_stage(ps)[i];
    }
}

// ---------------------------------------------------------------------------- warp-autonomous fused emit kernel (16-bit maps, single label)
// gfl_decode_emit_kernel walks 64-anchor tiles with the whole CTA: three CTA barriers per tile, every warp in the same phase at the same
// time (all in the MUFU-heavy softmax, then all in the min / max network, then all in the class scan), issue slots 59 % busy at 24 warps
// per SM, and the per-tile bookkeeping (tile coordinates, staging hand-over) runs in all eight warps.  Here a WARP owns a slab of 16
// anchors from the bulk copy to the candidate keys and synchronises with nobody but itself:
//   lane 0 issues two 1-D bulk copies (box slab 2 KB, class slab 32 nc bytes) onto the warp's own mbarriers -- the next slab's box copy as
//   soon as the softmax phase has consumed the current one, its class copy after the class scan;
//   phase 1  lanes <-> (anchor of 8, side) twice: softmax statistics -> the warp's 16 x 28 TF32 statistics tile, DFL distances;
//   phase 2  the warp's own m16 tile through the DGQP hidden layer: 24 mma.sync m16n8k8, partial sums kept in the order of the CTA kernel
//            (hidden units 0-31 and 32-63 separately, then added), so quality and scores are bit-identical to it and to the dense kernel;
//   phase 3  lanes <-> (anchor of 8, class quarter) twice: logit maximum, windowed scoring, first-max combine; boxes by the quarter-0 lanes;
//   one atomicAdd per slab reserves the key slots of its <= 16 candidates.
// With 24 independent warps per SM in different phases the schedulers always find an issuable warp; __syncwarp is the only barrier.
constexpr int kSlab = 16;
constexpr int kWarpsPerCta = 12;

struct WarpEmitLayout { uint32_t w, bias, warp0, warp_stride, box, cls, stat, dist, bar, total; int bias_ld, w_stride; };
__host__ __device__ inline WarpEmitLayout warp_emit_layout(int nc, int nl) {
    WarpEmitLayout L;
    L.w_stride = WL<true>::total;
    L.bias_ld = (4 * kRegMax + nc + 3) & ~3;
    uint32_t cur = 0;
    L.w = cur; cur += (uint32_t)nl * L.w_stride * 4;
    L.bias = cur; cur += (uint32_t)nl * L.bias_ld * 4;
    cur = (cur + 127u) & ~127u;
    L.warp0 = cur;
    uint32_t o = 0;
    L.box = o; o += kSlab * 4 * kRegMax * 2;
    L.cls = o; o += ((uint32_t)kSlab * nc * 2 + 127u) & ~127u;
    L.stat = o; o += kSlab * WL<true>::stat_ld * 4;
    L.dist = o; o += 4 * kSlab * 4;
    L.bar = o; o += 16;
    L.warp_stride = (o + 127u) & ~127u;
    L.total = L.warp0 + kWarpsPerCta * L.warp_stride;
    return L;
}

template <typename T>
__global__ void __launch_bounds__(kWarpsPerCta * 32, 2) gfl_decode_emit_warp_kernel(const __grid_constant__ DecodeParams P, const __grid_constant__ EmitArgs E) {
    static_assert(sizeof(T) == 2, "16-bit maps only");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int LD = WL<true>::stat_ld;
    const int nc = P.nc, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const WarpEmitLayout ML = warp_emit_layout(nc, P.nl);
    float* s_w = (float*)(smem_raw + ML.w);
    float* s_bias = (float*)(smem_raw + ML.bias);
    unsigned char* wbase = smem_raw + ML.warp0 + warp * ML.warp_stride;
    T* s_box = (T*)(wbase + ML.box);
    T* s_cls = (T*)(wbase + ML.cls);
    float* s_stat = (float*)(wbase + ML.stat);
    float (*s_dist)[kSlab] = (float (*)[kSlab])(wbase + ML.dist);
    uint64_t* bar = (uint64_t*)(wbase + ML.bar);  // [0] box slab, [1] class slab

    for (int l = 0; l < P.nl; ++l) {
        load_level_weights<true>(s_w + l * ML.w_stride, P.lv[l]);
        for (int i = tid; i < 4 * kRegMax + nc; i += kWarpsPerCta * 32)
            s_bias[l * ML.bias_ld + i] = i < 4 * kRegMax ? (P.lv[l].bb ? __ldg(P.lv[l].bb + i) : 0.f) : (P.lv[l].cb ? __ldg(P.lv[l].cb + i - 4 * kRegMax) : 0.f);
    }
    if (lane < kSlab) *reinterpret_cast<float4*>(&s_stat[lane * LD + kStat]) = make_float4(0.f, 0.f, 0.f, 0.f);  // K padding 20 -> 24, written once
    if (lane == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // slabs of this warp: u = first + k * stride; slab u = quarter (u & 3) of tile (u >> 2)
    const int total = E.B * P.tiles_per_image * 4;
    const int stride = gridDim.x * kWarpsPerCta, first = blockIdx.x * kWarpsPerCta + warp;
    struct Slab { int b, l, pix0, nval; };
    auto slab_of = [&](int u) {
        const TileInfo ti = tile_info(P, u >> 2);
        Slab sl;
        sl.b = ti.b; sl.l = ti.l; sl.pix0 = ti.pix0 + kSlab * (u & 3);
        sl.nval = max(min(ti.nvalid - kSlab * (u & 3), kSlab), 0);
        return sl;
    };
    auto issue_box = [&](const Slab& sl) {   // lane 0 only
        const DecodeLevel& L = P.lv[sl.l];
        const uint32_t nb = (uint32_t)sl.nval * 4 * kRegMax * sizeof(T);
        mbar_expect_tx(&bar[0], nb);
        bulk_g2s(s_box, reinterpret_cast<const T*>(L.box) + (int64_t)sl.b * L.bs.n + (int64_t)sl.pix0 * (4 * kRegMax), nb, &bar[0]);
    };
    auto issue_cls = [&](const Slab& sl) {   // lane 0 only
        const DecodeLevel& L = P.lv[sl.l];
        const uint32_t ncb = (uint32_t)sl.nval * nc * sizeof(T);
        mbar_expect_tx(&bar[1], ncb);
        bulk_g2s(s_cls, reinterpret_cast<const T*>(L.cls) + (int64_t)sl.b * L.cs.n + (int64_t)sl.pix0 * nc, ncb, &bar[1]);
    };
    // skip empty slabs (quarters past the end of a ragged tile)
    auto next_live = [&](int u, Slab& sl) {
        for (; u < total; u += stride) {
            sl = slab_of(u);
            if (sl.nval > 0) return u;
        }
        return total;
    };
    Slab cur, nxt;
    int u = next_live(first, cur);
    if (u < total && lane == 0) { issue_box(cur); issue_cls(cur); }
    const int g = lane >> 2, t4 = lane & 3;
    const int hswap = g & 1;  // 1: this lane reads bins 8..15 of a side first (conflict-free 16 B shared loads)
    const int cq = (nc + 3) >> 2;
    uint32_t phase = 0;
    for (; u < total; phase ^= 1) {
        const int un = next_live(u + stride, nxt);
        const DecodeLevel& L = P.lv[cur.l];
        const float* w = s_w + cur.l * ML.w_stride;
        const float* bias_b = s_bias + cur.l * ML.bias_ld;
        const float* bias_c = bias_b + 4 * kRegMax;
        mbar_wait(&bar[0], phase);
        // ---- phase 1: (anchor g / g + 8, side t4)
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int a = g + 8 * i;
            float lg[kRegMax];
            const T* pv = s_box + a * (4 * kRegMax) + t4 * kRegMax;
            float f0[8], f1[8];
            unpack<T>(*reinterpret_cast<const uint4*>(pv + 8 * hswap), f0);
            unpack<T>(*reinterpret_cast<const uint4*>(pv + 8 * (hswap ^ 1)), f1);
#pragma unroll
            for (int k = 0; k < 8; ++k) { lg[k] = f0[k]; lg[8 + k] = f1[k]; }
            const float4* bb4 = reinterpret_cast<const float4*>(bias_b + t4 * kRegMax);
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                const float4 bv = bb4[v ^ (2 * hswap)];
                lg[4 * v] += bv.x; lg[4 * v + 1] += bv.y; lg[4 * v + 2] += bv.z; lg[4 * v + 3] += bv.w;
            }
            if (a >= cur.nval) {
#pragma unroll
                for (int k = 0; k < kRegMax; ++k) lg[k] = 0.f;
            }
            float dist;
            side_stats<true>(lg, dist, &s_stat[a * LD + t4 * 5], hswap != 0);
            s_dist[t4][a] = dist;
        }
        __syncwarp();
        if (lane == 0 && un < total) issue_box(nxt);  // the box slab has been consumed: prefetch the next one behind phases 2 and 3
        // ---- phase 2: DGQP hidden layer of the warp's own 16-anchor tile; z of anchors g (q0) and g + 8 (q1)
        float q0, q1;
        {
            const uint32_t* S = reinterpret_cast<const uint32_t*>(s_stat) + g * LD + t4;
            uint32_t a[3][4];
#pragma unroll
            for (int ks = 0; ks < 3; ++ks) {
                a[ks][0] = S[8 * ks]; a[ks][1] = S[8 * LD + 8 * ks]; a[ks][2] = S[8 * ks + 4]; a[ks][3] = S[8 * LD + 8 * ks + 4];
            }
            const uint2* Wf = reinterpret_cast<const uint2*>(w) + lane;
            const float2* b1 = reinterpret_cast<const float2*>(w + WL<true>::b1);
            const float2* w2 = reinterpret_cast<const float2*>(w + WL<true>::w2);
            float zp[2][2];  // [hidden half][row g / g + 8]: the CTA kernel's two partial sums
#pragma unroll
            for (int nh = 0; nh < 2; ++nh) {
                float z0 = 0.f, z1 = 0.f;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int nt = 4 * nh + j;
                    float d[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int ks = 0; ks < 3; ++ks) {
                        const uint2 bfr = Wf[(nt * 3 + ks) * 32];
                        mma_tf32(d, a[ks], bfr.x, bfr.y);
                    }
                    const float2 bb = b1[4 * nt + t4], ww = w2[4 * nt + t4];
                    z0 = fmaf(ww.x, fmaxf(d[0] + bb.x, 0.f), z0); z0 = fmaf(ww.y, fmaxf(d[1] + bb.y, 0.f), z0);
                    z1 = fmaf(ww.x, fmaxf(d[2] + bb.x, 0.f), z1); z1 = fmaf(ww.y, fmaxf(d[3] + bb.y, 0.f), z1);
                }
                z0 += __shfl_xor_sync(0xffffffffu, z0, 1); z0 += __shfl_xor_sync(0xffffffffu, z0, 2);
                z1 += __shfl_xor_sync(0xffffffffu, z1, 1); z1 += __shfl_xor_sync(0xffffffffu, z1, 2);
                zp[nh][0] = z0; zp[nh][1] = z1;
            }
            const float b2 = w[WL<true>::b2];
            float z = b2; z += zp[0][0] + zp[1][0];
            q0 = fminf(fmaxf(sigmoidf_<true>(z), 1e-6f), 1.f - 1e-6f);  // clamp(1e-6, 1 - 1e-6), head.py:343
            z = b2; z += zp[0][1] + zp[1][1];
            q1 = fminf(fmaxf(sigmoidf_<true>(z), 1e-6f), 1.f - 1e-6f);
        }
        mbar_wait(&bar[1], phase);
        // ---- phase 3: (anchor g / g + 8, class quarter t4)
        unsigned long long key[2];
        bool pass[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int a = g + 8 * i, qd = t4;
            const bool av = a < cur.nval;
            const float q = i == 0 ? q0 : q1;
            if (av && qd == 0) {
                const int pix = cur.pix0 + a, py = pix / L.W, px = pix - py * L.W;
                E.boxes[(int64_t)cur.b * P.A + L.a_off + pix] = decode_box(px, py, s_dist[0][a], s_dist[1][a], s_dist[2][a], s_dist[3][a], L.stride);
            }
            const T* pc = s_cls + a * nc;
            float best = -INFINITY;
            int bc = 0;
            if (nc == 80) {  // COCO-sized head: class maximum on the logits, only logits within 1/64 of it are scored (see gfl_decode_emit_kernel)
                constexpr int CQ = 20;
                float xs[CQ];
                const uint2* pcv = reinterpret_cast<const uint2*>(pc + qd * CQ);
                const float4* bv4 = reinterpret_cast<const float4*>(bias_c + qd * CQ);
#pragma unroll
                for (int v = 0; v < CQ / 4; ++v) {
                    const uint2 r = pcv[v];
                    const float4 bb = bv4[v];
                    unpack2<T>(r.x, xs[4 * v], xs[4 * v + 1]);
                    unpack2<T>(r.y, xs[4 * v + 2], xs[4 * v + 3]);
                    xs[4 * v] += bb.x; xs[4 * v + 1] += bb.y; xs[4 * v + 2] += bb.z; xs[4 * v + 3] += bb.w;
                }
                float mx = xs[0];
#pragma unroll
                for (int j = 1; j < CQ; ++j) mx = fmaxf(mx, xs[j]);
                mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
                mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
                const float bound = sigmoidf_<true>(mx) * q;
                if (av && bound * 1.0001f > E.conf) {
                    const float thr = mx > 8.f ? -INFINITY : mx - 0.015625f;
#pragma unroll
                    for (int j = 0; j < CQ; ++j) {
                        if (xs[j] >= thr) {
                            const float sc = sigmoidf_<true>(xs[j]) * q;
                            if (sc > best) { best = sc; bc = qd * CQ + j; }
                        }
                    }
                }
            } else if (av) {
                const int c1 = min(nc, (qd + 1) * cq);
                for (int c = qd * cq; c < c1; ++c) {  // first maximum inside the quarter
                    const float sc = sigmoidf_<true>(to_f(pc[c]) + bias_c[c]) * q;
                    if (sc > best) { best = sc; bc = c; }
                }
            }
#pragma unroll
            for (int o = 1; o <= 2; o <<= 1) {  // combine the 4 quarters; ties go to the lower class (cls.max(1) first-max rule)
                const float ob = __shfl_xor_sync(0xffffffffu, best, o);
                const int oc = __shfl_xor_sync(0xffffffffu, bc, o);
                if (ob > best || (ob == best && oc < bc)) { best = ob; bc = oc; }
            }
            pass[i] = av && qd == 0 && best > E.conf && (!E.class_keep || E.class_keep[bc]);
            const uint32_t idx = (uint32_t)(L.a_off + cur.pix0 + a) * (uint32_t)nc + (uint32_t)bc;
            key[i] = ((unsigned long long)__float_as_uint(best) << 32) | (uint32_t)(~idx);
        }
        __syncwarp();
        if (lane == 0 && un < total) issue_cls(nxt);  // the class slab has been consumed
        {   // one slot reservation per slab; anchors keep their order inside it (any order is fine: the keys are sorted later)
            const unsigned m0 = __ballot_sync(0xffffffffu, pass[0]), m1 = __ballot_sync(0xffffffffu, pass[1]);
            const int n0 = __popc(m0), nn = n0 + __popc(m1);
            if (nn) {
                int base = 0;
                if (lane == 0) base = atomicAdd(E.counts + cur.b, nn);
                base = __shfl_sync(0xffffffffu, base, 0);
                unsigned long long* dst = E.keys + (int64_t)cur.b * E.key_stride + base;
                const unsigned below = (1u << lane) - 1u;
                if (pass[0]) dst[__popc(m0 & below)] = key[0];
                if (pass[1]) dst[n0 + __popc(m1 & below)] = key[1];
            }
        }
        u = un;
        cur = nxt;
    }
}

// can the warp kernel stage this problem?  (emit_supported + 16-byte slab granularity of the class maps)
static bool warp_emit_supported(const DecodeParams& P) {
    for (int l = 0; l < P.nl; ++l) {
        const int rem = (P.lv[l].H * P.lv[l].W) % kSlab;
        if (rem && ((rem * P.nc * 2) % 16)) return false;
    }
    return warp_emit_layout(P.nc, P.nl).total <= 110 * 1024;
}

static size_t emit_smem_bytes(int nc, size_t esz, int nl) {
    return (esz == 2 ? emit_layout<true>(nc, (uint32_t)esz, nl).total : emit_layout<false>(nc, (uint32_t)esz, nl).total) + 64;
}

static int fill_params(DecodeParams& P, int nl, const void* const* box, const int64_t* box_s, const void* const* cls, const int64_t* cls_s, const int32_t* hw,
                       const float* stride, const float* const* w1, const float* const* b1, const float* const* w2, const float* const* b2,
                       const float* const* box_bias, const float* const* cls_bias, int nc) {
    P.nl = nl; P.nc = nc;
    int a_off = 0, tile_off = 0;
    for (int l = 0; l < nl; ++l) {
        DecodeLevel& L = P.lv[l];
        if (!box[l] || !cls[l] || !w1[l] || !b1[l] || !w2[l] || !b2[l] || hw[2 * l] <= 0 || hw[2 * l + 1] <= 0) return EL_ERR_ARG;
        L.box = box[l]; L.bs = s4(box_s + 4 * l);
        L.cls = cls[l]; L.cs = s4(cls_s + 4 * l);
        L.w1 = w1[l]; L.b1 = b1[l]; L.w2 = w2[l]; L.b2 = b2[l];
        L.bb = box_bias ? box_bias[l] : nullptr;
        L.cb = cls_bias ? cls_bias[l] : nullptr;
        L.H = hw[2 * l]; L.W = hw[2 * l + 1];
        L.stride = stride[l];
        L.a_off = a_off; L.tile_off = tile_off;
        a_off += L.H * L.W;
        tile_off += (int)ceil_div((int64_t)L.H * L.W, kTile);
    }
    P.A = a_off;
    P.tiles_per_image = tile_off;
    return EL_OK;
}

// can the engine kernel stage this problem with 1-D bulk copies?  (dense NHWC maps, 16 B aligned tile rows)
static bool emit_supported(const DecodeParams& P, size_t esz) {
    if (P.nc > 256) return false;
    for (int l = 0; l < P.nl; ++l) {
        const DecodeLevel& L = P.lv[l];
        const int64_t bc = 4 * kRegMax;
        if (L.bs.c != 1 || L.bs.w != bc || L.bs.h != bc * L.W || L.cs.c != 1 || L.cs.w != P.nc || L.cs.h != (int64_t)P.nc * L.W) return false;
        if (!aligned16(L.box) || !aligned16(L.cls)) return false;
        if ((L.bs.n * esz) % 16 || (L.cs.n * esz) % 16) return false;
        if ((kTile * P.nc * esz) % 16) return false;
        const int rem = (L.H * L.W) % kTile;
        if (rem && ((rem * P.nc * esz) % 16 || (rem * bc * esz) % 16)) return false;
    }
    return true;
}

}  // namespace el

using namespace el;

extern "C" int el_gfl_decode_fwd(int nl, const void* const* box, const int64_t* box_s, const void* const* cls, const int64_t* cls_s,
                                 const int32_t* hw, const float* stride, const float* const* w1, const float* const* b1, const float* const* w2,
                                 const float* const* b2, const float* const* box_bias, const float* const* cls_bias, float* y, float* q_out, int B,
                                 int nc, int dtype, void* stream) {
    if (nl <= 0 || nl > kMaxLevels || !box || !cls || !box_s || !cls_s || !hw || !stride || !w1 || !b1 || !w2 || !b2 || !y || B <= 0 || nc <= 0)
        return EL_ERR_ARG;
    DecodeParams P;
    if (int e = fill_params(P, nl, box, box_s, cls, cls_s, hw, stride, w1, b1, w2, b2, box_bias, cls_bias, nc)) return e;
    bool box_fast = true, cls_fast = true;
    for (int l = 0; l < nl; ++l) {
        const DecodeLevel& L = P.lv[l];
        // 16 B vector reads of the 16 bins of a side need channel-contiguous, 16 B aligned views
        box_fast = box_fast && L.bs.c == 1 && aligned16(L.box) && L.bs.n % 8 == 0 && L.bs.h % 8 == 0 && L.bs.w % 8 == 0;
        cls_fast = cls_fast && L.cs.c == 1;
    }
    const size_t cls_smem = (size_t)kTile * (nc + 1) * sizeof(float);
    if (cls_smem > 160 * 1024) cls_fast = false;  // very wide heads: direct (uncoalesced) reads
    dim3 grid(P.tiles_per_image, B);
    cudaStream_t st = (cudaStream_t)stream;
#define EL_LAUNCH_DECODE(BF, CS)                                                                                                  \
    do {                                                                                                                          \
        auto kern = gfl_decode_kernel<T, BF, CS>;                                                                                 \
        size_t sm = CS ? cls_smem : 0;                                                                                            \
        if (sm > 16 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);                     \
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

extern "C" int el_gfl_detect_workspace_bytes(int B, int nc, int A, int multi_label, int max_nms, size_t* bytes) {
    if (!bytes || B <= 0 || nc <= 0 || A <= 0 || max_nms <= 0) return EL_ERR_ARG;
    if ((int64_t)A * nc >= (int64_t)1 << 31) return EL_ERR_UNSUPPORTED;
    const NmsLayout L = nms_layout(B, nc, A, multi_label && nc > 1, max_nms);
    *bytes = ((L.total + 255) & ~(size_t)255) + (size_t)B * A * sizeof(float4);
    return EL_OK;
}

extern "C" int el_gfl_detect_fwd(int nl, const void* const* box, const int64_t* box_s, const void* const* cls, const int64_t* cls_s, const int32_t* hw,
                                 const float* stride, const float* const* w1, const float* const* b1, const float* const* w2, const float* const* b2,
                                 const float* const* box_bias, const float* const* cls_bias, int B, int nc, int dtype, float conf, double iou, int multi_label, int agnostic, const int32_t* class_keep,
                                 int max_det, int max_nms, float max_wh, int stages, void* workspace, size_t workspace_bytes, float* out, int32_t* out_count,
                                 int64_t* out_index, void* stream) {
    if (stages < 1 || stages > 7) return EL_ERR_ARG;
    if (nl <= 0 || nl > kMaxLevels || !box || !cls || !box_s || !cls_s || !hw || !stride || !w1 || !b1 || !w2 || !b2 || B <= 0 || nc <= 0 ||
        !workspace || !out || !out_count || max_det <= 0 || max_nms <= 0)
        return EL_ERR_ARG;
    if (conf < 0.f || conf > 1.f || iou < 0.0 || iou > 1.0) return EL_ERR_ARG;
    DecodeParams P;
    if (int e = fill_params(P, nl, box, box_s, cls, cls_s, hw, stride, w1, b1, w2, b2, box_bias, cls_bias, nc)) return e;
    if ((int64_t)P.A * nc >= (int64_t)1 << 31) return EL_ERR_UNSUPPORTED;
    const size_t esz = dtype == EL_F32 ? 4 : 2;
    if (!emit_supported(P, esz)) return EL_ERR_UNSUPPORTED;  // caller falls back to el_gfl_decode_fwd + el_nms_batched
    const bool multi = multi_label && nc > 1;
    const NmsLayout L = nms_layout(B, nc, P.A, multi, max_nms);
    const size_t box_off = (L.total + 255) & ~(size_t)255;
    if (workspace_bytes < box_off + (size_t)B * P.A * sizeof(float4)) return EL_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    char* ws = (char*)workspace;
    const bool do_emit = stages & 1;
    if (do_emit) nms_prepare(L, ws, st);
    EmitArgs E{conf, class_keep, (unsigned long long*)(ws + L.keys), L.key_stride, (int*)(ws + L.counts), (float4*)(ws + box_off), B};
    const size_t sm = emit_smem_bytes(nc, esz, nl);
    const int total = B * P.tiles_per_image;
    // persistent: as many CTAs as stay resident (16-bit maps: 3 per SM -- 56 / 74 registers and 75 KB of shared memory since the DGQP layer
    // moved to the tensor cores; fp32 maps: 2 per SM)
#define EL_LAUNCH_EMIT(M)                                                                               \
    do {                                                                                                \
        auto kern = gfl_decode_emit_kernel<T, M>;                                                       \
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);               \
        int occ = 0;                                                                                    \
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, sm) != cudaSuccess || occ < 1) occ = 2; \
        const int grid = total < occ * kSMs ? total : occ * kSMs;                                       \
        kern<<<grid, 256, sm, st>>>(P, E);                                                              \
    } while (0)
    static const int use_warp = [] { const char* e = getenv("EL_DECODE_WARP"); return e ? atoi(e) : 1; }();
    if (do_emit && use_warp && !multi && esz == 2 && warp_emit_supported(P)) {
        const WarpEmitLayout WLo = warp_emit_layout(nc, nl);
        const int slabs = total * 4;
        int grid = (int)ceil_div(slabs, kWarpsPerCta);
        if (grid > 2 * kSMs) grid = 2 * kSMs;
        if (dtype == EL_BF16) {
            cudaFuncSetAttribute(gfl_decode_emit_warp_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WLo.total);
            gfl_decode_emit_warp_kernel<__nv_bfloat16><<<grid, kWarpsPerCta * 32, WLo.total, st>>>(P, E);
        } else {
            cudaFuncSetAttribute(gfl_decode_emit_warp_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WLo.total);
            gfl_decode_emit_warp_kernel<__half><<<grid, kWarpsPerCta * 32, WLo.total, st>>>(P, E);
        }
        note_launches(1);
    } else if (do_emit) {
        EL_DISPATCH_DTYPE(dtype, {
            if (multi) EL_LAUNCH_EMIT(true);
            else EL_LAUNCH_EMIT(false);
        });
        note_launches(1);
    }
#undef EL_LAUNCH_EMIT
    if (int e = check_launch()) return e;
    return nms_finish(L, ws, BoxSource{(const float*)(ws + box_off), (int64_t)P.A * 4, 4, 1}, B, nc, iou, agnostic, max_det, max_nms, max_wh, stages, out,
                      out_count, out_index, st);
}
