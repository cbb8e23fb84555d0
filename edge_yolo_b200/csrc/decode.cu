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

#include "el_internal.h"

namespace el {

constexpr int kMaxLevels = 4;
constexpr int kTile = 64;        // anchors per tile
constexpr int kRegMax = 16;
constexpr int kStat = 20;        // 4 sides x (top-4 + mean)
constexpr int kStatLd = kStat + 1;
constexpr int kHidden = 64;
constexpr int kWFloats = kHidden * kStat + kHidden + kHidden + 4;  // w1 | b1 | w2 | b2 (+pad) per level = 1412

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
template <bool FAST> __device__ __forceinline__ float exp_(float x) { return FAST ? __expf(x) : expf(x); }
template <bool FAST> __device__ __forceinline__ float rcp_(float x) { return FAST ? __frcp_rn(x) : 1.f / x; }
template <bool FAST> __device__ __forceinline__ float sigmoidf_(float x) { return rcp_<FAST>(1.f + exp_<FAST>(-x)); }

// ---- shared arithmetic ------------------------------------------------------------------------
// softmax over the 16 bins of one side: DFL integral, sorted top-4 probabilities and their mean
template <bool FAST>
__device__ __forceinline__ void side_stats(float (&lg)[kRegMax], float& dist, float* __restrict__ stat5) {
    float m = lg[0];
#pragma unroll
    for (int k = 1; k < kRegMax; ++k) m = fmaxf(m, lg[k]);
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < kRegMax; ++k) { lg[k] = exp_<FAST>(lg[k] - m); s += lg[k]; }
    const float inv = rcp_<FAST>(s);
    float d = 0.f, psum = 0.f;
    float t0 = -1.f, t1 = -1.f, t2 = -1.f, t3 = -1.f;  // running top-4, descending
#pragma unroll
    for (int k = 0; k < kRegMax; ++k) {
        float p = lg[k] * inv;
        d += (float)k * p;
        psum += p;
        float v = p;
        float n0 = fmaxf(t0, v); v = fminf(t0, v); t0 = n0;
        float n1 = fmaxf(t1, v); v = fminf(t1, v); t1 = n1;
        float n2 = fmaxf(t2, v); v = fminf(t2, v); t2 = n2;
        t3 = fmaxf(t3, v);
    }
    dist = d;
    stat5[0] = t0; stat5[1] = t1; stat5[2] = t2; stat5[3] = t3;
    stat5[4] = psum * (1.f / kRegMax);  // prob.mean(dim=2): == 1/16, carries no information but is reproduced
}

// DGQP hidden layer for one 64-anchor tile with 256 threads: thread = (anchor pair ap, ap+32 ; 8 hidden units).
// w = [w1 (64x20) | b1 (64) | w2 (64) | b2] in shared memory; partial z sums go to s_part[8][kTile].
__device__ __forceinline__ void dgqp_hidden(const float* __restrict__ w, const float (*s_stat)[kStatLd], float (*s_part)[kTile]) {
    const int ap = threadIdx.x & 31, hg = threadIdx.x >> 5;
    float st0[kStat], st1[kStat];
#pragma unroll
    for (int c = 0; c < kStat; ++c) { st0[c] = s_stat[ap][c]; st1[c] = s_stat[ap + 32][c]; }
    const float* b1 = w + kHidden * kStat;
    const float* w2 = b1 + kHidden;
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

template <bool FAST>
__device__ __forceinline__ float dgqp_quality(const float* __restrict__ w, const float (*s_part)[kTile], int a) {
    float z = w[kHidden * kStat + 2 * kHidden];  // b2
    z += ((s_part[0][a] + s_part[1][a]) + (s_part[2][a] + s_part[3][a])) + ((s_part[4][a] + s_part[5][a]) + (s_part[6][a] + s_part[7][a]));
    const float q = sigmoidf_<FAST>(z);
    return fminf(fmaxf(q, 1e-6f), 1.f - 1e-6f);  // clamp(1e-6, 1 - 1e-6), head.py:343
}

// dist2bbox (xywh) * stride, tal.py:348-357 + head.py:341
__device__ __forceinline__ float4 decode_box(int px, int py, float l, float t, float r, float b, float stride) {
    const float ax = (float)px + 0.5f, ay = (float)py + 0.5f;
    const float x1 = ax - l, y1 = ay - t, x2 = ax + r, y2 = ay + b;
    return make_float4(((x1 + x2) / 2.f) * stride, ((y1 + y2) / 2.f) * stride, (x2 - x1) * stride, (y2 - y1) * stride);
}

__device__ __forceinline__ void load_level_weights(float* __restrict__ dst, const DecodeLevel& L) {
    for (int i = threadIdx.x; i < kHidden * kStat; i += blockDim.x) dst[i] = __ldg(L.w1 + i);
    if (threadIdx.x < kHidden) {
        dst[kHidden * kStat + threadIdx.x] = __ldg(L.b1 + threadIdx.x);
        dst[kHidden * kStat + kHidden + threadIdx.x] = __ldg(L.w2 + threadIdx.x);
    }
    if (threadIdx.x == 0) dst[kHidden * kStat + 2 * kHidden] = __ldg(L.b2);
}

// ---------------------------------------------------------------------------------- dense kernel
template <typename T, bool BOX_CH_FAST, bool CLS_STAGE>
__global__ void __launch_bounds__(256) gfl_decode_kernel(const __grid_constant__ DecodeParams P, float* __restrict__ y, float* __restrict__ q_out) {
    extern __shared__ float s_cls[];  // [kTile][nc+1] when CLS_STAGE
    __shared__ __align__(16) float s_w[kWFloats];
    __shared__ float s_stat[kTile][kStatLd];
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
    load_level_weights(s_w, L);

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
        side_stats<sizeof(T) == 2>(lg, dist, &s_stat[a][side * 5]);
        s_dist[side][a] = dist;
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
    dgqp_hidden(s_w, s_stat, s_part);
    __syncthreads();
    if (tid < kTile) {
        const int a = tid, pix = pix0 + a;
        const float q = dgqp_quality<sizeof(T) == 2>(s_w, s_part, a);
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

constexpr int kWStride = ((kWFloats * 4 + 15) & ~15) / 4;
constexpr int kStageCap = 1024;  // staged candidate keys per tile (8 KiB per stage); the rare overflow goes straight to global memory

template <typename T, bool MULTI>
__global__ void __launch_bounds__(256, 2) gfl_decode_emit_kernel(const __grid_constant__ DecodeParams P, const __grid_constant__ EmitArgs E) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int nc = P.nc, tid = threadIdx.x, lane = tid & 31;
    const uint32_t box_bytes = kTile * 4 * kRegMax * sizeof(T);
    const uint32_t cls_bytes = (uint32_t)((kTile * nc * sizeof(T) + 127) & ~127u);
    T* s_box[2]; T* s_clsT[2];
    unsigned char* cur = smem_raw;
    s_box[0] = (T*)cur; cur += box_bytes; s_box[1] = (T*)cur; cur += box_bytes;
    s_clsT[0] = (T*)cur; cur += cls_bytes; s_clsT[1] = (T*)cur; cur += cls_bytes;
    float* s_w = (float*)cur; cur += kMaxLevels * kWStride * 4;
    float (*s_stat)[kStatLd] = (float (*)[kStatLd])cur; cur += kTile * kStatLd * 4;
    float (*s_dist)[kTile] = (float (*)[kTile])cur; cur += 4 * kTile * 4;
    float (*s_part)[kTile] = (float (*)[kTile])cur; cur += 8 * kTile * 4;
    float* s_q = (float*)cur; cur += kTile * 4;
    const int bias_ld = (4 * kRegMax + nc + 3) & ~3;
    float* s_bias = (float*)cur; cur += kMaxLevels * bias_ld * 4;  // per level: box bias (64) | class bias (nc)
    uint64_t* bar = (uint64_t*)cur; cur += 16;
    // candidate keys are staged per tile in shared memory and flushed to the image's key list one tile later, so the global
    // atomic that reserves the slots (one per tile instead of one per warp) has a whole tile of work to hide behind
    unsigned long long* s_stage[2];
    s_stage[0] = (unsigned long long*)cur; cur += kStageCap * 8; s_stage[1] = (unsigned long long*)cur; cur += kStageCap * 8;
    int* s_scnt = (int*)cur;      // [2] staged count (may overshoot the capacity)
    int* s_slimit = s_scnt + 2;   // [2] first position that did not fit (INT_MAX if none)
    int* s_fcnt = s_scnt + 4;     // [2] number of staged keys to flush
    int* s_fbase = s_scnt + 6;    // [2] reserved base slot in the image's key list
    int* s_fimg = s_scnt + 8;     // [2] image of the staged tile
    if (tid < 2) { s_scnt[tid] = 0; s_slimit[tid] = INT_MAX; s_fcnt[tid] = 0; }

    for (int l = 0; l < P.nl; ++l) {
        load_level_weights(s_w + l * kWStride, P.lv[l]);
        for (int i = tid; i < 4 * kRegMax + nc; i += 256)
            s_bias[l * bias_ld + i] = i < 4 * kRegMax ? (P.lv[l].bb ? __ldg(P.lv[l].bb + i) : 0.f) : (P.lv[l].cb ? __ldg(P.lv[l].cb + i - 4 * kRegMax) : 0.f);
    }
    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int total = E.B * P.tiles_per_image;
    auto issue = [&](int t, int stage) {  // one thread: arm the barrier with the byte count, then two bulk copies
        const TileInfo ti = tile_info(P, t);
        const DecodeLevel& L = P.lv[ti.l];
        const uint32_t nb = (uint32_t)ti.nvalid * 4 * kRegMax * sizeof(T), ncb = (uint32_t)ti.nvalid * nc * sizeof(T);
        const T* gb = reinterpret_cast<const T*>(L.box) + (int64_t)ti.b * L.bs.n + (int64_t)ti.pix0 * (4 * kRegMax);
        const T* gc = reinterpret_cast<const T*>(L.cls) + (int64_t)ti.b * L.cs.n + (int64_t)ti.pix0 * nc;
        mbar_expect_tx(&bar[stage], nb + ncb);
        bulk_g2s(s_box[stage], gb, nb, &bar[stage]);
        bulk_g2s(s_clsT[stage], gc, ncb, &bar[stage]);
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
            if (pass) s_stage[stage][pos + rank] = key;
        } else {
            int gb = 0;
            if (lane == 0) { atomicMin(&s_slimit[stage], pos); gb = atomicAdd(E.counts + img, nn); }
            gb = __shfl_sync(0xffffffffu, gb, 0);
            if (pass) E.keys[(int64_t)img * E.key_stride + gb + rank] = key;
        }
    };
    for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
        const int stage = it & 1;
        const TileInfo ti = tile_info(P, t);
        const DecodeLevel& L = P.lv[ti.l];
        const float* w = s_w + ti.l * kWStride;
        const float* bias_b = s_bias + ti.l * bias_ld;
        const float* bias_c = bias_b + 4 * kRegMax;
        mbar_wait(&bar[stage], (it >> 1) & 1);

        {   // phase 1: thread = (anchor a, side): 16 logits from the staged tile
            const int side = tid & 3, a = tid >> 2;
            float lg[kRegMax];
            const T* pv = s_box[stage] + a * (4 * kRegMax) + side * kRegMax;
            if constexpr (sizeof(T) == 2) {
                const int h = (lane >> 2) & 1;  // alternate the half read first: conflict-free 16 B shared loads
                float f0[8], f1[8];
                unpack<T>(*reinterpret_cast<const uint4*>(pv + 8 * h), f0);
                unpack<T>(*reinterpret_cast<const uint4*>(pv + 8 * (h ^ 1)), f1);
#pragma unroll
                for (int k = 0; k < 8; ++k) { lg[k] = h ? f1[k] : f0[k]; lg[8 + k] = h ? f0[k] : f1[k]; }
            } else {
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    float f[4];
                    unpack<T>(*reinterpret_cast<const uint4*>(pv + 4 * v), f);
#pragma unroll
                    for (int k = 0; k < 4; ++k) lg[4 * v + k] = f[k];
                }
            }
#pragma unroll
            for (int k = 0; k < kRegMax; ++k) lg[k] = a < ti.nvalid ? lg[k] + bias_b[side * kRegMax + k] : 0.f;
            float dist;
            side_stats<sizeof(T) == 2>(lg, dist, &s_stat[a][side * 5]);
            s_dist[side][a] = dist;
        }
        __syncthreads();
        if (tid == 0 && it > 0) s_fbase[stage ^ 1] = pending_base;  // the previous tile's atomic has had a whole phase to return
        dgqp_hidden(w, s_stat, s_part);
        __syncthreads();
        if (tid < kTile) {
            const int a = tid;
            s_q[a] = dgqp_quality<sizeof(T) == 2>(w, s_part, a);
            if (a < ti.nvalid) {
                const int pix = ti.pix0 + a, py = pix / L.W, px = pix - py * L.W;
                E.boxes[(int64_t)ti.b * P.A + L.a_off + pix] = decode_box(px, py, s_dist[0][a], s_dist[1][a], s_dist[2][a], s_dist[3][a], L.stride);
            }
        }
        if (it > 0) {  // flush the previous tile's staged keys (coalesced 8-byte stores)
            const int ps = stage ^ 1, fc = s_fcnt[ps];
            unsigned long long* dst = E.keys + (int64_t)s_fimg[ps] * E.key_stride + s_fbase[ps];
            for (int i = tid; i < fc; i += 256) dst[i] = s_stage[ps][i];
        }
        __syncthreads();
        {   // phase 3: thread = (anchor a, quarter qd of the classes); scores never leave the SM unless they are candidates
            const int a = tid >> 2, qd = tid & 3;
            const bool av = a < ti.nvalid;
            const float q = s_q[a];
            const T* pc = s_clsT[stage] + a * nc;
            const uint32_t anchor = (uint32_t)(L.a_off + ti.pix0 + a);
            if (MULTI) {
                for (int j = 0; j < cq; ++j) {
                    const int c = qd * cq + j;
                    float sc = 0.f;
                    bool pass = false;
                    if (av && c < nc) {
                        sc = sigmoidf_<sizeof(T) == 2>(to_f(pc[c]) + bias_c[c]) * q;
                        pass = sc > E.conf && (!E.class_keep || E.class_keep[c]);
                    }
                    const uint32_t idx = anchor * (uint32_t)nc + (uint32_t)c;
                    append(pass, ((unsigned long long)__float_as_uint(sc) << 32) | (uint32_t)(~idx), stage, ti.b);
                }
            } else {
                float best = -INFINITY;
                int bc = 0;
                if (av) {
                    const int c1 = min(nc, (qd + 1) * cq);
                    for (int c = qd * cq; c < c1; ++c) {  // first maximum inside the quarter
                        const float sc = sigmoidf_<sizeof(T) == 2>(to_f(pc[c]) + bias_c[c]) * q;
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
        for (int i = tid; i < fc; i += 256) dst[i] = s_stage[ps][i];
    }
}

static size_t emit_smem_bytes(int nc, size_t esz) {
    const size_t box_bytes = (size_t)kTile * 4 * kRegMax * esz;
    const size_t cls_bytes = ((size_t)kTile * nc * esz + 127) & ~(size_t)127;
    const size_t bias_ld = (4 * kRegMax + nc + 3) & ~3;
    return 2 * box_bytes + 2 * cls_bytes + (size_t)kMaxLevels * kWStride * 4 + kTile * kStatLd * 4 + 4 * kTile * 4 + 8 * kTile * 4 + kTile * 4 +
           kMaxLevels * bias_ld * 4 + 16 + 2 * kStageCap * 8 + 64;
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
                                 int max_det, int max_nms, float max_wh, void* workspace, size_t workspace_bytes, float* out, int32_t* out_count,
                                 int64_t* out_index, void* stream) {
    if (nl <= 0 || nl > kMaxLevels || !box || !cls || !box_s || !cls_s || !hw || !stride || !w1 || !b1 || !w2 || !b2 || B <= 0 || nc <= 0 ||
        !workspace || !out || !out_count || max_det <= 0 || max_nms <= 0)
        return EL_ERR_ARG;
    if (conf < 0.f || conf > 1.f || iou < 0.0 || iou > 1.0) return EL_ERR_ARG;
    if (max_det > kMaxDetSmem) return EL_ERR_UNSUPPORTED;
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
    const bool do_emit = g_detect_stages & 1;
    if (do_emit) nms_prepare(L, ws, st);
    EmitArgs E{conf, class_keep, (unsigned long long*)(ws + L.keys), L.key_stride, (int*)(ws + L.counts), (float4*)(ws + box_off), B};
    const size_t sm = emit_smem_bytes(nc, esz);
    const int total = B * P.tiles_per_image;
    const int grid = total < 2 * kSMs ? total : 2 * kSMs;  // persistent: two CTAs per SM
#define EL_LAUNCH_EMIT(M)                                                                               \
    do {                                                                                                \
        auto kern = gfl_decode_emit_kernel<T, M>;                                                       \
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);               \
        kern<<<grid, 256, sm, st>>>(P, E);                                                              \
    } while (0)
    if (do_emit) {
        EL_DISPATCH_DTYPE(dtype, {
            if (multi) EL_LAUNCH_EMIT(true);
            else EL_LAUNCH_EMIT(false);
        });
        note_launches(1);
    }
#undef EL_LAUNCH_EMIT
    if (int e = check_launch()) return e;
    return nms_finish(L, ws, BoxSource{(const float*)(ws + box_off), (int64_t)P.A * 4, 4, 1}, B, nc, iou, agnostic, max_det, max_nms, max_wh, out,
                      out_count, out_index, st);
}
