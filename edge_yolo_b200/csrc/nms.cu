// Batched, sync-free GPU NMS.  Replaces non_max_suppression (utils/ops.py:167-316) and the
// torchvision.ops.nms call inside it (ops.py:296) for detection outputs (nm=0, not rotated).
//
// Pipeline per call (all device-side, no host round trip, CUDA-graph capturable):
//   1. emit      : conf filter + multi-label expansion / first-max class -> 64-bit keys
//                  key = score_bits << 32 | ~(anchor*nc+cls)   (unique, so "descending key" ==
//                  "stable descending score, ties by ascending candidate index" == torchvision order)
//   2. select    : only if more than max_nms candidates are possible: exact radix select of the
//                  max_nms-th largest key (8 x 8-bit passes) + compaction (ops.py:285-286)
//   3. sort      : bitonic sort of the keys, descending (shared-memory for <= 16384 keys/image)
//   4. sweep     : one CTA per image walks the sorted candidates in chunks of 256, keeps the list
//                  of kept boxes in shared memory and stops at max_det keeps.  Work is
//                  n x kept IoUs instead of the n^2/2 of a bitmask matrix, and there is no
//                  n x n/64 mask in HBM.
// IoU arithmetic replicates torchvision's fp32 sequence with round-to-nearest intrinsics (no FMA
// contraction) so keep decisions are bit-exact against the CPU op.
#include <cstdio>
#include <cstdlib>

#include "el_internal.h"

namespace el {

constexpr int kSortChunk = 16384;  // keys sorted per CTA in shared memory (128 KiB)
constexpr int kSortThreads = 1024;
constexpr int kSweepThreads = 512;    // 16 warps: half an SM's registers, so the forward graph's CTAs co-reside with a sweep CTA (the sweep runs beside the next forward)

// ------------------------------------------------------------------------------- 1. emit
template <bool MULTI>
__global__ void __launch_bounds__(256) nms_emit(const float* __restrict__ pred, int nc, int A, float conf, const int32_t* __restrict__ class_keep,
                                                unsigned long long* __restrict__ keys, int64_t key_stride, int* __restrict__ counts) {
    const int b = blockIdx.y, a = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31;
    const float* p = pred + ((int64_t)b * (4 + nc) + 4) * A;
    unsigned long long* kb = keys + (int64_t)b * key_stride;
    const bool in = a < A;
    if (MULTI) {
        for (int c = 0; c < nc; ++c) {
            float s = in ? __ldg(p + (int64_t)c * A + a) : 0.f;
            bool pass = in && s > conf && (!class_keep || class_keep[c]);
            unsigned m = __ballot_sync(0xffffffffu, pass);
            if (m) {
                int base = 0;
                if (lane == 0) base = atomicAdd(counts + b, __popc(m));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (pass) {
                    uint32_t idx = (uint32_t)a * (uint32_t)nc + (uint32_t)c;
                    kb[base + __popc(m & ((1u << lane) - 1))] = ((unsigned long long)__float_as_uint(s) << 32) | (uint32_t)(~idx);
                }
            }
        }
    } else {
        float best = -INFINITY;
        int j = 0;
        if (in)
            for (int c = 0; c < nc; ++c) {  // first maximum, like cls.max(1) on the CPU (ops.py:274)
                float s = __ldg(p + (int64_t)c * A + a);
                if (s > best) { best = s; j = c; }
            }
        bool pass = in && best > conf && (!class_keep || class_keep[j]);
        unsigned m = __ballot_sync(0xffffffffu, pass);
        if (m) {
            int base = 0;
            if (lane == 0) base = atomicAdd(counts + b, __popc(m));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (pass) {
                uint32_t idx = (uint32_t)a * (uint32_t)nc + (uint32_t)j;
                kb[base + __popc(m & ((1u << lane) - 1))] = ((unsigned long long)__float_as_uint(best) << 32) | (uint32_t)(~idx);
            }
        }
    }
}

// keys for the plain nms(boxes, scores) entry: any float score, order-preserving bit map
__global__ void __launch_bounds__(256) nms_box_keys(const float* __restrict__ scores, int n, unsigned long long* __restrict__ keys, int* __restrict__ count) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) *count = n;
    if (i >= n) return;
    uint32_t u = __float_as_uint(scores[i]);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);  // monotone float -> uint
    keys[i] = ((unsigned long long)u << 32) | (uint32_t)(~(uint32_t)i);
}

// ------------------------------------------------------------------------------- 2. select

__global__ void __launch_bounds__(256) select_hist(const unsigned long long* __restrict__ keys, int64_t key_stride, const int* __restrict__ counts,
                                                   int max_nms, const SelectState* __restrict__ st, unsigned* __restrict__ hist, int shift) {
    const int b = blockIdx.y, n = counts[b];
    if (n <= max_nms) return;
    __shared__ unsigned h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const unsigned long long pre = st[b].prefix;
    const unsigned long long* kb = keys + (int64_t)b * key_stride;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        unsigned long long k = kb[i];
        bool match = shift >= 56 ? true : ((k >> (shift + 8)) == (pre >> (shift + 8)));
        if (match) atomicAdd(&h[(unsigned)(k >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (h[threadIdx.x]) atomicAdd(hist + b * 256 + threadIdx.x, h[threadIdx.x]);
}

__global__ void __launch_bounds__(256) select_pick(const int* __restrict__ counts, int max_nms, SelectState* __restrict__ st, unsigned* __restrict__ hist,
                                                   int shift) {
    const int b = blockIdx.x, d = threadIdx.x;
    if (counts[b] <= max_nms) return;
    __shared__ unsigned suf[257];
    unsigned mine = hist[b * 256 + d];
    hist[b * 256 + d] = 0;  // ready for the next pass
    suf[d] = mine;
    if (d == 0) suf[256] = 0;
    __syncthreads();
    // inclusive suffix sum (Hillis-Steele over 256 entries)
    for (int o = 1; o < 256; o <<= 1) {
        unsigned v = (d + o < 256) ? suf[d + o] : 0;
        __syncthreads();
        suf[d] += v;
        __syncthreads();
    }
    const unsigned k = (unsigned)st[b].k_rem;
    __syncthreads();
    if (suf[d] >= k && suf[d + 1] < k) {  // exactly one digit satisfies this
        st[b].prefix |= (unsigned long long)d << shift;
        st[b].k_rem = (int)(k - suf[d + 1]);
    }
}

__global__ void __launch_bounds__(256) select_compact(const unsigned long long* __restrict__ keys, int64_t key_stride, const int* __restrict__ counts,
                                                      int max_nms, SelectState* __restrict__ st, unsigned long long* __restrict__ keys2,
                                                      int64_t key2_stride) {
    const int b = blockIdx.y, n = counts[b], lane = threadIdx.x & 31;
    const unsigned long long thr = n > max_nms ? st[b].prefix : 0ull;  // keys are unique: exactly max_nms keys are >= thr
    const unsigned long long* kb = keys + (int64_t)b * key_stride;
    unsigned long long* ob = keys2 + (int64_t)b * key2_stride;
    const int n_round = (n + 31) & ~31;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += gridDim.x * blockDim.x) {
        unsigned long long k = i < n ? kb[i] : 0ull;
        bool pass = i < n && k >= thr;
        unsigned m = __ballot_sync(0xffffffffu, pass);
        if (m) {
            int base = 0;
            if (lane == 0) base = atomicAdd(&st[b].count2, __popc(m));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (pass) ob[base + __popc(m & ((1u << lane) - 1))] = k;
        }
    }
}

__global__ void init_select(SelectState* st, int B, int max_nms) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) { st[b].prefix = 0; st[b].k_rem = max_nms; st[b].count2 = 0; }
}

// ------------------------------------------------------------------------------- 3. sort
// count source: plain int array or SelectState.count2
__device__ __forceinline__ int seg_count(const int* counts, const SelectState* st, int b, int cap) {
    int n = st ? st[b].count2 : counts[b];
    return n < cap ? n : cap;
}

// compare-exchange network on shared memory; element i sorts descending iff (g & k) == 0, g = global index
__device__ __forceinline__ void bitonic_smem(unsigned long long* s, int n_local, int g_base, int k_lo, int k_hi, int j_hi_first) {
    for (int k = k_lo; k <= k_hi; k <<= 1) {
        int j0 = (k == k_lo && j_hi_first) ? j_hi_first : (k >> 1);
        if (j0 > (n_local >> 1)) j0 = n_local >> 1;
        for (int j = j0; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < (n_local >> 1); t += blockDim.x) {
                int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));  // lower index of the pair
                int p = i | j;
                bool desc = (((g_base + i) & k) == 0);
                unsigned long long a = s[i], c = s[p];
                if ((a < c) == desc) { s[i] = c; s[p] = a; }
            }
            __syncthreads();
        }
    }
}

// whole segment in one CTA (capacity <= kSortChunk)
__global__ void __launch_bounds__(kSortThreads) sort_single(unsigned long long* __restrict__ keys, int64_t key_stride, const int* __restrict__ counts,
                                                             const SelectState* __restrict__ st, int cap) {
    extern __shared__ unsigned long long sk[];
    const int b = blockIdx.x;
    const int n = seg_count(counts, st, b, cap);
    if (n <= 1) return;
    const int m = (int)pow2ceil((uint32_t)n);
    unsigned long long* kb = keys + (int64_t)b * key_stride;
    for (int i = threadIdx.x; i < m; i += blockDim.x) sk[i] = i < n ? kb[i] : 0ull;
    __syncthreads();
    bitonic_smem(sk, m, 0, 2, m, 0);
    for (int i = threadIdx.x; i < n; i += blockDim.x) kb[i] = sk[i];
}

// multi-CTA path: (a) sort every chunk, (b) global exchange steps for j >= chunk, (c) finish inside chunks
__global__ void __launch_bounds__(kSortThreads) sort_chunks(unsigned long long* __restrict__ keys, int64_t key_stride, const int* __restrict__ counts,
                                                             const SelectState* __restrict__ st, int cap) {
    extern __shared__ unsigned long long sk[];
    const int b = blockIdx.y, base = blockIdx.x * kSortChunk;
    const int n = seg_count(counts, st, b, cap);
    if (base >= (int)pow2ceil((uint32_t)(n > 1 ? n : 1)) && base > 0) return;
    unsigned long long* kb = keys + (int64_t)b * key_stride + base;
    for (int i = threadIdx.x; i < kSortChunk; i += blockDim.x) sk[i] = (base + i) < n ? kb[i] : 0ull;
    __syncthreads();
    bitonic_smem(sk, kSortChunk, base, 2, kSortChunk, 0);
    for (int i = threadIdx.x; i < kSortChunk; i += blockDim.x) kb[i] = sk[i];  // padded capacity is a chunk multiple
}

__global__ void __launch_bounds__(256) sort_global_step(unsigned long long* __restrict__ keys, int64_t key_stride, const int* __restrict__ counts,
                                                        const SelectState* __restrict__ st, int cap, int cap_pow2, int k, int j) {
    const int b = blockIdx.y;
    const int n = seg_count(counts, st, b, cap);
    const int m = (int)pow2ceil((uint32_t)(n > 1 ? n : 1));
    if (k > m) return;  // stage not needed for this image
    unsigned long long* kb = keys + (int64_t)b * key_stride;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < (m >> 1); t += gridDim.x * blockDim.x) {
        int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        int p = i | j;
        bool desc = ((i & k) == 0);
        unsigned long long a = kb[i], c = kb[p];
        if ((a < c) == desc) { kb[i] = c; kb[p] = a; }
    }
}

__global__ void __launch_bounds__(kSortThreads) sort_chunk_merge(unsigned long long* __restrict__ keys, int64_t key_stride, const int* __restrict__ counts,
                                                                  const SelectState* __restrict__ st, int cap, int k) {
    extern __shared__ unsigned long long sk[];
    const int b = blockIdx.y, base = blockIdx.x * kSortChunk;
    const int n = seg_count(counts, st, b, cap);
    const int m = (int)pow2ceil((uint32_t)(n > 1 ? n : 1));
    if (k > m || base >= m) return;
    unsigned long long* kb = keys + (int64_t)b * key_stride + base;
    for (int i = threadIdx.x; i < kSortChunk; i += blockDim.x) sk[i] = kb[i];
    __syncthreads();
    bitonic_smem(sk, kSortChunk, base, k, k, kSortChunk >> 1);
    for (int i = threadIdx.x; i < kSortChunk; i += blockDim.x) kb[i] = sk[i];
}

// ------------------------------------------------------------------------------- 4. sweep
// torchvision's IoU test with every fp32 rounding made explicit (no FMA contraction):
//   inter/(Sa+Sb-inter) > thr ; thr is the double threshold rounded DOWN to float, which makes the
//   float compare equivalent to torchvision's CPU compare of the float IoU against a double.
// Disjoint boxes (the common case, also every cross-class pair thanks to the class offset) leave
// before the division: inter == 0 gives 0/uni = 0 or 0/0 = NaN, neither of which is > thr >= 0.
__device__ __forceinline__ bool iou_gt(float ax1, float ay1, float ax2, float ay2, float aarea, float bx1, float by1, float bx2, float by2, float barea,
                                       float thr) {
    float w = __fsub_rn(fminf(ax2, bx2), fmaxf(ax1, bx1)), h = __fsub_rn(fminf(ay2, by2), fmaxf(ay1, by1));
    if (!(w > 0.f) || !(h > 0.f)) return false;
    float inter = __fmul_rn(w, h);
    float uni = __fsub_rn(__fadd_rn(aarea, barea), inter);
    // Division-free answer when the quotient is clearly on one side of the threshold: p = thr * uni carries one rounding (2^-24), the
    // quotient another, so a 2e-6 margin decides exactly as the rounded division would; only the band around the threshold (and
    // degenerate magnitudes) pays for the IEEE division.  The result is the division's in every case.
    if (uni > 1e-30f && uni < 1e30f && inter > 1e-30f && thr > 0.f) {
        const float p = __fmul_rn(thr, uni);
        if (inter > __fmul_rn(p, 1.000002f)) return true;
        if (inter < __fmul_rn(p, 0.999998f)) return false;
    }
    return __fdiv_rn(inter, uni) > thr;
}

// Branch-free form for unrolled loops: 0 = no, 1 = yes, 2 = inside the band around the threshold (the caller evaluates iou_gt, i.e. the
// IEEE division, for those).  Same arithmetic as iou_gt up to the division.
__device__ __forceinline__ int iou_class(float4 a, float aarea, float bx1, float by1, float bx2, float by2, float barea, float thr) {
    const float w = __fsub_rn(fminf(a.z, bx2), fmaxf(a.x, bx1)), h = __fsub_rn(fminf(a.w, by2), fmaxf(a.y, by1));
    const float inter = __fmul_rn(w, h);
    const float uni = __fsub_rn(__fadd_rn(aarea, barea), inter);
    const float p = __fmul_rn(thr, uni);
    const bool ov = (w > 0.f) & (h > 0.f);
    const bool ranged = (uni > 1e-30f) & (uni < 1e30f) & (inter > 1e-30f) & (thr > 0.f);
    const bool yes = ov & ranged & (inter > __fmul_rn(p, 1.000002f));
    const bool no = !ov | (ranged & (inter < __fmul_rn(p, 0.999998f)));
    return yes ? 1 : (no ? 0 : 2);
}

struct SweepArgs {
    const unsigned long long* keys; int64_t key_stride;
    const int* counts; const SelectState* st; int cap;
    // FROM_PRED: component k of anchor a of image b lives at box[b*sb + a*sa + k*sk] (xywh); else raw xyxy boxes (n,4)
    const float* box; int64_t sb, sa, sk;
    int nc; float max_wh; int agnostic;
    float thr; int max_det;
    float* out; int32_t* out_count; int64_t* out_index;  // batched outputs
    int64_t* keep;                                       // nms_boxes output
    float* gkept;                                        // global kept list (5 x cap) when !SMEM_KEPT
    const int* handled;                                  // per image: 1 = already done by nms_sweep_classes (may be NULL)
};

constexpr int kMaxBucketClasses = 2048;  // class tags are used up to this many classes

// raw (un-offset) xyxy box + class of sorted candidate i; shared by the bounds pre-pass and the sweep
struct Cand { float rx1, ry1, rx2, ry2, score, clsf; uint32_t idx; int cls; };
__device__ __forceinline__ Cand load_cand(const SweepArgs& P, const unsigned long long* kb, int b, int i) {
    Cand c;
    const unsigned long long k = kb[i];
    c.idx = ~(uint32_t)k;
    c.score = __uint_as_float((uint32_t)(k >> 32));
    const uint32_t a = c.idx / (uint32_t)P.nc;
    c.cls = (int)(c.idx - a * (uint32_t)P.nc);
    c.clsf = (float)c.cls;
    const float* pb = P.box + (int64_t)b * P.sb + (int64_t)a * P.sa;
    const float cx = __ldg(pb), cy = __ldg(pb + P.sk), w = __ldg(pb + 2 * P.sk), h = __ldg(pb + 3 * P.sk);
    const float hw = __fdiv_rn(w, 2.f), hh = __fdiv_rn(h, 2.f);  // xywh2xyxy, ops.py:416-433
    c.rx1 = __fsub_rn(cx, hw); c.ry1 = __fsub_rn(cy, hh); c.rx2 = __fadd_rn(cx, hw); c.ry2 = __fadd_rn(cy, hh);
    return c;
}

// Greedy sweep of one image's sorted candidates, 1024 per chunk, 32 (one warp) per step.  The serial chain of NMS is
//   "test the warp's candidates against the keeps of the previous step -> resolve the warp -> publish its keeps";
// everything else is kept off it.  History (bench.py regime: 2000 tied candidates per image, one dominant class, 54 steps; clock64
// traces with -DEL_SWEEP_TRACE): the first version spent ~6600 cycles per step -- the resolving warp wrote its output rows to global
// memory and fenced once per keep so that other warps could follow per-class linked lists concurrently, and evaluated an IEEE division
// per overlapping pair inside its lane-serial loop.  Now:
//   * kept boxes live in shared memory (float4 + area) with a class tag; a candidate tests the keeps it has not seen yet by INDEX
//     ([checked, nk), rounds of four branch-free IoU classes, class compare first), so new keeps are only read after the step's
//     barrier: no fences, no lists to link;
//   * the IEEE division is evaluated only inside a 2e-6 band around the threshold (iou_class / iou_gt), results unchanged;
//   * every warp precomputes, before its turn, which of its higher lanes each lane suppresses (32 x 32 pair masks), so resolving a
//     warp is a bit loop over its keeps: ffs / shfl / and;
//   * output rows are written once, at the end, by all threads (the kept candidates' sorted indices are remembered in shared memory).
// Also measured and dropped: only the next two resolvers walking (the pre-walk of a large backlog then sits on the barrier: 281 us
// against 169), and a barrier-free producer / consumer variant (one chain warp with a scheduler of its own, 24 tester warps covering the
// groups ahead of it, ld.acquire / st.release hand-over: 190 us -- a single warp issues a dependent instruction every ~5 cycles, so the
// chain's own 14 tests + bit loop per group cost what the contention had cost before).
template <bool SMEM_KEPT, bool FROM_PRED>
__global__ void __launch_bounds__(kSweepThreads) nms_sweep(const __grid_constant__ SweepArgs P) {
    extern __shared__ __align__(16) float s_kept[];  // SMEM_KEPT: kept boxes float4[kcap], areas[kcap]; then (FROM_PRED) class tags[kcap], sorted indices[kcap]
    __shared__ int s_nk[2];
    __shared__ float s_red[2][kSweepThreads / 32];
    __shared__ int s_bucketed;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (P.handled && P.handled[b]) return;
#ifdef EL_SWEEP_TRACE
    __shared__ long long s_tr[8];
    if (tid < 8) s_tr[tid] = 0;
    const long long tr_begin = clock64();
#define TR_MARK(var) const long long var = clock64()
#define TR_ADD(k, a, b_) do { if (lane == 0) atomicAdd((unsigned long long*)&s_tr[k], (unsigned long long)((b_) - (a))); } while (0)
#else
#define TR_MARK(var)
#define TR_ADD(k, a, b_)
#endif
    const int n = seg_count(P.counts, P.st, b, P.cap);
    const int kcap = P.max_det < P.cap ? P.max_det : P.cap;  // keeps never outnumber candidates
    // !SMEM_KEPT: the kept list lives in global memory (max_det beyond kMaxDetSmem, or el_nms_boxes where every box may be kept); the
    // batched form gives every image 8 * cap floats of it
    float4* kbox = reinterpret_cast<float4*>(SMEM_KEPT ? s_kept : P.gkept + (FROM_PRED ? (int64_t)b * 8 * P.cap : 0));
    float* kar = reinterpret_cast<float*>(kbox + kcap);
    const unsigned long long* kb = P.keys + (int64_t)b * P.key_stride;
    // one candidate against kept entry j: 0 / 1, the band around the threshold resolved by the exact division
    auto hits = [&](int j, float ox1, float oy1, float ox2, float oy2, float area) -> bool {
        const float4 kbx = kbox[j];
        const float ka = kar[j];
        const int r = iou_class(kbx, ka, ox1, oy1, ox2, oy2, area, P.thr);
        return r == 2 ? iou_gt(kbx.x, kbx.y, kbx.z, kbx.w, ka, ox1, oy1, ox2, oy2, area, P.thr) : r == 1;
    };
    // Class tags (FROM_PRED only).  The reference separates classes by adding cls*max_wh to the coordinates (ops.py:289-295); when
    // every candidate coordinate of this image lies in an interval narrower than max_wh, boxes of different classes are disjoint
    // after the offset, so a candidate only needs the IoU against kept boxes of its own class -- same result, the other pairs cost a
    // compare.  Otherwise (or class-agnostic) every keep carries tag 0.
    int* s_kcls = reinterpret_cast<int*>(kar + kcap);
    int* s_ki = s_kcls + kcap;
    const int nb = (FROM_PRED && !P.agnostic && P.nc <= kMaxBucketClasses) ? P.nc : 1;
    if (FROM_PRED) {
        float lo = INFINITY, hi = -INFINITY;
        for (int i = tid; i < n; i += kSweepThreads) {
            const Cand c = load_cand(P, kb, b, i);
            lo = fminf(lo, fminf(fminf(c.rx1, c.ry1), fminf(c.rx2, c.ry2)));
            hi = fmaxf(hi, fmaxf(fmaxf(c.rx1, c.ry1), fmaxf(c.rx2, c.ry2)));
        }
        lo = -warp_max(-lo); hi = warp_max(hi);
        if (lane == 0) { s_red[0][warp] = lo; s_red[1][warp] = hi; }
    }
    if (tid == 0) { s_nk[0] = 0; s_nk[1] = 0; }
    __syncthreads();
    if (FROM_PRED && tid == 0) {
        float lo = INFINITY, hi = -INFINITY;
        for (int w = 0; w < kSweepThreads / 32; ++w) { lo = fminf(lo, s_red[0][w]); hi = fmaxf(hi, s_red[1][w]); }
        // strict: NaN / inf coordinates fall back to the single list
        s_bucketed = (nb > 1 && (hi - lo) < P.max_wh && P.max_wh > 0.f) ? 1 : 0;
    }
    __syncthreads();
    const bool bucketed = FROM_PRED && s_bucketed;
    int step = 0;
    bool done = false;
    for (int c0 = 0; c0 < n && !done; c0 += kSweepThreads) {
        const int i = c0 + tid;
        const bool valid = i < n;
        uint32_t cidx = 0;
        float ox1 = 0.f, oy1 = 0.f, ox2 = 0.f, oy2 = 0.f, area = 0.f;
        int bucket = 0;
        if (valid) {
            if (FROM_PRED) {
                const Cand cd = load_cand(P, kb, b, i);
                const float off = P.agnostic ? 0.f : __fmul_rn(cd.clsf, P.max_wh);  // ops.py:289
                ox1 = __fadd_rn(cd.rx1, off); oy1 = __fadd_rn(cd.ry1, off); ox2 = __fadd_rn(cd.rx2, off); oy2 = __fadd_rn(cd.ry2, off);
                bucket = bucketed ? cd.cls : 0;
            } else {
                cidx = ~(uint32_t)kb[i];
                const float4 bx = __ldg(reinterpret_cast<const float4*>(P.box) + cidx);
                ox1 = bx.x; oy1 = bx.y; ox2 = bx.z; oy2 = bx.w;
            }
            area = __fmul_rn(__fsub_rn(ox2, ox1), __fsub_rn(oy2, oy1));
        }
        bool alive = valid;
        TR_MARK(tc0);
        // walk kept entries [j0, j1) for this thread's candidate: rounds of four branch-free IoU classes (the loads and the tests of a
        // round overlap; rounds of eight were measured slower: a step publishes ~6 keeps, the padding of the second half is wasted
        // issue slots for all 32 warps), class compare first
        auto walk = [&](int j0, int j1) {
            for (int j = j0; j < j1 && alive; j += 4) {
                int any1 = 0, any2 = 0;
                int code[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int jj = min(j + u, j1 - 1);  // the tail repeats the last entry (idempotent)
                    const int r = iou_class(kbox[jj], kar[jj], ox1, oy1, ox2, oy2, area, P.thr);
                    code[u] = (FROM_PRED && s_kcls[jj] != bucket) ? 0 : r;
                    any1 |= code[u] & 1; any2 |= code[u] & 2;
                }
                if (any1) alive = false;
                else if (any2) {  // rare: inside the band around the threshold -> the division decides
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (code[u] == 2 && hits(min(j + u, j1 - 1), ox1, oy1, ox2, oy2, area)) alive = false;
                }
            }
        };
        // ---- chunk start: everything kept by earlier chunks
        int checked = s_nk[step & 1];
        if (alive && checked > 0) walk(0, checked);
        TR_MARK(tc1);
        if (warp == 0) TR_ADD(4, tc0, tc1);
        // ---- pair masks of this warp: bit j of `sup` = this lane suppresses lane j > lane (same class list, IoU > thr)
        unsigned sup = 0;
        if (__any_sync(0xffffffffu, alive)) {
#pragma unroll 4
            for (int j = 1; j < 32; ++j) {
                const float jx1 = __shfl_sync(0xffffffffu, ox1, j), jy1 = __shfl_sync(0xffffffffu, oy1, j);
                const float jx2 = __shfl_sync(0xffffffffu, ox2, j), jy2 = __shfl_sync(0xffffffffu, oy2, j);
                const float jar = __shfl_sync(0xffffffffu, area, j);
                const int jb = FROM_PRED ? __shfl_sync(0xffffffffu, bucket, j) : 0;
                const bool jal = __shfl_sync(0xffffffffu, (int)alive, j);
                if (alive && jal && lane < j && jb == bucket && iou_gt(ox1, oy1, ox2, oy2, area, jx1, jy1, jx2, jy2, jar, P.thr)) sup |= 1u << j;
            }
        }
        TR_MARK(tc2);
        if (warp == 0) TR_ADD(5, tc1, tc2);
        for (int sub = 0; sub < kSweepThreads / 32; ++sub, ++step) {
            if (c0 + sub * 32 >= n) break;  // no candidates left for the remaining warps (uniform)
            const int nk = s_nk[step & 1];
            TR_MARK(ts0);
            if (warp >= sub && alive) {  // keeps published since this candidate last looked (all from this chunk)
                walk(checked, nk);
            }
            checked = nk;
            TR_MARK(ts1);
            if (warp == sub) TR_ADD(0, ts0, ts1);
            if (warp == sub) {
                // greedy resolution inside the warp from the precomputed masks: the lowest alive lane is kept and kills its overlaps
                unsigned m = __ballot_sync(0xffffffffu, alive), K = 0;
                while (m) {
                    const int j = __ffs(m) - 1;
                    K |= 1u << j;
                    m &= ~(__shfl_sync(0xffffffffu, sup, j) | (1u << j));
                }
                TR_MARK(ts2);
                TR_ADD(1, ts1, ts2);
                const int rank = nk + __popc(K & ((1u << lane) - 1u));
                const bool keep = ((K >> lane) & 1u) && rank < P.max_det;
                if (keep) {
                    kbox[rank] = make_float4(ox1, oy1, ox2, oy2); kar[rank] = area;
                    if (FROM_PRED) { s_kcls[rank] = bucket; s_ki[rank] = i; }
                    else P.keep[rank] = (int64_t)cidx;
                }
                if (lane == 0) {
                    int nn = nk + __popc(K);
                    s_nk[(step + 1) & 1] = nn < P.max_det ? nn : P.max_det;
                }
                TR_MARK(ts3);
                TR_ADD(2, ts2, ts3);
            }
            TR_MARK(ts4);
            __syncthreads();  // publishes the new keeps (shared or global memory) to the other warps of the CTA
            TR_MARK(ts5);
            if (warp == sub) TR_ADD(3, ts4, ts5);
            if (warp == ((sub + 16) & 31)) TR_ADD(6, ts0, ts5);
            if (s_nk[(step + 1) & 1] >= P.max_det) { done = true; ++step; break; }
        }
    }
    const int nkf = s_nk[step & 1];
#ifdef EL_SWEEP_TRACE
    __syncthreads();
    if (b == 0 && tid == 0)
        printf("sweep trace image 0: n %d steps %d kept %d | total %lld cyc | resolver: walk %lld resolve %lld publish %lld barrier %lld | chunk: bulk %lld pairmask %lld | other-warp step total %lld\n",
               n, step, nkf, clock64() - tr_begin, s_tr[0], s_tr[1], s_tr[2], s_tr[3], s_tr[4], s_tr[5], s_tr[6]);
#endif
    if (FROM_PRED) {  // output rows, in rank order, from the remembered sorted indices
        for (int r = tid; r < nkf; r += kSweepThreads) {
            const Cand cd = load_cand(P, kb, b, s_ki[r]);
            float* o = P.out + ((int64_t)b * P.max_det + r) * 6;
            o[0] = cd.rx1; o[1] = cd.ry1; o[2] = cd.rx2; o[3] = cd.ry2; o[4] = cd.score; o[5] = cd.clsf;
            if (P.out_index) P.out_index[(int64_t)b * P.max_det + r] = (int64_t)cd.idx;
        }
    }
    if (tid == 0) P.out_count[b] = nkf;
}

// ---- class-parallel sweep ---------------------------------------------------------------------------
// When the class offset makes boxes of different classes disjoint (see nms_sweep), NMS decomposes into independent
// per-class problems; only the final "first max_det in global score order" couples them.  One CTA per image:
//   1. classes of the sorted candidates -> stable counting sort by class in shared memory (per-class lists in score order)
//   2. each warp runs the greedy sweep of whole classes (32 candidates per step, kept boxes of the class in an L1-resident
//      scratch), a class stops after max_det keeps (later ones can never reach the global top max_det)
//   3. ordered block scan over the keep flags picks the first max_det survivors in global score order.
// The serial chain is the longest class instead of the whole image.  Images that fail the disjointness test (or
// class-agnostic calls) are left to nms_sweep, flagged through `handled`.
constexpr int kClsThreads = 1024;
constexpr int kClsMaxList = 512;  // longest per-class list one warp is allowed to walk

__global__ void __launch_bounds__(kClsThreads) nms_sweep_classes(const __grid_constant__ SweepArgs P, float* __restrict__ gkept_all,
                                                                 int* __restrict__ handled) {
    extern __shared__ unsigned char s_raw[];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = seg_count(P.counts, P.st, b, P.cap);
    const int nb = P.nc;
    uint16_t* s_cls = reinterpret_cast<uint16_t*>(s_raw);
    uint16_t* s_list = s_cls + P.cap;
    int* s_segcnt = reinterpret_cast<int*>(s_raw + (((size_t)4 * P.cap + 15) & ~(size_t)15));
    int* s_off = s_segcnt + kClsThreads;  // nb + 1
    uint8_t* s_keep = reinterpret_cast<uint8_t*>(s_off + nb + 1);
    __shared__ float s_red[2][kClsThreads / 32];
    __shared__ int s_wsum[kClsThreads / 32];
    __shared__ int s_flag, s_running;
    const unsigned long long* kb = P.keys + (int64_t)b * P.key_stride;

    // cheap test first: class histogram from the keys alone; a dominant class (long chain) is left to nms_sweep
    for (int i = tid; i <= nb; i += kClsThreads) s_off[i] = 0;
    if (tid == 0) { s_flag = 1; s_running = 0; }
    __syncthreads();
    for (int i = tid; i < n; i += kClsThreads) {
        const uint32_t idx = ~(uint32_t)kb[i];
        const int c = (int)(idx % (uint32_t)P.nc);
        s_cls[i] = (uint16_t)c;
        s_keep[i] = 0;
        if (atomicAdd(&s_off[c + 1], 1) + 1 > kClsMaxList) s_flag = 0;
    }
    __syncthreads();
    if (!s_flag) {
        if (tid == 0) handled[b] = 0;
        return;
    }
    float lo = INFINITY, hi = -INFINITY;
    for (int i = tid; i < n; i += kClsThreads) {
        const Cand c = load_cand(P, kb, b, i);
        lo = fminf(lo, fminf(fminf(c.rx1, c.ry1), fminf(c.rx2, c.ry2)));
        hi = fmaxf(hi, fmaxf(fmaxf(c.rx1, c.ry1), fmaxf(c.rx2, c.ry2)));
    }
    lo = -warp_max(-lo); hi = warp_max(hi);
    if (lane == 0) { s_red[0][warp] = lo; s_red[1][warp] = hi; }
    __syncthreads();
    if (tid == 0) {
        float l2 = INFINITY, h2 = -INFINITY;
        for (int w = 0; w < kClsThreads / 32; ++w) { l2 = fminf(l2, s_red[0][w]); h2 = fmaxf(h2, s_red[1][w]); }
        s_flag = (n == 0 || ((h2 - l2) < P.max_wh)) ? 1 : 0;  // NaN / inf coordinates fail the test -> serial kernel
        handled[b] = s_flag;
    }
    __syncthreads();
    if (!s_flag) return;
    if (n == 0) { if (tid == 0) P.out_count[b] = 0; return; }

    // ---- 1. stable counting sort by class: thread = (class c, segment of the candidate list)
    const int S = kClsThreads / nb, seglen = (n + S - 1) / S;
    const int c_t = tid % nb, seg_t = tid / nb;
    const bool worker = tid < nb * S;
    const int i0 = seg_t * seglen, i1 = min(n, i0 + seglen);
    if (worker) {
        int cnt = 0;
        for (int i = i0; i < i1; ++i) cnt += (s_cls[i] == c_t);  // all lanes read the same element: broadcast
        s_segcnt[seg_t * nb + c_t] = cnt;
    }
    __syncthreads();
    if (tid < nb) {  // per class: exclusive prefix over the segments, total into s_off[c + 1]
        int run = 0;
        for (int sg = 0; sg < S; ++sg) { const int t = s_segcnt[sg * nb + tid]; s_segcnt[sg * nb + tid] = run; run += t; }
        s_off[tid + 1] = run;
    }
    __syncthreads();
    if (warp == 0) {  // exclusive scan of the class totals (nb <= 1024): 32 classes per step
        int carry = 0;
        for (int base = 0; base < nb; base += 32) {
            int v = base + lane < nb ? s_off[base + lane + 1] : 0;
            int incl = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
            if (base + lane < nb) s_off[base + lane + 1] = carry + incl;
            carry += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (lane == 0) s_off[0] = 0;
    }
    __syncthreads();
    if (worker) {
        int pos = s_off[c_t] + s_segcnt[seg_t * nb + c_t];
        for (int i = i0; i < i1; ++i)
            if (s_cls[i] == c_t) s_list[pos++] = (uint16_t)i;
    }
    __syncthreads();

    // ---- 2. per-class greedy sweeps, one warp per class
    for (int c = warp; c < nb; c += kClsThreads / 32) {
        const int base = s_off[c], cnt = s_off[c + 1] - base;
        if (cnt == 0) continue;
        float* gk = gkept_all + ((int64_t)b * P.cap + base) * 5;  // this class's kept boxes: [k][x1,y1,x2,y2,area]
        const float off = __fmul_rn((float)c, P.max_wh);            // ops.py:289
        int kc = 0;
        for (int g0 = 0; g0 < cnt && kc < P.max_det; g0 += 32) {
            const int li = g0 + lane;
            const bool valid = li < cnt;
            int i = 0;
            float ox1 = 0.f, oy1 = 0.f, ox2 = 0.f, oy2 = 0.f, area = 0.f;
            if (valid) {
                i = s_list[base + li];
                const Cand cd = load_cand(P, kb, b, i);
                ox1 = __fadd_rn(cd.rx1, off); oy1 = __fadd_rn(cd.ry1, off); ox2 = __fadd_rn(cd.rx2, off); oy2 = __fadd_rn(cd.ry2, off);
                area = __fmul_rn(__fsub_rn(ox2, ox1), __fsub_rn(oy2, oy1));
            }
            bool alive = valid;
            for (int j = 0; j < kc; ++j) {  // kept boxes of this class (written by this warp, warp-uniform address)
                if (!__any_sync(0xffffffffu, alive)) break;
                const float* kj = gk + j * 5;
                if (alive && iou_gt(kj[0], kj[1], kj[2], kj[3], kj[4], ox1, oy1, ox2, oy2, area, P.thr)) alive = false;
            }
            unsigned m = __ballot_sync(0xffffffffu, alive), K = 0;
            while (m) {  // greedy resolution inside the group
                const int j = __ffs(m) - 1;
                K |= 1u << j;
                float jx1 = __shfl_sync(0xffffffffu, ox1, j), jy1 = __shfl_sync(0xffffffffu, oy1, j);
                float jx2 = __shfl_sync(0xffffffffu, ox2, j), jy2 = __shfl_sync(0xffffffffu, oy2, j);
                float jar = __shfl_sync(0xffffffffu, area, j);
                if (alive && lane > j && iou_gt(jx1, jy1, jx2, jy2, jar, ox1, oy1, ox2, oy2, area, P.thr)) alive = false;
                m = __ballot_sync(0xffffffffu, alive) & ~((2u << j) - 1u);
            }
            const int rank = kc + __popc(K & ((1u << lane) - 1u));
            if (((K >> lane) & 1u) && rank < P.max_det) {
                float* kj = gk + rank * 5;
                kj[0] = ox1; kj[1] = oy1; kj[2] = ox2; kj[3] = oy2; kj[4] = area;
                s_keep[i] = 1;
            }
            kc += __popc(K);
            __syncwarp();
        }
    }
    __syncthreads();

    // ---- 3. first max_det survivors in global (sorted) order
    for (int c0 = 0; c0 < n; c0 += kClsThreads) {
        const int i = c0 + tid;
        const bool k = i < n && s_keep[i];
        const unsigned m = __ballot_sync(0xffffffffu, k);
        if (lane == 0) s_wsum[warp] = __popc(m);
        __syncthreads();
        int pre = s_running, tot = 0;
        for (int w = 0; w < kClsThreads / 32; ++w) { const int v = s_wsum[w]; if (w < warp) pre += v; tot += v; }
        const int rank = pre + __popc(m & ((1u << lane) - 1u));
        if (k && rank < P.max_det) {
            const Cand cd = load_cand(P, kb, b, i);
            float* o = P.out + ((int64_t)b * P.max_det + rank) * 6;
            o[0] = cd.rx1; o[1] = cd.ry1; o[2] = cd.rx2; o[3] = cd.ry2; o[4] = cd.score; o[5] = cd.clsf;
            if (P.out_index) P.out_index[(int64_t)b * P.max_det + rank] = (int64_t)cd.idx;
        }
        __syncthreads();
        if (tid == 0) s_running += tot;
        __syncthreads();
        if (s_running >= P.max_det) break;
    }
    if (tid == 0) P.out_count[b] = s_running < P.max_det ? s_running : P.max_det;
}

NmsLayout nms_layout(int B, int nc, int A, int multi, int max_nms) {
    NmsLayout L{};
    const int64_t cand = multi ? (int64_t)A * nc : (int64_t)A;
    L.select = cand > max_nms;
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    size_t off = 0;
    L.counts = off; off = al(off + sizeof(int) * B);
    L.state = off; off = al(off + sizeof(SelectState) * B);
    L.hist = off; off = al(off + sizeof(unsigned) * 256 * B);
    if (L.select) {
        L.cap = (int)cand;
        L.key_stride = cand;
        L.cap2 = max_nms;
        uint32_t p2 = pow2ceil((uint32_t)max_nms);
        L.key2_stride = p2 <= (uint32_t)kSortChunk ? p2 : ceil_div(p2, kSortChunk) * kSortChunk;
    } else {
        L.cap = (int)cand;
        uint32_t p2 = pow2ceil((uint32_t)cand);
        L.key_stride = p2 <= (uint32_t)kSortChunk ? p2 : ceil_div(p2, kSortChunk) * kSortChunk;
        L.cap2 = 0; L.key2_stride = 0;
    }
    L.keys = off; off = al(off + sizeof(unsigned long long) * L.key_stride * B);
    L.keys2 = off; off = al(off + sizeof(unsigned long long) * L.key2_stride * B);
    const int sweep_cap = L.select ? L.cap2 : L.cap;  // candidates the sweep can see per image
    L.handled = off; off = al(off + sizeof(int) * B);
    L.gkept = off; off = al(off + sizeof(float) * 8 * (size_t)sweep_cap * B);  // kept boxes of the class-parallel sweep (5 floats per keep); global kept list of nms_sweep for max_det > kMaxDetSmem (8 per keep)
    L.total = off;
    return L;
}

// sort `B` segments of capacity `cap` (pow2-padded stride) in place
static void launch_sort(unsigned long long* keys, int64_t stride, const int* counts, const SelectState* st, int cap, int B, cudaStream_t s) {
    const uint32_t p2 = pow2ceil((uint32_t)(cap > 1 ? cap : 1));
    if (p2 <= (uint32_t)kSortChunk) {
        note_launches(1);
        size_t sm = (size_t)p2 * sizeof(unsigned long long);
        if (sm > 48 * 1024) cudaFuncSetAttribute(sort_single, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        sort_single<<<B, kSortThreads, sm, s>>>(keys, stride, counts, st, cap);
        return;
    }
    const size_t sm = (size_t)kSortChunk * sizeof(unsigned long long);
    cudaFuncSetAttribute(sort_chunks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    cudaFuncSetAttribute(sort_chunk_merge, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    const int chunks = (int)(p2 / kSortChunk);
    sort_chunks<<<dim3(chunks, B), kSortThreads, sm, s>>>(keys, stride, counts, st, cap);
    note_launches(1);
    for (uint32_t k = 2u * kSortChunk; k <= p2; k <<= 1) {
        for (uint32_t j = k >> 1; j >= (uint32_t)kSortChunk; j >>= 1) {
            int blocks = (int)(p2 / 2 / 256);
            if (blocks > kSMs * 8) blocks = kSMs * 8;
            sort_global_step<<<dim3(blocks, B), 256, 0, s>>>(keys, stride, counts, st, cap, (int)p2, (int)k, (int)j);
            note_launches(1);
        }
        sort_chunk_merge<<<dim3(chunks, B), kSortThreads, sm, s>>>(keys, stride, counts, st, cap, (int)k);
        note_launches(1);
    }
}

void sort_keys_desc(unsigned long long* keys, int64_t stride, const int* counts, int cap, int B, cudaStream_t s) { launch_sort(keys, stride, counts, nullptr, cap, B, s); }

static float threshold_round_down(double thr) {
    float f = (float)thr;
    if ((double)f > thr) f = nextafterf(f, -INFINITY);
    return f;
}

}  // namespace el

using namespace el;

extern "C" int el_nms_workspace_bytes(int B, int nc, int A, int multi_label, int max_nms, size_t* bytes) {
    if (!bytes || B <= 0 || nc <= 0 || A <= 0 || max_nms <= 0) return EL_ERR_ARG;
    if ((int64_t)A * nc >= (int64_t)1 << 31) return EL_ERR_UNSUPPORTED;
    *bytes = nms_layout(B, nc, A, multi_label && nc > 1, max_nms).total;
    return EL_OK;
}

namespace el {

void nms_prepare(const NmsLayout& L, void* ws, cudaStream_t s) { cudaMemsetAsync(ws, 0, L.keys, s); }  // counts, select state, histograms

int nms_finish(const NmsLayout& L, void* workspace, BoxSource src, int B, int nc, double iou, int agnostic, int max_det, int max_nms, float max_wh,
               int stages, float* out, int32_t* out_count, int64_t* out_index, cudaStream_t s) {
    char* ws = (char*)workspace;
    int* counts = (int*)(ws + L.counts);
    SelectState* st = (SelectState*)(ws + L.state);
    unsigned* hist = (unsigned*)(ws + L.hist);
    unsigned long long* keys = (unsigned long long*)(ws + L.keys);
    unsigned long long* keys2 = (unsigned long long*)(ws + L.keys2);
    SweepArgs P{};
    const bool do_sort = stages & 2, do_sweep = stages & 4;
    if (L.select && !do_sort) {
        P.keys = keys2; P.key_stride = L.key2_stride; P.counts = nullptr; P.st = st; P.cap = L.cap2;
    } else if (!do_sort) {
        P.keys = keys; P.key_stride = L.key_stride; P.counts = counts; P.st = nullptr; P.cap = L.cap;
    } else if (L.select) {
        init_select<<<(B + 127) / 128, 128, 0, s>>>(st, B, max_nms);
        int hb = (int)ceil_div(L.cap, 256 * 16);
        if (hb > 64) hb = 64;
        for (int shift = 56; shift >= 0; shift -= 8) {
            select_hist<<<dim3(hb, B), 256, 0, s>>>(keys, L.key_stride, counts, max_nms, st, hist, shift);
            select_pick<<<B, 256, 0, s>>>(counts, max_nms, st, hist, shift);
        }
        select_compact<<<dim3(hb, B), 256, 0, s>>>(keys, L.key_stride, counts, max_nms, st, keys2, L.key2_stride);
        note_launches(18);
        launch_sort(keys2, L.key2_stride, nullptr, st, L.cap2, B, s);
        P.keys = keys2; P.key_stride = L.key2_stride; P.counts = nullptr; P.st = st; P.cap = L.cap2;
    } else {
        launch_sort(keys, L.key_stride, counts, nullptr, L.cap, B, s);
        P.keys = keys; P.key_stride = L.key_stride; P.counts = counts; P.st = nullptr; P.cap = L.cap;
    }
    P.box = src.base; P.sb = src.sb; P.sa = src.sa; P.sk = src.sk;
    P.nc = nc; P.max_wh = max_wh; P.agnostic = agnostic;
    P.thr = threshold_round_down(iou); P.max_det = max_det;
    P.out = out; P.out_count = out_count; P.out_index = out_index;
    if (!do_sweep) return check_launch();
    // class-parallel sweep first (per-class independence needs class offsets and a class count that fits the CTA)
    const size_t sm_cls = (((size_t)4 * P.cap + 15) & ~(size_t)15) + sizeof(int) * (kClsThreads + nc + 1) + (size_t)P.cap + 16;
    const bool try_classes = !agnostic && nc > 1 && nc <= kClsThreads && P.cap <= 65535 && sm_cls <= 200 * 1024;
    int* handled = (int*)(ws + L.handled);
    if (try_classes) {
        if (sm_cls > 48 * 1024) cudaFuncSetAttribute(nms_sweep_classes, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_cls);
        nms_sweep_classes<<<B, kClsThreads, sm_cls, s>>>(P, (float*)(ws + L.gkept), handled);
        note_launches(1);
        P.handled = handled;
    }
    // images the class-parallel kernel declined (or all of them)
    const int kcap = max_det < P.cap ? max_det : P.cap;
    if (kcap <= kMaxDetSmem) {
        size_t sm = (size_t)8 * kcap * sizeof(float);  // kept boxes float4 + area, class tag, sorted index (+ one spare int per keep)
        if (sm > 48 * 1024) cudaFuncSetAttribute(nms_sweep<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        nms_sweep<true, true><<<B, kSweepThreads, sm, s>>>(P);
    } else {  // the reference accepts any max_det (ops.py:167): kept list in the workspace instead of shared memory
        P.gkept = (float*)(ws + L.gkept);
        nms_sweep<false, true><<<B, kSweepThreads, 0, s>>>(P);
    }
    note_launches(1);
    return check_launch();
}

}  // namespace el

extern "C" int el_nms_batched(const float* pred, int B, int nc, int A, float conf, double iou, int multi_label, int agnostic, const int32_t* class_keep,
                              int max_det, int max_nms, float max_wh, void* workspace, size_t workspace_bytes, float* out, int32_t* out_count,
                              int64_t* out_index, void* stream) {
    if (!pred || !workspace || !out || !out_count || B <= 0 || nc <= 0 || A <= 0 || max_det <= 0 || max_nms <= 0) return EL_ERR_ARG;
    if (conf < 0.f || conf > 1.f || iou < 0.0 || iou > 1.0) return EL_ERR_ARG;  // ops.py:217-218
    if ((int64_t)A * nc >= (int64_t)1 << 31) return EL_ERR_UNSUPPORTED;
    const bool multi = multi_label && nc > 1;  // ops.py:239
    const NmsLayout L = nms_layout(B, nc, A, multi, max_nms);
    if (workspace_bytes < L.total) return EL_ERR_WORKSPACE;
    cudaStream_t s = (cudaStream_t)stream;
    char* ws = (char*)workspace;
    nms_prepare(L, ws, s);
    dim3 eg((A + 255) / 256, B);
    if (multi)
        nms_emit<true><<<eg, 256, 0, s>>>(pred, nc, A, conf, class_keep, (unsigned long long*)(ws + L.keys), L.key_stride, (int*)(ws + L.counts));
    else
        nms_emit<false><<<eg, 256, 0, s>>>(pred, nc, A, conf, class_keep, (unsigned long long*)(ws + L.keys), L.key_stride, (int*)(ws + L.counts));
    note_launches(1);
    return nms_finish(L, ws, BoxSource{pred, (int64_t)(4 + nc) * A, 1, A}, B, nc, iou, agnostic, max_det, max_nms, max_wh, 7, out, out_count, out_index, s);
}

extern "C" int el_nms_boxes_workspace_bytes(int n, size_t* bytes) {
    if (!bytes || n < 0) return EL_ERR_ARG;
    uint32_t p2 = pow2ceil((uint32_t)(n > 1 ? n : 1));
    int64_t stride = p2 <= (uint32_t)kSortChunk ? p2 : ceil_div(p2, kSortChunk) * kSortChunk;
    *bytes = 256 + sizeof(unsigned long long) * stride + sizeof(float) * 5 * (size_t)(n > 0 ? n : 1) + 256;
    return EL_OK;
}

extern "C" int el_nms_boxes(const float* boxes, const float* scores, int n, double iou, void* workspace, size_t workspace_bytes, int64_t* keep,
                            int32_t* keep_count, void* stream) {
    if (!keep_count || n < 0 || iou < 0.0 || iou > 1.0) return EL_ERR_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    if (n == 0) { cudaMemsetAsync(keep_count, 0, sizeof(int32_t), s); return check_launch(); }
    if (!boxes || !scores || !workspace || !keep) return EL_ERR_ARG;
    if (!aligned16(boxes)) return EL_ERR_ARG;
    size_t need;
    el_nms_boxes_workspace_bytes(n, &need);
    if (workspace_bytes < need) return EL_ERR_WORKSPACE;
    uint32_t p2 = pow2ceil((uint32_t)(n > 1 ? n : 1));
    int64_t stride = p2 <= (uint32_t)kSortChunk ? p2 : ceil_div(p2, kSortChunk) * kSortChunk;
    char* ws = (char*)workspace;
    int* count = (int*)ws;
    unsigned long long* keys = (unsigned long long*)(ws + 256);
    float* gkept = (float*)(ws + 256 + sizeof(unsigned long long) * stride);
    nms_box_keys<<<(n + 255) / 256, 256, 0, s>>>(scores, n, keys, count);
    launch_sort(keys, stride, count, nullptr, n, 1, s);
    SweepArgs P{};
    P.keys = keys; P.key_stride = stride; P.counts = count; P.st = nullptr; P.cap = n;
    P.box = boxes; P.thr = threshold_round_down(iou); P.max_det = n;
    P.out_count = keep_count; P.keep = keep; P.gkept = gkept;
    nms_sweep<false, false><<<1, kSweepThreads, 0, s>>>(P);
    note_launches(2);  // keys + sweep
    return check_launch();
}
