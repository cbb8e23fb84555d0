// Internal (non-ABI) interfaces shared between translation units of libedgeline_b200.so.
#pragma once
#include "el_common.cuh"

namespace el {

struct SelectState {            // per image, radix-select state of the max_nms cut
    unsigned long long prefix;  // high bits of the k-th largest key found so far
    int k_rem;                  // rank still to resolve inside the current prefix
    int count2;                 // compacted count
};

struct NmsLayout {  // offsets into the caller's workspace
    size_t counts, state, hist, keys, keys2, handled, gkept, total;
    int64_t key_stride, key2_stride;
    int cap, cap2;
    bool select;
};
NmsLayout nms_layout(int B, int nc, int A, int multi, int max_nms);

// where the sweep reads xywh boxes: component k of anchor a of image b at base[b*sb + a*sa + k*sk]
struct BoxSource { const float* base; int64_t sb, sa, sk; };

// zero counters / select state (call before emitting keys)
void nms_prepare(const NmsLayout& L, void* ws, cudaStream_t s);
// keys + counts are in the workspace: max_nms cut + sort (stages bit 1), greedy sweep (stages bit 2)
int nms_finish(const NmsLayout& L, void* ws, BoxSource src, int B, int nc, double iou, int agnostic, int max_det, int max_nms, float max_wh,
               int stages, float* out, int32_t* out_count, int64_t* out_index, cudaStream_t s);

// tensor-core path of el_stem_conv_u8 for 16-bit activations (stem_tc.cu)
int stem_tc_launch(const uint8_t* src, const float* w, const float* bias, void* dst, Strides4 ds, int B, int C0, int H, int W, int dtype, cudaStream_t st);

constexpr int kMaxDetSmem = 4096;

__host__ __device__ inline uint32_t pow2ceil(uint32_t v) {
    uint32_t p = 1;
    while (p < v) p <<= 1;
    return p;
}

// descending in-place sort of B key segments (count per segment on the device, pow2 / 16384-multiple padded stride): nms.cu
void sort_keys_desc(unsigned long long* keys, int64_t stride, const int* counts, int cap, int B, cudaStream_t s);

}  // namespace el
