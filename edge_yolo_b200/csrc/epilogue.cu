// Streaming glue kernels around the cuDNN convolutions of the EdgeLine graph (inference engine only):
//   el_bias_act_fwd      out = act(x + bias[c]) [+ res] -- the BatchNorm-folded bias + SiLU of Conv / DSConv
//                                                        (nn/modules/conv.py:41-60, 87-104) in ONE pass instead of
//                                                        PyTorch's separate broadcast-add and SiLU kernels
//   el_upsample2x_cat    out = cat[nearest2x(x), skip] -- nn.Upsample + Concat pairs of the neck
//                                                        (cfg/models/11/yolo11-test.yaml:34-39) in one pass
// Both are pure HBM-bound elementwise kernels over channel-contiguous (NHWC) views with arbitrary pitches, so the
// result can be written straight into a slice of a pre-allocated concat buffer.
#include "el_internal.h"

namespace el {

template <typename T> __device__ __forceinline__ float silu_f(float v) {
    if constexpr (sizeof(T) == 2) {  // x * sigmoid(x) = h + h * tanh(h), h = x / 2: ONE transcendental (MUFU.TANH) instead of ex2 + rcp;
        const float h = 0.5f * v;    // its ~2^-11 error disappears in the rounding to 16 bits
        float t;
        asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
        return fmaf(h, t, h);
    } else {
        return v / (1.f + expf(-v));
    }
}

constexpr int kEpRowsMax = 4;

// thread <-> (channel vector, column) fixed; walks kEpRows rows.  grid = (col tiles, row chunks, images)
template <typename T, int ACT>
__global__ void __launch_bounds__(256) bias_act_tiled(const T* x, Strides4 xs, const float* __restrict__ bias, const T* res, Strides4 rs, T* o,
                                                      Strides4 os, T* o2, Strides4 os2, int split_cv, int CV, int cols_per_block, int rows, int H, int W) {
    constexpr int V = Vec16<T>::N;
    pdl_launch_dependents();
    const int xi = (int)threadIdx.x / CV, cv = (int)threadIdx.x - xi * CV, col = (int)blockIdx.x * cols_per_block + xi;
    if (xi >= cols_per_block || col >= W) return;
    float bv[V];
#pragma unroll
    for (int e = 0; e < V; ++e) bv[e] = bias ? __ldg(bias + cv * V + e) : 0.f;
    const int r0 = (int)blockIdx.y * rows, r1 = min(r0 + rows, H);
    const int64_t n = blockIdx.z;
    const T* p = x + n * xs.n + (int64_t)r0 * xs.h + (int64_t)col * xs.w + cv * V;
    T* q = o + n * os.n + (int64_t)r0 * os.h + (int64_t)col * os.w + cv * V;
    int64_t qh = os.h;
    if (o2 && cv >= split_cv) {  // channels >= split go to the second destination (e.g. a dense tensor next to a concat slice)
        q = o2 + n * os2.n + (int64_t)r0 * os2.h + (int64_t)col * os2.w + (cv - split_cv) * V;
        qh = os2.h;
    }
    const T* pr = res ? res + n * rs.n + (int64_t)r0 * rs.h + (int64_t)col * rs.w + cv * V : nullptr;
#pragma unroll
    for (int r = 0; r < kEpRowsMax; ++r) {
        if (r0 + r >= r1) break;
        float f[V];
        unpack<T>(*reinterpret_cast<const uint4*>(p + (int64_t)r * xs.h), f);  // x may alias o
#pragma unroll
        for (int e = 0; e < V; ++e) {
            float v = f[e] + bv[e];
            f[e] = ACT == 1 ? silu_f<T>(v) : (ACT == 2 ? fmaxf(v, 0.f) : v);
        }
        if (pr) {
            float g[V];
            unpack<T>(*reinterpret_cast<const uint4*>(pr + (int64_t)r * rs.h), g);
#pragma unroll
            for (int e = 0; e < V; ++e) f[e] += g[e];
        }
        *reinterpret_cast<uint4*>(q + (int64_t)r * qh) = pack<T>(f);
    }
}

// flat variant for small feature maps / very wide channel counts: one vector per thread, grid-stride
template <typename T, int ACT>
__global__ void __launch_bounds__(256) bias_act_flat(const T* x, Strides4 xs, const float* __restrict__ bias, const T* res, Strides4 rs, T* o, Strides4 os,
                                                     T* o2, Strides4 os2, int split_cv, int CV, int H, int W, uint32_t total) {
    constexpr int V = Vec16<T>::N;
    pdl_launch_dependents();
    for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        uint32_t cv = idx % (uint32_t)CV, t = idx / (uint32_t)CV;
        uint32_t col = t % (uint32_t)W; t /= (uint32_t)W;
        uint32_t row = t % (uint32_t)H, n = t / (uint32_t)H;
        float f[V];
        unpack<T>(*reinterpret_cast<const uint4*>(x + (int64_t)n * xs.n + (int64_t)row * xs.h + (int64_t)col * xs.w + cv * V), f);
#pragma unroll
        for (int e = 0; e < V; ++e) {
            float v = f[e] + (bias ? __ldg(bias + cv * V + e) : 0.f);
            f[e] = ACT == 1 ? silu_f<T>(v) : (ACT == 2 ? fmaxf(v, 0.f) : v);
        }
        if (res) {
            float g[V];
            unpack<T>(*reinterpret_cast<const uint4*>(res + (int64_t)n * rs.n + (int64_t)row * rs.h + (int64_t)col * rs.w + cv * V), g);
#pragma unroll
            for (int e = 0; e < V; ++e) f[e] += g[e];
        }
        if (o2 && (int)cv >= split_cv)
            *reinterpret_cast<uint4*>(o2 + (int64_t)n * os2.n + (int64_t)row * os2.h + (int64_t)col * os2.w + ((int)cv - split_cv) * V) = pack<T>(f);
        else
            *reinterpret_cast<uint4*>(o + (int64_t)n * os.n + (int64_t)row * os.h + (int64_t)col * os.w + cv * V) = pack<T>(f);
    }
}

template <typename T, int ACT>
__global__ void __launch_bounds__(256) bias_act_generic(const T* x, Strides4 xs, const float* __restrict__ bias, const T* res, Strides4 rs, T* o, Strides4 os,
                                                        T* o2, Strides4 os2, int split, int C, int H, int W, int64_t total, bool ch_fast) {
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        int64_t t = idx, n; int c, row, col;
        if (ch_fast) { c = (int)(t % C); t /= C; col = (int)(t % W); t /= W; row = (int)(t % H); n = t / H; }
        else { col = (int)(t % W); t /= W; row = (int)(t % H); t /= H; c = (int)(t % C); n = t / C; }
        float v = to_f(x[n * xs.n + (int64_t)c * xs.c + (int64_t)row * xs.h + (int64_t)col * xs.w]) + (bias ? __ldg(bias + c) : 0.f);
        v = ACT == 1 ? silu_f<T>(v) : (ACT == 2 ? fmaxf(v, 0.f) : v);
        if (res) v += to_f(res[n * rs.n + (int64_t)c * rs.c + (int64_t)row * rs.h + (int64_t)col * rs.w]);
        if (o2 && c >= split) o2[n * os2.n + (int64_t)(c - split) * os2.c + (int64_t)row * os2.h + (int64_t)col * os2.w] = from_f<T>(v);
        else o[n * os.n + (int64_t)c * os.c + (int64_t)row * os.h + (int64_t)col * os.w] = from_f<T>(v);
    }
}

// out[:, :C1] = nearest-2x(x), out[:, C1:] = skip; thread <-> (channel vector of out, column of out)
template <typename T>
__global__ void __launch_bounds__(256) upsample2x_cat_tiled(const T* __restrict__ x, Strides4 xs, const T* __restrict__ skip, Strides4 ss, T* __restrict__ o,
                                                            Strides4 os, int CV1, int CV, int cols_per_block, int H, int W) {
    constexpr int V = Vec16<T>::N;
    pdl_launch_dependents();
    pdl_wait();
    const int xi = (int)threadIdx.x / CV, cv = (int)threadIdx.x - xi * CV, col = (int)blockIdx.x * cols_per_block + xi;
    if (xi >= cols_per_block || col >= W) return;
    const int r0 = (int)blockIdx.y * kEpRowsMax, r1 = min(r0 + kEpRowsMax, H);
    const int64_t n = blockIdx.z;
    T* q = o + n * os.n + (int64_t)col * os.w + cv * V;
    if (cv < CV1) {
        const T* p = x + n * xs.n + (int64_t)(col >> 1) * xs.w + cv * V;
        for (int r = r0; r < r1; ++r) stg_stream(q + (int64_t)r * os.h, ldg_cached(p + (int64_t)(r >> 1) * xs.h));  // each source pixel is read 4x: L1
    } else {
        const T* p = skip + n * ss.n + (int64_t)col * ss.w + (cv - CV1) * V;
        for (int r = r0; r < r1; ++r) stg_stream(q + (int64_t)r * os.h, ldg_stream(p + (int64_t)r * ss.h));
    }
}


// SPPF pooling pyramid (nn/modules/block.py:204-223): out = cat[x, m(x), m(m(x)), m(m(m(x)))] with m = MaxPool2d(5, 1, 2).
// Chained 5x5 max-pools with -inf padding are exactly the clipped 5x5 / 9x9 / 13x13 window maxima, and max is separable,
// so one CTA per (image, channel vector) stages the map in shared memory, does the three row passes, then the three column
// passes, and writes all four concat slices.
template <typename T> __device__ __forceinline__ uint4 vmax16(uint4 a, uint4 b);
template <> __device__ __forceinline__ uint4 vmax16<float>(uint4 a, uint4 b) {
    return make_uint4(__float_as_uint(fmaxf(__uint_as_float(a.x), __uint_as_float(b.x))), __float_as_uint(fmaxf(__uint_as_float(a.y), __uint_as_float(b.y))),
                      __float_as_uint(fmaxf(__uint_as_float(a.z), __uint_as_float(b.z))), __float_as_uint(fmaxf(__uint_as_float(a.w), __uint_as_float(b.w))));
}
template <> __device__ __forceinline__ uint4 vmax16<__nv_bfloat16>(uint4 a, uint4 b) {
    uint4 r;
    const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&a);
    const __nv_bfloat162* pb = reinterpret_cast<const __nv_bfloat162*>(&b);
    __nv_bfloat162* pr = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) pr[i] = __hmax2(pa[i], pb[i]);
    return r;
}
template <> __device__ __forceinline__ uint4 vmax16<__half>(uint4 a, uint4 b) {
    uint4 r;
    const __half2* pa = reinterpret_cast<const __half2*>(&a);
    const __half2* pb = reinterpret_cast<const __half2*>(&b);
    __half2* pr = reinterpret_cast<__half2*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) pr[i] = __hmax2(pa[i], pb[i]);
    return r;
}

// CVL channel vectors per CTA: a pixel's CVL * 16 bytes are contiguous, so with CVL = 2 every global access is a whole 32-byte sector (one vector
// per CTA read and wrote half sectors: 26 us for 33 MB on the 20 x 20 map of EdgeLine-n); items = (pixel, vector), vector fastest.
template <typename T, int CVL>
__global__ void __launch_bounds__(256) sppf_pool_kernel(const T* __restrict__ x, Strides4 xs, T* __restrict__ o, Strides4 os, int C, int H, int W) {
    constexpr int V = Vec16<T>::N;
    extern __shared__ uint4 s_tiles[];  // [4][H*W*CVL]: input, row-max r=2, r=4, r=6
    const int HW = H * W, NI = HW * CVL, cv0 = blockIdx.x * CVL;
    const int64_t n = blockIdx.y;
    uint4* s_in = s_tiles;
    uint4* s_h[3] = {s_tiles + NI, s_tiles + 2 * NI, s_tiles + 3 * NI};
    pdl_launch_dependents();
    pdl_wait();
    for (int i = threadIdx.x; i < NI; i += blockDim.x) {
        const int p = i / CVL, c = i - p * CVL, yy = p / W, xx = p - yy * W;
        s_in[i] = ldg_stream(x + n * xs.n + (int64_t)yy * xs.h + (int64_t)xx * xs.w + (cv0 + c) * V);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < NI; i += blockDim.x) {  // row pass: windows |dx| <= 2, 4, 6 (nested)
        const int p = i / CVL, yy = p / W, xx = p - yy * W;
        uint4 m = s_in[i];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
#pragma unroll
            for (int d = 2 * r + 1; d <= 2 * r + 2; ++d) {
                if (xx - d >= 0) m = vmax16<T>(m, s_in[i - d * CVL]);
                if (xx + d < W) m = vmax16<T>(m, s_in[i + d * CVL]);
            }
            s_h[r][i] = m;
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < NI; i += blockDim.x) {  // column pass on the matching row-max tile
        const int p = i / CVL, c = i - p * CVL, yy = p / W, xx = p - yy * W;
        T* q = o + n * os.n + (int64_t)yy * os.h + (int64_t)xx * os.w + (cv0 + c) * V;
        *reinterpret_cast<uint4*>(q) = s_in[i];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            uint4 m = s_h[r][i];
            for (int d = 1; d <= 2 * r + 2; ++d) {
                if (yy - d >= 0) m = vmax16<T>(m, s_h[r][i - d * W * CVL]);
                if (yy + d < H) m = vmax16<T>(m, s_h[r][i + d * W * CVL]);
            }
            *reinterpret_cast<uint4*>(q + (int64_t)(r + 1) * C) = m;
        }
    }
}


// Stem: uint8 HWC image -> Conv(3, C0, k=3, s=2, p=1) + folded BatchNorm + SiLU -> NHWC activations, in one pass.
// Replaces, for the inference engine, the uint8->float preprocess (engine/predictor.py:117-135), layer 0 of the yaml
// (cfg/models/11/yolo11-test.yaml:21, a cuDNN conv that PyTorch runs through an fp32 NHWC round trip for 3 input channels)
// and its bias + SiLU epilogue.  The 1/255 scale is folded into the weights by the caller.
// CUDA-core FMA kernel balanced against the shared-memory pipe: a thread computes 8 adjacent output pixels x 8 output
// channels, so every 16-byte weight broadcast (LDS.128 = 4 LSU cycles per warp) feeds 32 FMAs per lane; the 17 x 129 pixel
// input patch of an 8 x 64 output tile is staged once per CTA as fp32 (uint8 -> float without the conversion pipe: 0x4B000000 | b).
constexpr int kStemTW = 64, kStemTH = 8, kStemIW = 2 * kStemTW + 1, kStemIH = 2 * kStemTH + 1, kStemRow = kStemIW * 3 + 1;  // 388 floats per patch row

template <typename T, int C0>
__global__ void __launch_bounds__(128) stem_conv_u8_kernel(const uint8_t* __restrict__ src, const float* __restrict__ w, const float* __restrict__ bias,
                                                           T* __restrict__ dst, Strides4 ds, int H, int W) {
    extern __shared__ __align__(16) float s_stem[];
    float* s_w = s_stem;                       // [tap = (ky*3+kx)*3+ci][co]
    float* s_b = s_w + 27 * C0;                // [co]
    float* s_in = s_b + C0;                    // [kStemIH][kStemRow]
    const int tid = threadIdx.x;
    const int ox0 = blockIdx.x * kStemTW, oy0 = blockIdx.y * kStemTH;
    const int64_t n = blockIdx.z;
    for (int i = tid; i < 27 * C0; i += 128) {  // caller layout (C0, 3, 3, 3) = [co][ci][ky][kx] -> [tap][co]
        const int co = i / 27, r = i - co * 27, ci = r / 9, k = r - ci * 9;
        s_w[(k * 3 + ci) * C0 + co] = __ldg(w + i);
    }
    if (tid < C0) s_b[tid] = __ldg(bias + tid);
    const uint8_t* img = src + n * (int64_t)H * W * 3;
    const int iy0 = 2 * oy0 - 1;
    // patch rows are staged with aligned 32-bit loads (coalesced; W % 4 == 0 so a word never straddles an image row):
    // byte range [seg0, seg0 + 387) of each image row, seg0 = (2*ox0 - 1) * 3 (negative at the left edge = zero padding)
    const int seg0 = (2 * ox0 - 1) * 3, row_bytes = W * 3;
    const int w_lo = (seg0 - (seg0 & 3)) >> 2;  // floor(seg0 / 4), also for seg0 = -3
    constexpr int NW = (kStemIW * 3 + 3) / 4 + 1;  // words that cover 387 bytes at any alignment
    for (int i = tid; i < kStemIH * NW; i += 128) {
        const int r = i / NW, wi = i - r * NW;
        const int iy = iy0 + r, byte0 = 4 * (w_lo + wi);
        uint32_t u = 0;
        if (iy >= 0 && iy < H && byte0 >= 0 && byte0 < row_bytes) u = __ldg(reinterpret_cast<const uint32_t*>(img + (int64_t)iy * row_bytes + byte0));
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int cb = byte0 + k - seg0;
            if (cb >= 0 && cb < kStemIW * 3) s_in[r * kStemRow + cb] = __uint_as_float(0x4B000000u | ((u >> (8 * k)) & 0xffu)) - 8388608.f;
        }
    }
    __syncthreads();
    // thread <-> (half of a 16-channel block, group of 8 adjacent output pixels, row): 64 fp32 accumulators; every 16-byte weight
    // broadcast feeds 8 pixels x 4 channels = 32 FMAs and every input float 8 channels, which keeps the shared-memory pipe
    // (LDS.128 = 4 cycles per warp) below the FMA pipe
    const int half = tid & 1, cg = (tid >> 1) & 7, ty = tid >> 4;
    const int ox = ox0 + 8 * cg, oy = oy0 + ty;
    if (ox >= W / 2 || oy >= H / 2) return;
    constexpr int V = Vec16<T>::N;
#pragma unroll 1
    for (int cb = 8 * half; cb < C0; cb += 16) {
        float acc[8][8];
#pragma unroll
        for (int p = 0; p < 8; ++p)
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[p][c] = s_b[cb + c];
#pragma unroll 1
        for (int ky = 0; ky < 3; ++ky) {
            float v[51];  // 17 input pixels x 3 channels of patch row 2*ty + ky, starting at pixel 16*cg
            const float* row = s_in + (2 * ty + ky) * kStemRow + 48 * cg;
#pragma unroll
            for (int t = 0; t < 48; t += 4) {
                const float4 q = *reinterpret_cast<const float4*>(row + t);
                v[t] = q.x; v[t + 1] = q.y; v[t + 2] = q.z; v[t + 3] = q.w;
            }
            v[48] = row[48]; v[49] = row[49]; v[50] = row[50];
#pragma unroll
            for (int t = 0; t < 9; ++t) {  // t = kx*3 + ci
                const float4* wp = reinterpret_cast<const float4*>(s_w + (ky * 9 + t) * C0 + cb);
                const float4 wa = wp[0], wb = wp[1];  // two addresses per warp (the channel halves): broadcast
#pragma unroll
                for (int p = 0; p < 8; ++p) {
                    const float x = v[6 * p + t];
                    acc[p][0] = fmaf(x, wa.x, acc[p][0]); acc[p][1] = fmaf(x, wa.y, acc[p][1]);
                    acc[p][2] = fmaf(x, wa.z, acc[p][2]); acc[p][3] = fmaf(x, wa.w, acc[p][3]);
                    acc[p][4] = fmaf(x, wb.x, acc[p][4]); acc[p][5] = fmaf(x, wb.y, acc[p][5]);
                    acc[p][6] = fmaf(x, wb.z, acc[p][6]); acc[p][7] = fmaf(x, wb.w, acc[p][7]);
                }
            }
        }
#pragma unroll
        for (int p = 0; p < 8; ++p) {
            if (ox + p >= W / 2) break;
            T* q = dst + n * ds.n + (int64_t)oy * ds.h + (int64_t)(ox + p) * ds.w + cb;
#pragma unroll
            for (int g = 0; g < 8 / V; ++g) {
                float f[V];
#pragma unroll
                for (int e = 0; e < V; ++e) f[e] = silu_f<T>(acc[p][g * V + e]);
                *reinterpret_cast<uint4*>(q + g * V) = pack<T>(f);
            }
        }
    }
}

}  // namespace el

using namespace el;

extern "C" int el_bias_act_fwd(const void* x, const int64_t xs_[4], const float* bias, const void* residual, const int64_t rs_[4], void* out,
                               const int64_t os_[4], void* out2, const int64_t os2_[4], int split, int B, int C, int H, int W, int act, int dtype,
                               void* stream) {
    if (!x || !out || B <= 0 || C <= 0 || H <= 0 || W <= 0 || act < 0 || act > 2 || (residual && !rs_)) return EL_ERR_ARG;
    if (out2 && (!os2_ || split <= 0 || split >= C)) return EL_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    Strides4 xs = s4(xs_), os = s4(os_), rs = residual ? s4(rs_) : Strides4{0, 0, 0, 0}, os2 = out2 ? s4(os2_) : Strides4{0, 0, 0, 0};
    const int C1 = out2 ? split : C;  // channels landing in `out`
#define EL_BIAS_ACT(KERNEL, ...)                                                           \
    do {                                                                                   \
        if (act == 1) KERNEL<T, 1> __VA_ARGS__;                                            \
        else if (act == 2) KERNEL<T, 2> __VA_ARGS__;                                       \
        else KERNEL<T, 0> __VA_ARGS__;                                                     \
    } while (0)
    EL_DISPATCH_DTYPE(dtype, {
        constexpr int V = Vec16<T>::N;
        if (channel_vectorisable<T>(x, xs, C) && channel_vectorisable<T>(out, os, C1) && (!residual || channel_vectorisable<T>(residual, rs, C)) &&
            (!out2 || channel_vectorisable<T>(out2, os2, C - split))) {
            const int CV = C / V;
            const int64_t total = (int64_t)B * H * W * CV;
            if (CV <= 256 && B <= 65535) {
                    const int cpb = 256 / CV < W ? 256 / CV : W;
                int rows = kEpRowsMax;
                while (rows > 1 && ceil_div(W, cpb) * ceil_div(H, rows) * B < (int64_t)kSMs * 16) rows >>= 1;
                dim3 g((unsigned)ceil_div(W, cpb), (unsigned)ceil_div(H, rows), (unsigned)B);
                EL_BIAS_ACT(bias_act_tiled, <<<g, 256, 0, st>>>((const T*)x, xs, bias, (const T*)residual, rs, (T*)out, os, (T*)out2, os2, split / V, CV, cpb, rows, H, W));
            } else if (total < ((int64_t)1 << 32)) {
                int grid = (int)(ceil_div(total, 256) < (int64_t)kSMs * 16 ? ceil_div(total, 256) : (int64_t)kSMs * 16);
                EL_BIAS_ACT(bias_act_flat, <<<grid, 256, 0, st>>>((const T*)x, xs, bias, (const T*)residual, rs, (T*)out, os, (T*)out2, os2, split / V, CV, H, W, (uint32_t)total));
            } else {
                return EL_ERR_UNSUPPORTED;
            }
        } else {
            const int64_t total = (int64_t)B * C * H * W;
            int grid = (int)(ceil_div(total, 256) < (int64_t)kSMs * 16 ? ceil_div(total, 256) : (int64_t)kSMs * 16);
            EL_BIAS_ACT(bias_act_generic, <<<grid, 256, 0, st>>>((const T*)x, xs, bias, (const T*)residual, rs, (T*)out, os, (T*)out2, os2, split, C, H, W, total, os.c == 1));
        }
    });
#undef EL_BIAS_ACT
    note_launches(1);
    return check_launch();
}

extern "C" int el_upsample2x_cat_fwd(const void* x, const int64_t xs_[4], const void* skip, const int64_t ss_[4], void* out, const int64_t os_[4], int B,
                                     int C1, int C2, int H, int W, int dtype, void* stream) {
    if (!x || !skip || !out || B <= 0 || C1 <= 0 || C2 <= 0 || H <= 0 || W <= 0 || (H & 1) || (W & 1)) return EL_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    Strides4 xs = s4(xs_), ss = s4(ss_), os = s4(os_);
    EL_DISPATCH_DTYPE(dtype, {
        constexpr int V = Vec16<T>::N;
        if (!(channel_vectorisable<T>(x, xs, C1) && channel_vectorisable<T>(skip, ss, C2) && channel_vectorisable<T>(out, os, C1 + C2)) ||
            (C1 + C2) / V > 256 || B > 65535)
            return EL_ERR_UNSUPPORTED;  // the engine only uses this on NHWC activations
        const int CV = (C1 + C2) / V, cpb = 256 / CV;
        dim3 g((unsigned)ceil_div(W, cpb), (unsigned)ceil_div(H, kEpRowsMax), (unsigned)B);
        launch_pdl(upsample2x_cat_tiled<T>, g, dim3(256), 0, st, (const T*)x, xs, (const T*)skip, ss, (T*)out, os, C1 / V, CV, cpb, H, W);
    });
    note_launches(1);
    return check_launch();
}

extern "C" int el_sppf_pool_fwd(const void* x, const int64_t xs_[4], void* out, const int64_t os_[4], int B, int C, int H, int W, int dtype, void* stream) {
    if (!x || !out || B <= 0 || C <= 0 || H <= 0 || W <= 0) return EL_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    Strides4 xs = s4(xs_), os = s4(os_);
    if (B > 65535) return EL_ERR_UNSUPPORTED;
    cudaError_t e = cudaSuccess;
    EL_DISPATCH_DTYPE(dtype, {
        constexpr int V = Vec16<T>::N;
        if (!(channel_vectorisable<T>(x, xs, C) && channel_vectorisable<T>(out, os, 4 * C))) return EL_ERR_UNSUPPORTED;
        // two channel vectors per CTA (whole 32-byte sectors) when the channel count allows it and four CTAs of that size still fit an SM
        const int cvl = ((C / V) % 2 == 0 && (size_t)8 * H * W * sizeof(uint4) <= 56 * 1024) ? 2 : 1;
        const size_t sm = (size_t)4 * cvl * H * W * sizeof(uint4);
        if (sm > 200 * 1024) return EL_ERR_UNSUPPORTED;  // maps above ~56x56: callers keep nn.MaxPool2d
        auto kern = cvl == 2 ? sppf_pool_kernel<T, 2> : sppf_pool_kernel<T, 1>;
        if (sm > 48 * 1024) e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        if (e == cudaSuccess) e = launch_pdl(kern, dim3(C / V / cvl, B), dim3(256), sm, st, (const T*)x, xs, (T*)out, os, C, H, W);
    });
    if (e != cudaSuccess) { g_last_cuda_error = (int)e; return EL_ERR_CUDA; }
    note_launches(1);
    return check_launch();
}

extern "C" int el_stem_conv_u8(const uint8_t* src, const float* w, const float* bias, void* dst, const int64_t ds_[4], int B, int C0, int H, int W, int dtype,
                               void* stream) {
    if (!src || !w || !bias || !dst || B <= 0 || H <= 0 || W <= 0 || (H & 1) || (W & 3) || (reinterpret_cast<uintptr_t>(src) & 3)) return EL_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    Strides4 ds = s4(ds_);
    if (B > 65535) return EL_ERR_UNSUPPORTED;
    if (dtype == EL_BF16 || dtype == EL_F16) {  // 16-bit activations: im2col GEMM on the tensor cores (stem_tc.cu)
        if (ds.c != 1 || ds.n % 8 || ds.h % 8 || ds.w % 8 || !aligned16(dst) || (C0 != 16 && C0 != 32 && C0 != 64)) return EL_ERR_UNSUPPORTED;
        const int rc = stem_tc_launch(src, w, bias, dst, ds, B, C0, H, W, dtype, st);
        if (rc != EL_OK) return rc;
        note_launches(1);
        return check_launch();
    }
    dim3 grid((unsigned)ceil_div(W / 2, kStemTW), (unsigned)ceil_div(H / 2, kStemTH), (unsigned)B);
    const size_t smem = (size_t)(28 * C0 + kStemIH * kStemRow) * sizeof(float);
    EL_DISPATCH_DTYPE(dtype, {
        if (!channel_vectorisable<T>(dst, ds, C0)) return EL_ERR_UNSUPPORTED;
        if (C0 == 16) stem_conv_u8_kernel<T, 16><<<grid, 128, smem, st>>>(src, w, bias, (T*)dst, ds, H, W);
        else if (C0 == 32) stem_conv_u8_kernel<T, 32><<<grid, 128, smem, st>>>(src, w, bias, (T*)dst, ds, H, W);
        else if (C0 == 64) stem_conv_u8_kernel<T, 64><<<grid, 128, smem, st>>>(src, w, bias, (T*)dst, ds, H, W);
        else return EL_ERR_UNSUPPORTED;
    });
    note_launches(1);
    return check_launch();
}
