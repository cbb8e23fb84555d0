// Streaming glue kernels around the cuDNN convolutions of the EdgeLine graph (inference engine only):
//   el_bias_act_fwd      out = act(x + bias[c]) [+ res] -- the BatchNorm-folded bias + SiLU of Conv / DSConv
//                                                        (nn/modules/conv.py:41-60, 87-104) in ONE pass instead of
//                                                        PyTorch's separate broadcast-add and SiLU kernels
//   el_upsample2x_cat    out = cat[nearest2x(x), skip] -- nn.Upsample + Concat pairs of the neck
//                                                        (cfg/models/11/yolo11-test.yaml:34-39) in one pass
// Both are pure HBM-bound elementwise kernels over channel-contiguous (NHWC) views with arbitrary pitches, so the
// result can be written straight into a slice of a pre-allocated concat buffer.
#include "el_common.cuh"

namespace el {

template <typename T> __device__ __forceinline__ float silu_f(float v) {
    if constexpr (sizeof(T) == 2) return __fdividef(v, 1.f + __expf(-v));  // rounding to 16 bits hides the fast-math error
    else return v / (1.f + expf(-v));
}

constexpr int kEpRows = 4;

// thread <-> (channel vector, column) fixed; walks kEpRows rows.  grid = (col tiles, row chunks, images)
template <typename T, int ACT>
__global__ void __launch_bounds__(256) bias_act_tiled(const T* x, Strides4 xs, const float* __restrict__ bias, const T* res, Strides4 rs, T* o,
                                                      Strides4 os, int CV, int cols_per_block, int H, int W) {
    constexpr int V = Vec16<T>::N;
    const int xi = (int)threadIdx.x / CV, cv = (int)threadIdx.x - xi * CV, col = (int)blockIdx.x * cols_per_block + xi;
    if (xi >= cols_per_block || col >= W) return;
    float bv[V];
#pragma unroll
    for (int e = 0; e < V; ++e) bv[e] = bias ? __ldg(bias + cv * V + e) : 0.f;
    const int r0 = (int)blockIdx.y * kEpRows, r1 = min(r0 + kEpRows, H);
    const int64_t n = blockIdx.z;
    const T* p = x + n * xs.n + (int64_t)r0 * xs.h + (int64_t)col * xs.w + cv * V;
    T* q = o + n * os.n + (int64_t)r0 * os.h + (int64_t)col * os.w + cv * V;
    const T* pr = res ? res + n * rs.n + (int64_t)r0 * rs.h + (int64_t)col * rs.w + cv * V : nullptr;
#pragma unroll
    for (int r = 0; r < kEpRows; ++r) {
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
        *reinterpret_cast<uint4*>(q + (int64_t)r * os.h) = pack<T>(f);
    }
}

// flat variant for small feature maps / very wide channel counts: one vector per thread, grid-stride
template <typename T, int ACT>
__global__ void __launch_bounds__(256) bias_act_flat(const T* x, Strides4 xs, const float* __restrict__ bias, const T* res, Strides4 rs, T* o, Strides4 os,
                                                     int CV, int H, int W, uint32_t total) {
    constexpr int V = Vec16<T>::N;
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
        *reinterpret_cast<uint4*>(o + (int64_t)n * os.n + (int64_t)row * os.h + (int64_t)col * os.w + cv * V) = pack<T>(f);
    }
}

template <typename T, int ACT>
__global__ void __launch_bounds__(256) bias_act_generic(const T* x, Strides4 xs, const float* __restrict__ bias, const T* res, Strides4 rs, T* o, Strides4 os,
                                                        int C, int H, int W, int64_t total, bool ch_fast) {
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        int64_t t = idx, n; int c, row, col;
        if (ch_fast) { c = (int)(t % C); t /= C; col = (int)(t % W); t /= W; row = (int)(t % H); n = t / H; }
        else { col = (int)(t % W); t /= W; row = (int)(t % H); t /= H; c = (int)(t % C); n = t / C; }
        float v = to_f(x[n * xs.n + (int64_t)c * xs.c + (int64_t)row * xs.h + (int64_t)col * xs.w]) + (bias ? __ldg(bias + c) : 0.f);
        v = ACT == 1 ? silu_f<T>(v) : (ACT == 2 ? fmaxf(v, 0.f) : v);
        if (res) v += to_f(res[n * rs.n + (int64_t)c * rs.c + (int64_t)row * rs.h + (int64_t)col * rs.w]);
        o[n * os.n + (int64_t)c * os.c + (int64_t)row * os.h + (int64_t)col * os.w] = from_f<T>(v);
    }
}

// out[:, :C1] = nearest-2x(x), out[:, C1:] = skip; thread <-> (channel vector of out, column of out)
template <typename T>
__global__ void __launch_bounds__(256) upsample2x_cat_tiled(const T* __restrict__ x, Strides4 xs, const T* __restrict__ skip, Strides4 ss, T* __restrict__ o,
                                                            Strides4 os, int CV1, int CV, int cols_per_block, int H, int W) {
    constexpr int V = Vec16<T>::N;
    const int xi = (int)threadIdx.x / CV, cv = (int)threadIdx.x - xi * CV, col = (int)blockIdx.x * cols_per_block + xi;
    if (xi >= cols_per_block || col >= W) return;
    const int r0 = (int)blockIdx.y * kEpRows, r1 = min(r0 + kEpRows, H);
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

}  // namespace el

using namespace el;

extern "C" int el_bias_act_fwd(const void* x, const int64_t xs_[4], const float* bias, const void* residual, const int64_t rs_[4], void* out,
                               const int64_t os_[4], int B, int C, int H, int W, int act, int dtype, void* stream) {
    if (!x || !out || B <= 0 || C <= 0 || H <= 0 || W <= 0 || act < 0 || act > 2 || (residual && !rs_)) return EL_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    Strides4 xs = s4(xs_), os = s4(os_), rs = residual ? s4(rs_) : Strides4{0, 0, 0, 0};
#define EL_BIAS_ACT(KERNEL, ...)                                                           \
    do {                                                                                   \
        if (act == 1) KERNEL<T, 1> __VA_ARGS__;                                            \
        else if (act == 2) KERNEL<T, 2> __VA_ARGS__;                                       \
        else KERNEL<T, 0> __VA_ARGS__;                                                     \
    } while (0)
    EL_DISPATCH_DTYPE(dtype, {
        constexpr int V = Vec16<T>::N;
        if (channel_vectorisable<T>(x, xs, C) && channel_vectorisable<T>(out, os, C) && (!residual || channel_vectorisable<T>(residual, rs, C))) {
            const int CV = C / V;
            const int64_t total = (int64_t)B * H * W * CV;
            if (CV <= 256 && W >= 2 * (256 / CV) && B <= 65535) {
                const int cpb = 256 / CV;
                dim3 g((unsigned)ceil_div(W, cpb), (unsigned)ceil_div(H, kEpRows), (unsigned)B);
                EL_BIAS_ACT(bias_act_tiled, <<<g, 256, 0, st>>>((const T*)x, xs, bias, (const T*)residual, rs, (T*)out, os, CV, cpb, H, W));
            } else if (total < ((int64_t)1 << 32)) {
                int grid = (int)(ceil_div(total, 256) < (int64_t)kSMs * 16 ? ceil_div(total, 256) : (int64_t)kSMs * 16);
                EL_BIAS_ACT(bias_act_flat, <<<grid, 256, 0, st>>>((const T*)x, xs, bias, (const T*)residual, rs, (T*)out, os, CV, H, W, (uint32_t)total));
            } else {
                return EL_ERR_UNSUPPORTED;
            }
        } else {
            const int64_t total = (int64_t)B * C * H * W;
            int grid = (int)(ceil_div(total, 256) < (int64_t)kSMs * 16 ? ceil_div(total, 256) : (int64_t)kSMs * 16);
            EL_BIAS_ACT(bias_act_generic, <<<grid, 256, 0, st>>>((const T*)x, xs, bias, (const T*)residual, rs, (T*)out, os, C, H, W, total, os.c == 1));
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
        dim3 g((unsigned)ceil_div(W, cpb), (unsigned)ceil_div(H, kEpRows), (unsigned)B);
        upsample2x_cat_tiled<T><<<g, 256, 0, st>>>((const T*)x, xs, (const T*)skip, ss, (T*)out, os, C1 / V, CV, cpb, H, W);
    });
    note_launches(1);
    return check_launch();
}
